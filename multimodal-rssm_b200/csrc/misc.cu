// Reconstruction loss, fused clip+Adam, and small layout helpers.
#include <algorithm>
#include "common.cuh"

namespace {

constexpr int RED_BLOCKS = 1184;   // 8 x 148 SMs
constexpr int RED_THREADS = 256;

__device__ __forceinline__ double block_sum_d(double v) {
    __shared__ double sh[32];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
    return t;   // valid on thread 0
}

// partial[b] = sum over the block's grid-stride slice of (y-o)^2  (or of g^2 when o == nullptr)
__global__ void sq_partial_kernel(const float* __restrict__ y, const float* __restrict__ o, long long n, float* partial) {
    float acc = 0.f;
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    long long stride = (long long)gridDim.x * blockDim.x * 4;
    bool vec = ((reinterpret_cast<uintptr_t>(y) | (o ? reinterpret_cast<uintptr_t>(o) : 0)) & 15) == 0;
    for (; i < n; i += stride) {
        if (vec && i + 3 < n) {
            float4 a = *reinterpret_cast<const float4*>(y + i);
            float4 b = o ? *reinterpret_cast<const float4*>(o + i) : make_float4(0, 0, 0, 0);
            float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
            acc += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
        } else {
            for (long long j = i; j < min(n, i + 4); ++j) {
                float d = y[j] - (o ? o[j] : 0.f);
                acc += d * d;
            }
        }
    }
    double t = block_sum_d((double)acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = (float)t;
}

// mode 0: out[0] = sum * scale ; mode 1: out[0] = sqrt(sum) * scale
__global__ void final_reduce_kernel(const float* partial, int nb, double scale, int mode, float* out) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) acc += partial[i];
    double t = block_sum_d(acc);
    if (threadIdx.x == 0) out[0] = (float)((mode ? sqrt(t) : t) * scale);
}

__global__ void mse_bwd_kernel(const float* __restrict__ y, const float* __restrict__ o, long long n, float inv_rows,
                               const float* g, float* dy) {
    float s = 2.f * inv_rows * (g ? g[0] : 1.f);
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dy[i] = s * (y[i] - o[i]);
}

__global__ void sqdiff_kernel(const float* __restrict__ y, const float* __restrict__ o, long long n, float* out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float d = y[i] - o[i];
        out[i] = d * d;
    }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr_over_bc1, float inv_sqrt_bc2, float b1, float b2, float eps,
                            float max_norm, float grad_scale, const float* norm) {
    // torch.nn.utils.clip_grad_norm_: coef = min(1, max_norm / (total_norm + 1e-6))
    float coef = fminf(1.f, max_norm / (norm[0] + 1e-6f)) * grad_scale;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float gi = g[i] * coef;
        float mi = b1 * m[i] + (1.f - b1) * gi;
        float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
        p[i] -= lr_over_bc1 * (mi / denom);
    }
}

__global__ void transpose_kernel(const float* __restrict__ src, long long rows, long long cols, long long ld, float* __restrict__ dst) {
    __shared__ float tile[32][33];
    long long c = (long long)blockIdx.x * 32 + threadIdx.x;
    for (int i = threadIdx.y; i < 32; i += 8) {
        long long r = (long long)blockIdx.y * 32 + i;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? src[r * ld + c] : 0.f;
    }
    __syncthreads();
    long long r2 = (long long)blockIdx.y * 32 + threadIdx.x;
    for (int i = threadIdx.y; i < 32; i += 8) {
        long long c2 = (long long)blockIdx.x * 32 + i;
        if (c2 < cols && r2 < rows) dst[c2 * rows + r2] = tile[threadIdx.x][i];
    }
}

__global__ void concat2_kernel(const float* __restrict__ a, long long ca, const float* __restrict__ b, long long cb,
                               long long rows, float* __restrict__ out) {
    long long n = rows * (ca + cb);
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        long long r = i / (ca + cb), c = i % (ca + cb);
        out[i] = c < ca ? a[r * ca + c] : b[r * cb + (c - ca)];
    }
}

__global__ void colsum_acc_kernel(const float* __restrict__ x, long long rows, long long cols, long long ld, float* out, long long chunk) {
    long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    long long r0 = (long long)blockIdx.y * chunk, r1 = min(rows, r0 + chunk);
    float acc = 0.f;
    for (long long r = r0; r < r1; ++r) acc += x[r * ld + c];
    atomicAdd(out + c, acc);
}

__global__ void act_bwd_kernel(const float* __restrict__ g, const float* __restrict__ y, long long n, int act, float* out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = g[i] * act_grad_from_out(y[i], act);
}

__global__ void fill_kernel(float* p, long long n, float v) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

inline int blocks_for(long long n, int threads = 256, int cap = 148 * 16) {
    return (int)std::max<long long>(1, std::min<long long>(cap, ceil_div64(n, threads)));
}

}  // namespace

extern "C" int mrssm_mse_fwd(const float* y, const float* o, int64_t n, int64_t rows, float* partial, float* out, void* stream) {
    MRSSM_CHECK(y && o && partial && out && n > 0 && rows > 0, "mse_fwd: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    sq_partial_kernel<<<RED_BLOCKS, RED_THREADS, 0, st>>>(y, o, n, partial);
    MRSSM_LAUNCH_CHECK();
    final_reduce_kernel<<<1, 1024, 0, st>>>(partial, RED_BLOCKS, 1.0 / (double)rows, 0, out);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

extern "C" int mrssm_mse_bwd(const float* y, const float* o, int64_t n, int64_t rows, const float* g, float* dy, void* stream) {
    MRSSM_CHECK(y && o && dy && n > 0 && rows > 0, "mse_bwd: bad args");
    mse_bwd_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(y, o, n, 1.f / (float)rows, g, dy);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

extern "C" int mrssm_sqdiff(const float* y, const float* o, int64_t n, float* out, void* stream) {
    MRSSM_CHECK(y && o && out && n > 0, "sqdiff: bad args");
    sqdiff_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(y, o, n, out);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

extern "C" int mrssm_clip_adam(float* p, const float* g, float* m, float* v, int64_t n, int32_t step, float lr,
                               float beta1, float beta2, float eps, float max_norm, float grad_scale,
                               float* partial, float* norm_out, void* stream) {
    MRSSM_CHECK(p && g && m && v && partial && norm_out && n > 0 && step > 0, "clip_adam: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    sq_partial_kernel<<<RED_BLOCKS, RED_THREADS, 0, st>>>(g, nullptr, n, partial);
    MRSSM_LAUNCH_CHECK();
    final_reduce_kernel<<<1, 1024, 0, st>>>(partial, RED_BLOCKS, (double)grad_scale, 1, norm_out);
    MRSSM_LAUNCH_CHECK();
    double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    adam_kernel<<<blocks_for(n), 256, 0, st>>>(p, g, m, v, n, (float)(lr / bc1), (float)(1.0 / sqrt(bc2)), beta1, beta2,
                                               eps, max_norm, grad_scale, norm_out);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

extern "C" int mrssm_transpose(const float* src, int64_t rows, int64_t cols, int64_t src_ld, float* dst, void* stream) {
    MRSSM_CHECK(src && dst && rows > 0 && cols > 0 && src_ld >= cols, "transpose: bad args");
    dim3 grid((unsigned)ceil_div64(cols, 32), (unsigned)ceil_div64(rows, 32)), block(32, 8);
    transpose_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, rows, cols, src_ld, dst);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

extern "C" int mrssm_concat2(const float* a, int64_t ca, const float* b, int64_t cb, int64_t rows, float* out, void* stream) {
    MRSSM_CHECK(a && b && out && ca > 0 && cb > 0 && rows > 0, "concat2: bad args");
    concat2_kernel<<<blocks_for(rows * (ca + cb)), 256, 0, (cudaStream_t)stream>>>(a, ca, b, cb, rows, out);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

extern "C" int mrssm_colsum_acc(const float* x, int64_t rows, int64_t cols, int64_t ld, float* out, void* stream) {
    MRSSM_CHECK(x && out && rows > 0 && cols > 0 && ld >= cols, "colsum: bad args");
    long long chunk = std::max<long long>(64, ceil_div64(rows, 592));
    dim3 grid((unsigned)ceil_div64(cols, 128), (unsigned)ceil_div64(rows, chunk));
    colsum_acc_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(x, rows, cols, ld, out, chunk);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

extern "C" int mrssm_fill(float* p, int64_t n, float v, void* stream) {
    MRSSM_CHECK(p && n > 0, "fill: bad args");
    fill_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(p, n, v);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

extern "C" int mrssm_act_bwd(const float* g, const float* y, int64_t n, int32_t act, float* out, void* stream) {
    MRSSM_CHECK(g && y && out && n > 0, "act_bwd: bad args");
    act_bwd_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(g, y, n, act, out);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

// ---- input pipeline: uint8 frames -> normalised fp32 (utils/processing/image_processing.py:5-11) ------------------
// dst = floor(u8 / 2^(8-bits)) / 2^bits - 0.5 + u / 2^bits,  u ~ U[0,1): from `noise` when given (parity tests),
// otherwise a counter-based hash of (seed, element index) (the reference draws torch.rand_like on the device).
namespace {
__device__ __forceinline__ float hash_uniform(uint64_t seed, uint64_t i) {
    uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ull;           // splitmix64
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (float)(uint32_t)(z >> 40) * (1.0f / 16777216.0f);      // 24 random bits -> [0,1)
}
__global__ void normalize_u8_kernel(const uint8_t* __restrict__ src, long long n, int bits, const float* __restrict__ noise,
                                    uint64_t seed, float* __restrict__ dst) {
    const float q = 1.f / (float)(1 << (8 - bits)), s = 1.f / (float)(1 << bits);
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const uchar4 u = reinterpret_cast<const uchar4*>(src)[i];
        float4 r;
        const float4 nz = noise ? reinterpret_cast<const float4*>(noise)[i]
                                : make_float4(hash_uniform(seed, 4 * i), hash_uniform(seed, 4 * i + 1), hash_uniform(seed, 4 * i + 2),
                                              hash_uniform(seed, 4 * i + 3));
        r.x = floorf(u.x * q) * s - 0.5f + nz.x * s;
        r.y = floorf(u.y * q) * s - 0.5f + nz.y * s;
        r.z = floorf(u.z * q) * s - 0.5f + nz.z * s;
        r.w = floorf(u.w * q) * s - 0.5f + nz.w * s;
        reinterpret_cast<float4*>(dst)[i] = r;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const long long i = (n4 << 2) + threadIdx.x;
        const float nz = noise ? noise[i] : hash_uniform(seed, i);
        dst[i] = floorf(src[i] * q) * s - 0.5f + nz * s;
    }
}
}  // namespace

extern "C" int mrssm_normalize_image_u8(const uint8_t* src, int64_t n, int32_t bit_depth, const float* noise, uint64_t seed,
                                        float* dst, void* stream) {
    MRSSM_CHECK(src && dst && n > 0 && bit_depth >= 1 && bit_depth <= 8, "normalize_image_u8: bad args");
    MRSSM_CHECK(((uintptr_t)src & 3) == 0 && ((uintptr_t)dst & 15) == 0 && (!noise || ((uintptr_t)noise & 15) == 0),
                "normalize_image_u8: buffers must be 16-byte aligned");
    const int blocks = (int)std::min<long long>(148 * 16, std::max<long long>(1, ceil_div64(n / 4, 256)));
    normalize_u8_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, n, bit_depth, noise, seed, dst);
    MRSSM_LAUNCH_CHECK();
    return 0;
}


namespace {
__global__ void copy2d_kernel(const float* __restrict__ src, long long rows, long long cols, long long ld, float* __restrict__ dst) {
    const long long total = rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cols, c = i - r * cols;
        dst[i] = src[r * ld + c];
    }
}
}  // namespace

extern "C" int mrssm_copy2d(const float* src, int64_t rows, int64_t cols, int64_t ld, float* dst, void* stream) {
    MRSSM_CHECK(src && dst && rows > 0 && cols > 0 && ld >= cols, "copy2d: bad args");
    copy2d_kernel<<<blocks_for(rows * cols), 256, 0, (cudaStream_t)stream>>>(src, rows, cols, ld, dst);
    MRSSM_LAUNCH_CHECK();
    return 0;
}
