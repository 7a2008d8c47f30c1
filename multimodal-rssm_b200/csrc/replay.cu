// Replay-buffer sampling on the device (reference utils/replay_buffer/memory.py:191-208, data_augment.py:178-210,
// utils/processing/image_processing.py:5-11).  The reference keeps the frame store on the host, gathers a [L*n] index list
// there, converts to fp32, copies, and then runs crop / colour shift / Gaussian noise / clip / quantise / dequantise as separate
// full-size elementwise passes.  Here the uint8 frame store lives in HBM (12 KB per 3x64x64 frame) and ONE kernel does
//   out[row] = normalise(clip(crop(frames[idx[row]]) + delta_c + gauss * scale * 255, 0, 255))
// reading each source byte once and writing each fp32 output once (1 B in, 4 B out per element: HBM-bound).
// Row order is the reference's: row = l * n + b (time-major).
#include "common.cuh"

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t seed, uint64_t i) {
    uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ull;           // splitmix64
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float hash_normal(uint64_t seed, uint64_t i) {        // Box-Muller on two 24-bit uniforms
    const uint64_t z = mix64(seed ^ 0xD1B54A32D192ED03ull, i);
    const float u1 = ((float)(uint32_t)(z >> 40) + 1.0f) * (1.0f / 16777216.0f);  // (0,1]
    const float u2 = (float)(uint32_t)((z >> 16) & 0xFFFFFFu) * (1.0f / 16777216.0f);
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

// Two uniforms from one 64-bit hash (bits 40..63 and 16..39).
__device__ __forceinline__ void hash_uniform2(uint64_t seed, uint64_t pair, float& u0, float& u1) {
    const uint64_t z = mix64(seed, pair);
    u0 = (float)(uint32_t)(z >> 40) * (1.0f / 16777216.0f);
    u1 = (float)(uint32_t)((z >> 16) & 0xFFFFFFu) * (1.0f / 16777216.0f);
}

// One CTA per gathered frame (its slot number is read once), one thread per V consecutive output pixels of an image row
// (V = 16: one 16-byte load, four 16-byte stores; V = 4 for windows that are not 16-byte friendly).  All index math is
// 32-bit and per frame; only the frame base uses 64 bits.
template <int V>
__global__ void __launch_bounds__(256) replay_gather_u8_kernel(mrssm_replay_gather_args a, bool aligned) {
    const bool raw = a.bit_depth == 0;                  // binary masks: gathered and cropped only (memory.py:199-201)
    const float q = 1.f / (float)(1 << (8 - a.bit_depth)), s = 1.f / (float)(1 << a.bit_depth);
    const bool augment = a.delta || a.gauss || a.gauss_scale > 0.f;
    const long long row = blockIdx.x;
    const long long frame = a.idx[row];
    const uint8_t* fbase = a.frames + frame * a.C * (long long)a.Hs * a.Ws;
    const unsigned wv = a.W / V, per_frame = (unsigned)a.C * a.H * wv;
    const long long obase = row * (long long)a.C * a.H * a.W;
    for (unsigned i = threadIdx.x; i < per_frame; i += blockDim.x) {
        const unsigned xv = i % wv, t = i / wv, y = t % a.H, c = t / a.H;
        const uint8_t* src = fbase + ((size_t)c * a.Hs + (y + a.dh)) * a.Ws + a.dw + V * xv;
        float v[V];
        if (aligned) {
            if (V == 16) {
                const uint4 w = *reinterpret_cast<const uint4*>(src);
                const unsigned ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int j = 0; j < V; ++j) v[j] = (float)((ww[j >> 2] >> (8 * (j & 3))) & 0xFFu);
            } else {
                const unsigned w = *reinterpret_cast<const unsigned*>(src);
#pragma unroll
                for (int j = 0; j < V; ++j) v[j] = (float)((w >> (8 * j)) & 0xFFu);
            }
        } else {
#pragma unroll
            for (int j = 0; j < V; ++j) v[j] = src[j];
        }
        const long long o = obase + (long long)V * i;           // flat index into out [rows, C, H, W]
        if (augment) {
            const float d = a.delta ? a.delta[c] : 0.f;
#pragma unroll
            for (int j4 = 0; j4 < V; j4 += 4) {
                float g[4] = {0.f, 0.f, 0.f, 0.f};
                if (a.gauss) {
                    const float4 gg = *reinterpret_cast<const float4*>(a.gauss + o + j4);
                    g[0] = gg.x; g[1] = gg.y; g[2] = gg.z; g[3] = gg.w;
                } else if (a.gauss_scale > 0.f) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) g[j] = hash_normal(a.seed, o + j4 + j);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    // the reference's rounding sequence, kept un-fused so the quantiser sees the same value:
                    const float n = __fmul_rn(__fmul_rn(g[j], a.gauss_scale), 255.0f);                      // data_augment.py:88-91
                    v[j4 + j] = fminf(fmaxf(__fadd_rn(__fadd_rn(v[j4 + j], d), n), 0.f), 255.f);            // :207, clipped
                }
            }
        }
#pragma unroll
        for (int j4 = 0; j4 < V; j4 += 4) {
            float4 r;
            if (raw) {
                r = make_float4(v[j4], v[j4 + 1], v[j4 + 2], v[j4 + 3]);
            } else {
                float u[4];
                if (a.uniform) {
                    const float4 uu = *reinterpret_cast<const float4*>(a.uniform + o + j4);
                    u[0] = uu.x; u[1] = uu.y; u[2] = uu.z; u[3] = uu.w;
                } else {
                    hash_uniform2(a.seed, (uint64_t)(o + j4) >> 1, u[0], u[1]);
                    hash_uniform2(a.seed, ((uint64_t)(o + j4) >> 1) + 1, u[2], u[3]);
                }
                r.x = floorf(v[j4] * q) * s - 0.5f + u[0] * s;
                r.y = floorf(v[j4 + 1] * q) * s - 0.5f + u[1] * s;
                r.z = floorf(v[j4 + 2] * q) * s - 0.5f + u[2] * s;
                r.w = floorf(v[j4 + 3] * q) * s - 0.5f + u[3] * s;
            }
            *reinterpret_cast<float4*>(a.out + o + j4) = r;
        }
    }
}

// fp32 rows (vector observations, actions, rewards, nonterminals): out[r, :] = src[idx[r], :]
__global__ void gather_rows_kernel(const float* __restrict__ src, const long long* __restrict__ idx, long long rows, int K,
                                   float* __restrict__ out) {
    const long long total = rows * K;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / K;
        out[i] = src[idx[r] * K + (i - r * K)];
    }
}

int grid_for(long long n) { return (int)std::min<long long>(148 * 16, std::max<long long>(1, ceil_div64(n, 256))); }

}  // namespace

extern "C" int mrssm_replay_gather_u8(const mrssm_replay_gather_args* a, void* stream) {
    MRSSM_CHECK(a && a->frames && a->idx && a->out && a->rows > 0, "replay_gather_u8: null tensor");
    MRSSM_CHECK(a->C > 0 && a->H > 0 && a->W > 0 && a->W % 4 == 0, "replay_gather_u8: the output width must be a multiple of 4");
    MRSSM_CHECK(a->dh >= 0 && a->dw >= 0 && a->dh + a->H <= a->Hs && a->dw + a->W <= a->Ws, "replay_gather_u8: crop outside the stored frame");
    MRSSM_CHECK(a->bit_depth >= 0 && a->bit_depth <= 8, "replay_gather_u8: bit depth (0 = no normalisation)");
    MRSSM_CHECK(((uintptr_t)a->out & 15) == 0 && (!a->gauss || ((uintptr_t)a->gauss & 15) == 0) &&
                    (!a->uniform || ((uintptr_t)a->uniform & 15) == 0), "replay_gather_u8: fp32 buffers must be 16-byte aligned");
    MRSSM_CHECK(a->rows < (1ll << 31), "replay_gather_u8: too many rows for one launch");
    cudaStream_t st = (cudaStream_t)stream;
    const bool wide = a->W % 16 == 0;
    const bool al16 = ((uintptr_t)a->frames & 15) == 0 && a->Ws % 16 == 0 && a->dw % 16 == 0 &&
                      ((long long)a->Hs * a->Ws) % 16 == 0;
    const bool al4 = ((uintptr_t)a->frames & 3) == 0 && a->Ws % 4 == 0 && a->dw % 4 == 0;
    if (wide && (al16 || !al4))      // 16 pixels per thread: one 16-byte load (or 16 byte loads when nothing is aligned)
        replay_gather_u8_kernel<16><<<(unsigned)a->rows, 256, 0, st>>>(*a, al16);
    else
        replay_gather_u8_kernel<4><<<(unsigned)a->rows, 256, 0, st>>>(*a, al4);
    MRSSM_LAUNCH_CHECK();
    return 0;
}

extern "C" int mrssm_gather_rows(const float* src, const int64_t* idx, int64_t rows, int32_t K, float* out, void* stream) {
    MRSSM_CHECK(src && idx && out && rows > 0 && K > 0, "gather_rows: bad args");
    gather_rows_kernel<<<grid_for(rows * K), 256, 0, (cudaStream_t)stream>>>(src, (const long long*)idx, rows, K, out);
    MRSSM_LAUNCH_CHECK();
    return 0;
}
