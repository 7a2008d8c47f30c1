// Microbenchmark: how fast does one SM's TMA unit move small activation planes into shared memory?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_rate tma_rate.cu -lcuda && ./tma_rate
// One CTA per SM, one converged warp issues `iters` rounds of `nplanes` loads into a double-buffered stage and waits for each
// round on an mbarrier.  Modes: 0 = tensor-map 4-D box {2*bx (u64), by, 1, bi} from a planar bf16 tensor [img][chunk][H][W][8]
// (what csrc/conv_plane.cu does), 1 = the same bytes as ONE linear cp.async.bulk per (plane, image) (needs bx == W: the box is
// a contiguous run), 2 = tensor-map box with an inner extent of 8 bf16 (16 bytes: NHWC element form).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(160, 1) rate_kernel(int NW, const __grid_constant__ CUtensorMap map, const char* base, int mode, int nplanes, int bi, int bx, int by,
                                                      int W, int H, int nchunk, int n_img, int iters, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar[2];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(NW));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int wid = threadIdx.x >> 5;
    if (wid >= NW) return;
    const uint32_t lead = elect_one();
    const uint32_t plane_bytes = (uint32_t)bi * by * bx * 16u;
    const uint32_t plane_stride = (plane_bytes + 127u) & ~127u;      // TMA destinations are 128-byte aligned
    const uint32_t stage_bytes = (plane_stride * nplanes + 1023u) & ~1023u;
    const long long img_bytes = (long long)nchunk * H * W * 16;
    long long t0 = clock64();
    for (int it = 0; it < iters + 1; ++it) {
        if (it < iters) {
            const uint32_t b = smem_u32(&bar[it & 1]);
            const uint32_t dst = smem0 + (uint32_t)(it & 1) * stage_bytes;
            const int img0 = (int)(((unsigned)(blockIdx.x + it * gridDim.x) * 67u) % (unsigned)(n_img - 64));
            asm volatile("{\n\t.reg .pred e;\n\tsetp.ne.b32 e, %2, 0;\n\t@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(b), "r"(plane_bytes * nplanes / NW), "r"(lead) : "memory");
            // no divisions / selects in the issue loop: one warp runs ~5 cycles per dependent instruction
            uint32_t d = dst;
            const int groups = nplanes / nchunk;                    // images (modes 0-2 with bi: box covers bi images)
            for (int g = 0; g < groups; ++g) {
                const int img = img0 + g * bi;
                if (mode == 3) {
                    asm volatile("{\n\t.reg .pred e;\n\tsetp.ne.b32 e, %7, 0;\n\t@e cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n\t}"
                                 ::"r"(d), "l"(&map), "r"(0), "r"(0), "r"(0), "r"(img), "r"(b), "r"(lead) : "memory");
                    d += plane_stride * nchunk;
                    continue;
                }
                for (int ch = 0; ch < nchunk; ++ch, d += plane_stride) {
                    if ((ch & (NW - 1)) != wid) continue;
                    if (mode == 0) {
                        asm volatile("{\n\t.reg .pred e;\n\tsetp.ne.b32 e, %7, 0;\n\t@e cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n\t}"
                                     ::"r"(d), "l"(&map), "r"(0), "r"(0), "r"(ch), "r"(img), "r"(b), "r"(lead) : "memory");
                    } else if (mode == 2) {
                        asm volatile("{\n\t.reg .pred e;\n\tsetp.ne.b32 e, %7, 0;\n\t@e cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n\t}"
                                     ::"r"(d), "l"(&map), "r"(ch * 8), "r"(0), "r"(0), "r"(img), "r"(b), "r"(lead) : "memory");
                    } else {
                        const char* src = base + (long long)img * img_bytes + (long long)ch * H * W * 16;
                        for (int i = 0; i < bi; ++i, src += img_bytes)
                            asm volatile("{\n\t.reg .pred e;\n\tsetp.ne.b32 e, %4, 0;\n\t@e cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}"
                                         ::"r"(d + i * by * bx * 16), "l"(src), "r"((uint32_t)by * bx * 16u), "r"(b), "r"(lead) : "memory");
                    }
                }
            }
        }
        if (it > 0) mbar_wait(smem_u32(&bar[(it - 1) & 1]), ((it - 1) >> 1) & 1);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;     // warp 0's view
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    long long* d_out;
    cudaMalloc(&d_out, 64);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct Case { const char* name; int W, H, nchunk, bx, by, bi, nplanes; };
    // (W, H) of the plane, chunks per image, box, images per box, planes (box loads) per round
    Case cases[] = {{"16x16 plane box 16x16, 16 planes (4 img)", 16, 16, 4, 16, 16, 1, 16}, {"16x16 plane box 16x8, 16 planes", 16, 16, 4, 16, 8, 1, 16},
                    {"16x16 plane box 16x4, 16 planes", 16, 16, 4, 16, 4, 1, 16},           {"16x16 plane box 16x16, 4 planes", 16, 16, 4, 16, 16, 1, 4},
                    {"D2wgrad 7x7, 48 planes, bi 2", 7, 7, 8, 7, 7, 2, 48},                {"8x8 plane, 8 chunks, 32 planes", 8, 8, 8, 8, 8, 1, 32},
                    {"wide 64x4 plane, 8 planes", 64, 4, 8, 64, 4, 1, 8},                  {"13x13 plane (up), 8 planes", 13, 13, 8, 13, 13, 1, 8}};
    const int NW = atoi(getenv("NW") ? getenv("NW") : "1");
    const int n_img = atoi(getenv("NIMG") ? getenv("NIMG") : "20000"), iters = 2000;
    printf("%-36s %6s %12s %12s %10s\n", "case", "mode", "cycles/round", "bytes/round", "B/clk/SM");
    for (const Case& c : cases) {
        const size_t bytes = (size_t)n_img * c.nchunk * c.H * c.W * 16;
        char* d;
        cudaMalloc(&d, bytes);
        cudaMemset(d, 0, bytes);
        for (int mode = 0; mode < 4; ++mode) {
            if (NW > 1 && mode != 0) continue;
            CUtensorMap map;
            if (mode == 2) {
                // NHWC element form of the same bytes: [img][H][W][C = 8 * nchunk], box {8, bx, by, bi}
                cuuint64_t dims[4] = {(cuuint64_t)(8 * c.nchunk), (cuuint64_t)c.W, (cuuint64_t)c.H, (cuuint64_t)n_img};
                cuuint64_t strides[3] = {(cuuint64_t)(16 * c.nchunk), (cuuint64_t)(16 * c.nchunk * c.W), (cuuint64_t)(16 * c.nchunk * c.W * c.H)};
                cuuint32_t box[4] = {8, (cuuint32_t)c.bx, (cuuint32_t)c.by, (cuuint32_t)c.bi}, es[4] = {1, 1, 1, 1};
                if (c.bx > 256 || enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("%-36s %6d  encode failed\n", c.name, mode); continue; }
            } else {
                cuuint64_t dims[4] = {(cuuint64_t)(2 * c.W), (cuuint64_t)c.H, (cuuint64_t)c.nchunk, (cuuint64_t)n_img};
                cuuint64_t strides[3] = {(cuuint64_t)(16 * c.W), (cuuint64_t)(16 * c.W * c.H), (cuuint64_t)(16 * c.W * c.H * c.nchunk)};
                cuuint32_t box[4] = {(cuuint32_t)(2 * c.bx), (cuuint32_t)c.by, (cuuint32_t)(mode == 3 ? c.nchunk : 1), (cuuint32_t)(mode == 3 ? 1 : c.bi)}, es[4] = {1, 1, 1, 1};
                if (mode == 3 && (c.bi != 1 || (c.by * c.bx * 16) % 128 != 0)) continue;
                if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("%-36s %6d  encode failed\n", c.name, mode); continue; }
            }
            if (mode == 1 && (c.bx != c.W)) continue;
            rate_kernel<<<148, 160, 200 * 1024>>>(NW, map, d, mode, c.nplanes, c.bi, c.bx, c.by, c.W, c.H, c.nchunk, n_img, iters, d_out);
            long long h = 0;
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%-36s %6d  failed: %s\n", c.name, mode, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
            const double rb = (double)c.nplanes * c.bi * c.by * c.bx * 16;
            printf("%-36s %6d %12.0f %12.0f %10.1f\n", c.name, mode, (double)h / iters, rb, rb * iters / (double)h);
        }
        cudaFree(d);
    }
    return 0;
}
