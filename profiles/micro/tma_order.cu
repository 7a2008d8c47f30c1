// Does cuTensorMapEncodeTiled accept NON-MONOTONIC strides, so that one box {x, y, img, chunk} of a planar tensor
// [img][chunk][H][W][8 bf16] lands in shared memory as [chunk][img][y][x] (the operand layout of csrc/conv_plane.cu) with ONE
// TMA instruction?  Checks the bytes and times it.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_order tma_order.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(32, 1) k(const __grid_constant__ CUtensorMap map, int bytes, int x0, int y0, int img0, int iters, uint4* out, long long* cyc) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar;
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (smem0 - smem_u32(smem_raw));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                         ::"r"(smem0), "l"(&map), "r"(2 * x0), "r"(y0), "r"(img0), "r"(0), "r"(smem_u32(&bar)) : "memory");
        }
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(it & 1) : "memory");
    }
    long long t1 = clock64();
    for (int i = threadIdx.x; i < bytes / 16; i += 32) out[i] = reinterpret_cast<uint4*>(sm)[i];
    if (threadIdx.x == 0) cyc[0] = (t1 - t0) / iters;
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    const int n = 64, C = 8, H = 13, W = 13;                 // planar [n][C chunks][H][W][16 B]
    const int bx = 15, by = 15, bi = 3, x0 = -1, y0 = -1, img0 = 5;     // box with a one-pixel zero border (the `up` kernels)
    std::vector<uint32_t> h((size_t)n * C * H * W * 4);
    for (int i = 0; i < n; ++i) for (int c = 0; c < C; ++c) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x)
        for (int e = 0; e < 4; ++e) h[((((size_t)i * C + c) * H + y) * W + x) * 4 + e] = (i << 24) | (c << 16) | (y << 8) | x;
    uint32_t* d;
    cudaMalloc(&d, h.size() * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap map;
    cuuint64_t dims[4] = {2 * W, H, n, C};
    cuuint64_t strides[3] = {16ull * W, 16ull * W * H * C, 16ull * W * H};        // row, IMAGE (large), chunk (small): non-monotonic
    cuuint32_t box[4] = {2 * bx, by, bi, C}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode with strides {row, image, chunk}: %s (%d)\n", r == CUDA_SUCCESS ? "accepted" : "REJECTED", (int)r);
    if (r != CUDA_SUCCESS) return 0;
    const int bytes = C * bi * by * bx * 16;
    uint4* d_out;
    long long* d_cyc;
    cudaMalloc(&d_out, bytes);
    cudaMalloc(&d_cyc, 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k<<<1, 32, bytes + 2048>>>(map, bytes, x0, y0, img0, 200, d_out, d_cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<uint32_t> o(bytes / 4);
    long long cyc;
    cudaMemcpy(o.data(), d_out, bytes, cudaMemcpyDeviceToHost);
    cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
    long long bad = 0;
    for (int c = 0; c < C; ++c) for (int i = 0; i < bi; ++i) for (int y = 0; y < by; ++y) for (int x = 0; x < bx; ++x) {
        const int gy = y0 + y, gx = x0 + x;
        const uint32_t want = (gy < 0 || gy >= H || gx < 0 || gx >= W) ? 0u : (uint32_t)(((img0 + i) << 24) | (c << 16) | (gy << 8) | gx);
        const uint32_t got = o[((((size_t)c * bi + i) * by + y) * bx + x) * 4];
        if (got != want && bad++ < 5) printf("mismatch chunk %d img %d y %d x %d: got %08x want %08x\n", c, i, y, x, got, want);
    }
    printf("one box {x %d, y %d, img %d, chunk %d} = %d bytes -> smem [chunk][img][y][x]: %lld mismatches; %lld cycles per load (issue + latency, L2-resident)\n",
           bx, by, bi, C, bytes, bad, cyc);
    return 0;
}
