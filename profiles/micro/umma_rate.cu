// Microbenchmark: tcgen05.mma issue/execute rate on sm_100a for the operand layouts used by csrc/conv_plane.cu.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu && ./umma_rate
// One CTA per SM; one thread issues `iters` MMAs (M=128, N, K=16, bf16 -> fp32) and waits for the commit.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../multimodal-rssm_b200/csrc/tc_common.cuh"

__device__ __forceinline__ uint64_t desc_plain(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// amode/bmode: 0 = un-swizzled K-major planes (LBO = plane stride, SBO = 128), 1 = 128B-swizzled K-major rows
// pattern: 0 = same accumulator & same operands; 1 = cycle K-steps (4 per 64-wide block) on one accumulator;
//          2 = switch accumulator every MMA (4 accumulators); 3 = switch every 4 MMAs
// whole warp executes; one elected lane issues (predicate produced and consumed inside one asm block)
__device__ __forceinline__ void umma_elect(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ int g_ps_a = 4224, g_lbo_b = 0;
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int amode, int bmode, int pattern, int iters, int conv, long long* out, int Mrows = 128) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_s;
    const uint32_t smem0 = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
    tc::fence_proxy_async();
    if (threadIdx.x == 0) {
        tc::mbar_init(tc::smem_u32(&bar), pattern == 6 ? 2 : 1);
        tc::fence_barrier_init();
    }
    if (threadIdx.x < 32) tc::tmem_alloc(tc::smem_u32(&tmem_s), 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tm = tmem_s;
    if (pattern == 6 && threadIdx.x == 64) {   // second issuer (warp 2), accumulator 1, tight loop
        const uint32_t idesc2 = tc::idesc_bf16(128, N, 0, 0);
        uint64_t ad = desc_plain(smem0, 4224, 128), bd = tc::smem_desc_sw128(smem0 + 96 * 1024, 16, 1024);
        for (int i = 0; i < iters; ++i) tc::umma_bf16(tm + 256, ad + (uint64_t)(2 * (i & 3)), bd + (uint64_t)(2 * (i & 3)), idesc2, 1);
        tc::umma_commit(tc::smem_u32(&bar));
    }
    if (threadIdx.x < 32 && (conv || threadIdx.x == 0)) {
        const bool leader = conv ? elect_one() : true;
        const uint32_t idesc = tc::idesc_bf16(Mrows, N, 0, 0);
        const uint32_t sA = smem0, sB = smem0 + 96 * 1024;
        const uint32_t PS = 4224;      // plane stride of a typical tile (odd multiple of 128)
        long long t0 = clock64();
        if (pattern == 6) {
            uint64_t ad = desc_plain(sA, PS, 128), bd = tc::smem_desc_sw128(sB, 16, 1024);
            for (int i = 0; i < iters; ++i) tc::umma_bf16(tm, ad + (uint64_t)(2 * (i & 3)), bd + (uint64_t)(2 * (i & 3)), idesc, 1);
        } else if (pattern == 5) {            // tight, converged warp, elect inside the asm
            uint64_t ad = tc::smem_desc_sw128(sA, 16, 1024), bd = tc::smem_desc_sw128(sB, 16, 1024);
            if (amode == 0) ad = desc_plain(sA, PS, 128);
            for (int i = 0; i < iters; i += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) umma_elect(tm, ad + (uint64_t)(2 * (u & 3)), bd + (uint64_t)(2 * (u & 3)), idesc, 1);
            }
        } else if (pattern == 4) {            // tight: descriptors precomputed, unrolled by 8, only adds between MMAs
            uint64_t ad = tc::smem_desc_sw128(sA, 16, 1024), bd = tc::smem_desc_sw128(sB, 16, 1024);
            if (amode == 0) ad = desc_plain(sA, g_ps_a, 128);
            if (bmode == 0) bd = desc_plain(sB, g_lbo_b ? g_lbo_b : N * 16, 128);
            for (int i = 0; i < iters; i += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (leader) tc::umma_bf16(tm, ad + (uint64_t)(2 * (u & 3)), bd + (uint64_t)(2 * (u & 3)), idesc, 1);
                }
            }
        } else
        for (int i = 0; i < iters; ++i) {
            const int j = (pattern == 1) ? (i & 3) : 0;
            const int acc = (pattern == 2) ? (i & 3) : (pattern == 3 ? ((i >> 2) & 3) : 0);
            uint64_t ad = amode == 0 ? desc_plain(sA + j * 2 * PS, PS, 128) : tc::smem_desc_sw128(sA + j * 32, 16, 1024);
            uint64_t bd = bmode == 0 ? desc_plain(sB + j * 2 * PS, PS, 128) : tc::smem_desc_sw128(sB + j * 32, 16, 1024);
            if (leader) tc::umma_bf16(tm + acc * 128 * (N <= 128), ad, bd, idesc, i != 0);
        }
        long long t1 = clock64();
        if (leader) tc::umma_commit(tc::smem_u32(&bar));
        tc::mbar_wait(tc::smem_u32(&bar), 0);
        long long t2 = clock64();
        if (blockIdx.x == 0 && leader) {
            out[0] = t1 - t0;
            out[1] = t2 - t0;
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tm, 512);
    }
}

int main(int argc, char** argv) {
    long long* d;
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2000;
    if (argc > 1 && argv[1][0] == 't') {   // tight loop (pattern 4), un-swizzled A with LBO = psa, B un-swizzled (LBO = lbob or N*16) or swizzled
        printf("M   N  psA  bmode lboB  issue_cyc/mma  total_cyc/mma\n");
        for (int M : {64, 128})
            for (int N : {64, 112, 208})
                for (int psa : {1024, 1152, 4224})
                    for (int bm = 0; bm < 3; ++bm) {
                        int lbob = bm == 1 ? (N * 16 + 128) : 0;      // bm 0: LBO = N*16; bm 1: N*16 + 128 (odd/even flip); bm 2: swizzled
                        cudaMemcpyToSymbol(g_ps_a, &psa, 4);
                        cudaMemcpyToSymbol(g_lbo_b, &lbob, 4);
                        rate_kernel<<<148, 128, 200 * 1024>>>(N, 0, bm == 2 ? 1 : 0, 4, iters, 0, d, M);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                        long long h[2];
                        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                        printf("%3d %3d %5d %5d %5d  %10.1f  %10.1f\n", M, N, psa, bm, lbob ? lbob : N * 16, (double)h[0] / iters, (double)h[1] / iters);
                    }
        return 0;
    }
    if (argc > 1) {   // sweep for csrc/rollout_tc.cu: M x N x operand layouts, one accumulator, same operands (pattern 0)
        printf("M   N  amode bmode  issue_cyc/mma  total_cyc/mma   (mode 0 = un-swizzled K-major planes, 1 = 128B swizzle)\n");
        for (int M : {64, 128})
            for (int N : {64, 96, 112, 208, 256})
                for (int amode = 0; amode < 2; ++amode)
                    for (int bmode = 0; bmode < 2; ++bmode) {
                        rate_kernel<<<148, 128, 200 * 1024>>>(N, amode, bmode, 0, iters, 0, d, M);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                        long long h[2];
                        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                        printf("%3d %3d %5d %5d  %10.1f  %10.1f\n", M, N, amode, bmode, (double)h[0] / iters, (double)h[1] / iters);
                    }
        return 0;
    }
    printf("N amode bmode pattern  issue_cyc/mma  total_cyc/mma   (M=128,K=16; floor N/2)\n");
    printf("(last column block: conv = 1 -> whole warp runs the loop, MMA under elect.sync)\n");
    for (int conv = 0; conv < 2; ++conv)
    for (int N : {32, 128, 256})
        for (int amode = 0; amode < 1; ++amode)
            for (int bmode = 1; bmode < 2; ++bmode)
                for (int pattern = 4; pattern < 7; ++pattern) {
                    if (pattern == 5 && !conv) continue;
                    if (pattern == 6 && conv) continue;
                    rate_kernel<<<148, 128, 200 * 1024>>>(N, amode, bmode, pattern, iters, conv, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    long long h[2];
                    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                    printf("conv=%d %3d %5d %5d %7d  %10.1f  %10.1f\n", conv, N, amode, bmode, pattern, (double)h[0] / iters, (double)h[1] / iters);
                }
    printf("M=64 tight loop:\n");
    for (int N : {16, 32, 64, 128}) {
        rate_kernel<<<148, 128, 200 * 1024>>>(N, 0, 1, 4, iters, 0, d, 64);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[2];
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("M=64 N=%3d  %10.1f  %10.1f\n", N, (double)h[0] / iters, (double)h[1] / iters);
    }
    for (int N : {16}) {
        rate_kernel<<<148, 128, 200 * 1024>>>(N, 0, 1, 4, iters, 0, d, 128);
        cudaDeviceSynchronize();
        long long h[2];
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("M=128 N=%3d  %10.1f  %10.1f\n", N, (double)h[0] / iters, (double)h[1] / iters);
    }
    return 0;
}
