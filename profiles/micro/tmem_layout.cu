// Micro-test: where does an M=64 cta_group::1 accumulator live in TMEM, and what is the register fragment of
// tcgen05.ld.16x256b?  (csrc/rollout_tc.cu relies on both.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_layout tmem_layout.cu && ./tmem_layout
// D[r][n] = 100 r + n  (A[r][0] = r, A[r][1] = 1; B[n][0] = 100, B[n][1] = n; bf16-exact).
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

__global__ void __launch_bounds__(128, 1) layout_kernel(int M, float* out32, float* out16) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_s;
    __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(smem);            // [2 chunks][M rows][8]
    __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(smem + 8192);     // [2 chunks][16 rows][8]
    const int N = 16;
    for (int i = threadIdx.x; i < 2 * M * 8; i += blockDim.x) {
        const int ch = i / (M * 8), r = (i / 8) % M, j = i % 8, k = ch * 8 + j;
        A[i] = __float2bfloat16(k == 0 ? (float)r : (k == 1 ? 1.f : 0.f));
    }
    for (int i = threadIdx.x; i < 2 * N * 8; i += blockDim.x) {
        const int ch = i / (N * 8), n = (i / 8) % N, j = i % 8, k = ch * 8 + j;
        B[i] = __float2bfloat16(k == 0 ? 100.f : (k == 1 ? (float)n : 0.f));
    }
    tc::fence_proxy_async();
    if (threadIdx.x == 0) {
        tc::mbar_init(tc::smem_u32(&bar), 1);
        tc::fence_barrier_init();
    }
    if (threadIdx.x < 32) tc::tmem_alloc(tc::smem_u32(&tmem_s), 32);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tm = tmem_s;
    // poison the accumulator columns first so untouched lanes are visible
    {
        const uint32_t taddr = tm + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);
        for (int c = 0; c < 16; ++c)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr + c), "r"(__float_as_uint(-7.f)) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    if (threadIdx.x == 0) {
        const uint64_t ad = desc_plain(tc::smem_u32(A), M * 16, 128), bd = desc_plain(tc::smem_u32(B), N * 16, 128);
        tc::umma_bf16(tm, ad, bd, tc::idesc_bf16(M, N, 0, 0), 0);
        tc::umma_commit(tc::smem_u32(&bar));
    }
    tc::mbar_wait(tc::smem_u32(&bar), 0);
    tc::tc_fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    {   // 32x32b: thread = TMEM lane, 16 columns
        float v[16];
        tc::tmem_ld16(tm + ((uint32_t)(warp * 32) << 16), v);
        for (int c = 0; c < 16; ++c) out32[threadIdx.x * 16 + c] = v[c];
    }
    for (int half = 0; half < 2; ++half) {   // 16x256b.x1 at lane base 32*warp + 16*half, columns 0..7
        uint32_t r[4];
        const uint32_t taddr = tm + ((uint32_t)(warp * 32 + half * 16) << 16);
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 4; ++i) out16[((warp * 2 + half) * 32 + lane) * 4 + i] = __uint_as_float(r[i]);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        tc::tc_fence_after();
        tc::tmem_dealloc(tm, 32);
    }
}

int main() {
    float *o32, *o16;
    cudaMalloc(&o32, 128 * 16 * 4);
    cudaMalloc(&o16, 8 * 32 * 4 * 4);
    static float h32[128 * 16], h16[8 * 32 * 4];
    for (int M = 64; M <= 128; M += 64) {
        layout_kernel<<<1, 128, 16384>>>(M, o32, o16);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        cudaMemcpy(h32, o32, sizeof(h32), cudaMemcpyDeviceToHost);
        cudaMemcpy(h16, o16, sizeof(h16), cudaMemcpyDeviceToHost);
        printf("M=%d  32x32b: lane -> (row, first col) [value = 100*row + col; -7 = untouched]\n", M);
        for (int l = 0; l < 128; ++l) printf("%s%3d:%5.0f,%5.0f", (l % 8 == 0) ? "\n" : "  ", l, h32[l * 16], h32[l * 16 + 1]);
        printf("\nM=%d  16x256b.x1 (lane base 32w+16h): thread -> 4 regs\n", M);
        for (int wh = 0; wh < 8; ++wh) {
            printf(" w%d h%d:", wh / 2, wh % 2);
            for (int t = 0; t < 32; ++t) {
                if (t % 8 == 0) printf("\n   ");
                printf(" t%02d[%4.0f %4.0f %4.0f %4.0f]", t, h16[(wh * 32 + t) * 4], h16[(wh * 32 + t) * 4 + 1], h16[(wh * 32 + t) * 4 + 2], h16[(wh * 32 + t) * 4 + 3]);
            }
            printf("\n");
        }
    }
    return 0;
}
