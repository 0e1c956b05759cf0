// Developer microbenchmark: issue rate of the tcgen05.mma shapes the attention kernel uses, one CTA per SM, operands
// resident (garbage data), one thread issuing back to back - cycles per MMA as a function of N, of where A lives
// (shared memory "SS" or tensor memory "TS") and of the kind (f8f6f4 / f16).  The question it answers: is the QK^T
// GEMM of the kernel (M128 x N64 x K32 e4m3, A = Q and B = K both from shared memory: 6 KB of operand reads per MMA)
// bound by the tensor pipe (N / 2 cycles) or by shared-memory operand bandwidth?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "../../quantumattention_b200/csrc/ptx.cuh"

using namespace qa;

struct Case {
    int kind_f16;  // 0: kind::f8f6f4 (K = 32 per MMA), 1: kind::f16 (K = 16 per MMA)
    int ts;        // A from tensor memory
    int N;
    int mix;       // > 0: interleave with `mix`-style second MMA (TS, N = 128) as the kernel's PV does
};

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(Case c, int iters, long long* out) {  // (c is modified locally)
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        tmem_alloc(&tmem_base_s, 512);
        tmem_relinquish();
    }
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_init(&bar2, 1);
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    // mix >= 10: a SECOND warp issues the same stream (mix - 10) into other accumulator columns at the same time - do the
    // A-operand loads of one issuer's MMAs overlap the other's products?
    const bool two = c.mix >= 10;
    if (two) c.mix -= 10;
    if (warp == 1 || (two && warp == 2)) {
        // A: 128 rows x 128 bytes (one 128B-swizzle box, K-major), B: N rows x 128 bytes behind it; 4 K slices of 32 bytes
        const uint64_t a_desc = make_smem_desc(smem_u32(smem), 16, 8 * 128, kSwz128);
        const uint64_t b_desc = make_smem_desc(smem_u32(smem + 16384), 16, 8 * 128, kSwz128);
        // second operand set for the mixed case: V-like MN-major B of 128 columns
        const uint64_t v_desc = make_smem_desc(smem_u32(smem + 65536), 16384, 8 * 128, kSwz128);
        const uint32_t idesc = make_idesc(c.kind_f16 ? 1 : 0, c.kind_f16 ? 1 : 0, 0, 0, 128, c.N);
        const uint32_t idesc_pv = make_idesc(c.mix == 2 ? 1 : 0, c.mix == 2 ? 1 : 0, 0, 1, 128, 128);
        const uint32_t d0 = tmem + (warp == 2 ? 128 : 0), d1 = tmem + 256 + (warp == 2 ? 128 : 0), a_t = tmem + 448;
        uint64_t* mybar = warp == 2 ? &bar2 : &bar;
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t ad = a_desc + uint64_t((k * 32) >> 4), bd = b_desc + uint64_t((k * 32) >> 4);
                    if (c.kind_f16) {
                        if (c.ts) umma_f16_ts(d0, a_t + k * 8, bd, idesc, 1u);
                        else umma_f16_ss(d0, ad, bd, idesc, 1u);
                    } else {
                        if (c.ts) umma_f8_ts(d0, a_t + k * 8, bd, idesc, 1u);
                        else umma_f8_ss(d0, ad, bd, idesc, 1u);
                    }
                }
                if (c.mix == 1) {  // fp8 PV: two TS MMAs of 32 keys, N = 128
#pragma unroll
                    for (int k = 0; k < 2; ++k) umma_f8_ts(d1, a_t + 32 + k * 8, v_desc + uint64_t((k * 32 * 128) >> 4), idesc_pv, 1u);
                } else if (c.mix == 2) {  // 16-bit PV: four TS MMAs of 16 keys, N = 128 (V rows of 256 bytes: two boxes)
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16_ts(d1, a_t + 32 + k * 8, v_desc + uint64_t((k * 16 * 128) >> 4), idesc_pv, 1u);
                }
            }
            umma_commit(mybar);
        }
        __syncwarp();
        mbar_wait(mybar, 0);
        if (elect_one()) {
            t1 = clock64();
            if (blockIdx.x == 0 && warp == 1) out[0] = t1 - t0;
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2000;
    const Case cases[] = {
        {0, 0, 64, 0},  {0, 0, 128, 0}, {0, 0, 256, 0}, {0, 1, 64, 0},  {0, 1, 128, 0}, {0, 1, 256, 0},
        {1, 0, 64, 0},  {1, 0, 128, 0}, {1, 0, 256, 0}, {1, 1, 64, 0},  {1, 1, 128, 0}, {1, 1, 256, 0},
        {0, 0, 64, 1},  {0, 1, 64, 1},  {0, 0, 64, 2},  {0, 1, 64, 2},  {0, 0, 128, 1}, {0, 0, 128, 2},
        {0, 0, 64, 10}, {0, 1, 64, 10}, {0, 0, 64, 11}, {0, 0, 64, 12}, {0, 1, 64, 12},  // two issuing warps
    };
    for (const Case& c : cases) {
        for (int grid : {148}) {
            mma_rate_kernel<<<grid, 128, 200 * 1024>>>(c, iters, d);
            cudaError_t e = cudaDeviceSynchronize();
            long long cyc = 0;
            cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
            const int m_ = c.mix % 10;
            const int per_iter = 4 + (m_ == 1 ? 2 : (m_ == 2 ? 4 : 0));
            printf("%s %s N=%3d mix=%d grid=%3d: %7.1f cycles per iteration (%d MMAs), %6.1f per QK-shaped MMA group  [%s]\n",
                   c.kind_f16 ? "f16" : "f8 ", c.ts ? "TS" : "SS", c.N, c.mix, grid, double(cyc) / iters, per_iter,
                   double(cyc) / iters / 4, cudaGetErrorString(e));
        }
    }
    return 0;
}
