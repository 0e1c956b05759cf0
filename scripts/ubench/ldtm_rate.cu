// Developer microbenchmark: tensor-memory read bandwidth seen by tcgen05.ld.32x32b (the softmax's S read: one row per
// thread).  W warps loop over loads into two alternating register sets (no write-after-write dependency between
// consecutive loads) and fold the values into an accumulator; prints cycles per load and bytes / clk / SM.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../quantumattention_b200/csrc/ptx.cuh"
using namespace qa;
template <int X>
__device__ __forceinline__ void ld(uint32_t a, float* r) {
    if constexpr (X == 64) tmem_ld_f64(a, r);
    else if constexpr (X == 32) tmem_ld_x32(a, r);
    else tmem_ld_f16(a, r);
}
template <int X>
__global__ void __launch_bounds__(512, 1) ldtm_kernel(int nwarps, int iters, long long* out, float* sink) {
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    float acc = 0.f;
    long long t0 = 0, t1 = 0;
    if (warp < nwarps) {
        const uint32_t addr = tmem + ((uint32_t(warp & 3) * 32u) << 16) + (warp >> 2) * 64;
        float z[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) z[i] = float(i);
        tmem_st_x32(addr, z); tmem_st_x32(addr + 32, z + 32); tmem_st_wait();
        float ra[X], rb[X];
        t0 = clock64();
        ld<X>(addr, ra);
        for (int i = 0; i < iters; i += 2) {
            ld<X>(addr, rb);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < X; k += 8) acc += ra[k];
            ld<X>(addr, ra);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < X; k += 8) acc += rb[k];
        }
        tmem_ld_wait();
        acc += ra[1];
        t1 = clock64();
    }
    if (acc == 123.f) sink[0] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}
template <int X>
void run(long long* d, float* s) {
    for (int nw : {1, 2, 4, 8}) {
        ldtm_kernel<X><<<148, 512>>>(nw, 2000, d, s);
        cudaError_t e = cudaDeviceSynchronize();
        long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        printf("x%-2d warps=%2d: %7.1f cycles per load per warp (%d B) -> %5.1f B/clk/SM  [%s]\n", X, nw, double(c) / 2000, X * 128,
               X * 128.0 * nw / (double(c) / 2000), cudaGetErrorString(e));
    }
}
int main() {
    long long* d; float* s; cudaMalloc(&d, 8); cudaMalloc(&s, 4);
    run<64>(d, s); run<32>(d, s); run<16>(d, s);
    return 0;
}
