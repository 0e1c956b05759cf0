// Developer microbenchmark: issue cost of packed fp32x2 FFMA2 against scalar FFMA on sm_100a (one warp-instruction each).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(float* out, long long* cyc, int rounds, float a, float b) {
    float2 x[8];
    for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    __syncthreads();
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) {  // 2 scalar FFMA per element pair
                    x[i].x = fmaf(x[i].x, a, b);
                    x[i].y = fmaf(x[i].y, a, b);
                } else {          // 1 FFMA2 per element pair
                    x[i] = __ffma2_rn(x[i], a2, b2);
                }
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int rounds = 2000;
    for (int warps : {4, 8, 16, 32}) {
        for (int mode = 0; mode < 2; ++mode) {
            if (mode == 0) k<0><<<148, warps * 32>>>(out, cyc, rounds, 1.0001f, 0.5f);
            else k<1><<<148, warps * 32>>>(out, cyc, rounds, 1.0001f, 0.5f);
            cudaDeviceSynchronize();
            long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
            const double pairs = double(rounds) * 32 * (warps / 4.0);  // element pairs per lane per SM sub-partition
            printf("%s warps/SM %2d: %.3f cycles per element pair per sub-partition\n", mode ? "FFMA2     " : "2 x FFMA  ", warps, avg / pairs);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
