// qa_probe: single-CTA known-answer tests of the tcgen05 / TMA building blocks the attention kernel is built from.
// Each case loads operand tiles with TMA (hardware swizzle), issues the MMAs with hand-built descriptors, reads the
// accumulator back with tcgen05.ld and compares EXACTLY against a host computation (operands are small dyadic
// rationals, so fp32 accumulation is exact in any order).  Run on the GPU box before trusting the big kernel:
//   ./qa_probe            -> prints PASS/FAIL per case, exit code = number of failures
// This is test infrastructure, not product code.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include "ptx.cuh"
#include "tma_host.h"

using namespace qa;

struct ProbeParams {
    int ts;            // 0: A from smem (SS), 1: A from TMEM (TS)
    int kind_f16;      // 0: kind::f8f6f4, 1: kind::f16
    int nk;            // number of MMA instructions along K
    int N;             // MMA N
    uint32_t idesc;
    // A (smem)
    int a_nbox, a_box_bytes, a_box_elems, a_kpb, a_kstep, a_lbo, a_sbo, a_swz;
    // A (tmem): words per row to stage, columns advanced per k step
    int a_words, a_tmem_kstep;
    // B (smem)
    int b_nbox, b_box_bytes, b_box_elems, b_kpb, b_kstep, b_lbo, b_sbo, b_swz;
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const uint32_t* __restrict__ a_words_g, float* __restrict__ out, ProbeParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + 65536;
    __shared__ uint64_t bar_load, bar_mma;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        tmem_alloc(&tmem_base_s, 512);
        tmem_relinquish();
    }
    if (tid == 0) {
        mbar_init(&bar_load, 1);
        mbar_init(&bar_mma, 1);
        fence_barrier_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_base = (uint32_t(warp) * 32u) << 16;

    if (tid == 0) {
        uint32_t bytes = p.b_nbox * p.b_box_bytes + (p.ts ? 0 : p.a_nbox * p.a_box_bytes);
        mbar_arrive_expect_tx(&bar_load, bytes);
        if (!p.ts)
            for (int i = 0; i < p.a_nbox; ++i)
                tma_load_3d(sA + i * p.a_box_bytes, &tmA, &bar_load, i * p.a_box_elems, 0, 0, kEvictNormal);
        for (int i = 0; i < p.b_nbox; ++i)
            tma_load_3d(sB + i * p.b_box_bytes, &tmB, &bar_load, i * p.b_box_elems, 0, 0, kEvictNormal);
    }
    if (p.ts) {
        // stage A rows into TMEM columns [256, 256 + a_words): thread t owns row t
        for (int c = 0; c < p.a_words; c += 8) {
            uint32_t r[8];
            for (int j = 0; j < 8; ++j) r[j] = a_words_g[tid * p.a_words + c + j];
            tmem_st_x8(tmem + lane_base + 256 + c, r);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (tid == 0) {
        mbar_wait(&bar_load, 0);
        tc_fence_after();
        for (int k = 0; k < p.nk; ++k) {
            uint32_t b_off = (k / p.b_kpb) * p.b_box_bytes + (k % p.b_kpb) * p.b_kstep;
            uint64_t bdesc = make_smem_desc(smem_u32(sB) + b_off, p.b_lbo, p.b_sbo, uint64_t(p.b_swz));
            if (p.ts) {
                uint32_t a_t = tmem + 256 + k * p.a_tmem_kstep;
                if (p.kind_f16) umma_f16_ts(tmem, a_t, bdesc, p.idesc, k > 0);
                else umma_f8_ts(tmem, a_t, bdesc, p.idesc, k > 0);
            } else {
                uint32_t a_off = (k / p.a_kpb) * p.a_box_bytes + (k % p.a_kpb) * p.a_kstep;
                uint64_t adesc = make_smem_desc(smem_u32(sA) + a_off, p.a_lbo, p.a_sbo, uint64_t(p.a_swz));
                if (p.kind_f16) umma_f16_ss(tmem, adesc, bdesc, p.idesc, k > 0);
                else umma_f8_ss(tmem, adesc, bdesc, p.idesc, k > 0);
            }
        }
        umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    for (int c = 0; c < p.N; c += 32) {
        uint32_t r[32];
        tmem_ld_x32(tmem + lane_base + c, r);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[tid * p.N + c + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------- host side
static float e4m3_decode(uint8_t b) {
    int s = b >> 7, e = (b >> 3) & 15, m = b & 7;
    float v;
    if (e == 0) v = ldexpf(float(m), -9);
    else if (e == 15 && m == 7) v = NAN;
    else v = ldexpf(float(8 + m), e - 10);
    return s ? -v : v;
}
static uint8_t e4m3_encode_exact(float f) {
    for (int b = 0; b < 256; ++b)
        if (e4m3_decode(uint8_t(b)) == f && !(b == 0x80 && f == 0.f)) return uint8_t(b);
    fprintf(stderr, "value %f not representable\n", f);
    exit(99);
}
static uint32_t rng_state = 12345u;
static uint32_t rnd() {
    rng_state = rng_state * 1664525u + 1013904223u;
    return rng_state >> 8;
}
static float rnd_val() {  // dyadic values in [-2, 2] step 0.25
    return float(int(rnd() % 17) - 8) * 0.25f;
}

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);     \
            return 1;                                                                           \
        }                                                                                       \
    } while (0)

// One case. A is logically [128][K] (row-major, "K-major"); B is either [N][K] row-major (K-major, b_mn = 0) or
// [K][N] row-major (MN-major, b_mn = 1).  elem = 1 (e4m3) or 2 (bf16) bytes.
static int run_case(const char* name, int elem, int ts, int N, int K, int b_mn, int swz_bytes_a, int swz_bytes_b,
                    int lbo_override = -1, int sbo_override = -1) {
    const int M = 128;
    std::vector<float> A(M * K), B(size_t(N) * K);  // B stored logically as B[n][k]
    for (auto& v : A) v = rnd_val();
    for (auto& v : B) v = rnd_val();
    std::vector<float> ref(size_t(M) * N, 0.f);
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float acc = 0;
            for (int k = 0; k < K; ++k) acc += A[m * K + k] * B[size_t(n) * K + k];
            ref[size_t(m) * N + n] = acc;
        }
    // device images
    std::vector<uint8_t> Ab(size_t(M) * K * elem), Bb(size_t(N) * K * elem);
    auto put = [&](std::vector<uint8_t>& dst, size_t idx, float v) {
        if (elem == 1) dst[idx] = e4m3_encode_exact(v);
        else {
            __nv_bfloat16 h = __float2bfloat16(v);
            memcpy(&dst[idx * 2], &h, 2);
        }
    };
    for (int m = 0; m < M; ++m)
        for (int k = 0; k < K; ++k) put(Ab, size_t(m) * K + k, A[m * K + k]);
    if (!b_mn) {
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < K; ++k) put(Bb, size_t(n) * K + k, B[size_t(n) * K + k]);
    } else {
        for (int k = 0; k < K; ++k)
            for (int n = 0; n < N; ++n) put(Bb, size_t(k) * N + n, B[size_t(n) * K + k]);
    }
    uint8_t *dA, *dB;
    float* dOut;
    CK(cudaMalloc(&dA, Ab.size()));
    CK(cudaMalloc(&dB, Bb.size()));
    CK(cudaMalloc(&dOut, size_t(M) * N * 4));
    CK(cudaMemcpy(dA, Ab.data(), Ab.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bb.data(), Bb.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dOut, 0xFF, size_t(M) * N * 4));

    ProbeParams p{};
    p.ts = ts;
    p.kind_f16 = (elem == 2);
    const int kper = 32 / elem;  // K elements per MMA
    p.nk = K / kper;
    p.N = N;
    p.idesc = make_idesc(elem == 2 ? 1 : 0, elem == 2 ? 1 : 0, 0, b_mn, 128, N);
    auto swz_enum = [](int b) { return b == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : b == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B; };
    auto swz_desc = [](int b) { return b == 128 ? int(kSwz128) : b == 64 ? int(kSwz64) : int(kSwz32); };
    CUtensorMapDataType dt = elem == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    CUtensorMap tmA{}, tmB{};
    // ---- A: [128 rows][K] K-major, boxes of swz_bytes_a along K
    {
        int row_bytes = K * elem;
        p.a_box_elems = swz_bytes_a / elem;
        p.a_nbox = row_bytes / swz_bytes_a;
        p.a_box_bytes = swz_bytes_a * 128;
        p.a_kpb = swz_bytes_a / 32;
        p.a_kstep = 32;
        p.a_lbo = 16;
        p.a_sbo = 8 * swz_bytes_a;
        p.a_swz = swz_desc(swz_bytes_a);
        p.a_words = row_bytes / 4;
        p.a_tmem_kstep = 8;
        if (!make_tmap_3d(&tmA, dt, elem, dA, K, M, 1, uint64_t(K) * elem, uint64_t(K) * elem * M, p.a_box_elems, 128,
                          swz_enum(swz_bytes_a))) {
            printf("%s: tensor map A failed\n", name);
            return 1;
        }
    }
    // ---- B
    if (!b_mn) {  // [N rows][K], boxes of swz_bytes_b along K, N rows each
        int row_bytes = K * elem;
        p.b_box_elems = swz_bytes_b / elem;
        p.b_nbox = row_bytes / swz_bytes_b;
        p.b_box_bytes = swz_bytes_b * N;
        p.b_kpb = swz_bytes_b / 32;
        p.b_kstep = 32;
        p.b_lbo = 16;
        p.b_sbo = 8 * swz_bytes_b;
        p.b_swz = swz_desc(swz_bytes_b);
        if (!make_tmap_3d(&tmB, dt, elem, dB, K, N, 1, uint64_t(K) * elem, uint64_t(K) * elem * N, p.b_box_elems, N,
                          swz_enum(swz_bytes_b))) {
            printf("%s: tensor map B failed\n", name);
            return 1;
        }
    } else {  // [K rows][N], boxes of swz_bytes_b along N, K rows each
        int row_bytes = N * elem;
        p.b_box_elems = swz_bytes_b / elem;
        p.b_nbox = row_bytes / swz_bytes_b;
        p.b_box_bytes = swz_bytes_b * K;
        p.b_kpb = 1 << 20;  // all k steps inside "box 0" addressing: step over rows
        p.b_kstep = kper * swz_bytes_b;  // kper rows of swz_bytes_b bytes
        p.b_lbo = p.b_box_bytes;         // distance between N-atoms (boxes)
        p.b_sbo = 8 * swz_bytes_b;       // distance between 8-row groups along K
        p.b_swz = swz_desc(swz_bytes_b);
        if (!make_tmap_3d(&tmB, dt, elem, dB, N, K, 1, uint64_t(N) * elem, uint64_t(N) * elem * K, p.b_box_elems, K,
                          swz_enum(swz_bytes_b))) {
            printf("%s: tensor map B failed\n", name);
            return 1;
        }
    }
    if (lbo_override >= 0) p.b_lbo = lbo_override;
    if (sbo_override >= 0) p.b_sbo = sbo_override;

    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 * 2 + 2048));
    probe_kernel<<<1, 128, 65536 * 2 + 2048>>>(tmA, tmB, reinterpret_cast<const uint32_t*>(dA), dOut, p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("%-44s FAIL (CUDA: %s)\n", name, cudaGetErrorString(e));
        return 1;
    }
    std::vector<float> got(size_t(M) * N);
    CK(cudaMemcpy(got.data(), dOut, got.size() * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    double maxerr = 0;
    for (size_t i = 0; i < got.size(); ++i) {
        double d = fabs(double(got[i]) - double(ref[i]));
        if (!(d == 0)) {
            if (bad < 4) printf("   mismatch m=%zu n=%zu got %g want %g\n", i / N, i % N, got[i], ref[i]);
            ++bad;
        }
        if (d > maxerr || d != d) maxerr = d;
    }
    printf("%-44s %s  (mismatches %zu / %zu, max err %g, b_lbo %d b_sbo %d)\n", name, bad ? "FAIL" : "PASS", bad,
           got.size(), maxerr, p.b_lbo, p.b_sbo);
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dOut);
    return bad ? 1 : 0;
}

int main() {
    int fails = 0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) {
        printf("no CUDA device\n");
        return 100;
    }
    printf("device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
    // S = Q K^T, e4m3, both operands K-major from shared memory
    fails += run_case("qk  e4m3 SS D=128 sw128", 1, 0, 128, 128, 0, 128, 128);
    fails += run_case("qk  e4m3 SS D=64  sw64", 1, 0, 128, 64, 0, 64, 64);
    fails += run_case("qk  e4m3 SS D=256 sw128 x2 boxes", 1, 0, 128, 256, 0, 128, 128);
    fails += run_case("qk  e4m3 SS D=128 N=256", 1, 0, 256, 128, 0, 128, 128);
    // O = P V, P (e4m3) from TMEM, V [kv][D] MN-major from shared memory
    fails += run_case("pv  e4m3 TS D=128 V mn-major sw128", 1, 1, 128, 128, 1, 128, 128);
    fails += run_case("pv  e4m3 TS D=64  V mn-major sw64", 1, 1, 64, 128, 1, 128, 64);
    fails += run_case("pv  e4m3 TS D=256 V mn-major sw128 x2", 1, 1, 256, 128, 1, 128, 128);
    // fallback layout: V^T [D][kv] K-major
    fails += run_case("pv  e4m3 TS D=128 V^T k-major sw128", 1, 1, 128, 128, 0, 128, 128);
    // same with P and V in bf16 (the reference's PV precision)
    fails += run_case("pv  bf16 TS D=128 V mn-major sw128 x2", 2, 1, 128, 128, 1, 128, 128);
    fails += run_case("pv  bf16 TS D=64  V mn-major sw128", 2, 1, 64, 128, 1, 128, 128);
    fails += run_case("qk  bf16 SS D=128 sw128 x2", 2, 0, 128, 128, 0, 128, 128);
    if (fails) {
        // diagnostics for the multi-atom MN-major cases: try the other LBO/SBO conventions
        printf("-- variants --\n");
        run_case("pv e4m3 D=256 mn: lbo<->sbo swapped", 1, 1, 256, 128, 1, 128, 128, 1024, 16384);
        run_case("pv bf16 D=128 mn: lbo<->sbo swapped", 2, 1, 128, 128, 1, 128, 128, 1024, 16384);
        run_case("pv e4m3 D=128 mn: sbo=128*8,lbo=16", 1, 1, 128, 128, 1, 128, 128, 16, 1024);
    }
    printf("probe: %d failing case(s)\n", fails);
    return fails;
}
