// Fused FP8 attention forward for sm_100a:  S = Q K^T (tcgen05 kind::f8f6f4, fp32 in TMEM)  ->  online softmax in
// registers (one thread per query row, base-2 domain, lazy rescale)  ->  P written back into TMEM as e4m3  ->
// O += P V (tcgen05, A operand from TMEM, V straight from its [kv][D] layout as an MN-major B operand).
//
// What it replaces: `fwd_attend_ker` of the reference (src/quantum_attn/tk/attention.py:97-349) and its launcher
// (:355-647).  Same function (dequant scale folded into the exp2 argument as in :204-210,248-250; top-left causal
// mask as in :252-263; ragged tails via TMA zero fill + -inf columns as in :269-271), different machine:
//
//   CTA = 2 query tiles of 128 rows (ping-pong), 12 warps (3 warpgroups; setmaxnreg moves registers to softmax):
//     warps 0-3  softmax + correction + epilogue of query tile 0      (thread r <-> TMEM lane r <-> query row r)
//     warps 4-7  the same for query tile 1
//     warp  8    MMA issuer (one thread): QK_0, QK_1, PV_0, PV_1 interleaved so one tile's softmax hides behind
//                the other tile's MMAs
//     warp  9    TMA producer: Q once, K/V tiles through an mbarrier ring        (warps 10-11 idle)
//   TMEM (512 columns): S0 | S1 | O0 | O1, 128 columns each at D=128; P_t aliases the first columns of S_t.
//   K and V tiles are shared by both query tiles, halving L2->SMEM traffic per FLOP.
//
// Deliberately absent: any non-sm_100 path, any fallback.
#include <cmath>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "ptx.cuh"
#include "qattn_internal.h"
#include "tma_host.h"

namespace qa {

constexpr int BM = 128;  // query rows per tile (= TMEM lanes)
constexpr int BN = 128;  // keys per K/V tile
constexpr float kLog2e = 1.4426950408889634f;

template <int D_, int PMODE_>
struct AttnCfg {
    static constexpr int D = D_;
    static constexpr int PMODE = PMODE_;
    static constexpr bool V16 = (PMODE_ == QA_P_16BIT);
    static constexpr int NQ = (D_ <= 128) ? 2 : 1;  // O for two tiles does not fit TMEM at D = 256
    static constexpr int VB = V16 ? 2 : 1;          // bytes per V element
    // shared-memory tiles are stored as "boxes" whose rows are one swizzle span (<= 128 bytes) wide
    static constexpr int QK_ROW = D_ < 128 ? D_ : 128;            // bytes per box row for Q / K (1 byte / element)
    static constexpr int QK_BOXES = D_ / QK_ROW;
    static constexpr int QK_BOX_BYTES = QK_ROW * 128;
    static constexpr int V_ROW = (D_ * VB) < 128 ? (D_ * VB) : 128;
    static constexpr int V_BOXES = D_ * VB / V_ROW;
    static constexpr int V_BOX_BYTES = V_ROW * BN;
    static constexpr int Q_TILE = BM * D_;
    static constexpr int K_TILE = BN * D_;
    static constexpr int V_TILE = BN * D_ * VB;
    static constexpr int O_BOXES = D_ / 64;  // 16-bit output, 64 elements = 128 bytes per box row
    static constexpr int O_TILE = BM * D_ * 2;
    static constexpr int STAGES = (D_ == 64) ? 4 : (D_ == 128 ? (V16 ? 2 : 3) : 2);
    static constexpr int SMEM_Q = 0;
    static constexpr int SMEM_K = SMEM_Q + NQ * Q_TILE;
    static constexpr int SMEM_V = SMEM_K + STAGES * K_TILE;
    // O staging for the TMA store: its own region when two query tiles finish at different times; with a single
    // tile (D = 256) every MMA has retired before the epilogue, so the dead K ring is reused
    static constexpr int SMEM_O = (NQ == 2) ? SMEM_V + STAGES * V_TILE : SMEM_K;
    static constexpr int SMEM_BAR = (NQ == 2) ? SMEM_O + NQ * O_TILE : SMEM_V + STAGES * V_TILE;
    static_assert(NQ == 2 || STAGES * K_TILE >= O_TILE, "K ring too small to stage O");
    static constexpr int SMEM_TOTAL = SMEM_BAR + 256 + 1024;  // + barriers + alignment slack
    static_assert(SMEM_TOTAL <= 232448, "shared memory budget exceeded");
    static constexpr int NTHREADS = (NQ * 4 + 4) * 32;  // softmax warpgroups + one warpgroup holding the MMA / TMA warps
    // TMEM columns
    static constexpr int TM_S = 0;                        // S_t at t * 128
    static constexpr int TM_O = 256;                      // O_t at 256 + t * 128 (D <= 128), single O at D = 256
    static constexpr int P_COLS = V16 ? 64 : 32;          // columns holding one P tile
    static constexpr int TM_P_LO = 64;                    // hi/lo mode: second P tile at S_t + 64
    // softmax range management: p' = 2^KOFF * exp2(s - m_used), m_used may lag the true max by <= TAU (log2 units)
    static constexpr float KOFF = V16 ? 0.f : 4.f;
    static constexpr float TAU = V16 ? 8.f : 4.f;
};

struct AttnParams {
    const float* scale_q;
    const float* scale_k;
    const float* scale_v;
    float* lse;
    int B, Hq, Hkv, Sq, Skv;
    int causal;
    float sm_scale_log2;  // sm_scale * log2(e)
    int out_fp16;
    float inv_group;  // Hkv / Hq
};

struct Barriers {
    uint64_t q_full[2];
    uint64_t k_full[4], k_empty[4], v_full[4], v_empty[4];
    uint64_t s_full[2], p_full[2], o_full[2];
    uint32_t tmem_base;
};
static_assert(sizeof(Barriers) <= 256, "barrier block too large");

template <class C>
__device__ __forceinline__ uint32_t qk_koff(int k) {  // byte offset of the k-th 32-byte K slice inside a Q/K tile
    constexpr int per_box = C::QK_ROW / 32;
    return uint32_t(k / per_box) * C::QK_BOX_BYTES + uint32_t(k % per_box) * 32u;
}

template <class C, bool CAUSAL, bool TOKEN>
__global__ void __launch_bounds__(C::NTHREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, AttnParams p) {
    constexpr int D = C::D;
    constexpr int NQ = C::NQ;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    Barriers* bars = reinterpret_cast<Barriers*>(smem + C::SMEM_BAR);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int h = blockIdx.y, b = blockIdx.z;
    const int hkv = int((float(h) + 0.5f) * p.inv_group);  // exact for h < 2^16; avoids an integer-division call
    const int bh = b * p.Hq + h;
    const int bhkv = b * p.Hkv + hkv;
    // heavy (late) causal query blocks are scheduled first
    const int mblk = CAUSAL ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;
    const int m0 = mblk * (BM * NQ);
    const int nkv = (p.Skv + BN - 1) / BN;
    // per-tile trip counts: under a causal mask tile t only needs K/V tiles up to its own diagonal
    const int n_iter0 = CAUSAL ? min(nkv, m0 / BN + 1) : nkv;
    const int n_iter1 = CAUSAL ? min(nkv, (m0 + BM) / BN + 1) : nkv;
    auto n_iter = [&](int t) { return t == 0 ? n_iter0 : n_iter1; };
    const int n_max = n_iter(NQ - 1);

    // ------------------------------------------------------------------ one-time setup
    if (warp == 0) {
        tmem_alloc(&bars->tmem_base, 512);
        tmem_relinquish();
    }
    if (threadIdx.x == 32) {
        for (int t = 0; t < 2; ++t) {
            mbar_init(&bars->q_full[t], 1);
            mbar_init(&bars->s_full[t], 1);
            mbar_init(&bars->p_full[t], 128);
            mbar_init(&bars->o_full[t], 1);
        }
        for (int s = 0; s < 4; ++s) {
            mbar_init(&bars->k_full[s], 1);
            mbar_init(&bars->k_empty[s], 1);
            mbar_init(&bars->v_full[s], 1);
            mbar_init(&bars->v_empty[s], 1);
        }
        fence_barrier_init();
    }
    if (warp == NQ * 4 + 1 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        tma_prefetch_desc(&tmO);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    // register rebalancing (two-tile configs run 384 threads -> 168 registers each at launch): the softmax
    // warpgroups keep a whole 128-wide score row per thread in registers, the MMA / TMA warps need almost nothing
    if (warp >= NQ * 4) {
        if constexpr (NQ == 2) reg_dealloc<56>();
        if (warp == NQ * 4 + 1) {
        // =============================================================== TMA producer
        if (lane == 0) {
            for (int t = 0; t < NQ; ++t) {
                mbar_arrive_expect_tx(&bars->q_full[t], C::Q_TILE);
                for (int x = 0; x < C::QK_BOXES; ++x)
                    tma_load_3d(smem + C::SMEM_Q + t * C::Q_TILE + x * C::QK_BOX_BYTES, &tmQ, &bars->q_full[t],
                                x * C::QK_ROW, m0 + t * BM, bh, kEvictFirst);
            }
            for (int n = 0; n < n_max; ++n) {
                const int s = n % C::STAGES;
                const uint32_t ph = (n / C::STAGES) & 1;
                mbar_wait(&bars->k_empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&bars->k_full[s], C::K_TILE);
                for (int x = 0; x < C::QK_BOXES; ++x)
                    tma_load_3d(smem + C::SMEM_K + s * C::K_TILE + x * C::QK_BOX_BYTES, &tmK, &bars->k_full[s],
                                x * C::QK_ROW, n * BN, bhkv, kEvictLast);
                mbar_wait(&bars->v_empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&bars->v_full[s], C::V_TILE);
                for (int x = 0; x < C::V_BOXES; ++x)
                    tma_load_3d(smem + C::SMEM_V + s * C::V_TILE + x * C::V_BOX_BYTES, &tmV, &bars->v_full[s],
                                x * (C::V_ROW / C::VB), n * BN, bhkv, kEvictLast);
            }
        }
    } else if (warp == NQ * 4) {
        // =============================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc_qk = make_idesc(0, 0, 0, 0, BM, BN);
            constexpr uint32_t idesc_pv = C::V16 ? 0u : make_idesc(0, 0, 0, 1, BM, D);
            const uint32_t idesc_pv16 = make_idesc(p.out_fp16 ? 0 : 1, p.out_fp16 ? 0 : 1, 0, 1, BM, D);
            constexpr uint64_t qk_swz = (C::QK_ROW == 128) ? kSwz128 : kSwz64;
            constexpr uint64_t v_swz = (C::V_ROW == 128) ? kSwz128 : kSwz64;
            const uint32_t q_base = smem_u32(smem + C::SMEM_Q);
            const uint32_t k_base = smem_u32(smem + C::SMEM_K);
            const uint32_t v_base = smem_u32(smem + C::SMEM_V);

            auto issue_qk = [&](int t, int stage) {
#pragma unroll
                for (int k = 0; k < D / 32; ++k) {
                    uint64_t ad = make_smem_desc(q_base + t * C::Q_TILE + qk_koff<C>(k), 16, 8 * C::QK_ROW, qk_swz);
                    uint64_t bd = make_smem_desc(k_base + stage * C::K_TILE + qk_koff<C>(k), 16, 8 * C::QK_ROW, qk_swz);
                    umma_f8_ss(tmem + C::TM_S + t * 128, ad, bd, idesc_qk, k > 0);
                }
            };
            auto issue_pv = [&](int t, int stage, bool acc) {
                const uint32_t o_t = tmem + C::TM_O + (NQ == 2 ? t * 128 : 0);
                const uint32_t p_t = tmem + C::TM_S + t * 128;
                if constexpr (!C::V16) {
#pragma unroll
                    for (int k = 0; k < BN / 32; ++k) {  // 32 keys per instruction
                        uint64_t bd = make_smem_desc(v_base + stage * C::V_TILE + k * 32 * C::V_ROW, C::V_BOX_BYTES,
                                                     8 * C::V_ROW, v_swz);
                        umma_f8_ts(o_t, p_t + k * 8, bd, idesc_pv, (acc || k > 0) ? 1u : 0u);
                        if constexpr (C::PMODE == QA_P_E4M3_HILO)
                            umma_f8_ts(o_t, p_t + C::TM_P_LO + k * 8, bd, idesc_pv, 1u);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < BN / 16; ++k) {  // 16 keys per instruction
                        uint64_t bd = make_smem_desc(v_base + stage * C::V_TILE + k * 16 * C::V_ROW, C::V_BOX_BYTES,
                                                     8 * C::V_ROW, v_swz);
                        umma_f16_ts(o_t, p_t + k * 8, bd, idesc_pv16, (acc || k > 0) ? 1u : 0u);
                    }
                }
            };

            // prologue: S_t = Q_t K_0^T
            mbar_wait(&bars->k_full[0], 0);
            for (int t = 0; t < NQ; ++t) {
                mbar_wait(&bars->q_full[t], 0);
                tc_fence_after();
                issue_qk(t, 0);
                umma_commit(&bars->s_full[t]);
            }
            umma_commit(&bars->k_empty[0]);

            for (int n = 0; n < n_max; ++n) {
                const int sv = n % C::STAGES;
                const int sk = (n + 1) % C::STAGES;
                const uint32_t ph_v = (n / C::STAGES) & 1;
                const uint32_t ph_k = ((n + 1) / C::STAGES) & 1;
                mbar_wait(&bars->v_full[sv], ph_v);
                bool k_ready = false;
#pragma unroll
                for (int t = 0; t < NQ; ++t) {
                    if (n < n_iter(t)) {
                        mbar_wait(&bars->p_full[t], n & 1);
                        tc_fence_after();
                        issue_pv(t, sv, n > 0);
                    }
                    if (t == NQ - 1) umma_commit(&bars->v_empty[sv]);
                    if (n + 1 < n_iter(t)) {
                        if (!k_ready) {
                            mbar_wait(&bars->k_full[sk], ph_k);
                            tc_fence_after();
                            k_ready = true;
                        }
                        issue_qk(t, sk);
                        umma_commit(&bars->s_full[t]);
                    }
                    if (t == NQ - 1 && n + 1 < n_max) umma_commit(&bars->k_empty[sk]);
                }
            }
            for (int t = 0; t < NQ; ++t) umma_commit(&bars->o_full[t]);
        }
    }
    } else {
        // =============================================================== softmax / correction / epilogue
        if constexpr (NQ == 2) reg_alloc<224>();
        const int t = warp >> 2;                       // query tile of this warpgroup
        const int row = ((warp & 3) << 5) | lane;      // row inside the tile == TMEM lane
        const uint32_t lane_base = uint32_t((warp & 3) * 32) << 16;
        const uint32_t s_addr = tmem + lane_base + C::TM_S + t * 128;
        const uint32_t o_addr = tmem + lane_base + C::TM_O + (NQ == 2 ? t * 128 : 0);
        const int row_g = m0 + t * BM + row;

        float c;  // multiplier taking raw fp8 dot products to the base-2 softmax domain
        if constexpr (TOKEN) {
            c = p.scale_q[size_t(bh) * p.Sq + min(row_g, p.Sq - 1)] * p.sm_scale_log2;
        } else {
            c = p.scale_q[bh] * p.scale_k[bhkv] * p.sm_scale_log2;
        }
        const float* sk_row = TOKEN ? p.scale_k + size_t(bhkv) * p.Skv : nullptr;

        float m_used = -INFINITY;  // running max in raw score units (times per-column scale in token mode)
        float l = 0.f;             // running sum of p' = 2^KOFF * exp2(c * (s - m_used))
        const int my_iters = n_iter(t);

        for (int n = 0; n < my_iters; ++n) {
            mbar_wait(&bars->s_full[t], n & 1);
            tc_fence_after();
            float s[128];
            tmem_ld_x32(s_addr + 0, &s[0]);
            tmem_ld_x32(s_addr + 32, &s[32]);
            tmem_ld_x32(s_addr + 64, &s[64]);
            tmem_ld_x32(s_addr + 96, &s[96]);
            tmem_ld_wait();
            const int col0 = n * BN;
            if constexpr (TOKEN) {
                if (col0 + BN <= p.Skv) {
#pragma unroll
                    for (int j = 0; j < 128; j += 4) {
                        float4 k4 = __ldg(reinterpret_cast<const float4*>(sk_row + col0 + j));
                        s[j] *= k4.x, s[j + 1] *= k4.y, s[j + 2] *= k4.z, s[j + 3] *= k4.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 128; ++j) s[j] *= __ldg(sk_row + min(col0 + j, p.Skv - 1));
                }
            }
            // masking: only tiles that touch the causal diagonal or the ragged tail pay for it
            const bool tail = col0 + BN > p.Skv;
            const bool diag = CAUSAL && (col0 + BN - 1 > m0 + t * BM);
            if (tail || diag) {
                const int lim = CAUSAL ? min(p.Skv - 1, row_g) : (p.Skv - 1);  // last visible column
#pragma unroll
                for (int j = 0; j < 128; ++j)
                    if (col0 + j > lim) s[j] = -INFINITY;
            }
            float mx0 = fmaxf(s[0], s[1]), mx1 = fmaxf(s[2], s[3]);
#pragma unroll
            for (int j = 4; j < 128; j += 4) {
                mx0 = fmaxf(mx0, fmaxf(s[j], s[j + 1]));
                mx1 = fmaxf(mx1, fmaxf(s[j + 2], s[j + 3]));
            }
            const float m_new = fmaxf(m_used, fmaxf(mx0, mx1));
            // lazy rescale: keep the stale max while the true max has grown by < 2^TAU
            const bool grow = (m_new - m_used) * c > C::TAU;
            if (__any_sync(0xffffffffu, grow)) {
                const float alpha = ex2_approx((m_used - m_new) * c);  // 0 on the first tile
                m_used = m_new;
                l *= alpha;
                if (n > 0) {
#pragma unroll
                    for (int cc = 0; cc < D; cc += 32) {
                        float o[32];
                        tmem_ld_x32(o_addr + cc, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) o[j] *= alpha;
                        tmem_st_x32(o_addr + cc, o);
                    }
                }
            }
            const float neg = C::KOFF - m_used * c;
            float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
            if constexpr (C::PMODE == QA_P_E4M3) {
                uint32_t pw[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float p0 = ex2_approx(fmaf(s[4 * j + 0], c, neg));
                    const float p1 = ex2_approx(fmaf(s[4 * j + 1], c, neg));
                    const float p2 = ex2_approx(fmaf(s[4 * j + 2], c, neg));
                    const float p3 = ex2_approx(fmaf(s[4 * j + 3], c, neg));
                    l0 += p0, l1 += p1, l2 += p2, l3 += p3;
                    pw[j] = pack_e4m3x4(p0, p1, p2, p3);
                }
                tmem_st_x32(s_addr, pw);
            } else if constexpr (C::PMODE == QA_P_E4M3_HILO) {
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float p0 = ex2_approx(fmaf(s[4 * j + 0], c, neg));
                    const float p1 = ex2_approx(fmaf(s[4 * j + 1], c, neg));
                    const float p2 = ex2_approx(fmaf(s[4 * j + 2], c, neg));
                    const float p3 = ex2_approx(fmaf(s[4 * j + 3], c, neg));
                    l0 += p0, l1 += p1, l2 += p2, l3 += p3;
                    const uint32_t h01 = cvt_e4m3x2(p0, p1), h23 = cvt_e4m3x2(p2, p3);
                    const float2 f01 = e4m3x2_to_float2(h01), f23 = e4m3x2_to_float2(h23);
                    hi[j] = h01 | (h23 << 16);
                    lo[j] = pack_e4m3x4(p0 - f01.x, p1 - f01.y, p2 - f23.x, p3 - f23.y);
                }
                tmem_st_x32(s_addr, hi);
                tmem_st_x32(s_addr + C::TM_P_LO, lo);
            } else {
                uint32_t pw[64];
#pragma unroll
                for (int j = 0; j < 64; j += 2) {
                    const float p0 = ex2_approx(fmaf(s[2 * j + 0], c, neg));
                    const float p1 = ex2_approx(fmaf(s[2 * j + 1], c, neg));
                    const float p2 = ex2_approx(fmaf(s[2 * j + 2], c, neg));
                    const float p3 = ex2_approx(fmaf(s[2 * j + 3], c, neg));
                    l0 += p0, l1 += p1, l2 += p2, l3 += p3;
                    pw[j] = p.out_fp16 ? pack_f16x2(p0, p1) : pack_bf16x2(p0, p1);
                    pw[j + 1] = p.out_fp16 ? pack_f16x2(p2, p3) : pack_bf16x2(p2, p3);
                }
                tmem_st_x32(s_addr, *reinterpret_cast<uint32_t(*)[32]>(&pw[0]));
                tmem_st_x32(s_addr + 32, *reinterpret_cast<uint32_t(*)[32]>(&pw[32]));
            }
            l += (l0 + l1) + (l2 + l3);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&bars->p_full[t]);
        }

        // ---------------------------------------------------------------- epilogue: O / l -> 16 bit -> smem -> TMA
        mbar_wait(&bars->o_full[t], 0);
        tc_fence_after();
        const float sv = C::V16 ? 1.f : p.scale_v[bhkv];
        const float inv = __fdividef(sv, l);
        uint8_t* o_smem = smem + C::SMEM_O + t * C::O_TILE;
#pragma unroll
        for (int cc = 0; cc < D; cc += 32) {
            float o[32];
            tmem_ld_x32(o_addr + cc, o);
            tmem_ld_wait();
            uint32_t w[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float a = o[2 * j] * inv, bb = o[2 * j + 1] * inv;
                w[j] = p.out_fp16 ? pack_f16x2(a, bb) : pack_bf16x2(a, bb);
            }
            // 64 output columns (128 bytes) per box row, 128B-swizzled so the TMA store un-swizzles it
            uint8_t* box = o_smem + (cc >> 6) * (BM * 128) + row * 128;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int chunk = ((cc & 63) >> 3) + q;
                *reinterpret_cast<uint4*>(box + ((chunk ^ (row & 7)) << 4)) =
                    make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
            }
        }
        if (p.lse != nullptr && row_g < p.Sq)
            p.lse[size_t(bh) * p.Sq + row_g] = (m_used * c + (__log2f(l) - C::KOFF)) * 0.6931471805599453f;
        fence_proxy_async_smem();
        named_bar_sync(1 + t, 128);
        if ((warp & 3) == 0 && lane == 0 && m0 + t * BM < p.Sq) {
            for (int x = 0; x < C::O_BOXES; ++x) tma_store_3d(&tmO, o_smem + x * (BM * 128), x * 64, m0 + t * BM, bh);
            tma_store_commit();
            tma_store_wait_all<0>();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ host side
template <class C, bool CAUSAL, bool TOKEN>
static int launch_cfg(const AttnArgs& a, cudaStream_t stream, int* launches) {
    CUtensorMap tmQ, tmK, tmV, tmO;
    const CUtensorMapSwizzle qk_swz = C::QK_ROW == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    const CUtensorMapSwizzle v_swz = C::V_ROW == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    const uint64_t D = C::D;
    bool ok = true;
    ok &= make_tmap_3d(&tmQ, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, a.q8, D, a.Sq, uint64_t(a.B) * a.Hq, D, D * a.Sq,
                       C::QK_ROW, BM, qk_swz);
    ok &= make_tmap_3d(&tmK, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, a.k8, D, a.Skv, uint64_t(a.B) * a.Hkv, D, D * a.Skv,
                       C::QK_ROW, BN, qk_swz);
    if (C::V16) {
        const CUtensorMapDataType dt =
            a.v_dtype == QA_DT_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
        ok &= make_tmap_3d(&tmV, dt, 2, a.v, D, a.Skv, uint64_t(a.B) * a.Hkv, D * 2, D * 2 * a.Skv, C::V_ROW / 2, BN,
                           v_swz);
    } else {
        ok &= make_tmap_3d(&tmV, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, a.v, D, a.Skv, uint64_t(a.B) * a.Hkv, D, D * a.Skv,
                           C::V_ROW, BN, v_swz);
    }
    const CUtensorMapDataType odt =
        a.out_dtype == QA_DT_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    ok &= make_tmap_3d(&tmO, odt, 2, a.out, D, a.Sq, uint64_t(a.B) * a.Hq, D * 2, D * 2 * a.Sq, 64, BM,
                       CU_TENSOR_MAP_SWIZZLE_128B);
    if (!ok) return set_error(QA_ERR_DEVICE, "cuTensorMapEncodeTiled failed or is unavailable (no CUDA driver?)");

    AttnParams p;
    p.scale_q = a.scale_q;
    p.scale_k = a.scale_k;
    p.scale_v = a.scale_v;
    p.lse = a.lse;
    p.B = a.B, p.Hq = a.Hq, p.Hkv = a.Hkv, p.Sq = a.Sq, p.Skv = a.Skv;
    p.causal = a.causal;
    p.sm_scale_log2 = a.sm_scale * kLog2e;
    p.out_fp16 = (a.out_dtype == QA_DT_FP16);
    p.inv_group = float(a.Hkv) / float(a.Hq);

    auto kern = attn_fwd_kernel<C, CAUSAL, TOKEN>;
    static bool attr_done = false;  // per instantiation; racing threads set the same value
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_TOTAL);
        if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(max dynamic smem)", e);
        attr_done = true;
    }
    dim3 grid((a.Sq + BM * C::NQ - 1) / (BM * C::NQ), a.Hq, a.B);
    kern<<<grid, C::NTHREADS, C::SMEM_TOTAL, stream>>>(tmQ, tmK, tmV, tmO, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("attn_fwd_kernel launch", e);
    *launches += 1;
    return QA_OK;
}

#ifndef QA_FAST_BUILD
template <class C>
static int launch_flags(const AttnArgs& a, cudaStream_t stream, int* launches) {
    const bool token = a.scale_mode == QA_SCALE_TOKEN;
    if (a.causal) return token ? launch_cfg<C, true, true>(a, stream, launches) : launch_cfg<C, true, false>(a, stream, launches);
    return token ? launch_cfg<C, false, true>(a, stream, launches) : launch_cfg<C, false, false>(a, stream, launches);
}

template <int PMODE>
static int launch_d(const AttnArgs& a, cudaStream_t stream, int* launches) {
    switch (a.D) {
        case 64: return launch_flags<AttnCfg<64, PMODE>>(a, stream, launches);
        case 128: return launch_flags<AttnCfg<128, PMODE>>(a, stream, launches);
        case 256: return launch_flags<AttnCfg<256, PMODE>>(a, stream, launches);
    }
    return set_error(QA_ERR_INVALID, "Unsupported head dimension: %d", a.D);
}
#endif

int attn_fwd_dispatch(const AttnArgs& a, cudaStream_t stream, int* launches) {
#ifdef QA_FAST_BUILD  // developer switch: one instantiation only, for quick ptxas / SASS inspection
    return launch_cfg<AttnCfg<128, QA_P_E4M3>, false, false>(a, stream, launches);
#else
    switch (a.p_mode) {
        case QA_P_E4M3: return launch_d<QA_P_E4M3>(a, stream, launches);
        case QA_P_E4M3_HILO: return launch_d<QA_P_E4M3_HILO>(a, stream, launches);
        case QA_P_16BIT: return launch_d<QA_P_16BIT>(a, stream, launches);
    }
    return set_error(QA_ERR_INVALID, "unknown p_mode %d", a.p_mode);
#endif
}

}  // namespace qa
