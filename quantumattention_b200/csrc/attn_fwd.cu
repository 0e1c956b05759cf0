// FP8 instantiations of the fused attention kernel (attn_fwd_kernel.cuh) and their dispatch: e4m3 Q / K with
// head-wise or token-wise scales, the three P modes, D in {64, 128, 256}, causal or not.
#include "attn_fwd_kernel.cuh"

namespace qa {

#ifdef QA_TRACE
long long* g_trace_ptr = nullptr;
int g_trace_x = 0, g_trace_y = 0;
#endif

#ifndef QA_FAST_BUILD
template <class C>
static int launch_flags(const AttnArgs& a, cudaStream_t stream, int* launches) {
    const bool token = a.scale_mode == QA_SCALE_TOKEN;
    if (a.causal) return token ? launch_cfg<C, true, true>(a, stream, launches) : launch_cfg<C, true, false>(a, stream, launches);
    return token ? launch_cfg<C, false, true>(a, stream, launches) : launch_cfg<C, false, false>(a, stream, launches);
}

template <int PMODE, bool PF16 = false>
static int launch_d(const AttnArgs& a, cudaStream_t stream, int* launches) {
    switch (a.D) {
        case 64: return launch_flags<AttnCfg<64, PMODE, false, PF16>>(a, stream, launches);
        case 128: return launch_flags<AttnCfg<128, PMODE, false, PF16>>(a, stream, launches);
        case 256: return launch_flags<AttnCfg<256, PMODE, false, PF16>>(a, stream, launches);
    }
    return set_error(QA_ERR_INVALID, "Unsupported head dimension: %d", a.D);
}
#endif

#ifdef QA_TRACE
extern "C" void qa_debug_set_trace(void* dev_ptr) { g_trace_ptr = static_cast<long long*>(dev_ptr); }
extern "C" void qa_debug_set_trace_cta(int x, int y) { g_trace_x = x, g_trace_y = y; }
#endif

int attn_fwd_dispatch(const AttnArgs& a, cudaStream_t stream, int* launches) {
#ifdef QA_FAST_BUILD  // developer switch: one instantiation only, for quick ptxas / SASS inspection
    if (a.p_mode != QA_P_E4M3 || a.D != 128 || a.causal || a.scale_mode == QA_SCALE_TOKEN)
        return set_error(QA_ERR_INVALID, "QA_FAST_BUILD library: only D=128, fp8 P, non-causal, head-wise scales");
    return launch_cfg<AttnCfg<128, QA_P_E4M3>, false, false>(a, stream, launches);
#else
    switch (a.p_mode) {
        case QA_P_E4M3: return launch_d<QA_P_E4M3>(a, stream, launches);
        case QA_P_E4M3_HILO: return launch_d<QA_P_E4M3_HILO>(a, stream, launches);
        case QA_P_16BIT:  // (P takes V's 16-bit type: a compile-time property of the kernel)
            return a.out_dtype == QA_DT_FP16 ? launch_d<QA_P_16BIT, true>(a, stream, launches)
                                             : launch_d<QA_P_16BIT, false>(a, stream, launches);
    }
    return set_error(QA_ERR_INVALID, "unknown p_mode %d", a.p_mode);
#endif
}

}  // namespace qa
