// 16-bit instantiations of the fused attention kernel (attn_fwd_kernel.cuh): bf16 / fp16 Q, K, V and P, QK^T and PV
// as tcgen05 kind::f16 - the `attn_func` path of the reference (src/quantum_attn/tk/attention.py with is_fp8 = false).
#include "attn_fwd_kernel.cuh"

namespace qa {

template <int D, bool PF16>
static int launch16t(const AttnArgs& a, cudaStream_t stream, int* launches) {
    using C = AttnCfg<D, QA_P_16BIT, true, PF16>;
    return a.causal ? launch_cfg<C, true, false>(a, stream, launches) : launch_cfg<C, false, false>(a, stream, launches);
}
template <int D>
static int launch16(const AttnArgs& a, cudaStream_t stream, int* launches) {
    // (P takes V's 16-bit type: a compile-time property of the kernel)
    return a.out_dtype == QA_DT_FP16 ? launch16t<D, true>(a, stream, launches) : launch16t<D, false>(a, stream, launches);
}

int attn16_fwd_dispatch(const AttnArgs& a, cudaStream_t stream, int* launches) {
    switch (a.D) {
        case 64: return launch16<64>(a, stream, launches);
        case 128: return launch16<128>(a, stream, launches);
        case 256: return launch16<256>(a, stream, launches);
    }
    return set_error(QA_ERR_INVALID, "Unsupported head dimension: %d", a.D);
}

}  // namespace qa
