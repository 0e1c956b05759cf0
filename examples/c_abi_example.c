/* Plain-C use of the drop-in boundary (include/qattn.h): what a C / cgo / JNI caller of the library writes.
 * Pointers are device pointers obtained from the caller's own allocator (cudaMalloc, a framework's caching allocator);
 * the library never allocates and never synchronises.  Compile check (no GPU needed):
 *     gcc -std=c99 -Wall -Werror -Iinclude -c examples/c_abi_example.c -o /dev/null
 * Link against quantumattention_b200/libqattn_sm100.so to run.
 * Reference counterpart of the call below: `_fp8_attention_wrapper` + the op it calls (src/quantum_attn/nn.py:394-430). */
#include <stddef.h>
#include <stdio.h>
#include "qattn.h"

/* One FLUX-shaped fp8_attn_func call on bf16 inputs: quantise Q and K head-wise, fused forward in the default mode. */
int flux_attention(const void* q, const void* k, const void* v, /* bf16 [B,H,S,D], dense */
                   void* q8, void* k8, float* scale_q, float* scale_k, /* scratch: B*H*S*D bytes each, B*H floats each */
                   float* amax_ws, /* qa_quantize_workspace_floats(B, H, S, D) floats */
                   void* out, /* bf16 [B,H,S,D], 32-byte aligned */
                   void* stream) {
    const int B = 1, H = 24, S = 4608, D = 128;
    const float sm_scale = 0.08838834764831845f; /* 1 / sqrt(D) */
    if (qa_abi_version() != QA_ABI_VERSION) {
        fprintf(stderr, "qattn: header %d, library %d\n", QA_ABI_VERSION, qa_abi_version());
        return -1;
    }
    int rc = qa_fp8_attn_func(q, k, v, QA_DT_BF16, NULL, NULL, NULL, q8, k8, NULL, scale_q, scale_k, NULL, amax_ws,
                              0 /* plain scratch workspace */, out, NULL /* no LSE */, B, H, H, S, S, D, 0 /* not causal */,
                              sm_scale, QA_SCALE_HEAD, QA_P_16BIT, stream);
    if (rc != 0) fprintf(stderr, "qattn: %s\n", qa_last_error());
    return rc;
}

size_t flux_workspace_floats(void) { return qa_quantize_workspace_floats(1, 24, 4608, 128); }
