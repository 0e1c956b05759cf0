/* Plain-C restatement of the reference's dynamic FP8 quantiser (TEST INFRASTRUCTURE - never linked into the product).
 *
 * Follows src/quantum_attn/nn.py:14-19 with fp32 intermediates:
 *     scale = max(amax(|t|) * (1/448), FLT_EPSILON);   t8 = e4m3fn_rne(clamp(t / scale, -448, 448))
 * and the two reduction layouts of src/quantum_attn/nn.py:410-418 (head-wise: over S and D; token-wise: over D).
 * The e4m3fn encoder is written from the format definition (1-4-3, bias 7, max 448, no inf) with
 * round-to-nearest-even, independently of the numpy version in oracle/e4m3.py, so the two check each other.
 * Parity is pinned by tests/golden/quantize_*.npz (outputs of the reference itself, see oracle/gen_golden.py).
 *
 * Build: gcc -O2 -shared -fPIC -ffp-contract=off -o oracle/_build/libqa_oracle.so oracle/quantize_ref.c -lm
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

static uint8_t e4m3_encode(float x) {
    uint32_t bits;
    memcpy(&bits, &x, 4);
    uint8_t sign = (uint8_t)((bits >> 31) << 7);
    if (x != x) return 0x7F;
    float a = fabsf(x);
    if (a >= 448.0f) return sign | 0x7E; /* saturate (callers clamp first) */
    if (a < 0.0009765625f) {             /* below half of the smallest subnormal 2^-9 ... handled by rounding */
        /* fallthrough to generic path: values < 2^-10 round to 0 or 2^-9 */
    }
    int e;
    (void)frexpf(a, &e); /* a = m * 2^e, m in [0.5,1) -> exponent of leading bit is e-1 */
    int lead = (a == 0.0f) ? -127 : e - 1;
    if (lead < -6) lead = -6;                 /* subnormal range shares the step 2^-9 */
    double step = ldexp(1.0, lead - 3);       /* 3 mantissa bits */
    double q = (double)a / step;              /* exact */
    double r = nearbyint(q);                  /* default rounding mode: half to even */
    double v = r * step;
    if (v >= 448.0) return sign | 0x7E;
    if (v == 0.0) return sign;
    int e2;
    (void)frexp(v, &e2);
    int lead2 = e2 - 1;
    if (lead2 < -6) { /* subnormal: value = m * 2^-9 */
        return sign | (uint8_t)(v / ldexp(1.0, -9));
    }
    int mant = (int)(v / ldexp(1.0, lead2 - 3)) - 8;
    return sign | (uint8_t)((lead2 + 7) << 3) | (uint8_t)mant;
}

static float scale_of(float amax) {
    float s = amax * (1.0f / 448.0f);
    return s < FLT_EPSILON ? FLT_EPSILON : s;
}

static void quant_span(const float* x, uint8_t* out, long n, float scale) {
    for (long i = 0; i < n; ++i) {
        float y = x[i] / scale;
        if (y > 448.0f) y = 448.0f;
        if (y < -448.0f) y = -448.0f;
        out[i] = e4m3_encode(y);
    }
}

/* x: fp32 [groups][len]; one scale per group.  head-wise: groups = B*H, len = S*D.  token-wise: groups = B*H*S, len = D */
void qa_oracle_quantize(const float* x, uint8_t* out, float* scale, long groups, long len) {
    for (long g = 0; g < groups; ++g) {
        const float* p = x + g * len;
        float amax = 0.0f;
        for (long i = 0; i < len; ++i) {
            float a = fabsf(p[i]);
            if (a > amax) amax = a;
        }
        scale[g] = scale_of(amax);
        quant_span(p, out + g * len, len, scale[g]);
    }
}

uint8_t qa_oracle_e4m3_encode(float x) { return e4m3_encode(x); }
