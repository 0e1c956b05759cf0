/* Developer tool: exhaustive check of the quantiser's shared-divisor division over its whole input domain.
 *
 * The head-wise quantiser divides 16-bit inputs x by scale = max(amax * (1/448), FLT_EPSILON) where amax is itself a
 * 16-bit value and |x| <= amax.  That is < 2^30 (amax, |x|) pairs per dtype, so "is q0 = RN(x * RN(1/scale)) followed by
 * N residual corrections equal to the IEEE quotient" can be decided by enumeration rather than by a proof.
 * Prints, per dtype and N in {0, 1, 2}: pairs whose fp32 quotient differs, and pairs whose e4m3 byte differs.
 *
 * gcc -O2 -march=native -ffp-contract=off -fopenmp -o /tmp/div_exhaustive scripts/ubench/div_exhaustive.c -lm
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static inline float bits2f(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }
static inline uint32_t f2bits(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }
static inline float bf16_to_f(uint16_t h) { return bits2f((uint32_t)h << 16); }
static float fp16_to_f(uint16_t h) {
    int e = (h >> 10) & 31, m = h & 1023;
    if (e == 0) return ldexpf((float)m, -24);
    return ldexpf((float)(m | 1024), e - 25);
}
/* e4m3fn, round to nearest even, saturating; a >= 0 finite */
static inline uint8_t e4m3_pos(float a) {
    if (a >= 448.0f) return 0x7E;
    uint32_t b = f2bits(a);
    int e = (int)(b >> 23) - 127;
    if (e < -6) { /* subnormal range: step 2^-9 */
        float q = a * 512.0f; /* exact */
        float r = nearbyintf(q);
        return (uint8_t)r; /* 0..8: 8 is 0x08 = 2^-6, the first normal */
    }
    uint32_t mant = b & 0x7FFFFF, keep = mant >> 20, rest = mant & 0xFFFFF;
    uint32_t v = ((uint32_t)(e + 7) << 3) | keep;
    if (rest > 0x80000 || (rest == 0x80000 && (keep & 1))) v += 1;
    return v > 0x7E ? 0x7E : (uint8_t)v;
}

int main(void) {
    for (int dt = 0; dt < 2; ++dt) {
        const int amax_top = dt == 0 ? 0x7F7F : 0x7BFF; /* largest finite bf16 / fp16 */
        unsigned long long pairs = 0, qdiff[3] = {0, 0, 0}, bdiff[3] = {0, 0, 0};
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : pairs, qdiff[:3], bdiff[:3])
        for (int ai = 0; ai <= amax_top; ++ai) {
            const float amax = dt == 0 ? bf16_to_f((uint16_t)ai) : fp16_to_f((uint16_t)ai);
            float s = amax * (1.0f / 448.0f);
            if (s < FLT_EPSILON) s = FLT_EPSILON;
            const float rcp = 1.0f / s; /* IEEE division: the correctly rounded reciprocal, as __frcp_rn */
            for (int xi = 0; xi <= ai; ++xi) {
                const float x = dt == 0 ? bf16_to_f((uint16_t)xi) : fp16_to_f((uint16_t)xi);
                const float ref = x / s;
                const uint8_t rb = e4m3_pos(ref);
                float q = x * rcp;
                for (int n = 0; n < 3; ++n) {
                    if (n) q = fmaf(-fmaf(s, q, -x), rcp, q);
                    if (f2bits(q) != f2bits(ref)) {
                        qdiff[n]++;
                        if (e4m3_pos(q) != rb) bdiff[n]++;
                    }
                }
                pairs++;
            }
        }
        printf("%s: %llu pairs;", dt == 0 ? "bf16" : "fp16", pairs);
        for (int n = 0; n < 3; ++n) printf("  %d corrections: fp32 quotient differs %llu, e4m3 byte differs %llu;", n, qdiff[n], bdiff[n]);
        printf("\n");
    }
    return 0;
}
