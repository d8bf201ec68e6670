/* tools/verify_sincosf.c -- is the restatement of glibc's sinf / cosf (sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c,
 * sincosf.h: Szabolcs Nagy's double-precision polynomial routines) that the FM transmit head uses on the GPU
 * (hrd_device.cuh glibc_sincosf) the SAME FUNCTION as the host's libm, bit for bit?
 *
 * The reference calls cos(phase) / sin(phase) on a float (Nco/Nco.cc:186-199; signals/pm.cc, fm.cc), i.e. libm's
 * cosf / sinf, whose results are not always the correctly rounded ones -- so "sin in double, rounded" differs from
 * them in the last bit now and then.  This program evaluates the restatement for EVERY float of magnitude below 8
 * (the phases are wrapped to +-pi, the prototype heads stay below 2*pi) and compares with libm:
 *     gcc -O2 -ffp-contract=off -o /tmp/verify_sincosf tools/verify_sincosf.c -lm && /tmp/verify_sincosf [stride]
 * The polynomial coefficients are the table of the libm the reference was built against (glibc 2.39, found in
 * libm.so.6 as the two 15-double records that start with {1,-1,-1,1}); x86-64 libm selects its FMA build of these
 * routines at load time on every CPU with FMA, and GCC contracts each "a + b * c" of the source there: variant 1
 * below.  Variant 0 (no contraction) is what a CPU without FMA would run.  Both are checked and reported. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    double sign[4], hpi_inv, hpi, c0, c1, c2, c3, c4, s1, s2, s3;
} sincos_t;
static const sincos_t T[2] = {
    {{1.0, -1.0, -1.0, 1.0}, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, 0x1p0, -0x1.ffffffd0c621cp-2, 0x1.55553e1068f19p-5,
     -0x1.6c087e89a359dp-10, 0x1.99343027bf8c3p-16, -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13},
    {{1.0, -1.0, -1.0, 1.0}, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, -0x1p0, 0x1.ffffffd0c621cp-2, -0x1.55553e1068f19p-5,
     0x1.6c087e89a359dp-10, -0x1.99343027bf8c3p-16, -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13}};

static inline uint32_t asuint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline uint32_t abstop12(float x) { return (asuint(x) >> 20) & 0x7ff; }
#define MADD(F, a, b, c) ((F) ? fma((a), (b), (c)) : (a) * (b) + (c))

#define DEFINE(FMA, NAME)                                                                                     \
    static inline float poly_##NAME(double x, double x2, const sincos_t *p, int n)                           \
    {                                                                                                         \
        if ((n & 1) == 0) {                                                                                   \
            const double x3 = x * x2, s1 = MADD(FMA, x2, p->s3, p->s2), x7 = x3 * x2, s = MADD(FMA, x3, p->s1, x);      \
            return (float)MADD(FMA, x7, s1, s);                                                                    \
        }                                                                                                     \
        const double x4 = x2 * x2, c2 = MADD(FMA, x2, p->c4, p->c3), c1 = MADD(FMA, x2, p->c1, p->c0), x6 = x4 * x2,    \
                     c = MADD(FMA, x4, p->c2, c1);                                                                 \
        return (float)MADD(FMA, x6, c2, c);                                                                        \
    }                                                                                                         \
    static inline float eval_##NAME(float y, int want_cos)                                                    \
    {                                                                                                         \
        double x = y;                                                                                         \
        const sincos_t *p = &T[0];                                                                            \
        if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {                                                         \
            if (abstop12(y) < abstop12(0x1p-12f)) return want_cos ? 1.0f : y;                                 \
            return poly_##NAME(x, x * x, p, want_cos);                                                        \
        }                                                                                                     \
        const double r = x * p->hpi_inv;                                                                      \
        const int n = ((int32_t)r + 0x800000) >> 24;                                                          \
        x = FMA ? fma(-(double)n, p->hpi, x) : x - n * p->hpi;                                                \
        const double s = p->sign[n & 3];                                                                      \
        if (n & 2) p = &T[1];                                                                                 \
        return poly_##NAME(x * s, x * x, p, n ^ want_cos);                                                    \
    }
DEFINE(0, plain)
DEFINE(1, fused)

int main(int argc, char **argv)
{
    unsigned long long n = 0, bad[2][2] = {{0, 0}, {0, 0}};
    const uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 1u; /* 1 = every float (22 s); the CPU test suite strides */
    for (uint32_t u = 0; u < 0x41000000u; u += stride) { /* 0 .. 8.0 */
        for (int neg = 0; neg < 2; neg++) {
            const uint32_t bits = u | (neg ? 0x80000000u : 0u);
            float y;
            memcpy(&y, &bits, 4);
            const float s = sinf(y), c = cosf(y);
            bad[0][0] += asuint(eval_plain(y, 0)) != asuint(s);
            bad[0][1] += asuint(eval_plain(y, 1)) != asuint(c);
            bad[1][0] += asuint(eval_fused(y, 0)) != asuint(s);
            bad[1][1] += asuint(eval_fused(y, 1)) != asuint(c);
            n++;
        }
    }
    printf("%llu floats with |x| < 8\n", n);
    printf("variant 0 (no contraction): sinf %llu mismatches, cosf %llu\n", bad[0][0], bad[0][1]);
    printf("variant 1 (every a + b * c fused): sinf %llu mismatches, cosf %llu\n", bad[1][0], bad[1][1]);
    return !(bad[1][0] == 0 && bad[1][1] == 0) && !(bad[0][0] == 0 && bad[0][1] == 0);
}
