/* tools/verify_fp_tricks.c -- exhaustive CPU check of the three "no double division / no double
 * compare on the hot path" identities the CUDA kernels rely on (hackrfdiags_b200/csrc/hrd_device.cuh,
 * hrd_tx.cu).  Each trick is restated here in plain C with the same constants and compared against
 * the reference expression (the C++ the reference evaluates) over EVERY float the kernels can feed it.
 *
 *   gcc -O2 -ffp-contract=off -fopenmp -o /tmp/verify_fp_tricks tools/verify_fp_tricks.c -lm && /tmp/verify_fp_tricks
 *
 * 1. wrap:   (float)((double)x - 2*M_PI) for x in [PI_UP, 12)   vs   (x - 2PI_HI) - 2PI_LO in fp32,
 *            with the kernels' fallback rule (|x - 2PI_HI| < 2^-10 -> use the double expression)
 *            (FmDemodulator.cc:511-519, WbFmDemodulator.cc:416-424, PhaseAccumulator.cc:165-177)
 * 1c. the chains' branch-free form (hrd_device.cuh phase_step_fast): k = sat((|a| - P_DN) * 2^30) is exactly 1.0 for
 *            |a| >= PI_UP and exactly 0.0 below, and r = fma(k, -+2PI_LO, fma(k, -+2PI_HI, a)) equals the reference's
 *            wrap for PI_UP <= |a| < 6.2 and a itself otherwise -- every float a with |a| < 6.2
 * 2. step:   (float)((2*M_PI*(double)f)/fs) for fs = 256000 (WBFM NCO) and 8000 (FM NCO), every finite float f   vs   multiply by the
 *            reciprocal with the kernels' "risky -> divide" rule (PhaseAccumulator.cc:103)
 * 3. index:  (int16_t)((double)(p*16384.0f)/(2*M_PI)) for every float p in [0, PI_UP]   vs   the
 *            threshold-table search (Nco.cc:231-233): the two-sided form (estimate within +-1, two
 *            compares) and the ONE-SIDED form tx_wbfm_kernel uses (hrd_tx.cu nco_fold_offset): an FMA with a
 *            slightly low constant and the magic addend 2^23 - 0.5 lands on k or k+1, one compare settles it
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static uint64_t d2u(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }

#define PI_UP 3.14159274101257324219f
#define TWO_PI_HI 6.28318548202514648438f
#define TWO_PI_LO (-1.74845553146951715e-07f)

static long check_wrap(void)
{
    long bad = 0, slow = 0, n = 0;
    const uint32_t lo = f2u(PI_UP), hi = f2u(12.0f);
#pragma omp parallel for reduction(+ : bad, slow, n)
    for (uint32_t u = lo; u < hi; u++) {
        const float x = u2f(u);
        const float want = (float)((double)x - 2 * M_PI);
        volatile float t = x - TWO_PI_HI;
        volatile float r = t - TWO_PI_LO;
        float got = r;
        if (fabsf(t) < 0x1p-10f) { got = want; slow++; }
        if (f2u(got) != f2u(want)) bad++;
        n++;
    }
    printf("wrap : %ld floats in [PI_UP,12), %ld through the double fallback, %ld mismatches\n", n, slow, bad);
    if (u2f(f2u(TWO_PI_HI)) != (float)(2 * M_PI)) { printf("wrap : TWO_PI_HI is not fl32(2*M_PI)\n"); bad++; }
    if (TWO_PI_LO != (float)(2 * M_PI - (double)TWO_PI_HI)) { printf("wrap : TWO_PI_LO wrong\n"); bad++; }
    if (!((double)PI_UP > M_PI && (double)u2f(f2u(PI_UP) - 1) < M_PI)) { printf("wrap : PI_UP is not the float just above pi\n"); bad++; }
    return bad;
}

#define PI_DN 3.14159250259399414062f
static float satf(float x) { return x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x); }
static long check_wrap_sat(void)
{
    long bad = 0, n = 0;
    const uint32_t hi = f2u(6.2f);
    if (f2u(PI_DN) + 1 != f2u(PI_UP)) { printf("wrap2: PI_DN is not the float below PI_UP\n"); bad++; }
#pragma omp parallel for reduction(+ : bad, n)
    for (uint32_t u = 0; u < hi; u++) {
        for (int neg = 0; neg < 2; neg++) {
            const float a = neg ? -u2f(u) : u2f(u);
            /* the reference: one pass of its two while loops (|a| < 6.2: at most one wrap) */
            float want = a;
            if ((double)a > M_PI) want = (float)((double)a - 2 * M_PI);
            else if ((double)a < -M_PI) want = (float)((double)a + 2 * M_PI);
            const float k = satf(fmaf(fabsf(a), 0x1p30f, -PI_DN * 0x1p30f));
            if (k != 0.0f && k != 1.0f) bad++;
            const float hs = neg ? TWO_PI_HI : -TWO_PI_HI, ls = neg ? TWO_PI_LO : -TWO_PI_LO;
            const float got = fmaf(k, ls, fmaf(k, hs, a));
            if (f2u(got) != f2u(want) && !(got == 0.0f && want == 0.0f)) bad++;
            n++;
        }
    }
    printf("wrap2: %ld floats with |a| < 6.2, branch-free chain step against the reference's loops: %ld mismatches\n", n, bad);
    return bad;
}

static float step_fast(float f, int *slow, double fs)
{
    const double a = (2 * M_PI) * (double)f;
    double r = a * (1.0 / fs);
    const uint64_t b = d2u(r);
    const uint32_t lo = (uint32_t)b, e = (uint32_t)(b >> 52) & 0x7ffu;
    const int risky = ((lo & 0x1fffffffu) - 0x0ffffff8u) <= 16u || (e - 898u) > 250u;
    if (risky && a != 0.0) { r = a / fs; (*slow)++; }
    return (float)r;
}

static long check_step(double fs)
{
    long bad = 0, slow_total = 0, n = 0;
#pragma omp parallel for reduction(+ : bad, slow_total, n)
    for (uint64_t u = 0; u < 0x7f800000ull; u++) { /* every non-negative finite float; negatives mirror */
        const float f = u2f((uint32_t)u);
        int slow = 0;
        const float want = (float)((2 * M_PI * (double)f) / fs);
        const float got = step_fast(f, &slow, fs);
        if (f2u(got) != f2u(want)) bad++;
        const float gn = step_fast(-f, &slow, fs);
        if (f2u(gn) != f2u((float)((2 * M_PI * (double)(-f)) / fs))) bad++;
        slow_total += slow;
        n += 2;
    }
    printf("step / %.0f: %ld floats, %ld through the division fallback, %ld mismatches\n", fs, n, slow_total, bad);
    return bad;
}

static int ref_index(float p) { return (int)(int16_t)(int32_t)((double)(p * 16384.0f) / (2 * M_PI)); }

static long check_index(void)
{
    /* thresholds exactly as hrd_api.cu builds them: T[k] = smallest float p >= 0 with ref_index(p) >= k */
    static float T[8194];
    T[0] = 0.0f;
    for (int k = 1; k <= 8192; k++) {
        float p = (float)(2 * M_PI * k / 16384.0);
        while (ref_index(p) >= k) p = nextafterf(p, 0.0f);
        while (ref_index(p) < k) p = nextafterf(p, 100.0f);
        T[k] = p;
    }
    T[8193] = INFINITY;
    long bad = 0, n = 0;
    const uint32_t hi = f2u(PI_UP);
#pragma omp parallel for reduction(+ : bad, n)
    for (uint32_t u = 0; u <= hi; u++) {
        const float p = u2f(u);
        int k = (int)(p * 2607.59448f); /* 16384/(2 pi) in float: within +-1 of the answer */
        if (k > 8192) k = 8192;
        if (p < T[k]) k--;
        else if (p >= T[k + 1]) k++;
        if (k != ref_index(p) || -k != ref_index(-p)) bad++;
        n++;
    }
    printf("index: %ld floats in [0,PI_UP], %ld mismatches (T[8192] = %.9g, PI_UP = %.9g)\n", n, bad, T[8192], PI_UP);
    return bad;
}

/* hrd_tx.cu nco_fold_offset: k1 = bits(fma(|p|, C_LO, 2^23 - 0.5)) - 0x4affffff is k or k + 1 (never less, never
 * more), clamped to 8193; k = k1 - (|p| < T[k1]).  The constant must be the one in the kernel. */
#define NCO_C_LO 2607.5920f
static long check_index_onesided(void)
{
    static float T[8194];
    T[0] = 0.0f;
    for (int k = 1; k <= 8192; k++) {
        float p = (float)(2 * M_PI * k / 16384.0);
        while (ref_index(p) >= k) p = nextafterf(p, 0.0f);
        while (ref_index(p) < k) p = nextafterf(p, 100.0f);
        T[k] = p;
    }
    T[8193] = INFINITY;
    long bad = 0, n = 0, exact = 0;
    const uint32_t hi = f2u(PI_UP);
#pragma omp parallel for reduction(+ : bad, n, exact)
    for (uint32_t u = 0; u <= hi; u++) {
        const float p = u2f(u);
        uint32_t k1 = f2u(fmaf(p, NCO_C_LO, 8388607.5f)) - 0x4affffffu;
        if (k1 > 8193u) k1 = 8193u;
        const int want = ref_index(p);
        if (!((int)k1 == want || (int)k1 == want + 1)) bad++;
        exact += (int)k1 == want;
        const int k = (int)k1 - (p < T[k1] ? 1 : 0);
        if (k != want || -k != ref_index(-p)) bad++;
        n++;
    }
    printf("index (one-sided): %ld floats in [0,PI_UP], estimate already right for %ld, %ld mismatches\n", n, exact, bad);
    return bad;
}

int main(void)
{
    long bad = check_wrap() + check_wrap_sat() + check_index() + check_index_onesided() + check_step(256000.0) + check_step(8000.0);
    printf(bad ? "FAILED\n" : "all identities hold\n");
    return bad != 0;
}
