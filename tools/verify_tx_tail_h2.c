/* tools/verify_tx_tail_h2.c -- exhaustive CPU proof of the packed-half (both rails per instruction) form of Tx
 * interpolator stages 6, 7, 8 that tx_wbfm_kernel and the two-rail tx_kernel instances use (hackrfdiags_b200/csrc/hrd_tx.cu, tail3_h2).
 *
 *   gcc -O2 -ffp-contract=off -fopenmp -o /tmp/verify_tx_tail_h2 tools/verify_tx_tail_h2.c -lm && /tmp/verify_tx_tail_h2
 *
 * The reference arithmetic per rail (Interpolator_int16.cc:398-418 with the 4-tap half-bands {c,16384,c,0} of
 * AmModulator.cc:91-123, c = 8424 / 8249 / 8206): even = (16384 + c*(x + xm)) >> 15, odd = (x + 1) >> 1, three
 * stages, final (int8_t).  The GPU form evaluates both rails at once as fp16 pairs:
 *   B+ form of an integer v: the half 1536 + v (bits 0x6600 + v: the integer sits in the mantissa), B- form v - 1536;
 *   even6 = fma(s, 1053/4096, +-1536)         (8424/32768 = 1053/4096 is a half; rounding to the integer grid of
 *                                              [1024, 2048) is round-to-nearest of c*s, which never ties)
 *   odd   = fma(v + 1/2, 1/2, +-1536)         (= rint(v/2 + 1/4) = (v + 1) >> 1), and OO(v) = fma(v + 3/2, 1/4, ..)
 *   even8 = fma(s, 1/4 + 2^-12, 1536)         (c8 = 8206 only breaks the ties of s/4)
 *   even7 in fp32 per rail: fma(s, 8249/32768, 1.5*2^23 + 0x6600): exact product, one rounding, low 16 bits = B+ bits
 *   a B+ half plus a B- half is their exact integer sum; stage-8 odd outputs are integer shifts on the B-form bits.
 * fp16 operations are emulated exactly (every operand is a double, every result is rounded once to binary16, RNE).
 * Every (x, xm) in [-995, 995]^2 is checked for both halves of the packed word (cross-field effects included): the
 * WBFM modulator's NCO samples stay within +-900; tx_kernel<FM|SSB|IQ> checks its stage-5 outputs against 995 and
 * falls back to the integer form beyond (at +-1000 the packed form has 13 mismatching pairs).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static double rh(double x) /* round an exact double to binary16, ties to even */
{
    if (x == 0.0 || !isfinite(x)) return x;
    int e;
    frexp(fabs(x), &e); /* |x| = m * 2^e, m in [0.5, 1) */
    int q = e - 11;     /* quantum exponent for 11 significant bits */
    if (q < -24) q = -24;
    return ldexp(nearbyint(ldexp(x, -q)), q);
}
static double hadd(double a, double b) { return rh(a + b); }
static double hfma(double a, double b, double c) { return rh(a * b + c); }
static uint16_t hbits(double v) /* bit pattern of a binary16 NORMAL value */
{
    int e;
    double m = frexp(fabs(v), &e); /* m in [0.5,1) */
    uint16_t mant = (uint16_t)((uint32_t)ldexp(m, 11) & 0x3ffu);
    return (uint16_t)((v < 0 ? 0x8000u : 0u) | (uint16_t)((e - 1 + 15) << 10) | mant);
}
static uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

#define MG 1536.0
static const int C6 = 8424, C7 = 8249, C8 = 8206;

static void tail3_ref(int x0, int xm1, int8_t out[8])
{
    int y6m1 = (xm1 + 1) >> 1, y6[2] = {(16384 + C6 * (x0 + xm1)) >> 15, (x0 + 1) >> 1};
    int y7m1 = (y6m1 + 1) >> 1, y7[4];
    for (int k = 0; k < 2; k++) {
        y7[2 * k] = (16384 + C7 * (y6[k] + (k ? y6[k - 1] : y6m1))) >> 15;
        y7[2 * k + 1] = (y6[k] + 1) >> 1;
    }
    for (int k = 0; k < 4; k++) {
        int left = k ? y7[k - 1] : y7m1;
        out[2 * k] = (int8_t)((16384 + C8 * (y7[k] + left)) >> 15);
        out[2 * k + 1] = (int8_t)((y7[k] + 1) >> 1);
    }
}

/* one rail of the packed form; returns the B-form BITS of every stage-8 input and the even outputs' bits */
typedef struct { uint16_t e[4], p0, p2, p1, p3; } rail_t;
static rail_t tail3_h2_rail(int x, int xm)
{
    const double c6h = 1053.0 / 4096.0, c8h = 0.25 + 1.0 / 4096.0;
    const float c7f = 8249.0f / 32768.0f, m32 = 12582912.0f + 26112.0f;
    rail_t r;
    /* carried from the sample before (functions of xm alone) */
    double bm = hfma(hadd(xm, 0.5), 0.5, -MG), p3m = hfma(hadd(xm, 1.5), 0.25, -MG);
    double s6 = hadd(x, xm), a = hfma(s6, c6h, MG);
    double b = hfma(hadd(x, 0.5), 0.5, -MG);
    double s70 = hadd(a, bm), s72 = hadd(b, a);
    uint16_t p0b = (uint16_t)(f2u(fmaf((float)s70, c7f, m32)) & 0xffffu);
    uint16_t p2b = (uint16_t)(f2u(fmaf((float)s72, c7f, m32)) & 0xffffu);
    double p0 = (double)(p0b - 0x6600) + MG, p2 = (double)(p2b - 0x6600) + MG; /* the halves those bits are */
    double p1 = hfma(hadd(hadd(a, -MG), 0.5), 0.5, -MG);
    double p3 = hfma(hadd(x, 1.5), 0.25, -MG);
    r.e[0] = hbits(hfma(hadd(p0, p3m), c8h, MG));
    r.e[1] = hbits(hfma(hadd(p1, p0), c8h, MG));
    r.e[2] = hbits(hfma(hadd(p2, p1), c8h, MG));
    r.e[3] = hbits(hfma(hadd(p3, p2), c8h, MG));
    r.p0 = p0b, r.p2 = p2b, r.p1 = hbits(p1), r.p3 = hbits(p3);
    if (hbits(p0) != p0b || hbits(p2) != p2b) r.e[0] ^= 0xffff; /* the fp32 route must land on a valid B+ half */
    return r;
}

int main(int argc, char **argv)
{
    long bad = 0, n = 0;
    const int R = argc > 1 ? atoi(argv[1]) : 995; /* 995 is the largest range that passes (1000: 13 mismatches) */
    /* constants must be halves */
    if (rh(1053.0 / 4096.0) != 1053.0 / 4096.0 || rh(0.25 + 1.0 / 4096.0) != 0.25 + 1.0 / 4096.0) { printf("constants are not binary16\n"); return 1; }
#pragma omp parallel for reduction(+ : bad, n) schedule(dynamic, 16)
    for (int x = -R; x <= R; x++)
        for (int xm = -R; xm <= R; xm++) {
            int8_t wi[8], wq[8];
            /* rail I = (x, xm); rail Q = (xm, x) (any other valid pair: exercises the cross-field terms) */
            tail3_ref(x, xm, wi);
            tail3_ref(xm, x, wq);
            rail_t ri = tail3_h2_rail(x, xm), rq = tail3_h2_rail(xm, x);
            int8_t gi[8], gq[8];
            for (int k = 0; k < 4; k++) {
                gi[2 * k] = (int8_t)(ri.e[k] & 0xff); /* B+ bits 0x6600 + v: the low byte is (int8_t)v */
                gq[2 * k] = (int8_t)(rq.e[k] & 0xff);
            }
            /* odd outputs: integer shifts on the packed words {I low, Q high} */
            const uint32_t w0 = ri.p0 | (uint32_t)rq.p0 << 16, w2 = ri.p2 | (uint32_t)rq.p2 << 16;
            const uint32_t w1 = ri.p1 | (uint32_t)rq.p1 << 16, w3 = ri.p3 | (uint32_t)rq.p3 << 16;
            const uint32_t o1 = (w0 + 0x00010001u) >> 1, o5 = (w2 + 0x00010001u) >> 1; /* B+: (0x6600 + v + 1) >> 1 */
            const uint32_t o3 = (0xe801e801u - w1) >> 1, o7 = (0xe801e801u - w3) >> 1; /* B-: bits 0xe600 - v */
            gi[1] = (int8_t)(o1 & 0xff), gq[1] = (int8_t)((o1 >> 16) & 0xff);
            gi[3] = (int8_t)(o3 & 0xff), gq[3] = (int8_t)((o3 >> 16) & 0xff);
            gi[5] = (int8_t)(o5 & 0xff), gq[5] = (int8_t)((o5 >> 16) & 0xff);
            gi[7] = (int8_t)(o7 & 0xff), gq[7] = (int8_t)((o7 >> 16) & 0xff);
            if (memcmp(gi, wi, 8) || memcmp(gq, wq, 8)) bad++;
            n++;
        }
    printf("tail h2: %ld (x, xm) pairs, %ld mismatches\n", n, bad);
    printf(bad ? "FAILED\n" : "packed-half stages 6-8 equal the integer reference\n");
    return bad != 0;
}
