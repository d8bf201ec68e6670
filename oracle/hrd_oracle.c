/*
 * hrd_oracle.c -- CPU restatement of the HackRfDiags baseband DSP hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see hrd_oracle.h).  Plain C99, single stream,
 * sample-serial, written for clarity not speed.  Build with
 *   gcc -O2 -ffp-contract=off -fwrapv
 * (no fast-math: the float sections are order- and rounding-sensitive).
 *
 * Parity pin: bit-compared with the compiled reference in
 * tests/test_oracle_vs_ref.py and with tests/golden/ (see header).
 *
 * Layout of this file
 *   1. scalar conversion helpers that pin down the reference's C++ casts
 *   2. the six DSP primitives (Q15 decimator / interpolator / FIR, float FIR,
 *      float IIR, phase accumulator + NCO)
 *   3. the Rx front end and the four demodulators
 *   4. the four modulators
 */
#define _DEFAULT_SOURCE /* M_PI */
#include "hrd_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* 1. conversions                                                      */
/* ------------------------------------------------------------------ */

/* (int16_t)someFloat as x86-64 g++ evaluates it: cvttss2si to int32
 * (0x80000000 when out of range or NaN), then the low 16 bits.  Used at
 * FmDemodulator.cc:565, WbFmDemodulator.cc:474, AmDemodulator.cc:465,
 * SsbDemodulator.cc:592, AmModulator.cc:601, FmModulator.cc:616,
 * WbFmModulator.cc:625, SsbModulator.cc:683. */
static int16_t f32_to_i16(float x)
{
    int32_t v;
    if (x >= -2147483648.0f && x < 2147483648.0f)
        v = (int32_t)x; /* truncation toward zero */
    else
        v = INT32_MIN;
    return (int16_t)(uint16_t)((uint32_t)v & 0xffffu);
}

/* (int16_t)someDouble, same rule through cvttsd2si (Nco.cc:231). */
static int16_t f64_to_i16(double x)
{
    int32_t v;
    if (x > -2147483649.0 && x < 2147483648.0)
        v = (int32_t)x;
    else
        v = INT32_MIN;
    return (int16_t)(uint16_t)((uint32_t)v & 0xffffu);
}

/* (int16_t)round(c * 32768) with c, the product and round() all in float
 * (Decimator_int16.cc:56-66, Interpolator_int16.cc:288-296,
 * FirFilter_int16.cc:44-54).  1.0f*32768 = 32768 -> 0x8000 = -32768. */
static int16_t quantise_q15(float c)
{
    float scaled = c * 32768;
    scaled = roundf(scaled);
    return f32_to_i16(scaled);
}

/* Q15 multiply-accumulate tail: accumulator starts at 1<<14, wraps like
 * int32 does on the reference's targets, ends with an arithmetic >>15 and an
 * (int16_t) narrowing (Decimator_int16.cc:192-247). */
static int16_t q15_round(uint32_t acc)
{
    int32_t s = (int32_t)acc;
    return (int16_t)(uint16_t)((uint32_t)(s >> 15) & 0xffffu);
}

/* ------------------------------------------------------------------ */
/* coefficient sources (float literals exactly as the reference spells  */
/* them; quantised at construction like the reference does)             */
/* ------------------------------------------------------------------ */
static const float k_fe1[3] = {0.2504357f, 0.5000000f, 0.2504357f};
static const float k_fe2[3] = {0.2517491f, 0.4999998f, 0.2517491f};
static const float k_fe3[3] = {0.2570951f, 0.5000000f, 0.2570951f};
static const float k_am1[8] = {0.0242683f, 0.0766338f, 0.1457589f, 0.1959036f,
                               0.1959036f, 0.1457589f, 0.0766338f, 0.0242683f};
static const float k_am2[12] = {0.0057496f, 0.0263853f, 0.0605301f, 0.1074406f,
                                0.1523486f, 0.1804951f, 0.1804951f, 0.1523486f,
                                0.1074406f, 0.0605301f, 0.0263853f, 0.0057496f};
static const float k_am3[16] = {0.0116487f,  0.0152694f,  -0.0109804f, -0.0611915f,
                                -0.0736143f, 0.0187617f,  0.1988190f,  0.3481364f,
                                0.3481364f,  0.1988190f,  0.0187617f,  -0.0736143f,
                                -0.0611915f, -0.0109804f, 0.0152694f,  0.0116487f};
static const float k_fm_tuner[32] = {
    0.0041331f, 0.0054174f, 0.0076016f, 0.0115481f, 0.0151685f, 0.0203192f, 0.0251608f,
    0.0311322f, 0.0366372f, 0.0427168f, 0.0480527f, 0.0533425f, 0.0575831f, 0.0611914f,
    0.0635413f, 0.0648239f, 0.0648239f, 0.0635413f, 0.0611914f, 0.0575831f, 0.0533425f,
    0.0480527f, 0.0427168f, 0.0366372f, 0.0311322f, 0.0251608f, 0.0203192f, 0.0151685f,
    0.0115481f, 0.0076016f, 0.0054174f, 0.0041331f};
static const float k_fm_post[12] = {0.0022977f, 0.0237042f, 0.0605386f, 0.1127073f,
                                    0.1645167f, 0.1971107f, 0.1971107f, 0.1645167f,
                                    0.1127073f, 0.0605386f, 0.0237042f, 0.0022977f};
static const float k_audio40[40] = {
    0.0015969f,  -0.0111080f, -0.0270501f, -0.0265610f, -0.0023190f, 0.0180618f,  0.0065495f,
    -0.0183409f, -0.0133345f, 0.0184489f,  0.0230891f,  -0.0161248f, -0.0363745f, 0.0091343f,
    0.0550219f,  0.0070312f,  -0.0862280f, -0.0497761f, 0.1793543f,  0.4145808f,  0.4145808f,
    0.1793543f,  -0.0497761f, -0.0862280f, 0.0070312f,  0.0550219f,  0.0091343f,  -0.0363745f,
    -0.0161248f, 0.0230891f,  0.0184489f,  -0.0133345f, -0.0183409f, 0.0065495f,  0.0180618f,
    -0.0023190f, -0.0265610f, -0.0270501f, -0.0111080f, 0.0015969f};
/* signals/interpolateSignal.cc:30-72: the stand-alone interpolator's OWN stage-1 prototype (stages 2..8 carry the
 * same numbers as the modulator classes, interpolateSignal.cc:74-140 vs AmModulator.cc:57-123) */
static const float k_sig40[40] = {
    -0.0011405f, 0.0183372f,  0.0030542f,  -0.0100052f, -0.0059350f, 0.0115377f,  0.0109293f,
    -0.0120883f, -0.0175779f, 0.0110390f,  0.0262645f,  -0.0074772f, -0.0377408f, -0.0003152f,
    0.0541009f,  0.0165897f,  -0.0829085f, 0.0587608f,  0.1736804f,  0.4222137f,  0.4222137f,
    0.1736804f,  -0.0587608f, -0.0829085f, 0.0165897f,  0.0541009f,  -0.0003152f, -0.0377408f,
    -0.0074772f, 0.0262645f,  0.0110390f,  -0.0175779f, -0.0120883f, 0.0109293f,  0.0115377f,
    -0.0059350f, -0.0100052f, 0.0030542f,  0.0183372f,  -0.0011405f};
static const float k_wbfm_post1[8] = {0.0243699f, 0.0769537f, 0.1463572f, 0.1967096f,
                                      0.1967096f, 0.1463572f, 0.0769537f, 0.0243699f};
static const float k_delay16[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1};
static const float k_hilbert31[31] = {
    -0.0033953f, 0, -0.0058652f, 0, -0.0134385f, 0, -0.0281423f, 0, -0.0534836f, 0, -0.0980394f,
    0, -0.1935638f, 0, -0.6302204f, 0, 0.6302204f, 0, 0.1935638f, 0, 0.0980394f, 0, 0.0534836f,
    0, 0.0281423f, 0, 0.0134385f, 0, 0.0058652f, 0, 0.0033953f};
static const float k_tx_hb8[8] = {-0.0440934f, 0, 0.2913764f, 0.5000000f,
                                  0.2913764f,  0, -0.0440934f, 0};
/* Tx stages 3,6 / 7 / 8 (AmModulator.cc:69-123): the Rx half-band triplets
 * with a trailing zero so that N is a multiple of L=2. */
static const float k_tx_hb4c[4] = {0.2570951f, 0.5000000f, 0.2570951f, 0};
static const float k_tx_hb4b[4] = {0.2517491f, 0.4999998f, 0.2517491f, 0};
static const float k_tx_hb4a[4] = {0.2504357f, 0.5000000f, 0.2504357f, 0};

/* ------------------------------------------------------------------ */
/* 2. primitives                                                       */
/* ------------------------------------------------------------------ */

#define MAXTAPS 40

/* Decimator_int16 (Decimator_int16.cc:321-362 decimate, :176-249
 * filterData, :269-285 shiftSampleIn).  The reference buffers M samples,
 * shifts M-1 of them into a ring, then convolves while inserting the last;
 * that is the same as keeping the last N inputs, newest first, and
 * convolving after every M-th arrival. */
typedef struct {
    int n, m, fill;
    int16_t q[MAXTAPS];
    int16_t x[MAXTAPS]; /* x[0] newest */
} dec16;

static void dec16_reset(dec16 *d)
{
    d->fill = 0;
    memset(d->x, 0, sizeof d->x);
}

static void dec16_init(dec16 *d, const float *c, int n, int m)
{
    d->n = n;
    d->m = m;
    for (int i = 0; i < n; i++) d->q[i] = quantise_q15(c[i]);
    dec16_reset(d);
}

static int dec16_push(dec16 *d, int16_t in, int16_t *out)
{
    memmove(d->x + 1, d->x, (size_t)(d->n - 1) * sizeof(int16_t));
    d->x[0] = in;
    if (++d->fill < d->m) return 0;
    d->fill = 0;
    uint32_t acc = 1u << 14;
    for (int k = 0; k < d->n; k++) acc += (uint32_t)((int32_t)d->q[k] * (int32_t)d->x[k]);
    *out = q15_round(acc);
    return 1;
}

/* Interpolator_int16 (Interpolator_int16.cc:398-418 interpolate, :149-211
 * filterData, :267-333 polyphase split): branch i uses taps q[i + k*L]. */
typedef struct {
    int n, l, plen;
    int16_t q[MAXTAPS];
    int16_t x[MAXTAPS]; /* x[0] newest, plen valid */
} int16i;

static void int16i_reset(int16i *p) { memset(p->x, 0, sizeof p->x); }

static void int16i_init(int16i *p, const float *c, int n, int l)
{
    p->n = n;
    p->l = l;
    p->plen = n / l;
    for (int i = 0; i < n; i++) p->q[i] = quantise_q15(c[i]);
    int16i_reset(p);
}

static void int16i_push(int16i *p, int16_t in, int16_t *out /* [l] */)
{
    memmove(p->x + 1, p->x, (size_t)(p->plen - 1) * sizeof(int16_t));
    p->x[0] = in;
    for (int i = 0; i < p->l; i++) {
        uint32_t acc = 1u << 14;
        for (int k = 0; k < p->plen; k++)
            acc += (uint32_t)((int32_t)p->q[i + k * p->l] * (int32_t)p->x[k]);
        out[i] = q15_round(acc);
    }
}

/* FirFilter_int16 (FirFilter_int16.cc:151-224) */
typedef struct {
    int n;
    int16_t q[MAXTAPS];
    int16_t x[MAXTAPS];
} fir16;

static void fir16_reset(fir16 *f) { memset(f->x, 0, sizeof f->x); }

static void fir16_init(fir16 *f, const float *c, int n)
{
    f->n = n;
    for (int i = 0; i < n; i++) f->q[i] = quantise_q15(c[i]);
    fir16_reset(f);
}

static int16_t fir16_push(fir16 *f, int16_t in)
{
    memmove(f->x + 1, f->x, (size_t)(f->n - 1) * sizeof(int16_t));
    f->x[0] = in;
    uint32_t acc = 1u << 14;
    for (int k = 0; k < f->n; k++) acc += (uint32_t)((int32_t)f->q[k] * (int32_t)f->x[k]);
    return q15_round(acc);
}

/* FirFilter (Filters/FirFilter.cc:144-185): y = 0; y = y + h[k]*x[n-k] in
 * that order, every operation rounded to float, no fused multiply-add. */
typedef struct {
    int n;
    float h[8];
    float x[8];
} firf;

static void firf_reset(firf *f) { memset(f->x, 0, sizeof f->x); }

static void firf_init(firf *f, const float *h, int n)
{
    f->n = n;
    memcpy(f->h, h, (size_t)n * sizeof(float));
    firf_reset(f);
}

static float firf_push(firf *f, float in)
{
    memmove(f->x + 1, f->x, (size_t)(f->n - 1) * sizeof(float));
    f->x[0] = in;
    volatile float y = 0; /* volatile: keep every partial sum a rounded float */
    for (int k = 0; k < f->n; k++) {
        volatile float p = f->h[k] * f->x[k];
        y = y + p;
    }
    return y;
}

/* IirFilter (Filters/IirFilter.cc:161-176, 199-229, 250-266) for the only
 * shape the hot path uses: a denominator of length 1.
 *   y = fir(x);  y -= (0 + a0 * yprev);  yprev = y                    */
typedef struct {
    firf num;
    float a0, yprev;
} iirf;

static void iirf_reset(iirf *f)
{
    firf_reset(&f->num);
    f->yprev = 0;
}

static void iirf_init(iirf *f, const float *b, int nb, float a0)
{
    firf_init(&f->num, b, nb);
    f->a0 = a0;
    f->yprev = 0;
}

static float iirf_push(iirf *f, float in)
{
    volatile float y = firf_push(&f->num, in);
    volatile float r = 0;
    volatile float p = f->a0 * f->yprev;
    r = r + p;
    y = y - r;
    f->yprev = y;
    return y;
}

/* PhaseAccumulator (Nco/PhaseAccumulator.cc:95-107 setFrequency, :157-181
 * run) and Nco (Nco/Nco.cc:33-72 tables, :186-199 run, :222-257 runFast). */
typedef struct {
    float fs, step, acc;
} phase_acc;

static void phase_set_frequency(phase_acc *p, float f)
{
    p->step = (float)((2 * M_PI * (double)f) / (double)p->fs);
}

static float phase_run(phase_acc *p)
{
    float phase = p->acc;
    volatile float a = p->acc + p->step;
    while ((double)a > M_PI) a = (float)((double)a - (2 * M_PI));
    while ((double)a < (-M_PI)) a = (float)((double)a + (2 * M_PI));
    p->acc = a;
    return phase;
}

static float g_sin[16384], g_cos[16384];
static float g_atan2[256][256];
static int g_tables_ready;

static void tables_init(void)
{
    if (g_tables_ready) return;
    /* Nco.cc:45-61: the angle is accumulated in float; sin()/cos() on a
     * float argument resolve to the float overloads (sinf/cosf) because the
     * reference is C++ and includes <math.h>. */
    float inc = (float)(2 * M_PI / 16384);
    volatile float ang = (float)(-M_PI);
    for (int i = 0; i < 16384; i++) {
        g_sin[i] = sinf(ang);
        g_cos[i] = cosf(ang);
        ang = ang + inc;
    }
    /* FmDemodulator.cc:158-170 / WbFmDemodulator.cc:136-148: double atan2
     * narrowed to float, table[y][x] with y = q+128, x = i+128. */
    for (int x = 0; x < 256; x++)
        for (int y = 0; y < 256; y++)
            g_atan2[y][x] = (float)atan2((double)y - 128, (double)x - 128);
    g_tables_ready = 1;
}

/* wrap a float phase difference into [-pi, pi] with double compares and a
 * double subtraction narrowed to float (FmDemodulator.cc:511-519). */
static float wrap_pi(float d)
{
    volatile float v = d;
    while ((double)v > M_PI) v = (float)((double)v - (2 * M_PI));
    while ((double)v < (-M_PI)) v = (float)((double)v + (2 * M_PI));
    return v;
}

/* ------------------------------------------------------------------ */
/* 3. receive                                                          */
/* ------------------------------------------------------------------ */
struct hro_rx {
    int mode;
    int ssb_lsb; /* SsbDemodulator::lsbDemodulationMode */
    /* IqDataProcessor front end: [rail][stage] */
    dec16 fe[2][3];
    /* AmDemodulator */
    dec16 am_dec[2][3];
    iirf am_dc;
    float am_gain;
    /* FmDemodulator */
    dec16 fm_tuner[2], fm_post, fm_audio;
    firf fm_diff;
    float fm_gain;
    /* WbFmDemodulator */
    dec16 wb_post1, wb_post2, wb_audio;
    iirf wb_deemph;
    float wb_prev_theta, wb_gain;
    /* SsbDemodulator */
    dec16 ssb_dec[2][3];
    fir16 ssb_delay, ssb_hilbert;
    iirf ssb_dc;
    float ssb_gain;
    /* Squelch (Squelch.cc, SignalDetector.cc, SignalTracker.cc, DbfsCalculator.cc) */
    int32_t sq_threshold;     /* IqDataProcessor::signalDetectThreshold, dBFS */
    uint32_t sq_gain_db;      /* radio_adjustableReceiveGainInDb (Radio.cc:413) */
    int sq_tracking;          /* SignalTracker::state == Tracking */
    uint32_t sq_magnitude;    /* SignalDetector::signalMagnitude of the latest block */
    int sq_allowed;           /* what Squelch::run returned for the latest block */
    /* scratch for the 256 kS/s stream */
    int8_t *scratch;
    size_t scratch_cap;
};

hro_rx *hro_rx_new(void)
{
    tables_init();
    hro_rx *rx = (hro_rx *)calloc(1, sizeof *rx);
    if (!rx) return NULL;
    static const float dc_b[2] = {1, -1};
    static const float de_b[2] = {0.0253863f, 0.0253863f};
    /* FmDemodulator.cc:116-125: "-1/16" and "1/16" are integer divisions */
    static const float diff[7] = {-1 / 16, 0, 1, 0, -1, 0, 1 / 16};
    rx->mode = HRO_NONE;   /* IqDataProcessor.cc:70 */
    rx->sq_threshold = -200; /* IqDataProcessor.cc:120-124 */
    rx->sq_gain_db = 16;     /* Radio.cc:413 */
    rx->sq_tracking = 0;     /* SignalTracker.cc: state = NoSignal */
    rx->sq_allowed = 1;
    rx->ssb_lsb = 1;       /* SsbDemodulator.cc:143 */
    for (int r = 0; r < 2; r++) {
        dec16_init(&rx->fe[r][0], k_fe1, 3, 2);
        dec16_init(&rx->fe[r][1], k_fe2, 3, 2);
        dec16_init(&rx->fe[r][2], k_fe3, 3, 2);
        dec16_init(&rx->am_dec[r][0], k_am1, 8, 4);
        dec16_init(&rx->am_dec[r][1], k_am2, 12, 4);
        dec16_init(&rx->am_dec[r][2], k_am3, 16, 2);
        dec16_init(&rx->ssb_dec[r][0], k_am1, 8, 4);
        dec16_init(&rx->ssb_dec[r][1], k_am2, 12, 4);
        dec16_init(&rx->ssb_dec[r][2], k_am3, 16, 2);
        dec16_init(&rx->fm_tuner[r], k_fm_tuner, 32, 4);
    }
    iirf_init(&rx->am_dc, dc_b, 2, -0.95f);
    rx->am_gain = 300;                              /* AmDemodulator.cc:102 */
    dec16_init(&rx->fm_post, k_fm_post, 12, 4);
    dec16_init(&rx->fm_audio, k_audio40, 40, 2);
    firf_init(&rx->fm_diff, diff, 7);
    rx->fm_gain = (float)(64000 / (2 * M_PI));      /* FmDemodulator.cc:173 */
    dec16_init(&rx->wb_post1, k_wbfm_post1, 8, 4);
    dec16_init(&rx->wb_post2, k_fm_post, 12, 4);
    dec16_init(&rx->wb_audio, k_audio40, 40, 2);
    iirf_init(&rx->wb_deemph, de_b, 2, -0.9492274f);
    rx->wb_prev_theta = 0;
    rx->wb_gain = (float)(256000 / (2 * M_PI));     /* WbFmDemodulator.cc:151 */
    fir16_init(&rx->ssb_delay, k_delay16, 16);
    fir16_init(&rx->ssb_hilbert, k_hilbert31, 31);
    iirf_init(&rx->ssb_dc, dc_b, 2, -0.95f);
    rx->ssb_gain = 300;                             /* SsbDemodulator.cc:146 */
    return rx;
}

void hro_rx_free(hro_rx *rx)
{
    if (!rx) return;
    free(rx->scratch);
    free(rx);
}

void hro_rx_set_mode(hro_rx *rx, int mode)
{
    rx->mode = mode;
    if (mode == HRO_LSB) rx->ssb_lsb = 1;
    if (mode == HRO_USB) rx->ssb_lsb = 0;
}

void hro_rx_set_gain(hro_rx *rx, int demod, float gain)
{
    switch (demod) {
    case HRO_DEMOD_AM: rx->am_gain = gain; break;
    case HRO_DEMOD_FM: rx->fm_gain = gain; break;
    case HRO_DEMOD_WBFM: rx->wb_gain = gain; break;
    case HRO_DEMOD_SSB: rx->ssb_gain = gain; break;
    }
}

void hro_rx_reset_demod(hro_rx *rx, int demod)
{
    switch (demod) {
    case HRO_DEMOD_AM: /* AmDemodulator.cc:249-263 */
        for (int r = 0; r < 2; r++)
            for (int s = 0; s < 3; s++) dec16_reset(&rx->am_dec[r][s]);
        iirf_reset(&rx->am_dc);
        break;
    case HRO_DEMOD_FM: /* FmDemodulator.cc:296-308 */
        dec16_reset(&rx->fm_tuner[0]);
        dec16_reset(&rx->fm_tuner[1]);
        dec16_reset(&rx->fm_post);
        dec16_reset(&rx->fm_audio);
        firf_reset(&rx->fm_diff);
        break;
    case HRO_DEMOD_WBFM: /* WbFmDemodulator.cc:265-297: the de-emphasis filter is NOT reset */
        dec16_reset(&rx->wb_post1);
        dec16_reset(&rx->wb_post2);
        dec16_reset(&rx->wb_audio);
        rx->wb_prev_theta = 0;
        break;
    case HRO_DEMOD_SSB: /* SsbDemodulator.cc:297-313 */
        for (int r = 0; r < 2; r++)
            for (int s = 0; s < 3; s++) dec16_reset(&rx->ssb_dec[r][s]);
        fir16_reset(&rx->ssb_delay);
        fir16_reset(&rx->ssb_hilbert);
        iirf_reset(&rx->ssb_dc);
        break;
    }
}

/* IqDataProcessor::reduceSampleRate (IqDataProcessor.cc:429-500) then
 * upconvertByFsOver4 (:771-815).  The (int8_t) narrowing wraps; the rotation
 * restarts at phase 0 on every call and negates in int8 (-(-128) = -128). */
size_t hro_rx_front_end(hro_rx *rx, const int8_t *iq, size_t nbytes, int8_t *out)
{
    size_t nout[2] = {0, 0};
    for (int r = 0; r < 2; r++) {
        for (size_t i = (size_t)r; i < nbytes; i += 2) {
            int16_t s;
            if (!dec16_push(&rx->fe[r][0], (int16_t)iq[i], &s)) continue;
            if (!dec16_push(&rx->fe[r][1], s, &s)) continue;
            if (!dec16_push(&rx->fe[r][2], s, &s)) continue;
            out[2 * nout[r] + (size_t)r] = (int8_t)(uint8_t)((uint16_t)s & 0xff);
            nout[r]++;
        }
    }
    size_t bytes = 2 * nout[0]; /* byteCount is taken from the I rail (:465) */
    /* :782-812 works on groups of 8 bytes and, like the reference, reads and
     * writes the whole group even when byteCount is not a multiple of 8; the
     * caller's buffer is sized for that (decimatedData[32768]). */
    for (size_t i = 0; i < bytes; i += 8) {
        int8_t x, y;
        x = out[i + 2]; y = out[i + 3];
        out[i + 2] = (int8_t)(uint8_t)(0u - (uint8_t)y); out[i + 3] = x;
        x = out[i + 4]; y = out[i + 5];
        out[i + 4] = (int8_t)(uint8_t)(0u - (uint8_t)x);
        out[i + 5] = (int8_t)(uint8_t)(0u - (uint8_t)y);
        x = out[i + 6]; y = out[i + 7];
        out[i + 6] = y; out[i + 7] = (int8_t)(uint8_t)(0u - (uint8_t)x);
    }
    return bytes;
}

/* AmDemodulator::acceptIqData (AmDemodulator.cc:297-315): reduceSampleRate
 * :339-408, demodulateSignal :434-471, createPcmData :492-504. */
static size_t am_accept(hro_rx *rx, const int8_t *iq, size_t nbytes, int16_t *pcm)
{
    int16_t *rail[2];
    size_t cnt[2] = {0, 0};
    size_t cap = nbytes / 64 + 2;
    rail[0] = (int16_t *)malloc(cap * sizeof(int16_t));
    rail[1] = (int16_t *)malloc(cap * sizeof(int16_t));
    for (int r = 0; r < 2; r++)
        for (size_t i = (size_t)r; i < nbytes; i += 2) {
            int16_t s;
            if (!dec16_push(&rx->am_dec[r][0], (int16_t)iq[i], &s)) continue;
            if (!dec16_push(&rx->am_dec[r][1], s, &s)) continue;
            if (!dec16_push(&rx->am_dec[r][2], s, &s)) continue;
            rail[r][cnt[r]++] = s;
        }
    size_t n = cnt[1]; /* the Q-rail count is what reduceSampleRate returns */
    for (size_t i = 0; i < n; i++) {
        /* abs() is int abs narrowed back to int16_t (:444-445) */
        int16_t im = (int16_t)(uint16_t)((uint32_t)abs((int)rail[0][i]) & 0xffff);
        int16_t qm = (int16_t)(uint16_t)((uint32_t)abs((int)rail[1][i]) & 0xffff);
        int16_t mag;
        if (im > qm)
            mag = (int16_t)(uint16_t)((uint32_t)((int)im + (qm >> 1)) & 0xffff);
        else
            mag = (int16_t)(uint16_t)((uint32_t)((int)qm + (im >> 1)) & 0xffff);
        float y = iirf_push(&rx->am_dc, (float)mag);
        volatile float g = rx->am_gain * y;
        pcm[i] = f32_to_i16(g);
    }
    free(rail[0]);
    free(rail[1]);
    return n;
}

/* FmDemodulator::acceptIqData (FmDemodulator.cc:353-371): reduceSampleRate
 * :395-442, demodulateSignal :479-529, createPcmData :551-585. */
static size_t fm_accept(hro_rx *rx, const int8_t *iq, size_t nbytes, int16_t *pcm)
{
    size_t cap = nbytes / 8 + 2;
    int16_t *rail[2];
    size_t cnt[2] = {0, 0};
    rail[0] = (int16_t *)malloc(cap * sizeof(int16_t));
    rail[1] = (int16_t *)malloc(cap * sizeof(int16_t));
    for (int r = 0; r < 2; r++)
        for (size_t i = (size_t)r; i < nbytes; i += 2) {
            int16_t s;
            if (dec16_push(&rx->fm_tuner[r], (int16_t)iq[i], &s)) rail[r][cnt[r]++] = s;
        }
    size_t n = cnt[1];
    volatile float scale = rx->fm_gain / 15000;
    scale = scale * 32767;
    size_t npcm = 0;
    for (size_t i = 0; i < n; i++) {
        /* uint8_t idx = (uint8_t)value + 128 : low byte, then +128 mod 256 */
        uint8_t ii = (uint8_t)((uint8_t)rail[0][i] + 128);
        uint8_t qi = (uint8_t)((uint8_t)rail[1][i] + 128);
        float theta = g_atan2[qi][ii];
        float d = wrap_pi(firf_push(&rx->fm_diff, theta));
        volatile float v = scale * d;
        int16_t s;
        if (!dec16_push(&rx->fm_post, f32_to_i16(v), &s)) continue;
        if (!dec16_push(&rx->fm_audio, s, &s)) continue;
        pcm[npcm++] = s;
    }
    free(rail[0]);
    free(rail[1]);
    return npcm;
}

/* WbFmDemodulator::acceptIqData (WbFmDemodulator.cc:341-356):
 * demodulateSignal :381-439, createPcmData :460-500. */
static size_t wbfm_accept(hro_rx *rx, const int8_t *iq, size_t nbytes, int16_t *pcm)
{
    volatile float scale = rx->wb_gain / 75000;
    scale = scale * 32767;
    size_t n = nbytes / 2, npcm = 0;
    for (size_t i = 0; i < n; i++) {
        uint8_t ii = (uint8_t)((uint8_t)iq[2 * i] + 128);
        uint8_t qi = (uint8_t)((uint8_t)iq[2 * i + 1] + 128);
        float theta = g_atan2[qi][ii];
        volatile float d0 = theta - rx->wb_prev_theta;
        float d = wrap_pi(d0);
        volatile float v = scale * d;
        float y = iirf_push(&rx->wb_deemph, v);
        rx->wb_prev_theta = theta;
        int16_t s;
        if (!dec16_push(&rx->wb_post1, f32_to_i16(y), &s)) continue;
        if (!dec16_push(&rx->wb_post2, s, &s)) continue;
        if (!dec16_push(&rx->wb_audio, s, &s)) continue;
        pcm[npcm++] = s;
    }
    return npcm;
}

/* SsbDemodulator::acceptIqData (SsbDemodulator.cc:420-438):
 * reduceSampleRate :462-529, demodulateSignal :563-598. */
static size_t ssb_accept(hro_rx *rx, const int8_t *iq, size_t nbytes, int16_t *pcm)
{
    int16_t *rail[2];
    size_t cnt[2] = {0, 0};
    size_t cap = nbytes / 64 + 2;
    rail[0] = (int16_t *)malloc(cap * sizeof(int16_t));
    rail[1] = (int16_t *)malloc(cap * sizeof(int16_t));
    for (int r = 0; r < 2; r++)
        for (size_t i = (size_t)r; i < nbytes; i += 2) {
            int16_t s;
            if (!dec16_push(&rx->ssb_dec[r][0], (int16_t)iq[i], &s)) continue;
            if (!dec16_push(&rx->ssb_dec[r][1], s, &s)) continue;
            if (!dec16_push(&rx->ssb_dec[r][2], s, &s)) continue;
            rail[r][cnt[r]++] = s;
        }
    size_t n = cnt[1];
    for (size_t i = 0; i < n; i++) {
        int16_t id = fir16_push(&rx->ssb_delay, rail[0][i]);
        int16_t qh = fir16_push(&rx->ssb_hilbert, rail[1][i]);
        float v = rx->ssb_lsb ? (float)((int)id - (int)qh) : (float)((int)id + (int)qh);
        float y = iirf_push(&rx->ssb_dc, v);
        volatile float g = rx->ssb_gain * y;
        pcm[i] = f32_to_i16(g);
    }
    free(rail[0]);
    free(rail[1]);
    return n;
}

size_t hro_rx_accept_256k(hro_rx *rx, const int8_t *iq, size_t nbytes, int16_t *pcm)
{
    switch (rx->mode) {
    case HRO_AM: return am_accept(rx, iq, nbytes, pcm);
    case HRO_FM: return fm_accept(rx, iq, nbytes, pcm);
    case HRO_WBFM: return wbfm_accept(rx, iq, nbytes, pcm);
    case HRO_LSB:
    case HRO_USB: return ssb_accept(rx, iq, nbytes, pcm);
    default: return 0;
    }
}

/* ---- squelch ------------------------------------------------------- */

/* DbfsCalculator::DbfsCalculator(7) + convertMagnitudeToDbFs (DbfsCalculator.cc:36-68, 111-147):
 * fullScaleValue = 127, fullScaleValueInDb = (uint32_t)(20*log10(127.0)) = 42,
 * dbTable[i] = (int32_t)(20 * log10((float)i)) for i = 1..256, dbTable[0] = dbTable[1]. */
static int32_t dbfs_of_magnitude(uint32_t magnitude)
{
    static int32_t table[257];
    static int ready = 0;
    if (!ready) {
        for (int i = 1; i <= 256; i++) {
            float level = 20 * log10f((float)i); /* C++ log10(float) is the float overload */
            table[i] = (int32_t)level;
        }
        table[0] = table[1];
        ready = 1;
    }
    int32_t decibels = 0;
    if (magnitude > 127) magnitude = 127; /* clip to fullScaleValue (:121-125) */
    while (magnitude > 256) {             /* never taken after the clip; kept as in :127-134 */
        magnitude /= 2;
        decibels += 6;
    }
    return table[magnitude] + decibels - (int32_t)(uint32_t)(20 * log10((double)127));
}

/* SignalDetector::detectSignal (SignalDetector.cc:205-273) on nbytes of 256 kS/s int8 I,Q,
 * SignalTracker::run (SignalTracker.cc:104-145) and Squelch::run (Squelch.cc:227-273) */
static int squelch_run(hro_rx *rx, const int8_t *iq256, size_t nbytes)
{
    uint32_t magnitude = 0;
    const uint32_t count = (uint32_t)(nbytes / 2);
    for (size_t i = 0; i < nbytes; i += 2) {
        uint8_t im = (uint8_t)abs(iq256[i]), qm = (uint8_t)abs(iq256[i + 1]);
        uint8_t m = im > qm ? (uint8_t)(im + (qm >> 1)) : (uint8_t)(qm + (im >> 1));
        magnitude += m;
    }
    magnitude /= count;
    int32_t dbfs = dbfs_of_magnitude(magnitude);
    dbfs -= rx->sq_gain_db; /* int32 -= uint32, as the reference writes it */
    const int present = dbfs >= rx->sq_threshold;
    rx->sq_magnitude = magnitude;
    /* NoSignal: present -> Tracking/START (allowed), else NOISE (not allowed);
     * Tracking: present -> SIGNALPRESENT (allowed), else -> NoSignal/ENDOFSIGNAL (allowed: the tail) */
    const int allowed = rx->sq_tracking ? 1 : present;
    rx->sq_tracking = present;
    rx->sq_allowed = allowed;
    return allowed;
}

void hro_rx_set_squelch_threshold(hro_rx *rx, int32_t threshold_dbfs) { rx->sq_threshold = threshold_dbfs; }
void hro_rx_set_rx_gain_db(hro_rx *rx, uint32_t gain_db) { rx->sq_gain_db = gain_db; }
uint32_t hro_rx_signal_magnitude(const hro_rx *rx) { return rx->sq_magnitude; }
int hro_rx_signal_allowed(const hro_rx *rx) { return rx->sq_allowed; }

/* IqDataProcessor::acceptIqData (IqDataProcessor.cc:926-1038): front end, squelch, gated demodulator */
size_t hro_rx_accept_2048k(hro_rx *rx, const int8_t *iq, size_t nbytes, int16_t *pcm)
{
    size_t need = nbytes / 8 + 16;
    if (rx->scratch_cap < need) {
        free(rx->scratch);
        rx->scratch = (int8_t *)calloc(need, 1);
        rx->scratch_cap = need;
    }
    size_t nb = hro_rx_front_end(rx, iq, nbytes, rx->scratch);
    if (nb == 0) return 0;
    if (!squelch_run(rx, rx->scratch, nb)) return 0; /* :991: the demodulator is not called */
    return hro_rx_accept_256k(rx, rx->scratch, nb, pcm);
}

/* ------------------------------------------------------------------ */
/* 4. transmit                                                         */
/* ------------------------------------------------------------------ */

/* the eight-stage x256 tree of one rail (AmModulator.cc:410-530) */
typedef struct {
    int16i st[8];
} tx_rail;

static void tx_rail_init(tx_rail *t)
{
    int16i_init(&t->st[0], k_audio40, 40, 2);
    int16i_init(&t->st[1], k_tx_hb8, 8, 2);
    int16i_init(&t->st[2], k_tx_hb4c, 4, 2);
    int16i_init(&t->st[3], k_tx_hb8, 8, 2);
    int16i_init(&t->st[4], k_tx_hb8, 8, 2);
    int16i_init(&t->st[5], k_tx_hb4c, 4, 2);
    int16i_init(&t->st[6], k_tx_hb4b, 4, 2);
    int16i_init(&t->st[7], k_tx_hb4a, 4, 2);
}

/* run stages [first,last] on one input sample; out gets 2^(last-first+1) */
static void tx_rail_run(tx_rail *t, int first, int last, int16_t in, int16_t *out)
{
    int16_t a[256], b[256];
    int16_t *src = a, *dst = b;
    int n = 1;
    a[0] = in;
    for (int s = first; s <= last; s++) {
        for (int i = 0; i < n; i++) int16i_push(&t->st[s], src[i], dst + 2 * i);
        n *= 2;
        int16_t *tmp = src; src = dst; dst = tmp;
    }
    memcpy(out, src, (size_t)n * sizeof(int16_t));
}

struct hro_tx {
    /* AmModulator */
    tx_rail am[2];
    float am_index;
    /* FmModulator */
    tx_rail fm[2];
    float fm_dev;
    phase_acc fm_phase;
    /* WbFmModulator: stages 1-5 on the real PCM, 6-8 on I and Q */
    tx_rail wb_pcm, wb_iq[2];
    float wb_dev;
    phase_acc wb_phase;
    /* SsbModulator */
    tx_rail ssb[2];
    fir16 ssb_delay, ssb_hilbert;
    /* signals/interpolateSignal.cc: the stand-alone I and Q interpolator trees */
    tx_rail sig[2];
    float sig_theta; /* signals/fm.cc: theta */
};

hro_tx *hro_tx_new(void)
{
    tables_init();
    hro_tx *tx = (hro_tx *)calloc(1, sizeof *tx);
    if (!tx) return NULL;
    for (int r = 0; r < 2; r++) {
        tx_rail_init(&tx->am[r]);
        tx_rail_init(&tx->fm[r]);
        tx_rail_init(&tx->wb_iq[r]);
        tx_rail_init(&tx->ssb[r]);
    }
    tx_rail_init(&tx->wb_pcm);
    for (int r = 0; r < 2; r++) {
        tx_rail_init(&tx->sig[r]);
        int16i_init(&tx->sig[r].st[0], k_sig40, 40, 2); /* interpolateSignal.cc:30-72, 206-208 */
    }
    tx->am_index = 0.8f;        /* AmModulator.cc:218 */
    tx->fm_dev = 3500;          /* FmModulator.cc:218 */
    tx->fm_phase.fs = 8000;     /* FmModulator.cc:221 */
    phase_set_frequency(&tx->fm_phase, 0);
    tx->wb_dev = 70000;         /* WbFmModulator.cc:204 */
    tx->wb_phase.fs = 256000;   /* WbFmModulator.cc:207 */
    phase_set_frequency(&tx->wb_phase, 0);
    fir16_init(&tx->ssb_delay, k_delay16, 16);
    fir16_init(&tx->ssb_hilbert, k_hilbert31, 31);
    return tx;
}

void hro_tx_free(hro_tx *tx) { free(tx); }

void hro_tx_set_am_index(hro_tx *tx, float m)
{
    if (m >= 0 && m <= 1) tx->am_index = m;
}

void hro_tx_set_fm_deviation(hro_tx *tx, float dev)
{
    if (tx->fm_dev >= 0 && tx->fm_dev <= 3500) tx->fm_dev = dev;
}

void hro_tx_set_wbfm_deviation(hro_tx *tx, float dev)
{
    if (tx->wb_dev >= 0 && tx->wb_dev <= 112000) tx->wb_dev = dev;
}

void hro_tx_reset_mod(hro_tx *tx, int mod)
{
    tx_rail *rails[3] = {0, 0, 0};
    switch (mod) {
    case HRO_MOD_AM: rails[0] = &tx->am[0]; rails[1] = &tx->am[1]; break;
    case HRO_MOD_FM: rails[0] = &tx->fm[0]; rails[1] = &tx->fm[1]; break;
    case HRO_MOD_WBFM: rails[0] = &tx->wb_iq[0]; rails[1] = &tx->wb_iq[1]; rails[2] = &tx->wb_pcm; break;
    case HRO_MOD_SSB:
        rails[0] = &tx->ssb[0]; rails[1] = &tx->ssb[1];
        fir16_reset(&tx->ssb_delay);
        fir16_reset(&tx->ssb_hilbert);
        break;
    }
    for (int r = 0; r < 3; r++)
        if (rails[r])
            for (int s = 0; s < 8; s++) int16i_reset(&rails[r]->st[s]);
}

static void emit_iq(const int16_t *i8, const int16_t *q8, int n, int8_t *out)
{
    for (int k = 0; k < n; k++) {
        out[2 * k] = (int8_t)(uint8_t)((uint16_t)i8[k] & 0xff);
        out[2 * k + 1] = (int8_t)(uint8_t)((uint16_t)q8[k] & 0xff);
    }
}

/* The tool chain of signals/ (generateBaseband.sh:  <head> < pcm | interpolateSignal > x.iq).
 * head HRO_SIG_IQ:  in = n int16 I,Q pairs at 8 kS/s, straight into interpolateSignal (interpolateSignal.cc:266-340)
 *      HRO_SIG_DSB: signals/dsb.cc:38-47   x/4 on both rails
 *      HRO_SIG_AM:  signals/am.cc:38-50    (x*0.8 + 65536)/4 on both rails (the *0.8 is a double multiply)
 *      HRO_SIG_PM:  signals/pm.cc:39-55    angle = x/60000*pi (double multiply), (cos, sin)*16000
 *      HRO_SIG_FM:  signals/fm.cc:41-62    theta += x/65536*3.5, wrapped at +-2*pi, (cos, sin)*16000
 * Output: n*512 bytes of int8 I,Q at 2.048 MS/s. */
size_t hro_tx_signals(hro_tx *tx, int head, const int16_t *in, size_t n, int8_t *iq)
{
    int16_t ibuf[256], qbuf[256];
    for (size_t j = 0; j < n; j++) {
        int16_t i16, q16;
        switch (head) {
        case HRO_SIG_IQ:
            i16 = in[2 * j];
            q16 = in[2 * j + 1];
            break;
        case HRO_SIG_DSB: {
            volatile float s = (float)in[j];
            s = s / 4;
            i16 = q16 = f32_to_i16(s);
            break;
        }
        case HRO_SIG_AM: {
            volatile float s = (float)in[j];
            s = (float)((double)s * 0.8);
            s = s + 65536;
            s = s / 4;
            i16 = q16 = f32_to_i16(s);
            break;
        }
        case HRO_SIG_PM: {
            volatile float s = (float)in[j];
            s = s / 60000;
            s = (float)((double)s * M_PI);
            /* C++ <math.h>: cos(float) is the float overload */
            volatile float c = cosf(s) * 16000, q = sinf(s) * 16000;
            i16 = f32_to_i16(c);
            q16 = f32_to_i16(q);
            break;
        }
        case HRO_SIG_FM: { /* signals/fm.cc:41-62: kF = 3.5, theta wrapped at +-2*pi with double constants */
            volatile float tn = (float)in[j];
            tn = tn / 65536;
            tn = tn * 3.5f;
            volatile float theta = tx->sig_theta + tn;
            while (theta > (2 * M_PI)) theta = (float)(theta - (2 * M_PI));
            while (theta < (-(2 * M_PI))) theta = (float)(theta + (2 * M_PI));
            tx->sig_theta = theta;
            volatile float c = cosf(theta) * 16000, q = sinf(theta) * 16000;
            i16 = f32_to_i16(c);
            q16 = f32_to_i16(q);
            break;
        }
        default: return 0;
        }
        tx_rail_run(&tx->sig[0], 0, 7, i16, ibuf);
        tx_rail_run(&tx->sig[1], 0, 7, q16, qbuf);
        emit_iq(ibuf, qbuf, 256, iq + j * 512);
    }
    return n * 512;
}

size_t hro_tx_accept(hro_tx *tx, int mode, const int16_t *pcm, size_t n, int8_t *iq)
{
    int16_t ibuf[256], qbuf[256];
    for (size_t j = 0; j < n; j++) {
        int8_t *out = iq + j * 512;
        switch (mode) {
        case HRO_AM: { /* AmModulator.cc:574-607 */
            volatile float s = (float)pcm[j] / 32768;
            s = s * tx->am_index;
            s = s + 1;
            s = s / 2;
            volatile float v = s * 128;
            v = v * 250;
            int16_t m = f32_to_i16(v);
            tx_rail_run(&tx->am[0], 0, 7, m, ibuf);
            tx_rail_run(&tx->am[1], 0, 7, m, qbuf);
            emit_iq(ibuf, qbuf, 256, out);
            break;
        }
        case HRO_FM: { /* FmModulator.cc:586-622 */
            volatile float f = tx->fm_dev * (float)pcm[j];
            f = f / 32768;
            phase_set_frequency(&tx->fm_phase, f);
            float ph = phase_run(&tx->fm_phase);
            /* Nco::run (Nco.cc:186-199): cos()/sin() of a float -> cosf/sinf */
            volatile float c = cosf(ph), s = sinf(ph);
            c = c * 16000;
            s = s * 16000;
            tx_rail_run(&tx->fm[0], 0, 7, f32_to_i16(c), ibuf);
            tx_rail_run(&tx->fm[1], 0, 7, f32_to_i16(s), qbuf);
            emit_iq(ibuf, qbuf, 256, out);
            break;
        }
        case HRO_WBFM: { /* WbFmModulator.cc:347-365, 389-441, 583-632, 471-531 */
            int16_t p32[32];
            tx_rail_run(&tx->wb_pcm, 0, 4, pcm[j], p32);
            for (int k = 0; k < 32; k++) {
                volatile float f = tx->wb_dev * (float)p32[k];
                f = f / 1024;
                phase_set_frequency(&tx->wb_phase, f);
                float ph = phase_run(&tx->wb_phase);
                /* Nco::runFast (Nco.cc:222-257) */
                volatile float scaled = ph * 16384;
                int idx = (int)f64_to_i16((double)scaled / (2 * M_PI));
                idx += 8192;
                if (idx < 0) idx = 0;
                else if (idx > 16383) idx = 16383;
                volatile float c = g_cos[idx] * 900;
                volatile float s = g_sin[idx] * 900;
                tx_rail_run(&tx->wb_iq[0], 5, 7, f32_to_i16(c), ibuf);
                tx_rail_run(&tx->wb_iq[1], 5, 7, f32_to_i16(s), qbuf);
                emit_iq(ibuf, qbuf, 8, out + 16 * k);
            }
            break;
        }
        case HRO_LSB:
        case HRO_USB: { /* SsbModulator.cc:667-707 */
            volatile float s = (float)pcm[j];
            s = s / 2;
            int16_t half = f32_to_i16(s);
            int16_t id = fir16_push(&tx->ssb_delay, half);
            int16_t qh = fir16_push(&tx->ssb_hilbert, half);
            if (mode == HRO_USB) qh = (int16_t)(uint16_t)((uint32_t)(-(int)qh) & 0xffff);
            tx_rail_run(&tx->ssb[0], 0, 7, id, ibuf);
            tx_rail_run(&tx->ssb[1], 0, 7, qh, qbuf);
            emit_iq(ibuf, qbuf, 256, out);
            break;
        }
        default:
            memset(out, 0, 512);
        }
    }
    return n * 512;
}

/* ------------------------------------------------------------------ */
/* tables                                                              */
/* ------------------------------------------------------------------ */
int hro_taps(int which, int16_t *out, int cap)
{
    static const struct { const float *c; int n; } t[HRO_TAPS_COUNT] = {
        {k_fe1, 3}, {k_fe2, 3}, {k_fe3, 3}, {k_am1, 8}, {k_am2, 12}, {k_am3, 16},
        {k_fm_tuner, 32}, {k_fm_post, 12}, {k_audio40, 40}, {k_wbfm_post1, 8},
        {k_delay16, 16}, {k_hilbert31, 31}, {k_tx_hb8, 8}};
    if (which < 0 || which >= HRO_TAPS_COUNT) return -1;
    for (int i = 0; i < t[which].n && i < cap; i++) out[i] = quantise_q15(t[which].c[i]);
    return t[which].n;
}

void hro_atan2_table(float *out)
{
    tables_init();
    memcpy(out, g_atan2, sizeof g_atan2);
}

void hro_nco_tables(float *s, float *c)
{
    tables_init();
    memcpy(s, g_sin, sizeof g_sin);
    memcpy(c, g_cos, sizeof g_cos);
}
