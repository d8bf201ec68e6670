// hrd_api.cu -- the C ABI of libhrd_b200.so (include/hrd.h): batch handles, per-stream
// parameters with the reference's setter semantics, table construction, host<->device
// staging and kernel dispatch.  No DSP happens on the host; without a CUDA device every
// entry point fails (there is deliberately no CPU fallback).
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/hrd.h"
#include "hrd_tables.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define HRD_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) return fail(HRD_ECUDA, "%s: %s", #call, cudaGetErrorString(e_));     \
    } while (0)

// ---- the reference's filter designs (float literals as the reference spells them) -----
const float k_fe1[3] = {0.2504357f, 0.5000000f, 0.2504357f};   // IqDataProcessor.cc:8-13
const float k_fe2[3] = {0.2517491f, 0.4999998f, 0.2517491f};   // :15-20
const float k_fe3[3] = {0.2570951f, 0.5000000f, 0.2570951f};   // :22-27
const float k_am1[8] = {0.0242683f, 0.0766338f, 0.1457589f, 0.1959036f,
                        0.1959036f, 0.1457589f, 0.0766338f, 0.0242683f};
const float k_am2[12] = {0.0057496f, 0.0263853f, 0.0605301f, 0.1074406f, 0.1523486f, 0.1804951f,
                         0.1804951f, 0.1523486f, 0.1074406f, 0.0605301f, 0.0263853f, 0.0057496f};
const float k_am3[16] = {0.0116487f, 0.0152694f, -0.0109804f, -0.0611915f, -0.0736143f, 0.0187617f,
                         0.1988190f, 0.3481364f, 0.3481364f,  0.1988190f,  0.0187617f,  -0.0736143f,
                         -0.0611915f, -0.0109804f, 0.0152694f, 0.0116487f};
const float k_fm_tuner[32] = {
    0.0041331f, 0.0054174f, 0.0076016f, 0.0115481f, 0.0151685f, 0.0203192f, 0.0251608f, 0.0311322f,
    0.0366372f, 0.0427168f, 0.0480527f, 0.0533425f, 0.0575831f, 0.0611914f, 0.0635413f, 0.0648239f,
    0.0648239f, 0.0635413f, 0.0611914f, 0.0575831f, 0.0533425f, 0.0480527f, 0.0427168f, 0.0366372f,
    0.0311322f, 0.0251608f, 0.0203192f, 0.0151685f, 0.0115481f, 0.0076016f, 0.0054174f, 0.0041331f};
const float k_fm_post[12] = {0.0022977f, 0.0237042f, 0.0605386f, 0.1127073f, 0.1645167f, 0.1971107f,
                             0.1971107f, 0.1645167f, 0.1127073f, 0.0605386f, 0.0237042f, 0.0022977f};
const float k_audio40[40] = {
    0.0015969f,  -0.0111080f, -0.0270501f, -0.0265610f, -0.0023190f, 0.0180618f,  0.0065495f,  -0.0183409f,
    -0.0133345f, 0.0184489f,  0.0230891f,  -0.0161248f, -0.0363745f, 0.0091343f,  0.0550219f,  0.0070312f,
    -0.0862280f, -0.0497761f, 0.1793543f,  0.4145808f,  0.4145808f,  0.1793543f,  -0.0497761f, -0.0862280f,
    0.0070312f,  0.0550219f,  0.0091343f,  -0.0363745f, -0.0161248f, 0.0230891f,  0.0184489f,  -0.0133345f,
    -0.0183409f, 0.0065495f,  0.0180618f,  -0.0023190f, -0.0265610f, -0.0270501f, -0.0111080f, 0.0015969f};
// signals/interpolateSignal.cc:30-72
const float k_sig40[40] = {
    -0.0011405f, 0.0183372f,  0.0030542f,  -0.0100052f, -0.0059350f, 0.0115377f,  0.0109293f,
    -0.0120883f, -0.0175779f, 0.0110390f,  0.0262645f,  -0.0074772f, -0.0377408f, -0.0003152f,
    0.0541009f,  0.0165897f,  -0.0829085f, 0.0587608f,  0.1736804f,  0.4222137f,  0.4222137f,
    0.1736804f,  -0.0587608f, -0.0829085f, 0.0165897f,  0.0541009f,  -0.0003152f, -0.0377408f,
    -0.0074772f, 0.0262645f,  0.0110390f,  -0.0175779f, -0.0120883f, 0.0109293f,  0.0115377f,
    -0.0059350f, -0.0100052f, 0.0030542f,  0.0183372f,  -0.0011405f};
const float k_wbfm_post1[8] = {0.0243699f, 0.0769537f, 0.1463572f, 0.1967096f,
                               0.1967096f, 0.1463572f, 0.0769537f, 0.0243699f};
const float k_delay16[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1};
const float k_hilbert31[31] = {-0.0033953f, 0, -0.0058652f, 0, -0.0134385f, 0, -0.0281423f, 0,
                               -0.0534836f, 0, -0.0980394f, 0, -0.1935638f, 0, -0.6302204f, 0,
                               0.6302204f,  0, 0.1935638f,  0, 0.0980394f,  0, 0.0534836f,  0,
                               0.0281423f,  0, 0.0134385f,  0, 0.0058652f,  0, 0.0033953f};
const float k_tx_hb8[8] = {-0.0440934f, 0, 0.2913764f, 0.5000000f, 0.2913764f, 0, -0.0440934f, 0};

struct TapSet {
    const float *c;
    int n;
};
// numbering shared with oracle/hrd_oracle.h HRO_TAPS_* so tests can compare table by table
const TapSet k_tapsets[] = {{k_fe1, 3},      {k_fe2, 3},     {k_fe3, 3},       {k_am1, 8},      {k_am2, 12},
                            {k_am3, 16},     {k_fm_tuner, 32}, {k_fm_post, 12}, {k_audio40, 40}, {k_wbfm_post1, 8},
                            {k_delay16, 16}, {k_hilbert31, 31}, {k_tx_hb8, 8}};
const int k_n_tapsets = (int)(sizeof k_tapsets / sizeof k_tapsets[0]);

// (int16_t)round(c*32768) evaluated in float; the narrowing keeps the low 16 bits, so
// 1.0 -> 32768 -> -32768 exactly as the reference build does
// taps of an N-tap decimator for mac_pair (hrd_tables.h): word w = {tl(q[N-1-2w]), tl(q[N-2-2w]), th(..), th(..)}
void split_taps(const int32_t *q, int n, uint32_t *out)
{
    for (int w = 0; w < n / 2; w++) {
        const int q0 = q[n - 1 - 2 * w], q1 = q[n - 2 - 2 * w];
        out[w] = (uint32_t)(q0 & 0xff) | (uint32_t)(q1 & 0xff) << 8 | (uint32_t)((q0 >> 8) & 0xff) << 16 |
                 (uint32_t)((q1 >> 8) & 0xff) << 24;
    }
}

int16_t quantise(float c)
{
    float scaled = c * 32768;
    scaled = roundf(scaled);
    int32_t v = (int32_t)scaled;
    return (int16_t)(uint16_t)((uint32_t)v & 0xffffu);
}

uint32_t pair16(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }

int build_tables(hrd::ConstTables &t)
{
    memset(&t, 0, sizeof t);
    const float *fe[3] = {k_fe1, k_fe2, k_fe3};
    for (int s = 0; s < 3; s++) {
        int q0 = quantise(fe[s][0]), q1 = quantise(fe[s][1]), q2 = quantise(fe[s][2]);
        // doubled taps must fit an unsigned 16-bit dp2a operand
        if (q0 < 0 || q1 < 0 || q2 < 0 || 2 * q0 > 65535 || 2 * q1 > 65535 || 2 * q2 > 65535)
            return fail(HRD_EINVAL, "front-end taps do not fit the dp2a encoding");
        t.fe_a[s] = pair16(2 * q1, 2 * q0);
        t.fe_b[s] = pair16(0, 2 * q2);
    }
    for (int i = 0; i < 16; i++) t.fm_tuner[i] = pair16(quantise(k_fm_tuner[31 - 2 * i]), quantise(k_fm_tuner[30 - 2 * i]));
    for (int i = 0; i < 4; i++) t.am1[i] = pair16(quantise(k_am1[7 - 2 * i]), quantise(k_am1[6 - 2 * i]));
    for (int i = 0; i < 12; i++) t.am2[i] = quantise(k_am2[i]);
    for (int i = 0; i < 16; i++) t.am3[i] = quantise(k_am3[i]);
    for (int i = 0; i < 12; i++) t.fm_post[i] = quantise(k_fm_post[i]);
    for (int i = 0; i < 40; i++) t.audio40[i] = quantise(k_audio40[i]);
    for (int i = 0; i < 40; i++) t.sig40[i] = quantise(k_sig40[i]);
    for (int i = 0; i < 8; i++) t.wbfm_post1[i] = quantise(k_wbfm_post1[i]);
    split_taps(t.wbfm_post1, 8, t.wb1_sp);
    split_taps(t.fm_post, 12, t.fm_post_sp);
    split_taps(t.audio40, 40, t.audio40_sp);
    split_taps(t.am2, 12, t.am2_sp);
    split_taps(t.am3, 16, t.am3_sp);
    for (int i = 0; i < 31; i++) t.hilbert[i] = quantise(k_hilbert31[i]);
    for (int i = 0; i < 16; i++) t.delay[i] = quantise(k_delay16[i]);
    for (int i = 0; i < 8; i++) t.tx_hb8[i] = quantise(k_tx_hb8[i]);
    t.k_32768 = 32768;
    // DbfsCalculator::DbfsCalculator (DbfsCalculator.cc:56-65): (int32_t)(20 * log10((float)i)), the float
    // overload of log10 as the C++ reference resolves it; entry 0 repeats entry 1
    for (int i = 1; i <= 256; i++) {
        volatile float level = 20 * log10f((float)i);
        t.db_table[i] = (int32_t)level;
    }
    t.db_table[0] = t.db_table[1];
    {
        const float hi = -6.28318548202514648438f, lo = 1.74845553146951715e-07f; // -fl32(2*pi), -(2*pi - fl32(2*pi))
        t.k_sign = 0x80000000u;
        memcpy(&t.k_m2pi_hi, &hi, 4);
        memcpy(&t.k_m2pi_lo, &lo, 4);
    }
    t.tx_c3 = quantise(k_fe3[0]);
    t.tx_m3 = quantise(k_fe3[1]);
    t.tx_c7 = quantise(k_fe2[0]);
    t.tx_m7 = quantise(k_fe2[1]);
    t.tx_c8 = quantise(k_fe1[0]);
    t.tx_m8 = quantise(k_fe1[1]);
    // structural facts the kernels rely on
    for (int i = 1; i < 31; i += 2)
        if (t.hilbert[i] != 0) return fail(HRD_EINVAL, "Hilbert odd taps expected to be zero");
    for (int i = 0; i < 15; i++)
        if (t.delay[i] != 0) return fail(HRD_EINVAL, "delay-line taps expected to be zero");
    // Tx half-band designs: 8 taps {a,0,b,16384,b,0,a,0} and 4 taps {c,16384,c,0} with 0 < c <= 16384.
    // With that structure the odd branches are (x+1)>>1 and no stage output can leave int16
    // (hrd_tx.cu hb4_even / hb4_odd / interp8 rely on it).
    const int32_t *h = t.tx_hb8;
    if (h[1] || h[5] || h[7] || h[3] != 16384 || h[0] != h[6] || h[2] != h[4] ||
        2 * (abs(h[0]) + abs(h[2])) > 32767)
        return fail(HRD_EINVAL, "Tx 8-tap half-band does not have the expected structure");
    if (t.tx_m3 != 16384 || t.tx_m7 != 16384 || t.tx_m8 != 16384)
        return fail(HRD_EINVAL, "Tx 4-tap half-band centre taps expected to be 16384");
    if (t.tx_c3 <= 0 || t.tx_c3 > 16384 || t.tx_c7 <= 0 || t.tx_c7 > 16384 || t.tx_c8 <= 0 || t.tx_c8 > 16384)
        return fail(HRD_EINVAL, "Tx 4-tap half-band outer taps out of the supported range");
    // the fp16-pair form of stages 6..8 (hrd_tx.cu tail3_h2, proved by tools/verify_tx_tail_h2.c) has these three
    // taps built into its constants
    if (t.tx_c3 != 8424 || t.tx_c7 != 8249 || t.tx_c8 != 8206)
        return fail(HRD_EINVAL, "Tx stage 6/7/8 taps are not 8424 / 8249 / 8206: the packed-half form does not apply");
    return HRD_OK;
}

// per-device tables ---------------------------------------------------------------------
struct DeviceTables {
    bool ready = false;
    float *atan2_lut = nullptr;
    float *nco_sin = nullptr, *nco_cos = nullptr;
    uint32_t *nco_iq900 = nullptr;
    float *nco_thr = nullptr;
};
std::mutex g_tab_mutex;
DeviceTables g_dev_tables[64];

int ensure_tables(int device)
{
    std::lock_guard<std::mutex> lock(g_tab_mutex);
    DeviceTables &d = g_dev_tables[device];
    if (d.ready) return HRD_OK;
    hrd::ConstTables t;
    int rc = build_tables(t);
    if (rc) return rc;
    hrd::upload_tables(t);
    hrd::upload_tables_tx(t);
    HRD_CUDA(cudaGetLastError());
    // atan2 table (FmDemodulator.cc:158-170): double atan2 narrowed to float, [q+128][i+128]
    std::vector<float> lut(65536);
    for (int x = 0; x < 256; x++)
        for (int y = 0; y < 256; y++) lut[(size_t)y * 256 + x] = (float)atan2((double)y - 128, (double)x - 128);
    // rx_wbfm_kernel keeps only the rows q >= 0 on chip and mirrors the rest: that needs the table
    // to be odd in q bit for bit (glibc's atan2 is; anything else is refused, not approximated)
    for (int q = 1; q < 128; q++)
        for (int x = 0; x < 256; x++) {
            const float up = lut[(size_t)(128 + q) * 256 + x], dn = lut[(size_t)(128 - q) * 256 + x];
            const float mirrored = -dn;
            if (memcmp(&up, &mirrored, sizeof up) != 0)
                return fail(HRD_EINVAL, "host atan2 table is not odd-symmetric at q=%d: unsupported libm", q);
        }
    // NCO tables (Nco.cc:45-61): the angle is accumulated in float, sinf/cosf of it
    std::vector<float> s(16384), c(16384);
    float inc = (float)(2 * M_PI / 16384);
    volatile float ang = (float)(-M_PI);
    for (int i = 0; i < 16384; i++) {
        s[i] = sinf(ang);
        c[i] = cosf(ang);
        ang = ang + inc;
    }
    HRD_CUDA(cudaMalloc(&d.atan2_lut, 65536 * sizeof(float)));
    HRD_CUDA(cudaMalloc(&d.nco_sin, 16384 * sizeof(float)));
    HRD_CUDA(cudaMalloc(&d.nco_cos, 16384 * sizeof(float)));
    HRD_CUDA(cudaMemcpy(d.atan2_lut, lut.data(), 65536 * sizeof(float), cudaMemcpyHostToDevice));
    HRD_CUDA(cudaMemcpy(d.nco_sin, s.data(), 16384 * sizeof(float), cudaMemcpyHostToDevice));
    HRD_CUDA(cudaMemcpy(d.nco_cos, c.data(), 16384 * sizeof(float), cudaMemcpyHostToDevice));
    // WbFmModulator.cc:606-626: iSample *= 900; (int16_t)iSample -- a pure function of the table
    // entry, so it is tabulated too (float multiply, truncation toward zero; |v| <= 900)
    std::vector<uint32_t> iq900(16384);
    for (int i = 0; i < 16384; i++) {
        volatile float ci = c[i] * 900.0f, si = s[i] * 900.0f;
        iq900[i] = pair16((int16_t)(int32_t)ci, (int16_t)(int32_t)si);
    }
    // Nco::runFast (Nco.cc:231-233): thr[k] = the smallest float phase >= 0 for which the
    // reference's own expression gives an index offset >= k; found by walking floats around
    // 2*pi*k/16384, so the table is exact by construction (hrd_tx.cu nco_index searches it)
    std::vector<float> thr(8194);
    auto ref_index = [](float ph) { return (int)(int16_t)(int32_t)((double)(ph * 16384.0f) / (2 * M_PI)); };
    thr[0] = 0.0f;
    for (int k = 1; k <= 8192; k++) {
        float ph = (float)(2 * M_PI * k / 16384.0);
        while (ref_index(ph) >= k) ph = nextafterf(ph, 0.0f);
        while (ref_index(ph) < k) ph = nextafterf(ph, 100.0f);
        thr[(size_t)k] = ph;
    }
    thr[8193] = INFINITY;
    HRD_CUDA(cudaMalloc(&d.nco_thr, thr.size() * sizeof(float)));
    HRD_CUDA(cudaMemcpy(d.nco_thr, thr.data(), thr.size() * sizeof(float), cudaMemcpyHostToDevice));
    // folded about phase 0 (Nco::runFast truncates toward zero, so the index is 8192 +- k with k from |phase|,
    // clamped to 16383): [k] serves phase >= 0, [8193 + k] phase < 0
    std::vector<uint32_t> fold(2 * 8193);
    for (int k = 0; k <= 8192; k++) {
        fold[(size_t)k] = iq900[(size_t)std::min(8192 + k, 16383)];
        fold[(size_t)(8193 + k)] = iq900[(size_t)(8192 - k)];
    }
#if HRD_TW_H2
    // stages 6-8 run on fp16 pairs (hrd_tx.cu tail3_h2): the table holds the same integers as binary16 numbers
    for (uint32_t &w : fold) {
        const int vi = (int16_t)(w & 0xffffu), vq = (int16_t)(w >> 16);
        if (abs(vi) > 900 || abs(vq) > 900) return fail(HRD_EINVAL, "NCO table entry beyond 900");
        const __half hi = __float2half((float)vi), hq = __float2half((float)vq);
        uint16_t bi, bq;
        memcpy(&bi, &hi, 2);
        memcpy(&bq, &hq, 2);
        w = (uint32_t)bi | (uint32_t)bq << 16;
    }
#endif
    HRD_CUDA(cudaMalloc(&d.nco_iq900, fold.size() * sizeof(uint32_t)));
    HRD_CUDA(cudaMemcpy(d.nco_iq900, fold.data(), fold.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    d.ready = true;
    return HRD_OK;
}

int kernel_kind_of_mode(int mode)
{
    switch (mode) {
    case HRD_MODE_AM: return hrd::K_AM;
    case HRD_MODE_FM: return hrd::K_FM;
    case HRD_MODE_WBFM: return hrd::K_WBFM;
    case HRD_MODE_LSB:
    case HRD_MODE_USB: return hrd::K_SSB;
    case HRD_MODE_IQ8K:
    case HRD_MODE_DSB:
    case HRD_MODE_PM:
    case HRD_MODE_AM_PROTO:
    case HRD_MODE_FM_PROTO: return hrd::K_IQ;
    default: return hrd::K_NONE;
    }
}

} // namespace

#define HRD_PROFILE_RING 32

struct hrd_batch {
    int device = 0, n = 0, kind = HRD_RX;
    // CONTROL STATE.  The reference's setters are plain member writes that the UI thread makes while the data
    // thread runs (Radio.cc:1973, 2404-2633).  Here they write these host vectors under `ctl` and bump `gen`; the
    // next process call snapshots them under the same lock and uploads the snapshot ON ITS STREAM (pinned
    // staging, stream-ordered: no device-wide synchronisation, nothing blocks).  A setter that lands while a call
    // is uploading simply leaves gen ahead of `uploaded`, and the call after picks it up.
    std::mutex ctl;
    std::vector<int32_t> mode;
    std::vector<uint8_t> lsb;
    std::vector<float> param[HRD_PARAM_COUNT];
    uint64_t gen = 1, uploaded = 0;
    void *h_stage[2] = {nullptr, nullptr}; // pinned: {kind[n], ids[n], lsb[n], mode[n], param[P][n]}
    cudaEvent_t ev_stage[2] = {nullptr, nullptr};
    int stage_at = 0;
    bool armed = false;                    // some stream's squelch can close (snapshot of squelch_armed)
    bool any_iq8k = false, any_fm_proto = false;
    // calls on one batch are serialised by the caller, but they may name different CUDA streams: every call waits
    // for the end of the one before it and marks its own end here
    cudaEvent_t ev_last = nullptr;
    void *d_state[2] = {nullptr, nullptr}; // double-buffered, see hrd::RxParams / hrd::TxParams
    int cur = 0;                           // the half the next call reads
    int sm_count = 148;
    int opt[HRD_OPT_COUNT] = {};
    float *d_pre = nullptr;                // Rx AM/SSB: IIR input scratch, [n][pre_stride]
    void *d_wbv = nullptr;                 // Rx WBFM: verification pairs, [n][n_tiles] float2
    void *d_fmph = nullptr;                // Tx FM: NCO phase per PCM sample, [n_fm][n8]
    size_t d_fmph_cap = 0;
    void *d_sigph = nullptr;               // Tx signals/fm.cc head: theta per PCM sample, [n_iq][n8]
    size_t d_sigph_cap = 0;
    size_t d_wbv_cap = 0;
    // [0] = streams that failed the verification in this call (tiled retry), [1] = those the retry failed for too
    // (serial re-run); +2 words: 64-bit total of [0], +4 words: 64-bit total of [1]
    uint32_t *d_wbflag = nullptr;
    int32_t *d_wbrerun = nullptr;          // their ids: [n] for the retry, [n] for the serial re-run
    float *d_wbguess = nullptr;            // [n] the retry's start values
    void *d_wbv2 = nullptr;                // the retry's own check pairs
    size_t d_wbv2_cap = 0;
    size_t d_pre_cap = 0, pre_stride = 0;
    int32_t *d_ids = nullptr; // streams grouped by kernel kind
    int32_t *d_all = nullptr; // 0..n-1
    uint8_t *d_lsb = nullptr;
    uint8_t *d_kind = nullptr; // kernel kind (hrd::K_*) of every stream
    float *d_param[HRD_PARAM_COUNT] = {};
    int group_off[hrd::K_COUNT] = {}, group_cnt[hrd::K_COUNT] = {};
    int32_t *d_mode = nullptr; // HRD_MODE_* of every stream
    void *d_in = nullptr, *d_out = nullptr;
    size_t d_in_cap = 0, d_out_cap = 0;
    // squelch gate (hrd_squelch.cu): the 256 kS/s stream of the call, per-(stream, block) magnitudes,
    // decisions and output offsets, one block of PCM, per-block stream lists, tracker state
    void *d_sq256 = nullptr, *d_sq_mag = nullptr, *d_sq_open = nullptr, *d_sq_n256 = nullptr;
    size_t d_sq256_cap = 0, d_sq_mag_cap = 0, d_sq_open_cap = 0;
    uint8_t *d_sq_track = nullptr;
    // what the gate decided, on its way to the host: pinned staging filled by stream-ordered copies, unpacked by a
    // host function on the stream, ev_report behind it (hrd_rx_squelch_report waits for that, nothing else does)
    void *h_sq = nullptr;
    size_t h_sq_cap = 0;
    cudaEvent_t ev_report = nullptr;
    std::vector<uint8_t> h_kind;           // kernel kind of every stream (host copy of d_kind)
    std::vector<uint32_t> sq_mag;          // latest squelched call: [n][sq_blocks]
    std::vector<uint8_t> sq_open;
    uint32_t sq_blocks = 0;
    uint64_t gated_calls = 0;
    cudaStream_t own = nullptr;
    // mixed-mode batches: the tile kernels of the different modes run side by side (run_demods)
    cudaStream_t aux[4] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[4] = {};
    uint64_t launches = 0;
    // HRD_OPT_PROFILE: events around the hot kernel(s) of the latest call and around its tail kernel
    // (a ring of the last HRD_PROFILE_RING calls, so back-to-back timed steps can all be read)
    cudaEvent_t ev[HRD_PROFILE_RING][3] = {};
    cudaEvent_t ev_kind[HRD_PROFILE_RING][4][2] = {}; // HRD_OPT_RX_SERIAL: around each kind's kernels
    bool ev_kind_set[HRD_PROFILE_RING][4] = {};
    uint64_t ev_calls = 0;
};

namespace {

size_t state_size(int kind) { return kind == HRD_RX ? sizeof(hrd::RxState) : sizeof(hrd::TxState); }

size_t stage_bytes(int n) { return (size_t)n * (1 + 4 + 1 + 4 + 4 * HRD_PARAM_COUNT) + 64; }

// Upload the control state if a setter ran since the last upload: snapshot under the lock, copy from pinned staging
// on the call's stream.  Everything queued before on that stream (and, through ev_last, every earlier call) is
// ordered in front of the copies, everything this call launches behind them.
int regroup(hrd_batch *b, cudaStream_t s)
{
    std::lock_guard<std::mutex> lock(b->ctl);
    if (b->uploaded == b->gen) return HRD_OK;
    const size_t n = (size_t)b->n;
    const int at = b->stage_at;
    HRD_CUDA(cudaEventSynchronize(b->ev_stage[at])); // the copies that last used this staging half are long done
    char *st = (char *)b->h_stage[at];
    uint8_t *kinds = (uint8_t *)st;
    int32_t *ids = (int32_t *)(st + ((n + 15) & ~(size_t)15));
    uint8_t *lsb = (uint8_t *)(ids + n);
    int32_t *mode = (int32_t *)((char *)lsb + ((n + 15) & ~(size_t)15));
    float *param = (float *)(mode + n);
    b->armed = b->opt[HRD_OPT_RX_SQUELCH] != 0;
    b->any_iq8k = b->any_fm_proto = false;
    for (size_t i = 0; i < n; i++) {
        kinds[i] = (uint8_t)kernel_kind_of_mode(b->mode[i]);
        lsb[i] = b->lsb[i];
        mode[i] = b->mode[i];
        b->any_iq8k |= b->mode[i] == HRD_MODE_IQ8K;
        b->any_fm_proto |= b->mode[i] == HRD_MODE_FM_PROTO;
        // can the gate of this stream close?  dBFS >= dbTable[0] - 42 - gain (DbfsCalculator.cc, SignalDetector.cc:263)
        if (b->kind == HRD_RX && (double)b->param[HRD_PARAM_SQUELCH_THRESHOLD][i] > -42.0 - (double)b->param[HRD_PARAM_RX_GAIN_DB][i])
            b->armed = true;
    }
    for (int p = 0; p < HRD_PARAM_COUNT; p++) memcpy(param + (size_t)p * n, b->param[p].data(), n * sizeof(float));
    // AM and SSB streams sit next to each other: on Rx one launch runs both
    static const int order[hrd::K_COUNT] = {hrd::K_NONE, hrd::K_FM, hrd::K_WBFM, hrd::K_AM, hrd::K_SSB, hrd::K_IQ};
    size_t fill = 0;
    for (int o = 0; o < hrd::K_COUNT; o++) {
        const int k = order[o];
        b->group_off[k] = (int)fill;
        for (size_t i = 0; i < n; i++)
            if (kinds[i] == k) ids[fill++] = (int32_t)i;
        b->group_cnt[k] = (int)fill - b->group_off[k];
    }
    b->h_kind.assign(kinds, kinds + n);
    HRD_CUDA(cudaMemcpyAsync(b->d_kind, kinds, n, cudaMemcpyHostToDevice, s));
    HRD_CUDA(cudaMemcpyAsync(b->d_ids, ids, n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    HRD_CUDA(cudaMemcpyAsync(b->d_lsb, lsb, n, cudaMemcpyHostToDevice, s));
    HRD_CUDA(cudaMemcpyAsync(b->d_mode, mode, n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    for (int p = 0; p < HRD_PARAM_COUNT; p++)
        HRD_CUDA(cudaMemcpyAsync(b->d_param[p], param + (size_t)p * n, n * sizeof(float), cudaMemcpyHostToDevice, s));
    HRD_CUDA(cudaEventRecord(b->ev_stage[at], s));
    b->stage_at ^= 1;
    b->uploaded = b->gen;
    return HRD_OK;
}

// every call: behind the end of the call before it, whatever stream that one ran on
int order_after_last(hrd_batch *b, cudaStream_t s)
{
    HRD_CUDA(cudaStreamWaitEvent(s, b->ev_last, 0));
    return HRD_OK;
}
int mark_end(hrd_batch *b, cudaStream_t s)
{
    HRD_CUDA(cudaEventRecord(b->ev_last, s));
    return HRD_OK;
}

int ensure_cap(void **ptr, size_t *cap, size_t need)
{
    if (*cap >= need) return HRD_OK;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr;
    *cap = 0;
    HRD_CUDA(cudaMalloc(ptr, need));
    *cap = need;
    // Scratch rows are padded (8 floats, 16-byte copies): kernels copy the padding along and never use it.  Zero it
    // once, on growth only, so that nothing ever reads uninitialised memory (compute-sanitizer initcheck stays clean).
    HRD_CUDA(cudaMemset(*ptr, 0, need));
    HRD_CUDA(cudaDeviceSynchronize());
    return HRD_OK;
}

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int check_stream_arg(hrd_batch *b, int stream)
{
    if (!b) return fail(HRD_EINVAL, "null batch");
    if (stream != HRD_ALL_STREAMS && (stream < 0 || stream >= b->n))
        return fail(HRD_EINVAL, "stream %d out of range [0,%d)", stream, b->n);
    return HRD_OK;
}

// zero [off, off+len) of every selected stream's state record
int zero_state(hrd_batch *b, int stream, size_t off, size_t len)
{
    const size_t pitch = state_size(b->kind);
    char *base = (char *)b->d_state[b->cur] + off;
    if (stream == HRD_ALL_STREAMS)
        HRD_CUDA(cudaMemset2DAsync(base, pitch, 0, len, (size_t)b->n, b->own));
    else
        HRD_CUDA(cudaMemsetAsync(base + (size_t)stream * pitch, 0, len, b->own));
    return HRD_OK;
}

// How a call is cut into time tiles (hrd_rx.cu header).  Auto mode searches the tile count
// that wastes least: slots left empty in the last wave of warps plus the halo every tile after
// the first re-reads and recomputes.
void choose_tiles(hrd_batch *b, int kind, int entry, int n_streams, uint32_t n_batches, int32_t *n_tiles,
                  uint32_t *tile_batches)
{
    const uint32_t halo = (uint32_t)hrd::rx_halo_batches(kind);
    uint32_t tb = n_batches ? n_batches : 1;
    const bool allowed = kind != hrd::K_WBFM || b->opt[HRD_OPT_RX_WBFM_TILING];
    if (allowed && n_batches > 1) {
        if (b->opt[HRD_OPT_RX_TILE_BATCHES] > 0) {
            tb = (uint32_t)b->opt[HRD_OPT_RX_TILE_BATCHES];
            if (tb < halo) tb = halo;
        } else {
            const double slots = (double)b->sm_count * hrd::rx_resident_warps_per_sm(kind, entry);
            double best = -1.0;
            for (uint32_t t = 1; t <= 512 && t <= n_batches; t++) {
                const uint32_t cand = (n_batches + t - 1) / t; // batches per tile
                if (t > 1 && cand < 2 * halo) break;           // halo would exceed half a tile
                const uint32_t tiles = (n_batches + cand - 1) / cand;
                const double items = (double)tiles * n_streams;
                const double waves = ceil(items / slots);
                double fill = items / (waves * slots);
                if (kind == hrd::K_WBFM) {
                    // one CTA per SM whose size is balanced over the waves (hrd_rx.cu launch_wbfm); an SM needs
                    // about 16 item warps to stay busy (measured: 7 items per CTA run at 0.45 of the rate of 28)
                    const int ipc = hrd::balanced_items_per_cta((long long)items, b->sm_count, hrd::rx_resident_warps_per_sm(kind, entry));
                    fill = ceil(items / ipc) / (waves * b->sm_count) * (ipc >= 16 ? 1.0 : ipc / 16.0);
                }
                const double useful = (double)cand / (double)(cand + (tiles > 1 ? halo : 0));
                const double score = fill * useful;
                if (score > best + 1e-9) {
                    best = score;
                    tb = cand;
                }
            }
        }
        if (tb > n_batches) tb = n_batches;
    }
    *tile_batches = tb;
    *n_tiles = (int32_t)((n_batches + tb - 1) / tb);
    if (*n_tiles < 1) *n_tiles = 1;
}

// WBFM launches run one CTA per SM, so a stream count just above a whole number of waves (4096 streams on
// 148 x 27 slots) would leave the last wave nearly empty, or force every stream into tiles.  Instead the list is
// cut in two: the streams that fill whole waves run untiled, the rest are tiled finely enough to share one more
// (short) wave.  Cost model, in batch-times of a full CTA: waves x (batches per tile + halo) x the step time of
// the CTA size the launcher will pick (flat below 16 items: measured, 7 items per CTA run at 0.45 of the rate of 28).
struct WbPart {
    int first, count;
    int32_t n_tiles;
    uint32_t tile_batches;
};
static double wbfm_uniform_cost(hrd_batch *b, int entry, int n_streams, uint32_t n_batches, int32_t *n_tiles, uint32_t *tile_batches)
{
    choose_tiles(b, hrd::K_WBFM, entry, n_streams, n_batches, n_tiles, tile_batches);
    const int cap = hrd::rx_resident_warps_per_sm(hrd::K_WBFM, entry);
    const double items = (double)*n_tiles * n_streams;
    const double waves = ceil(items / ((double)b->sm_count * cap));
    const int ipc = hrd::balanced_items_per_cta((long long)items, b->sm_count, cap);
    const uint32_t halo = *n_tiles > 1 ? (uint32_t)hrd::rx_halo_batches(hrd::K_WBFM) : 0u;
    return waves * (double)(*tile_batches + halo) * (double)std::max(ipc, 16) / (double)cap;
}
static int plan_wbfm(hrd_batch *b, int entry, int n_streams, uint32_t n_batches, bool one_tile, WbPart parts[2])
{
    parts[0].first = 0;
    parts[0].count = n_streams;
    if (one_tile) { // ragged calls (the squelched path): one warp walks a stream's whole call
        parts[0].n_tiles = 1;
        parts[0].tile_batches = n_batches ? n_batches : 1;
        return 1;
    }
    const double whole = wbfm_uniform_cost(b, entry, n_streams, n_batches, &parts[0].n_tiles, &parts[0].tile_batches);
    const int slots = b->sm_count * hrd::rx_resident_warps_per_sm(hrd::K_WBFM, entry);
    const int n_main = n_streams / slots * slots;
    if (n_main == 0 || n_main == n_streams || b->opt[HRD_OPT_RX_TILE_BATCHES] > 0 || !b->opt[HRD_OPT_RX_WBFM_TILING] || n_batches < 2)
        return 1;
    WbPart rest;
    rest.first = n_main;
    rest.count = n_streams - n_main;
    const double split = (double)(n_main / slots) * (double)n_batches
                         + wbfm_uniform_cost(b, entry, rest.count, n_batches, &rest.n_tiles, &rest.tile_batches);
    if (split >= whole) return 1;
    parts[0].count = n_main;
    parts[0].n_tiles = 1;
    parts[0].tile_batches = n_batches;
    parts[1] = rest;
    return 2;
}

// Tx: tiles of whole 32-sample batches; same trade-off as choose_tiles (fill of the last wave of
// resident warps against the halo every tile after the first recomputes)
void choose_tx_tiles(hrd_batch *b, int kind, int n_streams, uint32_t n8, int32_t *n_tiles, uint32_t *tile_len8)
{
    const uint32_t n_batches = (n8 + 31) / 32;
    const uint32_t halo = (uint32_t)hrd::tx_halo_samples(kind) / 32;
    uint32_t tb = n_batches ? n_batches : 1;
    if (n_batches > 1) {
        if (b->opt[HRD_OPT_TX_TILE_SAMPLES] > 0) {
            tb = ((uint32_t)b->opt[HRD_OPT_TX_TILE_SAMPLES] + 31) / 32;
            if (tb < halo) tb = halo;
        } else {
            const double slots = (double)b->sm_count * hrd::tx_resident_warps_per_sm(kind);
            double best = -1.0;
            for (uint32_t t = 1; t <= 512 && t <= n_batches; t++) {
                const uint32_t cand = (n_batches + t - 1) / t;
                if (t > 1 && cand < 4 * halo) break; // keep the halo below a fifth of the work
                const uint32_t tiles = (n_batches + cand - 1) / cand;
                const double items = (double)tiles * n_streams;
                const double waves = ceil(items / slots);
                const double fill = items / (waves * slots);
                const double useful = (double)cand / (double)(cand + (tiles > 1 ? halo : 0));
                const double score = fill * useful;
                if (score > best + 1e-9) {
                    best = score;
                    tb = cand;
                }
            }
        }
        if (tb > n_batches) tb = n_batches;
    }
    *tile_len8 = tb * 32;
    *n_tiles = (int32_t)((n_batches + tb - 1) / tb);
    if (*n_tiles < 1) *n_tiles = 1;
}

} // namespace

extern "C" {

int hrd_abi_version(void) { return HRD_ABI_VERSION; }

const char *hrd_last_error(void) { return g_err; }

size_t hrd_state_bytes_per_stream(int kind) { return state_size(kind); }

int hrd_get_taps(int which, int16_t *out, int cap)
{
    if (which < 0 || which >= k_n_tapsets || !out) return fail(HRD_EINVAL, "bad tap set %d", which);
    for (int i = 0; i < k_tapsets[which].n && i < cap; i++) out[i] = quantise(k_tapsets[which].c[i]);
    return k_tapsets[which].n;
}

int hrd_create(int device, int n_streams, int kind, hrd_batch_t **out)
{
    if (!out) return fail(HRD_EINVAL, "out is null");
    *out = nullptr;
    if (n_streams <= 0) return fail(HRD_EINVAL, "n_streams must be positive");
    if (kind != HRD_RX && kind != HRD_TX) return fail(HRD_EINVAL, "kind must be HRD_RX or HRD_TX");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return fail(HRD_ENODEV, "no CUDA device: libhrd_b200 has no CPU fallback");
    if (device < 0 || device >= count || device >= 64) return fail(HRD_ENODEV, "device %d not present", device);
    cudaDeviceProp prop;
    HRD_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(HRD_ENODEV, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                    prop.minor);
    DeviceGuard guard(device);
    if (!guard.ok) return fail(HRD_ECUDA, "cudaSetDevice(%d) failed", device);
    int rc = ensure_tables(device);
    if (rc) return rc;
    hrd_batch *b = new (std::nothrow) hrd_batch;
    if (!b) return fail(HRD_ENOMEM, "out of host memory");
    b->device = device;
    b->n = n_streams;
    b->kind = kind;
    b->mode.assign((size_t)n_streams, HRD_MODE_NONE); // IqDataProcessor.cc:70, BasebandDataProcessor ctor
    b->lsb.assign((size_t)n_streams, 1);              // SsbDemodulator.cc:143, SsbModulator.cc:278
    const float defaults[HRD_PARAM_COUNT] = {300.f,                              // AmDemodulator.cc:102
                                             (float)(64000 / (2 * M_PI)),        // FmDemodulator.cc:173
                                             (float)(256000 / (2 * M_PI)),       // WbFmDemodulator.cc:151
                                             300.f,                              // SsbDemodulator.cc:146
                                             0.8f,                               // AmModulator.cc:218
                                             3500.f,                             // FmModulator.cc:218
                                             70000.f,                            // WbFmModulator.cc:204
                                             -200.f,                             // IqDataProcessor.cc:120-124
                                             16.f};                              // Radio.cc:413
    for (int p = 0; p < HRD_PARAM_COUNT; p++) b->param[p].assign((size_t)n_streams, defaults[p]);
    cudaError_t e = cudaSuccess;
    const size_t ssz = state_size(kind) * (size_t)n_streams;
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&b->own, cudaStreamNonBlocking);
    for (int i = 0; i < 4 && e == cudaSuccess; i++) {
        e = cudaStreamCreateWithFlags(&b->aux[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->ev_join[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->ev_last, cudaEventDisableTiming);
    for (int h = 0; h < 2 && e == cudaSuccess; h++) {
        e = cudaHostAlloc(&b->h_stage[h], stage_bytes(n_streams), cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->ev_stage[h], cudaEventDisableTiming);
    }
    b->sm_count = prop.multiProcessorCount;
    for (int h = 0; h < 2; h++) {
        if (e == cudaSuccess) e = cudaMalloc(&b->d_state[h], ssz);
        if (e == cudaSuccess) e = cudaMemset(b->d_state[h], 0, ssz);
    }
    if (e == cudaSuccess) e = cudaMalloc(&b->d_wbflag, 32);
    if (e == cudaSuccess) e = cudaMemset(b->d_wbflag, 0, 32);
    if (e == cudaSuccess) e = cudaMalloc(&b->d_wbrerun, 2 * (size_t)n_streams * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&b->d_wbguess, (size_t)n_streams * sizeof(float));
    b->opt[HRD_OPT_RX_WBFM_TILING] = 1;
    b->opt[HRD_OPT_RX_WBFM_PACK] = 1;
    if (e == cudaSuccess) e = cudaMalloc(&b->d_ids, (size_t)n_streams * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&b->d_all, (size_t)n_streams * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&b->d_lsb, (size_t)n_streams);
    if (e == cudaSuccess) e = cudaMalloc(&b->d_mode, (size_t)n_streams * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&b->d_kind, (size_t)n_streams);
    if (e == cudaSuccess) e = cudaMalloc(&b->d_sq_n256, (size_t)n_streams * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->ev_report, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&b->d_sq_track, (size_t)n_streams);
    if (e == cudaSuccess) e = cudaMemset(b->d_sq_track, 0, (size_t)n_streams); // SignalTracker: NoSignal
    for (int p = 0; p < HRD_PARAM_COUNT && e == cudaSuccess; p++)
        e = cudaMalloc(&b->d_param[p], (size_t)n_streams * sizeof(float));
    if (e == cudaSuccess) {
        std::vector<int32_t> all((size_t)n_streams);
        for (int s = 0; s < n_streams; s++) all[(size_t)s] = s;
        e = cudaMemcpy(b->d_all, all.data(), (size_t)n_streams * sizeof(int32_t), cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        hrd_destroy(b);
        return fail(HRD_ECUDA, "hrd_create: %s", cudaGetErrorString(e));
    }
    *out = b;
    return HRD_OK;
}

int hrd_destroy(hrd_batch_t *b)
{
    if (!b) return HRD_OK;
    DeviceGuard guard(b->device);
    cudaDeviceSynchronize(); // calls may have named other streams
    cudaFree(b->d_state[0]);
    cudaFree(b->d_state[1]);
    cudaFree(b->d_pre);
    cudaFree(b->d_wbv);
    cudaFree(b->d_fmph);
    cudaFree(b->d_sigph);
    cudaFree(b->d_wbflag);
    cudaFree(b->d_wbrerun);
    cudaFree(b->d_wbguess);
    cudaFree(b->d_wbv2);
    cudaFree(b->d_ids);
    cudaFree(b->d_all);
    cudaFree(b->d_lsb);
    cudaFree(b->d_mode);
    cudaFree(b->d_kind);
    for (int p = 0; p < HRD_PARAM_COUNT; p++) cudaFree(b->d_param[p]);
    cudaFree(b->d_in);
    cudaFree(b->d_out);
    cudaFree(b->d_sq256);
    cudaFree(b->d_sq_mag);
    cudaFree(b->d_sq_open);
    cudaFree(b->d_sq_n256);
    cudaFree(b->d_sq_track);
    if (b->h_sq) cudaFreeHost(b->h_sq);
    if (b->ev_report) cudaEventDestroy(b->ev_report);
    for (int r = 0; r < HRD_PROFILE_RING; r++) {
        for (int i = 0; i < 3; i++)
            if (b->ev[r][i]) cudaEventDestroy(b->ev[r][i]);
        for (int k = 0; k < 4; k++)
            for (int i = 0; i < 2; i++)
                if (b->ev_kind[r][k][i]) cudaEventDestroy(b->ev_kind[r][k][i]);
    }
    for (int i = 0; i < 4; i++) {
        if (b->aux[i]) cudaStreamSynchronize(b->aux[i]), cudaStreamDestroy(b->aux[i]);
        if (b->ev_join[i]) cudaEventDestroy(b->ev_join[i]);
    }
    if (b->ev_fork) cudaEventDestroy(b->ev_fork);
    if (b->ev_last) cudaEventDestroy(b->ev_last);
    for (int h = 0; h < 2; h++) {
        if (b->h_stage[h]) cudaFreeHost(b->h_stage[h]);
        if (b->ev_stage[h]) cudaEventDestroy(b->ev_stage[h]);
    }
    if (b->own) cudaStreamDestroy(b->own);
    delete b;
    return HRD_OK;
}

int hrd_set_mode(hrd_batch_t *b, int stream, int mode)
{
    int rc = check_stream_arg(b, stream);
    if (rc) return rc;
    if (mode < HRD_MODE_NONE || mode > (b->kind == HRD_TX ? HRD_MODE_FM_PROTO : HRD_MODE_USB))
        return fail(HRD_EINVAL, "bad mode %d", mode);
    const int lo = stream == HRD_ALL_STREAMS ? 0 : stream, hi = stream == HRD_ALL_STREAMS ? b->n : stream + 1;
    std::lock_guard<std::mutex> lock(b->ctl);
    for (int s = lo; s < hi; s++) {
        b->mode[(size_t)s] = mode;
        // the sideband flag lives in the SSB object and survives mode changes
        if (mode == HRD_MODE_LSB) b->lsb[(size_t)s] = 1;
        if (mode == HRD_MODE_USB) b->lsb[(size_t)s] = 0;
    }
    b->gen++;
    return HRD_OK;
}

int hrd_get_mode(hrd_batch_t *b, int stream, int *mode)
{
    int rc = check_stream_arg(b, stream);
    if (rc) return rc;
    if (stream == HRD_ALL_STREAMS || !mode) return fail(HRD_EINVAL, "hrd_get_mode needs one stream");
    std::lock_guard<std::mutex> lock(b->ctl);
    *mode = b->mode[(size_t)stream];
    return HRD_OK;
}

int hrd_set_param(hrd_batch_t *b, int stream, int param, float value)
{
    int rc = check_stream_arg(b, stream);
    if (rc) return rc;
    if (param < 0 || param >= HRD_PARAM_COUNT) return fail(HRD_EINVAL, "bad param %d", param);
    const int lo = stream == HRD_ALL_STREAMS ? 0 : stream, hi = stream == HRD_ALL_STREAMS ? b->n : stream + 1;
    std::lock_guard<std::mutex> lock(b->ctl);
    for (int s = lo; s < hi; s++) {
        float &cur = b->param[param][(size_t)s];
        switch (param) {
        case HRD_PARAM_AM_INDEX: // AmModulator.cc:329-339: silently ignored outside 0..1
            if (value >= 0 && value <= 1) cur = value;
            break;
        case HRD_PARAM_FM_DEV: // FmModulator.cc:336-346: the guard reads the member, not the argument
            if (cur >= 0 && cur <= 3500) cur = value;
            break;
        case HRD_PARAM_WBFM_DEV: // WbFmModulator.cc:310-328
            if (cur >= 0 && cur <= 112000) cur = value;
            break;
        case HRD_PARAM_SQUELCH_THRESHOLD: // an int32_t in the reference
        case HRD_PARAM_RX_GAIN_DB:        // a uint32_t
            if (!(value >= -2147483000.f && value <= 2147483000.f) || (param == HRD_PARAM_RX_GAIN_DB && value < 0))
                return fail(HRD_EINVAL, "param %d: %g is not representable in the reference's integer", param, (double)value);
            cur = truncf(value);
            break;
        default:
            cur = value;
        }
    }
    b->gen++;
    return HRD_OK;
}

int hrd_get_param(hrd_batch_t *b, int stream, int param, float *value)
{
    int rc = check_stream_arg(b, stream);
    if (rc) return rc;
    if (stream == HRD_ALL_STREAMS || !value || param < 0 || param >= HRD_PARAM_COUNT)
        return fail(HRD_EINVAL, "hrd_get_param needs one stream and a valid param");
    std::lock_guard<std::mutex> lock(b->ctl);
    *value = b->param[param][(size_t)stream];
    return HRD_OK;
}

int hrd_set_option(hrd_batch_t *b, int option, int value)
{
    if (!b) return fail(HRD_EINVAL, "null batch");
    if (option < 0 || option >= HRD_OPT_COUNT) return fail(HRD_EINVAL, "bad option %d", option);
    if (value < 0) return fail(HRD_EINVAL, "option values are non-negative");
    if (option == HRD_OPT_RX_SQUELCH_BLOCK && value % 512)
        return fail(HRD_EINVAL, "squelch block of %d bytes is not a whole number of PCM samples (512 bytes)", value);
    std::lock_guard<std::mutex> lock(b->ctl);
    b->opt[option] = value;
    if (option == HRD_OPT_RX_SQUELCH) b->gen++; // part of the "armed" snapshot
    return HRD_OK;
}

int hrd_get_option(hrd_batch_t *b, int option, int *value)
{
    if (!b || !value) return fail(HRD_EINVAL, "null argument");
    if (option < 0 || option >= HRD_OPT_COUNT) return fail(HRD_EINVAL, "bad option %d", option);
    *value = b->opt[option];
    return HRD_OK;
}

int hrd_reset(hrd_batch_t *b, int stream, int unit)
{
    int rc = check_stream_arg(b, stream);
    if (rc) return rc;
    DeviceGuard guard(b->device);
    // stream-ordered like everything else: behind the call before, in front of the call after; nothing waits
    rc = order_after_last(b, b->own);
    if (rc) return rc;
    using hrd::RxState;
    using hrd::TxState;
#define RANGE(T, first, next) offsetof(T, first), offsetof(T, next) - offsetof(T, first)
    if (b->kind == HRD_RX) {
        switch (unit) {
        case HRD_UNIT_AM: rc = zero_state(b, stream, RANGE(RxState, am, fm_r256)); break;
        case HRD_UNIT_FM: rc = zero_state(b, stream, RANGE(RxState, fm_r256, wb_prev_theta)); break;
        case HRD_UNIT_WBFM: // WbFmDemodulator.cc:265-297: decimators and previousTheta, not the IIR
            rc = zero_state(b, stream, RANGE(RxState, wb_prev_theta, wb_x1));
            if (!rc) rc = zero_state(b, stream, RANGE(RxState, wb_d256, ssb));
            break;
        case HRD_UNIT_SSB:
            rc = zero_state(b, stream, offsetof(RxState, ssb), sizeof(RxState) - offsetof(RxState, ssb));
            break;
        case HRD_UNIT_FRONT_END: rc = zero_state(b, stream, RANGE(RxState, fe_t, am)); break;
        case HRD_UNIT_ALL:
            rc = zero_state(b, stream, 0, sizeof(RxState));
            // a freshly constructed IqDataProcessor also has a fresh SignalTracker (state NoSignal)
            if (!rc) {
                if (stream == HRD_ALL_STREAMS)
                    HRD_CUDA(cudaMemsetAsync(b->d_sq_track, 0, (size_t)b->n, b->own));
                else
                    HRD_CUDA(cudaMemsetAsync(b->d_sq_track + stream, 0, 1, b->own));
            }
            break;
        default: return fail(HRD_EINVAL, "bad unit %d", unit);
        }
    } else {
        switch (unit) {
        case HRD_UNIT_AM: rc = zero_state(b, stream, RANGE(TxState, am, fm)); break;
        case HRD_UNIT_FM: rc = zero_state(b, stream, RANGE(TxState, fm, ssb)); break; // phase kept
        case HRD_UNIT_SSB:
            rc = zero_state(b, stream, RANGE(TxState, ssb, wb));
            if (!rc) rc = zero_state(b, stream, RANGE(TxState, ssb_h8, pad));
            break;
        case HRD_UNIT_WBFM: rc = zero_state(b, stream, RANGE(TxState, wb, fm_phase)); break; // phase kept
        case HRD_UNIT_SIGNALS: rc = zero_state(b, stream, offsetof(TxState, sig), sizeof(hrd::TxRail8) + sizeof(float)); break;
        case HRD_UNIT_ALL: rc = zero_state(b, stream, 0, sizeof(TxState)); break;
        default: return fail(HRD_EINVAL, "bad unit %d", unit);
        }
    }
#undef RANGE
    if (rc) return rc;
    return mark_end(b, b->own);
}

int hrd_synchronize(hrd_batch_t *b)
{
    if (!b) return fail(HRD_EINVAL, "null batch");
    DeviceGuard guard(b->device);
    HRD_CUDA(cudaDeviceSynchronize());
    return HRD_OK;
}

int hrd_launch_count(hrd_batch_t *b, uint64_t *count)
{
    if (!b || !count) return fail(HRD_EINVAL, "null argument");
    *count = b->launches;
    return HRD_OK;
}

int hrd_wbfm_fallback_count(hrd_batch_t *b, uint64_t *count)
{
    if (!b || !count) return fail(HRD_EINVAL, "null argument");
    DeviceGuard guard(b->device);
    unsigned long long v = 0;
    HRD_CUDA(cudaDeviceSynchronize());
    HRD_CUDA(cudaMemcpy(&v, b->d_wbflag + 2, sizeof v, cudaMemcpyDeviceToHost));
    *count = (uint64_t)v;
    return HRD_OK;
}

int hrd_get_device(hrd_batch_t *b, int *device)
{
    if (!b || !device) return fail(HRD_EINVAL, "null argument");
    *device = b->device;
    return HRD_OK;
}

int hrd_wbfm_serial_count(hrd_batch_t *b, uint64_t *count)
{
    if (!b || !count) return fail(HRD_EINVAL, "null argument");
    DeviceGuard guard(b->device);
    unsigned long long v = 0;
    HRD_CUDA(cudaDeviceSynchronize());
    HRD_CUDA(cudaMemcpy(&v, b->d_wbflag + 4, sizeof v, cudaMemcpyDeviceToHost));
    *count = (uint64_t)v;
    return HRD_OK;
}

int hrd_kernel_ms(hrd_batch_t *b, int which, int age, float *ms)
{
    if (!b || !ms) return fail(HRD_EINVAL, "null argument");
    if (!(which == 0 || which == 1 || (which >= 10 && which < 14)))
        return fail(HRD_EINVAL, "which must be 0 (main kernels), 1 (tail kernel) or 10 + kind");
    if (age < 0 || age >= HRD_PROFILE_RING || (uint64_t)age >= b->ev_calls)
        return fail(HRD_EINVAL, "no profiled call of age %d: set HRD_OPT_PROFILE before the process calls", age);
    DeviceGuard guard(b->device);
    const size_t slot = (size_t)((b->ev_calls - 1 - (uint64_t)age) % HRD_PROFILE_RING);
    cudaEvent_t *e = b->ev[slot];
    HRD_CUDA(cudaEventSynchronize(e[2]));
    if (which >= 10) {
        if (!b->ev_kind_set[slot][which - 10])
            return fail(HRD_EINVAL, "kind %d was not timed in that call (HRD_OPT_RX_SERIAL + HRD_OPT_PROFILE, and the batch must hold it)", which - 10);
        HRD_CUDA(cudaEventElapsedTime(ms, b->ev_kind[slot][which - 10][0], b->ev_kind[slot][which - 10][1]));
        return HRD_OK;
    }
    HRD_CUDA(cudaEventElapsedTime(ms, e[which], e[which + 1]));
    return HRD_OK;
}

int hrd_get_table(hrd_batch_t *b, int which, float *out, size_t n)
{
    if (!b || !out) return fail(HRD_EINVAL, "null argument");
    DeviceGuard guard(b->device);
    const DeviceTables &d = g_dev_tables[b->device];
    const float *src = which == 0 ? d.atan2_lut : which == 1 ? d.nco_sin : which == 2 ? d.nco_cos : nullptr;
    const size_t have = which == 0 ? 65536 : 16384;
    if (!src || n > have) return fail(HRD_EINVAL, "bad table request");
    HRD_CUDA(cudaMemcpy(out, src, n * sizeof(float), cudaMemcpyDeviceToHost));
    return HRD_OK;
}

// The demodulator launches of one call (or, with the squelch armed, of one block): one tile-kernel launch
// per kernel kind over the given stream lists, then the AM/SSB recurrence pass.  after_tiles, when set, is
// recorded between the two (HRD_OPT_PROFILE).
// A batch that holds several kinds (BASELINE config 5, mixed-mode sweeps) FANS OUT: every kind's launches go
// to a stream of their own between a fork and a join event on the caller's stream, so the kernels fill each
// other's last waves and the latency-bound recurrence pass of the AM/SSB streams runs beside the other
// modes' tile kernels.  The kinds touch disjoint streams' records and disjoint scratch buffers.
static int run_demods(hrd_batch_t *b, hrd::RxParams &p, int entry, uint32_t n_batches, const int32_t *const ids_of[4],
                      const int cnt_of[4], cudaStream_t s, cudaEvent_t after_tiles, int prof_slot = -1, bool one_tile = false)
{
    static const int gain_of_kind[5] = {-1, HRD_PARAM_AM_GAIN, HRD_PARAM_FM_GAIN, HRD_PARAM_WBFM_GAIN, HRD_PARAM_SSB_GAIN};
    int rc;
    hrd::RxParams iir_p;
    bool iir = false;
    int kinds = 0;
    for (int k = 0; k < 4; k++) kinds += cnt_of[k] && !(k == hrd::K_NONE && entry == HRD_ENTRY_256K);
    const bool serial = b->opt[HRD_OPT_RX_SERIAL] != 0;
    const bool fan = kinds >= 2 && !serial;
    if (fan) HRD_CUDA(cudaEventRecord(b->ev_fork, s));
    if (prof_slot >= 0)
        for (int k = 0; k < 4; k++) b->ev_kind_set[prof_slot][k] = false;
    int lane = 0;
    // WBFM first: its launch is the longest and may end in a small exact re-run (a handful of CTAs walking whole calls
    // serially); started first, that tail runs beside the other kinds' kernels instead of behind them
    static const int launch_order[4] = {hrd::K_WBFM, hrd::K_FM, hrd::K_AM, hrd::K_NONE};
    for (int o = 0; o < 4; o++) {
        const int k = launch_order[o];
        const int cnt = cnt_of[k];
        if (!cnt) continue;
        if (k == hrd::K_NONE && entry == HRD_ENTRY_256K) continue;
        cudaStream_t ks = s;
        if (fan) {
            ks = b->aux[lane];
            HRD_CUDA(cudaStreamWaitEvent(ks, b->ev_fork, 0));
        }
        const bool time_kind = serial && prof_slot >= 0;
        if (time_kind) {
            for (int i = 0; i < 2; i++)
                if (!b->ev_kind[prof_slot][k][i]) HRD_CUDA(cudaEventCreate(&b->ev_kind[prof_slot][k][i]));
            HRD_CUDA(cudaEventRecord(b->ev_kind[prof_slot][k][0], ks));
        }
        WbPart parts[2];
        int n_parts = 1;
        parts[0].first = 0;
        parts[0].count = cnt;
        if (k == hrd::K_WBFM) n_parts = plan_wbfm(b, entry, cnt, n_batches, one_tile, parts);
        for (int part = 0; part < n_parts; part++) {
            p.stream_ids = ids_of[k] + parts[part].first;
            p.n_streams = parts[part].count;
            p.gain = k ? b->d_param[gain_of_kind[k]] : nullptr;
            if (k == hrd::K_WBFM) {
                p.n_tiles = parts[part].n_tiles;
                p.tile_batches = parts[part].tile_batches;
            } else {
                // (ragged calls too: tiled over the nominal length, hrd_rx.cu rx_kernel)
                choose_tiles(b, k, entry, p.n_streams, n_batches, &p.n_tiles, &p.tile_batches);
            }
        p.wb_pack = fan && b->opt[HRD_OPT_RX_WBFM_PACK] ? 1 : 0; // beside other kinds: full WBFM CTAs on fewer SMs (hrd_tables.h)
        const bool speculate = k == hrd::K_WBFM && p.n_tiles > 1;
            if (speculate) { // verified speculation (hrd_rx.cu): pairs to compare, this call's flag
                rc = ensure_cap(&b->d_wbv, &b->d_wbv_cap, sizeof(float2) * (size_t)p.n_streams * (size_t)p.n_tiles);
                if (rc) return rc;
                p.wb_verify = (float2 *)b->d_wbv;
                HRD_CUDA(cudaMemsetAsync(b->d_wbflag, 0, 2 * sizeof(uint32_t), ks));
            }
            int e = hrd::launch_rx(k, entry, p, ks);
            if (e) return fail(HRD_ECUDA, "rx launch (kind %d) failed: %s", k, cudaGetErrorString((cudaError_t)e));
            b->launches++;
            if (speculate) {
                const int force = b->opt[HRD_OPT_DEBUG_WBFM_FORCE_RERUN];
                // (short tiles: the retry is a handful of streams, its time is the LENGTH of a tile)
            const uint32_t tb2 = std::max<uint32_t>(3u, (n_batches + 31) / 32);
                const int nt2 = (int)((n_batches + tb2 - 1) / tb2);
                e = hrd::launch_rx_wbfm_verify(p, nullptr, 0, b->d_wbflag, b->d_wbrerun, b->d_wbguess, (unsigned long long *)(b->d_wbflag + 2),
                                               nt2 >= 2 ? nullptr : (unsigned long long *)(b->d_wbflag + 4), force, ks);
                if (e) return fail(HRD_ECUDA, "wbfm verify launch failed: %s", cudaGetErrorString((cudaError_t)e));
                // SECOND PASS for the streams that failed: tiled again (finer, they are few), every tile >= 1 taking the
                // TRUE value at the first check point as its recurrence value.  What fails in practice is a constant
                // input: the recurrence then sits on one of several neighbouring fixed points of its rounded map and
                // stays there, the warm-up from zero reaches another one, and no amount of warm-up brings them
                // together -- but the value tile 0 saw is the value at every later check point as well.  The retry is
                // verified like the first pass (each tile's value against the one the tile before it arrives at), so a
                // wrong guess costs time only: what fails again is walked serially by the third launch.
                hrd::RxParams again = p;
                again.wb_verify = nullptr;
                again.run_if = b->d_wbflag;
                again.rerun_ids = b->d_wbrerun;
                if (nt2 >= 2) { // (too short a call for two tiles: straight to the serial run)
                    hrd::RxParams retry = again;
                    retry.n_streams = std::min(p.n_streams, 512); // the part of the list it is sized for
                    retry.n_tiles = nt2;
                    retry.tile_batches = tb2;
                    rc = ensure_cap(&b->d_wbv2, &b->d_wbv2_cap, sizeof(float2) * (size_t)retry.n_streams * (size_t)nt2);
                    if (rc) return rc;
                    retry.wb_verify = (float2 *)b->d_wbv2;
                    retry.wb_guess = b->d_wbguess;
                    e = hrd::launch_rx(k, entry, retry, ks);
                    if (e) return fail(HRD_ECUDA, "wbfm retry launch failed: %s", cudaGetErrorString((cudaError_t)e));
                    e = hrd::launch_rx_wbfm_verify(retry, b->d_wbflag, p.n_streams, b->d_wbflag + 1, b->d_wbrerun + b->n, nullptr,
                                                   (unsigned long long *)(b->d_wbflag + 4), nullptr, force >= 2, ks);
                    if (e) return fail(HRD_ECUDA, "wbfm retry verify launch failed: %s", cudaGetErrorString((cudaError_t)e));
                    b->launches += 2;
                    again.run_if = b->d_wbflag + 1;
                    again.rerun_ids = b->d_wbrerun + b->n;
                }
                // the exact untiled run, from the untouched state_in; runs only for the streams still listed
                again.n_tiles = 1;
                again.tile_batches = n_batches;
                e = hrd::launch_rx(k, entry, again, ks);
                if (e) return fail(HRD_ECUDA, "wbfm re-run launch failed: %s", cudaGetErrorString((cudaError_t)e));
                b->launches += 2;
                p.wb_verify = nullptr;
            }
        }
        p.stream_ids = ids_of[k];
        p.n_streams = cnt;
        if (k == hrd::K_AM) {
            iir_p = p;
            iir = true;
            if (fan) { // the recurrence pass follows its tile kernel on the same stream, beside the other kinds
                int e = hrd::launch_rx_dc_iir(iir_p, ks);
                if (e) return fail(HRD_ECUDA, "rx IIR launch failed: %s", cudaGetErrorString((cudaError_t)e));
                b->launches++;
                iir = false;
            }
        }
        if (time_kind) {
            HRD_CUDA(cudaEventRecord(b->ev_kind[prof_slot][k][1], ks));
            b->ev_kind_set[prof_slot][k] = true;
        }
        if (fan) {
            HRD_CUDA(cudaEventRecord(b->ev_join[lane], ks));
            HRD_CUDA(cudaStreamWaitEvent(s, b->ev_join[lane], 0));
            lane++;
        }
    }
    if (after_tiles) HRD_CUDA(cudaEventRecord(after_tiles, s));
    if (iir) { // the serial 8 kS/s recurrence of the AM/SSB streams, after every tile kernel
        int e = hrd::launch_rx_dc_iir(iir_p, s);
        if (e) return fail(HRD_ECUDA, "rx IIR launch failed: %s", cudaGetErrorString((cudaError_t)e));
        b->launches++;
    }
    return HRD_OK;
}

// The squelched receive call (include/hrd.h "Squelch"; hrd_rx.cu rx_gate_kernel): ONE kernel runs the front end of
// every stream, the per-block magnitudes, the tracker, and leaves each stream's open blocks packed in a 256 kS/s
// scratch row; the demodulator kernels then take those rows as one ragged call at the 256 kS/s entry and write
// their PCM packed into the caller's rows.  Nothing waits for the host: decisions, magnitudes and PCM counts
// travel back through pinned staging and a host function on the stream.  p arrives with iq / tables set.
struct GateReport {
    hrd_batch *b;
    size_t n, cells;
    uint32_t n_blocks;
    uint32_t *pcm_counts; // the caller's array (may be null)
};
static void CUDART_CB gate_report_landed(void *user)
{
    GateReport *r = (GateReport *)user;
    hrd_batch *b = r->b;
    const uint32_t *mag = (const uint32_t *)b->h_sq;
    const uint32_t *n256 = mag + r->cells;
    const uint8_t *open = (const uint8_t *)(n256 + r->n);
    b->sq_mag.assign(mag, mag + r->cells);
    b->sq_open.assign(open, open + r->cells);
    b->sq_blocks = r->n_blocks;
    if (r->pcm_counts)
        for (size_t i = 0; i < r->n; i++) r->pcm_counts[i] = n256[i] / 32;
    delete r;
}

static int rx_gated(hrd_batch_t *b, hrd::RxParams p, size_t n256, uint32_t *pcm_counts, cudaStream_t s)
{
    // one reference call: 262144 bytes at 2.048 MS/s (hackRf/hackrf.c:101) = 16384 samples at 256 kS/s
    const size_t blk256 = b->opt[HRD_OPT_RX_SQUELCH_BLOCK] > 0 ? (size_t)b->opt[HRD_OPT_RX_SQUELCH_BLOCK] / 16 : 16384;
    const int n_blocks = (int)((n256 + blk256 - 1) / blk256);
    const size_t n = (size_t)b->n, cells = n * (size_t)n_blocks;
    const size_t row256 = (n256 * 2 + 31) & ~(size_t)31;
    int rc = ensure_cap(&b->d_sq256, &b->d_sq256_cap, row256 * n);
    if (!rc) rc = ensure_cap(&b->d_sq_mag, &b->d_sq_mag_cap, cells * sizeof(uint32_t));
    if (!rc) rc = ensure_cap(&b->d_sq_open, &b->d_sq_open_cap, cells);
    if (rc) return rc;
    const size_t report_bytes = cells * sizeof(uint32_t) + n * sizeof(uint32_t) + cells;
    HRD_CUDA(cudaEventSynchronize(b->ev_report)); // the report of the call before has been unpacked
    if (b->h_sq_cap < report_bytes) {
        if (b->h_sq) cudaFreeHost(b->h_sq);
        b->h_sq = nullptr, b->h_sq_cap = 0;
        HRD_CUDA(cudaHostAlloc(&b->h_sq, report_bytes, cudaHostAllocDefault));
        b->h_sq_cap = report_bytes;
    }

    // 1. IqDataProcessor::reduceSampleRate + upconvertByFsOver4 (:937-946) and Squelch::run (:961) for every stream
    hrd::GateParams g;
    g.scratch = (int8_t *)b->d_sq256;
    g.row256 = row256;
    g.blk256 = (uint32_t)blk256;
    g.n_blocks = n_blocks;
    g.threshold = b->d_param[HRD_PARAM_SQUELCH_THRESHOLD];
    g.gain_db = b->d_param[HRD_PARAM_RX_GAIN_DB];
    g.tracking = b->d_sq_track;
    g.magnitude = (uint32_t *)b->d_sq_mag;
    g.allowed = (uint8_t *)b->d_sq_open;
    g.n256_open = (uint32_t *)b->d_sq_n256;
    {
        hrd::RxParams fe = p;
        fe.stream_ids = b->d_all;
        fe.n_streams = b->n;
        fe.kind_of = b->d_kind;
        fe.state_in = (const hrd::RxState *)b->d_state[b->cur];
        fe.state_out = (hrd::RxState *)b->d_state[b->cur ^ 1];
        if (hrd::launch_rx_gate(fe, g, s)) return fail(HRD_ECUDA, "gate launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        b->launches++;
        b->cur ^= 1;
    }
    // 2. what the gate decided, on its way to the host -- on a side stream, so that the demodulators below do not
    // queue behind a host function; the caller's stream joins it at the end of the call
    {
        cudaStream_t rs = b->aux[3];
        HRD_CUDA(cudaEventRecord(b->ev_fork, s));
        HRD_CUDA(cudaStreamWaitEvent(rs, b->ev_fork, 0));
        char *h = (char *)b->h_sq;
        HRD_CUDA(cudaMemcpyAsync(h, b->d_sq_mag, cells * sizeof(uint32_t), cudaMemcpyDeviceToHost, rs));
        HRD_CUDA(cudaMemcpyAsync(h + cells * sizeof(uint32_t), b->d_sq_n256, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, rs));
        HRD_CUDA(cudaMemcpyAsync(h + (cells + n) * sizeof(uint32_t), b->d_sq_open, cells, cudaMemcpyDeviceToHost, rs));
        GateReport *r = new (std::nothrow) GateReport{b, n, cells, (uint32_t)n_blocks, pcm_counts};
        if (!r) return fail(HRD_ENOMEM, "out of host memory");
        HRD_CUDA(cudaLaunchHostFunc(rs, gate_report_landed, r));
        HRD_CUDA(cudaEventRecord(b->ev_report, rs));
    }
    // 3. the demodulators over every stream's open blocks: one ragged call at the 256 kS/s entry
    hrd::RxParams q = p;
    q.iq = (const int8_t *)b->d_sq256;
    q.iq_stride = row256;
    q.n256 = (uint32_t)n256;
    q.n256_of = (const uint32_t *)b->d_sq_n256;
    q.state_in = (const hrd::RxState *)b->d_state[b->cur];
    q.state_out = (hrd::RxState *)b->d_state[b->cur ^ 1];
    q.kind_of = b->d_kind;
    q.gain_ssb = b->d_param[HRD_PARAM_SSB_GAIN];
    if (b->group_cnt[hrd::K_NONE]) // no demodulator selected: nothing runs for those streams, their records carry over
        HRD_CUDA(cudaMemcpyAsync(b->d_state[b->cur ^ 1], b->d_state[b->cur], sizeof(hrd::RxState) * n, cudaMemcpyDeviceToDevice, s));
    const int32_t *ids_of[4];
    int cnt_of[4];
    for (int k = 0; k < 4; k++) {
        ids_of[k] = b->d_ids + b->group_off[k];
        cnt_of[k] = k ? b->group_cnt[k] + (k == hrd::K_AM ? b->group_cnt[hrd::K_SSB] : 0) : 0;
    }
    rc = run_demods(b, q, HRD_ENTRY_256K, (uint32_t)((n256 + 1023) / 1024), ids_of, cnt_of, s, nullptr, -1, true);
    if (rc) return rc;
    HRD_CUDA(cudaStreamWaitEvent(s, b->ev_report, 0)); // "after synchronising cuda_stream" covers the report too
    b->cur ^= 1;
    b->gated_calls++;
    return HRD_OK;
}

static int rx_common(hrd_batch_t *b, const int8_t *iq, size_t bytes, size_t iq_stride, int entry, int16_t *pcm,
                     size_t pcm_stride, int8_t *out256, size_t out_stride, uint32_t *pcm_counts, int mem,
                     void *cuda_stream, bool front_end_only)
{
    if (!b || b->kind != HRD_RX) return fail(HRD_EINVAL, "not an Rx batch");
    if (!iq) return fail(HRD_EINVAL, "iq is null");
    if (entry != HRD_ENTRY_2048K && entry != HRD_ENTRY_256K) return fail(HRD_EINVAL, "bad entry %d", entry);
    const size_t unit = entry == HRD_ENTRY_2048K ? 512 : 64;
    if (bytes % unit) return fail(HRD_EINVAL, "bytes_per_stream %zu is not a multiple of %zu", bytes, unit);
    if (mem != HRD_MEM_HOST && mem != HRD_MEM_DEVICE) return fail(HRD_EINVAL, "bad mem %d", mem);
    const size_t n256 = entry == HRD_ENTRY_2048K ? bytes / 16 : bytes / 2;
    const size_t npcm = n256 / 32;
    // the kernels address a stream's input with 32-bit byte offsets
    if (bytes >= ((size_t)1 << 32)) return fail(HRD_EINVAL, "bytes_per_stream must be below 4 GiB per call");
    if (iq_stride < bytes) return fail(HRD_EINVAL, "iq_stride smaller than bytes_per_stream");
    if (!front_end_only && !pcm) return fail(HRD_EINVAL, "pcm is null");
    if (!front_end_only && pcm_stride < npcm) return fail(HRD_EINVAL, "pcm_stride too small");
    DeviceGuard guard(b->device);
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : (mem == HRD_MEM_HOST ? b->own : (cudaStream_t) nullptr);
    int rc = order_after_last(b, s);
    if (!rc) rc = regroup(b, s);
    if (rc) return rc;
    if (pcm_counts)
        for (int i = 0; i < b->n; i++) pcm_counts[i] = b->h_kind[(size_t)i] == hrd::K_NONE ? 0u : (uint32_t)npcm;
    if (bytes == 0) return HRD_OK;

    const int8_t *d_iq = iq;
    int16_t *d_pcm = pcm;
    int8_t *d_o256 = out256;
    size_t d_iq_stride = iq_stride, d_pcm_stride = pcm_stride, d_o_stride = out_stride;
    const size_t out_row = front_end_only ? n256 * 2 : npcm * sizeof(int16_t);
    if (mem == HRD_MEM_HOST) {
        const size_t in_row = (bytes + 31) & ~(size_t)31;
        rc = ensure_cap(&b->d_in, &b->d_in_cap, in_row * (size_t)b->n);
        if (!rc) rc = ensure_cap(&b->d_out, &b->d_out_cap, ((out_row + 31) & ~(size_t)31) * (size_t)b->n);
        if (rc) return rc;
        HRD_CUDA(cudaMemcpy2DAsync(b->d_in, in_row, iq, iq_stride, bytes, (size_t)b->n, cudaMemcpyHostToDevice, s));
        d_iq = (const int8_t *)b->d_in;
        d_iq_stride = in_row;
        d_pcm = (int16_t *)b->d_out;
        d_pcm_stride = ((out_row + 31) & ~(size_t)31) / sizeof(int16_t);
        d_o256 = (int8_t *)b->d_out;
        d_o_stride = (out_row + 31) & ~(size_t)31;
    } else {
        const size_t align = entry == HRD_ENTRY_2048K ? 32 : 4;
        if (((uintptr_t)iq % align) || (iq_stride % align))
            return fail(HRD_EINVAL, "device iq pointer/stride must be %zu-byte aligned", align);
        if (front_end_only && (((uintptr_t)out256 % 4) || (out_stride % 4)))
            return fail(HRD_EINVAL, "out256 pointer/stride must be 4-byte aligned");
    }

    hrd::RxParams p;
    memset(&p, 0, sizeof p);
    p.iq = d_iq;
    p.iq_stride = d_iq_stride;
    p.n256 = (uint32_t)n256;
    p.pcm = d_pcm;
    p.pcm_stride = d_pcm_stride;
    p.state_in = (const hrd::RxState *)b->d_state[b->cur];
    p.state_out = (hrd::RxState *)b->d_state[b->cur ^ 1];
    p.lsb = b->d_lsb;
    p.atan2_lut = g_dev_tables[b->device].atan2_lut;
    p.sm_count = b->sm_count;
    const uint32_t n_batches = (uint32_t)((n256 + 1023) / 1024);
    bool gated = false;
    if (!front_end_only && entry == HRD_ENTRY_2048K && b->armed) {
        if (mem == HRD_MEM_HOST) HRD_CUDA(cudaMemsetAsync(b->d_out, 0, d_pcm_stride * sizeof(int16_t) * (size_t)b->n, s));
        if (b->group_cnt[hrd::K_AM] || b->group_cnt[hrd::K_SSB]) {
            const size_t stride = (npcm + 7) & ~(size_t)7;
            if (stride * (size_t)b->n >= ((size_t)1 << 32))
                return fail(HRD_EINVAL, "call too long for %d AM/SSB streams (IIR scratch is indexed with 32 bits)", b->n);
            if (b->pre_stride < stride) {
                rc = ensure_cap((void **)&b->d_pre, &b->d_pre_cap, stride * sizeof(float) * (size_t)b->n);
                if (rc) return rc;
                b->pre_stride = stride;
            }
            p.pre_iir = b->d_pre;
            p.pre_stride = b->pre_stride;
        }
        rc = rx_gated(b, p, n256, pcm_counts, s);
        if (rc) return rc;
        gated = true;
    } else if (front_end_only) {
        p.out256 = d_o256;
        p.out_stride = d_o_stride;
        p.stream_ids = b->d_all;
        p.n_streams = b->n;
        choose_tiles(b, hrd::K_NONE, HRD_ENTRY_2048K, b->n, n_batches, &p.n_tiles, &p.tile_batches);
        if (hrd::launch_rx(hrd::K_NONE, HRD_ENTRY_2048K, p, s)) return fail(HRD_ECUDA, "front-end launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        b->launches++;
    } else {
        if (b->group_cnt[hrd::K_AM] || b->group_cnt[hrd::K_SSB]) {
            // the DC-removal IIR's input, one float per PCM sample (rows 32-byte aligned)
            const size_t stride = (npcm + 7) & ~(size_t)7;
            if (stride * (size_t)b->n >= ((size_t)1 << 32))
                return fail(HRD_EINVAL, "call too long for %d AM/SSB streams (IIR scratch is indexed with 32 bits)", b->n);
            if (b->pre_stride < stride) {
                rc = ensure_cap((void **)&b->d_pre, &b->d_pre_cap, stride * sizeof(float) * (size_t)b->n);
                if (rc) return rc;
                b->pre_stride = stride;
            }
            p.pre_iir = b->d_pre;
            p.pre_stride = b->pre_stride;
        }
        if (b->group_cnt[hrd::K_NONE] && entry == HRD_ENTRY_256K)
            // no demodulator selected: nothing runs for those streams, their state just carries over
            HRD_CUDA(cudaMemcpyAsync(b->d_state[b->cur ^ 1], b->d_state[b->cur], sizeof(hrd::RxState) * (size_t)b->n,
                                     cudaMemcpyDeviceToDevice, s));
        p.kind_of = b->d_kind;
        p.gain_ssb = b->d_param[HRD_PARAM_SSB_GAIN];
        // Squelch::run still runs on every block of the reference (IqDataProcessor.cc:961); with no threshold able
        // to close the gate every block counts as present, so the tracker of every stream sits in Tracking --
        // which matters the moment a caller raises a threshold: the first quiet block is then the tail
        // (ENDOFSIGNAL) and still passes
        if (entry == HRD_ENTRY_2048K) HRD_CUDA(cudaMemsetAsync(b->d_sq_track, 1, (size_t)b->n, s));
        const bool prof = b->opt[HRD_OPT_PROFILE] != 0;
        cudaEvent_t *ev = b->ev[b->ev_calls % HRD_PROFILE_RING];
        if (prof) {
            for (int i = 0; i < 3; i++)
                if (!ev[i]) HRD_CUDA(cudaEventCreate(&ev[i]));
            HRD_CUDA(cudaEventRecord(ev[0], s));
        }
        const int32_t *ids_of[4];
        int cnt_of[4];
        for (int k = 0; k < 4; k++) { // K_NONE, K_AM (+ K_SSB: adjacent in d_ids, one launch), K_FM, K_WBFM
            ids_of[k] = b->d_ids + b->group_off[k];
            cnt_of[k] = b->group_cnt[k] + (k == hrd::K_AM ? b->group_cnt[hrd::K_SSB] : 0);
        }
        rc = run_demods(b, p, entry, n_batches, ids_of, cnt_of, s, prof ? ev[1] : nullptr,
                        prof ? (int)(b->ev_calls % HRD_PROFILE_RING) : -1);
        if (rc) return rc;
        if (prof) {
            HRD_CUDA(cudaEventRecord(ev[2], s));
            b->ev_calls++;
        }
    }
    if (!gated) { // what this call wrote is what the next one reads (rx_gated swaps twice itself)
        b->cur ^= 1;
        // a call that was not gated has no report: behind the last gated call's, so the two cannot cross
        if (b->gated_calls) {
            HRD_CUDA(cudaEventSynchronize(b->ev_report));
            b->sq_blocks = 0;
        }
    }
    if (mem == HRD_MEM_HOST) {
        if (front_end_only)
            HRD_CUDA(cudaMemcpy2DAsync(out256, out_stride, b->d_out, d_o_stride, out_row, (size_t)b->n,
                                       cudaMemcpyDeviceToHost, s));
        else if (npcm)
            HRD_CUDA(cudaMemcpy2DAsync(pcm, pcm_stride * sizeof(int16_t), b->d_out, d_pcm_stride * sizeof(int16_t),
                                       out_row, (size_t)b->n, cudaMemcpyDeviceToHost, s));
        HRD_CUDA(cudaStreamSynchronize(s));
    }
    return mark_end(b, s);
}

int hrd_rx_process(hrd_batch_t *b, const int8_t *iq, size_t bytes_per_stream, size_t iq_stride, int entry,
                   int16_t *pcm, size_t pcm_stride, uint32_t *pcm_counts, int mem, void *cuda_stream)
{
    return rx_common(b, iq, bytes_per_stream, iq_stride, entry, pcm, pcm_stride, nullptr, 0, pcm_counts, mem,
                     cuda_stream, false);
}

int hrd_rx_front_end(hrd_batch_t *b, const int8_t *iq, size_t bytes_per_stream, size_t iq_stride, int8_t *out256k,
                     size_t out_stride, int mem, void *cuda_stream)
{
    if (!out256k) return fail(HRD_EINVAL, "out256k is null");
    if (out_stride < bytes_per_stream / 8) return fail(HRD_EINVAL, "out_stride too small");
    return rx_common(b, iq, bytes_per_stream, iq_stride, HRD_ENTRY_2048K, nullptr, 0, out256k, out_stride, nullptr,
                     mem, cuda_stream, true);
}

int hrd_rx_fs4_rotate(hrd_batch_t *b, int8_t *iq, size_t bytes, int up, int mem, void *cuda_stream)
{
    if (!b || b->kind != HRD_RX) return fail(HRD_EINVAL, "not an Rx batch");
    if (!iq || bytes % 8) return fail(HRD_EINVAL, "iq is null or bytes is not a multiple of 8 (four I,Q samples)");
    if (mem != HRD_MEM_HOST && mem != HRD_MEM_DEVICE) return fail(HRD_EINVAL, "bad mem %d", mem);
    if (!bytes) return HRD_OK;
    DeviceGuard guard(b->device);
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : (mem == HRD_MEM_HOST ? b->own : (cudaStream_t) nullptr);
    int orc = order_after_last(b, s);
    if (orc) return orc;
    int8_t *d = iq;
    if (mem == HRD_MEM_HOST) {
        int rc = ensure_cap(&b->d_in, &b->d_in_cap, bytes);
        if (rc) return rc;
        d = (int8_t *)b->d_in;
        HRD_CUDA(cudaMemcpyAsync(d, iq, bytes, cudaMemcpyHostToDevice, s));
    } else if ((uintptr_t)iq % 4) {
        return fail(HRD_EINVAL, "device iq pointer must be 4-byte aligned");
    }
    if (hrd::launch_fs4_rotate(d, bytes / 8, up != 0, s)) return fail(HRD_ECUDA, "rotate launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    b->launches++;
    if (mem == HRD_MEM_HOST) {
        HRD_CUDA(cudaMemcpyAsync(iq, d, bytes, cudaMemcpyDeviceToHost, s));
        HRD_CUDA(cudaStreamSynchronize(s));
    }
    return mark_end(b, s);
}

int hrd_rx_squelch_report(hrd_batch_t *b, uint32_t *magnitudes, uint8_t *allowed, size_t blocks_cap, uint32_t *n_blocks)
{
    if (!b || b->kind != HRD_RX) return fail(HRD_EINVAL, "not an Rx batch");
    {
        DeviceGuard guard(b->device);
        HRD_CUDA(cudaEventSynchronize(b->ev_report)); // the latest gated call's report has been unpacked
    }
    if (n_blocks) *n_blocks = b->sq_blocks;
    if (b->sq_blocks > blocks_cap && (magnitudes || allowed)) return fail(HRD_EINVAL, "blocks_cap %zu < %u blocks", blocks_cap, b->sq_blocks);
    for (size_t st = 0; st < (size_t)b->n; st++)
        for (uint32_t k = 0; k < b->sq_blocks; k++) {
            if (magnitudes) magnitudes[st * blocks_cap + k] = b->sq_mag[st * b->sq_blocks + k];
            if (allowed) allowed[st * blocks_cap + k] = b->sq_open[st * b->sq_blocks + k];
        }
    return HRD_OK;
}

int hrd_tx_process(hrd_batch_t *b, const int16_t *pcm, size_t n_per_stream, size_t pcm_stride, int8_t *iq,
                   size_t iq_stride, int mem, void *cuda_stream)
{
    if (!b || b->kind != HRD_TX) return fail(HRD_EINVAL, "not a Tx batch");
    if (!pcm || !iq) return fail(HRD_EINVAL, "null buffer");
    if (mem != HRD_MEM_HOST && mem != HRD_MEM_DEVICE) return fail(HRD_EINVAL, "bad mem %d", mem);
    if (n_per_stream > 0x7fffffu) return fail(HRD_EINVAL, "call too long");
    if (pcm_stride < n_per_stream) return fail(HRD_EINVAL, "pcm_stride too small");
    DeviceGuard guard(b->device);
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : (mem == HRD_MEM_HOST ? b->own : (cudaStream_t) nullptr);
    int rc = order_after_last(b, s);
    if (!rc) rc = regroup(b, s);
    if (rc) return rc;
    const bool pairs = b->any_iq8k; // a stream of int16 I,Q pairs (HRD_MODE_IQ8K) doubles the row
    const size_t in_elems = pairs ? 2 * n_per_stream : n_per_stream;
    if (pairs && (pcm_stride < in_elems || (pcm_stride & 1) || ((uintptr_t)pcm & 3)))
        return fail(HRD_EINVAL, "HRD_MODE_IQ8K rows hold %zu int16: pcm_stride must be even and at least that, pcm 4-byte aligned", in_elems);
    const size_t out_row = n_per_stream * 512;
    if (iq_stride < out_row) return fail(HRD_EINVAL, "iq_stride too small");
    if (n_per_stream == 0) return HRD_OK;

    const int16_t *d_pcm = pcm;
    int8_t *d_iq = iq;
    size_t d_pcm_stride = pcm_stride, d_iq_stride = iq_stride;
    if (mem == HRD_MEM_HOST) {
        const size_t in_row = (in_elems * sizeof(int16_t) + 31) & ~(size_t)31;
        rc = ensure_cap(&b->d_in, &b->d_in_cap, in_row * (size_t)b->n);
        if (!rc) rc = ensure_cap(&b->d_out, &b->d_out_cap, out_row * (size_t)b->n);
        if (rc) return rc;
        HRD_CUDA(cudaMemcpy2DAsync(b->d_in, in_row, pcm, pcm_stride * sizeof(int16_t), in_elems * sizeof(int16_t),
                                   (size_t)b->n, cudaMemcpyHostToDevice, s));
        d_pcm = (const int16_t *)b->d_in;
        d_pcm_stride = in_row / sizeof(int16_t);
        d_iq = (int8_t *)b->d_out;
        d_iq_stride = out_row;
    } else {
        if (((uintptr_t)iq % 32) || (iq_stride % 32))
            return fail(HRD_EINVAL, "device iq pointer/stride must be 32-byte aligned");
        if ((uintptr_t)pcm % 2) return fail(HRD_EINVAL, "pcm pointer must be 2-byte aligned");
    }
    hrd::TxParams p;
    memset(&p, 0, sizeof p);
    p.pcm = d_pcm;
    p.pcm_stride = d_pcm_stride;
    p.n8 = (uint32_t)n_per_stream;
    p.iq = d_iq;
    p.iq_stride = d_iq_stride;
    p.state = (const hrd::TxState *)b->d_state[b->cur];
    p.state_out = (hrd::TxState *)b->d_state[b->cur ^ 1];
    // every record is carried over first; the kernels then write only what their modulator owns
    HRD_CUDA(cudaMemcpyAsync(b->d_state[b->cur ^ 1], b->d_state[b->cur], sizeof(hrd::TxState) * (size_t)b->n,
                             cudaMemcpyDeviceToDevice, s));
    p.lsb = b->d_lsb;
    p.mode_of = b->d_mode;
    p.nco_sin = g_dev_tables[b->device].nco_sin;
    p.nco_cos = g_dev_tables[b->device].nco_cos;
    p.nco_iq900 = g_dev_tables[b->device].nco_iq900;
    p.nco_thr = g_dev_tables[b->device].nco_thr;
    p.sm_count = b->sm_count;
    static const int param_of_kind[hrd::K_COUNT] = {-1, HRD_PARAM_AM_INDEX, HRD_PARAM_FM_DEV, HRD_PARAM_WBFM_DEV, -1, -1};
    // a batch holding several kinds fans out like the receive side's (run_demods): one stream per kind
    int kinds = 0;
    for (int k = 0; k < hrd::K_COUNT; k++) kinds += b->group_cnt[k] != 0;
    const bool fan = kinds >= 2;
    if (fan) HRD_CUDA(cudaEventRecord(b->ev_fork, s));
    int lane = 0;
    for (int k = 0; k < hrd::K_COUNT; k++) {
        if (!b->group_cnt[k]) continue;
        cudaStream_t ks = s;
        if (fan) {
            ks = b->aux[lane & 3];
            HRD_CUDA(cudaStreamWaitEvent(ks, b->ev_fork, 0));
        }
        p.stream_ids = b->d_ids + b->group_off[k];
        p.n_streams = b->group_cnt[k];
        p.param = param_of_kind[k] >= 0 ? b->d_param[param_of_kind[k]] : nullptr;
        p.n_tiles = 1;
        p.tile_len8 = (uint32_t)((n_per_stream + 31) & ~(size_t)31);
        if (k == hrd::K_AM || k == hrd::K_FM || k == hrd::K_SSB || k == hrd::K_IQ) choose_tx_tiles(b, k, p.n_streams, (uint32_t)n_per_stream, &p.n_tiles, &p.tile_len8);
        if (k == hrd::K_FM) { // the serial NCO phase pass first (hrd_tx.cu tx_fm_phase_kernel)
            rc = ensure_cap(&b->d_fmph, &b->d_fmph_cap, sizeof(float) * (size_t)p.n_streams * n_per_stream);
            if (rc) return rc;
            p.fm_phase = (float *)b->d_fmph;
            int e = hrd::launch_tx_fm_phase(p, ks);
            if (e) return fail(HRD_ECUDA, "tx FM phase launch failed: %s", cudaGetErrorString((cudaError_t)e));
            b->launches++;
        }
        if (k == hrd::K_IQ) { // signals/fm.cc streams: their theta recurrence first (own scratch: K_FM may run beside)
            if (b->any_fm_proto) {
                rc = ensure_cap(&b->d_sigph, &b->d_sigph_cap, sizeof(float) * (size_t)p.n_streams * n_per_stream);
                if (rc) return rc;
                p.fm_phase = (float *)b->d_sigph;
                int e = hrd::launch_tx_sig_phase(p, ks);
                if (e) return fail(HRD_ECUDA, "tx signals phase launch failed: %s", cudaGetErrorString((cudaError_t)e));
                b->launches++;
            }
        }
        int e = hrd::launch_tx(k, p, ks);
        if (e) return fail(HRD_ECUDA, "tx launch (kind %d) failed: %s", k, cudaGetErrorString((cudaError_t)e));
        b->launches++;
        if (fan) { // five kinds share four side streams: the fifth queues behind the first
            HRD_CUDA(cudaEventRecord(b->ev_join[lane & 3], ks));
            HRD_CUDA(cudaStreamWaitEvent(s, b->ev_join[lane & 3], 0));
            lane++;
        }
    }
    b->cur ^= 1; // what this call wrote is what the next one reads
    if (mem == HRD_MEM_HOST) {
        HRD_CUDA(cudaMemcpy2DAsync(iq, iq_stride, b->d_out, d_iq_stride, out_row, (size_t)b->n,
                                   cudaMemcpyDeviceToHost, s));
        HRD_CUDA(cudaStreamSynchronize(s));
    }
    return mark_end(b, s);
}

} // extern "C"
