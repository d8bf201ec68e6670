// hrd_rx.cu -- receive chains: int8 I,Q at 2.048 MS/s (or 256 kS/s) -> int16 PCM at 8 kS/s.
//
// Replaces, per stream (reference paths relative to radioDiags/):
//   src_diags/IqDataProcessor.cc:429-500  reduceSampleRate  (3 x half-band /2, (int8_t) wrap)
//   src_diags/IqDataProcessor.cc:771-815  upconvertByFsOver4
//   AmDemodulator/AmDemodulator.cc:297-504, FmDemodulator/FmDemodulator.cc:353-585,
//   WbFmDemodulator/WbFmDemodulator.cc:341-500, SsbDemodulator/SsbDemodulator.cc:420-598
// and underneath them Filters/Int16/{Decimator,FirFilter}_int16.cc, Filters/{Fir,Iir}Filter.cc.
//
// One warp owns one stream.  Work is cut into BATCHES of 1024 samples at 256 kS/s (= 8192
// input samples = 32 PCM samples):
//   A. front end, 16 iterations: every lane takes 32 input bytes (16 I,Q samples, one
//      LDG.E.256) through the three /2 stages with dp2a on packed int8, passes the one
//      boundary sample each stage needs to its neighbour lane by shuffle, applies the Fs/4
//      rotation and the (int8_t) wrap, and appends one packed word (2 samples) to a
//      shared-memory ring;
//   B. the mode's demodulator runs lane-parallel over the ring at 64 k, 16 k and 8 kS/s;
//      the only serial pieces are the float IIR recurrences, run by one lane;
//   C. 32 PCM samples leave with one coalesced 64-byte store.
// Integer sections are bit-exact by construction (same Q15 arithmetic, same wrap-around);
// float sections use the reference's operation order with FMA contraction disabled
// (-fmad=false) and IEEE division.
#include "hrd_device.cuh"

namespace hrd {

__constant__ ConstTables c_tab;

void upload_tables(const ConstTables &t) { cudaMemcpyToSymbol(c_tab, &t, sizeof t); }

namespace {

constexpr int BATCH256 = 1024; // 256 kS/s samples per batch
constexpr int IT_SAMPLES = 64; // 256 kS/s samples per warp iteration

// ------------------------------------------------------------------------------------
// per-warp shared memory, by mode
// ------------------------------------------------------------------------------------
struct SmemNone {
    uint32_t dummy[4];
};
struct SmemAm {
    uint32_t r256[2 + BATCH256 / 2]; // packed int8 words, 2 samples each
    uint32_t d64[8 + BATCH256 / 4];  // I/Q int16 pairs
    uint32_t a16[14 + BATCH256 / 16];
    float f8[32];
};
struct SmemSsb {
    uint32_t r256[2 + BATCH256 / 2];
    uint32_t d64[8 + BATCH256 / 4];
    uint32_t a16[14 + BATCH256 / 16];
    uint32_t d8[30 + 32];
    float f8[32];
};
struct SmemFm {
    uint32_t r256[14 + BATCH256 / 2];
    float th[4 + BATCH256 / 4];
    int16_t d64[8 + BATCH256 / 4];
    int16_t a16[38 + BATCH256 / 16];
};
struct SmemWbfm {
    float f256[BATCH256];
    int16_t d256[4 + BATCH256];
    int16_t d64[8 + BATCH256 / 4];
    int16_t a16[38 + BATCH256 / 16];
};

template <int KIND> struct SmemOf;
template <> struct SmemOf<K_NONE> { typedef SmemNone type; };
template <> struct SmemOf<K_AM> { typedef SmemAm type; };
template <> struct SmemOf<K_FM> { typedef SmemFm type; };
template <> struct SmemOf<K_WBFM> { typedef SmemWbfm type; };
template <> struct SmemOf<K_SSB> { typedef SmemSsb type; };

// ------------------------------------------------------------------------------------
// A. front end
// ------------------------------------------------------------------------------------
struct FeCarry {
    uint32_t t, v, u; // last transposed word of stage 1/2/3 input, as seen by this lane
};

// byte 2 of a and byte 2 of b into bytes 0,1 (the >>16 of the doubled-tap accumulators)
__device__ __forceinline__ uint32_t pack_b2(int a, int b) { return __byte_perm((uint32_t)a, (uint32_t)b, 0x0062); }
// low halves of lo and hi
__device__ __forceinline__ uint32_t merge16(uint32_t lo, uint32_t hi) { return __byte_perm(lo, hi, 0x5410); }

// previous lane's value; lane 0 gets what lane 31 kept from the previous iteration
__device__ __forceinline__ uint32_t from_left(uint32_t cur, uint32_t &kept, int lane)
{
    uint32_t sel = (lane == 31) ? kept : cur;
    uint32_t left = __shfl_sync(HRD_FULL_MASK, sel, (lane + 31) & 31);
    kept = cur;
    return left;
}

// One half-band /2 stage on N packed words {I_e, I_o, Q_e, Q_o}: output m is
//   (q0*x[2m+1] + q1*x[2m] + q2*x[2m-1] + (1<<14)) >> 15          (Decimator_int16.cc:176-249)
// evaluated with doubled taps so the result sits in bits 16.. of the accumulator.
template <int N>
__device__ __forceinline__ void halfband_stage(const uint32_t (&in)[N], uint32_t left, uint32_t a, uint32_t b,
                                               int (&pi)[N], int (&pq)[N])
{
#pragma unroll
    for (int r = 0; r < N; r++) {
        uint32_t l = r ? in[r - 1] : left;
        pi[r] = dp2a_lo_us(a, in[r], dp2a_lo_us(b, l, 32768));
        pq[r] = dp2a_hi_us(a, in[r], dp2a_hi_us(b, l, 32768));
    }
}

// 16 input samples (8 raw words {I,Q,I,Q}) -> one ring word {I0,I1,Q0,Q1} at 256 kS/s,
// rotated by +Fs/4 (IqDataProcessor.cc:771-815) and narrowed like (int8_t) does.
__device__ __forceinline__ uint32_t front_end_iter(const u32x8 &w, FeCarry &c, int lane)
{
    uint32_t t[8];
#pragma unroll
    for (int r = 0; r < 8; r++) t[r] = __byte_perm(w.v[r], 0, 0x3120);
    int pi8[8], pq8[8];
    halfband_stage<8>(t, from_left(t[7], c.t, lane), c_tab.fe_a[0], c_tab.fe_b[0], pi8, pq8);
    uint32_t v[4];
#pragma unroll
    for (int j = 0; j < 4; j++)
        v[j] = merge16(pack_b2(pi8[2 * j], pi8[2 * j + 1]), pack_b2(pq8[2 * j], pq8[2 * j + 1]));
    int pi4[4], pq4[4];
    halfband_stage<4>(v, from_left(v[3], c.v, lane), c_tab.fe_a[1], c_tab.fe_b[1], pi4, pq4);
    uint32_t u[2];
#pragma unroll
    for (int k = 0; k < 2; k++)
        u[k] = merge16(pack_b2(pi4[2 * k], pi4[2 * k + 1]), pack_b2(pq4[2 * k], pq4[2 * k + 1]));
    int pi2[2], pq2[2];
    halfband_stage<2>(u, from_left(u[1], c.u, lane), c_tab.fe_a[2], c_tab.fe_b[2], pi2, pq2);
    // rotation: this lane's samples are number 2*lane and 2*lane+1 of the iteration, so
    // their phases are {0,1} on even lanes and {2,3} on odd lanes:
    //   0:(x,y) 1:(-y,x) 2:(-x,-y) 3:(y,-x).  Negating "acc>>16" is (65535-acc)>>16.
    const int s = (lane & 1) ? -1 : 1;
    const int k = (lane & 1) ? 65535 : 0;
    const int kn = (lane & 1) ? 0 : 65535;
    int i0 = pi2[0] * s + k;
    int q0 = pq2[0] * s + k;
    int i1 = pq2[1] * (-s) + kn;
    int q1 = pi2[1] * s + k;
    return merge16(pack_b2(i0, i1), pack_b2(q0, q1));
}

// ------------------------------------------------------------------------------------
// generic lane-parallel decimators over shared-memory rings
// ------------------------------------------------------------------------------------
// int8-pair words -> I/Q int16, N taps, /4 (N/2 dp2a per rail).  ring index 0 is the
// oldest history word; output j reads words 2j .. 2j+N/2-1.
template <int N>
__device__ __forceinline__ void dec4_int8(const uint32_t *ring, const uint32_t *pairs, int j, int &yi, int &yq)
{
    int ai = 1 << 14, aq = 1 << 14;
#pragma unroll
    for (int i = 0; i < N / 2; i++) {
        uint32_t w = ring[2 * j + i];
        ai = dp2a_lo_ss(pairs[i], w, ai);
        aq = dp2a_hi_ss(pairs[i], w, aq);
    }
    yi = q15(ai);
    yq = q15(aq);
}

__device__ __forceinline__ int lo16(uint32_t w) { return (int)(short)(w & 0xffffu); }
__device__ __forceinline__ int hi16(uint32_t w) { return ((int)w) >> 16; }
__device__ __forceinline__ uint32_t pack16(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }

// I/Q int16 pairs, N taps, decimate by M: output k reads ring[M*k + (N-1) - t + (hist-(N-M))]
// with hist == N-M, i.e. ring[M*k + N - 1 - t].
template <int N, int M>
__device__ __forceinline__ uint32_t dec_pairs(const uint32_t *ring, const int32_t *taps, int k)
{
    unsigned ai = 1u << 14, aq = 1u << 14;
#pragma unroll
    for (int t = 0; t < N; t++) {
        uint32_t w = ring[M * k + N - 1 - t];
        ai += (unsigned)(taps[t] * lo16(w));
        aq += (unsigned)(taps[t] * hi16(w));
    }
    return pack16(q15((int)ai), q15((int)aq));
}

// real int16 samples, N taps, decimate by M
template <int N, int M>
__device__ __forceinline__ int dec_real(const int16_t *ring, const int32_t *taps, int k)
{
    unsigned acc = 1u << 14;
#pragma unroll
    for (int t = 0; t < N; t++) acc += (unsigned)(taps[t] * (int)ring[M * k + N - 1 - t]);
    return q15((int)acc);
}

// The serial part of IirFilter::filterData for a one-tap denominator
// (Filters/IirFilter.cc:161-176): y[n] = fir[n] - a0*y[n-1], one lane, in place.
__device__ __forceinline__ float iir_serial(float *f, int n, float a0, float y1)
{
    for (int i = 0; i < n; i++) {
        float y = __fsub_rn(f[i], __fmul_rn(a0, y1));
        f[i] = y;
        y1 = y;
    }
    return y1;
}

// ------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------
template <int KIND, int ENTRY>
__global__ void __launch_bounds__(HRD_WARPS_PER_CTA * 32) rx_kernel(const RxParams p)
{
    typedef typename SmemOf<KIND>::type Smem;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int slot = blockIdx.x * HRD_WARPS_PER_CTA + warp;
    if (slot >= p.n_streams) return;
    const int sid = p.stream_ids[slot];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw + (size_t)warp * sizeof(Smem));
    RxState &st = p.state[sid];
    const int8_t *src = p.iq + (size_t)sid * p.iq_stride;

    // ---- load state -------------------------------------------------------------
    FeCarry fc;
    fc.t = st.fe_t;
    fc.v = st.fe_v;
    fc.u = st.fe_u;
    float gain = 0.f, scale = 0.f;
    float x1 = 0.f, y1 = 0.f, th_keep = 0.f, v_keep = 0.f;
    bool lsb = true;
    if constexpr (KIND != K_NONE) gain = p.gain[sid];
    if constexpr (KIND == K_AM) {
        ring_load_hist(sm.r256, st.am_r256, 2, lane);
        ring_load_hist(sm.d64, st.am_d64, 8, lane);
        ring_load_hist(sm.a16, st.am_a16, 14, lane);
        x1 = st.am_x1;
        y1 = st.am_y1;
    }
    if constexpr (KIND == K_SSB) {
        ring_load_hist(sm.r256, st.ssb_r256, 2, lane);
        ring_load_hist(sm.d64, st.ssb_d64, 8, lane);
        ring_load_hist(sm.a16, st.ssb_a16, 14, lane);
        ring_load_hist(sm.d8, st.ssb_d8, 30, lane);
        x1 = st.ssb_x1;
        y1 = st.ssb_y1;
        lsb = p.lsb[sid] != 0;
    }
    if constexpr (KIND == K_FM) {
        ring_load_hist(sm.r256, st.fm_r256, 14, lane);
        ring_load_hist(sm.th, st.fm_theta, 4, lane);
        ring_load_hist(sm.d64, st.fm_d64, 8, lane);
        ring_load_hist(sm.a16, st.fm_a16, 38, lane);
        // FmDemodulator.cc:488-491
        scale = __fmul_rn(__fdiv_rn(gain, 15000.f), 32767.f);
    }
    if constexpr (KIND == K_WBFM) {
        ring_load_hist(sm.d256, st.wb_d256, 4, lane);
        ring_load_hist(sm.d64, st.wb_d64, 8, lane);
        ring_load_hist(sm.a16, st.wb_a16, 38, lane);
        x1 = st.wb_x1;
        y1 = st.wb_y1;
        th_keep = st.wb_prev_theta;
        v_keep = x1;
        // WbFmDemodulator.cc:392-395
        scale = __fmul_rn(__fdiv_rn(gain, 75000.f), 32767.f);
    }
    __syncwarp();

    const uint32_t n256 = p.n256;
    uint32_t done256 = 0;
    uint32_t last_active = 32;

    while (done256 < n256) {
        const uint32_t nb = min((uint32_t)BATCH256, n256 - done256); // multiple of 32
        const uint32_t n_it = (nb + IT_SAMPLES - 1) / IT_SAMPLES;

        // ---- A. front end (or plain load at the 256 kS/s entry) ----------------------
        for (uint32_t it = 0; it < n_it; it++) {
            const uint32_t s0 = done256 + it * IT_SAMPLES + 2 * lane; // first sample of this lane
            const bool active = s0 < n256;
            if (it + 1 == n_it) last_active = min(32u, (nb - it * IT_SAMPLES) / 2);
            uint32_t word;
            if constexpr (ENTRY == 0) {
                u32x8 w;
                if (active) {
                    w = ldg_stream_256(src + (size_t)s0 * 16);
                } else {
#pragma unroll
                    for (int r = 0; r < 8; r++) w.v[r] = 0;
                }
                word = front_end_iter(w, fc, lane);
            } else {
                uint32_t raw = active ? __ldg(reinterpret_cast<const uint32_t *>(src + (size_t)s0 * 2)) : 0u;
                word = __byte_perm(raw, 0, 0x3120);
            }
            const uint32_t widx = it * 32 + lane; // word index inside the batch
            if constexpr (KIND == K_NONE) {
                if (p.out256 && active)
                    *reinterpret_cast<uint32_t *>(p.out256 + (size_t)sid * p.out_stride + (size_t)s0 * 2) =
                        __byte_perm(word, 0, 0x3120);
            } else if constexpr (KIND == K_AM || KIND == K_SSB) {
                if (active) sm.r256[2 + widx] = word;
            } else if constexpr (KIND == K_FM) {
                if (active) sm.r256[14 + widx] = word;
            } else { // K_WBFM: discriminator + the FIR half of the de-emphasis filter, in place
                // uint8_t idx = (uint8_t)sample + 128  (WbFmDemodulator.cc:403-404)
                uint32_t x = word ^ 0x80808080u;
                uint32_t idx0 = __byte_perm(x, 0, 0x4420), idx1 = __byte_perm(x, 0, 0x4431);
                float th0 = __ldg(p.atan2_lut + idx0), th1 = __ldg(p.atan2_lut + idx1);
                // theta of the previous sample: previous lane's th1 (lane 0: kept from before)
                float sel = (lane == 31) ? th_keep : th1;
                float thp = __shfl_sync(HRD_FULL_MASK, sel, (lane + 31) & 31);
                float d0 = wrap_pi(__fsub_rn(th0, thp));
                float d1 = wrap_pi(__fsub_rn(th1, th0));
                float v0 = __fmul_rn(scale, d0), v1 = __fmul_rn(scale, d1);
                float selv = (lane == 31) ? v_keep : v1;
                float vp = __shfl_sync(HRD_FULL_MASK, selv, (lane + 31) & 31);
                if (active) {
                    th_keep = th1;
                    v_keep = v1;
                }
                // FirFilter::filterData order: y = 0 + b0*x[n]; y = y + b1*x[n-1]
                const float b = 0.0253863f;
                float f0 = __fadd_rn(__fmul_rn(b, v0), __fmul_rn(b, vp));
                float f1 = __fadd_rn(__fmul_rn(b, v1), __fmul_rn(b, v0));
                if (active) {
                    sm.f256[2 * widx] = f0;
                    sm.f256[2 * widx + 1] = f1;
                }
            }
        }
        __syncwarp();

        const int n64 = nb / 4, n16 = nb / 16, n8 = nb / 32;
        int16_t *pcm_out = p.pcm + (size_t)sid * p.pcm_stride + done256 / 32;

        // ---- B. demodulators ----------------------------------------------------------
        if constexpr (KIND == K_AM || KIND == K_SSB) {
            // AmDemodulator.cc:339-408 / SsbDemodulator.cc:462-529: /4 (8) /4 (12) /2 (16)
            for (int j = lane; j < n64; j += 32) {
                int yi, yq;
                dec4_int8<8>(sm.r256, c_tab.am1, j, yi, yq);
                sm.d64[8 + j] = pack16(yi, yq);
            }
            __syncwarp();
            for (int k = lane; k < n16; k += 32) sm.a16[14 + k] = dec_pairs<12, 4>(sm.d64, c_tab.am2, k);
            __syncwarp();
            uint32_t iq8 = 0;
            if (lane < n8) iq8 = dec_pairs<16, 2>(sm.a16, c_tab.am3, lane);
            float fir = 0.f;
            if constexpr (KIND == K_AM) {
                // AmDemodulator.cc:444-458: |I|,|Q| narrowed to int16, max + min/2
                int im = (int)(short)abs(lo16(iq8)), qm = (int)(short)abs(hi16(iq8));
                int mag = (im > qm) ? (int)(short)(im + (qm >> 1)) : (int)(short)(qm + (im >> 1));
                float x = (float)mag;
                float xp = __shfl_up_sync(HRD_FULL_MASK, x, 1);
                if (lane == 0) xp = x1;
                x1 = __shfl_sync(HRD_FULL_MASK, x, n8 - 1);
                fir = __fsub_rn(x, xp); // b = {1,-1}: 0 + 1*x[n], then + (-1)*x[n-1]
            } else {
                if (lane < n8) sm.d8[30 + lane] = iq8;
                __syncwarp();
                float x = 0.f;
                if (lane < n8) {
                    // SsbDemodulator.cc:576-590: delay line (= -I[n-15]) and 31-tap Hilbert on Q
                    const uint32_t *r = sm.d8 + lane; // r[30] is sample n
                    int id = q15((1 << 14) + c_tab.delay[15] * lo16(r[30 - 15]));
                    unsigned acc = 1u << 14;
#pragma unroll
                    for (int t = 0; t < 31; t += 2) acc += (unsigned)(c_tab.hilbert[t] * hi16(r[30 - t]));
                    int qh = q15((int)acc);
                    x = lsb ? (float)(id - qh) : (float)(id + qh);
                }
                float xp = __shfl_up_sync(HRD_FULL_MASK, x, 1);
                if (lane == 0) xp = x1;
                x1 = __shfl_sync(HRD_FULL_MASK, x, n8 - 1);
                fir = __fsub_rn(x, xp);
            }
            sm.f8[lane] = fir;
            __syncwarp();
            if (lane == 0) y1 = iir_serial(sm.f8, n8, -0.95f, y1);
            y1 = __shfl_sync(HRD_FULL_MASK, y1, 0);
            __syncwarp();
            if (lane < n8) pcm_out[lane] = (int16_t)f32_to_i16(__fmul_rn(gain, sm.f8[lane]));
            __syncwarp();
            ring_shift(sm.r256, 2, nb / 2, lane);
            ring_shift(sm.d64, 8, n64, lane);
            ring_shift(sm.a16, 14, n16, lane);
            if constexpr (KIND == K_SSB) ring_shift(sm.d8, 30, n8, lane);
        }

        if constexpr (KIND == K_FM) {
            // FmDemodulator.cc:395-442: tuner /4, 32 taps per rail, then the atan2 table
            for (int j = lane; j < n64; j += 32) {
                int yi, yq;
                dec4_int8<32>(sm.r256, c_tab.fm_tuner, j, yi, yq);
                uint32_t ii = ((uint32_t)yi + 128u) & 255u, qi = ((uint32_t)yq + 128u) & 255u;
                sm.th[4 + j] = __ldg(p.atan2_lut + qi * 256u + ii);
            }
            __syncwarp();
            // FmDemodulator.cc:479-529: FirFilter with taps {0,0,1,0,-1,0,0} = th[n-2]-th[n-4]
            for (int j = lane; j < n64; j += 32) {
                float d = wrap_pi(__fsub_rn(sm.th[4 + j - 2], sm.th[4 + j - 4]));
                sm.d64[8 + j] = (int16_t)f32_to_i16(__fmul_rn(scale, d));
            }
            __syncwarp();
            // FmDemodulator.cc:551-585: /4 (12 taps) then /2 (40 taps)
            for (int k = lane; k < n16; k += 32) sm.a16[38 + k] = (int16_t)dec_real<12, 4>(sm.d64, c_tab.fm_post, k);
            __syncwarp();
            if (lane < n8) pcm_out[lane] = (int16_t)dec_real<40, 2>(sm.a16, c_tab.audio40, lane);
            __syncwarp();
            ring_shift(sm.r256, 14, nb / 2, lane);
            ring_shift(sm.th, 4, n64, lane);
            ring_shift(sm.d64, 8, n64, lane);
            ring_shift(sm.a16, 38, n16, lane);
        }

        if constexpr (KIND == K_WBFM) {
            // the recursive half of the de-emphasis filter, serial at 256 kS/s
            if (lane == 0) y1 = iir_serial(sm.f256, (int)nb, -0.9492274f, y1);
            y1 = __shfl_sync(HRD_FULL_MASK, y1, 0);
            __syncwarp();
            // WbFmDemodulator.cc:460-500: (int16_t) cast, /4 (8) /4 (12) /2 (40)
            for (int i = lane; i < (int)nb; i += 32) sm.d256[4 + i] = (int16_t)f32_to_i16(sm.f256[i]);
            __syncwarp();
            for (int j = lane; j < n64; j += 32) sm.d64[8 + j] = (int16_t)dec_real<8, 4>(sm.d256, c_tab.wbfm_post1, j);
            __syncwarp();
            for (int k = lane; k < n16; k += 32) sm.a16[38 + k] = (int16_t)dec_real<12, 4>(sm.d64, c_tab.fm_post, k);
            __syncwarp();
            if (lane < n8) pcm_out[lane] = (int16_t)dec_real<40, 2>(sm.a16, c_tab.audio40, lane);
            __syncwarp();
            ring_shift(sm.d256, 4, (int)nb, lane);
            ring_shift(sm.d64, 8, n64, lane);
            ring_shift(sm.a16, 38, n16, lane);
        }
        done256 += nb;
    }

    // ---- save state ------------------------------------------------------------------
    if constexpr (ENTRY == 0) {
        if (lane == (int)last_active - 1) {
            st.fe_t = fc.t;
            st.fe_v = fc.v;
            st.fe_u = fc.u;
        }
    }
    if constexpr (KIND == K_AM) {
        ring_save_hist(sm.r256, st.am_r256, 2, lane);
        ring_save_hist(sm.d64, st.am_d64, 8, lane);
        ring_save_hist(sm.a16, st.am_a16, 14, lane);
        if (lane == 0) {
            st.am_x1 = x1;
            st.am_y1 = y1;
        }
    }
    if constexpr (KIND == K_SSB) {
        ring_save_hist(sm.r256, st.ssb_r256, 2, lane);
        ring_save_hist(sm.d64, st.ssb_d64, 8, lane);
        ring_save_hist(sm.a16, st.ssb_a16, 14, lane);
        ring_save_hist(sm.d8, st.ssb_d8, 30, lane);
        if (lane == 0) {
            st.ssb_x1 = x1;
            st.ssb_y1 = y1;
        }
    }
    if constexpr (KIND == K_FM) {
        ring_save_hist(sm.r256, st.fm_r256, 14, lane);
        ring_save_hist(sm.th, st.fm_theta, 4, lane);
        ring_save_hist(sm.d64, st.fm_d64, 8, lane);
        ring_save_hist(sm.a16, st.fm_a16, 38, lane);
    }
    if constexpr (KIND == K_WBFM) {
        ring_save_hist(sm.d256, st.wb_d256, 4, lane);
        ring_save_hist(sm.d64, st.wb_d64, 8, lane);
        ring_save_hist(sm.a16, st.wb_a16, 38, lane);
        if (lane == (int)last_active - 1) {
            st.wb_prev_theta = th_keep;
            st.wb_x1 = v_keep;
        }
        if (lane == 0) st.wb_y1 = y1;
    }
}

template <int KIND, int ENTRY>
int launch_one(const RxParams &p, cudaStream_t s)
{
    typedef typename SmemOf<KIND>::type Smem;
    const int grid = (p.n_streams + HRD_WARPS_PER_CTA - 1) / HRD_WARPS_PER_CTA;
    const size_t smem = sizeof(Smem) * HRD_WARPS_PER_CTA;
    rx_kernel<KIND, ENTRY><<<grid, HRD_WARPS_PER_CTA * 32, smem, s>>>(p);
    return (int)cudaGetLastError();
}

} // namespace

int launch_rx(int kind, int entry, const RxParams &p, cudaStream_t s)
{
    if (p.n_streams <= 0) return 0;
#define HRD_RX_CASE(K)                                                           \
    case K:                                                                      \
        return entry == 0 ? launch_one<K, 0>(p, s) : launch_one<K, 1>(p, s);
    switch (kind) {
        HRD_RX_CASE(K_NONE)
        HRD_RX_CASE(K_AM)
        HRD_RX_CASE(K_FM)
        HRD_RX_CASE(K_WBFM)
        HRD_RX_CASE(K_SSB)
    }
#undef HRD_RX_CASE
    return (int)cudaErrorInvalidValue;
}

} // namespace hrd
