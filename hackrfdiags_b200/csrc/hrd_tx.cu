// hrd_tx.cu -- transmit chains: int16 PCM at 8 kS/s -> int8 I,Q at 2.048 MS/s.
//
// Replaces, per stream (reference paths relative to radioDiags/):
//   AmModulator/AmModulator.cc:366-607, FmModulator/FmModulator.cc:353-622,
//   WbFmModulator/WbFmModulator.cc:347-632, SsbModulator/SsbModulator.cc:430-707
// and underneath them Filters/Int16/{Interpolator,FirFilter}_int16.cc and
// Nco/{Nco,PhaseAccumulator}.cc.
//
// One warp owns one stream.  A batch is 32 PCM samples (8192 output I,Q samples, 16 KiB):
//   1. the modulator head at 8 kS/s, one PCM sample per lane (the NCO phase recurrence is
//      the only serial piece);
//   2. interpolator stages 1..4 lane-parallel through small shared-memory rings;
//   3. the hot loop, 16 iterations: every lane takes ONE 128 kS/s sample through stages
//      5,6,7,8 in registers (the left-neighbour sample each half-band stage needs is an
//      odd-phase output, which depends on a single input, so it is recomputed locally - no
//      shuffles), narrows like (int8_t) does and writes 32 contiguous bytes with one
//      STG.E.256: 1 KiB per warp instruction.
// All interpolation is the reference's Q15 arithmetic stage by stage (each stage rounds to
// int16, so stages cannot be merged): y[nL+i] = (16384 + sum_k q[i+kL]*x[n-k]) >> 15
// (Interpolator_int16.cc:398-418).
#include "hrd_device.cuh"

namespace hrd {

__constant__ ConstTables c_tabtx;

void upload_tables_tx(const ConstTables &t) { cudaMemcpyToSymbol(c_tabtx, &t, sizeof t); }

namespace {

constexpr int NB8 = 32; // PCM samples per batch

__device__ __forceinline__ int lo16(uint32_t w) { return (int)(short)(w & 0xffffu); }
__device__ __forceinline__ int hi16(uint32_t w) { return ((int)w) >> 16; }
__device__ __forceinline__ uint32_t pack16(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }

struct SmemTx {
    uint32_t s0[19 + NB8];      // stage 1 input  @8k   I/Q pairs (WBFM: PCM in the low half)
    uint32_t s1[3 + 2 * NB8];   // stage 2 input  @16k
    uint32_t s2[1 + 4 * NB8];   // stage 3 input  @32k
    uint32_t s3[3 + 8 * NB8];   // stage 4 input  @64k
    uint32_t s4[3 + 16 * NB8];  // stage 5 input  @128k
    uint32_t h8[30 + NB8];      // SSB: PCM/2 history for delay line / Hilbert
    float ph[NB8];              // FM: NCO phases of the batch
};

struct SmemTxWb {
    uint32_t s0[19 + NB8];
    uint32_t s1[3 + 2 * NB8];
    uint32_t s2[1 + 4 * NB8];
    uint32_t s3[3 + 8 * NB8];
    uint32_t s4[3 + 16 * NB8];
    uint32_t s5[1 + 32 * NB8];  // stage 6 input @256k: first the real PCM, then I/Q pairs in place
    float ph[32 * NB8];         // NCO phase per 256 kS/s sample
};

// ---- generic polyphase stages over rings of I/Q pairs ------------------------------------
// stage 1: 40 taps, L = 2 (20 taps per branch); ring hist 19, input n at ring[19 + n]
__device__ __forceinline__ void interp40(const uint32_t *in, int n, uint32_t &even, uint32_t &odd)
{
    unsigned ei = 1u << 14, eq = 1u << 14, oi = 1u << 14, oq = 1u << 14;
#pragma unroll
    for (int k = 0; k < 20; k++) {
        uint32_t w = in[19 + n - k];
        int xi = lo16(w), xq = hi16(w);
        ei += (unsigned)(c_tabtx.audio40[2 * k] * xi);
        eq += (unsigned)(c_tabtx.audio40[2 * k] * xq);
        oi += (unsigned)(c_tabtx.audio40[2 * k + 1] * xi);
        oq += (unsigned)(c_tabtx.audio40[2 * k + 1] * xq);
    }
    even = pack16(q15((int)ei), q15((int)eq));
    odd = pack16(q15((int)oi), q15((int)oq));
}

// stages 2,4,5: 8 taps, L = 2; ring hist 3, input n at ring[3 + n]
__device__ __forceinline__ void interp8(const uint32_t *in, int n, uint32_t &even, uint32_t &odd)
{
    int ei = 1 << 14, eq = 1 << 14, oi = 1 << 14, oq = 1 << 14;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t w = in[3 + n - k];
        int xi = lo16(w), xq = hi16(w);
        ei += c_tabtx.tx_hb8[2 * k] * xi;
        eq += c_tabtx.tx_hb8[2 * k] * xq;
        oi += c_tabtx.tx_hb8[2 * k + 1] * xi;
        oq += c_tabtx.tx_hb8[2 * k + 1] * xq;
    }
    even = pack16(q15(ei), q15(eq));
    odd = pack16(q15(oi), q15(oq));
}

// stages 3,6,7,8: taps {c, m, c, 0}, L = 2:  even = c*(x[n]+x[n-1]), odd = m*x[n]
__device__ __forceinline__ int hb4_even(int c, int x, int xm1) { return q15((1 << 14) + c * x + c * xm1); }
__device__ __forceinline__ int hb4_odd(int m, int x) { return q15((1 << 14) + m * x); }

// ---- PhaseAccumulator::run (Nco/PhaseAccumulator.cc:157-181) -------------------------------
__device__ __forceinline__ float phase_advance(float acc, float step)
{
    const double pi = 3.14159265358979323846;
    acc = __fadd_rn(acc, step);
    while ((double)acc > pi) acc = (float)((double)acc - 2.0 * pi);
    while ((double)acc < -pi) acc = (float)((double)acc + 2.0 * pi);
    return acc;
}

// PhaseAccumulator::setFrequency (:95-107): (float)((2*M_PI*f)/fs) in double
__device__ __forceinline__ float phase_step(float f, double fs)
{
    return (float)((2.0 * 3.14159265358979323846 * (double)f) / fs);
}

// Stages 5..8 of one rail for one 128 kS/s input sample x (with its three predecessors):
// 16 outputs as accumulators whose byte 2 is the (int8_t) value (doubled taps, see hrd_rx.cu).
__device__ __forceinline__ void tail4(int x0, int x1, int x2, int x3, int (&out)[16])
{
    const int *h = c_tabtx.tx_hb8;
    // stage 5 (8 taps): even uses x[n..n-3], odd uses taps {q1,q3,q5,q7} on the same four
    int y5[2], y5m1;
    y5[0] = q15((1 << 14) + h[0] * x0 + h[2] * x1 + h[4] * x2 + h[6] * x3);
    y5[1] = q15((1 << 14) + h[1] * x0 + h[3] * x1 + h[5] * x2 + h[7] * x3);
    // previous odd output (input n-1): needs x[n-1..n-4]; q7 == 0 so x[n-4] drops out only
    // if the tap is zero -- it is (AmModulator.cc:57-67), asserted on the host.
    y5m1 = q15((1 << 14) + h[1] * x1 + h[3] * x2 + h[5] * x3);
    const int c6 = c_tabtx.tx_c3, m6 = c_tabtx.tx_m3;
    const int c7 = c_tabtx.tx_c7, m7 = c_tabtx.tx_m7;
    const int c8 = c_tabtx.tx_c8, m8 = c_tabtx.tx_m8;
    int y6[4], y6m1;
    y6m1 = hb4_odd(m6, y5m1);
    y6[0] = hb4_even(c6, y5[0], y5m1);
    y6[1] = hb4_odd(m6, y5[0]);
    y6[2] = hb4_even(c6, y5[1], y5[0]);
    y6[3] = hb4_odd(m6, y5[1]);
    int y7[8], y7m1;
    y7m1 = hb4_odd(m7, y6m1);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        y7[2 * k] = hb4_even(c7, y6[k], k ? y6[k - 1] : y6m1);
        y7[2 * k + 1] = hb4_odd(m7, y6[k]);
    }
    // stage 8 with doubled taps: (int8_t)(acc>>15) == byte 2 of 2*acc
#pragma unroll
    for (int k = 0; k < 8; k++) {
        int left = k ? y7[k - 1] : y7m1;
        out[2 * k] = (1 << 15) + 2 * c8 * (y7[k] + left);
        out[2 * k + 1] = (1 << 15) + 2 * m8 * y7[k];
    }
}

__device__ __forceinline__ uint32_t pack_b2(int a, int b) { return __byte_perm((uint32_t)a, (uint32_t)b, 0x0062); }
__device__ __forceinline__ uint32_t merge16(uint32_t lo, uint32_t hi) { return __byte_perm(lo, hi, 0x5410); }

// ---- stages 6..8 only (WBFM: the NCO sits between stage 5 and stage 6) ----------------------
__device__ __forceinline__ void tail3(int x0, int xm1, int (&out)[8])
{
    const int c6 = c_tabtx.tx_c3, m6 = c_tabtx.tx_m3;
    const int c7 = c_tabtx.tx_c7, m7 = c_tabtx.tx_m7;
    const int c8 = c_tabtx.tx_c8, m8 = c_tabtx.tx_m8;
    int y6[2], y6m1;
    y6m1 = hb4_odd(m6, xm1);
    y6[0] = hb4_even(c6, x0, xm1);
    y6[1] = hb4_odd(m6, x0);
    int y7[4], y7m1;
    y7m1 = hb4_odd(m7, y6m1);
#pragma unroll
    for (int k = 0; k < 2; k++) {
        y7[2 * k] = hb4_even(c7, y6[k], k ? y6[k - 1] : y6m1);
        y7[2 * k + 1] = hb4_odd(m7, y6[k]);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int left = k ? y7[k - 1] : y7m1;
        out[2 * k] = (1 << 15) + 2 * c8 * (y7[k] + left);
        out[2 * k + 1] = (1 << 15) + 2 * m8 * y7[k];
    }
}

// ------------------------------------------------------------------------------------
// AM / FM / SSB kernel
// ------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(HRD_WARPS_PER_CTA * 32) tx_kernel(const TxParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int slot = blockIdx.x * HRD_WARPS_PER_CTA + warp;
    if (slot >= p.n_streams) return;
    const int sid = p.stream_ids[slot];
    SmemTx &sm = *reinterpret_cast<SmemTx *>(smem_raw + (size_t)warp * sizeof(SmemTx));
    TxState &st = p.state[sid];
    TxRail8 &rs = KIND == K_AM ? st.am : (KIND == K_FM ? st.fm : st.ssb);
    const int16_t *src = p.pcm + (size_t)sid * p.pcm_stride;
    int8_t *dst = p.iq + (size_t)sid * p.iq_stride;

    ring_load_hist(sm.s0, rs.s0, 19, lane);
    ring_load_hist(sm.s1, rs.s1, 3, lane);
    ring_load_hist(sm.s2, rs.s2, 1, lane);
    ring_load_hist(sm.s3, rs.s3, 3, lane);
    ring_load_hist(sm.s4, rs.s4, 3, lane);
    if constexpr (KIND == K_SSB) ring_load_hist(sm.h8, st.ssb_h8, 30, lane);
    const float prm = (KIND == K_SSB) ? 0.f : p.param[sid];
    const bool lsb = (KIND == K_SSB) ? (p.lsb[sid] != 0) : true;
    float phase = (KIND == K_FM) ? st.fm_phase : 0.f;
    __syncwarp();

    for (uint32_t done = 0; done < p.n8; done += NB8) {
        const int nb = (int)min((uint32_t)NB8, p.n8 - done);
        // ---- 1. modulator head, one PCM sample per lane ---------------------------------
        const int x = (lane < nb) ? (int)src[done + lane] : 0;
        uint32_t head = 0;
        if constexpr (KIND == K_AM) {
            // AmModulator.cc:583-602
            float s = __fdiv_rn((float)x, 32768.f);
            s = __fmul_rn(s, prm);
            s = __fadd_rn(s, 1.f);
            s = __fdiv_rn(s, 2.f);
            int m = f32_to_i16(__fmul_rn(__fmul_rn(s, 128.f), 250.f));
            head = pack16(m, m);
        }
        if constexpr (KIND == K_FM) {
            // FmModulator.cc:596-617: frequency -> phase step, NCO at 8 kS/s
            float f = __fdiv_rn(__fmul_rn(prm, (float)x), 32768.f);
            float step = phase_step(f, 8000.0);
            // serial phase recurrence: lane n needs the phase BEFORE step n is added
            float my_phase = 0.f;
            for (int n = 0; n < nb; n++) {
                float sn = __shfl_sync(HRD_FULL_MASK, step, n);
                if (lane == n) my_phase = phase;
                phase = phase_advance(phase, sn);
            }
            // Nco::run (Nco.cc:186-199): cosf/sinf of the float phase.  Evaluated in double
            // and rounded to float (see DESIGN.md "float tolerance").
            double sd, cd;
            sincos((double)my_phase, &sd, &cd);
            int ci = f32_to_i16(__fmul_rn((float)cd, 16000.f));
            int si = f32_to_i16(__fmul_rn((float)sd, 16000.f));
            head = pack16(ci, si);
        }
        if constexpr (KIND == K_SSB) {
            // SsbModulator.cc:676-700
            int half = f32_to_i16(__fdiv_rn((float)x, 2.f));
            if (lane < nb) sm.h8[30 + lane] = (uint32_t)half;
            __syncwarp();
            const uint32_t *r = sm.h8 + lane; // r[30] is sample n
            int id = q15((1 << 14) + c_tabtx.delay[15] * (int)r[30 - 15]);
            unsigned acc = 1u << 14;
#pragma unroll
            for (int t = 0; t < 31; t += 2) acc += (unsigned)(c_tabtx.hilbert[t] * (int)r[30 - t]);
            int qh = q15((int)acc);
            if (!lsb) qh = (int)(short)(-qh);
            head = pack16(id, qh);
        }
        if (lane < nb) sm.s0[19 + lane] = head;
        __syncwarp();

        // ---- 2. stages 1..4 ---------------------------------------------------------------
        for (int n = lane; n < nb; n += 32) interp40(sm.s0, n, sm.s1[3 + 2 * n], sm.s1[3 + 2 * n + 1]);
        __syncwarp();
        for (int n = lane; n < 2 * nb; n += 32) interp8(sm.s1, n, sm.s2[1 + 2 * n], sm.s2[1 + 2 * n + 1]);
        __syncwarp();
        for (int n = lane; n < 4 * nb; n += 32) {
            uint32_t w = sm.s2[1 + n], wm = sm.s2[n];
            const int c = c_tabtx.tx_c3, m = c_tabtx.tx_m3;
            sm.s3[3 + 2 * n] = pack16(hb4_even(c, lo16(w), lo16(wm)), hb4_even(c, hi16(w), hi16(wm)));
            sm.s3[3 + 2 * n + 1] = pack16(hb4_odd(m, lo16(w)), hb4_odd(m, hi16(w)));
        }
        __syncwarp();
        for (int n = lane; n < 8 * nb; n += 32) interp8(sm.s3, n, sm.s4[3 + 2 * n], sm.s4[3 + 2 * n + 1]);
        __syncwarp();

        // ---- 3. hot loop: stages 5..8, 32 bytes per lane per iteration ------------------------
        int8_t *out = dst + (size_t)done * 512;
        for (int n = lane; n < 16 * nb; n += 32) {
            uint32_t w0 = sm.s4[3 + n], w1 = sm.s4[2 + n], w2 = sm.s4[1 + n], w3 = sm.s4[n];
            int oi[16], oq[16];
            tail4(lo16(w0), lo16(w1), lo16(w2), lo16(w3), oi);
            if constexpr (KIND == K_AM) {
#pragma unroll
                for (int k = 0; k < 16; k++) oq[k] = oi[k];
            } else {
                tail4(hi16(w0), hi16(w1), hi16(w2), hi16(w3), oq);
            }
            u32x8 o;
#pragma unroll
            for (int k = 0; k < 8; k++) // bytes {I[2k], Q[2k], I[2k+1], Q[2k+1]}
                o.v[k] = merge16(pack_b2(oi[2 * k], oq[2 * k]), pack_b2(oi[2 * k + 1], oq[2 * k + 1]));
            stg_stream_256(out + (size_t)n * 32, o);
        }
        __syncwarp();
        ring_shift(sm.s0, 19, nb, lane);
        ring_shift(sm.s1, 3, 2 * nb, lane);
        ring_shift(sm.s2, 1, 4 * nb, lane);
        ring_shift(sm.s3, 3, 8 * nb, lane);
        ring_shift(sm.s4, 3, 16 * nb, lane);
        if constexpr (KIND == K_SSB) ring_shift(sm.h8, 30, nb, lane);
    }

    ring_save_hist(sm.s0, rs.s0, 19, lane);
    ring_save_hist(sm.s1, rs.s1, 3, lane);
    ring_save_hist(sm.s2, rs.s2, 1, lane);
    ring_save_hist(sm.s3, rs.s3, 3, lane);
    ring_save_hist(sm.s4, rs.s4, 3, lane);
    if constexpr (KIND == K_SSB) ring_save_hist(sm.h8, st.ssb_h8, 30, lane);
    if constexpr (KIND == K_FM) {
        if (lane == 0) st.fm_phase = phase;
    }
    // stages 6,7,8 keep one input sample each; it is always the last output of the stage
    // before, which tail4 recomputes from s4's history, so nothing more needs saving.
}

// ------------------------------------------------------------------------------------
// WBFM kernel: stages 1..5 on the real PCM, NCO at 256 kS/s, stages 6..8 on I and Q
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HRD_WARPS_PER_CTA * 32) tx_wbfm_kernel(const TxParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int slot = blockIdx.x * HRD_WARPS_PER_CTA + warp;
    if (slot >= p.n_streams) return;
    const int sid = p.stream_ids[slot];
    SmemTxWb &sm = *reinterpret_cast<SmemTxWb *>(smem_raw + (size_t)warp * sizeof(SmemTxWb));
    TxState &st = p.state[sid];
    TxRail8 &rs = st.wb;
    const int16_t *src = p.pcm + (size_t)sid * p.pcm_stride;
    int8_t *dst = p.iq + (size_t)sid * p.iq_stride;

    ring_load_hist(sm.s0, rs.s0, 19, lane);
    ring_load_hist(sm.s1, rs.s1, 3, lane);
    ring_load_hist(sm.s2, rs.s2, 1, lane);
    ring_load_hist(sm.s3, rs.s3, 3, lane);
    ring_load_hist(sm.s4, rs.s4, 3, lane);
    ring_load_hist(sm.s5, rs.s5, 1, lane);
    const float dev = p.param[sid];
    float phase = st.wb_phase;
    __syncwarp();

    for (uint32_t done = 0; done < p.n8; done += NB8) {
        const int nb = (int)min((uint32_t)NB8, p.n8 - done);
        if (lane < nb) sm.s0[19 + lane] = (uint32_t)(int)src[done + lane] & 0xffffu;
        __syncwarp();
        // WbFmModulator.cc:389-441: stages 1..5 on the PCM (only the low halves are live)
        for (int n = lane; n < nb; n += 32) interp40(sm.s0, n, sm.s1[3 + 2 * n], sm.s1[3 + 2 * n + 1]);
        __syncwarp();
        for (int n = lane; n < 2 * nb; n += 32) interp8(sm.s1, n, sm.s2[1 + 2 * n], sm.s2[1 + 2 * n + 1]);
        __syncwarp();
        for (int n = lane; n < 4 * nb; n += 32) {
            uint32_t w = sm.s2[1 + n], wm = sm.s2[n];
            sm.s3[3 + 2 * n] = pack16(hb4_even(c_tabtx.tx_c3, lo16(w), lo16(wm)), 0);
            sm.s3[3 + 2 * n + 1] = pack16(hb4_odd(c_tabtx.tx_m3, lo16(w)), 0);
        }
        __syncwarp();
        for (int n = lane; n < 8 * nb; n += 32) interp8(sm.s3, n, sm.s4[3 + 2 * n], sm.s4[3 + 2 * n + 1]);
        __syncwarp();
        // stage 5 -> phase step per 256 kS/s sample (WbFmModulator.cc:596-604)
        for (int n = lane; n < 16 * nb; n += 32) {
            uint32_t e, o;
            interp8(sm.s4, n, e, o);
            float fe = __fdiv_rn(__fmul_rn(dev, (float)lo16(e)), 1024.f);
            float fo = __fdiv_rn(__fmul_rn(dev, (float)lo16(o)), 1024.f);
            sm.ph[2 * n] = phase_step(fe, 256000.0);
            sm.ph[2 * n + 1] = phase_step(fo, 256000.0);
        }
        __syncwarp();
        // serial NCO phase recurrence at 256 kS/s: ph[n] <- phase before step n
        if (lane == 0) {
            for (int n = 0; n < 32 * nb; n++) {
                float stp = sm.ph[n];
                sm.ph[n] = phase;
                phase = phase_advance(phase, stp);
            }
        }
        phase = __shfl_sync(HRD_FULL_MASK, phase, 0);
        __syncwarp();
        // Nco::runFast (Nco.cc:222-257) and the x900 scaling (WbFmModulator.cc:606-626)
        for (int n = lane; n < 32 * nb; n += 32) {
            float ph = sm.ph[n];
            double t = (double)__fmul_rn(ph, 16384.f) / (2.0 * 3.14159265358979323846);
            int v = __double2int_rz(t);
            if (!(t > -2147483649.0 && t < 2147483648.0)) v = (int)0x80000000;
            int idx = (int)(short)v + 8192;
            idx = idx < 0 ? 0 : (idx > 16383 ? 16383 : idx);
            int ci = f32_to_i16(__fmul_rn(__ldg(p.nco_cos + idx), 900.f));
            int si = f32_to_i16(__fmul_rn(__ldg(p.nco_sin + idx), 900.f));
            sm.s5[1 + n] = pack16(ci, si);
        }
        __syncwarp();
        // stages 6..8: one 256 kS/s sample -> 8 output samples = 16 bytes per lane
        int8_t *out = dst + (size_t)done * 512;
        for (int n = lane; n < 32 * nb; n += 32) {
            uint32_t w0 = sm.s5[1 + n], w1 = sm.s5[n];
            int oi[8], oq[8];
            tail3(lo16(w0), lo16(w1), oi);
            tail3(hi16(w0), hi16(w1), oq);
            uint4 o;
            o.x = merge16(pack_b2(oi[0], oq[0]), pack_b2(oi[1], oq[1]));
            o.y = merge16(pack_b2(oi[2], oq[2]), pack_b2(oi[3], oq[3]));
            o.z = merge16(pack_b2(oi[4], oq[4]), pack_b2(oi[5], oq[5]));
            o.w = merge16(pack_b2(oi[6], oq[6]), pack_b2(oi[7], oq[7]));
            __stcs(reinterpret_cast<uint4 *>(out + (size_t)n * 16), o);
        }
        __syncwarp();
        ring_shift(sm.s0, 19, nb, lane);
        ring_shift(sm.s1, 3, 2 * nb, lane);
        ring_shift(sm.s2, 1, 4 * nb, lane);
        ring_shift(sm.s3, 3, 8 * nb, lane);
        ring_shift(sm.s4, 3, 16 * nb, lane);
        ring_shift(sm.s5, 1, 32 * nb, lane);
    }
    ring_save_hist(sm.s0, rs.s0, 19, lane);
    ring_save_hist(sm.s1, rs.s1, 3, lane);
    ring_save_hist(sm.s2, rs.s2, 1, lane);
    ring_save_hist(sm.s3, rs.s3, 3, lane);
    ring_save_hist(sm.s4, rs.s4, 3, lane);
    ring_save_hist(sm.s5, rs.s5, 1, lane);
    if (lane == 0) st.wb_phase = phase;
}

// mode NONE: BasebandDataProcessor.cc:689-694 fills the block with 64
__global__ void tx_idle_kernel(const TxParams p)
{
    const int slot = blockIdx.y;
    const int sid = p.stream_ids[slot];
    uint4 *dst = reinterpret_cast<uint4 *>(p.iq + (size_t)sid * p.iq_stride);
    const size_t n16 = (size_t)p.n8 * 32;
    const uint4 v = make_uint4(0x40404040u, 0x40404040u, 0x40404040u, 0x40404040u);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = v;
}

template <int KIND>
int launch_one(const TxParams &p, cudaStream_t s)
{
    const int grid = (p.n_streams + HRD_WARPS_PER_CTA - 1) / HRD_WARPS_PER_CTA;
    const size_t smem = sizeof(SmemTx) * HRD_WARPS_PER_CTA;
    tx_kernel<KIND><<<grid, HRD_WARPS_PER_CTA * 32, smem, s>>>(p);
    return (int)cudaGetLastError();
}

} // namespace

int launch_tx(int kind, const TxParams &p, cudaStream_t s)
{
    if (p.n_streams <= 0 || p.n8 == 0) return 0;
    switch (kind) {
    case K_AM: return launch_one<K_AM>(p, s);
    case K_FM: return launch_one<K_FM>(p, s);
    case K_SSB: return launch_one<K_SSB>(p, s);
    case K_WBFM: {
        const int grid = (p.n_streams + HRD_WARPS_PER_CTA - 1) / HRD_WARPS_PER_CTA;
        const size_t smem = sizeof(SmemTxWb) * HRD_WARPS_PER_CTA;
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(tx_wbfm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            attr_set = true;
        }
        tx_wbfm_kernel<<<grid, HRD_WARPS_PER_CTA * 32, smem, s>>>(p);
        return (int)cudaGetLastError();
    }
    case K_NONE: {
        dim3 grid(8, p.n_streams);
        tx_idle_kernel<<<grid, 256, 0, s>>>(p);
        return (int)cudaGetLastError();
    }
    }
    return (int)cudaErrorInvalidValue;
}

} // namespace hrd
