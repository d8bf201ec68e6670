// hrd_tx.cu -- transmit chains: int16 PCM at 8 kS/s -> int8 I,Q at 2.048 MS/s.
//
// Replaces, per stream (reference paths relative to radioDiags/):
//   AmModulator/AmModulator.cc:366-607, FmModulator/FmModulator.cc:373-622,
//   WbFmModulator/WbFmModulator.cc:347-632, SsbModulator/SsbModulator.cc:455-707
// and underneath them Filters/Int16/{Interpolator,FirFilter}_int16.cc and
// Nco/{Nco,PhaseAccumulator}.cc.
//
// One warp owns one (stream, time tile).  Everything behind the modulator heads is an FIR, so a tile
// after the first starts HALO PCM samples early from all-zero histories (32 samples for AM / FM: the
// 40-tap stage 1 looks back 19 and the later stages less than 2 more; 64 for SSB, whose 31-tap
// Hilbert FIR sits in front) and simply does not store what the halo produces: bit-exact for any
// tile size.  The FM head's NCO phase is the one recurrence; tx_fm_phase_kernel walks it first
// (serially per stream, as the reference does) and leaves the phase before every PCM sample.
// A batch is 32 PCM samples (8192 output I,Q samples, 16 KiB):
//   1. the modulator head at 8 kS/s, one PCM sample per lane (the NCO phase recurrence is
//      the only serial piece);
//   2. interpolator stages 1..4 lane-parallel through small shared-memory rings;
//   3. the hot loop, 16 iterations: every lane takes ONE 128 kS/s sample through stages
//      5,6,7,8 in registers (the left-neighbour sample each half-band stage needs is an
//      odd-phase output, which depends on a single input, so it is recomputed locally - no
//      shuffles), narrows like (int8_t) does and writes 32 contiguous bytes with one
//      STG.E.256: 1 KiB per warp instruction.  (Since round 2 a lane takes two or four consecutive samples, and
//      the two-rail kinds -- FM, SSB, signals/ -- run stages 6,7,8 on both rails at once as fp16 pairs wherever
//      the samples are small enough for that form to be exact: see the loop itself.)
// All interpolation is the reference's Q15 arithmetic stage by stage (each stage rounds to
// int16, so stages cannot be merged): y[nL+i] = (16384 + sum_k q[i+kL]*x[n-k]) >> 15
// (Interpolator_int16.cc:398-418).
#include <cuda_fp16.h>

#include "hrd_device.cuh"

#ifndef HRD_EXP
#define HRD_EXP 0 // timing experiments only (tools/exp_build.sh)
#endif

namespace hrd {

__constant__ ConstTables c_tabtx;

void upload_tables_tx(const ConstTables &t) { cudaMemcpyToSymbol(c_tabtx, &t, sizeof t); }

namespace {

constexpr int NB8 = 32; // PCM samples per batch

__device__ __forceinline__ int lo16(uint32_t w) { return (int)(short)(w & 0xffffu); }
__device__ __forceinline__ int hi16(uint32_t w) { return ((int)w) >> 16; }
__device__ __forceinline__ uint32_t pack16(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }

// Samples per lane per iteration of the hot loop (stages 5..8): consecutive ones, so that they share ring words and
// the left-neighbour chain of the half-band stages is carried from one to the next.  Measured (4096 streams x 0.5 s):
// FM 1.931 ms (two, scalar LDS) / 1.924 (two, vector LDS) / 1.868 (four, LDS.128); SSB 1.895 / 1.917 / 1.845;
// AM 1.308 / 1.296 / 1.326 (four costs it registers: 40 -> 48).  So: four for the two-rail modulators, two with
// scalar loads for AM and for the signals/ kind (72 registers with four).  With the fp16-pair tail (HRD_TX_H2) the
// choice was measured again: FM two 1.695 / four 1.639 ms, SSB 1.655 / 1.657, signals/ PM two 1.464 / four 1.463.
#ifndef HRD_TX_SPL_FM
#define HRD_TX_SPL_FM 4
#endif
#ifndef HRD_TX_SPL_IQ
#define HRD_TX_SPL_IQ 2
#endif
// stages 6..8 of the two-rail kinds as fp16 pairs (see the hot loop of tx_kernel): 1 = on, 0 = the integer form only
#ifndef HRD_TX_H2
#define HRD_TX_H2 1
#endif
template <int KIND> struct TxSplOf {
    static constexpr int value = (KIND == K_FM || KIND == K_SSB) ? HRD_TX_SPL_FM : (KIND == K_IQ ? HRD_TX_SPL_IQ : 2);
};

struct alignas(16) SmemTx {
    alignas(16) uint32_t s4[4 + 16 * NB8];  // stage 5 input  @128k (the hot loop reads it with vector loads; one pad word)
    uint32_t s0[19 + NB8];      // stage 1 input  @8k   I/Q pairs (WBFM: PCM in the low half)
    uint32_t s1[3 + 2 * NB8];   // stage 2 input  @16k
    uint32_t s2[1 + 4 * NB8];   // stage 3 input  @32k
    uint32_t s3[3 + 8 * NB8];   // stage 4 input  @64k
    uint32_t h8[30 + NB8];      // SSB: PCM/2 history for delay line / Hilbert
};

// ---- generic polyphase stages over rings of I/Q pairs ------------------------------------
// stage 1: 40 taps, L = 2 (20 taps per branch); ring hist 19, input n at ring[19 + n]
__device__ __forceinline__ void interp40(const uint32_t *in, const int32_t *taps, int n, uint32_t &even, uint32_t &odd)
{
    unsigned ei = 1u << 14, eq = 1u << 14, oi = 1u << 14, oq = 1u << 14;
#pragma unroll
    for (int k = 0; k < 20; k++) {
        uint32_t w = in[19 + n - k];
        int xi = lo16(w), xq = hi16(w);
        ei += (unsigned)(taps[2 * k] * xi);
        eq += (unsigned)(taps[2 * k] * xq);
        oi += (unsigned)(taps[2 * k + 1] * xi);
        oq += (unsigned)(taps[2 * k + 1] * xq);
    }
    even = pack16(q15((int)ei), q15((int)eq));
    odd = pack16(q15((int)oi), q15((int)oq));
}

// stages 2,4,5: 8 taps {a,0,b,16384,b,0,a,0}, L = 2; ring hist 3, input n at ring[3 + n]:
// even = a*(x[n]+x[n-3]) + b*(x[n-1]+x[n-2]), odd = 16384*x[n-1]  (results fit int16, see hb4_*)
__device__ __forceinline__ void interp8(const uint32_t *in, int n, uint32_t &even, uint32_t &odd)
{
    const int ha = c_tabtx.tx_hb8[0], hb = c_tabtx.tx_hb8[2];
    const uint32_t w0 = in[3 + n], w1 = in[2 + n], w2 = in[1 + n], w3 = in[n];
    const int ei = ((1 << 14) + ha * (lo16(w0) + lo16(w3)) + hb * (lo16(w1) + lo16(w2))) >> 15;
    const int eq = ((1 << 14) + ha * (hi16(w0) + hi16(w3)) + hb * (hi16(w1) + hi16(w2))) >> 15;
    even = pack16(ei, eq);
    odd = pack16((lo16(w1) + 1) >> 1, (hi16(w1) + 1) >> 1);
}

// stages 3,6,7,8: taps {c, m, c, 0}, L = 2:  even = c*(x[n]+x[n-1]), odd = m*x[n].
// The host asserts m == 16384 and 0 < c <= 16384 (hrd_api.cu build_tables), so for int16 inputs
//   odd  = (16384 + 16384*x) >> 15 = (x + 1) >> 1                       (no multiply)
//   even = (16384 + c*(x + xm1)) >> 15, |even| <= 16848                  (one multiply)
// and neither can leave the int16 range: the reference's (int16_t) narrowing is the identity
// here and is not spent on (Interpolator_int16.cc:398-418 evaluated with these taps).
//
// PIPE BALANCE.  These kernels are issue-bound.  On sm_100a (tools/ubench/pipes.cu, profiles/):
// IMAD (FMA pipe) and SHF / LEA / PRMT / LOP3 (ALU pipe) each sustain 2 warp-instructions/clk/SM,
// plain adds 4 (ptxas spreads them over both pipes).  Left alone, the compiler turns the stage-8
// odd outputs, (y << 15) + 32768, into LEAs and the ALU pipe ends up at 85 % with the FMA pipe at
// 25 % (tx_kernel<SSB>).  Those 16 operations per iteration are therefore spelled as
// y * K + K with K = 32768 read from constant memory, which ptxas cannot strength-reduce: they
// issue as IMADs.  Measured on tx_kernel<SSB>, 4096 x 0.5 s: 2.15 ms -> 1.98 ms.  (Pinning the
// "+1" adds of the odd phases onto the FMA pipe as well was tried and was slower: 2.07 ms.)
__device__ __forceinline__ int hb4_even(int c, int x, int xm1) { return ((1 << 14) + c * (x + xm1)) >> 15; }
__device__ __forceinline__ int hb4_odd(int x) { return (x + 1) >> 1; }

// ---- PhaseAccumulator::run (Nco/PhaseAccumulator.cc:157-181) -------------------------------
// acc += step, then the two while loops that wrap into [-pi, pi] (hrd_device.cuh wrap_pi: fp32
// compares and an fp32 wrap that is bit-identical to the reference's double expression).
__device__ __forceinline__ float phase_advance(float acc, float step)
{
    return wrap_pi(__fadd_rn(acc, step));
}

// The same, for the 256 kS/s chain of the WBFM modulator, where up to 31 lanes advance 31
// independent accumulators in lock step and some lane wraps on almost every sample.  A divergent
// branch per sample costs > 100 cycles there (measured: the chain warp was the critical path of
// the whole CTA), so the common single wrap is evaluated on every lane and SELECTED: one opaque
// PTX block, five dependent operations per sample.  The cases the fp32 wrap does not cover
// (hrd_device.cuh: next to 2*pi, |acc| >= 12, or a second wrap needed) only raise `rare`; the
// caller then redoes its chunk with phase_advance, which handles everything.
__device__ __forceinline__ float phase_advance_lockstep(float acc, float step, unsigned &rare)
{
    float out;
    unsigned flag;
    asm("{\n\t"
        ".reg .f32 a, s, t, r, w, m;\n\t"
        ".reg .b32 ab, rb;\n\t"
        ".reg .pred pw, p1, p2, p3;\n\t"
        "add.rn.f32 a, %2, %3;\n\t"
        "abs.f32 s, a;\n\t"
        "add.rn.f32 t, s, 0fC0C90FDB;\n\t"        // |a| - 2PI_HI            (exact)
        "add.rn.f32 r, t, 0f343BBD2E;\n\t"        // ... - 2PI_LO            (one rounding)
        "mov.b32 ab, a;\n\t"
        "mov.b32 rb, r;\n\t"
        "lop3.b32 rb, rb, ab, 0x80000000, 0x78;\n\t" // rb ^ (ab & sign bit): back to a's side of zero
        "mov.b32 w, rb;\n\t"
        "setp.ge.f32 pw, s, 0f40490FDB;\n\t"      // (double)|a| > M_PI
        "selp.f32 %0, w, a, pw;\n\t"
        "abs.f32 m, t;\n\t"
        "setp.lt.f32 p1, m, 0f3A800000;\n\t"      // |t| < 2^-10: fp32 wrap not proven exact
        "setp.geu.f32 p2, s, 0f41400000;\n\t"     // !(|a| < 12)
        "abs.f32 m, r;\n\t"
        "setp.ge.f32 p3, m, 0f40490FDB;\n\t"      // still outside after one wrap
        "or.pred p1, p1, p2;\n\t"
        "or.pred p1, p1, p3;\n\t"
        "and.pred p1, p1, pw;\n\t"
        "selp.u32 %1, 1, 0, p1;\n\t"
        "}"
        : "=f"(out), "=r"(flag)
        : "f"(acc), "f"(step));
    rare |= flag;
    return out;
}


// Stages 5..8 of one rail for one 128 kS/s input sample x (with its three predecessors):
// 16 outputs as accumulators whose byte 2 is the (int8_t) value (doubled taps, see hrd_rx.cu).
// Stage 5 is the 8-tap half-band {a,0,b,16384,b,0,a,0} (AmModulator.cc:57-67; structure asserted
// on the host): its even branch is a*(x0+x3) + b*(x1+x2), its odd branch is 16384*x[n-1].
// m5 / m6 / m7 are the values "one output before" this sample's first output at stages 5 / 6 / 7 (the odd-phase
// outputs of the sample before); they come in from the caller and go out updated for the next sample, so a lane
// that takes consecutive samples computes them once.
__device__ __forceinline__ void tail4(int x0, int x1, int x2, int x3, int &m5, int &m6, int &m7, int (&out)[16])
{
    const int ha = c_tabtx.tx_hb8[0], hb = c_tabtx.tx_hb8[2];
    int y5[2];
    y5[0] = ((1 << 14) + ha * (x0 + x3) + hb * (x1 + x2)) >> 15; // |.| <= 21986: fits int16
    y5[1] = (x1 + 1) >> 1;
    const int c6 = c_tabtx.tx_c3, c7 = c_tabtx.tx_c7, c8d = 2 * c_tabtx.tx_c8;
    const int k15 = c_tabtx.k_32768;
    int y6[4];
    y6[0] = hb4_even(c6, y5[0], m5);
    y6[1] = hb4_odd(y5[0]);
    y6[2] = hb4_even(c6, y5[1], y5[0]);
    y6[3] = hb4_odd(y5[1]);
    int y7[8];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        y7[2 * k] = hb4_even(c7, y6[k], k ? y6[k - 1] : m6);
        y7[2 * k + 1] = hb4_odd(y6[k]);
    }
    // stage 8 with doubled taps: (int8_t)(acc>>15) == byte 2 of 2*acc
#pragma unroll
    for (int k = 0; k < 8; k++) {
        int left = k ? y7[k - 1] : m7;
        out[2 * k] = (1 << 15) + c8d * (y7[k] + left);
        out[2 * k + 1] = y7[k] * k15 + k15;
    }
    m5 = y5[1];
    m6 = y6[3];
    m7 = y7[7];
}
// the three carried values at the start of a run: the odd-phase chain of the sample before (x2 of the first call)
__device__ __forceinline__ void tail4_start(int x2, int &m5, int &m6, int &m7)
{
    m5 = (x2 + 1) >> 1;
    m6 = hb4_odd(m5);
    m7 = hb4_odd(m6);
}

__device__ __forceinline__ uint32_t pack_b2(int a, int b) { return __byte_perm((uint32_t)a, (uint32_t)b, 0x0062); }
__device__ __forceinline__ uint32_t merge16(uint32_t lo, uint32_t hi) { return __byte_perm(lo, hi, 0x5410); }

// ---- stages 6..8 only (WBFM: the NCO sits between stage 5 and stage 6) ----------------------
// (the integer form per rail: the reference arithmetic as written; the kernel runs tail3_h2 below unless HRD_TW_H2 == 0)
[[maybe_unused]] __device__ __forceinline__ void tail3(int x0, int xm1, int (&out)[8])
{
    const int c6 = c_tabtx.tx_c3, c7 = c_tabtx.tx_c7, c8d = 2 * c_tabtx.tx_c8;
    const int k15 = c_tabtx.k_32768;
    int y6[2], y6m1;
    y6m1 = hb4_odd(xm1);
    y6[0] = hb4_even(c6, x0, xm1);
    y6[1] = hb4_odd(x0);
    int y7[4], y7m1;
    y7m1 = hb4_odd(y6m1);
#pragma unroll
    for (int k = 0; k < 2; k++) {
        y7[2 * k] = hb4_even(c7, y6[k], k ? y6[k - 1] : y6m1);
        y7[2 * k + 1] = hb4_odd(y6[k]);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int left = k ? y7[k - 1] : y7m1;
        out[2 * k] = (1 << 15) + c8d * (y7[k] + left);
        out[2 * k + 1] = y7[k] * k15 + k15;
    }
}

// ---- stages 6..8 on BOTH rails at once (WBFM) ---------------------------------------------------------------
// The NCO samples are 11-bit ((int16_t)(cos*900)), every stage halves the range, and an integer below 2048 is an
// exact binary16 number: the two rails of one sample ride in one register as an fp16 pair and every operation
// below serves both.  Forms of an integer v:  U = v;  B+ = 1536 + v (bits 0x6600 + v: v sits in the mantissa, the
// low byte IS (int8_t)v);  B- = v - 1536.  Rounding a sum to the integer grid of [1024, 2048) is what turns an FMA
// into the reference's ">> 15" (the products never tie), and a B+ plus a B- is their exact integer sum:
//   even6 = fma(x + xm, 1053/4096, 1536)            8424/32768 = 1053/4096 is a binary16 number
//   odd(v) = fma(v + 1/2, 1/2, -1536)               = rint(v/2 + 1/4) = (v + 1) >> 1;  odd(odd(v)) = fma(v + 3/2, 1/4, ..)
//   even7: 8249/32768 needs 14 bits -> per rail in fp32, fma(s, 8249/32768, 1.5*2^23 + 0x6600): exact product, one
//          rounding, and the low 16 bits of the result are the B+ bits
//   even8 = fma(s, 1/4 + 2^-12, 1536)               c8 = 8206 only breaks the ties of s/4
//   stage-8 odd outputs: integer shifts on the B-form bits, (bits + 1) >> 1 resp. (0xe801 - bits) >> 1 per half
// tools/verify_tx_tail_h2.c proves the whole block equal to tail3 for every (x, xm) in [-995, 995]^2 (the largest square that passes).
struct TailCarry {
    __half2 xm;  // U:  the sample before
    __half2 bm;  // B-: its stage-6 odd output
    __half2 p3m; // B-: ... and the stage-7 odd output of that
};
__device__ __forceinline__ uint32_t h2_bits(__half2 v) { return *reinterpret_cast<uint32_t *>(&v); }
__device__ __forceinline__ __half2 h2_from(uint32_t w) { return *reinterpret_cast<__half2 *>(&w); }
__device__ __forceinline__ __half2 h2_const(float v) { return __float2half2_rn(v); }

__device__ __forceinline__ void tail_carry_from(__half2 xm, TailCarry &c)
{
    c.xm = xm;
    c.bm = __hfma2(__hadd2(xm, h2_const(0.5f)), h2_const(0.5f), h2_const(-1536.f));
    c.p3m = __hfma2(__hadd2(xm, h2_const(1.5f)), h2_const(0.25f), h2_const(-1536.f));
}
__device__ __forceinline__ __half2 tail_even7(__half2 s)
{
    const float c7 = 8249.0f / 32768.0f, m32 = 12582912.0f + 26112.0f;
    const uint32_t ri = __float_as_uint(__fmaf_rn(__low2float(s), c7, m32));
    const uint32_t rq = __float_as_uint(__fmaf_rn(__high2float(s), c7, m32));
    return h2_from(__byte_perm(ri, rq, 0x5410));
}
// one 256 kS/s sample (both rails) -> eight output samples {I, Q, I, Q} x 4 words
__device__ __forceinline__ void tail3_h2(__half2 x, TailCarry &c, uint32_t *o)
{
    const __half2 mg = h2_const(1536.f), nmg = h2_const(-1536.f), hf = h2_const(0.5f), qt = h2_const(0.25f);
    const __half2 c6 = h2_const(1053.0f / 4096.0f), c8 = h2_const(0.25f + 1.0f / 4096.0f);
    const __half2 a = __hfma2(__hadd2(x, c.xm), c6, mg);                             // B+  stage 6 even
    const __half2 b = __hfma2(__hadd2(x, hf), hf, nmg);                              // B-  stage 6 odd
    const __half2 p0 = tail_even7(__hadd2(a, c.bm));                                 // B+  stage 7
    const __half2 p2 = tail_even7(__hadd2(b, a));                                    // B+
    const __half2 p1 = __hfma2(__hadd2(__hadd2(a, nmg), hf), hf, nmg);               // B-
    const __half2 p3 = __hfma2(__hadd2(x, h2_const(1.5f)), qt, nmg);                 // B-
    const uint32_t e0 = h2_bits(__hfma2(__hadd2(p0, c.p3m), c8, mg));                // stage 8 even outputs, B+
    const uint32_t e1 = h2_bits(__hfma2(__hadd2(p1, p0), c8, mg));
    const uint32_t e2 = h2_bits(__hfma2(__hadd2(p2, p1), c8, mg));
    const uint32_t e3 = h2_bits(__hfma2(__hadd2(p3, p2), c8, mg));
    const uint32_t o0 = (h2_bits(p0) + 0x00010001u) >> 1, o2 = (h2_bits(p2) + 0x00010001u) >> 1; // odd outputs
    const uint32_t o1 = (0xe801e801u - h2_bits(p1)) >> 1, o3 = (0xe801e801u - h2_bits(p3)) >> 1;
    o[0] = __byte_perm(e0, o0, 0x6420); // bytes {I even, Q even, I odd, Q odd}
    o[1] = __byte_perm(e1, o1, 0x6420);
    o[2] = __byte_perm(e2, o2, 0x6420);
    o[3] = __byte_perm(e3, o3, 0x6420);
    c.xm = x;
    c.bm = b;
    c.p3m = p3;
}

// ------------------------------------------------------------------------------------
// AM / FM / SSB kernel
// ------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void ring_init(T *ring, const T *state, int hist, int lane, bool from_state)
{
    for (int i = lane; i < hist; i += 32) ring[i] = from_state ? state[i] : T();
}

template <int KIND> struct TxHaloOf { static constexpr uint32_t value = 32; };
template <> struct TxHaloOf<K_SSB> { static constexpr uint32_t value = 64; };

template <int KIND>
__global__ void __launch_bounds__(HRD_WARPS_PER_CTA * 32) tx_kernel(const TxParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int item = blockIdx.x * HRD_WARPS_PER_CTA + warp;
    if (item >= p.n_streams * p.n_tiles) return;
    const int tile = item / p.n_streams;
    const int slot = item - tile * p.n_streams;
    const int sid = p.stream_ids[slot];
    const bool first = tile == 0, last = tile == p.n_tiles - 1;
    SmemTx &sm = *reinterpret_cast<SmemTx *>(smem_raw + (size_t)warp * sizeof(SmemTx));
    const TxState &st = p.state[sid];
    const TxRail8 &rs = KIND == K_AM ? st.am : (KIND == K_FM ? st.fm : (KIND == K_IQ ? st.sig : st.ssb));
    // signals/interpolateSignal.cc has its own stage-1 prototype (:30-72); stages 2..8 are the modulators'
    const int32_t *taps40 = KIND == K_IQ ? c_tabtx.sig40 : c_tabtx.audio40;
    const int sub = KIND == K_IQ ? p.mode_of[sid] : 0;
    const int16_t *src = p.pcm + (size_t)sid * p.pcm_stride;
    int8_t *dst = p.iq + (size_t)sid * p.iq_stride;

    // this tile's range in PCM samples; a later tile rebuilds the FIR histories from zero in its halo
    const uint32_t emit_from = (uint32_t)tile * p.tile_len8;
    const uint32_t end = min(p.n8, emit_from + p.tile_len8);
    const uint32_t start = first ? 0u : emit_from - TxHaloOf<KIND>::value;

    ring_init(sm.s0, rs.s0, 19, lane, first);
    ring_init(sm.s1, rs.s1, 3, lane, first);
    ring_init(sm.s2, rs.s2, 1, lane, first);
    ring_init(sm.s3, rs.s3, 3, lane, first);
    ring_init(sm.s4, rs.s4, 3, lane, first);
    if constexpr (KIND == K_SSB) ring_init(sm.h8, st.ssb_h8, 30, lane, first);
    const float prm = (KIND == K_SSB || KIND == K_IQ) ? 0.f : p.param[sid];
    const bool lsb = (KIND == K_SSB) ? (p.lsb[sid] != 0) : true;
    const float *fm_phase = (KIND == K_FM) ? p.fm_phase + (size_t)slot * p.n8 : nullptr;
    __syncwarp();

    for (uint32_t done = start; done < end; done += NB8) {
        const int nb = (int)min((uint32_t)NB8, end - done);
        const bool emit = done >= emit_from; // halo batches compute, they do not store
        // ---- 1. modulator head, one PCM sample per lane ---------------------------------
        const int x = (lane < nb && !(KIND == K_IQ && sub == SIG_MODE_IQ8K)) ? (int)src[done + lane] : 0;
        uint32_t head = 0;
        if constexpr (KIND == K_IQ) {
            // the tool chain of signals/ (generateBaseband.sh): one of the prototype heads, or raw I,Q pairs
            if (sub == SIG_MODE_IQ8K) { // interpolateSignal.cc:266-340 reads int16 I,Q pairs
                head = (lane < nb) ? reinterpret_cast<const uint32_t *>(src)[done + lane] : 0u;
            } else if (sub == SIG_MODE_DSB) { // dsb.cc:38-47
                const int v = f32_to_i16(__fdiv_rn((float)x, 4.f));
                head = pack16(v, v);
            } else if (sub == SIG_MODE_AM_PROTO) { // am.cc:38-50: "scaledSample *= 0.8" is a double multiply
                float s = (float)((double)(float)x * 0.8);
                s = __fdiv_rn(__fadd_rn(s, 65536.f), 4.f);
                const int v = f32_to_i16(s);
                head = pack16(v, v);
            } else { // pm.cc:39-55: angle = x / 60000 * M_PI (double multiply), cosf/sinf, * 16000
                float a = __fdiv_rn((float)x, 60000.f);
                a = (float)((double)a * 3.14159265358979323846);
                // fm.cc:41-62: theta comes from the serial pre-pass (tx_fm_phase_kernel<true>)
                if (sub == SIG_MODE_FM_PROTO) a = (lane < nb) ? p.fm_phase[(size_t)slot * p.n8 + done + lane] : 0.f;
                float sf, cf;
                glibc_sincosf(a, sf, cf); // libm's cosf / sinf, bit for bit (hrd_device.cuh)
                head = pack16(f32_to_i16(__fmul_rn(cf, 16000.f)), f32_to_i16(__fmul_rn(sf, 16000.f)));
            }
        }
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
            // FmModulator.cc:596-617: the NCO phase before this sample's step (tx_fm_phase_kernel walked
            // PhaseAccumulator::run for the whole call already)
            const float my_phase = (lane < nb) ? fm_phase[done + lane] : 0.f;
            // Nco::run (Nco.cc:186-199): cosf/sinf of the float phase, as libm computes them (hrd_device.cuh
            // glibc_sincosf: bit for bit).  (Measured in round 1 with the heavier double sincos: moving it into
            // tx_fm_phase_kernel makes that kernel FP64-pipe-bound and the pair slower, 2.28 -> 2.36 ms.)
            float sf, cf;
            glibc_sincosf(my_phase, sf, cf);
            int ci = f32_to_i16(__fmul_rn(cf, 16000.f));
            int si = f32_to_i16(__fmul_rn(sf, 16000.f));
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
        for (int n = lane; n < nb; n += 32) interp40(sm.s0, taps40, n, sm.s1[3 + 2 * n], sm.s1[3 + 2 * n + 1]);
        __syncwarp();
        for (int n = lane; n < 2 * nb; n += 32) interp8(sm.s1, n, sm.s2[1 + 2 * n], sm.s2[1 + 2 * n + 1]);
        __syncwarp();
        for (int n = lane; n < 4 * nb; n += 32) {
            uint32_t w = sm.s2[1 + n], wm = sm.s2[n];
            const int c = c_tabtx.tx_c3;
            sm.s3[3 + 2 * n] = pack16(hb4_even(c, lo16(w), lo16(wm)), hb4_even(c, hi16(w), hi16(wm)));
            sm.s3[3 + 2 * n + 1] = pack16((lo16(w) + 1) >> 1, (hi16(w) + 1) >> 1);
        }
        __syncwarp();
        for (int n = lane; n < 8 * nb; n += 32) interp8(sm.s3, n, sm.s4[3 + 2 * n], sm.s4[3 + 2 * n + 1]);
        __syncwarp();

        // ---- 3. hot loop: stages 5..8, 32 bytes per lane per iteration ------------------------
        int8_t *out = dst + (size_t)done * 512;
        // a lane takes TX_SPL consecutive 128 kS/s samples (32 output bytes each): TX_SPL + 3 ring words
        // instead of 4 per sample, and each sample's left-neighbour chain is the tail of the one before
        constexpr int TX_SPL = TxSplOf<KIND>::value;
        for (int n = TX_SPL * lane; n < 16 * nb; n += 32 * TX_SPL) {
            uint32_t w[TX_SPL + 3]; // w[j] = s4[n + j]: sample n + h reads w[h + 3] (newest) .. w[h]
            if constexpr (TX_SPL == 4) { // s4 is 16-byte aligned and n a multiple of four: two LDS.128
                const uint4 a = *reinterpret_cast<const uint4 *>(sm.s4 + n), b = *reinterpret_cast<const uint4 *>(sm.s4 + n + 4);
                w[0] = a.x, w[1] = a.y, w[2] = a.z, w[3] = a.w, w[4] = b.x, w[5] = b.y, w[6] = b.z;
            } else {
#pragma unroll
                for (int j = 0; j < 5; j++) w[j] = sm.s4[n + j];
            }
#if HRD_TX_H2
            // STAGES 6..8 ON BOTH RAILS AT ONCE (tail3_h2, the form tx_wbfm_kernel runs): stage 5 stays integer per rail
            // (its inputs need 12-13 bits), its two outputs per sample go into one register as an fp16 pair, and
            // every later operation serves I and Q together -- 91 instead of 132 instructions per 128 kS/s sample.
            // The packed form is exact for stage-6 inputs up to +-995 (tools/verify_tx_tail_h2.c, every pair); the
            // modulators' rails stay near 500 there (16000 / 32), but nothing in the arithmetic FORBIDS more (raw
            // int16 I,Q pairs of the signals/ kind reach 1024 and beyond), so the warp checks its samples and takes
            // the integer form below when any lane is out of range: identical results either way.
            if constexpr (KIND != K_AM) {
                const int ha = c_tabtx.tx_hb8[0], hb = c_tabtx.tx_hb8[2];
                int xi[TX_SPL + 3], xq[TX_SPL + 3];
#pragma unroll
                for (int j = 0; j < TX_SPL + 3; j++) xi[j] = lo16(w[j]), xq[j] = hi16(w[j]);
                __half2 ev[TX_SPL], od[TX_SPL];
                const __half2 xm0 = __halves2half2(__int2half_rn((xi[1] + 1) >> 1), __int2half_rn((xq[1] + 1) >> 1));
                __half2 worst = __habs2(xm0);
#pragma unroll
                for (int h = 0; h < TX_SPL; h++) {
                    const int ei = ((1 << 14) + ha * (xi[h + 3] + xi[h]) + hb * (xi[h + 2] + xi[h + 1])) >> 15;
                    const int eq = ((1 << 14) + ha * (xq[h + 3] + xq[h]) + hb * (xq[h + 2] + xq[h + 1])) >> 15;
                    ev[h] = __halves2half2(__int2half_rn(ei), __int2half_rn(eq));
                    od[h] = __halves2half2(__int2half_rn((xi[h + 2] + 1) >> 1), __int2half_rn((xq[h + 2] + 1) >> 1));
                    worst = __hmax2(worst, __hmax2(__habs2(ev[h]), __habs2(od[h])));
                }
                // (the lanes of this iteration vote -- a short batch's last iteration runs on some lanes only -- so that a
                //  warp never runs both forms: measured on SSB with full-scale noise and square waves in the batch,
                //  1.66 ms with the vote against 1.82 ms with a decision per lane)
                if (__all_sync(__activemask(), __hble2(worst, h2_const(995.f)))) {
                    TailCarry tc;
                    tail_carry_from(xm0, tc);
#pragma unroll
                    for (int h = 0; h < TX_SPL; h++) {
                        u32x8 o;
                        tail3_h2(ev[h], tc, o.v);
                        tail3_h2(od[h], tc, o.v + 4);
                        if (emit) stg_stream_256(out + (size_t)(n + h) * 32, o);
                    }
                    continue;
                }
            }
#endif
            int mi5, mi6, mi7, mq5, mq6, mq7;
            tail4_start(lo16(w[1]), mi5, mi6, mi7);
            if constexpr (KIND != K_AM) tail4_start(hi16(w[1]), mq5, mq6, mq7);
#pragma unroll
            for (int h = 0; h < TX_SPL; h++) {
                int oi[16];
                tail4(lo16(w[h + 3]), lo16(w[h + 2]), lo16(w[h + 1]), lo16(w[h]), mi5, mi6, mi7, oi);
                u32x8 o;
                if constexpr (KIND == K_AM) { // both rails carry the same samples (AmModulator.cc:601-602)
#pragma unroll
                    for (int k = 0; k < 8; k++) o.v[k] = __byte_perm((uint32_t)oi[2 * k], (uint32_t)oi[2 * k + 1], 0x6622);
                } else {
                    int oq[16];
                    tail4(hi16(w[h + 3]), hi16(w[h + 2]), hi16(w[h + 1]), hi16(w[h]), mq5, mq6, mq7, oq);
#pragma unroll
                    for (int k = 0; k < 8; k++) // bytes {I[2k], Q[2k], I[2k+1], Q[2k+1]}
                        o.v[k] = merge16(pack_b2(oi[2 * k], oq[2 * k]), pack_b2(oi[2 * k + 1], oq[2 * k + 1]));
                }
                if (emit) stg_stream_256(out + (size_t)(n + h) * 32, o);
            }
        }
        __syncwarp();
        ring_shift(sm.s0, 19, nb, lane);
        ring_shift(sm.s1, 3, 2 * nb, lane);
        ring_shift(sm.s2, 1, 4 * nb, lane);
        ring_shift(sm.s3, 3, 8 * nb, lane);
        ring_shift(sm.s4, 3, 16 * nb, lane);
        if constexpr (KIND == K_SSB) ring_shift(sm.h8, 30, nb, lane);
    }

    if (!last) return;
    // the last tile leaves the interpolator histories for the next call (the rest of the record was
    // copied over by the host; the FM phase is tx_fm_phase_kernel's)
    TxState &so = p.state_out[sid];
    TxRail8 &ro = KIND == K_AM ? so.am : (KIND == K_FM ? so.fm : (KIND == K_IQ ? so.sig : so.ssb));
    ring_save_hist(sm.s0, ro.s0, 19, lane);
    ring_save_hist(sm.s1, ro.s1, 3, lane);
    ring_save_hist(sm.s2, ro.s2, 1, lane);
    ring_save_hist(sm.s3, ro.s3, 3, lane);
    ring_save_hist(sm.s4, ro.s4, 3, lane);
    if constexpr (KIND == K_SSB) ring_save_hist(sm.h8, so.ssb_h8, 30, lane);
    // stages 6,7,8 keep one input sample each; it is always the last output of the stage
    // before, which tail4 recomputes from s4's history, so nothing more needs saving.
}

// ------------------------------------------------------------------------------------
// WBFM (WbFmModulator.cc:347-632): stages 1..5 on the real PCM, NCO at 256 kS/s, stages 6..8
// on I and Q.
// ------------------------------------------------------------------------------------
// The NCO phase (PhaseAccumulator.cc:157-181) is a float accumulation at 256 kS/s with no decay:
// phase[n+1] = wrap(fl(phase[n] + step[n])).  It cannot be cut in time and it cannot be
// re-associated, so it is evaluated serially per stream -- but TRANSPOSED, exactly like the
// de-emphasis recurrence of the WBFM demodulator (hrd_rx.cu rx_wbfm_kernel):
//   CTA = up to 32 warps.  All but the last own one stream each: PCM -> stages 1..5 -> phase step per
//   256 kS/s sample (-> shared memory), and after the chain: table index -> (cos,sin)*900 as
//   int16 -> stages 6..8 on both rails -> 32-byte stores.  The last warp is the chain warp: lane r
//   walks stream r's row of TW_STEP phase steps in place, leaving the phase BEFORE each step.
//   Two row buffers, one __syncthreads per step of 8 PCM samples.
constexpr int TW_ITEMS = 31;
constexpr int TW_STEP8 = 8;               // PCM samples per pipeline step
constexpr int TW_STEP = TW_STEP8 * 32;    // 256 kS/s samples per step
constexpr int TW_PITCH = TW_STEP + 4;     // floats per row: conflict-free LDS.128 by row
constexpr int TW_SUPER8 = 32;             // PCM samples per super-step: stages 1 and 2 run once per four steps, on all lanes
constexpr int TW_LUT_HALF = 2048;         // the phase-step table covers stage-5 outputs in [-2048, 2047]
constexpr uint32_t TW_FOLD = 8193;        // entries per half of the folded NCO table

// Per item: what stages 1 and 2 need, and the 32 kS/s samples of the current super-step.  Stages 3, 4 and 5
// keep NO ring: a lane owns one 32 kS/s sample per step (eight 256 kS/s samples) and recomputes the few
// 64 and 128 kS/s values before its own from the three 32 kS/s samples before it -- no shuffles, no ring
// shifts, no __syncwarp between the stages, all lanes busy.
struct SmemTwItem {                       // real samples, sign-extended
    alignas(16) int32_t x32[4 + 4 * TW_SUPER8]; // stage 3 input @32k: [1..3] history, [4 + n] sample n of the super-step
    int32_t s0[19 + TW_SUPER8];           // stage 1 input @8k
    int32_t s1[3 + 2 * TW_SUPER8 + 2];    // stage 2 input @16k (+2: the struct stays a multiple of 16 bytes)
};
struct SmemTw {
    float ph[2][32][TW_PITCH];
    SmemTwItem item[TW_ITEMS];
    uint32_t big[2][32];                  // per row and step: some |phase step| >= 3 (the chain's slow path)
    // phase step by stage-5 output value, for a CTA whose streams share one deviation.  It sits in the MIDDLE of
    // the block on purpose: a value outside the table is detected after its (discarded) load, which then still
    // falls inside the CTA's shared memory (|stage-5 output| <= 5091, see produce).
    float step_lut[2 * TW_LUT_HALF];
    alignas(16) uint32_t iqfold[2 * TW_FOLD + 2]; // {(int16_t)(cos*900), (int16_t)(sin*900)}: [k] phase >= 0, [8193 + k] phase < 0
    float thr[8194 + 2];                  // nco_fold_offset thresholds
    uint32_t lut_dev;                     // the shared deviation (bits)
    int32_t lut_on;                       // every live stream of the CTA has that deviation
    int32_t e_lim;                        // |stage-5 output| below which the table applies and |step| < 3
};
static_assert(sizeof(SmemTwItem) % 16 == 0, "items are 16-byte aligned");
static_assert(sizeof(SmemTw) <= 227 * 1024, "tx_wbfm_kernel shared memory");

__device__ __forceinline__ void load_real_hist(int32_t *ring, const uint32_t *state, int hist, int lane)
{
    for (int i = lane; i < hist; i += 32) ring[i] = lo16(state[i]);
}
__device__ __forceinline__ void save_real_hist(const int32_t *ring, uint32_t *state, int hist, int lane)
{
    for (int i = lane; i < hist; i += 32) state[i] = (uint32_t)ring[i] & 0xffffu;
}

// 8-tap half-band {a,0,b,16384,b,0,a,0} (stages 2, 4, 5; Interpolator_int16.cc:398-418 with these taps): input n
// gives the even output a*(x[n]+x[n-3]) + b*(x[n-1]+x[n-2]) and the odd output 16384*x[n-1], i.e. (x[n-1]+1)>>1
__device__ __forceinline__ int hb8_even(int ha, int hb, int x0, int x1, int x2, int x3)
{
    return ((1 << 14) + ha * (x0 + x3) + hb * (x1 + x2)) >> 15;
}
// the same on a real ring (hist 3): input n -> outputs 2n, 2n+1
__device__ __forceinline__ void interp8_real(const int32_t *in, int n, int &even, int &odd)
{
    even = hb8_even(c_tabtx.tx_hb8[0], c_tabtx.tx_hb8[2], in[3 + n], in[2 + n], in[1 + n], in[n]);
    odd = (in[2 + n] + 1) >> 1;
}

// fl32(fl64(a / 256000.0)) -- PhaseAccumulator::setFrequency's double division narrowed to
// float -- without dividing in the common case.  r = a * (1/256000) is within 3 ulp64 of the
// correctly rounded quotient; the two round to the same float unless a float rounding midpoint
// (double mantissa bits 28..0 == 0x10000000) lies that close to r, or the result is not a
// normal float.  Only then is the real division evaluated.
// out of line on purpose: inlined, the compiler hoists most of the division sequence above the
// `risky` test and every lane pays for it (it was 10 % of the kernel's instructions)
__device__ __noinline__ float div_256000_exact(double a) { return (float)(a / 256000.0); }
__device__ __noinline__ float div_8000_exact(double a) { return (float)(a / 8000.0); }

// (the argument holds for any divisor: r is a * fl(1/D), two roundings away from the exact quotient)
template <int D>
__device__ __forceinline__ float div_const_to_float(double a)
{
    static_assert(D == 256000 || D == 8000, "one out-of-line exact division per divisor");
    double r = a * (1.0 / (double)D);
    const uint32_t lo = (uint32_t)__double2loint(r);
    const uint32_t e = ((uint32_t)__double2hiint(r) >> 20) & 0x7ffu;
    const bool risky = ((lo & 0x1fffffffu) - 0x0ffffff8u) <= 16u || (e - 898u) > 250u;
    if (risky && a != 0.0) return D == 256000 ? div_256000_exact(a) : div_8000_exact(a);
    return (float)r;
}

// WbFmModulator.cc:596-604 + PhaseAccumulator.cc:103 for one 256 kS/s sample e of the interpolated PCM:
//   ncoFrequency = frequencyDeviation * (float)e / 1024   (dividing by 2^10 is exact scaling)
//   phaseStepSize = (float)((2*M_PI*ncoFrequency)/256000.0)
__device__ __forceinline__ float wb_phase_step(float dev, int e)
{
    const float f = __fmul_rn(__fmul_rn(dev, (float)e), 0.0009765625f);
    return div_const_to_float<256000>(2.0 * 3.14159265358979323846 * (double)f);
}

// Nco::runFast's table index (Nco.cc:231-248): (int16_t)((double)(phase*16384.0f)/(2*M_PI)) + 8192, clamped to
// [0,16383].  The double division is replaced by ONE comparison against a table of thresholds built on the host
// WITH that very expression: T[k] = the smallest float p >= 0 whose reference index offset is >= k (hrd_api.cu;
// T[0] = 0, T[8193] = +inf).  An FMA with a slightly low constant and the addend 2^23 - 0.5 leaves k or k + 1 in
// the mantissa (never less, never more: tools/verify_fp_tricks.c checks every float in [0, pi]), so
// k = k1 - (|phase| < T[k1]).  Truncation toward zero makes negative phases mirror images: the table of
// (cos, sin) * 900 pairs is stored folded, [k] for phase >= 0 and [8193 + k] for phase < 0 (hrd_api.cu), and the
// result here is the BYTE offset into it.
#define HRD_NCO_C_LO 2607.5920f
__device__ __forceinline__ uint32_t nco_fold_offset(float phase, const float *T)
{
    const float ap = fabsf(phase);
    uint32_t k4 = (uint32_t)__float_as_int(__fmaf_rn(ap, HRD_NCO_C_LO, 8388607.5f)) * 4u - 0x4affffffu * 4u;
    k4 = min(k4, 8193u * 4u); // (NaN or a phase outside [-pi, pi] must not leave the table)
    const float t = *reinterpret_cast<const float *>(reinterpret_cast<const char *>(T) + k4);
    if (ap < t) k4 -= 4u;
    if (phase < 0.0f) k4 += TW_FOLD * 4u;
    return k4;
}

__global__ void __launch_bounds__(1024, 1) tx_wbfm_kernel(const TxParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemTw &sm = *reinterpret_cast<SmemTw *>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // WARP ROLES.  The chain (four dependent operations per sample) is the CTA's critical path, and what slows it
    // under load is not latency but losing the issue slot to the item warps of its scheduler (ncu: the chain warp
    // sat in "not selected" a third of the time when it shared a scheduler with seven item warps).  A warp's
    // scheduler is warp id % 4, so the CTA always has 32 warp slots: the chain is warp 31 (scheduler 3, and the
    // arbiter favours the highest warp id), items fill schedulers 0, 1, 2 first (24 slots) and only the items
    // beyond 24 join scheduler 3; unused slots help load the tables and leave.
    const bool chain_warp = warp == 31;
    const int item_of_warp = (warp & 3) != 3 ? (warp >> 2) * 3 + (warp & 3) : 24 + (warp >> 2);
    const int row = chain_warp ? lane : item_of_warp;
    const int slot = blockIdx.x * p.items_per_cta + row;
    const bool member = chain_warp || item_of_warp < p.items_per_cta; // takes part in the pipeline's barriers
    const bool live = row < p.items_per_cta && slot < p.n_streams;
    const int sid = live ? p.stream_ids[slot] : 0;
    TxState &st = p.state_out[sid]; // == state[sid]: the host copied the records over before the launch
    TxRail8 &rs = st.wb;
    SmemTwItem &it = sm.item[chain_warp || !member ? 0 : item_of_warp];
    const int16_t *src = p.pcm + (size_t)sid * p.pcm_stride;
    int8_t *dst = p.iq + (size_t)sid * p.iq_stride;
    const uint32_t n_steps = (p.n8 + TW_STEP8 - 1) / TW_STEP8;

    for (int i = threadIdx.x; i < (int)(2 * TW_FOLD); i += blockDim.x) sm.iqfold[i] = __ldg(p.nco_iq900 + i);
    for (int i = threadIdx.x; i < 8194; i += blockDim.x) sm.thr[i] = __ldg(p.nco_thr + i);

    float phase = 0.f, dev = 0.f;
    uint32_t iq_keep = 0; // the (cos,sin)*900 pair of the previous 256 kS/s sample (stage 6 history)
    int pcm_next = 0;     // this lane's PCM sample of the NEXT super-step, loaded one super-step ahead
    if (live) {
        dev = p.param[sid];
        if (!chain_warp && (uint32_t)lane < p.n8) pcm_next = (int)src[lane];
        if (chain_warp) {
            phase = st.wb_phase;
        } else {
            load_real_hist(it.s0, rs.s0, 19, lane);
            load_real_hist(it.s1, rs.s1, 3, lane);
            if (lane < 3) it.x32[1 + lane] = lo16(rs.s3[lane]); // the last three 32 kS/s samples
            iq_keep = rs.s5[0];
        }
    }
    if (chain_warp) { // do all streams of this CTA share one deviation?  (lane 0 always has a stream)
        const uint32_t d0 = __shfl_sync(HRD_FULL_MASK, __float_as_uint(dev), 0);
        const bool same = __all_sync(HRD_FULL_MASK, !live || __float_as_uint(dev) == d0);
        if (lane == 0) {
            sm.lut_dev = d0;
            sm.lut_on = same;
            sm.e_lim = TW_LUT_HALF;
        }
    }
    __syncthreads();
    // The phase step is a pure function of (deviation, stage-5 output), and the interpolated PCM stays small: the
    // reference's own double expression is evaluated ONCE per value here (IEEE division) instead of once per sample.
    const bool lut_on = sm.lut_on != 0;
    if (lut_on) {
        const float d = __uint_as_float(sm.lut_dev);
        for (int i = threadIdx.x; i < 2 * TW_LUT_HALF; i += blockDim.x) {
            const int e = i - TW_LUT_HALF;
            const float f = __fmul_rn(__fmul_rn(d, (float)e), 0.0009765625f);
            const float stp = (float)((2.0 * 3.14159265358979323846 * (double)f) / 256000.0);
            sm.step_lut[i] = stp;
            if (!(fabsf(stp) < 3.0f)) atomicMin(&sm.e_lim, abs(e)); // the chain's fast path needs |step| < 3
        }
    }
    __syncthreads();
    if (!member) return; // (a spare warp slot: exited warps do not count at later barriers)
    // the table applies to |e| < e_lim: index (e + e_lim - 1) against the bound 2 e_lim - 1, one unsigned compare
    const int e_lim4 = (sm.e_lim - 1) * 4;
    const uint32_t e_bound4 = (uint32_t)max(2 * sm.e_lim - 1, 0) * 4u;
    const char *lut_base = reinterpret_cast<const char *>(sm.step_lut) + (TW_LUT_HALF * 4 - e_lim4);

    // WbFmModulator.cc:389-441 + 596-604 on step t: PCM -> stages 1..5 -> phase steps
    auto produce = [&](uint32_t t) {
        const uint32_t j = t & 3;
        if (j == 0) { // stages 1 and 2 for the next (up to) 32 PCM samples, one PCM sample per lane
            const uint32_t done = t * TW_STEP8;
            const int nb8 = (int)min((uint32_t)TW_SUPER8, p.n8 - done);
            if (t) { // histories to the front (every super-step but the last is a full one)
                ring_shift(it.s0, 19, TW_SUPER8, lane);
                ring_shift(it.s1, 3, 2 * TW_SUPER8, lane);
                int v = 0;
                if (lane < 3) v = it.x32[1 + 4 * TW_SUPER8 + lane];
                __syncwarp();
                if (lane < 3) it.x32[1 + lane] = v;
            }
            if (lane < nb8) it.s0[19 + lane] = pcm_next;
            if (done + TW_SUPER8 + lane < p.n8) pcm_next = (int)src[done + TW_SUPER8 + lane];
            __syncwarp();
            if (lane < nb8) { // stage 1: 40 taps, L = 2 (the only stage whose output may wrap: keep q15)
                unsigned e = 1u << 14, o = 1u << 14;
#pragma unroll
                for (int k = 0; k < 20; k++) {
                    const int x = it.s0[19 + lane - k];
                    e += (unsigned)(c_tabtx.audio40[2 * k] * x);
                    o += (unsigned)(c_tabtx.audio40[2 * k + 1] * x);
                }
                it.s1[3 + 2 * lane] = q15((int)e);
                it.s1[3 + 2 * lane + 1] = q15((int)o);
            }
            __syncwarp();
            if (lane < nb8) { // stage 2: two inputs, four outputs
                int4 y;
                interp8_real(it.s1, 2 * lane, y.x, y.y);
                interp8_real(it.s1, 2 * lane + 1, y.z, y.w);
                *reinterpret_cast<int4 *>(it.x32 + 4 + 4 * lane) = y;
            }
            __syncwarp();
        }
        // stages 3, 4, 5 and the phase steps: lane <-> 32 kS/s sample m = 32 j + lane of the super-step
        const int nb = (int)min((uint32_t)TW_STEP8, p.n8 - t * TW_STEP8);
        const bool valid = lane < 4 * nb;
        const int32_t *xr = it.x32 + 4 + 32 * j + lane;
        int xa = xr[0], xb = xr[-1], xc = xr[-2], xd = xr[-3];
        if (!valid) xa = xb = xc = xd = 0; // (stale ring words must not reach the table index)
        const int c3 = c_tabtx.tx_c3, ha = c_tabtx.tx_hb8[0], hb = c_tabtx.tx_hb8[2];
        // stage 3 @64k: u[i] = y3[2m - 4 + i]
        const int u0 = hb4_even(c3, xc, xd), u1 = hb4_odd(xc), u2 = hb4_even(c3, xb, xc), u3 = hb4_odd(xb);
        const int u4 = hb4_even(c3, xa, xb), u5 = hb4_odd(xa);
        // stage 4 @128k: v[i] = y4[4m - 3 + i]
        int v[7];
        v[0] = hb4_odd(u1);
        v[1] = hb8_even(ha, hb, u3, u2, u1, u0);
        v[2] = hb4_odd(u2);
        v[3] = hb8_even(ha, hb, u4, u3, u2, u1);
        v[4] = hb4_odd(u3);
        v[5] = hb8_even(ha, hb, u5, u4, u3, u2);
        v[6] = hb4_odd(u4);
        // stage 5 @256k: e[i] = y5[8m + i].  |e| <= 5091: stage 1 wraps into int16, the 8-tap stages scale a
        // bound by 2(|a|+|b|)/32768 <= 0.671 and stage 3 by 2c/32768 <= 0.515.
        int e[8];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            e[2 * q] = hb8_even(ha, hb, v[3 + q], v[2 + q], v[1 + q], v[q]);
            e[2 * q + 1] = hb4_odd(v[2 + q]);
        }
        float stp[8];
        bool big = false;
        bool slow = !lut_on;
        if (lut_on) {
            bool out = false;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint32_t ix = (uint32_t)(e[i] * 4 + e_lim4);
                out |= ix >= e_bound4;
                stp[i] = *reinterpret_cast<const float *>(lut_base + ix);
            }
            slow = out;
        }
        if (slow) { // mixed deviations in this CTA, or a sample beyond the table / a step of 3 rad or more
#pragma unroll
            for (int i = 0; i < 8; i++) {
                stp[i] = wb_phase_step(dev, e[i]);
                big |= !(fabsf(stp[i]) < 3.0f);
            }
        }
        float *out = sm.ph[t & 1][row] + 8 * lane;
        *reinterpret_cast<float4 *>(out) = make_float4(stp[0], stp[1], stp[2], stp[3]);
        *reinterpret_cast<float4 *>(out + 4) = make_float4(stp[4], stp[5], stp[6], stp[7]);
        big = __any_sync(HRD_FULL_MASK, big && valid);
        if (lane == 0) sm.big[t & 1][row] = big;
    };

    // PhaseAccumulator::run for every sample of step t: row[n] <- phase before step n.
    // The item warps flag rows with a step of 3 rad or more (a deviation setting beyond the reference's
    // limits, or NaN).  Without one, |phase + step| < pi + 3 < 6.2, so the single-FMA wrap is the exact one
    // (hrd_device.cuh phase_step_fast) and the chain is three dependent operations per sample with no
    // bookkeeping; with one, the lock-step version with its exact per-chunk redo runs instead.
    const ChainConsts cc = {c_tabtx.k_sign, c_tabtx.k_m2pi_hi, c_tabtx.k_m2pi_lo};
    auto chain = [&](uint32_t t) {
        const uint32_t nb = min((uint32_t)TW_STEP8, p.n8 - t * TW_STEP8) * 32;
        float *r = sm.ph[t & 1][lane];
        // (every lane of the chain warp votes; lanes without a stream then leave)
        const bool slow = __any_sync(HRD_FULL_MASK, live && (sm.big[t & 1][lane] != 0 || !(fabsf(phase) < HRD_PI_UP)));
        if (!live) return;
        if (!slow) {
            // the row is read four groups (16 samples) ahead of its use.  (Measured: 2.07 ms against 2.22 ms with all
            // eight loads of a 32-sample round issued first; in rx_wbfm_kernel, whose chain is two operations per
            // sample instead of four, the same change lost.)
            float4 q[4];
#pragma unroll
            for (int g = 0; g < 4; g++) q[g] = *reinterpret_cast<float4 *>(r + 4 * g);
            for (uint32_t c = 0; c < nb; c += 16) {
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    float4 v = q[g];
                    if (c + 16 < nb) q[g] = *reinterpret_cast<float4 *>(r + c + 16 + 4 * g);
                    float s;
                    s = v.x; v.x = phase; phase = phase_step_fast(phase, s, cc);
                    s = v.y; v.y = phase; phase = phase_step_fast(phase, s, cc);
                    s = v.z; v.z = phase; phase = phase_step_fast(phase, s, cc);
                    s = v.w; v.w = phase; phase = phase_step_fast(phase, s, cc);
                    *reinterpret_cast<float4 *>(r + c + 4 * g) = v;
                }
            }
            return;
        }
        for (uint32_t c = 0; c < nb; c += 32) {
            float4 v[8];
#pragma unroll
            for (int g = 0; g < 8; g++) v[g] = *reinterpret_cast<float4 *>(r + c + 4 * g);
            const float phase0 = phase;
            unsigned rare = 0;
#pragma unroll
            for (int g = 0; g < 8; g++) {
                float s;
                s = v[g].x; v[g].x = phase; phase = phase_advance_lockstep(phase, s, rare);
                s = v[g].y; v[g].y = phase; phase = phase_advance_lockstep(phase, s, rare);
                s = v[g].z; v[g].z = phase; phase = phase_advance_lockstep(phase, s, rare);
                s = v[g].w; v[g].w = phase; phase = phase_advance_lockstep(phase, s, rare);
            }
            if (rare) { // redo this lane's 32 samples with the reference's own loops (steps still in the row)
                phase = phase0;
                for (int i = 0; i < 32; i++) {
                    const float s = r[c + i];
                    r[c + i] = phase;
                    phase = phase_advance(phase, s);
                }
            } else {
#pragma unroll
                for (int g = 0; g < 8; g++) *reinterpret_cast<float4 *>(r + c + 4 * g) = v[g];
            }
        }
    };

    // Nco::runFast + x900 (WbFmModulator.cc:606-626), stages 6..8 (:471-531): two 256 kS/s samples
    // per lane -> 16 output samples = 32 bytes
    // two samples of a lane: table words w0, w1 and the word before them -> 16 output samples = 32 bytes
    auto emit_pair = [&](uint32_t wm, uint32_t w0, uint32_t w1, int8_t *at) {
        u32x8 o;
#if HRD_TW_H2
        TailCarry c;
        tail_carry_from(h2_from(wm), c);
        tail3_h2(h2_from(w0), c, o.v);
        tail3_h2(h2_from(w1), c, o.v + 4);
#else
        int oi[8], oq[8];
        tail3(lo16(w0), lo16(wm), oi);
        tail3(hi16(w0), hi16(wm), oq);
#pragma unroll
        for (int k = 0; k < 4; k++)
            o.v[k] = merge16(pack_b2(oi[2 * k], oq[2 * k]), pack_b2(oi[2 * k + 1], oq[2 * k + 1]));
        tail3(lo16(w1), lo16(w0), oi);
        tail3(hi16(w1), hi16(w0), oq);
#pragma unroll
        for (int k = 0; k < 4; k++)
            o.v[4 + k] = merge16(pack_b2(oi[2 * k], oq[2 * k]), pack_b2(oi[2 * k + 1], oq[2 * k + 1]));
#endif
        stg_stream_256(at, o);
    };
    // Nco::runFast + x900 (WbFmModulator.cc:606-626), stages 6..8 (:471-531): two 256 kS/s samples
    // per lane -> 16 output samples = 32 bytes
    auto consume = [&](uint32_t t) {
        const uint32_t done = t * TW_STEP8;
        const int nb = (int)min((uint32_t)TW_STEP8, p.n8 - done) * 32;
        const float *ph = sm.ph[t & 1][row];
        const char *fold = reinterpret_cast<const char *>(sm.iqfold);
        int8_t *out = dst + (size_t)done * 512;
        if (nb == TW_STEP) { // a full step: four rounds, every lane busy, nothing to clip
#pragma unroll 1
            for (int base = 0; base < TW_STEP; base += 64) {
                const int n = base + 2 * lane;
                const float2 pp = *reinterpret_cast<const float2 *>(ph + n);
                const uint32_t w0 = *reinterpret_cast<const uint32_t *>(fold + nco_fold_offset(pp.x, sm.thr));
                const uint32_t w1 = *reinterpret_cast<const uint32_t *>(fold + nco_fold_offset(pp.y, sm.thr));
                // the sample before w0: previous lane's w1 (lane 0: kept from the previous round)
                const uint32_t wm = __shfl_sync(HRD_FULL_MASK, (lane == 31) ? iq_keep : w1, (lane + 31) & 31);
                iq_keep = __shfl_sync(HRD_FULL_MASK, w1, 31);
                emit_pair(wm, w0, w1, out + (size_t)n * 16);
            }
            return;
        }
        for (int base = 0; base < nb; base += 64) { // warp-uniform trip count: shuffles inside
            const int n = base + 2 * lane;
            const bool valid = n < nb;               // a short last step leaves the upper lanes idle
            const int last_lane = min(32, (nb - base) / 2) - 1;
            const float2 pp = *reinterpret_cast<const float2 *>(ph + (valid ? n : 0));
            const uint32_t w0 = *reinterpret_cast<const uint32_t *>(fold + nco_fold_offset(pp.x, sm.thr));
            const uint32_t w1 = *reinterpret_cast<const uint32_t *>(fold + nco_fold_offset(pp.y, sm.thr));
            const uint32_t sel = (lane == 31) ? iq_keep : w1;
            const uint32_t wm = __shfl_sync(HRD_FULL_MASK, sel, (lane + 31) & 31);
            iq_keep = __shfl_sync(HRD_FULL_MASK, w1, last_lane);
            if (valid) emit_pair(wm, w0, w1, out + (size_t)n * 16);
        }
    };

    // hand-over by named barriers, as in rx_wbfm_kernel (hrd_rx.cu): item warps arrive on "produced"
    // and wait on "chained"; the chain warp does the opposite
    const int bar_threads = (p.items_per_cta + 1) * 32;
    if (chain_warp) {
        for (uint32_t t = 0; t < n_steps; t++) {
            named_bar_sync(HRD_BAR_PRODUCED, t, bar_threads);
            chain(t);
            named_bar_arrive(HRD_BAR_CHAINED, t, bar_threads);
        }
    } else {
        if (live) produce(0);
        named_bar_arrive(HRD_BAR_PRODUCED, 0, bar_threads);
        for (uint32_t t = 0; t < n_steps; t++) {
            if (t + 1 < n_steps) {
                if (live) produce(t + 1);
                named_bar_arrive(HRD_BAR_PRODUCED, t + 1, bar_threads);
            }
            named_bar_sync(HRD_BAR_CHAINED, t, bar_threads);
            if (live) consume(t);
        }
    }
    __syncthreads();
    if (live) {
        if (chain_warp) {
            st.wb_phase = phase;
        } else {
            // the histories sit behind the samples of the last super-step
            const int nb8 = (int)(p.n8 - ((n_steps - 1) / 4) * TW_SUPER8);
            save_real_hist(it.s0 + nb8, rs.s0, 19, lane);
            save_real_hist(it.s1 + 2 * nb8, rs.s1, 3, lane);
            save_real_hist(it.x32 + 1 + 4 * nb8, rs.s3, 3, lane);
            if (lane == 0) rs.s5[0] = iq_keep;
        }
    }
}

// ------------------------------------------------------------------------------------
// FM: the NCO phase recurrence at 8 kS/s (FmModulator.cc:596-604, PhaseAccumulator.cc:95-181)
// ------------------------------------------------------------------------------------
// phase[n+1] = wrap(fl(phase[n] + step[n])) has no decay and cannot be cut in time, so every stream's
// call is one dependent chain of n8 additions (about 22 cycles each).  The first version gave a warp to
// every stream and broadcast the steps by shuffle: 10 warp instructions per sample, 0.34 ms for 4096
// streams x 0.5 s, 15 % of the FM modulator's time (profiles/).  Here the recurrences of 32 streams
// are transposed onto the lanes of ONE chain warp, as in the WBFM kernels: a CTA owns 32 streams; worker
// warps turn PCM into phase steps, lane-parallel and coalesced (frequency = deviation * pcm / 32768,
// step = (float)((2*M_PI*frequency)/8000.0), the double division included), into a shared-memory tile;
// the chain warp's lane r walks row r in place, leaving the phase BEFORE each step; worker warps write
// the finished tile out, coalesced.  Three tiles rotate through the three roles, one __syncthreads per
// chunk.  Same operations in the same order per stream as PhaseAccumulator::run.  tx_kernel<FM> reads the
// phases, which lets it cut the call into time tiles like the other modes.
constexpr int FP_ROWS = 32;              // streams per CTA
constexpr int FP_CH = 64;                // PCM samples per chunk
constexpr int FP_PITCH = FP_CH + 1;      // floats per row: lane r, sample n -> bank (r + n) % 32
constexpr int FP_WORKERS = 16;           // worker warps; warp FP_WORKERS is the chain warp
constexpr int FP_RPW = FP_ROWS / FP_WORKERS; // rows per worker warp

struct SmemFp {
    float t[3][FP_ROWS][FP_PITCH];
    uint32_t big[3][FP_ROWS];            // per row and chunk: some |step| >= 3 (see tx_wbfm_kernel's chain)
};

// PROTO: the same machine for the prototype FM head of signals/fm.cc:41-62 (streams of a K_IQ launch in mode
// FM_PROTO; the others' rows stay idle): step = x/65536*3.5 in float, theta += step, ONE wrap at +-2*pi with the
// reference's double expression (|step| < 1.75, so its while loops run at most once), and the value stored is
// theta AFTER the update, which is what fm.cc takes the cosine of.
template <bool PROTO>
__global__ void __launch_bounds__((FP_WORKERS + 1) * 32) tx_fm_phase_kernel(const TxParams p)
{
    __shared__ SmemFp sm;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const bool chain_warp = warp == FP_WORKERS;
    const int row0 = blockIdx.x * FP_ROWS;
    const int rows_live = min(FP_ROWS, p.n_streams - row0);
    const uint32_t n_chunks = (p.n8 + FP_CH - 1) / FP_CH;

    float phase = 0.f;
    int sid_chain = 0;
    const bool live = lane < rows_live; // chain warp: this lane has a stream
    bool live_row = live; // PROTO: only the streams whose head is fm.cc
    if (chain_warp && live) {
        sid_chain = p.stream_ids[row0 + lane];
        if (PROTO) live_row = p.mode_of[sid_chain] == SIG_MODE_FM_PROTO;
        phase = PROTO ? p.state[sid_chain].sig_theta : p.state[sid_chain].fm_phase;
    }

    // workers: PCM of chunk c -> phase steps, rows warp, warp + FP_WORKERS, ...  All loads first (the rows
    // are in different streams: four independent DRAM round trips instead of four in a row), and the
    // double division by the multiply-and-check of div_const_to_float.
    int w_sid[FP_RPW];
    float w_dev[FP_RPW];
#pragma unroll
    for (int i = 0; i < FP_RPW; i++) {
        const int r = warp + i * FP_WORKERS;
        w_sid[i] = (!chain_warp && r < rows_live) ? p.stream_ids[row0 + r] : -1;
        if (PROTO && w_sid[i] >= 0 && p.mode_of[w_sid[i]] != SIG_MODE_FM_PROTO) w_sid[i] = -1;
        w_dev[i] = (!PROTO && w_sid[i] >= 0) ? p.param[w_sid[i]] : 0.f;
    }
    auto produce = [&](uint32_t c) {
        int x[FP_RPW][FP_CH / 32];
#pragma unroll
        for (int i = 0; i < FP_RPW; i++) {
            const int16_t *src = p.pcm + (size_t)max(w_sid[i], 0) * p.pcm_stride;
#pragma unroll
            for (int j = 0; j < FP_CH / 32; j++) {
                const uint32_t at = c * FP_CH + lane + 32 * j;
                x[i][j] = (w_sid[i] >= 0 && at < p.n8) ? (int)src[at] : 0;
            }
        }
#pragma unroll
        for (int i = 0; i < FP_RPW; i++) {
            const int r = warp + i * FP_WORKERS;
            if (w_sid[i] < 0) continue; // warp-uniform
            bool big = false;
#pragma unroll
            for (int j = 0; j < FP_CH / 32; j++) {
                float step;
                if (PROTO) {
                    step = __fmul_rn(__fdiv_rn((float)x[i][j], 65536.f), 3.5f); // fm.cc:43-45
                } else {
                    const float f = __fdiv_rn(__fmul_rn(w_dev[i], (float)x[i][j]), 32768.f);
                    // PhaseAccumulator::setFrequency (:95-107): (float)((2*M_PI*f)/8000.0) in double
                    step = div_const_to_float<8000>(2.0 * 3.14159265358979323846 * (double)f);
                }
                sm.t[c % 3][r][lane + 32 * j] = step;
                big |= !(fabsf(step) < 3.0f);
            }
            big = __any_sync(HRD_FULL_MASK, big);
            if (lane == 0) sm.big[c % 3][r] = big;
        }
    };
    // chain warp: PhaseAccumulator::run over chunk c of row `lane`, in place
    const ChainConsts cc = {c_tabtx.k_sign, c_tabtx.k_m2pi_hi, c_tabtx.k_m2pi_lo};
    auto walk = [&](uint32_t c) {
        const int nb = (int)min((uint32_t)FP_CH, p.n8 - c * FP_CH);
        float *row = sm.t[c % 3][lane];
        if (PROTO) {
            if (!live_row) return;
            for (int n = 0; n < nb; n++) {
                // theta = theta + thetaNew; while (theta > 2*M_PI) theta -= 2*M_PI; (and the mirror image).
                // (double)a > 2*M_PI  <=>  a >= fl32(2*pi), the float just above 2*pi; |a| < 2*pi + 1.75: one wrap.
                const float a = __fadd_rn(phase, row[n]);
                const float w = (float)((double)a - copysign(2.0 * 3.14159265358979323846, (double)a));
                phase = fabsf(a) >= HRD_2PI_HI ? w : a;
                row[n] = phase;
            }
            return;
        }
        const bool slow = __any_sync(HRD_FULL_MASK, live && (sm.big[c % 3][lane] != 0 || !(fabsf(phase) < HRD_PI_UP)));
        if (!live) return;
        if (!slow && nb == FP_CH) {
            // |phase + step| < pi + 3 < 6.2: the single-FMA wrap is the exact one (phase_step_fast)
#pragma unroll 16
            for (int n = 0; n < FP_CH; n++) {
                const float s = row[n];
                row[n] = phase;
                phase = phase_step_fast(phase, s, cc);
            }
        } else {
            for (int n = 0; n < nb; n++) {
                const float s = row[n];
                row[n] = phase;
                phase = phase_advance(phase, s);
            }
        }
    };
    // workers: phases of chunk c -> global, coalesced
    auto flush = [&](uint32_t c) {
#pragma unroll
        for (int i = 0; i < FP_RPW; i++) {
            const int r = warp + i * FP_WORKERS;
            if (w_sid[i] < 0) continue;
            float *out = p.fm_phase + (size_t)(row0 + r) * p.n8;
#pragma unroll
            for (int n = lane; n < FP_CH; n += 32) {
                const uint32_t at = c * FP_CH + n;
                if (at < p.n8) out[at] = sm.t[c % 3][r][n];
            }
        }
    };

    if (!chain_warp) produce(0);
    __syncthreads();
    for (uint32_t t = 0; t < n_chunks; t++) {
        if (chain_warp) {
            walk(t);
        } else {
            if (t >= 1) flush(t - 1);
            if (t + 1 < n_chunks) produce(t + 1);
        }
        __syncthreads();
    }
    if (!chain_warp) flush(n_chunks - 1);
    if (chain_warp && live && live_row) (PROTO ? p.state_out[sid_chain].sig_theta : p.state_out[sid_chain].fm_phase) = phase;
}

// mode NONE: BasebandDataProcessor.cc:689-694 fills the block with 64
__global__ void tx_idle_kernel(const TxParams p)
{
    const int slot = blockIdx.x;
    const int sid = p.stream_ids[slot];
    uint4 *dst = reinterpret_cast<uint4 *>(p.iq + (size_t)sid * p.iq_stride);
    const size_t n16 = (size_t)p.n8 * 32;
    const uint4 v = make_uint4(0x40404040u, 0x40404040u, 0x40404040u, 0x40404040u);
    for (size_t i = (size_t)blockIdx.y * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.y * blockDim.x)
        dst[i] = v;
}

template <int KIND>
int launch_one(const TxParams &p, cudaStream_t s)
{
    const long long items = (long long)p.n_streams * p.n_tiles;
    const int grid = (int)((items + HRD_WARPS_PER_CTA - 1) / HRD_WARPS_PER_CTA);
    const size_t smem = sizeof(SmemTx) * HRD_WARPS_PER_CTA;
    tx_kernel<KIND><<<grid, HRD_WARPS_PER_CTA * 32, smem, s>>>(p);
    return (int)cudaGetLastError();
}

template <int KIND>
int tx_resident_warps()
{
    int blocks = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, tx_kernel<KIND>, HRD_WARPS_PER_CTA * 32,
                                                      sizeof(SmemTx) * HRD_WARPS_PER_CTA) != cudaSuccess || blocks < 1)
        blocks = 1;
    return blocks * HRD_WARPS_PER_CTA;
}

} // namespace

int tx_halo_samples(int kind) { return kind == K_SSB ? (int)TxHaloOf<K_SSB>::value : (int)TxHaloOf<K_AM>::value; }

int tx_resident_warps_per_sm(int kind)
{
    static int cache[K_COUNT] = {};
    if (kind != K_AM && kind != K_FM && kind != K_SSB && kind != K_IQ) return HRD_WARPS_PER_CTA;
    if (!cache[kind])
        cache[kind] = kind == K_AM ? tx_resident_warps<K_AM>() : kind == K_FM ? tx_resident_warps<K_FM>()
                      : kind == K_IQ ? tx_resident_warps<K_IQ>() : tx_resident_warps<K_SSB>();
    return cache[kind];
}

int launch_tx_fm_phase(const TxParams &p, cudaStream_t s)
{
    if (p.n_streams <= 0 || p.n8 == 0) return 0;
    const int grid = (p.n_streams + FP_ROWS - 1) / FP_ROWS;
    tx_fm_phase_kernel<false><<<grid, (FP_WORKERS + 1) * 32, 0, s>>>(p);
    return (int)cudaGetLastError();
}

int launch_tx_sig_phase(const TxParams &p, cudaStream_t s)
{
    if (p.n_streams <= 0 || p.n8 == 0) return 0;
    const int grid = (p.n_streams + FP_ROWS - 1) / FP_ROWS;
    tx_fm_phase_kernel<true><<<grid, (FP_WORKERS + 1) * 32, 0, s>>>(p);
    return (int)cudaGetLastError();
}

int launch_tx(int kind, const TxParams &p, cudaStream_t s)
{
    if (p.n_streams <= 0 || p.n8 == 0) return 0;
    switch (kind) {
    case K_AM: return launch_one<K_AM>(p, s);
    case K_FM: return launch_one<K_FM>(p, s);
    case K_SSB: return launch_one<K_SSB>(p, s);
    case K_IQ: return launch_one<K_IQ>(p, s);
    case K_WBFM: {
        TxParams q = p;
        q.items_per_cta = balanced_items_per_cta(p.n_streams, p.sm_count, TW_ITEMS);
        const int grid = (p.n_streams + q.items_per_cta - 1) / q.items_per_cta;
        static PerDeviceOnce optin;
        const cudaError_t e = optin.run([] { return cudaFuncSetAttribute(tx_wbfm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemTw)); });
        if (e != cudaSuccess) return (int)e;
        tx_wbfm_kernel<<<grid, 1024, sizeof(SmemTw), s>>>(q); // always 32 warp slots (see WARP ROLES)
        return (int)cudaGetLastError();
    }
    case K_NONE: {
        dim3 grid(p.n_streams, 8); // streams on x: no 65535 ceiling
        tx_idle_kernel<<<grid, 256, 0, s>>>(p);
        return (int)cudaGetLastError();
    }
    }
    return (int)cudaErrorInvalidValue;
}

} // namespace hrd
