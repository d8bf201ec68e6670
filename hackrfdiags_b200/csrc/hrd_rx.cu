// hrd_rx.cu -- receive chains: int8 I,Q at 2.048 MS/s (or 256 kS/s) -> int16 PCM at 8 kS/s.
//
// Replaces, per stream (reference paths relative to radioDiags/):
//   src_diags/IqDataProcessor.cc:429-500  reduceSampleRate  (3 x half-band /2, (int8_t) wrap)
//   src_diags/IqDataProcessor.cc:771-815  upconvertByFsOver4
//   AmDemodulator/AmDemodulator.cc:297-504, FmDemodulator/FmDemodulator.cc:353-585,
//   WbFmDemodulator/WbFmDemodulator.cc:341-500, SsbDemodulator/SsbDemodulator.cc:420-598
// and underneath them Filters/Int16/{Decimator,FirFilter}_int16.cc, Filters/{Fir,Iir}Filter.cc.
//
// Work decomposition.  A call is cut, per stream, into TIME TILES of
// tile_batches batches; one batch is 1024 samples at 256 kS/s (= 8192 input samples = 32 PCM
// samples).  One warp owns one (stream, tile):
//   * tile 0 starts from the stream's saved state (filter histories of the previous call);
//   * a later tile starts HALO batches before its first output with all-zero histories.  Every
//     stage up to the detector is an FIR, so after the halo its histories hold exactly what a
//     serial run would hold (the halo is longer than the summed look-back of the cascade); the
//     halo's own outputs are simply not stored;
//   * the last tile writes the stream's state for the next call (into the other half of the
//     double-buffered state array, see RxParams).
// The two float recurrences are the exception: the AM/SSB DC-removal IIR (8 kS/s) has no finite
// look-back, so the tile kernel stops at the IIR's INPUT (one int32 per PCM sample) and
// rx_dc_iir_kernel runs the recurrence afterwards, one thread per stream, serially, exactly as
// the reference does.  The WBFM de-emphasis IIR (256 kS/s) has its own kernel (rx_wbfm_kernel
// below: recurrences transposed onto a chain warp, time tiles by verified speculation).
//
// Inside a batch:
//   A. front end, 16 iterations: every lane takes 32 input bytes (16 I,Q samples, one
//      LDG.E.256, software-prefetched DEPTH iterations ahead) through the three /2 stages with
//      dp2a on packed int8, passes the one boundary sample each stage needs to its neighbour
//      lane by shuffle, applies the Fs/4 rotation and the (int8_t) wrap, and appends one packed
//      word (2 samples) to a shared-memory ring;
//   B. the mode's demodulator runs lane-parallel over the ring at 64 k, 16 k and 8 kS/s;
//   C. 32 PCM samples (or IIR inputs) leave with one coalesced store.
// Integer sections are bit-exact by construction (same Q15 arithmetic, same wrap-around);
// float sections use the reference's operation order with FMA contraction disabled
// (-fmad=false) and IEEE division.
#include <type_traits>

#include "hrd_device.cuh"

#ifndef HRD_EXP
#define HRD_EXP 0 // timing experiments only (tools/exp_build.sh): bits switch parts of a kernel off
#endif
#ifndef HRD_RX_WB_ROLES
#define HRD_RX_WB_ROLES 0 // rx_wbfm_kernel: 1 = chain on warp 31 / scheduler 3 with the fewest items (as tx_wbfm_kernel), 0 = behind the items
#endif
#ifndef HRD_RX_WB_DEP
#define HRD_RX_WB_DEP 1 // rx_wbfm_kernel: transpose_after's scheduling dependency on (1) or off (0)
#endif
#ifndef HRD_RX_WB_LDDEP
#define HRD_RX_WB_LDDEP 1 // rx_wbfm_kernel: the refill load waits for the first transposition (offset_after)
#endif

namespace hrd {

__constant__ ConstTables c_tab;

void upload_tables(const ConstTables &t) { cudaMemcpyToSymbol(c_tab, &t, sizeof t); }

namespace {

constexpr int BATCH256 = 1024; // 256 kS/s samples per batch
constexpr int IT_SAMPLES = 128; // 256 kS/s samples per warp iteration: FOUR per lane (32 input samples, 64 bytes)
// Register prefetch depth, in warp iterations.  ONE, on purpose: ptxas puts every global load of the loop on one
// counting scoreboard (tools/sass_ctl.py), so with two buffers the consumer of the older load also waits for
// the younger one and the second buffer buys nothing but register pressure (measured: depth 1 is 1-4 % faster
// than depth 2 in all three kernels).  The distance comes from the L2 prefetch below instead.
constexpr int RX_DEPTH = 1;
#ifndef HRD_WB_DEPTH
#define HRD_WB_DEPTH 1
#endif
constexpr int WB_DEPTH = HRD_WB_DEPTH;    // the same in rx_wbfm_kernel
#ifndef HRD_L2_AHEAD
#define HRD_L2_AHEAD 4
#endif
constexpr uint32_t RX_L2_AHEAD = HRD_L2_AHEAD; // ... and a 4 KiB chunk pulled into L2 this many iterations ahead (prefetch_chunk)

// batches a tile > 0 runs ahead of its first stored output.  Look-back of each cascade in PCM
// periods (one batch = 32): AM 10, FM 23, SSB 40 (the 31-tap Hilbert FIR at 8 kS/s),
// WBFM 22 after the de-emphasis IIR plus >= 1024 samples at 256 kS/s of IIR warm-up.
template <int KIND> struct HaloOf { static constexpr int value = 1; };
template <> struct HaloOf<K_WBFM> { static constexpr int value = 2; };

// ------------------------------------------------------------------------------------
// per-warp shared memory, by mode
// ------------------------------------------------------------------------------------
struct SmemNone {
    uint32_t dummy[4];
};
struct alignas(16) SmemAm { // AM and SSB streams (one launch runs both)
    uint32_t r256[2 + BATCH256 / 2]; // packed int8 words, 2 samples each
    // int16 samples per rail, so that a word holds two consecutive samples of ONE rail (mac_pair's operand)
    int16_t d64i[8 + BATCH256 / 4], d64q[8 + BATCH256 / 4];
    int16_t a16i[14 + BATCH256 / 16], a16q[14 + BATCH256 / 16];
    uint32_t d8[30 + 32];            // SSB only: I/Q pairs at 8 kS/s
};
struct alignas(16) SmemFm { // 16-byte multiples per warp: the tuner reads its 16 ring words as LDS.64 pairs
    uint32_t r256[14 + BATCH256 / 2];
    float th[4 + BATCH256 / 4];
    int16_t d64[8 + BATCH256 / 4];
    int16_t a16[38 + BATCH256 / 16];
};
template <int KIND> struct SmemOf;
template <> struct SmemOf<K_NONE> { typedef SmemNone type; };
template <> struct SmemOf<K_AM> { typedef SmemAm type; };
template <> struct SmemOf<K_FM> { typedef SmemFm type; };

// ------------------------------------------------------------------------------------
// A. front end
// ------------------------------------------------------------------------------------
struct FeCarry {
    uint32_t t, v, u; // last transposed word of stage 1/2/3 input, as seen by this lane
};

struct FeTaps {
    uint32_t a0, b0, a1, b1, a2, b2;
};

// The transposition {I,Q,I,Q} -> {I,I,Q,Q} of a raw input word, with a SCHEDULING dependency: selector 0x3120 takes
// all four bytes from `raw`, so the second operand changes nothing -- but it makes the instruction wait for `after`,
// a value the previous iteration produces at its END.  Without it ptxas hoists the next iteration's transpositions
// to right behind the load that feeds them (the iterations are unrolled), and since every global load of the loop
// counts on ONE scoreboard the warp then waits a full memory latency per iteration: the one-iteration register
// prefetch was not a prefetch at all (ncu: long_scoreboard on these PRMTs was the top stall of every Rx kernel).
__device__ __forceinline__ uint32_t transpose_after(uint32_t raw, uint32_t after)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, 0x3120;" : "=r"(d) : "r"(raw), "r"(after));
    return d;
}

// The same kind of dependency for the NEXT load's offset: selector 0x3210 returns `off` unchanged, but the load that
// uses the result cannot be issued before `after` (a transposed word of the buffer being consumed) exists.  ptxas
// renames the in-place refill to another register block and, left alone, hoists that LDG above the transpositions
// of the old block -- onto the SAME counting scoreboard, so the first transposition then waits for the load that
// was issued a few instructions earlier: a full memory latency, once per step (ncu: 12 % of rx_wbfm_kernel's warp
// samples sat on that one PRMT).
__device__ __forceinline__ uint32_t offset_after(uint32_t off, uint32_t after)
{
    asm("prmt.b32 %0, %0, %1, 0x3210;" : "+r"(off) : "r"(after));
    return off;
}

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

// The same stage on separate I-pair and Q-pair words {x_e, x_o, -, -}; `left` is the merged {I,I,Q,Q} word of
// the lane before.
template <int N>
__device__ __forceinline__ void halfband_split(const uint32_t (&ini)[N], const uint32_t (&inq)[N], uint32_t left, uint32_t a,
                                               uint32_t b, int (&pi)[N], int (&pq)[N])
{
    pi[0] = dp2a_lo_us(a, ini[0], dp2a_lo_us(b, left, 32768));
    pq[0] = dp2a_lo_us(a, inq[0], dp2a_hi_us(b, left, 32768));
#pragma unroll
    for (int r = 1; r < N; r++) {
        pi[r] = dp2a_lo_us(a, ini[r], dp2a_lo_us(b, ini[r - 1], 32768));
        pq[r] = dp2a_lo_us(a, inq[r], dp2a_lo_us(b, inq[r - 1], 32768));
    }
}

// 32 input samples (16 raw words {I,Q,I,Q}) -> two ring words {I0,I1,Q0,Q1}, {I2,I3,Q2,Q3} at 256 kS/s,
// rotated by +Fs/4 (IqDataProcessor.cc:771-815) and narrowed like (int8_t) does.
// t[] is the raw input already transposed to {I_e, I_o, Q_e, Q_o} (the caller does that first
// so that the raw registers are free to receive the next prefetch in place).
// A lane takes FOUR output samples, not two: one neighbour shuffle per stage serves twice the samples, and
// the lane's samples always sit at rotation phases 0,1,2,3, so the rotation needs no per-lane selects.
__device__ __forceinline__ uint2 front_end_iter(const uint32_t (&t)[16], FeCarry &c, const FeTaps &k, int lane)
{
    int pi16[16], pq16[16];
    halfband_stage<16>(t, from_left(t[15], c.t, lane), k.a0, k.b0, pi16, pq16);
    // Stages 2 and 3 keep the I pairs and the Q pairs in separate words (pack_b2 puts a pair into bytes 0,1,
    // which is all dp2a.lo reads): merging them into {I,I,Q,Q} words first would cost a third PRMT per word.
    // Only the word that crosses to the next lane is merged, so the carried state keeps its layout.
    uint32_t vi[8], vq[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        vi[j] = pack_b2(pi16[2 * j], pi16[2 * j + 1]);
        vq[j] = pack_b2(pq16[2 * j], pq16[2 * j + 1]);
    }
    int pi8[8], pq8[8];
    halfband_split<8>(vi, vq, from_left(merge16(vi[7], vq[7]), c.v, lane), k.a1, k.b1, pi8, pq8);
    uint32_t ui[4], uq[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        ui[j] = pack_b2(pi8[2 * j], pi8[2 * j + 1]);
        uq[j] = pack_b2(pq8[2 * j], pq8[2 * j + 1]);
    }
    int pi4[4], pq4[4];
    halfband_split<4>(ui, uq, from_left(merge16(ui[3], uq[3]), c.u, lane), k.a2, k.b2, pi4, pq4);
    // rotation, samples 0..3 of the lane = phases 0..3 of the call (a call starts at a multiple of four
    // samples):  0:(x,y) 1:(-y,x) 2:(-x,-y) 3:(y,-x).  Negating "acc>>16" is (65535-acc)>>16.
    const int i0 = pi4[0], q0 = pq4[0];
    const int i1 = 65535 - pq4[1], q1 = pi4[1];
    const int i2 = 65535 - pi4[2], q2 = 65535 - pq4[2];
    const int i3 = pq4[3], q3 = 65535 - pi4[3];
    return make_uint2(merge16(pack_b2(i0, i1), pack_b2(q0, q1)), merge16(pack_b2(i2, i3), pack_b2(q2, q3)));
}

// The narrow form, kept for rx_wbfm_kernel (whose item warps have no registers to spare: the wide form spills there
// and was 14 % slower): 16 input samples (8 raw words {I,Q,I,Q}) -> one ring word {I0,I1,Q0,Q1} at 256 kS/s,
// rotated by +Fs/4 (IqDataProcessor.cc:771-815) and narrowed like (int8_t) does.
// t[] is the raw input already transposed to {I_e, I_o, Q_e, Q_o} (the caller does that first
// so that the raw registers are free to receive the next prefetch in place).
__device__ __forceinline__ uint32_t front_end_iter_narrow(const uint32_t (&t)[8], FeCarry &c, const FeTaps &k, int lane)
{
    int pi8[8], pq8[8];
    halfband_stage<8>(t, from_left(t[7], c.t, lane), k.a0, k.b0, pi8, pq8);
    // Stages 2 and 3 keep the I pairs and the Q pairs in separate words (pack_b2 puts a pair into bytes 0,1,
    // which is all dp2a.lo reads): merging them into {I,I,Q,Q} words first would cost a third PRMT per word.
    // Only the word that crosses to the next lane is merged, so the carried state keeps its layout.
    uint32_t vi[4], vq[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        vi[j] = pack_b2(pi8[2 * j], pi8[2 * j + 1]);
        vq[j] = pack_b2(pq8[2 * j], pq8[2 * j + 1]);
    }
    int pi4[4], pq4[4];
    halfband_split<4>(vi, vq, from_left(merge16(vi[3], vq[3]), c.v, lane), k.a1, k.b1, pi4, pq4);
    uint32_t ui[2], uq[2];
#pragma unroll
    for (int j = 0; j < 2; j++) {
        ui[j] = pack_b2(pi4[2 * j], pi4[2 * j + 1]);
        uq[j] = pack_b2(pq4[2 * j], pq4[2 * j + 1]);
    }
    int pi2[2], pq2[2];
    halfband_split<2>(ui, uq, from_left(merge16(ui[1], uq[1]), c.u, lane), k.a2, k.b2, pi2, pq2);
    // rotation: this lane's samples are number 2*lane and 2*lane+1 of the iteration, so
    // their phases are {0,1} on even lanes and {2,3} on odd lanes:
    //   0:(x,y) 1:(-y,x) 2:(-x,-y) 3:(y,-x).  Negating "acc>>16" is (65535-acc)>>16.
    const int s = (lane & 1) ? -1 : 1;
    const int kk = (lane & 1) ? 65535 : 0;
    const int kn = (lane & 1) ? 0 : 65535;
    int i0 = pi2[0] * s + kk;
    int q0 = pq2[0] * s + kk;
    int i1 = pq2[1] * (-s) + kn;
    int q1 = pi2[1] * s + kk;
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

// I/Q int16 pairs, N taps, decimate by M: output k reads ring[M*k + N - 1 - t]
// (the ring keeps hist == N-M old samples in front of the new ones).
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

// Two int16 samples {x[2w], x[2w+1]} against their two taps in ONE pair of dp2a: the taps are split
// tap = th*256 + tl (th signed, tl unsigned byte) and packed {tl0, tl1, th0, th1} on the host
// (hrd_api.cu split_taps), so  acc_l += x0*tl0 + x1*tl1,  acc_h += x0*th0 + x1*th1  and the Q15
// accumulator is acc_l + 256*acc_h (mod 2^32, like the reference's int32).  Half the instructions of
// the unpack-and-IMAD form, and the samples stay packed from narrowing to the last decimator.
__device__ __forceinline__ void mac_pair(uint32_t x, uint32_t taps, int &acc_l, int &acc_h)
{
    asm("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(acc_l) : "r"(x), "r"(taps));
    asm("dp2a.hi.s32.s32 %0, %1, %2, %0;" : "+r"(acc_h) : "r"(x), "r"(taps));
}

template <typename T>
__device__ __forceinline__ void ring_init(T *ring, const T *state, int hist, int lane, bool from_state)
{
    for (int i = lane; i < hist; i += 32) ring[i] = from_state ? state[i] : T();
}

// input of one warp iteration: 64 bytes per lane at the 2.048 MS/s entry, 8 at the 256 kS/s one
struct u32x16 {
    u32x8 a, b;
};
template <int ENTRY> struct RawOf { typedef u32x16 type; };
template <> struct RawOf<1> { typedef uint2 type; };

// Unconditional load at a clamped offset: lanes past the end of the tile re-read its last
// chunk instead of being predicated off (a predicated LDG.256 costs 16 register moves to keep
// the old value alive).  What they compute is never stored and never reaches a live lane:
// data only flows from lower to higher lanes, and a partly filled iteration is the tile's last.
template <int ENTRY>
__device__ __forceinline__ typename RawOf<ENTRY>::type load_raw(const int8_t *src, uint32_t off, uint32_t off_last)
{
    const uint32_t o = min(off, off_last);
    if constexpr (ENTRY == 0) {
        u32x16 r;
        r.a = ldg_stream_256(src + o);
        r.b = ldg_stream_256(src + o + 32);
        return r;
    } else { // rows are only 4-byte aligned at this entry
        const uint32_t *q = reinterpret_cast<const uint32_t *>(src + o);
        return make_uint2(__ldg(q), __ldg(q + 1));
    }
}

// An L2 prefetch needs no register and no scoreboard: once every two iterations the warp pulls the 4 KiB it
// will read RX_L2_AHEAD KiB from now out of DRAM (lane l takes the 128-byte line l), and the LDG that
// follows later hits in L2 (a few hundred cycles instead of a DRAM round trip under load).
// pf is the lane's own next load offset (warp base + 64 * lane).
// (ncu shows only ~55 % of the LDG sectors hitting in L2.  One prefetch per 64 bytes instead of per 128-byte line --
// in case the L2 fetched 64 bytes per miss -- was measured and lost: AM 1.412 -> 1.465 ms, FM 1.588 -> 1.654 ms.)
__device__ __forceinline__ void prefetch_chunk(const int8_t *src, uint32_t pf, uint32_t pf_last, int lane)
{
    prefetch_l2(src + min(pf + 64u * (uint32_t)lane + RX_L2_AHEAD * 1024u, pf_last));
}

// ---- the same for the narrow form (two samples per lane, IT_NARROW samples per warp iteration) ----
constexpr int IT_NARROW = 64;
// input of one warp iteration: 32 bytes per lane at the 2.048 MS/s entry, 4 at the 256 kS/s one
template <int ENTRY> struct RawNarrowOf { typedef u32x8 type; };
template <> struct RawNarrowOf<1> { typedef uint32_t type; };

// Unconditional load at a clamped offset: lanes past the end of the tile re-read its last
// chunk instead of being predicated off (a predicated LDG.256 costs 16 register moves to keep
// the old value alive).  What they compute is never stored and never reaches a live lane:
// data only flows from lower to higher lanes, and a partly filled iteration is the tile's last.
template <int ENTRY>
__device__ __forceinline__ typename RawNarrowOf<ENTRY>::type load_raw_narrow(const int8_t *src, uint32_t off, uint32_t off_last)
{
    const uint32_t o = min(off, off_last);
    if constexpr (ENTRY == 0) {
        return ldg_stream_256(src + o);
    }
    else
        return __ldg(reinterpret_cast<const uint32_t *>(src + o));
}

// The same for the narrow form, once per step of four iterations (4 KiB): lane l takes the 128-byte line l of the
// step that starts RX_L2_AHEAD iterations from now.  pf is the lane's own next load offset (warp base + 32 * lane).
__device__ __forceinline__ void prefetch_chunk_narrow(const int8_t *src, uint32_t pf, uint32_t pf_last, int lane)
{
    prefetch_l2(src + min(pf + 96u * (uint32_t)lane + RX_L2_AHEAD * IT_NARROW * 16, pf_last));
}

// ------------------------------------------------------------------------------------
// the tile kernel
// ------------------------------------------------------------------------------------
#ifndef HRD_RX_MIN_CTAS
#define HRD_RX_MIN_CTAS 8
#endif
template <int KIND, int ENTRY>
__global__ void __launch_bounds__(HRD_WARPS_PER_CTA * 32, HRD_RX_MIN_CTAS) rx_kernel(const RxParams p)
{
    typedef typename SmemOf<KIND>::type Smem;
    typedef typename RawOf<ENTRY>::type Raw;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int item = blockIdx.x * HRD_WARPS_PER_CTA + warp;
    if (item >= p.n_streams * p.n_tiles) return;
    const int tile = item / p.n_streams;
    const int slot = item - tile * p.n_streams;
    const int sid = p.stream_ids[slot];
    const bool first = tile == 0;
    // the K_AM instance runs the AM and the SSB streams of a batch (same /32 decimator chain)
    const bool ssb = KIND == K_AM && p.kind_of[sid] == K_SSB;
    const uint32_t halo = (KIND == K_AM && ssb) ? 2u : (uint32_t)HaloOf<KIND>::value;
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw + (size_t)warp * sizeof(Smem));
    const RxState &st = p.state_in[sid];
    const int8_t *src = p.iq + (size_t)sid * p.iq_stride;
    asm volatile("" : "+l"(src)); // keep the row pointer in registers (else it is recomputed per load)

    // ---- this tile's range, in 256 kS/s samples --------------------------------------
    // (ragged calls -- the squelched path, every stream demodulates only its open blocks -- are tiled over the
    //  NOMINAL length: a stream's tiles past its own end have nothing to do, the tile that holds its end is its last)
    const uint32_t n256 = p.n256_of ? p.n256_of[sid] : p.n256;
    const uint32_t tile_len = p.tile_batches * BATCH256;
    const uint32_t emit_from = (uint32_t)tile * tile_len; // outputs before this are the halo's
    if (tile > 0 && emit_from >= n256) return;
    const uint32_t end256 = min(n256, emit_from + tile_len);
    const bool last = end256 == n256; // leaves the stream's state for the next call
    uint32_t done256 = first ? 0u : emit_from - halo * BATCH256;

    // ---- start state: saved (tile 0) or all-zero (later tiles, rebuilt by the halo) -----
    FeCarry fc;
    fc.t = first ? st.fe_t : 0u;
    fc.v = first ? st.fe_v : 0u;
    fc.u = first ? st.fe_u : 0u;
    FeTaps fk;
    fk.a0 = c_tab.fe_a[0]; fk.b0 = c_tab.fe_b[0];
    fk.a1 = c_tab.fe_a[1]; fk.b1 = c_tab.fe_b[1];
    fk.a2 = c_tab.fe_a[2]; fk.b2 = c_tab.fe_b[2];
    float gain = 0.f, scale = 0.f;
    float x1 = 0.f;
    bool lsb = true;
    if constexpr (KIND == K_FM) gain = p.gain[sid];
    if constexpr (KIND == K_AM) {
        const RxDec32 &d = ssb ? st.ssb : st.am;
        ring_init(sm.r256, d.r256, 2, lane, first);
        if (lane < 8) { // the state record keeps {I, Q} words per sample; the rings keep the rails apart
            const uint32_t w = first ? d.d64[lane] : 0u;
            sm.d64i[lane] = (int16_t)lo16(w);
            sm.d64q[lane] = (int16_t)hi16(w);
        }
        if (lane < 14) {
            const uint32_t w = first ? d.a16[lane] : 0u;
            sm.a16i[lane] = (int16_t)lo16(w);
            sm.a16q[lane] = (int16_t)hi16(w);
        }
        if (ssb) {
            ring_init(sm.d8, st.ssb_d8, 30, lane, first);
            lsb = p.lsb[sid] != 0;
        }
        x1 = first ? d.x1 : 0.f; // (float)x[n-1] of the DC-removal filter; a halo rebuilds it
    }
    if constexpr (KIND == K_FM) {
        ring_init(sm.r256, st.fm_r256, 14, lane, first);
        ring_init(sm.th, st.fm_theta, 4, lane, first);
        ring_init(sm.d64, st.fm_d64, 8, lane, first);
        ring_init(sm.a16, st.fm_a16, 38, lane, first);
        // FmDemodulator.cc:488-491
        scale = __fmul_rn(__fdiv_rn(gain, 15000.f), 32767.f);
    }
    __syncwarp();

    // ---- software prefetch: RX_DEPTH iterations of input in flight ------------------------
    constexpr uint32_t BPS = ENTRY == 0 ? 16 : 2;           // input bytes per 256 kS/s sample
    uint32_t pf = (done256 + 4 * lane) * BPS;               // this lane's next prefetch offset
    const uint32_t pf_last = end256 >= 4 ? (end256 - 4) * BPS : 0u; // its clamp: the tile's last lane chunk (an empty stream reads its row's first bytes)
    Raw buf[RX_DEPTH];
#pragma unroll
    for (int d = 0; d < RX_DEPTH; d++) {
        buf[d] = load_raw<ENTRY>(src, pf, pf_last);
        pf += IT_SAMPLES * BPS;
    }

    uint32_t last_active = 32;
    uint32_t sched_dep = 0; // see transpose_after

    while (done256 < end256) {
        const uint32_t nb = min((uint32_t)BATCH256, end256 - done256); // multiple of 32
        const uint32_t n_it = (nb + IT_SAMPLES - 1) / IT_SAMPLES;
        const bool emit = done256 >= emit_from;
        last_active = min(32u, (nb - (n_it - 1) * IT_SAMPLES) / 4);

        // (256 kS/s entry: a lane's loads are 8 bytes, a warp's 256 -- far too little in flight to hide DRAM; the
        //  batch after next is pulled into L2 here, 2 KiB = sixteen lines, so that the loads find it there)
        if constexpr (ENTRY == 1) {
            if (lane < 16) prefetch_l2(src + min((done256 + 2 * BATCH256) * BPS + 128u * (uint32_t)lane, pf_last));
        }
        // ---- A. front end (or plain load at the 256 kS/s entry) ----------------------
        // one iteration: consume b (loaded RX_DEPTH iterations ago), refill it in place
        auto step = [&](Raw &b, const uint32_t it) {
            uint2 words;
            if constexpr (ENTRY == 0) {
                uint32_t t[16];
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    t[r] = transpose_after(b.a.v[r], sched_dep);
                    t[8 + r] = transpose_after(b.b.v[r], sched_dep);
                }
                b = load_raw<ENTRY>(src, pf, pf_last); // in place: b is dead now
                if ((it & 1) == 0) prefetch_chunk(src, pf, pf_last, lane);
                words = front_end_iter(t, fc, fk, lane);
                sched_dep = words.y;
            } else {
                words = make_uint2(__byte_perm(b.x, 0, 0x3120), __byte_perm(b.y, 0, 0x3120));
                b = load_raw<ENTRY>(src, pf, pf_last);
            }
            pf += IT_SAMPLES * BPS;
            const uint32_t widx = it * (IT_SAMPLES / 2) + 2 * lane; // word index inside the batch (even)
            // lanes past the end of the tile write ring slots nothing reads
            if constexpr (KIND == K_NONE) {
                const uint32_t s0 = done256 + it * IT_SAMPLES + 4 * lane;
                if (p.out256 && s0 < end256 && emit) {
                    uint32_t *o = reinterpret_cast<uint32_t *>(p.out256 + (size_t)sid * p.out_stride + (size_t)s0 * 2);
                    o[0] = __byte_perm(words.x, 0, 0x3120);
                    o[1] = __byte_perm(words.y, 0, 0x3120);
                }
            } else if constexpr (KIND == K_AM) {
                *reinterpret_cast<uint2 *>(&sm.r256[2 + widx]) = words;
            } else {
                *reinterpret_cast<uint2 *>(&sm.r256[14 + widx]) = words;
            }
        };
#pragma unroll 2
        for (uint32_t it = 0; it < n_it; it++) step(buf[0], it);
        __syncwarp();

        const int n64 = nb / 4, n16 = nb / 16, n8 = nb / 32;
        const size_t pcm_at = done256 / 32;
        int16_t *pcm_out = p.pcm + (size_t)sid * p.pcm_stride + pcm_at;

        // ---- B. demodulators ----------------------------------------------------------
        if constexpr (KIND == K_AM) {
            // AmDemodulator.cc:339-408 / SsbDemodulator.cc:462-529: /4 (8) /4 (12) /2 (16)
            for (int j = lane; j < n64; j += 32) {
                int yi, yq;
                dec4_int8<8>(sm.r256, c_tab.am1, j, yi, yq);
                sm.d64i[8 + j] = (int16_t)yi;
                sm.d64q[8 + j] = (int16_t)yq;
            }
            __syncwarp();
            // /4 (12 taps): output k reads ring samples 4k .. 4k+11 = words 2k .. 2k+5 of each rail
            for (int k = lane; k < n16; k += 32) {
                const uint32_t *ri = reinterpret_cast<const uint32_t *>(sm.d64i) + 2 * k;
                const uint32_t *rq = reinterpret_cast<const uint32_t *>(sm.d64q) + 2 * k;
                int il = 1 << 14, ih = 0, ql = 1 << 14, qh = 0;
#pragma unroll
                for (int w = 0; w < 6; w++) {
                    mac_pair(ri[w], c_tab.am2_sp[w], il, ih);
                    mac_pair(rq[w], c_tab.am2_sp[w], ql, qh);
                }
                sm.a16i[14 + k] = (int16_t)q15(il + (ih << 8));
                sm.a16q[14 + k] = (int16_t)q15(ql + (qh << 8));
            }
            __syncwarp();
            // /2 (16 taps): output k reads ring samples 2k .. 2k+15 = words k .. k+7
            uint32_t iq8 = 0;
            if (lane < n8) {
                const uint32_t *ri = reinterpret_cast<const uint32_t *>(sm.a16i) + lane;
                const uint32_t *rq = reinterpret_cast<const uint32_t *>(sm.a16q) + lane;
                int il = 1 << 14, ih = 0, ql = 1 << 14, qh = 0;
#pragma unroll
                for (int w = 0; w < 8; w++) {
                    mac_pair(ri[w], c_tab.am3_sp[w], il, ih);
                    mac_pair(rq[w], c_tab.am3_sp[w], ql, qh);
                }
                iq8 = pack16(q15(il + (ih << 8)), q15(ql + (qh << 8)));
            }
            int x = 0; // what enters the DC-removal IIR (as float) in the reference
            if (!ssb) {
                // AmDemodulator.cc:444-458: |I|,|Q| narrowed to int16, max + min/2
                int im = (int)(short)abs(lo16(iq8)), qm = (int)(short)abs(hi16(iq8));
                x = (im > qm) ? (int)(short)(im + (qm >> 1)) : (int)(short)(qm + (im >> 1));
            } else {
                if (lane < n8) sm.d8[30 + lane] = iq8;
                __syncwarp();
                if (lane < n8) {
                    // SsbDemodulator.cc:576-590: delay line (= -I[n-15]) and 31-tap Hilbert on Q
                    const uint32_t *r = sm.d8 + lane; // r[30] is sample n
                    int id = q15((1 << 14) + c_tab.delay[15] * lo16(r[30 - 15]));
                    unsigned acc = 1u << 14;
#pragma unroll
                    for (int t = 0; t < 31; t += 2) acc += (unsigned)(c_tab.hilbert[t] * hi16(r[30 - t]));
                    int qh = q15((int)acc);
                    x = lsb ? (id - qh) : (id + qh);
                }
            }
            // the FIR half of the DC-removal filter (IirFilter.cc:161-164, b = {1,-1}):
            // 0 + 1*x[n], then + (-1)*x[n-1], all floats
            const float xf = (float)x;
            float xp = __shfl_up_sync(HRD_FULL_MASK, xf, 1);
            if (lane == 0) xp = x1;
            x1 = __shfl_sync(HRD_FULL_MASK, xf, n8 - 1);
            if (emit && lane < n8) p.pre_iir[(size_t)sid * p.pre_stride + pcm_at + lane] = __fsub_rn(xf, xp);
            __syncwarp();
            ring_shift(sm.r256, 2, nb / 2, lane);
            ring_shift(sm.d64i, 8, n64, lane);
            ring_shift(sm.d64q, 8, n64, lane);
            ring_shift(sm.a16i, 14, n16, lane);
            ring_shift(sm.a16q, 14, n16, lane);
            if (ssb) ring_shift(sm.d8, 30, n8, lane);
        }

        if constexpr (KIND == K_FM) {
            // FmDemodulator.cc:395-442: tuner /4, 32 taps per rail, then the atan2 table
#pragma unroll 4
            for (int j = lane; j < n64; j += 32) {
                int yi, yq;
                dec4_int8<32>(sm.r256, c_tab.fm_tuner, j, yi, yq);
                uint32_t ii = ((uint32_t)yi + 128u) & 255u, qi = ((uint32_t)yq + 128u) & 255u;
                sm.th[4 + j] = __ldg(p.atan2_lut + qi * 256u + ii);
            }
            __syncwarp();
            // FmDemodulator.cc:479-529: FirFilter with taps {0,0,1,0,-1,0,0} = th[n-2]-th[n-4]
            for (int j = lane; j < n64; j += 32) {
                float d = wrap_pi_select(__fsub_rn(sm.th[4 + j - 2], sm.th[4 + j - 4]));
                sm.d64[8 + j] = (int16_t)f32_to_i16(__fmul_rn(scale, d));
            }
            __syncwarp();
            // FmDemodulator.cc:551-585: /4 (12 taps) then /2 (40 taps), two int16 samples and their two taps per
            // mac_pair (output k of the first reads ring samples 4k .. 4k+11, of the second 2k .. 2k+39: whole words)
            for (int k = lane; k < n16; k += 32) {
                const uint32_t *r = reinterpret_cast<const uint32_t *>(sm.d64) + 2 * k;
                int al = 1 << 14, ah = 0;
#pragma unroll
                for (int w = 0; w < 6; w++) mac_pair(r[w], c_tab.fm_post_sp[w], al, ah);
                sm.a16[38 + k] = (int16_t)q15(al + (ah << 8));
            }
            __syncwarp();
            if (lane < n8) {
                const uint32_t *r = reinterpret_cast<const uint32_t *>(sm.a16) + lane;
                int al = 1 << 14, ah = 0;
#pragma unroll
                for (int w = 0; w < 20; w++) mac_pair(r[w], c_tab.audio40_sp[w], al, ah);
                if (emit) pcm_out[lane] = (int16_t)q15(al + (ah << 8));
            }
            __syncwarp();
            ring_shift(sm.r256, 14, nb / 2, lane);
            ring_shift(sm.th, 4, n64, lane);
            ring_shift(sm.d64, 8, n64, lane);
            ring_shift(sm.a16, 38, n16, lane);
        }

        done256 += nb;
    }


    // ---- the last tile leaves the stream's state for the next call -----------------------
    if (!last) return;
    RxState &so = p.state_out[sid];
    {
        // everything this launch does not own (other demodulators, the IIR pair of AM/SSB)
        // is carried over unchanged
        const uint32_t *a = reinterpret_cast<const uint32_t *>(&st);
        uint32_t *b = reinterpret_cast<uint32_t *>(&so);
        for (int i = lane; i < (int)(sizeof(RxState) / 4); i += 32) b[i] = a[i];
    }
    __syncwarp();
    if constexpr (ENTRY == 0) {
        if (lane == (int)last_active - 1) {
            so.fe_t = fc.t;
            so.fe_v = fc.v;
            so.fe_u = fc.u;
        }
    }
    if constexpr (KIND == K_AM) {
        RxDec32 &d = ssb ? so.ssb : so.am;
        ring_save_hist(sm.r256, d.r256, 2, lane);
        if (lane < 8) d.d64[lane] = pack16(sm.d64i[lane], sm.d64q[lane]);
        if (lane < 14) d.a16[lane] = pack16(sm.a16i[lane], sm.a16q[lane]);
        if (ssb) ring_save_hist(sm.d8, so.ssb_d8, 30, lane);
        if (lane == 0) d.x1 = x1; // y1 follows from rx_dc_iir_kernel
    }
    if constexpr (KIND == K_FM) {
        ring_save_hist(sm.r256, so.fm_r256, 14, lane);
        ring_save_hist(sm.th, so.fm_theta, 4, lane);
        ring_save_hist(sm.d64, so.fm_d64, 8, lane);
        ring_save_hist(sm.a16, so.fm_a16, 38, lane);
    }
}

// ------------------------------------------------------------------------------------
// The squelch gate, fused into the front end
// ------------------------------------------------------------------------------------
// Replaces, per stream and per block (one reference call of IqDataProcessor::acceptIqData, paths relative to
// radioDiags/src_diags/):
//   SignalDetector::detectSignal  SignalDetector.cc:205-273   mean of max(|I|,|Q|) + min(|I|,|Q|)/2 over the
//                                                             block's 256 kS/s samples, integer dBFS, threshold
//   DbfsCalculator                DbfsCalculator.cc:36-68,111-147  the dB table (ConstTables::db_table)
//   SignalTracker::run            SignalTracker.cc:104-145    two states; a block after a signal still passes
//   Squelch::run                  Squelch.cc:227-273          decision = START | PRESENT | END (the tail)
// and the gate of IqDataProcessor.cc:991: the demodulator is simply not called for a closed block, so its state
// does not move and no PCM comes out -- for the demodulator the closed blocks never existed.  So: one warp per
// stream runs the front end over the whole call (the front end itself is never gated, :937-946), accumulates the
// block's magnitudes in its epilogue (a few operations per lane and iteration, one warp reduction per BLOCK),
// walks the tracker at every block end, and writes the 256 kS/s samples SPECULATIVELY at the position the block
// takes if the gate lets it through; a closed block just does not advance that position and the next one
// overwrites it.  What is left in the stream's scratch row is the concatenation of its open blocks, which the
// demodulator kernels then take as one ragged call at the 256 kS/s entry (RxParams::n256_of), PCM landing packed
// in the caller's rows.  No host round trip, no per-block launches.  All integer: bit-exact.
__device__ __forceinline__ uint32_t magnitude_pair(uint32_t w) // w = {I0, I1, Q0, Q1} int8: the two samples' levels, summed
{
    const uint32_t a = __vabs4(w); // abs as uint8 (abs(-128) = 128, SignalDetector.cc:229-242)
    const uint32_t i2 = a & 0xffffu, q2 = a >> 16;
    const uint32_t m = __vmaxu4(i2, q2) + ((__vminu4(i2, q2) >> 1) & 0x7f7fu); // per byte <= 192: no carry
    return (m & 0xffu) + (m >> 8);
}

__global__ void __launch_bounds__(HRD_WARPS_PER_CTA * 32, HRD_RX_MIN_CTAS) rx_gate_kernel(const RxParams p, const GateParams g)
{
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * HRD_WARPS_PER_CTA + (threadIdx.x >> 5);
    if (slot >= p.n_streams) return;
    const int sid = p.stream_ids[slot];
    const RxState &st = p.state_in[sid];
    RxState &so = p.state_out[sid];
    { // the whole record moves to the other half of the double buffer; the front-end words are rewritten below
        const uint32_t *a = reinterpret_cast<const uint32_t *>(&st);
        uint32_t *b = reinterpret_cast<uint32_t *>(&so);
        for (int i = lane; i < (int)(sizeof(RxState) / 4); i += 32) b[i] = a[i];
    }
    __syncwarp();
    const int8_t *src = p.iq + (size_t)sid * p.iq_stride;
    asm volatile("" : "+l"(src));
    FeCarry fc;
    fc.t = st.fe_t, fc.v = st.fe_v, fc.u = st.fe_u;
    FeTaps fk;
    fk.a0 = c_tab.fe_a[0]; fk.b0 = c_tab.fe_b[0];
    fk.a1 = c_tab.fe_a[1]; fk.b1 = c_tab.fe_b[1];
    fk.a2 = c_tab.fe_a[2]; fk.b2 = c_tab.fe_b[2];
    const uint32_t n256 = p.n256;
    const int32_t thr = (int32_t)g.threshold[sid];
    const uint32_t gain = (uint32_t)g.gain_db[sid];
    bool track = g.tracking[sid] != 0;
    const bool store = p.kind_of[sid] != K_NONE; // no demodulator: magnitudes and decisions only
    uint32_t *mag_row = g.magnitude + (size_t)sid * g.n_blocks;
    uint8_t *open_row = g.allowed + (size_t)sid * g.n_blocks;
    int8_t *out = g.scratch + (size_t)sid * g.row256;

    uint32_t wp = 0;                                  // samples of open blocks written so far
    uint32_t blk_begin = 0, blk_end = min(g.blk256, n256), k = 0;
    uint32_t acc = 0;                                 // this lane's share of the running block's sum

    uint32_t pf = (4u * (uint32_t)lane) * 16u;
    const uint32_t pf_last = (n256 - 4) * 16u;
    u32x16 buf = load_raw<0>(src, pf, pf_last);
    pf += IT_SAMPLES * 16u;
    const uint32_t n_it = (n256 + IT_SAMPLES - 1) / IT_SAMPLES;
    uint32_t sched_dep = 0;
    for (uint32_t it = 0; it < n_it; it++) {
        uint32_t t[16];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            t[r] = transpose_after(buf.a.v[r], sched_dep);
            t[8 + r] = transpose_after(buf.b.v[r], sched_dep);
        }
        buf = load_raw<0>(src, pf, pf_last);
        if ((it & 1) == 0) prefetch_chunk(src, pf, pf_last, lane);
        pf += IT_SAMPLES * 16u;
        const uint2 words = front_end_iter(t, fc, fk, lane);
        sched_dep = words.y;

        const uint32_t s0 = it * IT_SAMPLES + 4u * (uint32_t)lane; // this lane's four samples: never across a block end
        const bool valid = s0 < n256;
        const uint32_t m = valid ? magnitude_pair(words.x) + magnitude_pair(words.y) : 0u;
        bool placed = false, keep = true; // keep: not known to be closed (a running block is written speculatively)
        uint32_t pos = 0;
        const uint32_t it_end = min(n256, (it + 1) * IT_SAMPLES);
        // every block that ends inside (or with) this iteration, in order: rarely any, a few with tiny blocks
        while (blk_begin < n256 && it_end >= blk_end) { // warp-uniform
            const bool mine = s0 >= blk_begin && s0 < blk_end;
            uint32_t total = acc + (mine ? m : 0u);
#pragma unroll
            for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(HRD_FULL_MASK, total, o);
            if (mine) { // where the block goes IF the gate lets it through
                pos = wp + (s0 - blk_begin);
                placed = true;
            }
            const uint32_t count = blk_end - blk_begin;
            const uint32_t magnitude = total / count; // magnitude /= magnitudeBufferLength (SignalDetector.cc:248)
            // DbfsCalculator::convertMagnitudeToDbFs, 7-bit words: clip to 127, table, minus 42; then "-= gain" on
            // an int32 with a uint32 operand, as SignalDetector.cc:263 writes it
            int32_t dbfs = c_tab.db_table[min(magnitude, 127u)] - 42;
            dbfs = (int32_t)((uint32_t)dbfs - gain);
            const bool present = dbfs >= thr;
            // NoSignal: present -> START (allowed), else NOISE (closed); Tracking: PRESENT, or END = the tail (allowed)
            const bool allowed = track || present;
            track = present;
            if (lane == 0) {
                mag_row[k] = magnitude;
                open_row[k] = allowed ? 1 : 0;
            }
            // the block's lanes of THIS iteration know the decision: a closed block's must not store, or they could
            // land on what an open block that ends in the same iteration has just written (tiny blocks)
            if (mine) keep = allowed;
            if (allowed) wp += count;
            blk_begin = blk_end;
            blk_end = min(blk_begin + g.blk256, n256);
            k++;
            acc = 0;
        }
        if (!placed) { // the lane's samples belong to the block still running
            acc += m;
            pos = wp + (s0 - blk_begin);
        }
        if (store && valid && keep) {
            uint2 o;
            o.x = __byte_perm(words.x, 0, 0x3120); // {I0, Q0, I1, Q1}
            o.y = __byte_perm(words.y, 0, 0x3120);
            *reinterpret_cast<uint2 *>(out + (size_t)pos * 2) = o;
        }
    }
    const uint32_t last_active = min(32u, (n256 - (n_it - 1) * IT_SAMPLES) / 4);
    if (lane == (int)last_active - 1) {
        so.fe_t = fc.t;
        so.fe_v = fc.v;
        so.fe_u = fc.u;
    }
    if (lane == 0) {
        g.n256_open[sid] = store ? wp : 0u;
        g.tracking[sid] = track ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------
// WBFM (WbFmDemodulator.cc:341-500)
// ------------------------------------------------------------------------------------
// The de-emphasis filter (WbFmDemodulator.cc:93-102, IirFilter.cc:161-176) is a float recurrence
// at 256 kS/s, y[n] = fir[n] - (-0.9492274f * y[n-1]): a dependent FMUL+FSUB per sample that no
// amount of lanes can shorten.  Run on one lane of the warp that owns the stream it idles the
// other 31 (the first version of this kernel: 18 % of the HBM roofline, issue-bound on one-lane
// instructions).  Here the recurrences of 31 items are TRANSPOSED onto the lanes of one warp:
//
//   CTA = up to 32 warps.  All but the last own one (stream, tile) item each and do everything that is
//   parallel in time: front end, atan2 table, phase difference, wrap, gain, the FIR half of the
//   de-emphasis filter (-> shared memory, floats), and after the recurrence the (int16_t)
//   narrowing and the /4 /4 /2 decimators.  The last warp is the CHAIN warp: lane r walks item r's row
//   of WB_STEP floats in place, 16 bytes at a time (row pitch 260 words: conflict-free LDS.128).
//   Two row buffers make a two-stage pipeline with one __syncthreads per step: while the chain
//   warp is on step t, every item warp narrows/decimates its step t-1 and then produces step t+1
//   into the buffer it has just drained.
//
// VERIFIED SPECULATION (time tiles).  The recurrence has no finite look-back, so a tile after the
// first cannot rebuild it exactly from a halo the way the FIR stages do.  It does not have to
// be exact to be CHECKED: a tile k >= 1 starts two batches (2048 samples) early from y = 0; after
// the first 1024 samples the pole (0.949^1024 ~ 1e-23) has long forgotten the start value, and in
// practice y is bit-identical to the serial one -- but that is a numerical argument, not an
// identity.  So the tile records y at sample emit_from-1025 (the end of its first halo batch,
// "speculated"), the tile before it records its own y at the same sample ("true": by induction
// from tile 0, which starts from the saved state), and rx_wbfm_verify_kernel compares the pairs
// bit for bit.  Equal at that sample and fed the same inputs, the two recurrences are identical
// from there on, so the second halo batch (1024 samples, more than the 704-sample look-back of
// the /4 /4 /2 decimators behind the filter) rebuilds the FIR histories from EXACT values and the
// tile's output is the serial output.  Streams with a differing pair are run again untiled from
// the untouched state_in (run_if / rerun_ids), so the result is bit-exact in every case; the
// re-run costs one empty launch when it is not needed.
//
// The chain costs WB_STEP x ~10 cycles per step and runs beside ~31 x 4 warp-iterations of front
// end, so it is hidden as long as the item warps have at least that much to do (they do: the
// step is HBM- or issue-bound on them).  Results are bit-identical to the serial evaluation:
// same operations, same order, per stream.
#ifndef HRD_WB_THREADS
#define HRD_WB_THREADS 896 // threads per CTA the kernel is compiled for: 896 -> 72 registers, 27 item warps + the chain warp = 7 warps on each of the four schedulers (1024 -> 64 registers; 28 items put 8 warps on one scheduler and ran 8 % slower per item)
#endif
// Row buffers of the item-warp / chain-warp pipeline.  Two: an item warp can be one step ahead of the chain warp.
// Three: two steps (28 rows each then, so that they still fit beside the atan2 table).
#ifndef HRD_RX_WB_NBUF
#define HRD_RX_WB_NBUF 2
#endif
constexpr int WB_NBUF = HRD_RX_WB_NBUF;
// The item warps' front half: 1 = the wide form (four 256 kS/s samples per lane and iteration: 64-byte loads, one
// neighbour shuffle per stage and per detector quantity for twice the samples; needs ~80 registers), 0 = the narrow
// one (two samples per lane; fits 64 registers).
#ifndef HRD_RX_WB_WIDE
#define HRD_RX_WB_WIDE 0
#endif
constexpr bool WB_WIDE = HRD_RX_WB_WIDE != 0;
constexpr int WB_IT = WB_WIDE ? IT_SAMPLES : IT_NARROW; // 256 kS/s samples per warp iteration
constexpr int WB_SPL = WB_IT / 32;                        // ... per lane
constexpr int WB_ROWS = WB_NBUF == 3 ? 28 : 32;
constexpr int WB_ITEMS = WB_NBUF == 3 ? (HRD_WB_THREADS / 32 - 1 < WB_ROWS ? HRD_WB_THREADS / 32 - 1 : WB_ROWS) : HRD_WB_THREADS / 32 - 1;
__device__ __forceinline__ uint32_t wb_buf(uint32_t t) { return WB_NBUF == 3 ? t % 3u : t & 1u; }
constexpr int WB_STEP = 256;            // 256 kS/s samples per pipeline step (4 warp iterations)
constexpr int WB_PITCH = WB_STEP + 4;   // floats per row

struct SmemWbItem {                    // int16 samples, two per word {x[2w], x[2w+1]} (mac_pair's operand)
    uint32_t d64[4 + WB_STEP / 8];     // @64k: 8 samples of history + 64 new per step
    uint32_t a16[19 + WB_STEP / 32];   // @16k: 38 samples of history + 16 new per step
};
// SMALL: the exact re-run after a failed verification (a handful of streams, four per CTA, each walked serially over
// the whole call): eight rows and the atan2 table left in global memory -- 19 KB instead of 207 KB, so that its
// CTAs fit beside the other modes' kernels of a mixed batch instead of waiting for whole SMs to drain (measured on
// the mixed 4096-stream batch: the re-run used to run BEHIND the AM and FM kernels, 2.95 ms per step).
constexpr int WB_RERUN_ITEMS = 4;
template <bool SMALL> struct SmemWbT {
    float f[WB_NBUF][SMALL ? 8 : WB_ROWS][WB_PITCH];
    SmemWbItem item[SMALL ? WB_RERUN_ITEMS : WB_ITEMS];
    // split taps of the 12-tap and the 40-tap decimator, by the lane's share of the taps (see consume)
    alignas(16) uint32_t t12[2][4];
    alignas(16) uint32_t t40[4][8];
    // Upper half of the atan2 table, rows q = 0..128 (row 128 = minus the table's q = -128 row).
    // atan2 is odd in q and the table is the host libm's (double, narrowed to float), which is odd
    // bit for bit -- checked when the table is built (hrd_api.cu ensure_tables) -- so
    // theta(q, i) = sign(q) * lut[|q|][i + 128].  132 KB: the whole table (256 KB) fits neither
    // shared memory nor L1, and at two scattered 4-byte gathers per lane per iteration the L1
    // tag stage, not HBM, was the limiter of this kernel (profiles/: L1 hit rate 52 %).
    float lut[SMALL ? 4 : 129 * 256];
};
typedef SmemWbT<false> SmemWb;

// TILED: the call is cut into time tiles (n_tiles > 1); only that instance carries the verification stores
template <int ENTRY, bool TILED, bool SMALL = false>
__global__ void __launch_bounds__(SMALL ? (WB_RERUN_ITEMS + 1) * 32 : HRD_WB_THREADS, 1) rx_wbfm_kernel(const RxParams p)
{
    typedef typename std::conditional<WB_WIDE, typename RawOf<ENTRY>::type, typename RawNarrowOf<ENTRY>::type>::type Raw;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // the exact re-run after a failed verification: only the streams the verifier listed
    // (p.n_streams bounds it: the tiled retry is sized for a part of the batch, what does not fit goes on to the serial run)
    const int n_streams = p.run_if ? min((int)*p.run_if, p.n_streams) : p.n_streams;
    const int32_t *stream_ids = p.run_if ? p.rerun_ids : p.stream_ids;
    if ((int)blockIdx.x * p.items_per_cta >= n_streams * p.n_tiles) return; // uniform over the CTA
    SmemWbT<SMALL> &sm = *reinterpret_cast<SmemWbT<SMALL> *>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // WARP ROLES.  A warp's scheduler is warp id % 4.  tx_wbfm_kernel puts its chain on the scheduler with the fewest
    // items (its chain is two items' worth of instructions); here the chain is light (two operations per sample,
    // less than one item's worth), so the even split -- items on warps 0.., the chain right behind them, seven
    // or eight warps per scheduler -- balances better: measured 2.445 ms against 2.567 ms (4096 streams x 0.5 s).
#if HRD_RX_WB_ROLES
    const bool chain_warp = warp == 31;
    const int item_of_warp = (warp & 3) != 3 ? (warp >> 2) * 3 + (warp & 3) : 24 + (warp >> 2);
#else // items on warps 0.., the chain right behind them
    const bool chain_warp = warp == p.items_per_cta;
    const int item_of_warp = warp;
#endif
    const bool member = chain_warp || item_of_warp < p.items_per_cta; // takes part in the pipeline's barriers
    const int n_items = n_streams * p.n_tiles;

    // ---- this thread's item: the warp's (item warps) or the lane's (chain warp) ---------
    const uint32_t tile_len = p.tile_batches * BATCH256;
    const uint32_t halo = (uint32_t)HaloOf<K_WBFM>::value * BATCH256;
    const int row = chain_warp ? lane : item_of_warp;
    const int item = blockIdx.x * p.items_per_cta + row;
    const bool live = row < p.items_per_cta && item < n_items;
    int tile = 0, sid = 0, slot = 0;
    uint32_t start = 0, end = 0, emit_from = 0;
    if (live) {
        tile = item / n_streams;
        slot = item - tile * n_streams;
        sid = stream_ids[slot];
        emit_from = (uint32_t)tile * tile_len;
        end = min(p.n256_of ? p.n256_of[sid] : p.n256, emit_from + tile_len);
        start = tile == 0 ? 0u : emit_from - halo;
    }
    const bool first = tile == 0, last = tile == p.n_tiles - 1;
    // the same for every thread of the CTA: steps of the longest item
    const uint32_t max_len = min(p.n256, tile_len + (p.n_tiles > 1 ? halo : 0u));
    const uint32_t n_steps = (max_len + WB_STEP - 1) / WB_STEP;
    const RxState &st = p.state_in[sid];

    // ---- chain warp state ----------------------------------------------------------------
    float y1 = 0.f;
    if (chain_warp && live && first) y1 = st.wb_y1;
    // A later tile's HINT: the value the call started from.  A stream that has been silent (discriminator output
    // exactly zero) since before the call sits on a vanishing value -- a denormal the rounded recurrence no longer
    // shrinks, or zero; a tile's warm-up from zero histories begins with a start transient and is still on that
    // transient's tail (~1e-22) at the check point: different bits in every call, although both are nothing.
    // When the hint vanishes and the warmed-up value is that small, the tile takes the hint.  (Any value may be
    // tried at the check point: the verification decides, so this costs time at worst.)
    const float y_hint = (TILED && chain_warp && live && !first) ? st.wb_y1 : 1.f;

    // ---- item warp state -----------------------------------------------------------------
    SmemWbItem &it = sm.item[chain_warp || !member ? 0 : item_of_warp];
    const int8_t *src = p.iq + (size_t)sid * p.iq_stride;
    asm volatile("" : "+l"(src));
    FeCarry fc;
    fc.t = fc.v = fc.u = 0u;
    FeTaps fk;
    fk.a0 = c_tab.fe_a[0]; fk.b0 = c_tab.fe_b[0];
    fk.a1 = c_tab.fe_a[1]; fk.b1 = c_tab.fe_b[1];
    fk.a2 = c_tab.fe_a[2]; fk.b2 = c_tab.fe_b[2];
    float scale = 0.f, th_keep = 0.f, v_keep = 0.f, m_keep = 0.f;
    uint32_t keep2 = 0u, keep3 = 0u; // the last four narrowed samples @256k (history of the /4 decimator)
    int last_nl = 32;                // lanes that were live in the most recent consume
    uint32_t sched_dep = 0;          // see transpose_after
    bool narrow_fast = false;
    constexpr uint32_t BPS = ENTRY == 0 ? 16 : 2;
    uint32_t pf = 0, pf_last = 0, last_active = 32;
    Raw buf[WB_DEPTH];
    if (!chain_warp && live) {
        if (first) {
            fc.t = st.fe_t; fc.v = st.fe_v; fc.u = st.fe_u;
            th_keep = st.wb_prev_theta;
            v_keep = st.wb_x1;
        }
        if (first) {
            keep2 = reinterpret_cast<const uint32_t *>(st.wb_d256)[0];
            keep3 = reinterpret_cast<const uint32_t *>(st.wb_d256)[1];
        }
        ring_init(it.d64, reinterpret_cast<const uint32_t *>(st.wb_d64), 4, lane, first);
        ring_init(it.a16, reinterpret_cast<const uint32_t *>(st.wb_a16), 19, lane, first);
        // WbFmDemodulator.cc:392-395
        scale = __fmul_rn(__fdiv_rn(p.gain[sid], 75000.f), 32767.f);
        m_keep = __fmul_rn(0.0253863f, v_keep);
        // |y| <= max(|y[-1]|, pi * scale * (1 + 1e-6)): the de-emphasis filter has unit DC gain and a
        // positive impulse response, and |x| <= pi * scale.  Below 2^31 the (int16_t) narrowing needs no
        // out-of-range patch (f32_to_i16).  NaN gains fail the test and take the patched path.
        narrow_fast = scale < 0x1p27f && fabsf(first ? st.wb_y1 : (TILED && p.wb_guess) ? p.wb_guess[slot] : 0.f) < 0x1p30f;
        pf = (start + WB_SPL * lane) * BPS;
        pf_last = end >= WB_SPL ? (end - WB_SPL) * BPS : 0u;
#pragma unroll
        for (int d = 0; d < WB_DEPTH; d++) {
            if constexpr (WB_WIDE && ENTRY >= 0) buf[d] = load_raw<ENTRY>(src, pf, pf_last);
            else buf[d] = load_raw_narrow<ENTRY>(src, pf, pf_last);
            pf += WB_IT * BPS;
        }
    }
    __syncwarp();

    // one warp iteration: 64 samples at 256 kS/s -> 64 floats of the row (FIR half done)
    auto iter_narrow = [&](auto &b, float *dst) {
        uint32_t word;
        if constexpr (ENTRY == 0) {
            uint32_t t[8];
#pragma unroll
            for (int r = 0; r < 8; r++) t[r] = transpose_after(b.v[r], sched_dep);
#if HRD_RX_WB_LDDEP
            b = load_raw_narrow<ENTRY>(src, offset_after(pf, t[0]), pf_last);
#else
            b = load_raw_narrow<ENTRY>(src, pf, pf_last);
#endif
            word = front_end_iter_narrow(t, fc, fk, lane);
#if HRD_RX_WB_DEP
            sched_dep = word;
#endif
        } else {
            word = __byte_perm(b, 0, 0x3120);
            b = load_raw_narrow<ENTRY>(src, pf, pf_last);
        }
        pf += IT_NARROW * BPS;
#if HRD_EXP & 4
        *reinterpret_cast<float2 *>(dst + 2 * lane) = make_float2(__int_as_float(word), 0.f);
        return;
#endif
        // theta = atan2LookupTable[(uint8_t)Q + 128][(uint8_t)I + 128]  (WbFmDemodulator.cc:403-406)
        // word = {I0, I1, Q0, Q1}: sign-extend Q, fold the table on |Q|, put the sign back
        int q0, q1; // prmt with the sign-replicate bit (8) set in three selector nibbles: one-instruction sext of a byte
        asm("prmt.b32 %0, %1, 0, 0xaaa2;" : "=r"(q0) : "r"(word));
        asm("prmt.b32 %0, %1, 0, 0xbbb3;" : "=r"(q1) : "r"(word));
        const uint32_t x = word ^ 0x00008080u;
        // (the small instance reads the table where it lies in global memory: rows q >= 0 are rows 128 + q there,
        //  and |q| = 128 is minus row 0)
        auto lut_at = [&](int aq, int col) {
            if constexpr (SMALL) return aq < 128 ? __ldg(p.atan2_lut + (128 + aq) * 256 + col) : -__ldg(p.atan2_lut + col);
            else return sm.lut[aq * 256 + col];
        };
        const float a0 = lut_at(abs(q0), (int)(x & 0xffu));
        const float a1 = lut_at(abs(q1), (int)__byte_perm(x, 0, 0x4441));
        const float th0 = __int_as_float(__float_as_int(a0) ^ (q0 & (int)0x80000000));
        const float th1 = __int_as_float(__float_as_int(a1) ^ (q1 & (int)0x80000000));
        // theta of the previous sample: previous lane's th1 (lane 0: kept from before)
        const float sel = (lane == 31) ? th_keep : th1;
        const float thp = __shfl_sync(HRD_FULL_MASK, sel, (lane + 31) & 31);
        const float d0 = wrap_pi_select(__fsub_rn(th0, thp));
        const float d1 = wrap_pi_select(__fsub_rn(th1, th0));
        const float v0 = __fmul_rn(scale, d0), v1 = __fmul_rn(scale, d1);
        // FirFilter::filterData order: y = 0 + b0*x[n]; y = y + b1*x[n-1]   (FirFilter.cc:161-164).
        // b0 == b1, so b1*x[n-1] is the previous sample's b0*x[n]: the product crosses lanes, not x.
        const float bb = 0.0253863f;
        const float m0 = __fmul_rn(bb, v0), m1 = __fmul_rn(bb, v1);
        const float selm = (lane == 31) ? m_keep : m1;
        const float mp = __shfl_sync(HRD_FULL_MASK, selm, (lane + 31) & 31);
        th_keep = th1; // the state save reads them from the last live lane
        v_keep = v1;
        m_keep = m1;
        float2 o;
        o.x = __fadd_rn(m0, mp);
        o.y = __fadd_rn(m1, m0);
        *reinterpret_cast<float2 *>(dst + 2 * lane) = o;
    };

    // the same for four samples per lane: 128 samples at 256 kS/s -> 128 floats of the row
    auto iter_wide = [&](auto &b, float *dst) {
        uint2 words;
        if constexpr (ENTRY == 0) {
            uint32_t t[16];
#pragma unroll
            for (int r = 0; r < 8; r++) {
                t[r] = transpose_after(b.a.v[r], sched_dep);
                t[8 + r] = transpose_after(b.b.v[r], sched_dep);
            }
            b = load_raw<ENTRY>(src, pf, pf_last);
            words = front_end_iter(t, fc, fk, lane);
#if HRD_RX_WB_DEP
            sched_dep = words.y;
#endif
        } else {
            words = make_uint2(__byte_perm(b.x, 0, 0x3120), __byte_perm(b.y, 0, 0x3120));
            b = load_raw<ENTRY>(src, pf, pf_last);
        }
        pf += IT_SAMPLES * BPS;
        // theta = atan2LookupTable[(uint8_t)Q + 128][(uint8_t)I + 128]  (WbFmDemodulator.cc:403-406), as in iter_narrow
        auto lut_at = [&](int aq, int col) {
            if constexpr (SMALL) return aq < 128 ? __ldg(p.atan2_lut + (128 + aq) * 256 + col) : -__ldg(p.atan2_lut + col);
            else return sm.lut[aq * 256 + col];
        };
        float th[4];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t word = h ? words.y : words.x;
            int q0, q1;
            asm("prmt.b32 %0, %1, 0, 0xaaa2;" : "=r"(q0) : "r"(word));
            asm("prmt.b32 %0, %1, 0, 0xbbb3;" : "=r"(q1) : "r"(word));
            const uint32_t x = word ^ 0x00008080u;
            const float a0 = lut_at(abs(q0), (int)(x & 0xffu));
            const float a1 = lut_at(abs(q1), (int)__byte_perm(x, 0, 0x4441));
            th[2 * h] = __int_as_float(__float_as_int(a0) ^ (q0 & (int)0x80000000));
            th[2 * h + 1] = __int_as_float(__float_as_int(a1) ^ (q1 & (int)0x80000000));
        }
        const float sel = (lane == 31) ? th_keep : th[3];
        const float thp = __shfl_sync(HRD_FULL_MASK, sel, (lane + 31) & 31);
        float m[4], v3 = 0.f;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float d = wrap_pi_select(__fsub_rn(th[i], i ? th[i - 1] : thp));
            const float v = __fmul_rn(scale, d);
            m[i] = __fmul_rn(0.0253863f, v); // FirFilter::filterData, b0 == b1: see iter_narrow
            if (i == 3) v3 = v;
        }
        const float selm = (lane == 31) ? m_keep : m[3];
        const float mp = __shfl_sync(HRD_FULL_MASK, selm, (lane + 31) & 31);
        th_keep = th[3];
        v_keep = v3;
        m_keep = m[3];
        float4 o;
        o.x = __fadd_rn(m[0], mp);
        o.y = __fadd_rn(m[1], m[0]);
        o.z = __fadd_rn(m[2], m[1]);
        o.w = __fadd_rn(m[3], m[2]);
        *reinterpret_cast<float4 *>(dst + 4 * lane) = o;
    };
    auto iter = [&](Raw &b, float *dst) {
        if constexpr (WB_WIDE && ENTRY >= 0) iter_wide(b, dst); // (a dependent condition: the other form is not instantiated)
        else iter_narrow(b, dst);
    };

    // step t of this item: samples [start + t*WB_STEP, ...) -> row of buffer t&1
    auto produce = [&](uint32_t t) {
        const uint32_t done = start + t * WB_STEP;
        if (done >= end) return;
        const uint32_t nb = min((uint32_t)WB_STEP, end - done);
        const uint32_t n_it = (nb + WB_IT - 1) / WB_IT;
        last_active = min(32u, (nb - (n_it - 1) * WB_IT) / WB_SPL);
        if constexpr (ENTRY == 1) { // 256 kS/s entry: a step is 512 bytes, four lines; pulled into L2 eight steps ahead
            if (lane < 4) prefetch_l2(src + min((done + 8u * WB_STEP) * BPS + 128u * (uint32_t)lane, pf_last));
        }
        if constexpr (ENTRY == 0) { // a step is 4 KiB: lane l pulls its line l, one step ahead (pf = the lane's own next offset)
            if constexpr (WB_WIDE) prefetch_l2(src + min(pf + 64u * (uint32_t)lane + RX_L2_AHEAD * 1024u, pf_last));
            else prefetch_chunk_narrow(src, pf, pf_last, lane);
        }
        float *dst = sm.f[wb_buf(t)][row];
        if (n_it == WB_STEP / WB_IT) { // a full step, unrolled: the in-place refill of buf needs no register moves
#pragma unroll
            for (uint32_t i = 0; i < WB_STEP / WB_IT; i++) iter(buf[i % WB_DEPTH], dst + i * WB_IT);
        } else {
            for (uint32_t i = 0; i < n_it; i++) iter(buf[i % WB_DEPTH], dst + i * WB_IT);
        }
        __syncwarp();
    };

    // WbFmDemodulator.cc:460-500 on step t: (int16_t) narrowing, /4 (8 taps) /4 (12) /2 (40).
    // Register-blocked: lane L owns samples 8L..8L+7 of the step, narrows them, packs them two per word
    // and feeds the decimators with mac_pair (two taps per instruction pair, no unpacking).
    auto consume = [&](uint32_t t) {
        const uint32_t done = start + t * WB_STEP;
        if (done >= end) return;
        const int nb = (int)min((uint32_t)WB_STEP, end - done); // multiple of 32
        const int nl = nb >> 3;                                   // live lanes, a multiple of 4
        last_nl = nl;
        const float *y = sm.f[wb_buf(t)][row] + 8 * lane;         // lanes >= nl read stale floats and store nothing
        const float4 ya = *reinterpret_cast<const float4 *>(y), yb = *reinterpret_cast<const float4 *>(y + 4);
        const float yv[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
        int v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = __float2int_rz(yv[i]);
        if (!narrow_fast) { // (int16_t)float of an out-of-range value (hrd_device.cuh f32_to_i16)
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (!(yv[i] < 2147483648.0f)) v[i] = 0;
        }
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; i++) w[i] = merge16((uint32_t)v[2 * i], (uint32_t)v[2 * i + 1]);
        // /4, 8 taps: output j reads samples 4j-4 .. 4j+3; this lane has outputs 2L (half from the lane
        // before) and 2L+1
        const uint32_t l2 = from_left(w[2], keep2, lane), l3 = from_left(w[3], keep3, lane);
        int al = 1 << 14, ah = 0;
        mac_pair(l2, c_tab.wb1_sp[0], al, ah);
        mac_pair(l3, c_tab.wb1_sp[1], al, ah);
        mac_pair(w[0], c_tab.wb1_sp[2], al, ah);
        mac_pair(w[1], c_tab.wb1_sp[3], al, ah);
        const int e0 = (al + (ah << 8)) >> 15;
        al = 1 << 14, ah = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) mac_pair(w[i], c_tab.wb1_sp[i], al, ah);
        const int e1 = (al + (ah << 8)) >> 15;
        if (lane < nl) it.d64[4 + lane] = merge16((uint32_t)e0, (uint32_t)e1);
        __syncwarp();
        {
            // /4, 12 taps: output k reads ring words 2k .. 2k+5; two lanes share an output (the int32
            // accumulation wraps, so the order of the sum is free)
            const int k = lane >> 1, h = lane & 1;
            const uint4 b = *reinterpret_cast<const uint4 *>(sm.t12[h]);
            const uint32_t *r = it.d64 + 2 * k + 3 * h;
            al = h ? 0 : (1 << 14), ah = 0;
            mac_pair(r[0], b.x, al, ah);
            mac_pair(r[1], b.y, al, ah);
            mac_pair(r[2], b.z, al, ah);
            int acc = al + (ah << 8);
            acc += __shfl_xor_sync(HRD_FULL_MASK, acc, 1);
            const int o = acc >> 15;
            const int o_next = __shfl_down_sync(HRD_FULL_MASK, o, 2);
            if ((lane & 3) == 0 && lane < nl) it.a16[19 + (lane >> 2)] = merge16((uint32_t)o, (uint32_t)o_next);
        }
        __syncwarp();
        {
            // audio decimator /2, 40 taps: output k reads ring words k .. k+19; four lanes share an output
            const int k = lane >> 2, part = lane & 3;
            const uint4 b = *reinterpret_cast<const uint4 *>(sm.t40[part]);
            const uint32_t b4 = sm.t40[part][4];
            const uint32_t *r = it.a16 + k + 5 * part;
            al = part ? 0 : (1 << 14), ah = 0;
            mac_pair(r[0], b.x, al, ah);
            mac_pair(r[1], b.y, al, ah);
            mac_pair(r[2], b.z, al, ah);
            mac_pair(r[3], b.w, al, ah);
            mac_pair(r[4], b4, al, ah);
            int acc = al + (ah << 8);
            acc += __shfl_xor_sync(HRD_FULL_MASK, acc, 1);
            acc += __shfl_xor_sync(HRD_FULL_MASK, acc, 2);
            if (done >= emit_from && part == 0 && lane < nl)
                p.pcm[(size_t)sid * p.pcm_stride + done / 32 + k] = (int16_t)q15(acc);
        }
        __syncwarp();
        ring_shift(it.d64, 4, nl, lane);
        ring_shift(it.a16, 19, nb / 32, lane);
    };

    // IirFilter.cc:161-176 with a = {-0.9492274f}: lane = item, in place, 32 samples per round
    auto chain = [&](uint32_t t) {
        const uint32_t done = start + t * WB_STEP;
        if (done >= end) return;
        const uint32_t nb = min((uint32_t)WB_STEP, end - done); // multiple of 32
        float *r = sm.f[wb_buf(t)][lane];
        // (All eight loads of a 32-sample round first.  Reading the row a few groups AHEAD of its use instead was
        // measured and lost: ptxas puts every shared-memory load of the loop on one counting scoreboard, so the
        // consumer of an old load also waits for the one just issued -- a full load latency per group of four
        // samples instead of one per round.)
        for (uint32_t c = 0; c < nb; c += 32) {
            float4 v[8];
#pragma unroll
            for (int g = 0; g < 8; g++) v[g] = *reinterpret_cast<float4 *>(r + c + 4 * g);
#pragma unroll
            for (int g = 0; g < 8; g++) {
                v[g].x = y1 = __fsub_rn(v[g].x, __fmul_rn(-0.9492274f, y1));
                v[g].y = y1 = __fsub_rn(v[g].y, __fmul_rn(-0.9492274f, y1));
                v[g].z = y1 = __fsub_rn(v[g].z, __fmul_rn(-0.9492274f, y1));
                v[g].w = y1 = __fsub_rn(v[g].w, __fmul_rn(-0.9492274f, y1));
                *reinterpret_cast<float4 *>(r + c + 4 * g) = v[g];
            }
        }
        // verified speculation: y at the two check points of this tile (see the kernel's header)
        if constexpr (TILED) {
            const uint32_t pos = done + nb;
            if (tile >= 1 && pos == start + BATCH256) {
                if (p.wb_guess) y1 = p.wb_guess[slot]; // the retry: a given value instead of the warmed-up one
                else if (fabsf(y_hint) < 0x1p-100f && fabsf(y1) < 0x1p-60f) y1 = y_hint;
                p.wb_verify[(size_t)slot * p.n_tiles + tile].x = y1;
            }
            if (tile + 1 < p.n_tiles && pos == end - BATCH256) p.wb_verify[(size_t)slot * p.n_tiles + tile + 1].y = y1;
        }
    };

    if (threadIdx.x < 6) sm.t12[threadIdx.x / 3][threadIdx.x % 3] = c_tab.fm_post_sp[threadIdx.x];
    if (threadIdx.x < 20) sm.t40[threadIdx.x / 5][threadIdx.x % 5] = c_tab.audio40_sp[threadIdx.x];
    if constexpr (!SMALL)
        for (int i = threadIdx.x; i < (int)(sizeof(sm.lut) / 4); i += blockDim.x) // table rows q = 0..127 are rows 128..255; row 128 = -row 0
            sm.lut[i] = i < 128 * 256 ? __ldg(p.atan2_lut + 128 * 256 + i) : -__ldg(p.atan2_lut + (i - 128 * 256));
    __syncthreads(); // the tables are complete before any warp looks an angle up
    if (!member) return; // (a spare warp slot: exited warps do not count at later barriers)
    // Hand-over between the item warps and the chain warp: two pairs of named barriers (by step parity)
    // instead of one __syncthreads per step.  Item warps ARRIVE on "produced" and go on; only the chain
    // warp waits there.  The chain warp arrives on "chained" after its step; item warps wait there before
    // they narrow that step.  No item warp ever waits for another item warp's consume, so a slow warp
    // costs the CTA nothing as long as it keeps within a step of the others.
    const int bar_threads = (p.items_per_cta + 1) * 32;
    if constexpr (WB_NBUF == 3) { // three row buffers: the item warps run up to two steps ahead of the chain warp
        if (chain_warp) {
            for (uint32_t t = 0; t < n_steps; t++) {
                named_bar_sync3(HRD_BAR3_PRODUCED, t % 3u, bar_threads);
#if !(HRD_EXP & 1)
                if (live) chain(t);
#endif
                named_bar_arrive3(HRD_BAR3_CHAINED, t % 3u, bar_threads);
            }
        } else {
            if (live) produce(0);
            named_bar_arrive3(HRD_BAR3_PRODUCED, 0, bar_threads);
            if (n_steps > 1) {
                if (live) produce(1);
                named_bar_arrive3(HRD_BAR3_PRODUCED, 1, bar_threads);
            }
            for (uint32_t t = 0; t < n_steps; t++) {
                if (t + 2 < n_steps) {
                    if (live) produce(t + 2); // into the buffer this warp drained in consume(t - 1)
                    named_bar_arrive3(HRD_BAR3_PRODUCED, (t + 2) % 3u, bar_threads);
                }
                named_bar_sync3(HRD_BAR3_CHAINED, t % 3u, bar_threads);
#if !(HRD_EXP & 2)
                if (live) consume(t);
#endif
            }
        }
    } else if (chain_warp) {
        for (uint32_t t = 0; t < n_steps; t++) {
            named_bar_sync(HRD_BAR_PRODUCED, t, bar_threads);
#if !(HRD_EXP & 1)
            if (live) chain(t);
#endif
            named_bar_arrive(HRD_BAR_CHAINED, t, bar_threads);
        }
    } else {
        if (live) produce(0);
        named_bar_arrive(HRD_BAR_PRODUCED, 0, bar_threads);
        for (uint32_t t = 0; t < n_steps; t++) {
            if (t + 1 < n_steps) {
                if (live) produce(t + 1); // into the buffer this warp drained in consume(t - 1)
                named_bar_arrive(HRD_BAR_PRODUCED, t + 1, bar_threads);
            }
            named_bar_sync(HRD_BAR_CHAINED, t, bar_threads);
#if !(HRD_EXP & 2)
            if (live) consume(t);
#endif
        }
    }

    // ---- the last tile leaves the stream's state for the next call -----------------------
    RxState &so = p.state_out[sid];
    if (!chain_warp && live && last) {
        const uint32_t *a = reinterpret_cast<const uint32_t *>(&st);
        uint32_t *b = reinterpret_cast<uint32_t *>(&so);
        for (int i = lane; i < (int)(sizeof(RxState) / 4); i += 32) b[i] = a[i];
        __syncwarp();
        ring_save_hist(it.d64, reinterpret_cast<uint32_t *>(so.wb_d64), 4, lane);
        ring_save_hist(it.a16, reinterpret_cast<uint32_t *>(so.wb_a16), 19, lane);
        if (lane == last_nl - 1) { // from_left left every lane's own last two words in keep2/keep3
            reinterpret_cast<uint32_t *>(so.wb_d256)[0] = keep2;
            reinterpret_cast<uint32_t *>(so.wb_d256)[1] = keep3;
        }
        if (lane == (int)last_active - 1) {
            if constexpr (ENTRY == 0) {
                so.fe_t = fc.t;
                so.fe_v = fc.v;
                so.fe_u = fc.u;
            }
            so.wb_prev_theta = th_keep;
            so.wb_x1 = v_keep;
        }
    }
    __syncthreads(); // the record copy above must land before the chain warp's y[n-1]
    if (chain_warp && live && last) so.wb_y1 = y1;
}

// One thread per stream: speculated vs true recurrence value of every tile >= 1, bit for bit.  A stream with
// any difference is appended to the re-run list.  (Differences are not an error: a stream whose
// discriminator output is exactly zero for long stretches -- constant input -- leaves the filter in a
// slowly decaying denormal tail that a warm-up from zero cannot reproduce bit for bit.)
// VANISHING VALUES count as equal.  With a constant input the discriminator's output is exactly zero, the true
// recurrence value is a decaying tail that has long left the normal range and the warmed-up one is exactly zero:
// different bits, but the same output for ever after.  While the filter's input stays zero both values only shrink,
// and (int16_t) of either is 0; the first non-zero input f has |f| >= bb * scale * 2^-24 (one ulp of an atan2 table
// value times the two gains), so with scale >= 2^-20 adding 0.949 * y with |y| < 2^-100 does not change a bit of
// it, and from there on the two recurrences are identical.  Without this rule every silent stream took the
// untiled re-run (measured: 32 of 1024 streams, the launch 4 x slower).
__global__ void rx_wbfm_verify_kernel(const float2 *pairs, const int32_t *stream_ids, const float *gain, int n_streams, int n_tiles,
                                      const uint32_t *n_if, uint32_t *count, int32_t *rerun_ids, float *guess_out,
                                      unsigned long long *fallbacks, unsigned long long *fallbacks2, int force)
{
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= (n_if ? (int)*n_if : n_streams)) return;
    if (slot >= n_streams) { // listed, but the retry had no room for it
        rerun_ids[atomicAdd(count, 1u)] = stream_ids[slot];
        atomicAdd(fallbacks, 1ull);
        return;
    }
    bool bad = force != 0;
    const bool tiny_ok = fabsf(gain[stream_ids[slot]]) >= 1e-3f; // scale = gain / 75000 * 32767 >= 2^-20 with room (NaN: false)
    for (int t = 1; t < n_tiles; t++) {
        const float2 v = pairs[(size_t)slot * n_tiles + t];
        const bool vanishing = tiny_ok && fabsf(v.x) < 0x1p-100f && fabsf(v.y) < 0x1p-100f;
        bad |= __float_as_uint(v.x) != __float_as_uint(v.y) && !vanishing;
    }
    if (bad) {
        const uint32_t at = atomicAdd(count, 1u);
        rerun_ids[at] = stream_ids[slot];
        if (guess_out) guess_out[at] = pairs[(size_t)slot * n_tiles + 1].y; // tile 0's value: true by construction
        atomicAdd(fallbacks, 1ull);
        if (fallbacks2) atomicAdd(fallbacks2, 1ull);
    }
}

// ------------------------------------------------------------------------------------
// AM / SSB DC-removal IIR (AmDemodulator.cc:67-68,460-465, SsbDemodulator.cc:104-105,586-592;
// Filters/IirFilter.cc:161-176): one thread per stream walks the call's PCM samples in order,
//   fir = x[n] - x[n-1];  y = fir - (-0.95f * y[n-1]);  pcm = (int16_t)(gain * y)
// every operation a rounded fp32 one, as the reference evaluates it.  Must run after the tile
// kernel of the same call (same CUDA stream); it finishes the state record the tile kernel
// started in state_out.
// ------------------------------------------------------------------------------------
// One CTA runs 32 streams.  The recurrence is a dependent FMUL+FSUB per sample, so the time of
// a call is (samples per stream) x (chain latency) no matter how many streams there are; what
// can be done is to keep everything else OFF the warp that walks the chain.  Six warps form a
// software pipeline over chunks of 64 samples per stream, one __syncthreads per step:
//   warp 0   loader: cp.async brings chunk t+3 in ([32 rows][64 floats], coalesced);
//   warp 1   chain: lane r walks row r of chunk t in place, y = fir - (-0.95f * y1), reading
//            and writing shared memory 16 bytes at a time (pitch 68 words: conflict-free);
//   warps 2..5 post: chunk t-1, element-parallel: pcm = (int16_t)(gain * y), 16-byte stores.
constexpr int IIR_CHUNK = 64;   // samples per stream per pipeline step
constexpr int IIR_AHEAD = 3;    // chunks in flight ahead of the chain
constexpr int IIR_STAGES = IIR_AHEAD + 2;
constexpr int IIR_PITCH = 68;   // words per staged row (16-byte aligned, 17 x 4: conflict-free LDS.128)
constexpr int IIR_POST_THREADS = 128; // warps 2..5: the narrowing/store stage must not outlast the chain warp

struct SmemIir {
    float f[IIR_STAGES][32][IIR_PITCH];
    int16_t *prow[32];
    float gain[32];
    uint32_t cnt[32]; // PCM samples of each row (ragged calls: RxParams::n256_of)
};


__global__ void __launch_bounds__(64 + IIR_POST_THREADS) rx_dc_iir_kernel(const RxParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemIir &sm = *reinterpret_cast<SmemIir *>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int rows_live = min(32, p.n_streams - (int)blockIdx.x * 32);
    const uint32_t n = p.n256 / 32;
    const uint32_t n_pad = (n + 7) & ~7u; // rows of pre_iir are padded to 8 samples
    const uint32_t n_chunks = (n + IIR_CHUNK - 1) / IIR_CHUNK;

    float y1 = 0.f; // warp 1: the recurrence state of row `lane`
    bool vec_out = true;
    // warp 0: element offsets of the 16 rows this lane copies segments of (rows 2i + lane/16)
    uint32_t row_off[16];
    {
        // rows past the end of the launch shadow its last stream and store nothing
        const int sid = p.stream_ids[blockIdx.x * 32 + min(lane, rows_live - 1)];
        const bool ssb = p.kind_of[sid] == K_SSB;
        int16_t *pcm = p.pcm + (size_t)sid * p.pcm_stride;
        vec_out = __all_sync(HRD_FULL_MASK, (reinterpret_cast<uintptr_t>(pcm) & 15) == 0);
        if (warp == 0) {
            sm.prow[lane] = pcm;
            sm.gain[lane] = ssb ? p.gain_ssb[sid] : p.gain[sid];
            sm.cnt[lane] = lane < rows_live ? (p.n256_of ? p.n256_of[sid] : p.n256) / 32 : 0u;
        }
        y1 = ssb ? p.state_in[sid].ssb.y1 : p.state_in[sid].am.y1;
        const uint32_t my_row = (uint32_t)((size_t)sid * p.pre_stride); // the host keeps n * pre_stride < 2^32
#pragma unroll
        for (int i = 0; i < 16; i++) // row 2i + lane/16 belongs to lane 2i + lane/16; keep this lane's column
            row_off[i] = __shfl_sync(HRD_FULL_MASK, my_row, 2 * i + (lane >> 4)) + (uint32_t)(lane & 15) * 4;
    }

    auto issue = [&](uint32_t c) { // warp 0: chunk c -> f[c % IIR_STAGES], 16 x 16 bytes per row
        if (c < n_chunks) {
            float(*dst)[IIR_PITCH] = sm.f[c % IIR_STAGES];
            const uint32_t at = c * IIR_CHUNK + (lane & 15) * 4;
            if (at < n_pad) {
#pragma unroll
                for (int i = 0; i < 16; i++)
                    cp_async16(&dst[2 * i + (lane >> 4)][(lane & 15) * 4], p.pre_iir + row_off[i] + c * IIR_CHUNK);
            }
        }
        cp_async_commit(); // an empty group keeps the wait arithmetic uniform
    };
    if (warp == 0) {
#pragma unroll
        for (int c = 0; c < IIR_AHEAD; c++) issue(c);
        cp_async_wait<IIR_AHEAD - 1>(); // chunk 0 has landed
    }
    __syncthreads();

    for (uint32_t t = 0; t < n_chunks + 1; t++) {
        if (warp == 0) {
            issue(t + IIR_AHEAD);
            cp_async_wait<IIR_AHEAD - 1>(); // chunk t+1 has landed (visible to the others after the barrier)
        } else if (warp == 1) {
            if (t < n_chunks) {
                float *row = sm.f[t % IIR_STAGES][lane];
                const uint32_t mine = sm.cnt[lane], from = t * IIR_CHUNK; // this row's samples in this chunk
                const uint32_t m = mine > from ? min((uint32_t)IIR_CHUNK, mine - from) : 0u;
                uint32_t k = 0;
                if (__all_sync(HRD_FULL_MASK, m == IIR_CHUNK)) {
                    // the whole row into registers first: one shared-memory latency per chunk, then
                    // nothing but the dependent FMUL+FSUB pairs (IirFilter.cc:161-176, a0 = -0.95f)
                    float4 v[IIR_CHUNK / 4];
#pragma unroll
                    for (int g = 0; g < IIR_CHUNK / 4; g++) v[g] = *reinterpret_cast<float4 *>(row + 4 * g);
#pragma unroll
                    for (int g = 0; g < IIR_CHUNK / 4; g++) {
                        v[g].x = y1 = __fsub_rn(v[g].x, __fmul_rn(-0.95f, y1));
                        v[g].y = y1 = __fsub_rn(v[g].y, __fmul_rn(-0.95f, y1));
                        v[g].z = y1 = __fsub_rn(v[g].z, __fmul_rn(-0.95f, y1));
                        v[g].w = y1 = __fsub_rn(v[g].w, __fmul_rn(-0.95f, y1));
                        *reinterpret_cast<float4 *>(row + 4 * g) = v[g];
                    }
                    k = IIR_CHUNK;
                }
                for (; k < m; k++) row[k] = y1 = __fsub_rn(row[k], __fmul_rn(-0.95f, y1));
            }
        } else {
            if (t >= 1) {
                const uint32_t c = t - 1;
                const float(*src)[IIR_PITCH] = sm.f[c % IIR_STAGES];
                const uint32_t base = c * IIR_CHUNK;
                const int pl = threadIdx.x - 64; // 0..IIR_POST_THREADS-1: (row, 8 samples) items, 256 per chunk
#pragma unroll
                for (int i = 0; i < 256 / IIR_POST_THREADS; i++) {
                    const int r = (IIR_POST_THREADS / 8) * i + (pl >> 3), part = pl & 7;
                    const uint32_t at = base + part * 8;
                    const uint32_t n_r = sm.cnt[r];
                    if (r < rows_live && at < n_r) {
                        const float g = sm.gain[r];
                        const float4 a = *reinterpret_cast<const float4 *>(&src[r][part * 8]);
                        const float4 b = *reinterpret_cast<const float4 *>(&src[r][part * 8 + 4]);
                        // AmDemodulator.cc:465 / SsbDemodulator.cc:592: (int16_t)(gain * y)
                        const int o0 = f32_to_i16(__fmul_rn(g, a.x)), o1 = f32_to_i16(__fmul_rn(g, a.y));
                        const int o2 = f32_to_i16(__fmul_rn(g, a.z)), o3 = f32_to_i16(__fmul_rn(g, a.w));
                        const int o4 = f32_to_i16(__fmul_rn(g, b.x)), o5 = f32_to_i16(__fmul_rn(g, b.y));
                        const int o6 = f32_to_i16(__fmul_rn(g, b.z)), o7 = f32_to_i16(__fmul_rn(g, b.w));
                        int16_t *dst = sm.prow[r] + at;
                        if (vec_out && at + 8 <= n_r) {
                            int4 w;
                            w.x = (int)(((uint32_t)o0 & 0xffffu) | ((uint32_t)o1 << 16));
                            w.y = (int)(((uint32_t)o2 & 0xffffu) | ((uint32_t)o3 << 16));
                            w.z = (int)(((uint32_t)o4 & 0xffffu) | ((uint32_t)o5 << 16));
                            w.w = (int)(((uint32_t)o6 & 0xffffu) | ((uint32_t)o7 << 16));
                            *reinterpret_cast<int4 *>(dst) = w;
                        } else {
                            const int o[8] = {o0, o1, o2, o3, o4, o5, o6, o7};
#pragma unroll
                            for (int k = 0; k < 8; k++)
                                if (at + k < n_r) dst[k] = (int16_t)o[k];
                        }
                    }
                }
            }
        }
        __syncthreads();
    }

    // y[n-1] for the next call (x[n-1] was left by the tile kernel)
    if (warp == 1 && lane < rows_live) {
        const int sid = p.stream_ids[blockIdx.x * 32 + lane];
        const bool ssb = p.kind_of[sid] == K_SSB;
        (ssb ? p.state_out[sid].ssb : p.state_out[sid].am).y1 = y1;
    }
}

template <int KIND, int ENTRY>
int launch_one(const RxParams &p, cudaStream_t s)
{
    const long long items = (long long)p.n_streams * p.n_tiles;
    const int grid = (int)((items + HRD_WARPS_PER_CTA - 1) / HRD_WARPS_PER_CTA);
    typedef typename SmemOf<KIND>::type Smem;
    const size_t smem = sizeof(Smem) * HRD_WARPS_PER_CTA;
    rx_kernel<KIND, ENTRY><<<grid, HRD_WARPS_PER_CTA * 32, smem, s>>>(p);
    return (int)cudaGetLastError();
}

} // namespace

int rx_halo_batches(int kind)
{
    switch (kind) {
    case K_AM: return 2; // the AM launch also carries the SSB streams (31-tap Hilbert at 8 kS/s)
    case K_WBFM: return HaloOf<K_WBFM>::value;
    default: return 1;
    }
}

template <int KIND, int ENTRY>
int resident_warps()
{
    int blocks = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, rx_kernel<KIND, ENTRY>, HRD_WARPS_PER_CTA * 32,
                                                      sizeof(typename SmemOf<KIND>::type) * HRD_WARPS_PER_CTA) != cudaSuccess || blocks < 1)
        blocks = 1;
    return blocks * HRD_WARPS_PER_CTA;
}

// (stream, tile) items one SM works on at a time
int rx_resident_warps_per_sm(int kind, int entry)
{
    static int cache[5][2] = {};
    if (kind < 0 || kind > 4 || entry < 0 || entry > 1) return HRD_WARPS_PER_CTA;
    if (kind == K_WBFM) return WB_ITEMS; // one 32-warp CTA per SM
    if (cache[kind][entry]) return cache[kind][entry];
    int w = HRD_WARPS_PER_CTA;
#define HRD_RW(K, E) (w = resident_warps<K, E>())
    switch (kind) {
    case K_NONE: entry == 0 ? HRD_RW(K_NONE, 0) : HRD_RW(K_NONE, 1); break;
    case K_AM: entry == 0 ? HRD_RW(K_AM, 0) : HRD_RW(K_AM, 1); break;
    case K_FM: entry == 0 ? HRD_RW(K_FM, 0) : HRD_RW(K_FM, 1); break;
    }
#undef HRD_RW
    cache[kind][entry] = w;
    return w;
}

template <int ENTRY, bool TILED, bool SMALL>
int launch_wbfm_as(const RxParams &q, int grid, cudaStream_t s)
{
    static PerDeviceOnce optin; // per template instance
    const cudaError_t e = optin.run([] { return cudaFuncSetAttribute(rx_wbfm_kernel<ENTRY, TILED, SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemWbT<SMALL>)); });
    if (e != cudaSuccess) return (int)e;
    rx_wbfm_kernel<ENTRY, TILED, SMALL><<<grid, HRD_RX_WB_ROLES ? 1024 : (q.items_per_cta + 1) * 32, sizeof(SmemWbT<SMALL>), s>>>(q);
    return (int)cudaGetLastError();
}

template <int ENTRY>
int launch_wbfm(const RxParams &p, cudaStream_t s)
{
    const long long items = (long long)p.n_streams * p.n_tiles;
    RxParams q = p;
    // (the re-run's stream count is only known on the device: small CTAs, surplus ones exit at once)
    // (packing pays when it frees a good part of the SMs for the other kinds: measured, mixed batches of 4k-16k
    //  streams gain 1-3 %; when the full CTAs would cover > 85 % of the SMs anyway it only unbalances them: -2 %)
    const bool pack = p.wb_pack && (items + WB_ITEMS - 1) / WB_ITEMS * 100 <= (long long)p.sm_count * 85;
    q.items_per_cta = p.run_if ? WB_RERUN_ITEMS : pack ? WB_ITEMS : balanced_items_per_cta(items, p.sm_count, WB_ITEMS);
    const int grid = (int)((items + q.items_per_cta - 1) / q.items_per_cta);
    if (p.run_if) return p.n_tiles > 1 ? launch_wbfm_as<ENTRY, true, true>(q, grid, s) : launch_wbfm_as<ENTRY, false, true>(q, grid, s);
    return p.n_tiles > 1 ? launch_wbfm_as<ENTRY, true, false>(q, grid, s) : launch_wbfm_as<ENTRY, false, false>(q, grid, s);
}

// kind: K_NONE, K_AM (AM and SSB streams together), K_FM or K_WBFM
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
    case K_WBFM:
        return entry == 0 ? launch_wbfm<0>(p, s) : launch_wbfm<1>(p, s);
    }
#undef HRD_RX_CASE
    return (int)cudaErrorInvalidValue;
}

int launch_rx_wbfm_verify(const RxParams &p, const uint32_t *n_if, int most, uint32_t *count, int32_t *rerun_ids, float *guess_out,
                          unsigned long long *fallbacks, unsigned long long *fallbacks2, int force, cudaStream_t s)
{
    // (the retry's verification: the list is the first pass's, p.rerun_ids, *n_if <= most long; p.n_streams of them were run)
    const int32_t *ids = n_if ? p.rerun_ids : p.stream_ids;
    rx_wbfm_verify_kernel<<<((n_if ? most : p.n_streams) + 127) / 128, 128, 0, s>>>(p.wb_verify, ids, p.gain, p.n_streams, p.n_tiles, n_if,
                                                                                   count, rerun_ids, guess_out, fallbacks, fallbacks2, force);
    return (int)cudaGetLastError();
}

int launch_rx_gate(const RxParams &p, const GateParams &g, cudaStream_t s)
{
    if (p.n_streams <= 0 || p.n256 == 0) return 0;
    const int grid = (p.n_streams + HRD_WARPS_PER_CTA - 1) / HRD_WARPS_PER_CTA;
    rx_gate_kernel<<<grid, HRD_WARPS_PER_CTA * 32, 0, s>>>(p, g);
    return (int)cudaGetLastError();
}

int launch_rx_dc_iir(const RxParams &p, cudaStream_t s)
{
    if (p.n_streams <= 0 || p.n256 == 0) return 0;
    const int grid = (p.n_streams + 31) / 32;
    static PerDeviceOnce optin;
    const cudaError_t e = optin.run([] { return cudaFuncSetAttribute(rx_dc_iir_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemIir)); });
    if (e != cudaSuccess) return (int)e;
    rx_dc_iir_kernel<<<grid, 64 + IIR_POST_THREADS, sizeof(SmemIir), s>>>(p);
    return (int)cudaGetLastError();
}

} // namespace hrd
