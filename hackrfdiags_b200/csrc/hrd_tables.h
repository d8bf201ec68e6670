// hrd_tables.h -- coefficient tables, per-stream state records and kernel parameter blocks
// shared by the host side (hrd_api.cu) and the kernels (hrd_rx.cu, hrd_tx.cu).
//
// The float tap values are the reference's filter designs (the data of the spec); they are
// quantised at library load with the reference's rule (int16_t)round(c*32768) evaluated in
// float (radioDiags/Filters/Int16/Decimator_int16.cc:56-66) -- including its overflow of
// 1.0 to -32768 for the SSB delay line.
#pragma once

// tx_wbfm_kernel, stages 6-8: 1 = both rails per instruction as fp16 pairs (hrd_tx.cu tail3_h2, proved equal to the
// integer form by tools/verify_tx_tail_h2.c), 0 = the integer form per rail (tail3).  The NCO table the host builds
// holds halves in the first case and int16 in the second.
#ifndef HRD_TW_H2
#define HRD_TW_H2 1
#endif

#include <cstddef>
#include <cstdint>
#include <mutex>

#include <cuda_runtime.h>

namespace hrd {

// ---------------------------------------------------------------- taps (constant memory)
struct ConstTables {
    // Front end, three 3-tap half-band /2 stages (IqDataProcessor.cc:8-27), as dp2a
    // operand pairs with DOUBLED taps so that the >>15 of the reference becomes ">>16",
    // i.e. "take bytes 2..3":  A = {2*q[1] (centre, even sample), 2*q[0] (odd sample)},
    // B = {0, 2*q[2]} applied to the previous word's odd sample.
    uint32_t fe_a[3], fe_b[3];
    // FM tuner /4, 32 taps (FmDemodulator.cc:17-51): pair i multiplies ring word 2j+i,
    // lo half = q[31-2i] (even sample), hi half = q[30-2i] (odd sample)
    uint32_t fm_tuner[16];
    // AM/SSB stage 1 /4, 8 taps (AmDemodulator.cc:14-24), same pairing: lo q[7-2i], hi q[6-2i]
    uint32_t am1[4];
    // int16-domain decimators, plain taps
    int32_t am2[12];        // AmDemodulator.cc:27-41    /4
    int32_t am3[16];        // AmDemodulator.cc:44-62    /2
    int32_t fm_post[12];    // FmDemodulator.cc:54-68    /4 (also WBFM post-demod 2)
    int32_t audio40[40];    // FmDemodulator.cc:71-113   /2 (also WBFM audio, Tx stage 1)
    int32_t sig40[40];      // signals/interpolateSignal.cc:30-72: the stand-alone interpolator's own stage 1
    int32_t wbfm_post1[8];  // WbFmDemodulator.cc:17-27  /4
    // The same taps for mac_pair (hrd_rx.cu): word w serves ring samples 2w (low half) and 2w+1, i.e.
    // taps q[N-1-2w] and q[N-2-2w], each split q = th*256 + tl and packed {tl0, tl1, th0, th1}
    uint32_t wb1_sp[4], fm_post_sp[6], audio40_sp[20], am2_sp[6], am3_sp[8];
    int32_t hilbert[31];    // SsbDemodulator.cc:68-101, SsbModulator.cc:129-162
    int32_t delay[16];      // SsbDemodulator.cc:65: {0,...,0,-32768}
    // Tx half-band interpolators (AmModulator.cc:57-123)
    int32_t tx_hb8[8];      // stages 2,4,5
    int32_t tx_c3, tx_c7, tx_c8; // outer tap of stages 3&6 / 7 / 8 (8424 / 8249 / 8206)
    int32_t tx_m3, tx_m7, tx_m8; // their centre taps (16384 each after quantisation)
    int32_t k_32768;             // 32768 as a value ptxas cannot see (hrd_tx.cu PIPE BALANCE)
    // the NCO chains (hrd_device.cuh phase_step_fast): sign mask, -2PI_HI and -2PI_LO as REGISTER operands, so that
    // "(step & mask) ^ constant" is one LOP3 (with immediates ptxas needs two)
    uint32_t k_sign, k_m2pi_hi, k_m2pi_lo;
    // DbfsCalculator::DbfsCalculator (DbfsCalculator.cc:56-65): (int32_t)(20 * log10((float)i)), entry 0 = entry 1
    int32_t db_table[257];
};

// ---------------------------------------------------------------- Rx per-stream state
// Everything a reference IqDataProcessor + its four demodulators remember between calls,
// as filter HISTORIES (the last N-M inputs of every decimator) instead of ring buffers
// and indices.  32-bit slots; int16 samples sit in the low half, I/Q pairs are packed
// {I = low 16, Q = high 16}; 256 kS/s int8 samples are packed two per word as
// {I[2p], I[2p+1], Q[2p], Q[2p+1]} (the layout dp2a wants).
// AmDemodulator and SsbDemodulator share the /4 /4 /2 decimator chain and the DC-removal IIR
struct RxDec32 {
    uint32_t r256[2];      // 4 samples @256k
    uint32_t d64[8];       // 8 I/Q pairs @64k
    uint32_t a16[14];      // 14 I/Q pairs @16k
    float x1, y1;          // DC-removal IIR: x[n-1], y[n-1]
};

struct RxState {
    // IqDataProcessor stage 1/2/3 decimators: the last word each stage saw
    uint32_t fe_t, fe_v, fe_u, fe_pad;
    // AmDemodulator
    RxDec32 am;
    // FmDemodulator
    uint32_t fm_r256[14];  // 28 samples @256k
    float fm_theta[4];     // differentiator pipeline: theta[n-1..n-4], oldest first
    int16_t fm_d64[8];     // demodulated int16 @64k
    int16_t fm_a16[38];    // @16k
    // WbFmDemodulator
    float wb_prev_theta;   // previousTheta
    float wb_x1, wb_y1;    // de-emphasis IIR: x[n-1], y[n-1]
    uint32_t wb_pad;
    int16_t wb_d256[4];    // int16 @256k
    int16_t wb_d64[8];
    int16_t wb_a16[38];
    // SsbDemodulator
    RxDec32 ssb;
    uint32_t ssb_d8[30];   // I/Q pairs @8k for the delay line / Hilbert FIR
};

// ---------------------------------------------------------------- Tx per-stream state
struct TxRail8 {           // histories of the eight interpolators of one I/Q pair
    uint32_t s0[19];       // stage 1 input history @8k   (I/Q pairs)
    uint32_t s1[3];        // stage 2 input history @16k
    uint32_t s2[1];        // stage 3 @32k
    uint32_t s3[3];        // stage 4 @64k
    uint32_t s4[3];        // stage 5 @128k
    uint32_t s5[1];        // stage 6 @256k
    uint32_t s6[1];        // stage 7 @512k
    uint32_t s7[1];        // stage 8 @1024k
};

struct TxState {
    TxRail8 am, fm, ssb;
    // WbFmModulator: stages 1-5 on the real PCM (low halves used), 6-8 on I/Q.  tx_wbfm_kernel keeps s0 (PCM),
    // s1 (16 kS/s), s3 = the last THREE 32 kS/s samples (stages 3-5 are recomputed from them, so s2 and s4 stay
    // unused) and s5[0] = the last (cos, sin) * 900 pair.
    TxRail8 wb;
    float fm_phase;        // PhaseAccumulator::phaseAccumulator of the FM NCO (8 kS/s)
    float wb_phase;        // ... of the WBFM NCO (256 kS/s)
    uint32_t ssb_h8[30];   // PCM/2 history for the delay line / Hilbert FIR
    uint32_t pad[2];
    TxRail8 sig;           // signals/interpolateSignal.cc: its I and Q interpolator trees
    float sig_theta;       // signals/fm.cc: theta
    uint32_t pad2[3];
};

// ---------------------------------------------------------------- kernel parameters
struct RxParams {
    const int8_t *iq;
    size_t iq_stride;          // bytes between streams
    uint32_t n256;             // 256 kS/s samples per stream in this call (multiple of 32)
    int16_t *pcm;
    size_t pcm_stride;         // samples between streams
    int8_t *out256;            // front-end-only output (may be null)
    size_t out_stride;
    // Per-stream state is double-buffered: a call reads state_in and writes state_out, so the
    // warp that finishes a stream's LAST time tile can never overwrite what the warp of its
    // FIRST tile has yet to read.  The host swaps the two after every call.
    const RxState *state_in;
    RxState *state_out;
    const int32_t *stream_ids; // streams of this launch (one mode per launch)
    int32_t n_streams;
    const float *gain;         // [n_streams_total] gain of this launch's demodulator
    const uint8_t *lsb;        // [n_streams_total] SSB sideband flag
    const uint8_t *kind_of;    // [n_streams_total] K_* of every stream (the AM launch also runs SSB streams)
    const float *gain_ssb;     // [n_streams_total] SSB gain (AM/SSB launch: gain is the AM one)
    const float *atan2_lut;    // [256*256]
    // Time tiling: every stream's call is cut into n_tiles tiles of
    // tile_batches batches (1 batch = 1024 samples at 256 kS/s = 32 PCM samples); one warp owns
    // one (stream, tile).  Tile 0 starts from the saved state, later tiles rebuild the FIR
    // histories by running HALO batches ahead of their first output.
    int32_t n_tiles;
    uint32_t tile_batches;
    int32_t sm_count;          // multiprocessors of the device (grid sizing)
    // AM / SSB: the FIR half of the DC-removal filter, fir[n] = x[n] - x[n-1] as floats, one per
    // PCM sample.  The recurrence itself is serial per stream and runs in rx_dc_iir_kernel
    // afterwards.
    float *pre_iir;
    size_t pre_stride;         // elements between streams (multiple of 8)
    // rx_wbfm_kernel: (stream, tile) items per CTA, <= 31; chosen by the launcher so that the grid
    // fills whole waves of SMs (4096 items: 147 CTAs of 28 instead of 133 of 31)
    int32_t items_per_cta;
    // rx_wbfm_kernel, time-tiled calls: one {speculated, true} pair of de-emphasis outputs per
    // (stream slot, tile); see "verified speculation" in hrd_rx.cu.  run_if: when non-null the
    // kernel runs only if *run_if != 0 (the exact untiled re-run after a failed verification).
    float2 *wb_verify;
    const uint32_t *run_if;    // [0] = number of streams to re-run
    const int32_t *rerun_ids;  // their stream ids (replaces stream_ids / n_streams in the re-run)
    // the TILED retry of those streams (hrd_rx.cu "second pass"): per listed stream the value every tile >= 1 puts
    // into the recurrence at its check point instead of the warmed-up one; null otherwise
    const float *wb_guess;
    // rx_wbfm_kernel beside other kinds' kernels (a mixed batch fanned out over streams): 1 = full CTAs on as many SMs
    // as that takes instead of every SM with a partly filled one -- a CTA's step time falls more slowly than its item
    // count (21 items: 3.53 us, 27: 3.92 us), and the SMs left over are not idle, the other kinds run there
    int32_t wb_pack;
    // Ragged calls (the squelched path: every stream demodulates only the blocks its gate let through): when
    // non-null, stream sid has n256_of[sid] <= n256 samples in its row; rx_kernel tiles such calls over the nominal length, rx_wbfm_kernel runs them with n_tiles == 1.
    const uint32_t *n256_of;
};

// The squelch gate, fused into the front end (hrd_rx.cu rx_gate_kernel): per stream and per block of blk256
// samples at 256 kS/s the average magnitude, the tracker's decision, and the 256 kS/s samples of the OPEN blocks
// packed one behind the other in the stream's scratch row -- what the demodulators then read as one ragged call.
struct GateParams {
    int8_t *scratch;           // [n_streams][row256] int8 I,Q at 256 kS/s, open blocks only
    size_t row256;             // bytes between rows
    uint32_t blk256;           // samples per block (a multiple of 32); the last block of a call may be short
    int32_t n_blocks;
    const float *threshold;    // per stream: IqDataProcessor::setSignalDetectThreshold (an int32 in the reference)
    const float *gain_db;      // per stream: radio_adjustableReceiveGainInDb (a uint32)
    uint8_t *tracking;         // per stream: SignalTracker state (1 = Tracking), carried across calls
    uint32_t *magnitude;       // [n_streams][n_blocks] Squelch::getSignalMagnitude()
    uint8_t *allowed;          // [n_streams][n_blocks] Squelch::run()'s result
    uint32_t *n256_open;       // [n_streams] samples written to the scratch row (0 for streams without a demodulator)
};

struct TxParams {
    const int16_t *pcm;
    size_t pcm_stride;
    uint32_t n8;               // PCM samples per stream in this call
    int8_t *iq;
    size_t iq_stride;
    // Per-stream state is double-buffered like the receive side's: a call reads `state` and writes
    // `state_out` (the host copies the records over first, so kernels only write what they own).
    const TxState *state;
    TxState *state_out;
    // Time tiles (AM / FM / SSB; hrd_tx.cu): every stream's call is cut into n_tiles tiles of
    // tile_len8 PCM samples (a multiple of 32); one warp owns one (stream, tile).
    int32_t n_tiles;
    uint32_t tile_len8;
    // FM: NCO phase BEFORE each PCM sample's step, [slot of this launch][n8] (tx_fm_phase_kernel)
    float *fm_phase;
    const int32_t *stream_ids;
    int32_t n_streams;
    const int32_t *mode_of;    // [n_streams_total] HRD_MODE_* of every stream (the signals/ launch picks its head by it)
    const float *param;        // AM index / FM deviation / WBFM deviation
    const uint8_t *lsb;
    const float *nco_sin, *nco_cos;
    // WBFM: {(int16_t)(cos*900), (int16_t)(sin*900)} per NCO table entry, packed I = low half
    // (WbFmModulator.cc:606-626 applied to Nco.cc's tables once, at table build), FOLDED about phase 0:
    // [k] = entry min(8192 + k, 16383), [8193 + k] = entry 8192 - k, k = 0..8192 (hrd_tx.cu nco_fold_offset).
    // With HRD_TW_H2 the two values are binary16 numbers (exact: |v| <= 900) instead of int16.
    const uint32_t *nco_iq900;
    const float *nco_thr;      // [8194] Nco::runFast index thresholds (hrd_tx.cu nco_index)
    int32_t items_per_cta;     // tx_wbfm_kernel: streams per CTA, <= 31 (see RxParams)
    int32_t sm_count;
};

enum { K_NONE = 0, K_AM = 1, K_FM = 2, K_WBFM = 3, K_SSB = 4,
       K_IQ = 5 /* Tx only: the tool chain of signals/ (heads + interpolateSignal) */, K_COUNT = 6 };
// HRD_MODE_* values of the signals/ heads (include/hrd.h), as the kernel sees them
enum { SIG_MODE_IQ8K = 6, SIG_MODE_DSB = 7, SIG_MODE_PM = 8, SIG_MODE_AM_PROTO = 9, SIG_MODE_FM_PROTO = 10 };

// Items per CTA for the chain-warp kernels (at most `cap`): the smallest count that still needs no
// more waves of CTAs than `cap` per CTA would, so that the last wave is as full as the others.
static inline int balanced_items_per_cta(long long items, int sms, int cap)
{
    if (items <= 0 || sms <= 0) return cap;
    const long long waves = (items + (long long)cap * sms - 1) / ((long long)cap * sms);
    long long ipc = (items + waves * sms - 1) / (waves * sms);
    if (ipc < 1) ipc = 1;
    if (ipc > cap) ipc = cap;
    return (int)ipc;
}

// cudaFuncSetAttribute (the opt-in to more than 48 KB of dynamic shared memory) is per DEVICE: one of these per
// launch site remembers which devices have had it, so a process that opens batches on several GPUs gets it on each.
struct PerDeviceOnce {
    std::mutex m;
    uint64_t done = 0; // bit d: device d is set up (hrd_create admits devices 0..63)
    template <class Fn> cudaError_t run(Fn fn)
    {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        std::lock_guard<std::mutex> lock(m);
        if (dev >= 0 && dev < 64 && (done >> dev & 1)) return cudaSuccess;
        e = fn();
        if (e == cudaSuccess && dev >= 0 && dev < 64) done |= 1ull << dev;
        return e;
    }
};

void upload_tables(const ConstTables &t);            // hrd_rx.cu (owns the __constant__ copy)
void upload_tables_tx(const ConstTables &t);         // hrd_tx.cu
int launch_rx(int kind, int entry, const RxParams &p, cudaStream_t s);
int launch_rx_dc_iir(const RxParams &p, cudaStream_t s);  // AM + SSB streams, after their launch_rx
// WBFM, tiled call: compare every tile's speculated recurrence value with the true one; streams with any
// difference are appended to rerun_ids (*count of them) and added to *fallbacks (or always, when force is set: test hook)
// n_if: when non-null this is the verification of the tiled RETRY: the streams are the first pass's list p.rerun_ids,
// *n_if of them (at most `most`), of which the retry ran the first p.n_streams -- the others fail without a look.
// guess_out: when non-null, receives per failing stream the TRUE value at the first check point (the retry's wb_guess).
// fallbacks2: a second counter to add the failures to, or null.
int launch_rx_wbfm_verify(const RxParams &p, const uint32_t *n_if, int most, uint32_t *count, int32_t *rerun_ids, float *guess_out,
                          unsigned long long *fallbacks, unsigned long long *fallbacks2, int force, cudaStream_t s);
int rx_halo_batches(int kind);                            // batches a tile > 0 runs ahead
int rx_resident_warps_per_sm(int kind, int entry);        // occupancy of that kernel (cached)
int launch_rx_gate(const RxParams &p, const GateParams &g, cudaStream_t s); // hrd_rx.cu: front end + squelch gate, every stream
int launch_fs4_rotate(int8_t *iq, size_t n_groups, int up, cudaStream_t s);
int launch_tx(int kind, const TxParams &p, cudaStream_t s);
int launch_tx_fm_phase(const TxParams &p, cudaStream_t s);   // FM streams, before their launch_tx
int launch_tx_sig_phase(const TxParams &p, cudaStream_t s);  // signals/fm.cc streams of a K_IQ launch, before it
int tx_resident_warps_per_sm(int kind);                      // occupancy of tx_kernel<kind> (cached)
int tx_halo_samples(int kind);                               // PCM samples a tile > 0 runs ahead

} // namespace hrd
