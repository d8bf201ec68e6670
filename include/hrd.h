/*
 * hrd.h -- C ABI of libhrd_b200.so: the HackRfDiags baseband DSP hot path on
 * NVIDIA B200 (sm_100a), batched over thousands of independent streams.
 *
 * One "stream" is what one reference object graph processes:
 *   Rx: one IqDataProcessor plus its AmDemodulator, FmDemodulator,
 *       WbFmDemodulator and SsbDemodulator
 *       (radioDiags/hdr_diags/IqDataProcessor.h:21-39,
 *        radioDiags/{Am,Fm,WbFm,Ssb}Demodulator/ headers)
 *   Tx: one AmModulator, FmModulator, WbFmModulator and SsbModulator
 *       (radioDiags/{Am,Fm,WbFm,Ssb}Modulator/ headers) as owned by
 *       BasebandDataProcessor (radioDiags/hdr_diags/BasebandDataProcessor.h)
 * A batch owns n_streams of them; every stream keeps its own filter state,
 * mode and parameters between calls exactly like the reference objects do.
 *
 * The reference has no FFI; its boundary is those C++ classes.  The shim
 * classes in hackrfdiags_b200/shim/ re-create them (same names, signatures,
 * callback behaviour) on top of this ABI with n_streams = 1; INTEGRATION.md
 * shows how they drop into buildRadioDiags.sh.
 *
 * Conventions: every function returns 0 on success or a negative HRD_E*
 * code; hrd_last_error() gives the message (thread-local).  No exceptions
 * cross the boundary.  Calls on one batch must be serialised by the caller
 * (same rule as the reference objects: one data thread per object).  There
 * is NO CPU fallback: without a CUDA device hrd_create fails.
 */
#ifndef HRD_H
#define HRD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HRD_ABI_VERSION 1

typedef struct hrd_batch hrd_batch_t;

enum { HRD_RX = 0, HRD_TX = 1 };

/* IqDataProcessor::demodulatorType (IqDataProcessor.h:21); the same numbers
 * select the modulator on Tx (BasebandDataProcessor.h modulatorType). */
enum {
    HRD_MODE_NONE = 0,
    HRD_MODE_AM = 1,
    HRD_MODE_FM = 2,
    HRD_MODE_WBFM = 3,
    HRD_MODE_LSB = 4,
    HRD_MODE_USB = 5,
    /* Tx batches only: the stand-alone tool chain of signals/ (generateBaseband.sh: <head> < pcm | interpolateSignal),
     * i.e. signals/interpolateSignal.cc's eight-stage interpolator with ITS stage-1 taps (:30-72) behind ... */
    HRD_MODE_IQ8K = 6,     /* nothing: the input rows are int16 I,Q PAIRS at 8 kS/s (interpolateSignal.cc:266-340) */
    HRD_MODE_DSB = 7,      /* signals/dsb.cc:38-47  x/4 on both rails                                            */
    HRD_MODE_PM = 8,       /* signals/pm.cc:39-55   (cos, sin)(x/60000*pi) * 16000 (libm's cosf / sinf, bit for bit) */
    HRD_MODE_AM_PROTO = 9, /* signals/am.cc:38-50   (x*0.8 + 65536)/4 on both rails                              */
    HRD_MODE_FM_PROTO = 10 /* signals/fm.cc:41-62   theta += x/65536*3.5 wrapped at +-2*pi, (cos, sin)*16000 (libm)   */
};

/* per-stream scalar parameters and the reference setter each one replaces */
enum {
    HRD_PARAM_AM_GAIN = 0,   /* AmDemodulator::setDemodulatorGain   (AmDemodulator.cc:267)   default 300        */
    HRD_PARAM_FM_GAIN = 1,   /* FmDemodulator::setDemodulatorGain   (FmDemodulator.cc:323)   default 64000/2pi  */
    HRD_PARAM_WBFM_GAIN = 2, /* WbFmDemodulator::setDemodulatorGain (WbFmDemodulator.cc:299) default 256000/2pi */
    HRD_PARAM_SSB_GAIN = 3,  /* SsbDemodulator::setDemodulatorGain  (SsbDemodulator.cc:390)  default 300        */
    HRD_PARAM_AM_INDEX = 4,  /* AmModulator::setModulationIndex     (AmModulator.cc:329)  accepted iff 0<=m<=1, default 0.8 */
    HRD_PARAM_FM_DEV = 5,    /* FmModulator::setFrequencyDeviation  (FmModulator.cc:336)  guard tests the OLD value vs 0..3500   */
    HRD_PARAM_WBFM_DEV = 6,  /* WbFmModulator::setFrequencyDeviation(WbFmModulator.cc:310) guard tests the OLD value vs 0..112000 */
    /* the squelch gate of IqDataProcessor::acceptIqData (2.048 MS/s entry only; see hrd_rx_process) */
    HRD_PARAM_SQUELCH_THRESHOLD = 7, /* IqDataProcessor::setSignalDetectThreshold (IqDataProcessor.cc:392-405), dBFS as an
                                        integer value, default -200 = always open (IqDataProcessor.cc:120-124) */
    HRD_PARAM_RX_GAIN_DB = 8,        /* radio_adjustableReceiveGainInDb, the gain Squelch::run refers the level to
                                        (IqDataProcessor.cc:961; Radio.cc:413 default 16) */
    HRD_PARAM_COUNT = 9
};

/* which object hrd_reset addresses */
enum {
    HRD_UNIT_AM = 0,   /* AmDemodulator::resetDemodulator / AmModulator::resetModulator     */
    HRD_UNIT_FM = 1,   /* Fm...   (modulator: filters only, the NCO phase is kept)          */
    HRD_UNIT_WBFM = 2, /* WbFm... (demodulator: the de-emphasis filter is NOT reset)        */
    HRD_UNIT_SSB = 3,  /* Ssb...                                                            */
    HRD_UNIT_FRONT_END = 4, /* (ours) IqDataProcessor's three half-band decimators          */
    HRD_UNIT_ALL = 5,       /* (ours) everything, i.e. freshly constructed objects          */
    HRD_UNIT_SIGNALS = 6    /* (ours) Tx: the interpolator trees of the signals/ tool chain (a fresh process) */
};

enum { HRD_ENTRY_2048K = 0, /* IqDataProcessor::acceptIqData  (IqDataProcessor.cc:926)  */
       HRD_ENTRY_256K = 1   /* <X>Demodulator::acceptIqData   (e.g. FmDemodulator.cc:353) */ };

enum { HRD_MEM_HOST = 0, HRD_MEM_DEVICE = 1 };

/* Execution options (ours; they change how a call is scheduled on the GPU, never its result,
 * with the one documented exception).  The reference has no counterpart: its objects are
 * single-stream and single-threaded. */
enum {
    /* Rx: batches (of 8192 input samples = 32 PCM samples) per time tile; one warp owns one
     * (stream, tile).  0 = choose from the stream count and the SM count. */
    HRD_OPT_RX_TILE_BATCHES = 0,
    /* Rx WBFM: time tiling by verified speculation.  1 (default): tiles after the first warm the
     * 256 kS/s de-emphasis recurrence up from zero, every tile's warmed-up value is compared bit
     * for bit with the true one, and streams where any differs are run again (hrd_wbfm_fallback_count)
     * -- the result is bit-exact either way.
     * 0: never tile WBFM calls. */
    HRD_OPT_RX_WBFM_TILING = 1,
    /* Tx: PCM samples per time tile (multiple of 32).  0 = choose automatically. */
    HRD_OPT_TX_TILE_SAMPLES = 2,
    /* record CUDA events around the kernels of every process call (bench.py's roofline):
     * hrd_kernel_ms() then reports the main and tail kernel times of the latest calls */
    HRD_OPT_PROFILE = 3,
    /* test hook: 1 = make the WBFM verification fail, so that the retry (and, where its guess is wrong, the serial
     * re-run) is exercised; 2 = make the retry's verification fail as well */
    HRD_OPT_DEBUG_WBFM_FORCE_RERUN = 4,
    /* Rx, 2.048 MS/s entry: 1 = take the squelched path (per-block magnitudes and decisions, see
     * hrd_rx_squelch_report) even when no stream's threshold can close the gate.  The path is taken
     * automatically as soon as one stream's threshold can (threshold > -42 - gain dB). */
    HRD_OPT_RX_SQUELCH = 5,
    /* Rx, 2.048 MS/s entry: input bytes per squelch decision, i.e. the size of the reference call being
     * modelled (a multiple of 512).  0 (default) = 262144, the HackRF transfer block. */
    HRD_OPT_RX_SQUELCH_BLOCK = 6,
    /* Rx, measurement only: 1 = the kernels of a mixed-mode batch run one kind after the other on the caller's
     * stream instead of side by side, and (with HRD_OPT_PROFILE) each kind's own span is timed: hrd_kernel_ms
     * which = 10 + kind (1 AM+SSB, 2 FM, 3 WBFM).  Results are identical either way. */
    HRD_OPT_RX_SERIAL = 7,
    /* Rx, mixed-mode batches: 1 (default) = the WBFM launch fills whole CTAs (one per SM, 27 streams each) on as many
     * SMs as that takes, leaving the others to the AM / NBFM kernels running beside it; 0 = it spreads over all SMs
     * as it does when it runs alone.  Results are identical either way.  (Measured: +3 % on the mixed 4096-stream batch
     * in a process of its own; -3 % inside a process that holds a multi-rank NCCL communicator -- turn it off there.) */
    HRD_OPT_RX_WBFM_PACK = 8,
    HRD_OPT_COUNT = 9
};

#define HRD_ALL_STREAMS (-1)

enum {
    HRD_OK = 0,
    HRD_EINVAL = -1,   /* bad argument (size not a whole number of PCM samples, bad enum, ...) */
    HRD_ENODEV = -2,   /* no usable CUDA device / wrong architecture */
    HRD_ECUDA = -3,    /* CUDA runtime error, see hrd_last_error()   */
    HRD_ENOMEM = -4
};

/* ---- lifetime ------------------------------------------------------- */
int hrd_abi_version(void);
int hrd_create(int device, int n_streams, int kind, hrd_batch_t **out);
int hrd_destroy(hrd_batch_t *b);
const char *hrd_last_error(void);

/* ---- control (replaces the reference's unsynchronised member writes) - */
/* Setters may be called from another thread than the one that makes the process calls (as Radio.cc:1973,
 * 2404-2633 does from the UI thread): they take a per-batch lock, and the next process call uploads a consistent
 * snapshot on its own stream.  Nothing here synchronises the device. */
/* IqDataProcessor::setDemodulatorMode (IqDataProcessor.cc:346-372; LSB/USB
 * also flip the SSB demodulator's sideband) or, on a Tx batch,
 * BasebandDataProcessor::setModulatorMode (+ Ssb set{Lsb,Usb}ModulationMode). */
int hrd_set_mode(hrd_batch_t *b, int stream, int mode);
int hrd_get_mode(hrd_batch_t *b, int stream, int *mode);
int hrd_set_param(hrd_batch_t *b, int stream, int param, float value);
int hrd_get_param(hrd_batch_t *b, int stream, int param, float *value);
int hrd_reset(hrd_batch_t *b, int stream, int unit);
int hrd_set_option(hrd_batch_t *b, int option, int value);
int hrd_get_option(hrd_batch_t *b, int option, int *value);

/* ---- receive --------------------------------------------------------- */
/*
 * One call = one reference call on every stream:
 *   entry 2048K: IqDataProcessor::acceptIqData(ts, iq, bytes_per_stream)
 *   entry 256K : <mode's demodulator>::acceptIqData(iq, bytes_per_stream)
 * iq        int8 interleaved I,Q; stream s starts at iq + s*iq_stride bytes
 *           (HRD_MEM_DEVICE: pointer and stride 32-byte aligned at the 2048K entry, 4-byte aligned at the 256K one)
 * bytes_per_stream  multiple of 512 (2048K) or 64 (256K): whole PCM samples.
 *           Unlike the reference there is no 262144 / 32768 byte ceiling; one call takes less than 4 GiB per stream.
 * pcm       int16 out; stream s starts at pcm + s*pcm_stride samples; gets
 *           bytes_per_stream/512 (resp. /64) samples, 0 when the mode is NONE
 * pcm_counts  optional, host memory, n_streams entries
 *
 * Squelch (entry 2048K; Squelch.cc:227-273, SignalDetector.cc:205-273, SignalTracker.cc:104-145,
 * DbfsCalculator.cc): the reference takes one squelch decision per acceptIqData call, i.e. per
 * 262144-byte transfer block (hackRf/hackrf.c:101).  A call here is cut into blocks of 262144 bytes (a
 * shorter last block counts as its own call); a block the gate closes is not demodulated: the
 * demodulator state does not move and the stream's PCM is shorter by that block (pcm_counts tells).
 * With every threshold at its default (-200 dBFS) the gate cannot close and the call runs fused as one
 * pass; otherwise one kernel runs the front end, the magnitudes and the tracker of every stream and packs each
 * stream's open blocks, and the demodulators take those as one ragged call: still nothing waits for the host.
 * With HRD_MEM_DEVICE pcm_counts (and hrd_rx_squelch_report) are then filled by a host function on cuda_stream:
 * read them after synchronising that stream; the array must stay valid until then.
 * mem       HRD_MEM_HOST: pointers are host memory; the copies are cudaMemcpy2DAsync on the batch's own
 *           stream straight from / to the caller's buffers and the call returns when pcm is ready.  Pinned
 *           buffers (cudaHostAlloc / cudaHostRegister) get the full PCIe rate; pageable ones work but are
 *           staged by the driver and slower.
 *           HRD_MEM_DEVICE: device pointers, work is queued on cuda_stream
 *           and the call returns without synchronising
 * cuda_stream  a cudaStream_t (NULL = the legacy default stream).  Calls on one batch may name different
 *           streams: each call is ordered behind the one before it by an event.
 */
int hrd_rx_process(hrd_batch_t *b, const int8_t *iq, size_t bytes_per_stream, size_t iq_stride,
                   int entry, int16_t *pcm, size_t pcm_stride, uint32_t *pcm_counts, int mem,
                   void *cuda_stream);

/* IqDataProcessor::reduceSampleRate + upconvertByFsOver4 only
 * (IqDataProcessor.cc:429-500, 771-815): the 256 kS/s int8 I,Q stream the
 * reference can dump over UDP.  out256k gets bytes_per_stream/8 bytes per
 * stream at out_stride.  Advances the front-end state like the real call. */
int hrd_rx_front_end(hrd_batch_t *b, const int8_t *iq, size_t bytes_per_stream, size_t iq_stride,
                     int8_t *out256k, size_t out_stride, int mem, void *cuda_stream);

/* IqDataProcessor::upconvertByFsOver4 (up != 0, IqDataProcessor.cc:771-815) or downconvertByFsOver4 (up == 0,
 * :715-759) on their own, in place on bytes (a multiple of 8) of int8 I,Q: the rotation hrd_rx_process fuses into
 * its front end, for callers that use the two public methods separately. */
int hrd_rx_fs4_rotate(hrd_batch_t *b, int8_t *iq, size_t bytes, int up, int mem, void *cuda_stream);
/* What the reference reports through registerSignalMagnitudeCallback / registerSignalStateCallback
 * (IqDataProcessor.cc:961-988), for every stream and block of the latest squelched hrd_rx_process call:
 * magnitudes[s * blocks_cap + b] = Squelch::getSignalMagnitude(), allowed[...] = Squelch::run()'s result.
 * Either array may be NULL.  *n_blocks gets the number of blocks of that call (0 if it was not squelched). */
int hrd_rx_squelch_report(hrd_batch_t *b, uint32_t *magnitudes, uint8_t *allowed, size_t blocks_cap, uint32_t *n_blocks);
/* ---- transmit -------------------------------------------------------- */
/*
 * One call = <mode's modulator>::acceptData(pcm, n_per_stream, iq, &len) on
 * every stream (e.g. AmModulator.cc:366-381).  pcm is int16 at 8 kS/s,
 * stream s at pcm + s*pcm_stride samples; iq gets n_per_stream*512 bytes of
 * int8 I,Q at 2.048 MS/s at iq + s*iq_stride bytes (32-byte aligned starts).
 * Mode NONE writes the idle carrier BasebandDataProcessor uses: every byte
 * 64 (BasebandDataProcessor.cc:689-694).  No 512-sample ceiling.
 * A stream in mode HRD_MODE_IQ8K reads n_per_stream int16 I,Q PAIRS from its row (4-byte aligned, so
 * pcm_stride must be even and at least 2*n_per_stream when the batch holds one).
 */
int hrd_tx_process(hrd_batch_t *b, const int16_t *pcm, size_t n_per_stream, size_t pcm_stride,
                   int8_t *iq, size_t iq_stride, int mem, void *cuda_stream);

/* ---- ingest / egress adapters (host side; hrd_adapt.cc) ---------------- */
/*
 * The reference puts a 16-slot block pool + queue between the libusb thread and IqDataProcessor (DataConsumer,
 * src_diags/DataConsumer.cc:219-261, 319-351) and a 16-block PCM ring with drop / repeat rate matching between
 * the stdin reader thread and the transmit callback (BasebandDataProcessor.cc:416-433, 482-605).  These are the
 * same structures for MANY streams: producers push per stream, the consumer takes one block of EVERY stream per
 * round -- the row matrix one hrd_rx_process / hrd_tx_process call reads.  One producer and one consumer thread
 * may run concurrently, as in the reference.
 */
typedef struct hrd_pcm_ring hrd_pcm_ring_t;
typedef struct hrd_iq_queue hrd_iq_queue_t;
int hrd_pcm_ring_create(int n_streams, hrd_pcm_ring_t **out);
int hrd_pcm_ring_destroy(hrd_pcm_ring_t *r);
/* BasebandDataProcessor::start / stop: an idle stream sends zero blocks (BasebandDataProcessor.cc:588-601) */
int hrd_pcm_ring_start(hrd_pcm_ring_t *r, int stream, int running);
/* the reader thread's step: getNextUnfilledBuffer + fread of up to 512 samples (:416-433, :862-865) */
int hrd_pcm_ring_write(hrd_pcm_ring_t *r, int stream, const int16_t *pcm, uint32_t n_samples);
/* the transmit callback's getNextFilledBuffer (:482-605) for every stream: rows[s*row_stride .. +512); slots
 * (optional) gets the ring slot each stream sent, -1 for the zero block */
int hrd_pcm_ring_read_all(hrd_pcm_ring_t *r, int16_t *rows, size_t row_stride, int32_t *slots);
/* buffersProduced, buffersConsumed, pcmBlocksDropped, pcmBlocksAdded */
int hrd_pcm_ring_stats(hrd_pcm_ring_t *r, int stream, uint32_t out[4]);
/* BasebandDataProcessor::getIqData for every stream of a Tx batch: one ring block each -> 262144 bytes each */
int hrd_tx_from_ring(hrd_batch_t *b, hrd_pcm_ring_t *r, int8_t *iq, size_t iq_stride, int mem, void *cuda_stream);
int hrd_iq_queue_create(int n_streams, hrd_iq_queue_t **out);
int hrd_iq_queue_destroy(hrd_iq_queue_t *q);
/* DataConsumer::acceptData (:219-261): at most 262144 bytes into the stream's next pool slot, queued */
int hrd_iq_queue_push(hrd_iq_queue_t *q, int stream, uint32_t time_stamp, const void *data, uint32_t bytes);
/* the same for `count` consecutive streams in one call: row i (rows + i * row_stride) is the block of stream first + i */
int hrd_iq_queue_push_rows(hrd_iq_queue_t *q, int first, int count, uint32_t time_stamp, const void *rows, size_t row_stride,
                           uint32_t bytes);
/* one round: returns 1 with one block of every stream in rows (bytes / time_stamps optional), 0 if a stream has none */
int hrd_iq_queue_pop_all(hrd_iq_queue_t *q, int8_t *rows, size_t row_stride, uint32_t *bytes, uint32_t *time_stamps);
/* blocks queued, shortBlockCount, lastTimeStamp */
int hrd_iq_queue_stats(hrd_iq_queue_t *q, int stream, uint32_t out[3]);
/* the consumer thread's step for every stream: one round through hrd_rx_process at the 2.048 MS/s entry
 * (host pcm buffers); 1 = a round was processed, 0 = nothing to do */
int hrd_rx_from_queue(hrd_batch_t *b, hrd_iq_queue_t *q, int16_t *pcm, size_t pcm_stride, uint32_t *pcm_counts);

/* Pipelined rounds (the adapters at rate).  The block pool of hrd_iq_queue_* is page-locked and laid out
 * [slot][stream][262144], so a round whose streams sit at the same slot (producers in step) is ONE asynchronous
 * host-to-device copy straight from the pool.  A pipe keeps up to `depth` rounds in flight, each on a CUDA stream of
 * its own: the copy of round k+1 runs beside the kernels of round k and the PCM copy of round k-1.  One consumer
 * thread drives a pipe (submit / collect), producers keep pushing concurrently; depth <= 8 (half the pool: a slot
 * is not reused while its copy may still be reading it, provided the consumer keeps up as the reference requires). */
typedef struct hrd_rx_pipe hrd_rx_pipe_t;
typedef struct hrd_tx_pipe hrd_tx_pipe_t;
int hrd_rx_pipe_create(hrd_batch_t *b, hrd_iq_queue_t *q, int depth, hrd_rx_pipe_t **out);
int hrd_rx_pipe_destroy(hrd_rx_pipe_t *p);
/* 1 = a round was started (every stream had a block queued and the pipe had room), 0 = not now; HRD_EINVAL when the
 * blocks at the heads of the queues differ in size or are not whole PCM samples (nothing is dequeued then) */
int hrd_rx_pipe_submit(hrd_rx_pipe_t *p);
/* waits for the OLDEST round in flight: 1 and its PCM (pinned rows of 512 samples, valid until depth - 1 more
 * rounds have been submitted) and per-stream sample counts; 0 = no round in flight */
int hrd_rx_pipe_collect(hrd_rx_pipe_t *p, const int16_t **pcm, size_t *pcm_stride, const uint32_t **pcm_counts);
/* rounds started, host-to-device copies issued for them */
int hrd_rx_pipe_stats(hrd_rx_pipe_t *p, uint64_t out[2]);
/* the transmit side: one getIqData of every stream per round (ring policy, modulators, 262144 bytes per stream
 * back into pinned rows) */
int hrd_tx_pipe_create(hrd_batch_t *b, hrd_pcm_ring_t *r, int depth, hrd_tx_pipe_t **out);
int hrd_tx_pipe_destroy(hrd_tx_pipe_t *p);
int hrd_tx_pipe_submit(hrd_tx_pipe_t *p);
int hrd_tx_pipe_collect(hrd_tx_pipe_t *p, const int8_t **iq, size_t *iq_stride);

/* ---- one job on several GPUs (host side; hrd_shard.cc) ----------------- */
/*
 * Streams are independent, so N streams on G GPUs are G disjoint batches: shard g owns streams [g*N/G, (g+1)*N/G)
 * on devices[g] (a device may be listed twice: two shards on it).  One worker thread per shard; the process calls
 * take HOST rows of the whole job, hand every shard its rows at the same time and return when all are done.  Nothing
 * crosses GPUs.  Stream numbers are the job's.  hrd_sharded_shard gives a shard's device, range and batch handle
 * for anything else (device-memory calls on that GPU, options, introspection).
 */
typedef struct hrd_sharded hrd_sharded_t;
int hrd_sharded_create(const int *devices, int n_devices, int n_streams, int kind, hrd_sharded_t **out);
int hrd_sharded_destroy(hrd_sharded_t *s);
int hrd_sharded_count(hrd_sharded_t *s);
int hrd_sharded_shard(hrd_sharded_t *s, int shard, int *device, int *lo, int *hi, hrd_batch_t **batch);
const char *hrd_sharded_last_error(hrd_sharded_t *s); /* the failing shard's message (the workers' are thread-local) */
int hrd_sharded_set_mode(hrd_sharded_t *s, int stream, int mode);
int hrd_sharded_set_param(hrd_sharded_t *s, int stream, int param, float value);
int hrd_sharded_reset(hrd_sharded_t *s, int stream, int unit);
int hrd_sharded_set_option(hrd_sharded_t *s, int option, int value);
int hrd_sharded_rx_process(hrd_sharded_t *s, const int8_t *iq, size_t bytes_per_stream, size_t iq_stride, int entry,
                           int16_t *pcm, size_t pcm_stride, uint32_t *pcm_counts);
int hrd_sharded_tx_process(hrd_sharded_t *s, const int16_t *pcm, size_t n_per_stream, size_t pcm_stride, int8_t *iq,
                           size_t iq_stride);

/* ---- introspection (tests, bench) ------------------------------------ */
int hrd_synchronize(hrd_batch_t *b);
/* the CUDA device the batch lives on (hrd_create's argument) */
int hrd_get_device(hrd_batch_t *b, int *device);
/* with HRD_OPT_PROFILE set: device time of a recent process call's kernels; age 0 = the latest
 * call, up to 31 calls back; which = 0 the main kernels (Rx tile kernels / Tx kernels), 1 the
 * tail kernel (Rx AM/SSB IIR pass), 10 + kind = that kind's own kernels (HRD_OPT_RX_SERIAL).
 * Waits for that call to finish. */
int hrd_kernel_ms(hrd_batch_t *b, int which, int age, float *ms);
/* kernels this batch has launched so far (bench.py's gpu_launches) */
int hrd_launch_count(hrd_batch_t *b, uint64_t *count);
/* Rx WBFM (HRD_OPT_RX_WBFM_TILING): how many stream-calls failed the tile verification so far and were run again --
 * tiled once more from the true value at the first check point (constant inputs pass that), and serially from the
 * saved state when that fails too (hrd_wbfm_serial_count); both wait for the batch's queued work */
int hrd_wbfm_fallback_count(hrd_batch_t *b, uint64_t *count);
int hrd_wbfm_serial_count(hrd_batch_t *b, uint64_t *count);
/* copy a device table back: 0 = atan2 LUT (65536 floats), 1 = NCO sin,
 * 2 = NCO cos (16384 floats each) */
int hrd_get_table(hrd_batch_t *b, int which, float *out, size_t n);
/* quantised taps as uploaded, same numbering as oracle/hrd_oracle.h HRO_TAPS_* */
int hrd_get_taps(int which, int16_t *out, int cap);
/* bytes of per-stream state the batch keeps in HBM */
size_t hrd_state_bytes_per_stream(int kind);

#ifdef __cplusplus
}
#endif
#endif /* HRD_H */
