// hrd_adapt.cc -- batched ingest / egress adapters for the two callers of the hot path.
//
// Host code only (no kernels): what sits between the device threads and the DSP objects in the reference,
// restated for MANY streams so that one hrd_rx_process / hrd_tx_process call serves a whole round.
//   Rx  DataConsumer (src_diags/DataConsumer.cc:219-261, 319-351; hdr_diags/DataConsumer.h:15-27): the libusb
//       thread copies each 262144-byte transfer into the next of 16 pool slots and queues it; the consumer thread
//       dequeues and calls IqDataProcessor::acceptIqData.  -> hrd_iq_queue_*: one pool + FIFO per stream; a ROUND
//       is one block of every stream, gathered into a row matrix for hrd_rx_process.
//   Tx  BasebandDataProcessor's PCM ring (src_diags/BasebandDataProcessor.cc:416-433 writer, 482-605 reader with
//       the drop / repeat rate matching, :19-20 start table; hdr_diags/BasebandDataProcessor.h:16-17): the reader
//       thread fills 512-sample blocks, the transmit callback takes one per 262144-byte transfer.
//       -> hrd_pcm_ring_*: one ring per stream with the reference's index arithmetic, and a gather of one block per
//       stream into the row matrix hrd_tx_process reads.
// Like the reference objects: one producer thread and one consumer thread per stream may run concurrently (the
// writer index is under a mutex, as BasebandDataProcessor::writerLock; the queue under its own).
//
// PINNED AND PIPELINED (hrd_rx_pipe_* / hrd_tx_pipe_*).  The block pools are page-locked (cudaHostAlloc; plain memory
// when there is no CUDA device, as in the CPU tests) and laid out slot-major -- [slot][stream][262144] -- so that a
// round in which the streams sit at the same slot (producers running at the same rate, the normal case) goes to the
// GPU as ONE asynchronous copy straight from the pool, no gather.  A pipe keeps `depth` rounds in flight, each on its
// own CUDA stream: the H2D copy of round k+1 runs while the kernels of round k do and the D2H copy of round k-1
// does (the batch orders the kernels of consecutive calls by an event, nothing else).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <deque>
#include <mutex>
#include <new>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/hrd.h"

namespace {
constexpr int RING = 16;          // PCM_RING_SIZE
constexpr int BLOCK = 512;        // PCM_BLOCK_SIZE
constexpr uint32_t IQ_BLOCK = 262144; // DATA_CONSUMER_BUFFER_SIZE
constexpr int IQ_SLOTS = 16;      // DATA_CONSUMER_NUMBER_OF_MESSAGES
// BasebandDataProcessor.cc:19-20: the reader starts eight blocks behind the writer
const int k_reader_start[RING] = {8, 9, 10, 11, 12, 13, 14, 15, 0, 1, 2, 3, 4, 5, 6, 7};

struct PcmStream {
    std::mutex writer_lock;
    uint32_t writer = RING - 1;                    // :78
    uint32_t reader = (uint32_t)k_reader_start[RING - 1]; // :79
    bool synchronized = false, running = false;
    uint32_t produced = 0, consumed = 0, dropped = 0, added = 0;
    int16_t block[RING][BLOCK];
};

struct IqMessage {
    uint32_t time_stamp, byte_count;
};
struct IqStream {
    std::mutex lock;
    std::deque<int> queue;       // MessageQueue of slot numbers
    unsigned long index = 0;     // messageIndex
    uint32_t short_blocks = 0, last_time_stamp = 0;
    IqMessage meta[IQ_SLOTS];
};

// page-locked when a CUDA device is there, plain otherwise
void *alloc_host(size_t bytes, bool *pinned)
{
    void *p = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) == cudaSuccess && count > 0 && cudaHostAlloc(&p, bytes, cudaHostAllocPortable) == cudaSuccess) {
        *pinned = true;
        return p;
    }
    cudaGetLastError();
    *pinned = false;
    return malloc(bytes);
}
void free_host(void *p, bool pinned)
{
    if (!p) return;
    if (pinned) cudaFreeHost(p);
    else free(p);
}
} // namespace

struct hrd_pcm_ring {
    int n = 0;
    PcmStream *s = nullptr;
    std::vector<int16_t> rows;   // hrd_tx_from_ring: one gathered block per stream
};
struct hrd_iq_queue {
    int n = 0;
    IqStream *s = nullptr;
    int8_t *pool = nullptr;      // [IQ_SLOTS][n][IQ_BLOCK], page-locked when possible
    bool pool_pinned = false;
    int8_t *rows = nullptr;      // hrd_rx_from_queue: one gathered block per stream, page-locked when possible
    bool rows_pinned = false;
    std::vector<uint32_t> bytes;
    int8_t *slot_ptr(int slot, int stream) const { return pool + ((size_t)slot * (size_t)n + (size_t)stream) * IQ_BLOCK; }
};

extern "C" {

int hrd_pcm_ring_create(int n_streams, hrd_pcm_ring_t **out)
{
    if (!out || n_streams <= 0) return HRD_EINVAL;
    hrd_pcm_ring *r = new (std::nothrow) hrd_pcm_ring;
    if (!r) return HRD_ENOMEM;
    r->n = n_streams;
    r->s = new (std::nothrow) PcmStream[(size_t)n_streams];
    if (!r->s) {
        delete r;
        return HRD_ENOMEM;
    }
    for (int i = 0; i < n_streams; i++) memset(r->s[i].block, 0, sizeof r->s[i].block);
    *out = r;
    return HRD_OK;
}

int hrd_pcm_ring_destroy(hrd_pcm_ring_t *r)
{
    if (r) {
        delete[] r->s;
        delete r;
    }
    return HRD_OK;
}

int hrd_pcm_ring_start(hrd_pcm_ring_t *r, int stream, int running)
{
    if (!r || stream < HRD_ALL_STREAMS || stream >= r->n) return HRD_EINVAL;
    for (int i = (stream < 0 ? 0 : stream); i < (stream < 0 ? r->n : stream + 1); i++) r->s[i].running = running != 0;
    return HRD_OK;
}

// getNextUnfilledBuffer (:416-433) + the fread of basebandReaderProcedure (:862-865)
int hrd_pcm_ring_write(hrd_pcm_ring_t *r, int stream, const int16_t *pcm, uint32_t n_samples)
{
    if (!r || stream < 0 || stream >= r->n || !pcm || n_samples > BLOCK) return HRD_EINVAL;
    PcmStream &p = r->s[stream];
    uint32_t w;
    {
        std::lock_guard<std::mutex> g(p.writer_lock);
        p.writer++;
        p.writer %= RING;
        w = p.writer;
    }
    p.produced++;
    memcpy(p.block[w], pcm, n_samples * sizeof(int16_t)); // a short read leaves the rest of the slot as it was
    return HRD_OK;
}

// getNextFilledBuffer (:482-605) for one stream; returns the slot to send, or -1 for the zero block
static int next_filled(PcmStream &p)
{
    int32_t u;
    {
        std::lock_guard<std::mutex> g(p.writer_lock);
        u = (int32_t)p.writer;
    }
    int32_t l = (int32_t)p.reader;
    if (u < l) u += RING - 1; // (sic: the reference adds PCM_RING_SIZE - 1)
    const int32_t lag = u - l;
    if (lag > 10) { // the writer runs ahead: drop a block
        p.reader++;
        p.reader %= RING;
        p.dropped++;
    } else if (lag < 6) { // the writer falls behind: send the previous block again
        int32_t d = (int32_t)p.reader - 1;
        if (d < 0) d += RING;
        p.reader = (uint32_t)d;
        p.added++;
    }
    if (!p.running) return -1;
    if (!p.synchronized) {
        p.synchronized = true;
        std::lock_guard<std::mutex> g(p.writer_lock);
        p.reader = (uint32_t)k_reader_start[p.writer];
    }
    const int slot = (int)p.reader;
    p.reader++;
    p.reader %= RING;
    p.consumed++;
    return slot;
}

// one transmit callback of EVERY stream: rows[s * row_stride .. +512) <- the block getNextFilledBuffer picks
int hrd_pcm_ring_read_all(hrd_pcm_ring_t *r, int16_t *rows, size_t row_stride, int32_t *slots)
{
    if (!r || !rows || row_stride < BLOCK) return HRD_EINVAL;
    for (int i = 0; i < r->n; i++) {
        const int slot = next_filled(r->s[i]);
        if (slots) slots[i] = slot;
        if (slot < 0)
            memset(rows + (size_t)i * row_stride, 0, BLOCK * sizeof(int16_t)); // zeroPcmBuffer
        else
            memcpy(rows + (size_t)i * row_stride, r->s[i].block[slot], BLOCK * sizeof(int16_t));
    }
    return HRD_OK;
}

int hrd_pcm_ring_stats(hrd_pcm_ring_t *r, int stream, uint32_t out[4])
{
    if (!r || stream < 0 || stream >= r->n || !out) return HRD_EINVAL;
    const PcmStream &p = r->s[stream];
    out[0] = p.produced, out[1] = p.consumed, out[2] = p.dropped, out[3] = p.added;
    return HRD_OK;
}

// BasebandDataProcessor::getIqData (:381-389 -> modulateBasebandData :630-697) for every stream of a Tx batch
int hrd_tx_from_ring(hrd_batch_t *b, hrd_pcm_ring_t *r, int8_t *iq, size_t iq_stride, int mem, void *cuda_stream)
{
    if (!b || !r) return HRD_EINVAL;
    if (mem != HRD_MEM_HOST) return HRD_EINVAL; // the gathered rows live in host memory
    r->rows.resize((size_t)r->n * BLOCK);
    int rc = hrd_pcm_ring_read_all(r, r->rows.data(), BLOCK, nullptr);
    if (rc) return rc;
    return hrd_tx_process(b, r->rows.data(), BLOCK, BLOCK, iq, iq_stride, HRD_MEM_HOST, cuda_stream);
}

// ---------------------------------------------------------------------------------------------------------
int hrd_iq_queue_create(int n_streams, hrd_iq_queue_t **out)
{
    if (!out || n_streams <= 0) return HRD_EINVAL;
    hrd_iq_queue *q = new (std::nothrow) hrd_iq_queue;
    if (!q) return HRD_ENOMEM;
    q->n = n_streams;
    q->s = new (std::nothrow) IqStream[(size_t)n_streams];
    if (!q->s) {
        delete q;
        return HRD_ENOMEM;
    }
    q->pool = (int8_t *)alloc_host((size_t)IQ_SLOTS * (size_t)n_streams * IQ_BLOCK, &q->pool_pinned);
    if (!q->pool) {
        delete[] q->s;
        delete q;
        return HRD_ENOMEM;
    }
    *out = q;
    return HRD_OK;
}

int hrd_iq_queue_destroy(hrd_iq_queue_t *q)
{
    if (q) {
        free_host(q->pool, q->pool_pinned);
        free_host(q->rows, q->rows_pinned);
        delete[] q->s;
        delete q;
    }
    return HRD_OK;
}

// DataConsumer::acceptData (:219-261): clamp, count short blocks, copy into the next slot, queue it.  Like the
// reference there is no overflow check: the seventeenth unconsumed block overwrites the first.
int hrd_iq_queue_push(hrd_iq_queue_t *q, int stream, uint32_t time_stamp, const void *data, uint32_t bytes)
{
    if (!q || stream < 0 || stream >= q->n || !data) return HRD_EINVAL;
    IqStream &s = q->s[stream];
    std::lock_guard<std::mutex> g(s.lock);
    s.last_time_stamp = time_stamp;
    if (bytes > IQ_BLOCK) bytes = IQ_BLOCK;
    else if (bytes < IQ_BLOCK) s.short_blocks++;
    const int slot = (int)s.index;
    s.meta[slot].time_stamp = time_stamp;
    s.meta[slot].byte_count = bytes;
    memcpy(q->slot_ptr(slot, stream), data, bytes);
    s.queue.push_back(slot);
    s.index = (s.index + 1) % IQ_SLOTS;
    return HRD_OK;
}

// One transfer of each of `count` consecutive streams (a receive thread that serves several radios): row i is the
// block of stream first + i.  The same as `count` hrd_iq_queue_push calls.
int hrd_iq_queue_push_rows(hrd_iq_queue_t *q, int first, int count, uint32_t time_stamp, const void *rows, size_t row_stride, uint32_t bytes)
{
    if (!q || first < 0 || count < 0 || first + count > q->n || !rows) return HRD_EINVAL;
    for (int i = 0; i < count; i++) {
        const int rc = hrd_iq_queue_push(q, first + i, time_stamp, (const char *)rows + (size_t)i * row_stride, bytes);
        if (rc) return rc;
    }
    return HRD_OK;
}

// the consumer thread's dequeue (:319-351) for a whole round: returns 1 and one block per stream when every
// stream has one queued, 0 (and takes nothing) otherwise
int hrd_iq_queue_pop_all(hrd_iq_queue_t *q, int8_t *rows, size_t row_stride, uint32_t *bytes, uint32_t *time_stamps)
{
    if (!q || !rows || row_stride < IQ_BLOCK) return HRD_EINVAL;
    for (int i = 0; i < q->n; i++) {
        std::lock_guard<std::mutex> g(q->s[i].lock);
        if (q->s[i].queue.empty()) return 0;
    }
    for (int i = 0; i < q->n; i++) {
        IqStream &s = q->s[i];
        std::lock_guard<std::mutex> g(s.lock);
        const int slot = s.queue.front();
        s.queue.pop_front();
        memcpy(rows + (size_t)i * row_stride, q->slot_ptr(slot, i), s.meta[slot].byte_count);
        if (bytes) bytes[i] = s.meta[slot].byte_count;
        if (time_stamps) time_stamps[i] = s.meta[slot].time_stamp;
    }
    return 1;
}

int hrd_iq_queue_stats(hrd_iq_queue_t *q, int stream, uint32_t out[3])
{
    if (!q || stream < 0 || stream >= q->n || !out) return HRD_EINVAL;
    IqStream &s = q->s[stream];
    std::lock_guard<std::mutex> g(s.lock);
    out[0] = (uint32_t)s.queue.size(), out[1] = s.short_blocks, out[2] = s.last_time_stamp;
    return HRD_OK;
}

// the sizes of the blocks at the head of every queue, WITHOUT taking them: 1 = every stream has one (bytes[] filled,
// *uniform tells whether they are all equal), 0 = some stream has none
static int peek_round(hrd_iq_queue_t *q, uint32_t *bytes, bool *uniform, int *slots)
{
    *uniform = true;
    for (int i = 0; i < q->n; i++) {
        IqStream &s = q->s[i];
        std::lock_guard<std::mutex> g(s.lock);
        if (s.queue.empty()) return 0;
        const int slot = s.queue.front();
        bytes[i] = s.meta[slot].byte_count;
        if (slots) slots[i] = slot;
        if (bytes[i] != bytes[0]) *uniform = false;
    }
    return 1;
}

// dataConsumerProcedure for every stream: one round through hrd_rx_process (HRD_ENTRY_2048K); returns 1 when a round
// was processed, 0 when some stream had nothing queued, negative on error.  A round needs equal block sizes, a whole
// number of PCM samples each; otherwise HRD_EINVAL comes back and NOTHING is dequeued (the caller decides).
int hrd_rx_from_queue(hrd_batch_t *b, hrd_iq_queue_t *q, int16_t *pcm, size_t pcm_stride, uint32_t *pcm_counts)
{
    if (!b || !q) return HRD_EINVAL;
    if (!q->rows) q->rows = (int8_t *)alloc_host((size_t)q->n * IQ_BLOCK, &q->rows_pinned); // once, reused every round
    if (!q->rows) return HRD_ENOMEM;
    q->bytes.resize((size_t)q->n);
    bool uniform = true;
    const int ready = peek_round(q, q->bytes.data(), &uniform, nullptr);
    if (ready <= 0) return ready;
    if (!uniform || q->bytes[0] % 512) return HRD_EINVAL;
    const int got = hrd_iq_queue_pop_all(q, q->rows, IQ_BLOCK, q->bytes.data(), nullptr);
    if (got <= 0) return got;
    const int rc = hrd_rx_process(b, q->rows, q->bytes[0], IQ_BLOCK, HRD_ENTRY_2048K, pcm, pcm_stride, pcm_counts,
                                  HRD_MEM_HOST, nullptr);
    return rc ? rc : 1;
}

// ---------------------------------------------------------------------------------------------------------
// pipelined rounds (see the header of this file)
// ---------------------------------------------------------------------------------------------------------
} // extern "C"

namespace {
// the pipes' CUDA objects live on the batch's device, whatever the caller's current device is
struct OnDevice {
    int prev = -1;
    explicit OnDevice(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~OnDevice()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
int device_of(hrd_batch_t *b)
{
    int d = 0;
    hrd_get_device(b, &d);
    return d;
}
struct RxRound {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    int8_t *d_iq = nullptr;      // [n][IQ_BLOCK]
    int16_t *d_pcm = nullptr;    // [n][512]
    int16_t *h_pcm = nullptr;    // pinned
    uint32_t *counts = nullptr;  // host, n entries
    uint32_t bytes = 0;
    bool busy = false;
};
struct TxRound {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    int16_t *h_rows = nullptr;   // pinned [n][512]
    int16_t *d_rows = nullptr;
    int8_t *d_iq = nullptr;      // [n][262144]
    int8_t *h_iq = nullptr;      // pinned
    bool busy = false;
};
} // namespace

struct hrd_rx_pipe {
    hrd_batch_t *b = nullptr;
    hrd_iq_queue_t *q = nullptr;
    int depth = 0, head = 0, tail = 0, in_flight = 0, device = 0;
    std::vector<RxRound> r;
    std::vector<uint32_t> bytes;
    std::vector<int> slots;
    uint64_t rounds = 0, copies = 0;
};
struct hrd_tx_pipe {
    hrd_batch_t *b = nullptr;
    hrd_pcm_ring_t *ring = nullptr;
    int depth = 0, head = 0, tail = 0, in_flight = 0, device = 0;
    std::vector<TxRound> r;
};

extern "C" {

int hrd_rx_pipe_destroy(hrd_rx_pipe_t *p)
{
    if (!p) return HRD_OK;
    OnDevice guard(p->device);
    for (RxRound &k : p->r) {
        if (k.stream) cudaStreamSynchronize(k.stream);
        if (k.done) cudaEventDestroy(k.done);
        cudaFree(k.d_iq);
        cudaFree(k.d_pcm);
        if (k.h_pcm) cudaFreeHost(k.h_pcm);
        if (k.counts) cudaFreeHost(k.counts);
        if (k.stream) cudaStreamDestroy(k.stream);
    }
    delete p;
    return HRD_OK;
}

int hrd_rx_pipe_create(hrd_batch_t *b, hrd_iq_queue_t *q, int depth, hrd_rx_pipe_t **out)
{
    if (!b || !q || !out || depth < 1 || depth > IQ_SLOTS / 2) return HRD_EINVAL;
    hrd_rx_pipe *p = new (std::nothrow) hrd_rx_pipe;
    if (!p) return HRD_ENOMEM;
    p->b = b, p->q = q, p->depth = depth, p->device = device_of(b);
    OnDevice guard(p->device);
    p->r.resize((size_t)depth);
    p->bytes.resize((size_t)q->n);
    p->slots.resize((size_t)q->n);
    const size_t n = (size_t)q->n;
    cudaError_t e = cudaSuccess;
    for (RxRound &k : p->r) {
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&k.stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&k.done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaMalloc(&k.d_iq, n * IQ_BLOCK);
        if (e == cudaSuccess) e = cudaMalloc(&k.d_pcm, n * 512 * sizeof(int16_t));
        if (e == cudaSuccess) e = cudaHostAlloc(&k.h_pcm, n * 512 * sizeof(int16_t), cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaHostAlloc(&k.counts, n * sizeof(uint32_t), cudaHostAllocDefault);
    }
    if (e != cudaSuccess) {
        hrd_rx_pipe_destroy(p);
        return HRD_ECUDA;
    }
    *out = p;
    return HRD_OK;
}

// Start one round if every stream has a block queued and fewer than `depth` rounds are in flight: the blocks go to
// the GPU straight from their pool slots (one copy per run of streams that sit at the same slot), the kernels and the
// PCM's way back are queued behind them; returns 1 (started), 0 (nothing to do / pipe full), negative on error.
int hrd_rx_pipe_submit(hrd_rx_pipe_t *p)
{
    if (!p) return HRD_EINVAL;
    if (p->in_flight == p->depth) return 0;
    OnDevice guard(p->device);
    hrd_iq_queue_t *q = p->q;
    bool uniform = true;
    const int ready = peek_round(q, p->bytes.data(), &uniform, p->slots.data());
    if (ready <= 0) return ready;
    if (!uniform || p->bytes[0] % 512 || p->bytes[0] == 0) return HRD_EINVAL;
    RxRound &k = p->r[(size_t)p->head];
    k.bytes = p->bytes[0];
    // H2D: runs of consecutive streams at the same slot are contiguous in the slot-major pool
    for (int i = 0; i < q->n;) {
        int j = i + 1;
        while (j < q->n && p->slots[(size_t)j] == p->slots[(size_t)i]) j++;
        if (cudaMemcpy2DAsync(k.d_iq + (size_t)i * IQ_BLOCK, IQ_BLOCK, q->slot_ptr(p->slots[(size_t)i], i), IQ_BLOCK, k.bytes,
                              (size_t)(j - i), cudaMemcpyHostToDevice, k.stream) != cudaSuccess)
            return HRD_ECUDA;
        p->copies++;
        i = j;
    }
    // the blocks are on their way: take them off the queues (their slots stay untouched until the producer has
    // gone around the pool, as in the reference)
    for (int i = 0; i < q->n; i++) {
        IqStream &s = q->s[i];
        std::lock_guard<std::mutex> g(s.lock);
        s.queue.pop_front();
    }
    int rc = hrd_rx_process(p->b, k.d_iq, k.bytes, IQ_BLOCK, HRD_ENTRY_2048K, k.d_pcm, 512, k.counts, HRD_MEM_DEVICE, k.stream);
    if (rc) return rc;
    if (cudaMemcpyAsync(k.h_pcm, k.d_pcm, (size_t)q->n * 512 * sizeof(int16_t), cudaMemcpyDeviceToHost, k.stream) != cudaSuccess ||
        cudaEventRecord(k.done, k.stream) != cudaSuccess)
        return HRD_ECUDA;
    k.busy = true;
    p->head = (p->head + 1) % p->depth;
    p->in_flight++;
    p->rounds++;
    return 1;
}

// The oldest round in flight: waits for it and hands out its PCM (rows of 512, valid until the slot is reused,
// i.e. for the next depth - 1 submits) and per-stream sample counts; 1 = a round, 0 = none in flight.
int hrd_rx_pipe_collect(hrd_rx_pipe_t *p, const int16_t **pcm, size_t *pcm_stride, const uint32_t **pcm_counts)
{
    if (!p) return HRD_EINVAL;
    if (!p->in_flight) return 0;
    OnDevice guard(p->device);
    RxRound &k = p->r[(size_t)p->tail];
    if (cudaEventSynchronize(k.done) != cudaSuccess) return HRD_ECUDA;
    if (pcm) *pcm = k.h_pcm;
    if (pcm_stride) *pcm_stride = 512;
    if (pcm_counts) *pcm_counts = k.counts;
    k.busy = false;
    p->tail = (p->tail + 1) % p->depth;
    p->in_flight--;
    return 1;
}

// rounds started, host-to-device copies issued for them (1 per round when the streams move in step)
int hrd_rx_pipe_stats(hrd_rx_pipe_t *p, uint64_t out[2])
{
    if (!p || !out) return HRD_EINVAL;
    out[0] = p->rounds, out[1] = p->copies;
    return HRD_OK;
}

int hrd_tx_pipe_destroy(hrd_tx_pipe_t *p)
{
    if (!p) return HRD_OK;
    OnDevice guard(p->device);
    for (TxRound &k : p->r) {
        if (k.stream) cudaStreamSynchronize(k.stream);
        if (k.done) cudaEventDestroy(k.done);
        if (k.h_rows) cudaFreeHost(k.h_rows);
        cudaFree(k.d_rows);
        cudaFree(k.d_iq);
        if (k.h_iq) cudaFreeHost(k.h_iq);
        if (k.stream) cudaStreamDestroy(k.stream);
    }
    delete p;
    return HRD_OK;
}

int hrd_tx_pipe_create(hrd_batch_t *b, hrd_pcm_ring_t *r, int depth, hrd_tx_pipe_t **out)
{
    if (!b || !r || !out || depth < 1 || depth > 8) return HRD_EINVAL;
    hrd_tx_pipe *p = new (std::nothrow) hrd_tx_pipe;
    if (!p) return HRD_ENOMEM;
    p->b = b, p->ring = r, p->depth = depth, p->device = device_of(b);
    OnDevice guard(p->device);
    p->r.resize((size_t)depth);
    const size_t n = (size_t)r->n;
    cudaError_t e = cudaSuccess;
    for (TxRound &k : p->r) {
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&k.stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&k.done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaHostAlloc(&k.h_rows, n * BLOCK * sizeof(int16_t), cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaMalloc(&k.d_rows, n * BLOCK * sizeof(int16_t));
        if (e == cudaSuccess) e = cudaMalloc(&k.d_iq, n * IQ_BLOCK);
        if (e == cudaSuccess) e = cudaHostAlloc(&k.h_iq, n * IQ_BLOCK, cudaHostAllocDefault);
    }
    if (e != cudaSuccess) {
        hrd_tx_pipe_destroy(p);
        return HRD_ECUDA;
    }
    *out = p;
    return HRD_OK;
}

// One transmit callback of every stream, started: the ring policy picks a block per stream (pinned rows), the
// modulators run, the 262144 bytes per stream come back into pinned memory; 1 = started, 0 = the pipe is full.
int hrd_tx_pipe_submit(hrd_tx_pipe_t *p)
{
    if (!p) return HRD_EINVAL;
    if (p->in_flight == p->depth) return 0;
    OnDevice guard(p->device);
    TxRound &k = p->r[(size_t)p->head];
    const size_t n = (size_t)p->ring->n;
    int rc = hrd_pcm_ring_read_all(p->ring, k.h_rows, BLOCK, nullptr);
    if (rc) return rc;
    if (cudaMemcpyAsync(k.d_rows, k.h_rows, n * BLOCK * sizeof(int16_t), cudaMemcpyHostToDevice, k.stream) != cudaSuccess) return HRD_ECUDA;
    rc = hrd_tx_process(p->b, k.d_rows, BLOCK, BLOCK, k.d_iq, IQ_BLOCK, HRD_MEM_DEVICE, k.stream);
    if (rc) return rc;
    if (cudaMemcpyAsync(k.h_iq, k.d_iq, n * IQ_BLOCK, cudaMemcpyDeviceToHost, k.stream) != cudaSuccess ||
        cudaEventRecord(k.done, k.stream) != cudaSuccess)
        return HRD_ECUDA;
    k.busy = true;
    p->head = (p->head + 1) % p->depth;
    p->in_flight++;
    return 1;
}

int hrd_tx_pipe_collect(hrd_tx_pipe_t *p, const int8_t **iq, size_t *iq_stride)
{
    if (!p) return HRD_EINVAL;
    if (!p->in_flight) return 0;
    OnDevice guard(p->device);
    TxRound &k = p->r[(size_t)p->tail];
    if (cudaEventSynchronize(k.done) != cudaSuccess) return HRD_ECUDA;
    if (iq) *iq = k.h_iq;
    if (iq_stride) *iq_stride = IQ_BLOCK;
    k.busy = false;
    p->tail = (p->tail + 1) % p->depth;
    p->in_flight--;
    return 1;
}

} // extern "C"
