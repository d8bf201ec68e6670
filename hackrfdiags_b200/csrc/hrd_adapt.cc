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
#include <stdint.h>
#include <string.h>

#include <deque>
#include <mutex>
#include <new>
#include <vector>

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
    std::vector<int8_t> pool;    // IQ_SLOTS x IQ_BLOCK
};
} // namespace

struct hrd_pcm_ring {
    int n = 0;
    PcmStream *s = nullptr;
    std::vector<int16_t> rows;   // hrd_tx_from_ring: one gathered block per stream
};
struct hrd_iq_queue {
    int n = 0;
    IqStream *s = nullptr;
    std::vector<int8_t> rows;    // hrd_rx_from_queue: one gathered block per stream
    std::vector<uint32_t> bytes;
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
    for (int i = 0; i < n_streams; i++) q->s[i].pool.assign((size_t)IQ_SLOTS * IQ_BLOCK, 0);
    *out = q;
    return HRD_OK;
}

int hrd_iq_queue_destroy(hrd_iq_queue_t *q)
{
    if (q) {
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
    memcpy(s.pool.data() + (size_t)slot * IQ_BLOCK, data, bytes);
    s.queue.push_back(slot);
    s.index = (s.index + 1) % IQ_SLOTS;
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
        memcpy(rows + (size_t)i * row_stride, s.pool.data() + (size_t)slot * IQ_BLOCK, s.meta[slot].byte_count);
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

// dataConsumerProcedure for every stream: one round through hrd_rx_process (HRD_ENTRY_2048K); returns 1 when a round
// was processed, 0 when some stream had nothing queued, negative on error.  Rounds need equal block sizes.
int hrd_rx_from_queue(hrd_batch_t *b, hrd_iq_queue_t *q, int16_t *pcm, size_t pcm_stride, uint32_t *pcm_counts)
{
    if (!b || !q) return HRD_EINVAL;
    q->rows.resize((size_t)q->n * IQ_BLOCK); // allocated once, reused every round
    q->bytes.resize((size_t)q->n);
    const int got = hrd_iq_queue_pop_all(q, q->rows.data(), IQ_BLOCK, q->bytes.data(), nullptr);
    if (got <= 0) return got;
    for (int i = 1; i < q->n; i++)
        if (q->bytes[(size_t)i] != q->bytes[0]) return HRD_EINVAL;
    const int rc = hrd_rx_process(b, q->rows.data(), q->bytes[0] / 512 * 512, IQ_BLOCK, HRD_ENTRY_2048K, pcm, pcm_stride, pcm_counts,
                                  HRD_MEM_HOST, nullptr);
    return rc ? rc : 1;
}

} // extern "C"
