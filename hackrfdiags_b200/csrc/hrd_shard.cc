// hrd_shard.cc -- the multi-GPU host entry: one job of N streams on G GPUs of one box.
//
// Streams are independent -- every piece of state belongs to one reference object graph (one IqDataProcessor with
// its demodulators, or one set of modulators) -- so a job of N streams on G GPUs is G disjoint jobs: shard g owns the
// contiguous range [g*N/G, (g+1)*N/G) (the same rule as hackrfdiags_b200/shard.py) and a batch of its own on its
// device, inputs go straight from the caller's host rows to the owning GPU, and NOTHING crosses GPUs: no collective,
// no peer traffic.  What this file adds to G hand-made batches is the plumbing a caller would otherwise repeat: one
// worker thread per shard (each with its device and the batch's own CUDA stream), global stream numbers for the
// setters, and process calls that hand every shard its rows at the same time and return when all are done.
//
// A device may appear more than once in the list (two shards on one GPU): that is how the single-GPU tests
// exercise the partition.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "../../include/hrd.h"

namespace {
struct Shard {
    int device = 0, lo = 0, hi = 0;
    hrd_batch_t *batch = nullptr;
    std::thread worker;
    std::mutex m;
    std::condition_variable cv;
    std::function<int()> task; // set by the caller, run by the worker
    bool has_task = false, done = false, quit = false;
    int rc = 0;
    char err[256] = "";
};
} // namespace

struct hrd_sharded {
    int n = 0, kind = HRD_RX;
    std::vector<Shard *> shards;
    char err[320] = "";
};

namespace {
void worker_loop(Shard *s)
{
    for (;;) {
        std::function<int()> task;
        {
            std::unique_lock<std::mutex> lk(s->m);
            s->cv.wait(lk, [&] { return s->has_task || s->quit; });
            if (s->quit) return;
            task = s->task;
            s->has_task = false;
        }
        const int rc = task();
        {
            std::lock_guard<std::mutex> lk(s->m);
            s->rc = rc;
            if (rc) { // hrd_last_error() is thread-local: keep the worker's message
                strncpy(s->err, hrd_last_error(), sizeof s->err - 1);
                s->err[sizeof s->err - 1] = 0;
            }
            s->done = true;
        }
        s->cv.notify_all();
    }
}

// run f(shard) on every shard's worker at the same time; first non-zero status wins
int run_all(hrd_sharded *sh, const std::function<int(Shard &)> &f)
{
    for (Shard *s : sh->shards) {
        std::lock_guard<std::mutex> lk(s->m);
        s->task = [s, &f] { return f(*s); };
        s->has_task = true;
        s->done = false;
    }
    for (Shard *s : sh->shards) s->cv.notify_all();
    int rc = 0;
    for (Shard *s : sh->shards) {
        std::unique_lock<std::mutex> lk(s->m);
        s->cv.wait(lk, [&] { return s->done; });
        if (s->rc && !rc) {
            rc = s->rc;
            snprintf(sh->err, sizeof sh->err, "shard on device %d (streams %d..%d): %s", s->device, s->lo, s->hi - 1, s->err);
        }
    }
    return rc;
}

// the shards a (global) stream argument addresses: one, or all of them
template <class F> int for_stream(hrd_sharded *sh, int stream, F f)
{
    if (!sh) return HRD_EINVAL;
    if (stream != HRD_ALL_STREAMS && (stream < 0 || stream >= sh->n)) return HRD_EINVAL;
    for (Shard *s : sh->shards) {
        if (stream == HRD_ALL_STREAMS) {
            const int rc = f(*s, HRD_ALL_STREAMS);
            if (rc) return rc;
        } else if (stream >= s->lo && stream < s->hi) {
            return f(*s, stream - s->lo);
        }
    }
    return HRD_OK;
}
} // namespace

extern "C" {

int hrd_sharded_destroy(hrd_sharded_t *sh)
{
    if (!sh) return HRD_OK;
    for (Shard *s : sh->shards) {
        if (s->worker.joinable()) {
            {
                std::lock_guard<std::mutex> lk(s->m);
                s->quit = true;
            }
            s->cv.notify_all();
            s->worker.join();
        }
        if (s->batch) hrd_destroy(s->batch);
        delete s;
    }
    delete sh;
    return HRD_OK;
}

int hrd_sharded_create(const int *devices, int n_devices, int n_streams, int kind, hrd_sharded_t **out)
{
    if (!out) return HRD_EINVAL;
    *out = nullptr;
    if (!devices || n_devices <= 0 || n_streams < n_devices) return HRD_EINVAL;
    hrd_sharded *sh = new (std::nothrow) hrd_sharded;
    if (!sh) return HRD_ENOMEM;
    sh->n = n_streams;
    sh->kind = kind;
    for (int g = 0; g < n_devices; g++) {
        Shard *s = new (std::nothrow) Shard;
        if (!s) {
            hrd_sharded_destroy(sh);
            return HRD_ENOMEM;
        }
        s->device = devices[g];
        s->lo = (int)((long long)g * n_streams / n_devices);
        s->hi = (int)((long long)(g + 1) * n_streams / n_devices);
        sh->shards.push_back(s);
        const int rc = hrd_create(s->device, s->hi - s->lo, kind, &s->batch);
        if (rc) {
            hrd_sharded_destroy(sh);
            return rc;
        }
        s->worker = std::thread(worker_loop, s);
    }
    *out = sh;
    return HRD_OK;
}

int hrd_sharded_count(hrd_sharded_t *sh) { return sh ? (int)sh->shards.size() : HRD_EINVAL; }

int hrd_sharded_shard(hrd_sharded_t *sh, int shard, int *device, int *lo, int *hi, hrd_batch_t **batch)
{
    if (!sh || shard < 0 || shard >= (int)sh->shards.size()) return HRD_EINVAL;
    const Shard *s = sh->shards[(size_t)shard];
    if (device) *device = s->device;
    if (lo) *lo = s->lo;
    if (hi) *hi = s->hi;
    if (batch) *batch = s->batch;
    return HRD_OK;
}

const char *hrd_sharded_last_error(hrd_sharded_t *sh) { return sh ? sh->err : "null handle"; }

int hrd_sharded_set_mode(hrd_sharded_t *sh, int stream, int mode)
{
    return for_stream(sh, stream, [&](Shard &s, int local) { return hrd_set_mode(s.batch, local, mode); });
}

int hrd_sharded_set_param(hrd_sharded_t *sh, int stream, int param, float value)
{
    return for_stream(sh, stream, [&](Shard &s, int local) { return hrd_set_param(s.batch, local, param, value); });
}

int hrd_sharded_reset(hrd_sharded_t *sh, int stream, int unit)
{
    return for_stream(sh, stream, [&](Shard &s, int local) { return hrd_reset(s.batch, local, unit); });
}

int hrd_sharded_set_option(hrd_sharded_t *sh, int option, int value)
{
    return for_stream(sh, HRD_ALL_STREAMS, [&](Shard &s, int) { return hrd_set_option(s.batch, option, value); });
}

// hrd_rx_process on host rows, every shard at once: stream s of the job reads iq + s*iq_stride and writes
// pcm + s*pcm_stride (and pcm_counts[s]) wherever it lives
int hrd_sharded_rx_process(hrd_sharded_t *sh, const int8_t *iq, size_t bytes_per_stream, size_t iq_stride, int entry, int16_t *pcm,
                           size_t pcm_stride, uint32_t *pcm_counts)
{
    if (!sh || sh->kind != HRD_RX || !iq || !pcm) return HRD_EINVAL;
    return run_all(sh, [&](Shard &s) {
        return hrd_rx_process(s.batch, iq + (size_t)s.lo * iq_stride, bytes_per_stream, iq_stride, entry, pcm + (size_t)s.lo * pcm_stride,
                              pcm_stride, pcm_counts ? pcm_counts + s.lo : nullptr, HRD_MEM_HOST, nullptr);
    });
}

int hrd_sharded_tx_process(hrd_sharded_t *sh, const int16_t *pcm, size_t n_per_stream, size_t pcm_stride, int8_t *iq, size_t iq_stride)
{
    if (!sh || sh->kind != HRD_TX || !iq || !pcm) return HRD_EINVAL;
    return run_all(sh, [&](Shard &s) {
        return hrd_tx_process(s.batch, pcm + (size_t)s.lo * pcm_stride, n_per_stream, pcm_stride, iq + (size_t)s.lo * iq_stride, iq_stride,
                              HRD_MEM_HOST, nullptr);
    });
}

} // extern "C"
