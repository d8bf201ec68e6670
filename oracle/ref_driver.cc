/*
 * ref_driver.cc -- C-ABI harness around the UNMODIFIED reference classes.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is ours; it contains no reference
 * code.  oracle/Makefile compiles it together with the reference sources
 * where they lie under /root/reference (g++ -O3, the reference's own flags,
 * see radioDiags/build*Lib.sh) into oracle/_ref/libhrd_ref.so.  That library
 * is (a) what the C restatement in hrd_oracle.c is pinned against, (b) the
 * generator of tests/golden/, and (c) the "reference" CPU baseline that
 * bench.py times.  It is never linked into, or called from, the product.
 *
 * The reference prints through an extern nprintf() and reads an extern gain
 * variable; both are defined here the way its own test apps do
 * (radioDiags/AmModulator/am.cc:93-110).
 */
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <chrono>
#include <thread>
#include <vector>

/* reach decimatedData[] and the quantised taps for table dumps */
#define private public
#include "IqDataProcessor.h"
#include "AmModulator.h"
#include "FmModulator.h"
#include "WbFmModulator.h"
#include "SsbModulator.h"
#undef private

uint32_t radio_adjustableReceiveGainInDb = 16; /* Radio.cc default */

void nprintf(FILE *s, const char *formatPtr, ...)
{
    (void)s;
    (void)formatPtr;
}

namespace {

struct PcmSink {
    int16_t *dst;
    size_t count;
};
thread_local PcmSink *g_sink = nullptr;

void pcm_callback(int16_t *bufferPtr, uint32_t bufferLength)
{
    if (!g_sink) return;
    memcpy(g_sink->dst + g_sink->count, bufferPtr, bufferLength * sizeof(int16_t));
    g_sink->count += bufferLength;
}

struct RefRx {
    IqDataProcessor *iqdp;
    AmDemodulator *am;
    FmDemodulator *fm;
    WbFmDemodulator *wbfm;
    SsbDemodulator *ssb;
    int mode;
};

struct RefTx {
    AmModulator *am;
    FmModulator *fm;
    WbFmModulator *wbfm;
    SsbModulator *ssb;
};

} // namespace

/* the PCM ring of BasebandDataProcessor is private; the harness below drives it event by event (no threads) */
#define private public
#include "BasebandDataProcessor.h"
#undef private

extern "C" {

void *ref_rx_new(void)
{
    RefRx *rx = new RefRx;
    char host[] = "127.0.0.1";
    rx->iqdp = new IqDataProcessor(host, 8000);
    rx->am = new AmDemodulator(pcm_callback);
    rx->fm = new FmDemodulator(pcm_callback);
    rx->wbfm = new WbFmDemodulator(pcm_callback);
    rx->ssb = new SsbDemodulator(pcm_callback);
    rx->iqdp->setAmDemodulator(rx->am);
    rx->iqdp->setFmDemodulator(rx->fm);
    rx->iqdp->setWbFmDemodulator(rx->wbfm);
    rx->iqdp->setSsbDemodulator(rx->ssb);
    rx->mode = 0;
    return rx;
}

void ref_rx_free(void *h)
{
    RefRx *rx = (RefRx *)h;
    delete rx->iqdp;
    delete rx->am;
    delete rx->fm;
    delete rx->wbfm;
    delete rx->ssb;
    delete rx;
}

void ref_rx_set_mode(void *h, int mode)
{
    RefRx *rx = (RefRx *)h;
    rx->mode = mode;
    rx->iqdp->setDemodulatorMode((IqDataProcessor::demodulatorType)mode);
}

void ref_rx_set_gain(void *h, int demod, float gain)
{
    RefRx *rx = (RefRx *)h;
    switch (demod) {
    case 0: rx->am->setDemodulatorGain(gain); break;
    case 1: rx->fm->setDemodulatorGain(gain); break;
    case 2: rx->wbfm->setDemodulatorGain(gain); break;
    case 3: rx->ssb->setDemodulatorGain(gain); break;
    }
}

void ref_rx_reset_demod(void *h, int demod)
{
    RefRx *rx = (RefRx *)h;
    switch (demod) {
    case 0: rx->am->resetDemodulator(); break;
    case 1: rx->fm->resetDemodulator(); break;
    case 2: rx->wbfm->resetDemodulator(); break;
    case 3: rx->ssb->resetDemodulator(); break;
    }
}

/* reduceSampleRate + upconvertByFsOver4 only; copies decimatedData out */
size_t ref_rx_front_end(void *h, const int8_t *iq, size_t nbytes, int8_t *out256k)
{
    RefRx *rx = (RefRx *)h;
    uint32_t n = rx->iqdp->reduceSampleRate((int8_t *)iq, (uint32_t)nbytes);
    rx->iqdp->upconvertByFsOver4(rx->iqdp->decimatedData, n);
    memcpy(out256k, rx->iqdp->decimatedData, n);
    return n;
}

/* IqDataProcessor::acceptIqData, one call (nbytes <= 262144) */
size_t ref_rx_accept_2048k(void *h, const int8_t *iq, size_t nbytes, int16_t *pcm)
{
    RefRx *rx = (RefRx *)h;
    PcmSink sink = {pcm, 0};
    g_sink = &sink;
    rx->iqdp->acceptIqData(0, (int8_t *)iq, nbytes);
    g_sink = nullptr;
    return sink.count;
}

/* <X>Demodulator::acceptIqData, one call (nbytes <= 32768) */
size_t ref_rx_accept_256k(void *h, const int8_t *iq, size_t nbytes, int16_t *pcm)
{
    RefRx *rx = (RefRx *)h;
    PcmSink sink = {pcm, 0};
    g_sink = &sink;
    switch (rx->mode) {
    case 1: rx->am->acceptIqData((int8_t *)iq, (uint32_t)nbytes); break;
    case 2: rx->fm->acceptIqData((int8_t *)iq, (uint32_t)nbytes); break;
    case 3: rx->wbfm->acceptIqData((int8_t *)iq, (uint32_t)nbytes); break;
    case 4:
    case 5: rx->ssb->acceptIqData((int8_t *)iq, (uint32_t)nbytes); break;
    }
    g_sink = nullptr;
    return sink.count;
}

/* IqDataProcessor::setSignalDetectThreshold (IqDataProcessor.cc:392-405) */
void ref_rx_set_squelch_threshold(void *h, int32_t threshold) { ((RefRx *)h)->iqdp->setSignalDetectThreshold(threshold); }
/* the global Squelch::run is given (Radio.cc:413 sets 16); shared by every IqDataProcessor of the process */
void ref_set_rx_gain_db(uint32_t gain_db) { radio_adjustableReceiveGainInDb = gain_db; }

static void magnitude_cb(uint32_t magnitude, void *ctx) { *(uint32_t *)ctx = magnitude; }
static void state_cb(bool present, void *ctx) { *(uint8_t *)ctx = present ? 1 : 0; }

/* like ref_rx_run_2048k, and reports per block what the reference's own notification callbacks deliver:
 * the average magnitude (registerSignalMagnitudeCallback) and the squelch decision
 * (registerSignalStateCallback) of IqDataProcessor.cc:961-988 */
size_t ref_rx_run_2048k_squelch(void *h, const int8_t *iq, size_t nbytes, size_t block, int16_t *pcm,
                                uint32_t *magnitudes, uint8_t *allowed)
{
    RefRx *rx = (RefRx *)h;
    uint32_t mag = 0;
    uint8_t open = 0;
    rx->iqdp->registerSignalMagnitudeCallback(magnitude_cb, &mag);
    rx->iqdp->registerSignalStateCallback(state_cb, &open);
    rx->iqdp->enableSignalMagnitudeNotification();
    rx->iqdp->enableSignalNotification();
    size_t total = 0, b = 0;
    for (size_t off = 0; off < nbytes; off += block, b++) {
        size_t n = nbytes - off < block ? nbytes - off : block;
        total += ref_rx_accept_2048k(h, iq + off, n, pcm + total);
        magnitudes[b] = mag;
        allowed[b] = open;
    }
    rx->iqdp->disableSignalMagnitudeNotification();
    rx->iqdp->disableSignalNotification();
    rx->iqdp->registerSignalMagnitudeCallback(NULL, NULL);
    rx->iqdp->registerSignalStateCallback(NULL, NULL);
    return total;
}

/* ---- BasebandDataProcessor's PCM ring, event by event (BasebandDataProcessor.cc:416-433, 482-605) ---- */
void *ref_bbp_new(void)
{
    BasebandDataProcessor *b = new BasebandDataProcessor();
    /* the reference never initialises pcmBuffer; slots read before they are written hold heap garbage there.
     * Zero them so that traces are reproducible (hrd_pcm_ring_create zeroes its rings). */
    memset(b->pcmBuffer, 0, sizeof b->pcmBuffer);
    return b;
}
void ref_bbp_free(void *h)
{
    BasebandDataProcessor *b = (BasebandDataProcessor *)h;
    b->streamState = BasebandDataProcessor::Idle; /* no reader thread was started: stop() must not join one */
    delete b;
}
void ref_bbp_run(void *h, int running)
{
    ((BasebandDataProcessor *)h)->streamState = running ? BasebandDataProcessor::Running : BasebandDataProcessor::Idle;
}
/* the reader thread's step: returns the slot written */
int ref_bbp_write(void *h, const int16_t *pcm512)
{
    BasebandDataProcessor *b = (BasebandDataProcessor *)h;
    int16_t *p = b->getNextUnfilledBuffer();
    memcpy(p, pcm512, PCM_BLOCK_SIZE * sizeof(int16_t));
    return (int)((p - &b->pcmBuffer[0][0]) / PCM_BLOCK_SIZE);
}
/* the transmit callback's step: returns the slot sent (-1 = the zero buffer) and copies the block */
int ref_bbp_read(void *h, int16_t *out512)
{
    BasebandDataProcessor *b = (BasebandDataProcessor *)h;
    int16_t *p = b->getNextFilledBuffer();
    memcpy(out512, p, PCM_BLOCK_SIZE * sizeof(int16_t));
    return p == b->zeroPcmBuffer ? -1 : (int)((p - &b->pcmBuffer[0][0]) / PCM_BLOCK_SIZE);
}
void ref_bbp_stats(void *h, uint32_t out[4])
{
    BasebandDataProcessor *b = (BasebandDataProcessor *)h;
    out[0] = b->buffersProduced, out[1] = b->buffersConsumed, out[2] = b->pcmBlocksDropped, out[3] = b->pcmBlocksAdded;
}

/* stream a long buffer through in reference-sized blocks */
size_t ref_rx_run_2048k(void *h, const int8_t *iq, size_t nbytes, size_t block, int16_t *pcm)
{
    size_t total = 0;
    for (size_t off = 0; off < nbytes; off += block) {
        size_t n = nbytes - off < block ? nbytes - off : block;
        total += ref_rx_accept_2048k(h, iq + off, n, pcm + total);
    }
    return total;
}

size_t ref_rx_run_256k(void *h, const int8_t *iq, size_t nbytes, size_t block, int16_t *pcm)
{
    size_t total = 0;
    for (size_t off = 0; off < nbytes; off += block) {
        size_t n = nbytes - off < block ? nbytes - off : block;
        total += ref_rx_accept_256k(h, iq + off, n, pcm + total);
    }
    return total;
}

void *ref_tx_new(void)
{
    RefTx *tx = new RefTx;
    tx->am = new AmModulator();
    tx->fm = new FmModulator();
    tx->wbfm = new WbFmModulator();
    tx->ssb = new SsbModulator();
    return tx;
}

void ref_tx_free(void *h)
{
    RefTx *tx = (RefTx *)h;
    delete tx->am;
    delete tx->fm;
    delete tx->wbfm;
    delete tx->ssb;
    delete tx;
}

void ref_tx_set_am_index(void *h, float m) { ((RefTx *)h)->am->setModulationIndex(m); }
void ref_tx_set_fm_deviation(void *h, float d) { ((RefTx *)h)->fm->setFrequencyDeviation(d); }
void ref_tx_set_wbfm_deviation(void *h, float d) { ((RefTx *)h)->wbfm->setFrequencyDeviation(d); }

void ref_tx_reset_mod(void *h, int mod)
{
    RefTx *tx = (RefTx *)h;
    switch (mod) {
    case 0: tx->am->resetModulator(); break;
    case 1: tx->fm->resetModulator(); break;
    case 2: tx->wbfm->resetModulator(); break;
    case 3: tx->ssb->resetModulator(); break;
    }
}

/* <X>Modulator::acceptData in blocks of <= 512 PCM samples */
size_t ref_tx_accept(void *h, int mode, const int16_t *pcm, size_t n, int8_t *iq)
{
    RefTx *tx = (RefTx *)h;
    size_t total = 0;
    if (mode == 4) tx->ssb->setLsbModulationMode();
    if (mode == 5) tx->ssb->setUsbModulationMode();
    for (size_t off = 0; off < n; off += 512) {
        uint32_t cnt = (uint32_t)(n - off < 512 ? n - off : 512);
        uint32_t outBytes = 0;
        int16_t *src = (int16_t *)pcm + off;
        int8_t *dst = iq + total;
        switch (mode) {
        case 1: tx->am->acceptData(src, cnt, dst, &outBytes); break;
        case 2: tx->fm->acceptData(src, cnt, dst, &outBytes); break;
        case 3: tx->wbfm->acceptData(src, cnt, dst, &outBytes); break;
        case 4:
        case 5: tx->ssb->acceptData(src, cnt, dst, &outBytes); break;
        default: memset(dst, 0, (size_t)cnt * 512); outBytes = cnt * 512;
        }
        total += (size_t)cnt * 512;
        (void)outBytes;
    }
    return total;
}

/* ---- table dumps from the live reference objects ---------------------- */
int ref_dump_decimator_taps(void *dec, int16_t *out, int cap)
{
    Decimator_int16 *d = (Decimator_int16 *)dec;
    for (int i = 0; i < d->filterLength && i < cap; i++) out[i] = d->coefficientStoragePtr[i];
    return d->filterLength;
}

/* which: same enum as HRO_TAPS_* in hrd_oracle.h */
int ref_taps(void *hrx, void *htx, int which, int16_t *out, int cap)
{
    RefRx *rx = (RefRx *)hrx;
    RefTx *tx = (RefTx *)htx;
    switch (which) {
    case 0: return ref_dump_decimator_taps(rx->iqdp->stage1IDecimatorPtr, out, cap);
    case 1: return ref_dump_decimator_taps(rx->iqdp->stage2IDecimatorPtr, out, cap);
    case 2: return ref_dump_decimator_taps(rx->iqdp->stage3IDecimatorPtr, out, cap);
    case 3: return ref_dump_decimator_taps(rx->am->stage1IDecimatorPtr, out, cap);
    case 4: return ref_dump_decimator_taps(rx->am->stage2IDecimatorPtr, out, cap);
    case 5: return ref_dump_decimator_taps(rx->am->stage3IDecimatorPtr, out, cap);
    case 6: return ref_dump_decimator_taps(rx->fm->iTunerDecimatorPtr, out, cap);
    case 7: return ref_dump_decimator_taps(rx->fm->postDemodDecimatorPtr, out, cap);
    case 8: return ref_dump_decimator_taps(rx->fm->audioDecimatorPtr, out, cap);
    case 9: return ref_dump_decimator_taps(rx->wbfm->postDemodDecimator1Ptr, out, cap);
    case 10: {
        FirFilter_int16 *f = rx->ssb->delayLinePtr;
        for (int i = 0; i < f->filterLength && i < cap; i++) out[i] = f->coefficientStoragePtr[i];
        return f->filterLength;
    }
    case 11: {
        FirFilter_int16 *f = rx->ssb->phaseShifterPtr;
        for (int i = 0; i < f->filterLength && i < cap; i++) out[i] = f->coefficientStoragePtr[i];
        return f->filterLength;
    }
    case 12: {
        /* polyphase storage is p0 then p1 (Interpolator_int16.cc:311-322);
         * undo it so the dump is in prototype order */
        Interpolator_int16 *p = tx->am->iInterpolator2Ptr;
        int plen = p->polyphaseFilterLength, l = p->interpolationFactor;
        for (int i = 0; i < l; i++)
            for (int j = 0; j < plen; j++)
                if (i + j * l < cap) out[i + j * l] = p->coefficientStoragePtr[i * plen + j];
        return plen * l;
    }
    }
    return -1;
}

void ref_nco_tables(void *htx, float *sin16384, float *cos16384)
{
    RefTx *tx = (RefTx *)htx;
    memcpy(sin16384, tx->wbfm->ncoPtr->Sin, 16384 * sizeof(float));
    memcpy(cos16384, tx->wbfm->ncoPtr->Cos, 16384 * sizeof(float));
}

/* ---- CPU baseline: the reference chain, one instance per stream, streams
 *      round-robined over n_threads host threads, inputs resident in RAM.
 *      Returns wall seconds (slowest worker). ------------------------------ */
double ref_bench_rx(int mode, const int8_t *iq, size_t bytes_per_stream, size_t stride,
                    int n_streams, int n_threads, int16_t *pcm, size_t pcm_stride)
{
    std::vector<void *> rx((size_t)n_streams);
    for (int s = 0; s < n_streams; s++) {
        rx[(size_t)s] = ref_rx_new();
        ref_rx_set_mode(rx[(size_t)s], mode);
    }
    std::vector<int16_t> scratch((size_t)n_threads * (bytes_per_stream / 512 + 1024));
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; t++)
        pool.emplace_back([&, t]() {
            int16_t *local = scratch.data() + (size_t)t * (bytes_per_stream / 512 + 1024);
            for (int s = t; s < n_streams; s += n_threads) {
                int16_t *dst = pcm ? pcm + (size_t)s * pcm_stride : local;
                ref_rx_run_2048k(rx[(size_t)s], iq + (size_t)s * stride, bytes_per_stream, 262144, dst);
            }
        });
    for (auto &th : pool) th.join();
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (int s = 0; s < n_streams; s++) ref_rx_free(rx[(size_t)s]);
    return dt;
}

double ref_bench_tx(int mode, const int16_t *pcm, size_t n_per_stream, size_t stride,
                    int n_streams, int n_threads, int8_t *iq, size_t iq_stride)
{
    std::vector<void *> tx((size_t)n_streams);
    for (int s = 0; s < n_streams; s++) tx[(size_t)s] = ref_tx_new();
    std::vector<int8_t> scratch(iq ? 0 : (size_t)n_threads * n_per_stream * 512);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; t++)
        pool.emplace_back([&, t]() {
            for (int s = t; s < n_streams; s += n_threads) {
                int8_t *dst = iq ? iq + (size_t)s * iq_stride
                                 : scratch.data() + (size_t)t * n_per_stream * 512;
                ref_tx_accept(tx[(size_t)s], mode, pcm + (size_t)s * stride, n_per_stream, dst);
            }
        });
    for (auto &th : pool) th.join();
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (int s = 0; s < n_streams; s++) ref_tx_free(tx[(size_t)s]);
    return dt;
}

} // extern "C"
