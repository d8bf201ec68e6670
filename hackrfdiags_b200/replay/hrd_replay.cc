// hrd_replay.cc -- the reference's file formats through the batched library.
//
// The reference moves three raw formats around the hot path (radioDiags/README.txt:107-126,
// src_diags/DataProvider.cc, radioApp.cc:103-111, IqDataProcessor.cc:953-957):
//   *.iq   int8 interleaved I,Q at 2.048 MS/s   what DataProvider loops into the transmitter and what
//                                               hackrf_transfer records; the modulator test programs
//                                               (AmModulator/am.cc:31-68 ...) write it to stdout
//   *.pcm  S16_LE at 8 kS/s                     what radioApp writes to stdout (aplay -f S16_LE -r 8000)
//                                               and what the modulator programs read from stdin
//   256 kS/s int8 I,Q                           the decimated stream the reference dumps over UDP
// This program replays MANY such files at once: every file is one stream of one batch, fed in blocks of
// 262144 input bytes (the HackRF transfer size, hackRf/hackrf.c:101 -- one IqDataProcessor::acceptIqData
// call each) or 512 PCM samples (BasebandDataProcessor.h:16), exactly as the reference's threads would.
// Files may have different lengths: a round is as long as the shortest stream still running allows, and a
// stream that has ended drops out (a tail shorter than one PCM sample = 512 bytes is ignored, as the
// reference's readers ignore a short last read).
//
//   hrd_replay rx <am|fm|wbfm|lsb|usb> [-s squelch_dBFS] [-g demod_gain] <out_dir> <file.iq>...   -> <out_dir>/<name>.pcm
//   hrd_replay fe <out_dir> <file.iq>...                                                           -> <out_dir>/<name>.iq256k
//   hrd_replay tx <am|fm|wbfm|lsb|usb|dsb|pm|amproto|fmproto> [-p index_or_deviation] <out_dir> <file.pcm>...      -> <out_dir>/<name>.iq
// -l <blocks> (rx, fe): LOOP replay, DataProvider's way (src_diags/DataProvider.cc:230-286 loadIqFile, :163-212
// retrieveIqDataFromBuffer): every file is read whole into a ring and handed out 262144 bytes at a time modulo its
// length -- a block may straddle the end of the file, files of any (even odd) length keep their own phase -- for
// exactly <blocks> rounds.
// There is no CPU fallback: without a B200 hrd_create fails and so does this program.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "hrd.h"

static void die(const char *what)
{
    fprintf(stderr, "hrd_replay: %s: %s\n", what, hrd_last_error());
    exit(1);
}

static int mode_of(const char *name, bool tx)
{
    static const struct { const char *n; int m; bool tx_only; } table[] = {
        {"am", HRD_MODE_AM, false}, {"fm", HRD_MODE_FM, false}, {"wbfm", HRD_MODE_WBFM, false}, {"lsb", HRD_MODE_LSB, false},
        {"usb", HRD_MODE_USB, false}, {"dsb", HRD_MODE_DSB, true}, {"pm", HRD_MODE_PM, true},
        {"amproto", HRD_MODE_AM_PROTO, true}, {"fmproto", HRD_MODE_FM_PROTO, true}};
    for (const auto &e : table)
        if (!strcmp(name, e.n) && (tx || !e.tx_only)) return e.m;
    fprintf(stderr, "hrd_replay: unknown mode '%s'\n", name);
    exit(2);
}

static std::string out_name(const std::string &dir, const char *path, const char *ext)
{
    std::string base(path);
    const size_t slash = base.find_last_of('/');
    if (slash != std::string::npos) base = base.substr(slash + 1);
    const size_t dot = base.find_last_of('.');
    if (dot != std::string::npos) base = base.substr(0, dot);
    return dir + "/" + base + ext;
}

struct Stream {
    FILE *in = nullptr, *out = nullptr;
    std::vector<char> ring; // -l: the whole file (iqSampleBufferPtr)
    size_t ring_at = 0;     //     iqSampleBufferIndex
    long left = 0; // input units (bytes for rx/fe, PCM samples for tx) still to read
    bool live = true;
};

int main(int argc, char **argv)
{
    if (argc < 4) {
        fprintf(stderr, "usage: hrd_replay rx|fe|tx [mode] [options] <out_dir> <files...>  (see the header of hrd_replay.cc)\n");
        return 2;
    }
    const std::string what = argv[1];
    const bool tx = what == "tx", fe = what == "fe";
    if (!tx && !fe && what != "rx") return fprintf(stderr, "hrd_replay: rx, fe or tx\n"), 2;
    int a = 2;
    const int mode = fe ? HRD_MODE_NONE : mode_of(argv[a++], tx);
    float squelch = -200.f, gain = 0.f, param = 0.f;
    bool have_gain = false, have_param = false;
    long loop_blocks = 0;
    while (a + 1 < argc && argv[a][0] == '-' && argv[a][1] && !argv[a][2]) {
        const char opt = argv[a][1];
        const float v = (float)atof(argv[a + 1]);
        if (opt == 's') squelch = v;
        else if (opt == 'g') gain = v, have_gain = true;
        else if (opt == 'p') param = v, have_param = true;
        else if (opt == 'l') loop_blocks = atol(argv[a + 1]);
        else return fprintf(stderr, "hrd_replay: unknown option -%c\n", opt), 2;
        a += 2;
    }
    if (argc - a < 2) return fprintf(stderr, "hrd_replay: need an output directory and at least one file\n"), 2;
    if (loop_blocks < 0 || (loop_blocks && tx)) return fprintf(stderr, "hrd_replay: -l takes a block count, with rx or fe\n"), 2;
    const std::string dir = argv[a++];
    const int n = argc - a;
    const long unit = tx ? 1 : 512;              // a whole PCM sample
    const long block = tx ? 512 : 262144;        // one reference call
    const size_t in_elem = tx ? sizeof(int16_t) : 1;

    std::vector<Stream> st((size_t)n);
    for (int i = 0; i < n; i++) {
        st[(size_t)i].in = fopen(argv[a + i], "rb");
        if (!st[(size_t)i].in) return perror(argv[a + i]), 1;
        fseek(st[(size_t)i].in, 0, SEEK_END);
        st[(size_t)i].left = ftell(st[(size_t)i].in) / (long)in_elem / unit * unit;
        fseek(st[(size_t)i].in, 0, SEEK_SET);
        if (loop_blocks) { // DataProvider::loadIqFile: the whole file, every byte of it
            Stream &s = st[(size_t)i];
            fseek(s.in, 0, SEEK_END);
            s.ring.resize((size_t)ftell(s.in));
            fseek(s.in, 0, SEEK_SET);
            if (s.ring.empty() || fread(s.ring.data(), 1, s.ring.size(), s.in) != s.ring.size())
                return fprintf(stderr, "hrd_replay: cannot load %s\n", argv[a + i]), 1;
            s.left = block; // never runs out
        }
        const std::string o = out_name(dir, argv[a + i], tx ? ".iq" : (fe ? ".iq256k" : ".pcm"));
        st[(size_t)i].out = fopen(o.c_str(), "wb");
        if (!st[(size_t)i].out) return perror(o.c_str()), 1;
    }

    hrd_batch_t *b = nullptr;
    const char *dev = getenv("HRD_DEVICE");
    if (hrd_create(dev ? atoi(dev) : 0, n, tx ? HRD_TX : HRD_RX, &b)) die("hrd_create");
    if (!fe && hrd_set_mode(b, HRD_ALL_STREAMS, mode)) die("hrd_set_mode");
    if (!tx && !fe) {
        static const int gain_param[6] = {-1, HRD_PARAM_AM_GAIN, HRD_PARAM_FM_GAIN, HRD_PARAM_WBFM_GAIN, HRD_PARAM_SSB_GAIN, HRD_PARAM_SSB_GAIN};
        if (have_gain && hrd_set_param(b, HRD_ALL_STREAMS, gain_param[mode], gain)) die("hrd_set_param");
        if (hrd_set_param(b, HRD_ALL_STREAMS, HRD_PARAM_SQUELCH_THRESHOLD, squelch)) die("hrd_set_param");
    }
    if (tx && have_param) {
        const int p = mode == HRD_MODE_AM ? HRD_PARAM_AM_INDEX : mode == HRD_MODE_FM ? HRD_PARAM_FM_DEV : mode == HRD_MODE_WBFM ? HRD_PARAM_WBFM_DEV : -1;
        if (p >= 0 && hrd_set_param(b, HRD_ALL_STREAMS, p, param)) die("hrd_set_param");
    }

    const size_t in_stride = (size_t)block * in_elem;                 // bytes per stream row of a round
    const size_t out_units = tx ? (size_t)block * 512 : (fe ? (size_t)block / 8 : (size_t)block / 512);
    const size_t out_elem = tx || fe ? 1 : sizeof(int16_t);
    std::vector<char> in((size_t)n * in_stride), out((size_t)n * out_units * out_elem);
    std::vector<uint32_t> counts((size_t)n);
    unsigned long long total_in = 0, rounds = 0;
    for (;;) {
        long len = block;
        int running = 0;
        for (auto &s : st)
            if (s.live) {
                if (s.left < unit) s.live = false;
                else len = std::min(len, s.left), running++;
            }
        if (!running || (loop_blocks && (long)rounds >= loop_blocks)) break;
        for (int i = 0; i < n; i++) {
            Stream &s = st[(size_t)i];
            char *row = in.data() + (size_t)i * in_stride;
            if (loop_blocks) { // DataProvider::retrieveIqDataFromBuffer
                size_t todo = (size_t)len;
                while (todo) {
                    const size_t part = std::min(todo, s.ring.size() - s.ring_at);
                    memcpy(row + ((size_t)len - todo), s.ring.data() + s.ring_at, part);
                    todo -= part;
                    s.ring_at = (s.ring_at + part) % s.ring.size();
                }
            } else if (s.live) {
                if (fread(row, in_elem, (size_t)len, s.in) != (size_t)len) return fprintf(stderr, "hrd_replay: short read on %s\n", argv[a + i]), 1;
                s.left -= len;
            } else {
                memset(row, 0, (size_t)len * in_elem); // an ended stream idles; what it produces is not written
            }
        }
        int rc;
        if (tx)
            rc = hrd_tx_process(b, (const int16_t *)in.data(), (size_t)len, (size_t)block, (int8_t *)out.data(), out_units, HRD_MEM_HOST, nullptr);
        else if (fe)
            rc = hrd_rx_front_end(b, (const int8_t *)in.data(), (size_t)len, in_stride, (int8_t *)out.data(), out_units, HRD_MEM_HOST, nullptr);
        else
            rc = hrd_rx_process(b, (const int8_t *)in.data(), (size_t)len, in_stride, HRD_ENTRY_2048K, (int16_t *)out.data(), out_units, counts.data(), HRD_MEM_HOST, nullptr);
        if (rc) die("process");
        for (int i = 0; i < n; i++) {
            Stream &s = st[(size_t)i];
            if (!s.live) continue;
            const size_t produced = tx ? (size_t)len * 512 : (fe ? (size_t)len / 8 : counts[(size_t)i]); // squelch: fewer
            fwrite(out.data() + (size_t)i * out_units * out_elem, out_elem, produced, s.out);
        }
        total_in += (unsigned long long)len * (unsigned long long)running;
        rounds++;
    }
    for (auto &s : st) fclose(s.in), fclose(s.out);
    hrd_destroy(b);
    fprintf(stderr, "hrd_replay: %d stream(s), %llu round(s), %llu input %s\n", n, rounds, total_in, tx ? "PCM samples" : "bytes");
    return 0;
}
