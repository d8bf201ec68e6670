// hrd_squelch.cu -- the squelch gate and signal magnitude of the receive path.
//
// Replaces, per stream and per block (one reference call of IqDataProcessor::acceptIqData, reference
// paths relative to radioDiags/src_diags/):
//   SignalDetector::detectSignal  SignalDetector.cc:205-273   mean of max(|I|,|Q|) + min(|I|,|Q|)/2 over the
//                                                             block's 256 kS/s samples, integer dBFS, threshold
//   DbfsCalculator                DbfsCalculator.cc:36-68,111-147  the dB table (built on the host, hrd_api.cu)
//   SignalTracker::run            SignalTracker.cc:104-145    two states; a block after a signal still passes
//   Squelch::run                  Squelch.cc:227-273          decision = START | PRESENT | END (the tail)
// and the gate of IqDataProcessor.cc:991: the demodulator is simply not called for a closed block, so its
// state does not move and no PCM comes out.  Everything is integer arithmetic: bit-exact.
//
// Three small kernels around the demodulator launches (hrd_api.cu rx_squelched):
//   squelch_magnitude_kernel  one CTA per (stream, block): the block's average magnitude
//   squelch_track_kernel      one thread per stream: dBFS, threshold, tracker over the call's blocks in order
//   squelch_scatter_kernel    PCM of the blocks that were let through -> the caller's rows, packed
#include "hrd_device.cuh"

namespace hrd {

namespace {

__constant__ int32_t c_db_table[257];

// DETECT: bytes of one block = 2 * samples; words hold two I,Q samples {I0,Q0,I1,Q1}
__global__ void __launch_bounds__(256) squelch_magnitude_kernel(const int8_t *iq256, size_t stride, uint32_t block_bytes,
                                                                uint32_t total_bytes, int n_blocks, uint32_t *magnitude)
{
    const int stream = blockIdx.x, blk = blockIdx.y;
    const uint32_t begin = (uint32_t)blk * block_bytes;
    const uint32_t bytes = min(block_bytes, total_bytes - begin);
    const uint32_t *src = reinterpret_cast<const uint32_t *>(iq256 + (size_t)stream * stride + begin);
    uint32_t sum = 0;
    for (uint32_t w = threadIdx.x; w < bytes / 4; w += blockDim.x) {
        // iMagnitude = abs(I), qMagnitude = abs(Q) as uint8 (abs(-128) = 128); the larger plus half the smaller
        const uint32_t a = __vabs4(__ldg(src + w));
        const uint32_t i0 = a & 0xffu, q0 = (a >> 8) & 0xffu, i1 = (a >> 16) & 0xffu, q1 = a >> 24;
        sum += max(i0, q0) + (min(i0, q0) >> 1);
        sum += max(i1, q1) + (min(i1, q1) >> 1);
    }
    __shared__ uint32_t part[8];
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(HRD_FULL_MASK, sum, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t total = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) total += part[i];
        magnitude[(size_t)stream * n_blocks + blk] = total / (bytes / 2); // magnitude /= magnitudeBufferLength
    }
}

__global__ void squelch_track_kernel(const uint32_t *magnitude, int n_streams, int n_blocks, const float *threshold,
                                     const float *gain_db, uint8_t *tracking, uint8_t *allowed)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    const int32_t thr = (int32_t)threshold[s];
    const uint32_t gain = (uint32_t)gain_db[s];
    bool track = tracking[s] != 0;
    for (int b = 0; b < n_blocks; b++) {
        // DbfsCalculator::convertMagnitudeToDbFs with wordLengthInBits = 7: clip to 127, table, minus 42
        const uint32_t m = min(magnitude[(size_t)s * n_blocks + b], 127u);
        int32_t dbfs = c_db_table[m] - 42;
        dbfs -= gain; // int32 -= uint32, as SignalDetector.cc:263 writes it
        const bool present = dbfs >= thr;
        // NoSignal: present -> START (allowed), else NOISE (closed); Tracking: PRESENT, or END = the tail (allowed)
        allowed[(size_t)s * n_blocks + b] = (track || present) ? 1 : 0;
        track = present;
    }
    tracking[s] = track ? 1 : 0;
}

// one CTA per (block-of-this-launch, stream): copies n samples of an open block to the caller's row
__global__ void squelch_scatter_kernel(const int16_t *scratch, size_t scratch_stride, int16_t *pcm, size_t pcm_stride,
                                       const uint32_t *out_at, const uint8_t *allowed, const uint8_t *kind_of, int n_blocks,
                                       int blk, uint32_t n)
{
    const int s = blockIdx.x;
    if (kind_of[s] == K_NONE || !allowed[(size_t)s * n_blocks + blk]) return;
    const int16_t *src = scratch + (size_t)s * scratch_stride;
    int16_t *dst = pcm + (size_t)s * pcm_stride + out_at[(size_t)s * n_blocks + blk];
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

// IqDataProcessor::upconvertByFsOver4 / downconvertByFsOver4 on their own (IqDataProcessor.cc:771-815, 715-759):
// groups of four I,Q samples times {1, j, -1, -j} (up) or {1, -j, -1, j} (down), int8 negation wrapping as
// the reference's does.  One thread per group of 8 bytes, in place.
__global__ void fs4_rotate_kernel(int8_t *iq, size_t n_groups, int up)
{
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    char4 *p = reinterpret_cast<char4 *>(iq + g * 8);
    char4 a = p[0], b = p[1]; // a = {I0,Q0,I1,Q1}, b = {I2,Q2,I3,Q3}
    const signed char x1 = a.z, y1 = a.w, x3 = b.z, y3 = b.w;
    if (up) {
        a.z = (signed char)-y1, a.w = x1;
        b.z = y3, b.w = (signed char)-x3;
    } else {
        a.z = y1, a.w = (signed char)-x1;
        b.z = (signed char)-y3, b.w = x3;
    }
    b.x = (signed char)-b.x, b.y = (signed char)-b.y;
    p[0] = a, p[1] = b;
}

} // namespace

int launch_fs4_rotate(int8_t *iq, size_t n_groups, int up, cudaStream_t s)
{
    if (!n_groups) return 0;
    fs4_rotate_kernel<<<(unsigned)((n_groups + 255) / 256), 256, 0, s>>>(iq, n_groups, up);
    return (int)cudaGetLastError();
}

void upload_db_table(const int32_t *table257) { cudaMemcpyToSymbol(c_db_table, table257, 257 * sizeof(int32_t)); }

int launch_squelch_magnitude(const int8_t *iq256, size_t stride, uint32_t block_bytes, uint32_t total_bytes, int n_streams,
                             int n_blocks, uint32_t *magnitude, cudaStream_t s)
{
    dim3 grid((unsigned)n_streams, (unsigned)n_blocks); // at most 65535 blocks = 70 minutes per call
    squelch_magnitude_kernel<<<grid, 256, 0, s>>>(iq256, stride, block_bytes, total_bytes, n_blocks, magnitude);
    return (int)cudaGetLastError();
}

int launch_squelch_track(const uint32_t *magnitude, int n_streams, int n_blocks, const float *threshold, const float *gain_db,
                         uint8_t *tracking, uint8_t *allowed, cudaStream_t s)
{
    squelch_track_kernel<<<(n_streams + 127) / 128, 128, 0, s>>>(magnitude, n_streams, n_blocks, threshold, gain_db, tracking,
                                                                allowed);
    return (int)cudaGetLastError();
}

int launch_squelch_scatter(const int16_t *scratch, size_t scratch_stride, int16_t *pcm, size_t pcm_stride, const uint32_t *out_at,
                           const uint8_t *allowed, const uint8_t *kind_of, int n_streams, int n_blocks, int blk, uint32_t n,
                           cudaStream_t s)
{
    squelch_scatter_kernel<<<n_streams, 128, 0, s>>>(scratch, scratch_stride, pcm, pcm_stride, out_at, allowed, kind_of, n_blocks,
                                                    blk, n);
    return (int)cudaGetLastError();
}

} // namespace hrd
