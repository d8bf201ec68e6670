// hrd_squelch.cu -- the Fs/4 rotations on their own.  (The squelch gate and signal magnitude of the receive path,
// SURVEY.md section 8f row 1, live in hrd_rx.cu: rx_gate_kernel fuses them into the front end.)
#include "hrd_device.cuh"

namespace hrd {

namespace {

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

} // namespace hrd
