// hrd_device.cuh -- shared device-side pieces of libhrd_b200 (sm_100a only).
//
// Execution model: ONE WARP OWNS ONE STREAM for the whole call.
// Lanes split the time axis inside a 64-sample (256 kS/s) iteration; stage hand-over
// between lanes is by warp shuffle, between rates by small per-warp shared-memory
// rings with the filter history kept in front of the new samples.  There is no
// block-level synchronisation anywhere; warps never talk to each other.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "hrd_tables.h"

#define HRD_FULL_MASK 0xffffffffu
#define HRD_WARPS_PER_CTA 4

namespace hrd {

// ------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------
// dp2a: two 16-bit values of `a` (unsigned or signed) times two signed bytes of `b`
// (.lo = bytes 0,1  .hi = bytes 2,3), accumulated on c.  IDP.2A on sm_100a, same issue
// rate as IMAD (tools/ubench/pipes.cu: 64 lanes/clk/SM) but two MACs per instruction and
// it eats packed int8 samples directly.
__device__ __forceinline__ int dp2a_lo_us(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_us(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.hi.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_lo_ss(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_ss(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// 256-bit streaming accesses (LDG.E.256 / STG.E.256 on sm_100a): one full 32-byte
// sector per lane, 1 KiB contiguous per warp instruction.
struct __align__(32) u32x8 {
    uint32_t v[8];
};

__device__ __forceinline__ u32x8 ldg_stream_256(const void *p)
{
    u32x8 r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
                   "=r"(r.v[6]), "=r"(r.v[7])
                 : "l"(p));
    return r;
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ void stg_stream_256(void *p, const u32x8 &r)
{
    asm volatile("st.global.L1::no_allocate.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r.v[0]),
                 "r"(r.v[1]), "r"(r.v[2]), "r"(r.v[3]), "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7])
                 : "memory");
}

// cp.async (LDGSTS): global -> shared copies that need no registers and complete in order per thread,
// so wait_group N is exact ("all but the N newest groups have landed").
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Bulk asynchronous copy (the TMA engine, 1-D form): ONE instruction moves a whole chunk global -> shared
// and signals an mbarrier with the byte count; no registers, no per-lane requests, and every waiter sees
// the data once its try_wait on the barrier's phase succeeds.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 "HRD_WAIT:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra HRD_DONE;\n\t"
                 "bra HRD_WAIT;\n\t"
                 "HRD_DONE:\n\t"
                 "}" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}

// Named barriers (bar.sync / bar.arrive with an id and a thread count): the producer / consumer
// hand-over of the chain-warp kernels.  Threads that ARRIVE do not wait; shared-memory writes made
// before the arrive are visible to the threads the barrier releases.  Id 0 is __syncthreads.
#define HRD_BAR_PRODUCED 1 // and 2: item warps -> chain warp, by step parity
#define HRD_BAR_CHAINED 3  // and 4: chain warp -> item warps
// (the id is an immediate chosen by the step's parity: a register id makes ptxas reserve all 16 barriers)
__device__ __forceinline__ void named_bar_sync(int base, uint32_t parity, int threads)
{
    if (parity & 1)
        asm volatile("bar.sync %0, %1;" ::"r"(base + 1), "r"(threads) : "memory");
    else
        asm volatile("bar.sync %0, %1;" ::"r"(base), "r"(threads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int base, uint32_t parity, int threads)
{
    if (parity & 1)
        asm volatile("bar.arrive %0, %1;" ::"r"(base + 1), "r"(threads) : "memory");
    else
        asm volatile("bar.arrive %0, %1;" ::"r"(base), "r"(threads) : "memory");
}

// the same with three ids per direction (rx_wbfm_kernel with three row buffers: ids base .. base + 2 by step mod 3)
#define HRD_BAR3_PRODUCED 1
#define HRD_BAR3_CHAINED 4
__device__ __forceinline__ void named_bar_sync3(int base, uint32_t phase, int threads)
{
    if (phase == 0)
        asm volatile("bar.sync %0, %1;" ::"r"(base), "r"(threads) : "memory");
    else if (phase == 1)
        asm volatile("bar.sync %0, %1;" ::"r"(base + 1), "r"(threads) : "memory");
    else
        asm volatile("bar.sync %0, %1;" ::"r"(base + 2), "r"(threads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive3(int base, uint32_t phase, int threads)
{
    if (phase == 0)
        asm volatile("bar.arrive %0, %1;" ::"r"(base), "r"(threads) : "memory");
    else if (phase == 1)
        asm volatile("bar.arrive %0, %1;" ::"r"(base + 1), "r"(threads) : "memory");
    else
        asm volatile("bar.arrive %0, %1;" ::"r"(base + 2), "r"(threads) : "memory");
}

// ------------------------------------------------------------------------------------
// the reference's C++ casts, restated for the GPU
// ------------------------------------------------------------------------------------
// (int16_t)someFloat on x86-64: cvttss2si (0x80000000 when out of range / NaN), then the
// low 16 bits.  CUDA's cvt.rzi.s32.f32 saturates instead, so the out-of-range case is
// patched by hand (e.g. FmDemodulator.cc:565).
// F2I saturates; of the values it can return only +overflow (0x7fffffff, low half 0xffff) differs
// in the low 16 bits from cvttss2si's 0x80000000, and NaN gives 0 on both: one compare suffices.
__device__ __forceinline__ int f32_to_i16(float x)
{
    int v = __float2int_rz(x);
    if (!(x < 2147483648.0f)) v = 0;
    return (int)(short)v;
}

// Q15 tail: (int16_t)(acc >> 15), acc already holds the 1<<14 rounding constant
__device__ __forceinline__ int q15(int acc) { return (int)(short)(acc >> 15); }

// ------------------------------------------------------------------------------------
// phase wrapping without double arithmetic on the hot path
// ------------------------------------------------------------------------------------
// The reference wraps float phases with double constants: `while (x > M_PI) x -= 2*M_PI;`
// i.e. compare (double)x with M_PI, subtract in double, narrow to float
// (FmDemodulator.cc:511-519, WbFmDemodulator.cc:416-424, PhaseAccumulator.cc:165-177).
// On sm_100a F2F.F64.F32 runs at 0.47 warp-instructions/clk/SM (tools/ubench/pipes.cu), so:
//   * (double)x >  M_PI  <=>  x >=  HRD_PI_UP   (the float just above pi; the one below is < pi)
//     (double)x < -M_PI  <=>  x <= -HRD_PI_UP
//   * for |x| in [HRD_PI_UP, 12):  (float)((double)|x| - 2*M_PI) == (|x| - 2PI_HI) - 2PI_LO in
//     fp32, where 2PI_HI = fl32(2*M_PI) and 2PI_LO = fl32(2*M_PI - 2PI_HI): the first
//     subtraction is exact (Sterbenz), the second rounds once, and the rounding agrees with the
//     double expression whenever | |x| - 2PI_HI | >= 2^-10.  Otherwise (4095 floats next to
//     2*pi, or |x| >= 12) the double expression itself is evaluated.
// tools/verify_fp_tricks.c checks both statements over every float in range (0 mismatches).
#define HRD_PI_UP 3.14159274101257324219f
#define HRD_2PI_HI 6.28318548202514648438f
#define HRD_2PI_LO (-1.74845553146951715e-07f)

// one wrap step toward zero of an x with |x| >= HRD_PI_UP: (float)((double)x -+ 2*M_PI)
__device__ __forceinline__ float wrap_2pi_once(float x)
{
    const float s = fabsf(x);
    const float t = __fsub_rn(s, HRD_2PI_HI);
    float r = __fsub_rn(t, HRD_2PI_LO);
    if (!(s < 12.0f) || fabsf(t) < 0x1p-10f) r = (float)((double)s - 2.0 * 3.14159265358979323846);
    // x > 0: r;  x < 0: -(r)   (IEEE rounding is symmetric)
    return __int_as_float(__float_as_int(r) ^ (__float_as_int(x) & (int)0x80000000));
}

// wrap a float phase (difference) into [-pi, pi] exactly as the reference's two while loops do
__device__ __forceinline__ float wrap_pi(float d)
{
    while (fabsf(d) >= HRD_PI_UP) d = wrap_2pi_once(d);
    return d;
}

// The same for the discriminators, whose argument is a difference of two atan2 TABLE values.
// The table holds atan2(q, i) for integer q, i in [-128, 127]: its largest entry is +pi (q = 0,
// i < 0) and its smallest is atan2(-1, -128) = -3.13378 (q is never -0.0), so
//   |d| <= pi + 3.13378 = 6.27537 < 2*pi - 2^-10 :
// one wrap at most, the fp32 wrap is always the exact one (hrd_device.cuh above) and no fallback
// is needed (tests/test_capi_host.py checks the bound on the table itself).  Loops and divergent
// branches cost more than the arithmetic here, so the wrap is computed on all lanes and selected:
// two FADDs, a LOP3, a compare and a select.
__device__ __forceinline__ float wrap_pi_select(float d)
{
    const float s = fabsf(d);
    const float t = __fsub_rn(s, HRD_2PI_HI);
    const float r = __fsub_rn(t, HRD_2PI_LO);
    const float w = __int_as_float(__float_as_int(r) ^ (__float_as_int(d) & (int)0x80000000));
    return (s >= HRD_PI_UP) ? w : d;
}

// cosf / sinf as the reference's libm computes them (Nco::run, Nco/Nco.cc:186-199: cos(phase), sin(phase) of a float;
// signals/pm.cc:39-55, fm.cc:41-62).  glibc's routines (sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, sincosf.h, unchanged
// since 2.28: Szabolcs Nagy's double-precision polynomials) are not always correctly rounded, so "sin in double,
// rounded to float" -- what this library did in round 1 -- differs from them in the last bit now and then, and a last
// bit times 16000, truncated and interpolated, is the occasional 1 LSB of int8 IQ that FM transmit used to be allowed.
// This is that algorithm restated: range reduction x - n * (pi/2) with n = round(x * 2/pi) taken from the integer part
// of x * (2/pi * 2^24), then the odd / even polynomial of the quadrant, all in double, one rounding to float at the
// end.  tools/verify_sincosf.c evaluates the same restatement on the host for EVERY float with |x| < 8 (2.18e9 of
// them) against libm's sinf and cosf: 0 mismatches, with and without contraction of the a + b * c forms.  The phases
// that reach it are wrapped to +-pi (the prototype FM head: +-2 pi); beyond 8 (never) the double sincos is used.
__device__ __forceinline__ void glibc_sincosf(float y, float &sin_out, float &cos_out)
{
    const uint32_t top = (__float_as_uint(y) >> 20) & 0x7ffu; // abstop12
    if (!(top < 0x410u)) { // |y| >= 8, NaN: outside the verified range
        double sd, cd;
        sincos((double)y, &sd, &cd);
        sin_out = (float)sd, cos_out = (float)cd;
        return;
    }
    double x = (double)y;
    int n = 0;
    if (top >= 0x3f4u) { // abstop12(pi/4): reduce_fast
        const double r = __dmul_rn(x, 0x1.45F306DC9C883p+23);
        n = (__double2int_rz(r) + 0x800000) >> 24;
        x = __fma_rn(-(double)n, 0x1.921FB54442D18p0, x);
    } else if (top < 0x398u) { // |y| < 2^-12
        sin_out = y, cos_out = 1.0f;
        return;
    }
    const double x2 = __dmul_rn(x, x);
    // quadrant: sign[n & 3] = {1, -1, -1, 1} on the odd polynomial's argument; table 1 (n & 2) negates the even one
    const double xs = ((n + 1) & 2) ? -x : x;
    const double neg = (n & 2) ? -1.0 : 1.0;
    const double x3 = __dmul_rn(xs, x2), s1 = __fma_rn(x2, -0x1.994eb3774cf24p-13, 0x1.1107605230bc4p-7), x5 = __dmul_rn(x3, x2),
                 s = __fma_rn(x3, -0x1.555545995a603p-3, xs);
    const float odd = (float)__fma_rn(x5, s1, s);
    const double x4 = __dmul_rn(x2, x2), c2 = __fma_rn(x2, neg * 0x1.99343027bf8c3p-16, neg * -0x1.6c087e89a359dp-10),
                 c1 = __fma_rn(x2, neg * -0x1.ffffffd0c621cp-2, neg), x6 = __dmul_rn(x4, x2), c = __fma_rn(x4, neg * 0x1.55553e1068f19p-5, c1);
    const float even = (float)__fma_rn(x6, c2, c);
    // sinf uses the odd polynomial in even quadrants, cosf the other way round
    sin_out = (n & 1) ? even : odd;
    cos_out = (n & 1) ? odd : even;
}

// One step of PhaseAccumulator::run (Nco/PhaseAccumulator.cc:157-181) on the serial NCO chains of hrd_tx.cu, for
// |phase| < pi and |step| < 3 (so |phase + step| < 6.2 and at most one wrap, in the direction of the step):
// acc += step; then (float)((double)acc -+ 2*M_PI) when |acc| passed pi.
// The chain is the critical path of those kernels and its cost is LATENCY, so the form below has no predicate and no
// select in the dependency chain (measured on B200: a compare feeding a predicated FMA or an FSEL costs ~25 cycles
// per sample, as much as five dependent adds): four dependent FMA-pipe operations,
//   a = acc + step
//   k = sat((|a| - P_DN) * 2^30)            exactly 0.0 or 1.0: P_DN is the float just below pi, floats are >= 2^-22 apart
//   t = fma(k, -+2PI_HI, a)                 k = 1: exact (Sterbenz);      k = 0: a
//   r = fma(k, -+2PI_LO, t)                 k = 1: one rounding, the fp32 wrap proved equal to the double expression
//                                           for |t| >= 2^-10 (here |t| > 0.14; tools/verify_fp_tricks.c checks 1 and 1c)
// with the signs of the two constants taken from the step (off the chain).  (fma(0, c, -0.0) is +0.0: only the sign
// of a zero phase can differ from the reference, and nothing downstream reads it.)
#define HRD_PI_DN 3.14159250259399414062f /* 0x40490fda */
// the chain's three constants as values ptxas cannot see (ConstTables k_sign / k_m2pi_hi / k_m2pi_lo): kept in
// registers, "(step & sign) ^ constant" is ONE LOP3 per constant instead of an AND and two XORs with immediates
struct ChainConsts {
    uint32_t sign, m2pi_hi, m2pi_lo;
};
__device__ __forceinline__ uint32_t lop3_and_xor(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x6a;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); // (a & b) ^ c
    return d;
}
__device__ __forceinline__ float phase_step_fast(float phase, float step, const ChainConsts &cc)
{
    const float a = __fadd_rn(phase, step);
    const float hs = __uint_as_float(lop3_and_xor(__float_as_uint(step), cc.sign, cc.m2pi_hi)); // -2PI_HI for a step >= 0
    const float ls = __uint_as_float(lop3_and_xor(__float_as_uint(step), cc.sign, cc.m2pi_lo));
    const float k = __saturatef(__fmaf_rn(fabsf(a), 0x1p30f, -HRD_PI_DN * 0x1p30f));
    return __fmaf_rn(k, ls, __fmaf_rn(k, hs, a));
}

// ------------------------------------------------------------------------------------
// per-warp shared-memory rings: [hist old samples][new samples of this batch]
// ------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void ring_load_hist(T *ring, const T *state, int hist, int lane)
{
    for (int i = lane; i < hist; i += 32) ring[i] = state[i];
}

template <typename T>
__device__ __forceinline__ void ring_save_hist(const T *ring, T *state, int hist, int lane)
{
    for (int i = lane; i < hist; i += 32) state[i] = ring[i];
}

// move the last `hist` entries (after n_new were appended) to the front; hist <= 64
template <typename T>
__device__ __forceinline__ void ring_shift(T *ring, int hist, int n_new, int lane)
{
    T a = T(), b = T();
    if (lane < hist) a = ring[n_new + lane];
    if (lane + 32 < hist) b = ring[n_new + lane + 32];
    __syncwarp();
    if (lane < hist) ring[lane] = a;
    if (lane + 32 < hist) ring[lane + 32] = b;
    __syncwarp();
}

} // namespace hrd
