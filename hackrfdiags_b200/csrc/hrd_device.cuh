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

__device__ __forceinline__ void stg_stream_256(void *p, const u32x8 &r)
{
    asm volatile("st.global.L1::no_allocate.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r.v[0]),
                 "r"(r.v[1]), "r"(r.v[2]), "r"(r.v[3]), "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7])
                 : "memory");
}

// ------------------------------------------------------------------------------------
// the reference's C++ casts, restated for the GPU
// ------------------------------------------------------------------------------------
// (int16_t)someFloat on x86-64: cvttss2si (0x80000000 when out of range / NaN), then the
// low 16 bits.  CUDA's cvt.rzi.s32.f32 saturates instead, so the out-of-range case is
// patched by hand (e.g. FmDemodulator.cc:565).
__device__ __forceinline__ int f32_to_i16(float x)
{
    int v = __float2int_rz(x);
    if (!(x >= -2147483648.0f && x < 2147483648.0f)) v = (int)0x80000000;
    return (int)(short)v;
}

// Q15 tail: (int16_t)(acc >> 15), acc already holds the 1<<14 rounding constant
__device__ __forceinline__ int q15(int acc) { return (int)(short)(acc >> 15); }

// wrap a float phase difference into [-pi, pi]: double compares, double subtraction
// narrowed to float (FmDemodulator.cc:511-519, WbFmDemodulator.cc:416-424)
__device__ __forceinline__ float wrap_pi(float d)
{
    const double pi = 3.14159265358979323846;
    while ((double)d > pi) d = (float)((double)d - 2.0 * pi);
    while ((double)d < -pi) d = (float)((double)d + 2.0 * pi);
    return d;
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
