// Instruction-throughput probe for the pipes the Rx/Tx kernels lean on (B200, sm_100a).
// Each kernel runs ITER x UNROLL independent ops per thread over 8 accumulator chains;
// reports warp-instructions / clk / SM so the kernel design (dp2a vs imad, prmt cost) is
// grounded in this part's numbers, not a guess.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITER 4096

template <int OP>
__global__ void __launch_bounds__(256) probe(int *out, int a0, int b0)
{
    int acc[CHAINS];
    int a = a0 + threadIdx.x, b = b0 ^ threadIdx.x;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) acc[c] = threadIdx.x * (c + 1);
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
            if (OP == 0) asm volatile("dp2a.lo.u32.s32 %0, %1, %2, %0;" : "+r"(acc[c]) : "r"(a), "r"(b));
            if (OP == 1) asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(acc[c]) : "r"(a), "r"(b));
            if (OP == 2) asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(acc[c]) : "r"(a), "r"(b));
            if (OP == 3) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(acc[c]) : "r"(a), "r"(b));
            if (OP == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(acc[c]) : "r"(a), "r"(b));
            if (OP == 5) asm volatile("shf.r.clamp.b32 %0, %0, %1, %2;" : "+r"(acc[c]) : "r"(a), "r"(b));
            if (OP == 6) asm volatile("add.s32 %0, %0, %1;" : "+r"(acc[c]) : "r"(a));
            if (OP == 7) { // mixed: dp2a + prmt alternating (fma pipe + alu pipe)
                if (c & 1) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(acc[c]) : "r"(a), "r"(b));
                else asm volatile("dp2a.lo.u32.s32 %0, %1, %2, %0;" : "+r"(acc[c]) : "r"(a), "r"(b));
            }
            if (OP == 8) { // mixed: imad + prmt
                if (c & 1) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(acc[c]) : "r"(a), "r"(b));
                else asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(acc[c]) : "r"(a), "r"(b));
            }
            if (OP == 9) { float f = __int_as_float(acc[c]); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(__int_as_float(a)), "f"(__int_as_float(b))); acc[c] = __float_as_int(f); }
            if (OP == 10) asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(acc[c]) : "r"(a & 31));
            if (OP == 12) { // fp64 multiply (two chains share one 64-bit register pair)
                if (c & 1) { double x = __hiloint2double(acc[c], acc[c - 1]); asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(x) : "d"(__hiloint2double(a, b))); acc[c] = __double2hiint(x); acc[c - 1] = __double2loint(x); }
            }
            if (OP == 13) { float f = __int_as_float(acc[c]); double x; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(x) : "f"(f)); acc[c] ^= __double2hiint(x); }
            if (OP == 14) { int v; asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(v) : "f"(__int_as_float(acc[c]))); acc[c] += v; }
            if (OP == 15) { float f; asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f) : "r"(acc[c])); acc[c] ^= __float_as_int(f); }
            if (OP == 16) asm volatile("mul.hi.s32 %0, %0, %1;" : "+r"(acc[c]) : "r"(a));
            if (OP == 17) { long long w; asm volatile("mad.wide.s32 %0, %1, %2, %3;" : "=l"(w) : "r"(acc[c]), "r"(a), "l"((long long)b)); acc[c] = (int)(w >> 32); }
            if (OP == 11) { // dp2a + imad alternating (both fma pipe?)
                if (c & 1) asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(acc[c]) : "r"(a), "r"(b));
                else asm volatile("dp2a.lo.u32.s32 %0, %1, %2, %0;" : "+r"(acc[c]) : "r"(a), "r"(b));
            }
        }
    }
    int s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s ^= acc[c];
    if (s == 0x12345678) out[threadIdx.x] = s;
}

template <int OP>
void run(const char *name, int *d_out, int sms, int clock_khz)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int blocks = sms * 8;
    probe<OP><<<blocks, 256>>>(d_out, 3, 5);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        probe<OP><<<blocks, 256>>>(d_out, 3, 5);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double warp_instr = (double)blocks * 8 /*warps*/ * ITER * CHAINS;
    double per_s = warp_instr / (best * 1e-3);
    printf("%-22s %8.3f ms  %7.2f Gwarp-instr/s  = %5.2f warp-instr/clk/SM @%d MHz (max clock)\n", name, best,
           per_s * 1e-9, per_s / sms / (clock_khz * 1e3), clock_khz / 1000);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%s  SMs=%d  clockRate=%d kHz\n", p.name, p.multiProcessorCount, clk);
    int *d;
    cudaMalloc(&d, 4096);
    int sms = p.multiProcessorCount;
    run<0>("dp2a.lo.u32.s32", d, sms, clk);
    run<1>("dp4a.s32.s32", d, sms, clk);
    run<2>("mad.lo.s32 (IMAD)", d, sms, clk);
    run<3>("prmt", d, sms, clk);
    run<4>("lop3", d, sms, clk);
    run<5>("shf", d, sms, clk);
    run<6>("add.s32", d, sms, clk);
    run<9>("fma.rn.f32", d, sms, clk);
    run<10>("shfl.idx", d, sms, clk);
    run<7>("dp2a+prmt 1:1", d, sms, clk);
    run<8>("imad+prmt 1:1", d, sms, clk);
    run<11>("dp2a+imad 1:1", d, sms, clk);
    run<12>("mul.rn.f64 (x0.5: every other chain)", d, sms, clk);
    run<13>("cvt.f64.f32 + xor", d, sms, clk);
    run<14>("cvt.rzi.s32.f32 + add", d, sms, clk);
    run<15>("cvt.rn.f32.s32 + xor", d, sms, clk);
    run<16>("mul.hi.s32", d, sms, clk);
    run<17>("mad.wide.s32 (hi word)", d, sms, clk);
    return 0;
}
