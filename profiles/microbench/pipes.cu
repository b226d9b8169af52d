// Micro-benchmark: which issue pipe do the candidate byte->float conversions of the BVH node test use on sm_100a?
// Each kernel runs ITER x 8 independent instructions per thread of the op(s) under test; the time per warp
// instruction per SM sub-partition tells the pipe: ops on different pipes overlap when interleaved, ops on the same
// pipe add up.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run: ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed, float fa, float fb)
{
    uint32_t x[8];
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = seed + threadIdx.x * 8 + i; f[i] = (float)(threadIdx.x + i); }
    uint32_t one = seed | 0x3f800000u, sel = seed | 0x00008000u;
    for (int it = 0; it < ITER; it++)
    {
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
            if (MODE == 0 || MODE == 3 || MODE == 5) asm volatile("prmt.b32 %0, %0, %1, 0x7614;" : "+r"(x[i]) : "r"(one));
            if (MODE == 1 || MODE == 4 || MODE == 5) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(sel), "r"(one));
            if (MODE == 2 || MODE == 3 || MODE == 4) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb));
            if (MODE == 6) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fa));
            if (MODE == 7) { asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(fa)); asm volatile("prmt.b32 %0, %0, %1, 0x7614;" : "+r"(x[i]) : "r"(one)); }
            if (MODE == 8) { asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f[i]) : "r"(x[i] & 0xffu)); }
            if (MODE == 9) { asm volatile("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, lo; }" : "=f"(f[i]) : "r"(x[i])); x[i] += 1; }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i] + __float_as_uint(f[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, int ops_per_slot)
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* out;
    cudaMalloc(&out, (size_t)sms * 8 * 256 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<sms * 8, 256>>>(out, 0, 1.0001f, 0.5f);
    cudaEventRecord(e0);
    k<MODE><<<sms * 8, 256>>>(out, 0, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    // warp instructions per SM sub-partition: 8 blocks x 8 warps / 4 SMSPs = 16 warps per SMSP
    const double winst = 16.0 * ITER * 8 * ops_per_slot;
    printf("%-28s %8.3f ms  %6.2f cycles per warp-instruction per SMSP (at %d MHz nominal)\n", name, ms,
           ms * 1e-3 * clk * 1e3 / winst, clk / 1000);
    cudaFree(out);
}
int main()
{
    run<0>("PRMT", 1);
    run<1>("IDP4A", 1);
    run<2>("FFMA", 1);
    run<3>("PRMT+FFMA", 2);
    run<4>("IDP4A+FFMA", 2);
    run<5>("PRMT+IDP4A", 2);
    run<6>("FMNMX", 1);
    run<7>("FMNMX+PRMT", 2);
    run<8>("LOP+I2F.U32", 2);
    run<9>("F2F.F32.F16 (+IADD)", 2);
    return 0;
}
