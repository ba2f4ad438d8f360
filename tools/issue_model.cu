// tools/issue_model.cu -- micro-benchmark of the SM's issue model on B200: how many warp-instructions per clock one SM sustains
// for pure and mixed streams of FP64 (DFMA), FP32 (FFMA), integer ALU (LOP3/IADD3) and IMAD instructions.  Used to state the
// issue-slot roofline of K1.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/issue_model tools/issue_model.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP8(x) x x x x x x x x

template <int ND, int NF, int NI, int NM>
__global__ void mix_kernel(double* out, int iters, double a, double b, float fa, float fb, uint32_t ia, uint32_t ib, long long* cycles)
{
    double d[8];
    float f[8];
    uint32_t i[8], m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { d[k] = threadIdx.x + k; f[k] = threadIdx.x + 2 * k; i[k] = threadIdx.x * 7 + k; m[k] = threadIdx.x * 3 + k; }
    const long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (k < ND) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[k]) : "d"(a), "d"(b));
            if (k < NF) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[k]) : "f"(fa), "f"(fb));
            if (k < NI) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(i[k]) : "r"(ia), "r"(ib));
            if (k < NM) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(m[k]) : "r"(ia), "r"(ib));
        }
    }
    const long long c1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += d[k] + f[k] + i[k] + m[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = c1 - c0;
}

template <int ND, int NF, int NI, int NM>
void run(const char* name, int sms)
{
    const int blocks = sms * 4, threads = 256, iters = 20000;  // 32 warps per SM = 8 per scheduler
    double* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaMalloc(&cyc, sizeof(long long));
    mix_kernel<ND, NF, NI, NM><<<blocks, threads>>>(out, 100, 0.999, 1e-9, 0.999f, 1e-9f, 0x1234567u, 0x89abcdeu, cyc);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    mix_kernel<ND, NF, NI, NM><<<blocks, threads>>>(out, iters, 0.999, 1e-9, 0.999f, 1e-9f, 0x1234567u, 0x89abcdeu, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long c;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const double instr_per_warp = (double)iters * (ND + NF + NI + NM);
    // per scheduler: 8 warps each issuing instr_per_warp instructions in c cycles
    printf("%-28s FP64:%d FP32:%d ALU:%d IMAD:%d  %.3f warp-instr/clk/scheduler (%.3f ms, %lld cycles)\n", name, ND, NF, NI, NM,
           8.0 * instr_per_warp / (double)c, ms, c);
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    const int s = p.multiProcessorCount;
    run<8, 0, 0, 0>("DFMA only", s);
    run<0, 8, 0, 0>("FFMA only", s);
    run<0, 0, 8, 0>("LOP3 only", s);
    run<0, 0, 0, 8>("IMAD only", s);
    run<4, 4, 0, 0>("DFMA:FFMA 1:1", s);
    run<4, 0, 4, 0>("DFMA:LOP3 1:1", s);
    run<4, 0, 0, 4>("DFMA:IMAD 1:1", s);
    run<2, 4, 0, 0>("DFMA:FFMA 1:2", s);
    run<2, 0, 4, 0>("DFMA:LOP3 1:2", s);
    run<2, 2, 2, 2>("DFMA:FFMA:LOP3:IMAD 1:1:1:1", s);
    run<2, 4, 2, 0>("DFMA:FFMA:LOP3 1:2:1", s);
    run<0, 4, 4, 0>("FFMA:LOP3 1:1", s);
    run<0, 4, 0, 4>("FFMA:IMAD 1:1", s);
    run<0, 0, 4, 4>("LOP3:IMAD 1:1", s);
    run<3, 2, 3, 1>("K1-like 3:2:3:1", s);
    return 0;
}
