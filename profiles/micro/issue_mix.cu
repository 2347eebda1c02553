// Which pipes can issue in the shadow of an fp64 warp instruction (2 pipe cycles on B200)?
// Mixes an fp64 stream (DMUL/DADD, ILP 4) with NI independent instructions of another pipe per
// NF fp64 instructions and prints the scheduler cycles per group:
//   max(2*NF, cost*NI) => the other pipe issues in the shadow; 2*NF + cost*NI => it does not.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o issue_mix issue_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

// KIND 0: LOP3 (ALU)  1: IMAD (FMA-heavy)  2: FFMA  3: ISETP+SEL (ALU)  4: MUFU.RCP (XU)  5: IADD3 (ALU)
template <int NF, int NI, int KIND>
__global__ void __launch_bounds__(256) k(double *out, const unsigned *in, int iters, double a, double b)
{
    double v[4];
    unsigned w[4];
    float f[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { v[i] = threadIdx.x * 1e-9 + i; w[i] = in[threadIdx.x + i]; f[i] = 1.f + in[threadIdx.x + i + 4]; }
    const unsigned c = in[300];
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int j = 0; j < NF; j++)
#pragma unroll
                for (int i = 0; i < 4; i++) v[i] = (j & 1) ? __dadd_rn(v[i], b) : __dmul_rn(v[i], a);
#pragma unroll
            for (int j = 0; j < NI; j++)
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (KIND == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[i]) : "r"(w[(i + 1) & 3]), "r"(c));
                    else if (KIND == 1) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(w[i]) : "r"(c));
                    else if (KIND == 2) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"((float)a));
                    else if (KIND == 3) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %1, %2, p;}" : "+r"(w[i]) : "r"(w[(i + 1) & 3]), "r"(c));
                    else if (KIND == 4) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
                    else asm volatile("add.u32 %0, %0, %1;" : "+r"(w[i]) : "r"(w[(i + 1) & 3]));
                }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) s += v[i] + w[i] + f[i];
    if (s == 123.456) out[0] = s;
}

static const char *names[] = {"LOP3", "IMAD", "FFMA", "ISETP+SEL", "MUFU", "IADD"};
template <int NF, int NI, int KIND>
float run(int blocks_per_sm, int sms, double *d, unsigned *in)
{
    const int iters = 2048;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NF, NI, KIND><<<sms * blocks_per_sm, 256>>>(d, in, 16, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<NF, NI, KIND><<<sms * blocks_per_sm, 256>>>(d, in, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double cyc = ms * 1e-3 * 1.965e9 / ((double)iters * 8 * 4 * blocks_per_sm * 2);
    printf("%d fp64 + %d %-9s warps/SM=%2d : %6.2f scheduler cycles per group\n", NF, NI, names[KIND], blocks_per_sm * 8, cyc);
    return ms;
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *d;
    unsigned *in;
    cudaMalloc(&d, 64);
    cudaMalloc(&in, 4096);
    cudaMemset(in, 1, 4096);
    const int b = 4;
    run<2, 0, 0>(b, sms, d, in);
#define KINDS(K) run<0, 2, K>(b, sms, d, in); run<0, 4, K>(b, sms, d, in); run<2, 1, K>(b, sms, d, in); \
                 run<2, 2, K>(b, sms, d, in); run<2, 4, K>(b, sms, d, in);
    KINDS(0) KINDS(5) KINDS(3) KINDS(1) KINDS(2) KINDS(4)
    return 0;
}
