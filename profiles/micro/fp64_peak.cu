// FP64 pipe microbenchmark for the roofline denominator of the Newton kernels (SURVEY.md 8d:
// "FP64 peak is not in MEASURED_PEAKS.json -- measure it with a DFMA-chain microbenchmark").
// Reports warp-level instruction throughput for dependent chains of DFMA, DMUL, DADD and the
// no-contraction DMUL+DADD pair at several ILP / occupancy points.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP, int ILP>
__global__ void __launch_bounds__(256) k(double *out, int iters, double a, double b)
{
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) v[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                if (OP == 0) v[i] = __fma_rn(v[i], a, b);
                else if (OP == 1) v[i] = __dmul_rn(v[i], a);
                else if (OP == 2) v[i] = __dadd_rn(v[i], b);
                else v[i] = __dadd_rn(__dmul_rn(v[i], a), b);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += v[i];
    if (s == 123.456) out[0] = s;
}

template <int OP, int ILP>
void run(const char *name, int blocks_per_sm, int sms, double *d)
{
    const int iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP, ILP><<<sms * blocks_per_sm, 256>>>(d, 16, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP, ILP><<<sms * blocks_per_sm, 256>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double per_thread = (double)iters * 8 * ILP * (OP == 3 ? 2 : 1);
    double inst = per_thread * 256.0 * sms * blocks_per_sm;          // thread-level fp64 instructions
    printf("%-10s ILP=%d warps/SM=%2d : %8.3f T thread-instr/s  (%.1f%% of 148x64x1.965e9)\n", name, ILP,
           blocks_per_sm * 8, inst / ms * 1e-9, 100. * inst / (ms * 1e-3) / (sms * 64 * 1.965e9));
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *d;
    cudaMalloc(&d, 64);
    printf("SMs: %d\n", sms);
    for (int b : {1, 2, 4, 8}) {
        run<0, 1>("DFMA", b, sms, d); run<0, 2>("DFMA", b, sms, d); run<0, 4>("DFMA", b, sms, d);
        run<1, 1>("DMUL", b, sms, d); run<1, 4>("DMUL", b, sms, d);
        run<2, 1>("DADD", b, sms, d); run<2, 4>("DADD", b, sms, d);
        run<3, 1>("DMUL+DADD", b, sms, d); run<3, 2>("DMUL+DADD", b, sms, d); run<3, 4>("DMUL+DADD", b, sms, d);
    }
    return 0;
}
