// fp64 pipe throughput as a function of how many DISTINCT 64-bit REGISTER source operands an
// instruction reads (constant-bank / immediate operands and repeated registers are free).
// Hypothesis from the trace kernel: it runs at ~3.1 pipe cycles per fp64 instruction although the
// pipe issues one warp instruction every 2 cycles from constant-operand streams.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp64_operands fp64_operands.cu
#include <cstdio>
#include <cuda_runtime.h>

// OP: 0 DADD r,c   1 DADD r,r   2 DMUL r,c   3 DMUL r,r   4 DFMA v,c,v  5 DFMA v,w,c
//     6 DFMA v,w,u (3 distinct)  7 DFMA v,c,u   8 DFMA v,w,v
template <int OP, int ILP>
__global__ void __launch_bounds__(256) k(double *out, const double *in, int iters, double ca, double cb)
{
    double v[ILP], w[ILP], u[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) {
        v[i] = in[threadIdx.x + i];
        w[i] = in[threadIdx.x + 8 + i];
        u[i] = in[threadIdx.x + 16 + i];
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int t = 0; t < 16; t++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                if (OP == 0) v[i] = __dadd_rn(v[i], cb);
                if (OP == 1) v[i] = __dadd_rn(v[i], w[i]);
                if (OP == 2) v[i] = __dmul_rn(v[i], ca);
                if (OP == 3) v[i] = __dmul_rn(v[i], w[i]);
                if (OP == 4) v[i] = __fma_rn(v[i], ca, v[i]);
                if (OP == 5) v[i] = __fma_rn(v[i], w[i], cb);
                if (OP == 6) v[i] = __fma_rn(v[i], w[i], u[i]);
                if (OP == 7) v[i] = __fma_rn(v[i], ca, u[i]);
                if (OP == 8) v[i] = __fma_rn(v[i], w[i], v[i]);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += v[i];
    if (s == 123.456) out[0] = s;
}

template <int OP, int ILP>
void run(const char *name, int blocks_per_sm, int sms, double *d, double *in)
{
    const int iters = 2048;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP, ILP><<<sms * blocks_per_sm, 256>>>(d, in, 16, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP, ILP><<<sms * blocks_per_sm, 256>>>(d, in, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double cyc = ms * 1e-3 * 1.965e9 / ((double)iters * 16 * ILP * blocks_per_sm * 2);
    printf("%-28s ILP=%d warps/SM=%2d : %6.3f pipe cycles per warp instruction\n", name, ILP, blocks_per_sm * 8, cyc);
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *d, *in;
    cudaMalloc(&d, 64);
    cudaMalloc(&in, 4096);
    cudaMemset(in, 0, 4096);
    for (int b : {2, 4}) {
#define ALL(I) \
        run<0, I>("DADD r,c", b, sms, d, in); run<1, I>("DADD r,r", b, sms, d, in); \
        run<2, I>("DMUL r,c", b, sms, d, in); run<3, I>("DMUL r,r", b, sms, d, in); \
        run<4, I>("DFMA v,c,v (1 reg)", b, sms, d, in); run<5, I>("DFMA v,w,c (2 regs)", b, sms, d, in); \
        run<7, I>("DFMA v,c,u (2 regs)", b, sms, d, in); run<8, I>("DFMA v,w,v (2 regs)", b, sms, d, in); \
        run<6, I>("DFMA v,w,u (3 regs)", b, sms, d, in);
        ALL(1) ALL(2) ALL(4)
    }
    return 0;
}
