// Pure-compute throughput of the Wolter-I per-ray chain (no global loads in the loop): every thread
// traces rays it synthesises from its id, over and over.  Sweeps resident warps per SM to see where
// the SM saturates and at what instruction rate -- the compute ceiling of the fused trace kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true \
//        -I../../pyxfocus_b200/csrc -o trace_compute trace_compute.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "pxf_internal.h"
#include "pxf_params.h"

using namespace pxf;

struct Prm { TransformP t; WolterP w; };

template <int WHAT, int MAXB>
__global__ void __launch_bounds__(256, MAXB) k(double *out, int reps, const __grid_constant__ Prm p)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0.;
    for (int it = 0; it < reps; it++) {
        // annular source at z=0 travelling in -z (the bench's subannulus), deterministic per thread/rep
        const double u = ((gid * 977 + it * 131) & 0xffff) * (1. / 65536.);
        const double v = ((gid * 331 + it * 57) & 0xffff) * (1. / 65536.);
        Ray r;
        const double rho = 220. + .6 * u;
        r.x = rho * (1. - .5 * v * v); r.y = rho * v * .999; r.z = 0.;
        r.l = 0.; r.m = 0.; r.n = -1.; r.ux = r.uy = r.uz = 0.; r.opd = 0.;
        if (WHAT >= 1) op_transform(r, p.t);
        if (WHAT >= 2) { op_wolterprimary(r, p.w); op_reflect(r); }
        if (WHAT >= 3) { op_woltersecondary(r, p.w); op_reflect(r); }
        if (WHAT >= 4) op_flat(r, false, 0.);
        acc += r.x + r.y + r.z + r.l + r.m + r.n + r.ux + r.uy + r.uz;
    }
    if (acc == 123.456) out[0] = acc;
}

template <int WHAT, int MAXB>
void run(int blocks_per_sm, int sms, double *d, const Prm &p)
{
    if (blocks_per_sm > MAXB) return;
    const int reps = 64;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<WHAT, MAXB><<<sms * blocks_per_sm, 256>>>(d, 2, p);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<WHAT, MAXB><<<sms * blocks_per_sm, 256>>>(d, reps, p);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double rays = (double)reps * 256 * sms * blocks_per_sm;
    printf("what=%d regcap=%3d warps/SM=%2d : %7.3f ms  %6.2f Grays/s  => 1.25e8 rays in %.3f ms\n", WHAT,
           65536 / (256 * MAXB), blocks_per_sm * 8, ms, rays / ms * 1e-6, 1.25e8 / (rays / ms));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *d;
    cudaMalloc(&d, 64);
    Prm p;
    p.t = make_transform(-0., -0., 8400., -0., -0., -0.);
    p.t.groups = 3;
    p.w = make_wolter(220., 8400., 1., false, 0.);
    for (int b : {1, 2, 3, 4}) run<4, 4>(b, sms, d, p);
    for (int b : {1, 2, 3, 4, 5, 6, 8}) run<4, 8>(b, sms, d, p);
    for (int b : {3, 4}) { run<1, 4>(b, sms, d, p); run<2, 4>(b, sms, d, p); run<3, 4>(b, sms, d, p); }
    return 0;
}
