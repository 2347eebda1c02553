// Row access and the per-CTA centroid reduction shared by the interpreter (pxf_fused.cu), the statically specialised
// chains (pxf_chain.cu) and the run-time specialised ones (pxf_jit.cu compiles this header with NVRTC): device code
// only, no host headers.
#pragma once
#include "pxf_ray.cuh"
#ifndef PXF_BLOCK
#define PXF_BLOCK 256
#endif

namespace pxf {

// ---- row access shared by the interpreter and the specialised chains (device code) ----
struct RowPtrs { double *p[10]; };

PXF_DEV void fload1(Ray &r, const RowPtrs &P, unsigned M, int64_t i)
{
    r.opd = (M & R_OPD) ? P.p[0][i] : 0.;
    r.x = (M & R_X) ? P.p[1][i] : 0.;  r.y = (M & R_Y) ? P.p[2][i] : 0.;  r.z = (M & R_Z) ? P.p[3][i] : 0.;
    r.l = (M & R_L) ? P.p[4][i] : 0.;  r.m = (M & R_M) ? P.p[5][i] : 0.;  r.n = (M & R_N) ? P.p[6][i] : 0.;
    r.ux = (M & R_UX) ? P.p[7][i] : 0.; r.uy = (M & R_UY) ? P.p[8][i] : 0.; r.uz = (M & R_UZ) ? P.p[9][i] : 0.;
}
PXF_DEV void fstore1(const Ray &r, const RowPtrs &P, unsigned M, int64_t i)
{
    if (M & R_OPD) P.p[0][i] = r.opd;
    if (M & R_X) P.p[1][i] = r.x;   if (M & R_Y) P.p[2][i] = r.y;   if (M & R_Z) P.p[3][i] = r.z;
    if (M & R_L) P.p[4][i] = r.l;   if (M & R_M) P.p[5][i] = r.m;   if (M & R_N) P.p[6][i] = r.n;
    if (M & R_UX) P.p[7][i] = r.ux; if (M & R_UY) P.p[8][i] = r.uy; if (M & R_UZ) P.p[9][i] = r.uz;
}
#define FLD2(bit, k, f)                                                               \
    if (M & bit) { double2 v = *reinterpret_cast<const double2 *>(P.p[k] + i); a.f = v.x; b.f = v.y; } \
    else { a.f = 0.; b.f = 0.; }
#define FST2(bit, k, f) if (M & bit) *reinterpret_cast<double2 *>(P.p[k] + i) = make_double2(a.f, b.f);
PXF_DEV void fload2(Ray &a, Ray &b, const RowPtrs &P, unsigned M, int64_t i)
{
    FLD2(R_OPD, 0, opd) FLD2(R_X, 1, x) FLD2(R_Y, 2, y) FLD2(R_Z, 3, z) FLD2(R_L, 4, l)
    FLD2(R_M, 5, m) FLD2(R_N, 6, n) FLD2(R_UX, 7, ux) FLD2(R_UY, 8, uy) FLD2(R_UZ, 9, uz)
}
PXF_DEV void fstore2(const Ray &a, const Ray &b, const RowPtrs &P, unsigned M, int64_t i)
{
    FST2(R_OPD, 0, opd) FST2(R_X, 1, x) FST2(R_Y, 2, y) FST2(R_Z, 3, z) FST2(R_L, 4, l)
    FST2(R_M, 5, m) FST2(R_N, 6, n) FST2(R_UX, 7, ux) FST2(R_UY, 8, uy) FST2(R_UZ, 9, uz)
}

// Per-CTA partial sums of (count, x, y) over the final state of the surviving rays -- the
// centroid that analyses.hpd / rmsCentroid need next -- written by the trace kernel itself so
// that no extra pass over x,y is needed.  Layout matches k_sums: partial[blockIdx*9 + {0,1,2,3}]
// = {count, sum x, sum y, count}.
#define PXF_NSUM 9
PXF_DEV void centroid_block_reduce(double cnt, double sx, double sy, double *__restrict__ partial)
{
    __shared__ double sh[3][PXF_BLOCK / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double v[3] = {cnt, sx, sy};
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
        if (lane == 0) sh[k][warp] = v[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            double t = lane < PXF_BLOCK / 32 ? sh[k][lane] : 0.;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
            if (lane == 0) {
                partial[(int64_t)blockIdx.x * PXF_NSUM + k] = t;
                if (k == 0) partial[(int64_t)blockIdx.x * PXF_NSUM + 3] = t;
            }
        }
    }
}

}  // namespace pxf
