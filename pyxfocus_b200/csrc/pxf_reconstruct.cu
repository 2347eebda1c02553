// Southwell wavefront reconstruction (reconstruct.f95:1-128) and lenslet binning (:136-187) -- SURVEY.md 8(f)
// rank 4, the consumer of traced bundles in the reference's metrology scripts (analyses.wavefront,
// southwell.southwell).
//
// The reference iterates successive over-relaxation in LEXICOGRAPHIC Gauss-Seidel order (xi outer, yi inner):
// cell (xi,yi) of sweep s reads its west/south neighbours from sweep s and its east/north neighbours from
// sweep s-1.  That order looks sequential but pipelines exactly: give cell (xi,yi) of sweep s the time stamp
//     t = (xi + yi) + 2 s .
// Its west/south neighbours (diagonal xi+yi-1, sweep s) carry t-1, its east/north neighbours (diagonal
// xi+yi+1, sweep s-1) carry t-1 as well, and nothing else touches what it reads or writes at time t.  So at
// every time step ALL cells whose diagonal has the parity of t are updated at once -- each diagonal in a
// different sweep -- in place, and the result is bit-identical to the sequential loop, including the
// in-sweep invalidation of lenslets without valid neighbours (:88-94: it is just a value that later cells read).
// 2 time steps = 1 sweep of the whole grid; the pipeline depth is half the number of diagonals.
//
// The per-sweep convergence test (rms of the update < criteria, :112-117) is evaluated per sweep index from
// per-diagonal partial sums added in diagonal order (deterministic; the reference adds cell by cell, so the two
// rms values agree to rounding).  Sweeps run in batches; when a batch contains the first converged sweep k*,
// the arrays are restored from the batch's checkpoint and exactly k*+1 sweeps are re-run, so the number of
// sweeps -- and with it every bit of the result -- is the reference's.
//
// A warp per 32 cells of an active diagonal.  Tiny grids run in one CTA (barrier = __syncthreads); anything
// larger runs as a cooperative grid of up to one CTA per SM with grid.sync() between time steps and L2 loads
// (one CTA alone is issue bound: 75 us per time step at 256 x 256).
#include <cooperative_groups.h>
#include <math.h>
#include <vector>
#include "pxf_internal.h"
#include "pxf_ray.cuh"

namespace cg = cooperative_groups;

namespace pxf {

#define RECON_THREADS 512
#define RA(a, xi, yi) (a)[((xi) - 1) + (int64_t)((yi) - 1) * xdim]

// COOP = false: one CTA, __syncthreads between time steps, cached loads (the CTA's own SM wrote the data).
// COOP = true : a cooperative grid, grid.sync() between time steps, L2 loads (__ldcg: other SMs wrote the data).
// Work item of a time step = (sweep k, chunk c of 32 consecutive cells of its diagonal); one warp per item.
//
// Convergence (:112-117) is tested in the kernel: acc/cnt hold, per sweep in flight (ring of `ring` sweeps) and
// chunk, the sum of the squared updates and the number of updated cells -- slot (k,c) is touched by exactly one
// warp per time step, so the accumulation order is fixed.  One step after sweep k has finished, warp 0 of CTA 0
// adds its chunks in order and, if sqrt(sum/n) < criteria, publishes ctl->stop = k; everybody leaves two steps
// later.  By then the sweeps behind k have run ahead by up to half the pipeline depth, so the state "after
// exactly k+1 sweeps" no longer exists in the arrays.  It is rebuilt from a SNAPSHOT: every `every` sweeps the
// sweep that passes writes each cell it has just processed (updated or not) to one of two snapshot buffers
// (ping-pong; every >= depth/2 + 2 guarantees that the snapshot at or before k is complete and not yet being
// overwritten when the kernel stops).  The host restores the latest snapshot at or before k+1 sweeps and re-runs
// the few missing sweeps with a plain pipelined launch.
struct ReconCtl { int stop; int slot[2]; int pad; };    // slot[t&1]: written in step t, read at the top of step t+1
struct ReconSnap { double *x[2], *y[2], *p[2]; int every; };

template <bool COOP>
__global__ void __launch_bounds__(RECON_THREADS)
k_reconstruct(double *xang, double *yang, double *ph, const int xdim, const int ydim,
              const double w, const double h, const int nsweep, const int nchunk, const int ring, double *acc, int *cnt,
              const double criteria, ReconCtl *ctl, const ReconSnap snap)
{
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31;
    const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int nwarp = (int)((gridDim.x * blockDim.x) >> 5);
    const int dmin = 4, dmax = xdim + ydim - 2;             // diagonals xi+yi of the interior 2..dim-1
    const int R = dmax - dmin;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < ring * nchunk; k += gridDim.x * blockDim.x) { acc[k] = 0.; cnt[k] = 0; }
    if (blockIdx.x == 0 && threadIdx.x == 0) { ctl->stop = -1; ctl->slot[0] = -1; ctl->slot[1] = -1; }
    if (COOP) grid.sync(); else __syncthreads();
    auto ld = [](const double *p) { return COOP ? __ldcg(p) : *p; };
    const int T = (R + 1) + 2 * (nsweep - 1);
    for (int t = 0; t < T; t++) {
        // grid-uniform: the slot read here was written during step t-1, before its barrier; step t writes the other one
        if (t > 0 && (COOP ? __ldcg(&ctl->slot[(t - 1) & 1]) : *(volatile int *)&ctl->slot[(t - 1) & 1]) >= 0) break;
        // sweep kd finished in the previous step: test it (one warp) and recycle its ring slot
        if (warp == 0 && t - 1 - R >= 0 && ((t - 1 - R) & 1) == 0) {
            const int kd = (t - 1 - R) >> 1;
            if (lane == 0) {
                double r2 = 0.;
                long long nn = 0;
                for (int c = 0; c < nchunk; c++) {
                    double *pa = acc + (kd % ring) * nchunk + c;
                    int *pn = cnt + (kd % ring) * nchunk + c;
                    r2 += COOP ? __ldcg(pa) : *pa;
                    nn += COOP ? __ldcg(pn) : *pn;
                    *pa = 0.;
                    *pn = 0;
                }
                if (sqrt(r2 / (double)nn) < criteria) { ctl->stop = kd; ctl->slot[t & 1] = kd; }   // 0/0 = NaN: false
            }
        }
        // sweep k works on diagonal d = dmin + t - 2k
        int klo = (t - R + 1) / 2;                          // ceil((t - R) / 2)
        if (klo < 0) klo = 0;
        int khi = t / 2;
        if (khi > nsweep - 1) khi = nsweep - 1;
        const int nitem = (khi - klo + 1) * nchunk;
        for (int item = warp; item < nitem; item += nwarp) {
            const int k = klo + item / nchunk, c = item % nchunk;
            const int d = dmin + t - 2 * k;
            int x0 = d - (ydim - 1);
            if (x0 < 2) x0 = 2;
            int x1 = d - 2;
            if (x1 > xdim - 1) x1 = xdim - 1;
            const int xi = x0 + 32 * c + lane;
            double a2 = 0.;
            int n = 0;
            if (xi <= x1) {
                const int yi = d - xi;
                double pc = ld(&RA(ph, xi, yi));
                double ya = ld(&RA(yang, xi, yi)), xa = ld(&RA(xang, xi, yi));
                if (pc != 100.) {                                                     // :33-35
                    double yplus = ld(&RA(yang, xi, yi + 1)); if (yplus == 100.) yplus = -ya;   // :39-54
                    double yneg = ld(&RA(yang, xi, yi - 1));  if (yneg == 100.) yneg = -ya;
                    double xplus = ld(&RA(xang, xi + 1, yi)); if (xplus == 100.) xplus = -xa;
                    double xneg = ld(&RA(xang, xi - 1, yi));  if (xneg == 100.) xneg = -xa;
                    const double bk = .5 * (yplus - yneg + xplus - xneg) * h;         // :57
                    double goodpix = 4.;
                    double pyplus = ld(&RA(ph, xi, yi + 1)); if (pyplus == 100.) { pyplus = 0.; goodpix = goodpix - 1; }
                    double pyneg = ld(&RA(ph, xi, yi - 1));  if (pyneg == 100.) { pyneg = 0.; goodpix = goodpix - 1; }
                    double pxplus = ld(&RA(ph, xi + 1, yi)); if (pxplus == 100.) { pxplus = 0.; goodpix = goodpix - 1; }
                    double pxneg = ld(&RA(ph, xi - 1, yi));  if (pxneg == 100.) { pxneg = 0.; goodpix = goodpix - 1; }
                    const double psum = pyplus + pyneg + pxplus + pxneg;              // :85
                    if (goodpix == 0.) {                                              // :88-94
                        RA(ph, xi, yi) = 100.;
                        RA(xang, xi, yi) = 100.;
                        RA(yang, xi, yi) = 100.;
                        pc = 100.; xa = 100.; ya = 100.;
                    } else {
                        const double nv = pc + w * ((psum + bk) / goodpix - pc);      // :99
                        RA(ph, xi, yi) = nv;
                        n = 1;
                        a2 = sq(nv - pc);                                             // :103 (phase = value before the sweep)
                        pc = nv;
                    }
                }
                if (snap.every > 0 && (k + 1) % snap.every == 0) {
                    const int sb = ((k + 1) / snap.every) & 1;
                    RA(snap.p[sb], xi, yi) = pc;
                    RA(snap.x[sb], xi, yi) = xa;
                    RA(snap.y[sb], xi, yi) = ya;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a2 += __shfl_down_sync(0xffffffffu, a2, o);
                n += __shfl_down_sync(0xffffffffu, n, o);
            }
            if (lane == 0 && n) {
                double *pa = acc + (k % ring) * nchunk + c;
                int *pn = cnt + (k % ring) * nchunk + c;
                *pa = (COOP ? __ldcg(pa) : *pa) + a2;
                *pn = (COOP ? __ldcg(pn) : *pn) + n;
            }
        }
        if (COOP) grid.sync(); else __syncthreads();
    }
    // the last sweep of a run that was not stopped finishes in the final step: test it here
    if (blockIdx.x == 0 && threadIdx.x == 0 && (COOP ? __ldcg(&ctl->stop) : *(volatile int *)&ctl->stop) < 0) {
        const int kd = nsweep - 1;
        double r2 = 0.;
        long long nn = 0;
        for (int c = 0; c < nchunk; c++) {
            r2 += COOP ? __ldcg(acc + (kd % ring) * nchunk + c) : acc[(kd % ring) * nchunk + c];
            nn += COOP ? __ldcg(cnt + (kd % ring) * nchunk + c) : cnt[(kd % ring) * nchunk + c];
        }
        if (sqrt(r2 / (double)nn) < criteria) ctl->stop = kd;
    }
}

// ---------------------------------------------------------------- southwellbin
// Bin index per ray (Fortran 1-based xb,yb folded to a 0-based cell number, -1 outside the array).
__global__ void __launch_bounds__(PXF_BLOCK)
k_swb_cell(const double *__restrict__ x, const double *__restrict__ y, int64_t num, double binsize, int xdim, int ydim,
           double *__restrict__ key)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < num; i += nthr) {
        int xb, yb;
        if (xdim % 2 == 0) xb = (int)floor(x[i] / binsize) + xdim / 2 + 1 + 1;      // :155-156
        else xb = (int)floor((x[i] + binsize / 2) / binsize) + (xdim - 1) / 2 + 1;
        if (ydim % 2 == 0) yb = (int)floor(y[i] / binsize) + ydim / 2 + 1 + 1;
        else yb = (int)floor((y[i] + binsize / 2) / binsize) + (ydim - 1) / 2 + 1;
        const bool in = xb >= 1 && xb <= xdim && yb >= 1 && yb <= ydim;
        key[i] = in ? (double)((xb - 1) + (int64_t)(yb - 1) * xdim) : (double)((int64_t)xdim * ydim);
    }
}

// first sorted position of every cell that has rays
__global__ void __launch_bounds__(PXF_BLOCK)
k_swb_starts(const double *__restrict__ skey, int64_t num, long long *__restrict__ start, long long ncell)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = tid; p < num; p += nthr) {
        const long long c = (long long)skey[p];
        if (c < ncell && (p == 0 || (long long)skey[p - 1] != c)) start[c] = p;
    }
}

// one thread per cell: sum its rays in ray order (the stable sort kept it), then normalise (:172-184)
__global__ void __launch_bounds__(PXF_BLOCK)
k_swb_sum(const double *__restrict__ skey, const long long *__restrict__ idx, int64_t num, const long long *__restrict__ start,
          const double *__restrict__ l, const double *__restrict__ m, long long ncell, double *__restrict__ xang,
          double *__restrict__ yang, double *__restrict__ phase)
{
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    const long long s = start[c];
    if (s < 0) { phase[c] = 100.; xang[c] = 100.; yang[c] = 100.; return; }
    double sx = 0., sy = 0.;
    int n = 0;
    for (long long p = s; p < num && (long long)skey[p] == c; p++) {
        const long long i = idx[p];
        sx = sx + l[i];
        sy = sy + m[i];
        n++;
    }
    xang[c] = tan(asin(sx / n));
    yang[c] = tan(asin(sy / n));
    phase[c] = 0.;
}

}  // namespace pxf

using namespace pxf;

extern "C" {

// xang, yang, phase, phasec: device, column-major [xdim][ydim] (xi fastest).  Synchronises the stream.
int pxf_reconstruct(double *xang, double *yang, int32_t xdim, int32_t ydim, double criteria, double h, double *phase,
                    double *phasec, int32_t maxiter, int64_t *sweeps_host, pxf_stream_t stream)
{
    if (!xang || !yang || !phase || !phasec || xdim < 1 || ydim < 1 || maxiter < 0) { set_error("pxf_reconstruct: bad argument"); return PXF_ERR_INVALID; }
    if (sm_count() <= 0) { set_error("no CUDA device available (libpxf has no CPU fallback)"); return PXF_ERR_CUDA; }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const size_t bytes = (size_t)xdim * ydim * sizeof(double);
    const double pi = (double)3.1415926535897931f;          // reconstruct.f95:13, a default-real literal
    const double w = 2 / (1 + sin(pi / (sqrt((double)xdim * (double)ydim) + 1)));
    const int64_t max_sweeps = (int64_t)maxiter + 1;
    PXF_CUDA(cudaMemcpyAsync(phasec, phase, bytes, cudaMemcpyDeviceToDevice, s));        // :17
    int64_t done = 0;
    if (xdim >= 3 && ydim >= 3) {
        const int mind = xdim < ydim ? xdim : ydim;
        const int nchunk = (mind - 2 + 31) / 32;                    // chunks of 32 cells on the longest diagonal
        const int ndiag = xdim + ydim - 5;                          // pipeline depth in time steps
        const int every = ndiag / 2 + 2;                            // snapshot interval (sweeps)
        const int ring = ndiag / 2 + 4;                             // sweeps in flight (+ the one being tested)
        // warps that can be busy in one time step: every second diagonal, nchunk items each
        const int64_t peak_items = (int64_t)((ndiag + 1) / 2) * nchunk;
        int ctas = (int)((peak_items * 32 + RECON_THREADS - 1) / RECON_THREADS);
        int coop = 0, maxb = 0;
        int devid = 0;
        cudaGetDevice(&devid);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, devid);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxb, k_reconstruct<true>, RECON_THREADS, 0) != cudaSuccess) { cudaGetLastError(); maxb = 0; }
        const int cap_ctas = sm_count() * (maxb > 0 ? 1 : 0);       // one CTA per SM keeps the grid barrier cheap
        if (ctas > cap_ctas) ctas = cap_ctas;
        const bool use_coop = coop && ctas >= 2;
        Scratch sc;
        int rc = sc.alloc(9 * bytes + (size_t)ring * nchunk * (sizeof(double) + sizeof(int)) + 512, s);
        if (rc) return rc;
        char *p = static_cast<char *>(sc.p);
        double *init[3], *snapb[2][3];
        for (int k = 0; k < 3; k++) { init[k] = (double *)p; p += bytes; }
        for (int b2 = 0; b2 < 2; b2++)
            for (int k = 0; k < 3; k++) { snapb[b2][k] = (double *)p; p += bytes; }
        double *acc = (double *)p; p += (size_t)ring * nchunk * sizeof(double);
        int *cnt = (int *)p; p += (size_t)ring * nchunk * sizeof(int);
        ReconCtl *ctl = (ReconCtl *)(((uintptr_t)p + 15) & ~(uintptr_t)15);
        double *live[3] = {xang, yang, phasec};
        // the initial state is snapshot 0; the borders (never written by a sweep) of both snapshot buffers come from it
        for (int k = 0; k < 3; k++) {
            PXF_CUDA(cudaMemcpyAsync(init[k], live[k], bytes, cudaMemcpyDeviceToDevice, s));
            PXF_CUDA(cudaMemcpyAsync(snapb[0][k], live[k], bytes, cudaMemcpyDeviceToDevice, s));
            PXF_CUDA(cudaMemcpyAsync(snapb[1][k], live[k], bytes, cudaMemcpyDeviceToDevice, s));
        }
        auto run = [&](int ns, double crit, int snap_every) -> int {
            double *xa = xang, *ya = yang, *pc = phasec;
            int xd = xdim, yd = ydim, nss = ns, nch = nchunk, rg = ring;
            double ww = w, hh = h, cr = crit;
            ReconSnap sn;
            for (int b2 = 0; b2 < 2; b2++) { sn.x[b2] = snapb[b2][0]; sn.y[b2] = snapb[b2][1]; sn.p[b2] = snapb[b2][2]; }
            sn.every = snap_every;
            if (use_coop) {
                void *args[] = {&xa, &ya, &pc, &xd, &yd, &ww, &hh, &nss, &nch, &rg, &acc, &cnt, &cr, &ctl, &sn};
                cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_reconstruct<true>, dim3(ctas), dim3(RECON_THREADS), args, 0, s);
                if (e != cudaSuccess) { set_error("k_reconstruct (cooperative): %s", cudaGetErrorString(e)); return PXF_ERR_CUDA; }
            } else {
                k_reconstruct<false><<<1, RECON_THREADS, 0, s>>>(xa, ya, pc, xd, yd, ww, hh, nss, nch, rg, acc, cnt, cr, ctl, sn);
            }
            count_launch();
            return check_launch("k_reconstruct");
        };
        // one pipelined run over all the sweeps the reference could make; it stops itself two steps after the
        // first converged sweep has been recognised
        const int ns_all = (int)(max_sweeps < 0x1fffffff ? max_sweeps : 0x1fffffff);   // 2*ns_all time steps must fit an int
        if ((rc = run(ns_all, criteria, every))) return rc;
        ReconCtl hc;
        PXF_CUDA(cudaMemcpyAsync(&hc, ctl, sizeof(hc), cudaMemcpyDeviceToHost, s));
        PXF_CUDA(cudaStreamSynchronize(s));
        if (hc.stop < 0 || hc.stop == ns_all - 1) {
            done = ns_all;                                          // ran to the cap (or converged on the very last sweep): state is exact
        } else {
            done = (int64_t)hc.stop + 1;
            const int j = (int)(done / every);                      // latest snapshot at or before `done` sweeps
            double **from = j == 0 ? init : snapb[j & 1];
            for (int k = 0; k < 3; k++) PXF_CUDA(cudaMemcpyAsync(live[k], from[k], bytes, cudaMemcpyDeviceToDevice, s));
            const int missing = (int)(done - (int64_t)j * every);
            if (missing > 0 && (rc = run(missing, -1., 0))) return rc;      // criteria -1: never stops early
        }
    } else {
        // no interior cell: every sweep is empty, rms = 0/0 = NaN, the loop runs to maxiter
        done = max_sweeps;
    }
    PXF_CUDA(cudaMemcpyAsync(phase, phasec, bytes, cudaMemcpyDeviceToDevice, s));        // :114
    PXF_CUDA(cudaStreamSynchronize(s));
    if (sweeps_host) *sweeps_host = done;
    return PXF_OK;
}

size_t pxf_southwellbin_scratch_bytes(int64_t num, int32_t xdim, int32_t ydim)
{
    const size_t n = (size_t)(num > 0 ? num : 1);
    return 3 * ((n * 8 + 255) & ~(size_t)255) + pxf_sort_scratch_bytes(num) + (((size_t)xdim * ydim * 8 + 255) & ~(size_t)255) + 1024;
}

// x,y,l,m: device rows of the bundle; xang,yang,phase: device, column-major [xdim][ydim]
int pxf_southwellbin(const double *x, const double *y, const double *l, const double *m, int64_t num, double binsize,
                     double *xang, double *yang, double *phase, int32_t xdim, int32_t ydim, void *scratch,
                     pxf_stream_t stream)
{
    if (num < 0 || xdim < 1 || ydim < 1 || !xang || !yang || !phase || !scratch || (num > 0 && (!x || !y || !l || !m))) {
        set_error("pxf_southwellbin: bad argument");
        return PXF_ERR_INVALID;
    }
    if (sm_count() <= 0) { set_error("no CUDA device available (libpxf has no CPU fallback)"); return PXF_ERR_CUDA; }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const long long ncell = (long long)xdim * ydim;
    const size_t n = (size_t)(num > 0 ? num : 1);
    const size_t rb = (n * 8 + 255) & ~(size_t)255;
    char *p = static_cast<char *>(scratch);
    double *key = (double *)p; p += rb;
    double *skey = (double *)p; p += rb;
    long long *idx = (long long *)p; p += rb;
    void *sort_scr = p; p += pxf_sort_scratch_bytes(num);
    long long *start = (long long *)p;
    PXF_CUDA(cudaMemsetAsync(start, 0xff, (size_t)ncell * 8, s));            // -1
    if (num > 0) {
        k_swb_cell<<<grid_for(num, PXF_BLOCK, 8), PXF_BLOCK, 0, s>>>(x, y, num, binsize, xdim, ydim, key);
        count_launch();
        int rc = pxf_argsort(key, num, skey, reinterpret_cast<int64_t *>(idx), sort_scr, stream);   // stable: ray order kept
        if (rc) return rc;
        k_swb_starts<<<grid_for(num, PXF_BLOCK, 8), PXF_BLOCK, 0, s>>>(skey, num, start, ncell);
        count_launch();
    }
    k_swb_sum<<<(int)((ncell + PXF_BLOCK - 1) / PXF_BLOCK), PXF_BLOCK, 0, s>>>(skey, idx, num, start, l, m, ncell, xang, yang, phase);
    count_launch();
    return check_launch("pxf_southwellbin");
}

}  // extern "C"
