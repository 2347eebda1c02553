// Weighted half-power diameter without sorting the bundle (analyses.hpd weighted branch,
// analyses.py:88-97: r,cdf = rhocdf(...); r[argmin|cdf-.75|] - r[argmin|cdf-.25|]).
//
// The reference argsorts all radii, cumsums the permuted weights and takes two argmins.  The two
// argmins sit where the cdf crosses .25 and .75, so only the radii next to the two crossings have to
// be ordered exactly:
//   1. weighted centroid sums (24 B/ray, k_sums) -> centroid and total weight W on the device;
//   2. a strided sample of (radius, weight) pairs is sorted (pxf_argsort on <= 2^18 keys) and its
//      weighted quantiles at q -+ delta give one bracket [lo,hi] per crossing; delta = Z standard
//      errors of a weighted sample cdf (Kish design factor measured on the sample);
//   3. ONE pass over the bundle (k_wq_collect, 24 B/ray) sums the weight strictly below each bracket
//      (fixed-shape tree: deterministic) and appends the ~1-2 % of (radius, weight) pairs inside each
//      bracket to a candidate buffer;
//   4. each candidate buffer is sorted, its weights are cumsummed in sorted order on top of the
//      weight below the bracket, and argmin |cdf - q| (first minimiser, like numpy) is taken inside.
// The window result equals the global argmin when the cdf of the first candidate is < q and that of
// the last is > q (the cdf is non-decreasing for weights >= 0, so |cdf-q| falls then rises and its
// minimum is next to the crossing).  Anything else -- bracket miss, buffer overflow, a negative or NaN
// weight, a NaN radius -- clears `valid` and the caller runs the full sort.  Radii are bit-identical
// to k_rho's; the cdf differs from a sequential np.cumsum only by summation order.
#include <string.h>
#include "pxf_internal.h"
#include "pxf_ray.cuh"

namespace pxf {

typedef unsigned long long u64;

#define WQ_Z 6.0                 // bracket half width in standard errors of the sample cdf
#define WQ_SCAP 2048             // shared staging capacity per bracket (pairs)
#define WQ_U 2                   // independent 16-byte load triples per thread per batch

struct WqState {
    double lohi[4];              // lo25, hi25, lo75, hi75 (radii; brackets are closed)
    double below[2];             // weight of the rays with r < lo_b
    u64 count[2];                // candidates appended (may exceed the capacity)
    u64 nbad;                    // NaN radii, NaN or negative weights
    double design;               // Kish design factor of the sample (diagnostic)
};

// ---------------------------------------------------------------- 2. sample
__global__ void __launch_bounds__(PXF_BLOCK)
k_wq_sample(const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ w, int64_t num,
            const double *__restrict__ cxy, int nsamp, double *__restrict__ rs, double *__restrict__ ws)
{
    const double cx = cxy[0], cy = cxy[1];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nsamp; j += gridDim.x * blockDim.x) {
        const int64_t i = (int64_t)(((u64)j * (u64)num) / (u64)nsamp);
        rs[j] = sqrt(sq(x[i] - cx) + sq(y[i] - cy));
        ws[j] = w[i];
    }
}

// first index with cum[i] >= target (n if none)
PXF_DEV int wq_lower_bound(const double *cum, int n, double target)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cum[mid] >= target) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// One CTA.  rs: sorted sample radii, cum: inclusive prefix sums of the sample weights in that order.
__global__ void __launch_bounds__(PXF_BLOCK)
k_wq_brackets(const double *__restrict__ rs, const double *__restrict__ cum, int n, double z, long long lowmask,
              WqState *__restrict__ st)
{
    __shared__ double sh[PXF_BLOCK / 32];
    // Kish design factor n*sum(w^2)/W^2 from <= 16384 evenly spaced weights (it only sets the bracket width)
    const int stride = n > 16384 ? n / 16384 : 1;
    const int m = (n + stride - 1) / stride;
    double s2 = 0.;
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        const int i = j * stride;
        const double wi = cum[i] - (i ? cum[i - 1] : 0.);
        s2 += wi * wi;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s2 += __shfl_down_sync(0xffffffffu, s2, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s2;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.;
        for (int q = 0; q < PXF_BLOCK / 32; q++) t += sh[q];
        t *= (double)n / (double)m;
        const double W = cum[n - 1];
        double design = (double)n * t / (W * W);           // >= 1; 1 for equal weights
        if (!(design >= 1.)) design = 1.;                  // also catches NaN
        const double inf = __longlong_as_double(0x7ff0000000000000ll);
        const double qs[2] = {.25, .75};
        for (int b = 0; b < 2; b++) {
            const double q = qs[b];
            const double delta = z * sqrt(q * (1. - q) * design / (double)n) + 2. / (double)n;
            double lo = 0., hi = inf;
            if (q - delta > 0.) {
                const int p = wq_lower_bound(cum, n, (q - delta) * W);
                lo = p > 0 ? rs[p - 1] : 0.;
                // the sample is ordered by the top `keybits` bits of the radius pattern only: round down
                lo = __longlong_as_double(__double_as_longlong(lo) & ~lowmask);
            }
            if (q + delta < 1.) {
                const int p = wq_lower_bound(cum, n, (q + delta) * W);
                hi = p + 1 < n ? rs[p + 1] : inf;
                if (hi < inf) hi = __longlong_as_double(__double_as_longlong(hi) | lowmask);   // ... and up
                if (!(hi <= inf)) hi = inf;
            }
            if (!(lo == lo)) lo = 0.;
            if (!(hi == hi)) hi = inf;
            st->lohi[2 * b] = lo;
            st->lohi[2 * b + 1] = hi;
        }
        st->below[0] = st->below[1] = 0.;
        st->count[0] = st->count[1] = 0ull;
        st->nbad = 0ull;
        st->design = design;
    }
}

// ---------------------------------------------------------------- 3. collect
struct WqStage {
    double r[2][WQ_SCAP];
    double w[2][WQ_SCAP];
};

template <bool VEC2>
__global__ void __launch_bounds__(PXF_BLOCK)
k_wq_collect(const double *__restrict__ x, const double *__restrict__ y, const double *__restrict__ w, int64_t num,
             const double *__restrict__ cxy, WqState *__restrict__ st,
             double *__restrict__ cr0, double *__restrict__ cw0, double *__restrict__ cr1, double *__restrict__ cw1,
             u64 cap, double *__restrict__ partial /*[grid][2]*/)
{
    extern __shared__ __align__(16) unsigned char wq_smem[];
    WqStage &S = *reinterpret_cast<WqStage *>(wq_smem);
    __shared__ unsigned int scount[2];
    __shared__ u64 sbase[2];
    __shared__ double shb[2][PXF_BLOCK / 32];
    const double cx = cxy[0], cy = cxy[1];
    const double lo0 = st->lohi[0], hi0 = st->lohi[1], lo1 = st->lohi[2], hi1 = st->lohi[3];
    double below0 = 0., below1 = 0.;
    unsigned int bad = 0;
    constexpr int PER = VEC2 ? 2 : 1;
    constexpr int NE = WQ_U * PER;
    static_assert(NE * PXF_BLOCK <= WQ_SCAP / 2, "staging buffer too small for one batch");
    const int64_t items = VEC2 ? (num >> 1) : num;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * WQ_U;
    if (threadIdx.x < 2) scount[threadIdx.x] = 0;
    __syncthreads();
    auto flush = [&]() {                 // block-uniform
        if (threadIdx.x < 2) sbase[threadIdx.x] = atomicAdd(&st->count[threadIdx.x], (u64)scount[threadIdx.x]);
        __syncthreads();
        for (int b = 0; b < 2; b++) {
            const unsigned int n = scount[b];
            double *cr = b ? cr1 : cr0, *cw = b ? cw1 : cw0;
            for (unsigned int t = threadIdx.x; t < n; t += blockDim.x) {
                const u64 dst = sbase[b] + t;
                if (dst < cap) { cr[dst] = S.r[b][t]; cw[dst] = S.w[b][t]; }
            }
        }
        __syncthreads();
        if (threadIdx.x < 2) scount[threadIdx.x] = 0;
        __syncthreads();
    };
    auto classify = [&](double r, double wt) {
        if (r != r || !(wt >= 0.)) { bad++; return; }
        if (r < lo0) below0 += wt;
        else if (r <= hi0) { const unsigned int t = atomicAdd(&scount[0], 1u); S.r[0][t] = r; S.w[0][t] = wt; }
        if (r < lo1) below1 += wt;
        else if (r <= hi1) { const unsigned int t = atomicAdd(&scount[1], 1u); S.r[1][t] = r; S.w[1][t] = wt; }
    };
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x * WQ_U; base < items; base += stride) {
        if (VEC2) {
            double2 xv[WQ_U], yv[WQ_U], wv[WQ_U];
#pragma unroll
            for (int u = 0; u < WQ_U; u++) {
                const int64_t q = base + (int64_t)u * blockDim.x + threadIdx.x;
                if (q < items) {
                    xv[u] = *reinterpret_cast<const double2 *>(x + 2 * q);
                    yv[u] = *reinterpret_cast<const double2 *>(y + 2 * q);
                    wv[u] = *reinterpret_cast<const double2 *>(w + 2 * q);
                } else {
                    xv[u] = yv[u] = wv[u] = make_double2(0., 0.);
                }
            }
#pragma unroll
            for (int u = 0; u < WQ_U; u++) {
                const int64_t q = base + (int64_t)u * blockDim.x + threadIdx.x;
                if (q < items) {
                    classify(sqrt(sq(xv[u].x - cx) + sq(yv[u].x - cy)), wv[u].x);
                    classify(sqrt(sq(xv[u].y - cx) + sq(yv[u].y - cy)), wv[u].y);
                }
            }
        } else {
            double xv[WQ_U], yv[WQ_U], wv[WQ_U];
#pragma unroll
            for (int u = 0; u < WQ_U; u++) {
                const int64_t q = base + (int64_t)u * blockDim.x + threadIdx.x;
                xv[u] = q < items ? x[q] : 0.;
                yv[u] = q < items ? y[q] : 0.;
                wv[u] = q < items ? w[q] : 0.;
            }
#pragma unroll
            for (int u = 0; u < WQ_U; u++) {
                const int64_t q = base + (int64_t)u * blockDim.x + threadIdx.x;
                if (q < items) classify(sqrt(sq(xv[u] - cx) + sq(yv[u] - cy)), wv[u]);
            }
        }
        __syncthreads();
        if (scount[0] > WQ_SCAP / 2 || scount[1] > WQ_SCAP / 2) flush();
    }
    if (VEC2 && (num & 1) && blockIdx.x == 0 && threadIdx.x == 0)
        classify(sqrt(sq(x[num - 1] - cx) + sq(y[num - 1] - cy)), w[num - 1]);
    __syncthreads();
    if (scount[0] > 0 || scount[1] > 0) flush();
    // fixed-shape reduction of the weights below the brackets
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        below0 += __shfl_down_sync(0xffffffffu, below0, o);
        below1 += __shfl_down_sync(0xffffffffu, below1, o);
        bad += __shfl_down_sync(0xffffffffu, bad, o);
    }
    if ((threadIdx.x & 31) == 0) {
        shb[0][threadIdx.x >> 5] = below0;
        shb[1][threadIdx.x >> 5] = below1;
        if (bad) atomicAdd(&st->nbad, (u64)bad);
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        double t = 0.;
        for (int q = 0; q < PXF_BLOCK / 32; q++) t += shb[threadIdx.x][q];
        partial[2 * blockIdx.x + threadIdx.x] = t;
    }
}

__global__ void k_wq_below_final(const double *__restrict__ partial, int nblocks, WqState *__restrict__ st)
{
    // pairwise tree over the block partials in a fixed order (one warp per bracket)
    const int b = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (b >= 2) return;
    double acc = 0.;
    for (int i = lane; i < nblocks; i += 32) acc += partial[2 * i + b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (lane == 0) st->below[b] = acc;
}

// ---------------------------------------------------------------- 4. argmin inside the window
struct WqArg { double v; long long i; };
PXF_DEV bool wq_better(double v, long long i, double bv, long long bi)
{
    const bool vn = v != v, bn = bv != bv;
    if (vn || bn) { if (vn && bn) return i < bi; return vn; }
    return v < bv || (v == bv && i < bi);
}

__global__ void __launch_bounds__(PXF_BLOCK)
k_wq_argmin(const double *__restrict__ cum, int64_t n, const double *__restrict__ below_ptr,
            const double *__restrict__ total_ptr, double q, WqArg *__restrict__ partial)
{
    const double P = *below_ptr, W = *total_ptr;
    double bv = __longlong_as_double(0x7ff0000000000000ll);
    long long bi = 0x7fffffffffffffffll;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < n; i += nthr) {
        const double d = fabs((P + cum[i]) / W - q);
        if (wq_better(d, i, bv, bi)) { bv = d; bi = i; }
    }
    __shared__ double shv[PXF_BLOCK / 32];
    __shared__ long long shi[PXF_BLOCK / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, bv, o);
        const long long oi = __shfl_down_sync(0xffffffffu, bi, o);
        if (wq_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { shv[threadIdx.x >> 5] = bv; shi[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < PXF_BLOCK / 32; k++)
            if (wq_better(shv[k], shi[k], bv, bi)) { bv = shv[k]; bi = shi[k]; }
        partial[blockIdx.x].v = bv;
        partial[blockIdx.x].i = bi;
    }
}

// out[0] = r[argmin], out[1] = valid (1/0), out[2] = cdf at the argmin, out[3] = argmin index in the window
__global__ void k_wq_result(const WqArg *__restrict__ partial, int nblk, const double *__restrict__ rs,
                            const double *__restrict__ cum, int64_t n, const double *__restrict__ below_ptr,
                            const double *__restrict__ total_ptr, double q, double *__restrict__ out)
{
    // one warp
    const int lane = threadIdx.x & 31;
    double v = __longlong_as_double(0x7ff0000000000000ll);
    long long i = 0x7fffffffffffffffll;
    for (int b = lane; b < nblk; b += 32)
        if (wq_better(partial[b].v, partial[b].i, v, i)) { v = partial[b].v; i = partial[b].i; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, v, o);
        const long long oi = __shfl_down_sync(0xffffffffu, i, o);
        if (wq_better(ov, oi, v, i)) { v = ov; i = oi; }
    }
    if (lane) return;
    const double P = *below_ptr, W = *total_ptr;
    const double cfirst = (P + cum[0]) / W, clast = (P + cum[n - 1]) / W;
    const bool ok = i >= 0 && i < n && v == v && cfirst < q && clast > q && W > 0.;
    out[0] = ok ? rs[i] : __longlong_as_double(0x7ff8000000000000ll);
    out[1] = ok ? 1. : 0.;
    out[2] = ok ? (P + cum[i]) / W : 0.;
    out[3] = (double)i;
}

// ---------------------------------------------------------------- sharded merge: (K+1)-ary key-space search
// dist.weighted_quantile_radii on the device.  Each rank holds, per quantile i (NQ = 2), a sorted run of radius
// patterns keys[i][0..n[i]) with the prefix sums of their weights.  The GLOBAL answer -- smallest key whose
// global cumulative weight reaches q*W -- is found without moving a ray: per step K pivots split [lo,hi], every
// rank reports its weight at or below each pivot (k_wq_merge_probe), ONE all-reduce of [2][K] doubles merges
// them, and k_wq_merge_narrow shrinks [lo,hi] by a factor K+1 identically on every rank.
#define WQ_MERGE_MAXK 255
struct WqMerge {
    long long lo[2], hi[2];      // invariant: answer in [lo, hi]
    long long klo[2];            // final phase: largest local key below hi (or -1)
    double res[2], valid[2];     // final result
};

PXF_DEV long long wq_pivot(long long lo, long long hi, int j, int K)
{
    // j-th of K pivots splitting [lo,hi] into K+1 nearly equal parts (the torch driver's formula)
    const long long n = hi - lo + 1;
    const long long step = n / (K + 1), rem = n % (K + 1);
    long long p = lo + (long long)(j + 1) * step + ((long long)(j + 1) < rem ? (long long)(j + 1) : rem) - 1;
    return p < hi ? p : hi;
}
// number of keys <= key (strict: < key)
PXF_DEV long long wq_count(const long long *__restrict__ keys, long long n, long long key, bool strict)
{
    long long lo = 0, hi = n;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        const long long v = keys[mid];
        if (strict ? v < key : v <= key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

struct WqRuns { const long long *keys[2]; const double *cum[2]; long long n[2]; };

__global__ void k_wq_merge_begin(WqMerge *m, long long lo0, long long hi0, long long lo1, long long hi1)
{
    m->lo[0] = lo0; m->hi[0] = hi0; m->lo[1] = lo1; m->hi[1] = hi1;
    m->klo[0] = m->klo[1] = -1;
    m->res[0] = m->res[1] = 0.; m->valid[0] = m->valid[1] = 0.;
}

// thread t = i*K + j
__global__ void k_wq_merge_probe(const WqRuns R, const WqMerge *__restrict__ m, int K, double *__restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * K) return;
    const int i = t / K, j = t % K;
    const long long key = wq_pivot(m->lo[i], m->hi[i], j, K);
    const long long cnt = wq_count(R.keys[i], R.n[i], key, false);
    out[t] = cnt > 0 ? R.cum[i][cnt - 1] : 0.;
}

// one warp per quantile
__global__ void k_wq_merge_narrow(WqMerge *m, const double *__restrict__ sum, const double *__restrict__ offsets,
                                  const double *__restrict__ total, double q0, double q1, int K)
{
    const int i = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (i >= 2) return;
    const double q = i ? q1 : q0, W = *total, off = offsets ? offsets[i] : 0.;
    const long long lo = m->lo[i], hi = m->hi[i];
    int first = K;                                     // first pivot whose global cdf reaches q
    for (int j = lane; j < K; j += 32)
        if ((sum[i * K + j] + off) / W >= q) { first = j; break; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int other = __shfl_down_sync(0xffffffffu, first, o);
        first = other < first ? other : first;
    }
    if (lane == 0) {
        long long nlo, nhi;
        if (first < K) {
            nhi = wq_pivot(lo, hi, first, K);
            nlo = first > 0 ? wq_pivot(lo, hi, first - 1, K) + 1 : lo;
        } else {
            nhi = hi;
            nlo = wq_pivot(lo, hi, K - 1, K) + 1;
        }
        if (nlo > nhi) nlo = nhi;
        m->lo[i] = nlo; m->hi[i] = nhi;
    }
}

// final phase, local part: out[i*2+0] = weight at or below hi, out[i*2+1] = weight strictly below hi,
// m->klo[i] = largest local key below hi (or -1)
__global__ void k_wq_merge_final_probe(const WqRuns R, WqMerge *m, double *__restrict__ out)
{
    const int i = threadIdx.x;
    if (i >= 2) return;
    const long long key = m->hi[i];
    const long long le = wq_count(R.keys[i], R.n[i], key, false), lt = wq_count(R.keys[i], R.n[i], key, true);
    out[2 * i] = le > 0 ? R.cum[i][le - 1] : 0.;
    out[2 * i + 1] = lt > 0 ? R.cum[i][lt - 1] : 0.;
    m->klo[i] = lt > 0 ? R.keys[i][lt - 1] : -1;
}

// after all-reduce(sum) of out and all-reduce(max) of klo
__global__ void k_wq_merge_finish(WqMerge *m, const double *__restrict__ sum, const double *__restrict__ offsets,
                                  const double *__restrict__ total, double q0, double q1)
{
    const int i = threadIdx.x;
    if (i >= 2) return;
    const double q = i ? q1 : q0, W = *total, off = offsets ? offsets[i] : 0.;
    const double chi = (sum[2 * i] + off) / W, clo = (sum[2 * i + 1] + off) / W;
    const long long klo = m->klo[i], khi = m->hi[i];
    const bool take_lo = klo >= 0 && fabs(clo - q) <= fabs(chi - q);
    m->res[i] = __longlong_as_double(take_lo ? klo : khi);
    m->valid[i] = (chi >= q && (klo >= 0 || off == 0.)) ? 1. : 0.;
}

static size_t a256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace pxf

using namespace pxf;

extern "C" {

int pxf_hpd_weighted_sorted(const double *x, const double *y, const double *w, int64_t num, double *hpd_host,
                            pxf_stream_t stream);   // pxf_sort.cu

size_t pxf_wq_state_bytes(void) { return a256(sizeof(WqState)); }
int64_t pxf_wq_min_num(void) { return (int64_t)1 << 21; }
int32_t pxf_wq_samples(int64_t num)
{
    int64_t n = num / 16;
    if (n < 65536) n = 65536;
    if (n > 262144) n = 262144;
    return (int32_t)n;
}
int64_t pxf_wq_capacity(int64_t num) { int64_t c = num / 8; return c < 65536 ? 65536 : c; }

int pxf_wq_sample(const double *x, const double *y, const double *w, int64_t num, const double *cxy_dev,
                  int32_t nsamp, double *rs_out, double *ws_out, pxf_stream_t stream)
{
    if (num <= 0 || nsamp <= 0 || !x || !y || !w || !cxy_dev || !rs_out || !ws_out) { set_error("pxf_wq_sample: bad argument"); return PXF_ERR_INVALID; }
    if (sm_count() <= 0) { set_error("no CUDA device available (libpxf has no CPU fallback)"); return PXF_ERR_CUDA; }
    k_wq_sample<<<(nsamp + PXF_BLOCK - 1) / PXF_BLOCK, PXF_BLOCK, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        x, y, w, num, cxy_dev, nsamp, rs_out, ws_out);
    count_launch();
    return check_launch("k_wq_sample");
}

int pxf_wq_brackets(const double *rs_sorted, const double *cum, int32_t nsamp, int32_t keybits, void *state,
                    pxf_stream_t stream)
{
    if (nsamp <= 0 || !rs_sorted || !cum || !state || keybits < 16 || keybits > 64) { set_error("pxf_wq_brackets: bad argument"); return PXF_ERR_INVALID; }
    const long long lowmask = keybits >= 64 ? 0ll : (long long)((~0ull) >> keybits);
    k_wq_brackets<<<1, PXF_BLOCK, 0, reinterpret_cast<cudaStream_t>(stream)>>>(rs_sorted, cum, nsamp, WQ_Z, lowmask,
                                                                               static_cast<WqState *>(state));
    count_launch();
    return check_launch("k_wq_brackets");
}

size_t pxf_wq_collect_scratch_bytes(void) { return (size_t)2 * 8 * (size_t)(sm_count() > 0 ? sm_count() * 8 : 8192); }

int pxf_wq_collect(const double *x, const double *y, const double *w, int64_t num, const double *cxy_dev, void *state,
                   double *cand_r0, double *cand_w0, double *cand_r1, double *cand_w1, int64_t cap, void *scratch,
                   pxf_stream_t stream)
{
    if (num < 0 || !cxy_dev || !state || !cand_r0 || !cand_w0 || !cand_r1 || !cand_w1 || cap <= 0 || !scratch) {
        set_error("pxf_wq_collect: bad argument");
        return PXF_ERR_INVALID;
    }
    if (sm_count() <= 0) { set_error("no CUDA device available (libpxf has no CPU fallback)"); return PXF_ERR_CUDA; }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    WqState *st = static_cast<WqState *>(state);
    double *partial = static_cast<double *>(scratch);
    static bool attr[64] = {};
    if (first_on_device(attr)) {
        PXF_CUDA(cudaFuncSetAttribute(k_wq_collect<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WqStage)));
        PXF_CUDA(cudaFuncSetAttribute(k_wq_collect<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WqStage)));
    }
    int grid = 1;
    if (num > 0) {
        const bool aligned = (((uintptr_t)x | (uintptr_t)y | (uintptr_t)w) & 15) == 0;
        const int maxg = sm_count() * 3;               // 64 KB of staging per CTA: 3 resident CTAs per SM
        if (aligned) {
            grid = grid_for((num + 1) >> 1, PXF_BLOCK * WQ_U, 3);
            if (grid > maxg) grid = maxg;
            k_wq_collect<true><<<grid, PXF_BLOCK, sizeof(WqStage), s>>>(x, y, w, num, cxy_dev, st, cand_r0, cand_w0,
                                                                       cand_r1, cand_w1, (u64)cap, partial);
        } else {
            grid = grid_for(num, PXF_BLOCK * WQ_U, 3);
            if (grid > maxg) grid = maxg;
            k_wq_collect<false><<<grid, PXF_BLOCK, sizeof(WqStage), s>>>(x, y, w, num, cxy_dev, st, cand_r0, cand_w0,
                                                                        cand_r1, cand_w1, (u64)cap, partial);
        }
        count_launch();
    } else {
        PXF_CUDA(cudaMemsetAsync(partial, 0, 2 * sizeof(double), s));
    }
    k_wq_below_final<<<1, 64, 0, s>>>(partial, grid, st);
    count_launch();
    return check_launch("k_wq_collect");
}

double *pxf_wq_below_ptr(void *state, int32_t b) { return &static_cast<WqState *>(state)->below[b ? 1 : 0]; }

/* argmin |(below + cum[i]) / total - q| over a sorted candidate window; out_dev[4] = [r, valid, cdf, index] */
int pxf_wq_argmin(const double *rs_sorted, const double *cum, int64_t n, const double *below_dev, const double *total_dev,
                  double q, double *out_dev, void *scratch, pxf_stream_t stream)
{
    if (n <= 0 || !rs_sorted || !cum || !below_dev || !total_dev || !out_dev || !scratch) { set_error("pxf_wq_argmin: bad argument"); return PXF_ERR_INVALID; }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    int grid = grid_for(n, PXF_BLOCK * 4, 4);
    if (grid > 1024) grid = 1024;
    WqArg *partial = static_cast<WqArg *>(scratch);
    k_wq_argmin<<<grid, PXF_BLOCK, 0, s>>>(cum, n, below_dev, total_dev, q, partial);
    k_wq_result<<<1, 32, 0, s>>>(partial, grid, rs_sorted, cum, n, below_dev, total_dev, q, out_dev);
    count_launch(2);
    return check_launch("k_wq_argmin");
}
size_t pxf_wq_argmin_scratch_bytes(void) { return 1024 * sizeof(WqArg); }

/* ---- sharded merge of sorted runs (dist.weighted_quantile_radii on the device) ---- */
size_t pxf_wq_merge_state_bytes(void) { return a256(sizeof(WqMerge)); }
int64_t *pxf_wq_merge_klo_ptr(void *state) { return reinterpret_cast<int64_t *>(static_cast<WqMerge *>(state)->klo); }
double *pxf_wq_merge_result_ptr(void *state) { return static_cast<WqMerge *>(state)->res; }

static int wq_runs(WqRuns &R, const int64_t *keys0, const double *cum0, int64_t n0, const int64_t *keys1,
                   const double *cum1, int64_t n1)
{
    if (n0 < 0 || n1 < 0 || (n0 > 0 && (!keys0 || !cum0)) || (n1 > 0 && (!keys1 || !cum1))) { set_error("pxf_wq_merge: bad run"); return PXF_ERR_INVALID; }
    R.keys[0] = reinterpret_cast<const long long *>(keys0); R.cum[0] = cum0; R.n[0] = n0;
    R.keys[1] = reinterpret_cast<const long long *>(keys1); R.cum[1] = cum1; R.n[1] = n1;
    return PXF_OK;
}

int pxf_wq_merge_begin(void *state, int64_t lo0, int64_t hi0, int64_t lo1, int64_t hi1, pxf_stream_t stream)
{
    if (!state || lo0 < 0 || lo1 < 0 || hi0 < lo0 || hi1 < lo1) { set_error("pxf_wq_merge_begin: bad argument"); return PXF_ERR_INVALID; }
    if (sm_count() <= 0) { set_error("no CUDA device available (libpxf has no CPU fallback)"); return PXF_ERR_CUDA; }
    k_wq_merge_begin<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(static_cast<WqMerge *>(state), lo0, hi0, lo1, hi1);
    count_launch();
    return check_launch("k_wq_merge_begin");
}

int pxf_wq_merge_probe(const int64_t *keys0, const double *cum0, int64_t n0, const int64_t *keys1, const double *cum1,
                       int64_t n1, const void *state, int32_t K, double *out_dev, pxf_stream_t stream)
{
    WqRuns R;
    int rc = wq_runs(R, keys0, cum0, n0, keys1, cum1, n1);
    if (rc) return rc;
    if (!state || !out_dev || K < 1 || K > WQ_MERGE_MAXK) { set_error("pxf_wq_merge_probe: bad argument"); return PXF_ERR_INVALID; }
    k_wq_merge_probe<<<(2 * K + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        R, static_cast<const WqMerge *>(state), K, out_dev);
    count_launch();
    return check_launch("k_wq_merge_probe");
}

int pxf_wq_merge_narrow(void *state, const double *sum_dev, const double *offsets_dev, const double *total_dev,
                        double q0, double q1, int32_t K, pxf_stream_t stream)
{
    if (!state || !sum_dev || !total_dev || K < 1 || K > WQ_MERGE_MAXK) { set_error("pxf_wq_merge_narrow: bad argument"); return PXF_ERR_INVALID; }
    k_wq_merge_narrow<<<1, 64, 0, reinterpret_cast<cudaStream_t>(stream)>>>(static_cast<WqMerge *>(state), sum_dev, offsets_dev,
                                                                            total_dev, q0, q1, K);
    count_launch();
    return check_launch("k_wq_merge_narrow");
}

int pxf_wq_merge_final_probe(const int64_t *keys0, const double *cum0, int64_t n0, const int64_t *keys1,
                             const double *cum1, int64_t n1, void *state, double *out_dev, pxf_stream_t stream)
{
    WqRuns R;
    int rc = wq_runs(R, keys0, cum0, n0, keys1, cum1, n1);
    if (rc) return rc;
    if (!state || !out_dev) { set_error("pxf_wq_merge_final_probe: bad argument"); return PXF_ERR_INVALID; }
    k_wq_merge_final_probe<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(R, static_cast<WqMerge *>(state), out_dev);
    count_launch();
    return check_launch("k_wq_merge_final_probe");
}

int pxf_wq_merge_finish(void *state, const double *sum_dev, const double *offsets_dev, const double *total_dev,
                        double q0, double q1, pxf_stream_t stream)
{
    if (!state || !sum_dev || !total_dev) { set_error("pxf_wq_merge_finish: bad argument"); return PXF_ERR_INVALID; }
    k_wq_merge_finish<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(static_cast<WqMerge *>(state), sum_dev, offsets_dev,
                                                                            total_dev, q0, q1);
    count_launch();
    return check_launch("k_wq_merge_finish");
}

// Bracketed weighted HPD; *valid_host = 0 means "run the full sort".
int pxf_hpd_weighted_bracket(const double *x, const double *y, const double *w, int64_t num, double *hpd_host,
                             int32_t *valid_host, pxf_stream_t stream)
{
    if (num <= 0 || !x || !y || !w || !hpd_host || !valid_host) { set_error("pxf_hpd_weighted_bracket: bad argument"); return PXF_ERR_INVALID; }
    if (sm_count() <= 0) { set_error("no CUDA device available (libpxf has no CPU fallback)"); return PXF_ERR_CUDA; }
    *valid_host = 0;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    int rc;
    const int nsamp = num < pxf_wq_samples(num) ? (int)num : pxf_wq_samples(num);
    const int64_t cap = pxf_wq_capacity(num);
    const size_t ns = (size_t)nsamp, nc = (size_t)cap;
    const int64_t nsort = cap > nsamp ? cap : nsamp;
    Scratch sc;
    const size_t bytes = pxf_sums_scratch_bytes() + 64 * 8 + pxf_wq_state_bytes() + a256(pxf_wq_collect_scratch_bytes()) +
                         4 * a256(ns * 8) + 4 * a256(nc * 8) + 3 * a256((size_t)nsort * 8) +
                         a256(pxf_sort_scratch_bytes(nsort)) + a256(pxf_scan_scratch_bytes(nsort)) +
                         a256(pxf_wq_argmin_scratch_bytes()) + 4096;
    if ((rc = sc.alloc(bytes, s))) return rc;
    char *p = static_cast<char *>(sc.p);
    void *sum_scr = p; p += pxf_sums_scratch_bytes();
    double *sums = (double *)p; p += 32 * 8;            // [0]=W, [1]=sum w x, [2]=sum w y
    double *small = (double *)p; p += 32 * 8;           // [0..1]=cxy, [8..11]/[12..15]=argmin results
    void *state = p; p += pxf_wq_state_bytes();
    void *col_scr = p; p += a256(pxf_wq_collect_scratch_bytes());
    double *srs = (double *)p; p += a256(ns * 8);
    double *sws = (double *)p; p += a256(ns * 8);
    double *cand[4];
    for (int k = 0; k < 4; k++) { cand[k] = (double *)p; p += a256(nc * 8); }
    double *sorted = (double *)p; p += a256((size_t)nsort * 8);
    int64_t *idx = (int64_t *)p; p += a256((size_t)nsort * 8);
    double *cum = (double *)p; p += a256((size_t)nsort * 8);
    void *sort_scr = p; p += a256(pxf_sort_scratch_bytes(nsort));
    void *scan_scr = p; p += a256(pxf_scan_scratch_bytes(nsort));
    void *am_scr = p;
    // 1. weighted centroid and total weight, kept on the device
    if ((rc = pxf_sums(PXF_SUMS_CENTROID, x, y, nullptr, nullptr, nullptr, w, num, 0., 0., sums, sum_scr, stream))) return rc;
    if ((rc = pxf_centroid_from_sums(sums, small, stream))) return rc;
    // 2. sample -> brackets
    if ((rc = pxf_wq_sample(x, y, w, num, small, nsamp, srs, sws, stream))) return rc;
    // the sample only has to be ordered well enough for a bracket: sort on the top 32 bits of the pattern
    if ((rc = pxf_argsort_digits(srs, nsamp, sorted, idx, sort_scr, 0xF0, stream))) return rc;
    if ((rc = pxf_cumsum_gather(sws, idx, nsamp, cum, scan_scr, stream))) return rc;
    if ((rc = pxf_wq_brackets(sorted, cum, nsamp, 32, state, stream))) return rc;
    // 3. one pass over the bundle
    if ((rc = pxf_wq_collect(x, y, w, num, small, state, cand[0], cand[1], cand[2], cand[3], cap, col_scr, stream))) return rc;
    WqState h;
    PXF_CUDA(cudaMemcpyAsync(&h, state, sizeof(h), cudaMemcpyDeviceToHost, s));
    PXF_CUDA(cudaStreamSynchronize(s));
    if (h.nbad != 0 || h.count[0] == 0 || h.count[1] == 0 || h.count[0] > (u64)cap || h.count[1] > (u64)cap) return PXF_OK;
    // 4. exact order inside the two windows
    const double qs[2] = {.25, .75};
    for (int b = 0; b < 2; b++) {
        const int64_t n = (int64_t)h.count[b];
        // every candidate pattern lies in [lo, hi]: bytes above the highest differing bit are constant
        unsigned long long klo, khi;
        memcpy(&klo, &h.lohi[2 * b], 8);
        memcpy(&khi, &h.lohi[2 * b + 1], 8);
        int digits = 0;
        for (int d = 0; d < 8; d++)
            if (((klo ^ khi) >> (8 * d)) != 0) digits |= 1 << d;
        if (digits == 0) digits = 1;
        if ((rc = pxf_argsort_digits(cand[2 * b], n, sorted, idx, sort_scr, digits, stream))) return rc;
        if ((rc = pxf_cumsum_gather(cand[2 * b + 1], idx, n, cum, scan_scr, stream))) return rc;
        if ((rc = pxf_wq_argmin(sorted, cum, n, pxf_wq_below_ptr(state, b), sums, qs[b], small + 8 + 4 * b, am_scr, stream))) return rc;
    }
    double r[8];
    PXF_CUDA(cudaMemcpyAsync(r, small + 8, sizeof(r), cudaMemcpyDeviceToHost, s));
    PXF_CUDA(cudaStreamSynchronize(s));
    if (r[1] == 0. || r[5] == 0.) return PXF_OK;
    *hpd_host = r[4] - r[0];
    *valid_host = 1;
    return PXF_OK;
}

// analyses.hpd weighted branch: bracketed path for large bundles, full sort otherwise or on a miss.
int pxf_hpd_weighted(const double *x, const double *y, const double *w, int64_t num, double *hpd_host,
                     pxf_stream_t stream)
{
    if (num <= 0 || !x || !y || !w || !hpd_host) { set_error("pxf_hpd_weighted: bad argument"); return PXF_ERR_INVALID; }
    if (num >= pxf_wq_min_num()) {
        int32_t valid = 0;
        int rc = pxf_hpd_weighted_bracket(x, y, w, num, hpd_host, &valid, stream);
        if (rc) return rc;
        if (valid) return PXF_OK;
    }
    return pxf_hpd_weighted_sorted(x, y, w, num, hpd_host, stream);
}

}  // extern "C"
