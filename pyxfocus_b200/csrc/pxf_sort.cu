// Hand-written stable LSD radix sort (8-bit digits) of fp64 keys with an index payload, the
// gathered inclusive scan, and the weighted-HPD reduction built on them
// (analyses.rhocdf / hpd weighted branch, analyses.py:73-97: argsort -> cumsum -> argmin).
//
// Layout: the array is cut into G contiguous chunks, one persistent CTA per chunk
// (G = SM count x 4).  Per pass: (1) every CTA histograms the digit over its chunk,
// (2) one CTA turns the digit-major [256][G] table into exclusive offsets, (3) every CTA
// re-reads its chunk tile by tile and scatters; ranks inside a tile come from
// __match_any_sync per 32-key slice (stable), running digit offsets live in shared memory.
// Passes whose digit is constant over the whole array (typically the sign/exponent byte)
// are skipped after one up-front 8-digit histogram pass.
#include "pxf_internal.h"
#include "pxf_ray.cuh"

namespace pxf {

#define SORT_THREADS 256
#define SORT_WARPS (SORT_THREADS / 32)
#define SORT_ITEMS 8                       // 32-key slices per warp per tile
#define SORT_TILE (SORT_THREADS * SORT_ITEMS)
#define SORT_MAXG 1024

typedef unsigned long long u64;
typedef unsigned int u32;

// np.sort order: -inf < ... < -0 <= +0 < ... < +inf < NaN (all NaNs last)
PXF_DEV u64 sort_key(double v)
{
    if (v != v) return ~0ull;
    u64 b = (u64)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
PXF_DEV double unsort_key(u64 k, double nanv)
{
    if (k == ~0ull) return nanv;
    u64 b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

struct Chunking { int64_t num; int64_t per; int G; };   // chunk g = [g*per, min(num,(g+1)*per)), per % SORT_TILE == 0

// Up-front: convert keys, init indices, global histogram of all 8 digits.
__global__ void __launch_bounds__(SORT_THREADS)
k_sort_prepare(const double *__restrict__ in, u64 *__restrict__ keys, u32 *__restrict__ idx, Chunking ck,
               u64 *__restrict__ ghist /*[8][256]*/)
{
    __shared__ u32 sh[8 * 256];
    for (int t = threadIdx.x; t < 8 * 256; t += blockDim.x) sh[t] = 0;
    __syncthreads();
    const int64_t lo = (int64_t)blockIdx.x * ck.per;
    const int64_t hi = lo + ck.per < ck.num ? lo + ck.per : ck.num;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        u64 k = sort_key(in[i]);
        keys[i] = k;
        idx[i] = (u32)i;
#pragma unroll
        for (int d = 0; d < 8; d++) atomicAdd(&sh[d * 256 + (int)((k >> (8 * d)) & 255)], 1u);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 8 * 256; t += blockDim.x)
        if (sh[t]) atomicAdd(&ghist[t], (u64)sh[t]);
}

__global__ void __launch_bounds__(SORT_THREADS)
k_sort_hist(const u64 *__restrict__ keys, Chunking ck, int shift, u32 *__restrict__ table /*[256][G]*/)
{
    __shared__ u32 sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const int64_t lo = (int64_t)blockIdx.x * ck.per;
    const int64_t hi = lo + ck.per < ck.num ? lo + ck.per : ck.num;
    const int64_t span = hi > lo ? hi - lo : 0;
    const int64_t nround = (span + blockDim.x - 1) / blockDim.x;
    // four rounds at a time: the loads are issued together (one key per thread per round was latency bound:
    // 1.2 TB/s, profiles/r01h_sort_launches_summary.txt)
    for (int64_t r = 0; r < nround; r += 4) {
        int d[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int64_t i = lo + (r + u) * blockDim.x + threadIdx.x;
            d[u] = (r + u < nround && i < hi) ? (int)((keys[i] >> shift) & 255) : -1;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            unsigned peers = __match_any_sync(0xffffffffu, d[u]);
            if (d[u] >= 0 && (peers & ((1u << (threadIdx.x & 31)) - 1)) == 0) atomicAdd(&sh[d[u]], __popc(peers));
        }
    }
    __syncthreads();
    table[(size_t)threadIdx.x * ck.G + blockIdx.x] = sh[threadIdx.x];
}

// Exclusive scan of the digit-major table [256][G]: CTA d scans row d (G <= 1024 entries, one per thread) and
// writes the row total; the scatter kernel adds the exclusive scan of the 256 row totals itself.  (A single CTA
// chaining through all 256*G entries took 91 us per pass -- four times the histogram and scatter together.)
__global__ void __launch_bounds__(1024)
k_sort_scan_rows(const u32 *__restrict__ table, int G, u64 *__restrict__ offs, u64 *__restrict__ dtot)
{
    __shared__ u64 wsum[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t row = (size_t)blockIdx.x * G;
    const u64 v = (int)threadIdx.x < G ? (u64)table[row + threadIdx.x] : 0ull;
    u64 incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u64 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        u64 w = wsum[lane], iw = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u64 t = __shfl_up_sync(0xffffffffu, iw, o);
            if (lane >= o) iw += t;
        }
        wsum[lane] = iw - w;
    }
    __syncthreads();
    const u64 excl = wsum[warp] + incl - v;
    if ((int)threadIdx.x < G) offs[row + threadIdx.x] = excl;
    if (threadIdx.x == 1023) dtot[blockIdx.x] = excl + v;
}

__global__ void __launch_bounds__(SORT_THREADS)
k_sort_scatter(const u64 *__restrict__ kin, const u32 *__restrict__ vin, u64 *__restrict__ kout,
               u32 *__restrict__ vout, Chunking ck, int shift, const u64 *__restrict__ offs /*[256][G]*/,
               const u64 *__restrict__ dtot /*[256]*/)
{
    __shared__ u64 run[256];                     // running global offset per digit for this chunk
    __shared__ u32 whist[SORT_WARPS][256];       // per-warp digit counts inside the tile
    __shared__ u64 dwarp[SORT_WARPS];
    __shared__ u64 sk[SORT_TILE];                // the tile, sorted by digit
    __shared__ u32 sv[SORT_TILE];
    __shared__ u32 tbase[256];
    __shared__ u32 twarp[SORT_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {
        // digit base = exclusive scan of the 256 row totals (SORT_THREADS == 256: one digit per thread)
        const u64 v = dtot[threadIdx.x];
        u64 incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u64 t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) dwarp[warp] = incl;
        __syncthreads();
        u64 base = 0;
        for (int q = 0; q < warp; q++) base += dwarp[q];
        run[threadIdx.x] = base + incl - v + offs[(size_t)threadIdx.x * ck.G + blockIdx.x];
    }
    const int64_t lo = (int64_t)blockIdx.x * ck.per;
    const int64_t hi = lo + ck.per < ck.num ? lo + ck.per : ck.num;
    for (int64_t t0 = lo; t0 < hi; t0 += SORT_TILE) {
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) whist[w][threadIdx.x] = 0;
        __syncthreads();
        u64 k[SORT_ITEMS];
        u32 v[SORT_ITEMS];
        u32 rank[SORT_ITEMS];
        // warp w owns tile elements [w*256, (w+1)*256) as 8 consecutive 32-key slices
#pragma unroll
        for (int j = 0; j < SORT_ITEMS; j++) {
            const int64_t i = t0 + warp * (32 * SORT_ITEMS) + j * 32 + lane;
            int d = -1;
            if (i < hi) { k[j] = kin[i]; v[j] = vin[i]; d = (int)((k[j] >> shift) & 255); }
            unsigned peers = __match_any_sync(0xffffffffu, d);
            int leader = __ffs(peers) - 1;
            u32 base = 0;
            if (d >= 0 && lane == leader) {
                base = whist[warp][d];
                whist[warp][d] = base + __popc(peers);
            }
            base = __shfl_sync(0xffffffffu, base, leader);
            rank[j] = base + __popc(peers & ((1u << lane) - 1));
            __syncwarp();
        }
        __syncthreads();
        // digit d = threadIdx.x: exclusive prefix over warps (whist), exclusive prefix over digits of the tile
        // totals (tbase) = where each digit's run starts inside the tile once it is sorted by digit
        {
            const int d = threadIdx.x;
            u32 acc = 0;
#pragma unroll
            for (int w = 0; w < SORT_WARPS; w++) {
                u32 c = whist[w][d];
                whist[w][d] = acc;
                acc += c;
            }
            u32 incl = acc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                u32 t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) twarp[warp] = incl;
            __syncthreads();
            u32 base = 0;
            for (int q = 0; q < warp; q++) base += twarp[q];
            tbase[d] = base + incl - acc;
            __syncthreads();
            // stage the tile in shared memory in digit order ...
#pragma unroll
            for (int j = 0; j < SORT_ITEMS; j++) {
                const int64_t i = t0 + warp * (32 * SORT_ITEMS) + j * 32 + lane;
                if (i < hi) {
                    const int dd = (int)((k[j] >> shift) & 255);
                    const u32 pos = tbase[dd] + whist[warp][dd] + rank[j];
                    sk[pos] = k[j];
                    sv[pos] = v[j];
                }
            }
            __syncthreads();
            // ... and write it out position by position: keys of one digit are consecutive in the tile AND at
            // their destination, so a warp's store covers a few contiguous runs instead of 32 scattered keys
            // (the direct scatter ran at 0.85 TB/s, profiles/r01h_sort_launches_summary.txt)
            const int cnt = (int)(hi - t0 < SORT_TILE ? hi - t0 : SORT_TILE);
            for (int p = threadIdx.x; p < cnt; p += SORT_THREADS) {
                const u64 kk = sk[p];
                const int dd = (int)((kk >> shift) & 255);
                const u64 dst = run[dd] + (u32)(p - tbase[dd]);
                kout[dst] = kk;
                vout[dst] = sv[p];
            }
            __syncthreads();
            run[d] += acc;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SORT_THREADS)
k_sort_finish(const u64 *__restrict__ keys, const u32 *__restrict__ idx, int64_t num, double *__restrict__ keys_out,
              long long *__restrict__ idx_out)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    const double nanv = __longlong_as_double(0x7ff8000000000000ll);
    for (int64_t i = tid; i < num; i += nthr) {
        if (keys_out) keys_out[i] = unsort_key(keys[i], nanv);
        if (idx_out) idx_out[i] = (long long)idx[i];
    }
}

// ------------------------------------------------------------------ gathered inclusive scan
// out[i] = sum_{j<=i} (w ? w[idx[j]] : 1).  Same chunking: chunk sums, one-CTA scan, rescan.
__global__ void __launch_bounds__(SORT_THREADS)
k_scan_chunk_sums(const double *__restrict__ w, const long long *__restrict__ idx, Chunking ck,
                  double *__restrict__ csum)
{
    __shared__ double sh[SORT_WARPS];
    const int64_t lo = (int64_t)blockIdx.x * ck.per;
    const int64_t hi = lo + ck.per < ck.num ? lo + ck.per : ck.num;
    double acc = 0.;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) acc += w ? w[idx ? idx[i] : i] : 1.;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.;
        for (int q = 0; q < SORT_WARPS; q++) t += sh[q];
        csum[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024) k_scan_chunk_offsets(double *__restrict__ csum, int G)
{
    // serial exclusive scan of <= 1024 chunk sums (left-to-right like np.cumsum), staged through shared
    // memory so that the one scanning thread does not chain G global-memory round trips
    __shared__ double sh[SORT_MAXG + 1];
    for (int g = threadIdx.x; g < G; g += blockDim.x) sh[g] = csum[g];
    __syncthreads();
    if (threadIdx.x == 0) {
        double run = 0.;
        for (int g = 0; g < G; g++) { double t = sh[g]; sh[g] = run; run += t; }
        sh[G] = run;
    }
    __syncthreads();
    for (int g = threadIdx.x; g <= G; g += blockDim.x) csum[g] = sh[g];
}

__global__ void __launch_bounds__(SORT_THREADS)
k_scan_apply(const double *__restrict__ w, const long long *__restrict__ idx, Chunking ck,
             const double *__restrict__ csum, double *__restrict__ out)
{
    __shared__ double wtot[SORT_WARPS];
    __shared__ double carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = csum[blockIdx.x];
    __syncthreads();
    const int64_t lo = (int64_t)blockIdx.x * ck.per;
    const int64_t hi = lo + ck.per < ck.num ? lo + ck.per : ck.num;
    for (int64_t t0 = lo; t0 < hi; t0 += SORT_THREADS) {
        const int64_t i = t0 + threadIdx.x;
        double v = 0.;
        if (i < hi) v = w ? w[idx ? idx[i] : i] : 1.;
        double incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();
        double base = carry;
        for (int q = 0; q < warp; q++) base += wtot[q];
        if (i < hi) out[i] = base + incl;
        __syncthreads();
        if (threadIdx.x == SORT_THREADS - 1) carry = base + incl;
        __syncthreads();
    }
}

// ------------------------------------------------------------------ argmin |cdf/max - q|
// np.argmin returns the first minimiser; NaN handling follows numpy (first NaN wins).
struct ArgMin { double v; long long i; };
PXF_DEV bool am_better(double v, long long i, double bv, long long bi)
{
    const bool vn = v != v, bn = bv != bv;
    if (vn || bn) { if (vn && bn) return i < bi; return vn; }
    return v < bv || (v == bv && i < bi);
}

__global__ void __launch_bounds__(SORT_THREADS)
k_cdf_argmin(const double *__restrict__ cdf, int64_t num, const double *__restrict__ maxv_ptr, double q0, double q1,
             ArgMin *__restrict__ partial /*[grid][2]*/)
{
    const double mx = *maxv_ptr;
    double bv[2] = {__longlong_as_double(0x7ff0000000000000ll), __longlong_as_double(0x7ff0000000000000ll)};
    long long bi[2] = {0x7fffffffffffffffll, 0x7fffffffffffffffll};
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < num; i += nthr) {
        const double c = cdf[i] / mx;
        const double d0 = fabs(c - q0), d1 = fabs(c - q1);
        if (am_better(d0, i, bv[0], bi[0])) { bv[0] = d0; bi[0] = i; }
        if (am_better(d1, i, bv[1], bi[1])) { bv[1] = d1; bi[1] = i; }
    }
    __shared__ double shv[2][SORT_WARPS];
    __shared__ long long shi[2][SORT_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < 2; c++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_down_sync(0xffffffffu, bv[c], o);
            long long oi = __shfl_down_sync(0xffffffffu, bi[c], o);
            if (am_better(ov, oi, bv[c], bi[c])) { bv[c] = ov; bi[c] = oi; }
        }
        if (lane == 0) { shv[c][warp] = bv[c]; shi[c][warp] = bi[c]; }
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        const int c = threadIdx.x;
        double v = shv[c][0];
        long long i = shi[c][0];
        for (int w = 1; w < SORT_WARPS; w++)
            if (am_better(shv[c][w], shi[c][w], v, i)) { v = shv[c][w]; i = shi[c][w]; }
        partial[blockIdx.x * 2 + c].v = v;
        partial[blockIdx.x * 2 + c].i = i;
    }
}

// cdf.max(): numpy max propagates NaN
__global__ void __launch_bounds__(SORT_THREADS)
k_max_partial(const double *__restrict__ a, int64_t num, double *__restrict__ partial)
{
    double m = -__longlong_as_double(0x7ff0000000000000ll);
    bool nan = false;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < num; i += nthr) {
        double v = a[i];
        if (v != v) nan = true;
        else if (v > m) m = v;
    }
    if (nan) m = __longlong_as_double(0x7ff8000000000000ll);
    __shared__ double sh[SORT_WARPS];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_down_sync(0xffffffffu, m, o);
        if (ov != ov || (m == m && ov > m)) m = ov;
    }
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = sh[0];
        for (int w = 1; w < SORT_WARPS; w++) {
            double ov = sh[w];
            if (ov != ov || (r == r && ov > r)) r = ov;
        }
        partial[blockIdx.x] = r;
    }
}

__global__ void k_hpdw_final(const double *__restrict__ maxpart, int nmax, double *__restrict__ maxv)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double r = maxpart[0];
        for (int w = 1; w < nmax; w++) {
            double ov = maxpart[w];
            if (ov != ov || (r == r && ov > r)) r = ov;
        }
        *maxv = r;
    }
}

__global__ void k_hpdw_result(const ArgMin *__restrict__ partial, int nblk, const double *__restrict__ rsorted,
                              double *__restrict__ out /*[3]: hpd, r75, r25*/)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double v0 = partial[0].v, v1 = partial[1].v;
        long long i0 = partial[0].i, i1 = partial[1].i;
        for (int b = 1; b < nblk; b++) {
            if (am_better(partial[2 * b].v, partial[2 * b].i, v0, i0)) { v0 = partial[2 * b].v; i0 = partial[2 * b].i; }
            if (am_better(partial[2 * b + 1].v, partial[2 * b + 1].i, v1, i1)) { v1 = partial[2 * b + 1].v; i1 = partial[2 * b + 1].i; }
        }
        double r75 = rsorted[i0], r25 = rsorted[i1];
        out[0] = r75 - r25;
        out[1] = r75;
        out[2] = r25;
    }
}

static Chunking make_chunking(int64_t num)
{
    Chunking ck;
    int G = sm_count() * 4;
    if (G > SORT_MAXG) G = SORT_MAXG;
    if (G < 1) G = 1;
    int64_t per = (num + G - 1) / G;
    per = ((per + SORT_TILE - 1) / SORT_TILE) * SORT_TILE;
    if (per < SORT_TILE) per = SORT_TILE;
    ck.num = num; ck.per = per;
    ck.G = (int)((num + per - 1) / per);
    if (ck.G < 1) ck.G = 1;
    return ck;
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace pxf

using namespace pxf;

extern "C" {

size_t pxf_sort_scratch_bytes(int64_t num)
{
    size_t n = (size_t)(num > 0 ? num : 1);
    return 2 * align256(n * 8) + 2 * align256(n * 4) + align256((size_t)256 * SORT_MAXG * 4) +
           align256((size_t)256 * SORT_MAXG * 8) + align256(9 * 256 * 8) + 1024;
}

// digits: bit d set = sort on byte d of the 64-bit key (LSD order).  0 = automatic: one up-front histogram of all
// eight bytes is read back (host sync) and bytes that are constant over the array are skipped.  A caller that
// knows which bytes can differ (e.g. keys from a narrow bracket) passes the mask and nothing is read back.
int pxf_argsort_digits(const double *keys_in, int64_t num, double *keys_out, int64_t *idx_out,
                       void *scratch, int32_t digits, pxf_stream_t stream)
{
    if (num < 0 || !keys_in || !scratch || num > 0xffffffffll || digits < 0 || digits > 255) { set_error("pxf_argsort: bad argument"); return PXF_ERR_INVALID; }
    if (sm_count() <= 0) { set_error("no CUDA device available (libpxf has no CPU fallback)"); return PXF_ERR_CUDA; }
    if (num == 0) return PXF_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const size_t n = (size_t)num;
    char *p = static_cast<char *>(scratch);
    u64 *kA = (u64 *)p; p += align256(n * 8);
    u64 *kB = (u64 *)p; p += align256(n * 8);
    u32 *vA = (u32 *)p; p += align256(n * 4);
    u32 *vB = (u32 *)p; p += align256(n * 4);
    u32 *table = (u32 *)p; p += align256((size_t)256 * SORT_MAXG * 4);
    u64 *offs = (u64 *)p; p += align256((size_t)256 * SORT_MAXG * 8);
    u64 *ghist = (u64 *)p;                 // [8][256] digit histogram, then [256] row totals of the current pass
    u64 *dtot = ghist + 8 * 256;
    Chunking ck = make_chunking(num);
    PXF_CUDA(cudaMemsetAsync(ghist, 0, 8 * 256 * 8, s));
    k_sort_prepare<<<ck.G, SORT_THREADS, 0, s>>>(keys_in, kA, vA, ck, ghist);
    count_launch();
    if (digits == 0) {
        u64 hg[8 * 256];
        PXF_CUDA(cudaMemcpyAsync(hg, ghist, sizeof(hg), cudaMemcpyDeviceToHost, s));
        PXF_CUDA(cudaStreamSynchronize(s));
        for (int d = 0; d < 8; d++) {
            bool trivial = false;
            for (int b = 0; b < 256; b++)
                if (hg[d * 256 + b] == (u64)num) { trivial = true; break; }
            if (!trivial) digits |= 1 << d;
        }
    }
    u64 *ki = kA, *ko = kB;
    u32 *vi = vA, *vo = vB;
    for (int d = 0; d < 8; d++) {
        if (!(digits & (1 << d))) continue;
        k_sort_hist<<<ck.G, SORT_THREADS, 0, s>>>(ki, ck, 8 * d, table);
        k_sort_scan_rows<<<256, 1024, 0, s>>>(table, ck.G, offs, dtot);
        k_sort_scatter<<<ck.G, SORT_THREADS, 0, s>>>(ki, vi, ko, vo, ck, 8 * d, offs, dtot);
        count_launch(3);
        u64 *tk = ki; ki = ko; ko = tk;
        u32 *tv = vi; vi = vo; vo = tv;
    }
    k_sort_finish<<<grid_for(num, SORT_THREADS * 2, 8), SORT_THREADS, 0, s>>>(
        ki, vi, num, keys_out, reinterpret_cast<long long *>(idx_out));
    count_launch();
    return check_launch("pxf_argsort");
}

int pxf_argsort(const double *keys_in, int64_t num, double *keys_out, int64_t *idx_out,
                void *scratch, pxf_stream_t stream)
{
    return pxf_argsort_digits(keys_in, num, keys_out, idx_out, scratch, 0, stream);
}

size_t pxf_scan_scratch_bytes(int64_t num) { (void)num; return (size_t)(SORT_MAXG + 8) * 8; }

int pxf_cumsum_gather(const double *w, const int64_t *idx, int64_t num, double *out,
                      void *scratch, pxf_stream_t stream)
{
    if (num < 0 || !out || !scratch) { set_error("pxf_cumsum_gather: bad argument"); return PXF_ERR_INVALID; }
    if (sm_count() <= 0) { set_error("no CUDA device available (libpxf has no CPU fallback)"); return PXF_ERR_CUDA; }
    if (num == 0) return PXF_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    Chunking ck = make_chunking(num);
    double *csum = static_cast<double *>(scratch);
    const long long *ix = reinterpret_cast<const long long *>(idx);
    k_scan_chunk_sums<<<ck.G, SORT_THREADS, 0, s>>>(w, ix, ck, csum);
    k_scan_chunk_offsets<<<1, 1024, 0, s>>>(csum, ck.G);
    k_scan_apply<<<ck.G, SORT_THREADS, 0, s>>>(w, ix, ck, csum, out);
    count_launch(3);
    return check_launch("pxf_cumsum_gather");
}

// analyses.hpd weighted branch (analyses.py:88-94): r,cdf = rhocdf(...); r[argmin|cdf-.75|]-r[argmin|cdf-.25|]
int pxf_hpd_weighted_sorted(const double *x, const double *y, const double *w, int64_t num, double *hpd_host,
                     pxf_stream_t stream)
{
    if (num <= 0 || !x || !y || !w || !hpd_host) { set_error("pxf_hpd_weighted_sorted: bad argument"); return PXF_ERR_INVALID; }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    int rc;
    const size_t n = (size_t)num;
    const int RG = 1024;
    Scratch sc;
    size_t bytes = pxf_sums_scratch_bytes() + 64 * 8 + 3 * align256(n * 8) + pxf_sort_scratch_bytes(num) +
                   pxf_scan_scratch_bytes(num) + align256(RG * 2 * sizeof(ArgMin)) + align256(RG * 8) + 1024;
    if ((rc = sc.alloc(bytes, s))) return rc;
    char *p = static_cast<char *>(sc.p);
    void *sum_scr = p; p += pxf_sums_scratch_bytes();
    double *sums = (double *)p; p += 32 * 8;
    double *small = (double *)p; p += 32 * 8;       // [0]=max, [4..6]=result
    double *rho = (double *)p; p += align256(n * 8);   // later reused for the cdf
    double *rs = (double *)p; p += align256(n * 8);
    long long *idx = (long long *)p; p += align256(n * 8);
    void *sort_scr = p; p += pxf_sort_scratch_bytes(num);
    void *scan_scr = p; p += align256(pxf_scan_scratch_bytes(num));
    ArgMin *am = (ArgMin *)p; p += align256(RG * 2 * sizeof(ArgMin));
    double *maxpart = (double *)p;
    // centroid with weights, then radii about it (analyses.py:60-71 with cent=True)
    if ((rc = pxf_sums(PXF_SUMS_CENTROID, x, y, nullptr, nullptr, nullptr, w, num, 0., 0., sums, sum_scr, stream))) return rc;
    double h[4];
    PXF_CUDA(cudaMemcpyAsync(h, sums, sizeof(h), cudaMemcpyDeviceToHost, s));
    PXF_CUDA(cudaStreamSynchronize(s));
    double cx = h[1] / h[0], cy = h[2] / h[0];
    if ((rc = pxf_rho(x, y, num, cx, cy, rho, stream))) return rc;
    if ((rc = pxf_argsort(rho, num, rs, reinterpret_cast<int64_t *>(idx), sort_scr, stream))) return rc;
    double *cdf = rho;
    if ((rc = pxf_cumsum_gather(w, reinterpret_cast<int64_t *>(idx), num, cdf, scan_scr, stream))) return rc;
    int grid = grid_for(num, SORT_THREADS * 4, 4);
    if (grid > RG) grid = RG;
    k_max_partial<<<grid, SORT_THREADS, 0, s>>>(cdf, num, maxpart);
    k_hpdw_final<<<1, 32, 0, s>>>(maxpart, grid, small);
    k_cdf_argmin<<<grid, SORT_THREADS, 0, s>>>(cdf, num, small, .75, .25, am);
    k_hpdw_result<<<1, 32, 0, s>>>(am, grid, rs, small + 4);
    count_launch(4);
    if ((rc = check_launch("pxf_hpd_weighted_sorted"))) return rc;
    double r[3];
    PXF_CUDA(cudaMemcpyAsync(r, small + 4, sizeof(r), cudaMemcpyDeviceToHost, s));
    PXF_CUDA(cudaStreamSynchronize(s));
    *hpd_host = r[0];
    return PXF_OK;
}

}  // extern "C"
