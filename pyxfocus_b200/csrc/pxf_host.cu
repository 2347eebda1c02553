// Host-buffer entry point: the call an f2py user makes.  HOST arrays in, HOST arrays mutated
// in place, the device in between.  The bundle is cut into chunks that stream through a ring
// of device slots on three streams -- H2D of chunk c+1, the fused per-ray program on chunk c
// and D2H of chunk c-1 overlap (PCIe is full duplex) -- so the transfer cost is
// max(H2D, D2H), not their sum, and device memory is bounded by the ring, not the bundle.
// Only the rows the program reads are uploaded and only the rows it writes are downloaded
// (the liveness masks of build_program).  When the caller also wants the HPD, final x,y stay
// resident in full-length device rows and the radix-select HPD runs on them at the end.
#include "pxf_program.h"

namespace pxf {

#define HOST_SLOTS 3
#define HOST_CHUNK (int64_t(1) << 22)   // rays per chunk: 32 MiB per row, ~2.5 ms of PCIe per row

struct HostRing {
    cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
    cudaEvent_t in_done[HOST_SLOTS] = {}, run_done[HOST_SLOTS] = {}, out_done[HOST_SLOTS] = {};
    double *slot_rows[HOST_SLOTS][10] = {};
    uint8_t *slot_alive[HOST_SLOTS] = {};
    double *full_x = nullptr, *full_y = nullptr;
    uint8_t *full_alive = nullptr;
    void *misc = nullptr;

    ~HostRing()
    {
        for (int k = 0; k < HOST_SLOTS; k++) {
            for (int r = 0; r < 10; r++)
                if (slot_rows[k][r]) cudaFree(slot_rows[k][r]);
            if (slot_alive[k]) cudaFree(slot_alive[k]);
            if (in_done[k]) cudaEventDestroy(in_done[k]);
            if (run_done[k]) cudaEventDestroy(run_done[k]);
            if (out_done[k]) cudaEventDestroy(out_done[k]);
        }
        if (full_x) cudaFree(full_x);
        if (full_y) cudaFree(full_y);
        if (full_alive) cudaFree(full_alive);
        if (misc) cudaFree(misc);
        if (s_in) cudaStreamDestroy(s_in);
        if (s_run) cudaStreamDestroy(s_run);
        if (s_out) cudaStreamDestroy(s_out);
    }
};

}  // namespace pxf

using namespace pxf;

extern "C" {

int pxf_hpd_unweighted_dev(const double *x, const double *y, int64_t num, double *out_dev, void *workspace,
                           pxf_stream_t stream);
size_t pxf_hpd_workspace_bytes(void);

int pxf_host_trace_program(double *const rows_host[10], int64_t num, const pxf_op *ops, int32_t nops,
                           int32_t write_back, double *hpd_host, uint8_t *alive_host,
                           int64_t *alive_count_host, double *x_dev_keep, double *y_dev_keep)
{
    if (!rows_host || num < 0) { set_error("pxf_host_trace_program: bad argument"); return PXF_ERR_INVALID; }
    FusedProgram fp;
    int rc = build_program(fp, ops, nops);
    if (rc) return rc;
    if (sm_count() <= 0) { set_error("no CUDA device available (libpxf has no CPU fallback)"); return PXF_ERR_CUDA; }
    const unsigned LM = fp.load_mask, SM = fp.store_mask, USED = LM | SM;
    for (int r = 0; r < 10; r++)
        if ((USED & (1u << r)) && !rows_host[r]) {
            set_error("pxf_host_trace_program: row %d is used by the program but NULL", r);
            return PXF_ERR_INVALID;
        }
    const bool want_alive = fp.has_vignette != 0;
    const bool want_hpd = hpd_host != nullptr;
    if (want_hpd && !rows_host[1]) { set_error("pxf_host_trace_program: hpd needs the x,y rows"); return PXF_ERR_INVALID; }
    if (num == 0) {
        if (alive_count_host) *alive_count_host = 0;
        if (hpd_host) *hpd_host = __builtin_nan("");
        return PXF_OK;
    }

    HostRing R;
    PXF_CUDA(cudaStreamCreateWithFlags(&R.s_in, cudaStreamNonBlocking));
    PXF_CUDA(cudaStreamCreateWithFlags(&R.s_run, cudaStreamNonBlocking));
    PXF_CUDA(cudaStreamCreateWithFlags(&R.s_out, cudaStreamNonBlocking));
    const int64_t chunk = num < HOST_CHUNK ? ((num + 1) & ~int64_t(1)) : HOST_CHUNK;
    const int64_t nchunks = (num + chunk - 1) / chunk;
    const int nslots = nchunks < HOST_SLOTS ? (int)nchunks : HOST_SLOTS;
    // x,y of the final bundle stay resident when the HPD is wanted (rows 1,2 then live in the
    // full-length arrays and the slots alias into them)
    const bool keep_xy = x_dev_keep && y_dev_keep;      // caller-owned device rows of length num
    if (keep_xy && ((reinterpret_cast<uintptr_t>(x_dev_keep) | reinterpret_cast<uintptr_t>(y_dev_keep)) & 15)) {
        set_error("pxf_host_trace_program: x_dev_keep/y_dev_keep must be 16-byte aligned");
        return PXF_ERR_INVALID;
    }
    const bool xy_full = want_hpd || keep_xy;
    double *fx = nullptr, *fy = nullptr;
    if (xy_full) {
        if (keep_xy) { fx = x_dev_keep; fy = y_dev_keep; }
        else {
            PXF_CUDA(cudaMalloc(&R.full_x, (size_t)(nchunks * chunk) * 8));
            PXF_CUDA(cudaMalloc(&R.full_y, (size_t)(nchunks * chunk) * 8));
            fx = R.full_x; fy = R.full_y;
        }
        if (want_alive) PXF_CUDA(cudaMalloc(&R.full_alive, (size_t)(nchunks * chunk)));
    }
    for (int k = 0; k < nslots; k++) {
        PXF_CUDA(cudaEventCreateWithFlags(&R.in_done[k], cudaEventDisableTiming));
        PXF_CUDA(cudaEventCreateWithFlags(&R.run_done[k], cudaEventDisableTiming));
        PXF_CUDA(cudaEventCreateWithFlags(&R.out_done[k], cudaEventDisableTiming));
        for (int r = 0; r < 10; r++) {
            if (!(USED & (1u << r))) continue;
            if (xy_full && (r == 1 || r == 2)) continue;
            PXF_CUDA(cudaMalloc(&R.slot_rows[k][r], (size_t)chunk * 8));
        }
        if (want_alive && !xy_full) PXF_CUDA(cudaMalloc(&R.slot_alive[k], (size_t)chunk));
    }

    int64_t alive_total = 0;
    // per-chunk survivor counts are reduced on the host from the flags when the caller asks
    // for them; the flags themselves are tiny next to the rows (1 B/ray)
    for (int64_t c = 0; c < nchunks; c++) {
        const int k = (int)(c % nslots);
        const int64_t lo = c * chunk;
        const int64_t n = (lo + chunk <= num) ? chunk : (num - lo);
        double *rows[10];
        for (int r = 0; r < 10; r++) rows[r] = R.slot_rows[k][r];
        if (xy_full) { rows[1] = fx + lo; rows[2] = fy + lo; }
        uint8_t *alive = want_alive ? (xy_full ? R.full_alive + lo : R.slot_alive[k]) : nullptr;
        // the slot is free once its previous D2H finished
        if (c >= nslots) PXF_CUDA(cudaStreamWaitEvent(R.s_in, R.out_done[k], 0));
        for (int r = 0; r < 10; r++)
            if (LM & (1u << r))
                PXF_CUDA(cudaMemcpyAsync(rows[r], rows_host[r] + lo, (size_t)n * 8, cudaMemcpyHostToDevice, R.s_in));
        PXF_CUDA(cudaEventRecord(R.in_done[k], R.s_in));
        PXF_CUDA(cudaStreamWaitEvent(R.s_run, R.in_done[k], 0));
        if (c >= nslots) PXF_CUDA(cudaStreamWaitEvent(R.s_run, R.out_done[k], 0));
        if ((rc = launch_program(rows, n, fp, alive, R.s_run))) return rc;
        PXF_CUDA(cudaEventRecord(R.run_done[k], R.s_run));
        PXF_CUDA(cudaStreamWaitEvent(R.s_out, R.run_done[k], 0));
        if (write_back)
            for (int r = 0; r < 10; r++)
                if (SM & (1u << r))
                    PXF_CUDA(cudaMemcpyAsync(rows_host[r] + lo, rows[r], (size_t)n * 8, cudaMemcpyDeviceToHost, R.s_out));
        if (want_alive && alive_host)
            PXF_CUDA(cudaMemcpyAsync(alive_host + lo, alive, (size_t)n, cudaMemcpyDeviceToHost, R.s_out));
        PXF_CUDA(cudaEventRecord(R.out_done[k], R.s_out));
    }
    PXF_CUDA(cudaStreamSynchronize(R.s_run));

    if (want_hpd) {
        const double *hx = fx, *hy = fy;
        int64_t hn = num;
        double *cx = nullptr, *cy = nullptr;
        if (want_alive) {
            // HPD over the surviving rays only: compact x,y by the alive flags first
            size_t sb = pxf_compact_scratch_bytes(num);
            void *scr = nullptr;
            PXF_CUDA(cudaMalloc(&scr, sb));
            int64_t cnt = 0;
            rc = pxf_compact_count(R.full_alive, num, scr, &cnt, reinterpret_cast<pxf_stream_t>(R.s_run));
            if (!rc && cnt > 0) {
                if (cudaMalloc(&cx, (size_t)cnt * 8) != cudaSuccess || cudaMalloc(&cy, (size_t)cnt * 8) != cudaSuccess) {
                    cudaFree(scr); if (cx) cudaFree(cx);
                    set_error("pxf_host_trace_program: out of device memory"); return PXF_ERR_NOMEM;
                }
                const double *in2[2] = {fx, fy};
                double *out2[2] = {cx, cy};
                rc = pxf_compact_scatter(in2, out2, 2, R.full_alive, num, scr, reinterpret_cast<pxf_stream_t>(R.s_run));
                cudaStreamSynchronize(R.s_run);
            }
            cudaFree(scr);
            if (rc) { if (cx) cudaFree(cx); if (cy) cudaFree(cy); return rc; }
            hx = cx; hy = cy; hn = cnt;
            alive_total = cnt;
        }
        if (hn > 0) {
            if (cudaMalloc(&R.misc, pxf_hpd_workspace_bytes() + 64) != cudaSuccess) {
                if (cx) cudaFree(cx); if (cy) cudaFree(cy);
                set_error("pxf_host_trace_program: out of device memory"); return PXF_ERR_NOMEM;
            }
            double *out = reinterpret_cast<double *>((char *)R.misc + pxf_hpd_workspace_bytes());
            rc = pxf_hpd_unweighted_dev(hx, hy, hn, out, R.misc, reinterpret_cast<pxf_stream_t>(R.s_run));
            double h[3] = {0, 0, 0};
            if (!rc && cudaMemcpyAsync(h, out, sizeof(h), cudaMemcpyDeviceToHost, R.s_run) != cudaSuccess) rc = PXF_ERR_CUDA;
            cudaStreamSynchronize(R.s_run);
            *hpd_host = h[0];
        } else {
            *hpd_host = __builtin_nan("");
        }
        if (cx) cudaFree(cx);
        if (cy) cudaFree(cy);
        if (rc) return rc;
    }
    PXF_CUDA(cudaStreamSynchronize(R.s_out));
    if (alive_count_host) {
        if (!want_alive) *alive_count_host = num;
        else if (want_hpd) *alive_count_host = alive_total;
        else if (alive_host) {
            int64_t t = 0;
            for (int64_t i = 0; i < num; i++) t += alive_host[i] != 0;
            *alive_count_host = t;
        } else {
            *alive_count_host = -1;   // flags were not requested anywhere
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("pxf_host_trace_program: %s", cudaGetErrorString(e)); return PXF_ERR_CUDA; }
    return PXF_OK;
}

}  // extern "C"
