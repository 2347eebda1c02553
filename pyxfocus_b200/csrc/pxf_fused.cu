// Fused per-ray program: one kernel loads a ray once, runs a whole op list in registers and
// stores once -- no HBM round trip per optical element.  The per-op arithmetic is the same
// device code as the per-routine kernels (pxf_ray.cuh), so the result is bit-identical to
// issuing the routines one by one.  Every ray runs the same program, so the opcode switch is
// warp-uniform; only Newton trip counts and vignetting diverge.
#include <stdlib.h>
#include "pxf_program.h"

namespace pxf {

PXF_DEV double ray_row(const Ray &r, int row)
{
    switch (row) {
        case 0: return r.opd; case 1: return r.x; case 2: return r.y; case 3: return r.z;
        case 4: return r.l; case 5: return r.m; case 6: return r.n; case 7: return r.ux;
        case 8: return r.uy; default: return r.uz;
    }
}

// returns false when the ray is vignetted at this op
// ZERN: whether this instantiation carries the Zernike evaluation at all -- it needs ~3x the registers of every
// other op, so programs without a PXF_OP_ZERNSURF op run a kernel compiled without it
// one atomic per warp: the largest grating count seen by the lanes that are here together
PXF_DEV void count_max_update(int *dst, int k)
{
    const unsigned act = __activemask();
    const int m = __reduce_max_sync(act, k);
    if ((threadIdx.x & 31) == (__ffs(act) - 1)) atomicMax(dst, m);
}

// i: the ray's index in the bundle (side arrays); aux_*: FusedProgram::aux_* (null unless the program uses them)
// AUX: whether this instantiation carries the side-array ops (the grating fan: asin + a data-dependent loop) -- like
// ZERN they get their own kernel so that every other program keeps its register budget
template <bool ZERN = false, bool AUX = false>
PXF_DEV bool run_op(Ray &r, const FusedOp &op, const ZernP *zt = nullptr, int64_t i = 0,
                    const double *aux_wave = nullptr, int *aux_count = nullptr, int *aux_count_max = nullptr)
{
    switch (op.code) {
        case PXF_OP_KICKN: op_kickn(r, *reinterpret_cast<const KickNP *>(op.q)); break;
        case PXF_OP_VIGNETTE_RHOGT: {
            double rho = sqrt(sq(r.x) + sq(r.y));
            return rho > op.q[0];
        }
        case PXF_OP_GRATFAN:
            if constexpr (AUX) {
                const GratFanP &p = *reinterpret_cast<const GratFanP *>(op.q);
                const int k = op_gratfan(r, p, p.wave_array ? aux_wave[i] : p.g.wave);
                if (aux_count) aux_count[i] = k;
                if (aux_count_max) count_max_update(aux_count_max, k < 0 ? p.cap + 1 : k);
            }
            break;
        case PXF_OP_ROTX_REMAINING:
            if constexpr (AUX) {
                const int k = aux_count[i];
                if (k >= 0) op_rotx_repeat(r, *reinterpret_cast<const TransformP *>(op.q), op.row - k);
            }
            break;
        case PXF_OP_TRANSFORM: op_transform(r, *reinterpret_cast<const TransformP *>(op.q)); break;
        case PXF_OP_ITRANSFORM: op_itransform(r, *reinterpret_cast<const TransformP *>(op.q)); break;
        case PXF_OP_REFLECT: op_reflect(r); break;
        case PXF_OP_REFRACT: op_refract(r, *reinterpret_cast<const RefractP *>(op.q)); break;
        case PXF_OP_RADGRAT: {
            const RadgratP &p = *reinterpret_cast<const RadgratP *>(op.q);
            op_radgrat(r, p, p.wave, false);
            break;
        }
        case PXF_OP_FLAT: op_flat(r, false, 0.); break;
        case PXF_OP_FLATOPD: op_flat(r, true, op.q[0]); break;
        case PXF_OP_CONIC:
        case PXF_OP_CONICOPD: op_conic(r, *reinterpret_cast<const ConicP *>(op.q)); break;
        // (parameter blocks of the Newton surfaces are copied to registers once: read through the
        // op-indexed constant bank they would be re-fetched by an LDC on every use in every pass)
        case PXF_OP_WOLTERPRIMARY:
        case PXF_OP_WOLTERPRIMARYOPD: { const WolterP p = *reinterpret_cast<const WolterP *>(op.q); op_wolterprimary(r, p); break; }
        case PXF_OP_WOLTERSECONDARY: { const WolterP p = *reinterpret_cast<const WolterP *>(op.q); op_woltersecondary(r, p); break; }
        case PXF_OP_WOLTERSINE: op_woltersine(r, *reinterpret_cast<const WolterSineP *>(op.q)); break;
        case PXF_OP_WSPRIMARY: op_wsprimary(r, *reinterpret_cast<const WSP *>(op.q)); break;
        case PXF_OP_WSSECONDARY: op_wssecondary(r, *reinterpret_cast<const WSP *>(op.q)); break;
        case PXF_OP_SPOCONE: op_spocone(r, *reinterpret_cast<const SpoP *>(op.q)); break;
        case PXF_OP_VIGNETTE_MAG: {
            double mag = sq(r.l) + sq(r.m) + sq(r.n);
            return mag > .1;
        }
        case PXF_OP_VIGNETTE_BOX: {
            double v = ray_row(r, op.row);
            return (v > op.q[0]) && (v < op.q[1]);
        }
        case PXF_OP_VIGNETTE_ABS: {
            double v = ray_row(r, op.row);
            if (op.q[2] != 0.) v = v - op.q[1];        // |row - centre| < hi (the centre is optional: v - 0 is v, but skip it)
            return fabs(v) < op.q[0];
        }
        case PXF_OP_ZERNSURF:
            // zt: the table in shared memory (k_program); same device code as the per-routine kernel (nmax <= 7)
            if constexpr (ZERN) {
                if (zt) op_tracezern<7>(r, zt->rad, zt->nr, zt->tol, zt->nmax, op.q[1] != 0., reinterpret_cast<const double *>(zt->e));
            }
            break;
        case PXF_OP_KICK: {
            r.l = r.l + op.q[0];
            r.m = r.m + op.q[1];
            r.n = op.q[2] * sqrt(1. - sq(r.l) - sq(r.m));
            break;
        }
        default: break;
    }
    return true;
}

template <bool ZERN = false, bool AUX = false>
PXF_DEV bool run_program(Ray &r, const FusedProgram &prog, const ZernP *zt = nullptr, int64_t i = 0)
{
    for (int k = 0; k < prog.nops; k++)
        if (!run_op<ZERN, AUX>(r, prog.ops[k], zt, i, prog.aux_wave, prog.aux_count, prog.aux_count_max)) return false;
    return true;
}

template <bool VEC2, int MINB = 1, bool ZERN = false, bool AUX = false>
__global__ void __launch_bounds__(PXF_BLOCK, MINB)
k_program(const RowPtrs P, const RowPtrs Q, const int64_t num, uint8_t *__restrict__ alive,
          double *__restrict__ partials, const __grid_constant__ FusedProgram prog)
{
    double cnt = 0., sx = 0., sy = 0.;
    // the Zernike table of a PXF_OP_ZERNSURF op, staged once per CTA
    __shared__ __align__(16) unsigned char zraw[ZERN ? sizeof(ZernP) : 16];
    const ZernP *zt = nullptr;
    if (ZERN && prog.zern) {
        const double *src = reinterpret_cast<const double *>(prog.zern);
        double *dst = reinterpret_cast<double *>(zraw);
        for (int t = threadIdx.x; t < (int)(sizeof(ZernP) / 8); t += blockDim.x) dst[t] = src[t];
        __syncthreads();
        zt = reinterpret_cast<const ZernP *>(zraw);
    }
    // P: rows read, Q: rows written (Q == P for the in-place f2py semantics)
    const unsigned LM = prog.load_mask, SM = prog.store_mask;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    if (VEC2) {
        const int64_t npair = num >> 1;
        for (int64_t q = tid; q < npair; q += nthr) {
            const int64_t i = q << 1;
            Ray a, b;
            fload2(a, b, P, LM, i);
            const bool ka = run_program<ZERN, AUX>(a, prog, zt, i);
            const bool kb = run_program<ZERN, AUX>(b, prog, zt, i + 1);
            fstore2(a, b, Q, SM, i);
            if (alive) { alive[i] = ka ? 1 : 0; alive[i + 1] = kb ? 1 : 0; }
            if (ka) { cnt += 1.; sx += a.x; sy += a.y; }
            if (kb) { cnt += 1.; sx += b.x; sy += b.y; }
        }
        if ((num & 1) && tid == 0) {
            const int64_t i = num - 1;
            Ray a;
            fload1(a, P, LM, i);
            const bool ka = run_program<ZERN, AUX>(a, prog, zt, i);
            fstore1(a, Q, SM, i);
            if (alive) alive[i] = ka ? 1 : 0;
            if (ka) { cnt += 1.; sx += a.x; sy += a.y; }
        }
    } else {
        for (int64_t i = tid; i < num; i += nthr) {
            Ray a;
            fload1(a, P, LM, i);
            const bool ka = run_program<ZERN, AUX>(a, prog, zt, i);
            fstore1(a, Q, SM, i);
            if (alive) alive[i] = ka ? 1 : 0;
            if (ka) { cnt += 1.; sx += a.x; sy += a.y; }
        }
    }
    if (partials) centroid_block_reduce(cnt, sx, sy, partials);
}

// ---------------------------------------------------------------- segmented execution
// Nested assemblies (BASELINE config 5; examples/axro/axialHeights.py:215-322, SMARTX.py:163-259): the bundle
// is a concatenation of segments (one mirror shell each) and every segment runs the SAME opcode sequence with
// its own folded scalars.  The reference loops over shells in Python; one launch per shell is launch bound at
// 260 shells (measured 12 ms of launches for 1e7 rays).  Here one persistent grid walks the bundle in tiles of
// SEG_TILE consecutive rays; a CTA stages the op table of the segment its tile lies in into shared memory
// (re-staged only when the segment changes; a tile that straddles segments is processed piecewise).
#define SEG_TILE (PXF_BLOCK * 8)
struct SegHeader { unsigned load_mask, store_mask; int nops, nseg; int chain_id, jit_pack_bytes, pad[2]; };

template <int MINB>
__global__ void __launch_bounds__(PXF_BLOCK, MINB)
k_program_seg(const RowPtrs P, const RowPtrs Q, const int64_t num, uint8_t *__restrict__ alive,
              const long long *__restrict__ seg_start, const FusedOp *__restrict__ ops, const int nops, const int nseg,
              const unsigned LM, const unsigned SM)
{
    extern __shared__ __align__(16) unsigned char seg_smem[];
    FusedOp *sops = reinterpret_cast<FusedOp *>(seg_smem);
    int staged = -1;
    const int words = nops * (int)(sizeof(FusedOp) / 8);
    for (int64_t t0 = (int64_t)blockIdx.x * SEG_TILE; t0 < num; t0 += (int64_t)gridDim.x * SEG_TILE) {
        const int64_t t1 = t0 + SEG_TILE < num ? t0 + SEG_TILE : num;
        int lo = 0, hi = nseg - 1;                     // last segment starting at or before t0
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (seg_start[mid] <= t0) lo = mid; else hi = mid - 1;
        }
        int seg = lo;
        int64_t pos = t0;
        while (pos < t1 && seg < nseg) {
            int64_t send = seg_start[seg + 1];
            if (send <= pos) { seg++; continue; }      // empty segment
            if (send > t1) send = t1;
            if (seg != staged) {                       // block-uniform
                __syncthreads();
                const double *src = reinterpret_cast<const double *>(ops + (int64_t)seg * nops);
                double *dst = reinterpret_cast<double *>(sops);
                for (int t = threadIdx.x; t < words; t += blockDim.x) dst[t] = src[t];
                __syncthreads();
                staged = seg;
            }
            for (int64_t i = pos + threadIdx.x; i < send; i += blockDim.x) {
                Ray a;
                fload1(a, P, LM, i);
                bool keep = true;
                for (int k = 0; k < nops; k++)
                    if (!run_op(a, sops[k])) { keep = false; break; }
                fstore1(a, Q, SM, i);
                if (alive) alive[i] = keep ? 1 : 0;
            }
            pos = send;
            seg++;
        }
    }
}

// rows read (use) / possibly written (st) / unconditionally overwritten (kill) by each op
static void op_masks(int code, int row, unsigned &use, unsigned &st, unsigned &kill)
{
    switch (code) {
        case PXF_OP_TRANSFORM: case PXF_OP_ITRANSFORM: use = R_NINE; st = R_NINE; kill = R_NINE; break;
        case PXF_OP_REFLECT: use = R_DIR | R_NRM; st = R_DIR; kill = R_DIR; break;
        case PXF_OP_REFRACT: use = R_DIR | R_NRM; st = R_DIR | R_NRM; kill = 0; break;
        case PXF_OP_RADGRAT: use = R_X | R_Y | R_DIR; st = R_DIR; kill = R_DIR; break;
        case PXF_OP_FLAT: use = R_POS | R_DIR; st = R_POS | R_NRM; kill = R_POS | R_NRM; break;
        case PXF_OP_FLATOPD: use = R_POS | R_DIR | R_OPD; st = R_POS | R_NRM | R_OPD; kill = R_POS | R_NRM; break;
        case PXF_OP_CONIC: case PXF_OP_SPOCONE: use = R_NINE; st = R_NINE; kill = 0; break;
        case PXF_OP_CONICOPD: use = R_ALL; st = R_ALL; kill = 0; break;
        case PXF_OP_WOLTERPRIMARY: case PXF_OP_WOLTERSECONDARY: case PXF_OP_WOLTERSINE:
            use = R_POS | R_DIR; st = R_POS | R_NRM; kill = R_POS | R_NRM; break;
        case PXF_OP_WOLTERPRIMARYOPD:
            use = R_POS | R_DIR | R_OPD; st = R_POS | R_NRM | R_OPD; kill = R_POS | R_NRM; break;
        case PXF_OP_WSPRIMARY: case PXF_OP_WSSECONDARY: use = R_NINE; st = R_POS | R_NRM; kill = R_POS; break;
        case PXF_OP_VIGNETTE_MAG: use = R_DIR; st = 0; kill = 0; break;
        case PXF_OP_VIGNETTE_BOX: case PXF_OP_VIGNETTE_ABS: use = 1u << row; st = 0; kill = 0; break;
        case PXF_OP_KICK: use = R_L | R_M; st = R_DIR; kill = R_N; break;
        case PXF_OP_KICKN: use = R_DIR; st = R_DIR; kill = 0; break;
        case PXF_OP_VIGNETTE_RHOGT: use = R_X | R_Y; st = 0; kill = 0; break;
        case PXF_OP_GRATFAN: case PXF_OP_ROTX_REMAINING: use = R_NINE; st = R_NINE; kill = 0; break;
        case PXF_OP_ZERNSURF: use = R_POS | R_DIR | (row ? R_OPD : 0u); st = R_POS | R_NRM | (row ? R_OPD : 0u); kill = R_POS | R_NRM; break;
        default: use = 0; st = 0; kill = 0; break;
    }
}

int build_program(FusedProgram &fp, const pxf_op *ops, int nops, const pxf_program_aux *aux)
{
    if (!ops || nops < 1 || nops > PXF_MAX_OPS) { set_error("program: need 1..%d ops", PXF_MAX_OPS); return PXF_ERR_INVALID; }
    memset(&fp, 0, sizeof(fp));
    fp.nops = nops;
    if (aux) { fp.aux_wave = aux->wave; fp.aux_count = aux->count; fp.aux_count_max = aux->count_max; }
    unsigned store = 0;
    for (int k = 0; k < nops; k++) {
        const pxf_op &o = ops[k];
        FusedOp &f = fp.ops[k];
        f.code = o.code;
        f.row = 0;
        const double *p = o.p;
        switch (o.code) {
            case PXF_OP_TRANSFORM: { TransformP t = make_transform(p[0], p[1], p[2], p[3], p[4], p[5]); memcpy(f.q, &t, sizeof(t)); break; }
            case PXF_OP_ITRANSFORM: { TransformP t = make_itransform(p[0], p[1], p[2], p[3], p[4], p[5]); memcpy(f.q, &t, sizeof(t)); break; }
            case PXF_OP_REFLECT: case PXF_OP_FLAT: case PXF_OP_VIGNETTE_MAG: break;
            case PXF_OP_REFRACT: { RefractP t; t.ratio = p[0] / p[1]; memcpy(f.q, &t, sizeof(t)); break; }
            case PXF_OP_RADGRAT: { RadgratP t = make_radgrat(p[0], p[1], p[2]); memcpy(f.q, &t, sizeof(t)); break; }
            case PXF_OP_FLATOPD: f.q[0] = p[0]; break;
            case PXF_OP_CONIC: { ConicP t = make_conic(p[0], p[1], false, 0.); memcpy(f.q, &t, sizeof(t)); break; }
            case PXF_OP_CONICOPD: { ConicP t = make_conic(p[0], p[1], true, p[2]); memcpy(f.q, &t, sizeof(t)); break; }
            case PXF_OP_WOLTERPRIMARY: { WolterP t = make_wolter(p[0], p[1], p[2], false, 0.); memcpy(f.q, &t, sizeof(t)); break; }
            case PXF_OP_WOLTERPRIMARYOPD: { WolterP t = make_wolter(p[0], p[1], p[2], true, p[3]); memcpy(f.q, &t, sizeof(t)); break; }
            case PXF_OP_WOLTERSECONDARY: { WolterP t = make_wolter(p[0], p[1], p[2], false, 0.); memcpy(f.q, &t, sizeof(t)); break; }
            case PXF_OP_WOLTERSINE: { WolterSineP t = make_woltersine(p[0], p[1], p[2], p[3]); memcpy(f.q, &t, sizeof(t)); break; }
            case PXF_OP_WSPRIMARY: case PXF_OP_WSSECONDARY: { WSP t = make_ws(p[0], p[1], p[2]); memcpy(f.q, &t, sizeof(t)); break; }
            case PXF_OP_SPOCONE: { SpoP t = make_spo(p[0], p[1]); memcpy(f.q, &t, sizeof(t)); break; }
            case PXF_OP_VIGNETTE_BOX:
                f.row = (int)p[0];
                if (f.row < 0 || f.row > 9) { set_error("program: bad row in VIGNETTE_BOX"); return PXF_ERR_INVALID; }
                f.q[0] = p[1]; f.q[1] = p[2];
                break;
            case PXF_OP_VIGNETTE_ABS:
                f.row = (int)p[0];
                if (f.row < 0 || f.row > 9) { set_error("program: bad row in VIGNETTE_ABS"); return PXF_ERR_INVALID; }
                f.q[0] = p[1]; f.q[1] = p[2]; f.q[2] = (p[2] != 0.) ? 1. : 0.;
                break;
            case PXF_OP_KICK: f.q[0] = p[0]; f.q[1] = p[1]; f.q[2] = p[2]; break;
            case PXF_OP_KICKN: { KickNP t; t.dl = p[0]; t.dm = p[1]; t.dl2 = p[0] * p[0]; t.dm2 = p[1] * p[1]; memcpy(f.q, &t, sizeof(t)); break; }
            case PXF_OP_VIGNETTE_RHOGT: f.q[0] = p[0]; break;
            case PXF_OP_GRATFAN: {
                // p: ang, hubdist, l, dpermm, order, wave (NaN: per-ray wavelengths).  The script's
                // tran.transform(rays,0,0,0,ang,0,0) reaches the Fortran negated (transformations.py:29).
                GratFanP t;
                memset(&t, 0, sizeof(t));
                t.rot = make_transform(0., 0., 0., -p[0], 0., 0.);
                t.wave_array = (p[5] != p[5]) ? 1 : 0;
                t.g = make_radgrat(t.wave_array ? 0. : p[5], p[3], p[4]);
                t.hub = p[1]; t.hub_l = p[2] + p[1]; t.thresh = .001;
                t.cap = (aux && aux->cap > 0) ? aux->cap : 4096;
                if (t.wave_array && !(aux && aux->wave)) { set_error("program: PXF_OP_GRATFAN with wave = NaN needs aux.wave"); return PXF_ERR_INVALID; }
                memcpy(f.q, &t, sizeof(t));
                fp.uses_aux = 1;
                break;
            }
            case PXF_OP_ROTX_REMAINING: {
                if (!(aux && aux->count)) { set_error("program: PXF_OP_ROTX_REMAINING needs aux.count"); return PXF_ERR_INVALID; }
                if (!(p[1] >= 0. && p[1] <= 1.e6)) { set_error("program: PXF_OP_ROTX_REMAINING: bad total"); return PXF_ERR_INVALID; }
                TransformP t = make_transform(0., 0., 0., -p[0], 0., 0.);
                memcpy(f.q, &t, sizeof(t));
                f.row = (int)p[1];                     // total rotations of the loop
                f.q[sizeof(TransformP) / 8] = (double)f.row;     // (RotxP of the specialised chains)
                fp.uses_aux = 1;
                break;
            }
            case PXF_OP_ZERNSURF: {
                const ZernP *tab;
                memcpy(&tab, &p[0], sizeof(tab));
                if (!tab || fp.zern) { set_error("program: PXF_OP_ZERNSURF needs a device table, and only one per program"); return PXF_ERR_INVALID; }
                if (!(p[2] >= 0. && p[2] <= 7.)) { set_error("program: PXF_OP_ZERNSURF supports radial orders <= 7 (use pxf_tracezern)"); return PXF_ERR_UNSUPPORTED; }
                fp.zern = tab;
                f.q[0] = 0.; f.q[1] = p[1] != 0. ? 1. : 0.;
                f.row = p[1] != 0. ? 1 : 0;            // opd flag: decides whether row 0 is touched
                break;
            }
            default: set_error("program: unknown opcode %d", o.code); return PXF_ERR_INVALID;
        }
        if (o.code == PXF_OP_VIGNETTE_MAG || o.code == PXF_OP_VIGNETTE_BOX || o.code == PXF_OP_VIGNETTE_ABS ||
            o.code == PXF_OP_VIGNETTE_RHOGT)
            fp.has_vignette = 1;
        unsigned use, st, kill;
        op_masks(o.code, f.row, use, st, kill);
        store |= st;
        // forward constant tracking: whatever this op writes is no longer a known constant, except
        // the plane's own outputs z = 0, u = (0,0,1) (surfacesf.f95:17-23)
        fp.const_mask &= ~st;
        if (o.code == PXF_OP_FLAT || o.code == PXF_OP_FLATOPD) {
            fp.const_mask |= R_Z | R_NRM;
            fp.const_val[3] = 0.; fp.const_val[7] = 0.; fp.const_val[8] = 0.; fp.const_val[9] = 1.;
        }
    }
    if (fp.has_vignette) fp.const_mask = 0;
    // Backward liveness: a row is read from HBM only if some op consumes its incoming value
    // before an op overwrites it unconditionally (e.g. the normals entering a transform that is
    // followed by a surface are dead: never loaded, never rotated).  Every row any op may write
    // is live at the end (it is stored).
    unsigned live = store;
    for (int k = nops - 1; k >= 0; k--) {
        FusedOp &f = fp.ops[k];
        unsigned use, st, kill;
        op_masks(f.code, f.row, use, st, kill);
        if (f.code == PXF_OP_TRANSFORM || f.code == PXF_OP_ITRANSFORM) {
            // position / direction / normal triplets transform independently
            unsigned groups = 0, in = live & ~R_NINE;
            if (live & R_POS) { groups |= 1; in |= R_POS; }
            if (live & R_DIR) { groups |= 2; in |= R_DIR; }
            if (live & R_NRM) { groups |= 4; in |= R_NRM; }
            reinterpret_cast<TransformP *>(f.q)->groups = (int)groups;
            live = in;
        } else {
            const bool side_effect = (st == 0);   // vignette predicates
            unsigned in = live & ~kill;
            if ((st & live) || side_effect) in |= use;
            live = in;
        }
    }
    fp.load_mask = live;
    fp.store_mask = store;
    // A ray stopped by a vignette predicate is stored in the state it had AT the predicate (the contract for dead
    // rays: whatever the ops before the predicate made of them, nothing after).  Rows a later op would have
    // overwritten are dead for the survivors but not for the stopped rays, so they must be loaded as well.
    if (fp.has_vignette) fp.load_mask |= store;
    return PXF_OK;
}

int launch_program(double *const rays[10], int64_t num, const FusedProgram &fp, uint8_t *alive, cudaStream_t s,
                   double *const rays_out[10], double *partials, int *grid_out)
{
    if (num < 0 || !rays) { set_error("program: bad argument"); return PXF_ERR_INVALID; }
    if (sm_count() <= 0) { set_error("no CUDA device available (libpxf has no CPU fallback)"); return PXF_ERR_CUDA; }
    if (fp.has_vignette && !alive) { set_error("program with a VIGNETTE op needs an alive array"); return PXF_ERR_INVALID; }
    if (num == 0) return PXF_OK;
    RowPtrs P, Q;
    bool aligned = true;
    for (int k = 0; k < 10; k++) {
        P.p[k] = rays[k];
        Q.p[k] = rays_out ? rays_out[k] : rays[k];
        if (fp.load_mask & (1u << k)) {
            if (!P.p[k]) { set_error("program: null input row pointer (row %d)", k); return PXF_ERR_INVALID; }
            if (reinterpret_cast<uintptr_t>(P.p[k]) & 15) aligned = false;
        }
        if (fp.store_mask & (1u << k)) {
            if (!Q.p[k]) { set_error("program: null output row pointer (row %d)", k); return PXF_ERR_INVALID; }
            if (reinterpret_cast<uintptr_t>(Q.p[k]) & 15) aligned = false;
        }
    }
    {
        // statically specialised kernel for the canonical chains, else the interpreter below
        int rc = launch_chain(P, Q, num, fp, alive, aligned, s, partials, grid_out);
        if (rc != PXF_ERR_UNSUPPORTED) return rc;
        // ... else a kernel specialised for this op list at run time (NVRTC, cached), else the interpreter
        rc = jit_launch_chain(P, Q, num, fp, alive, aligned, s, partials, grid_out);
        if (rc != PXF_ERR_UNSUPPORTED) return rc;
    }
    // PXF_PROGRAM_VARIANT (tuning): 0 = two rays per thread (double2 rows), 1/3/4 = one ray per thread with
    // the register allocation capped for 1 / 3 / 4 resident CTAs per SM.  Measured on config 3's 12-op tail at
    // 5e7 rays (profiles/r01h_notes.md): 3.82 / 3.75 / 3.27 / 3.00 ms for 0 / 1 / 3 / 4 (5 and 6 CTAs/SM spill: 4.6 / 5.7 ms).
    static int variant = -1;
    if (variant < 0) { const char *e = getenv("PXF_PROGRAM_VARIANT"); variant = e ? atoi(e) : 4; }
    auto go = [&](auto kern, int64_t items) {
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, PXF_BLOCK, 0) != cudaSuccess || nb <= 0) { cudaGetLastError(); nb = 2; }
        const int grid = grid_for(items, PXF_BLOCK, nb);
        if (grid_out) *grid_out = grid;
        kern<<<grid, PXF_BLOCK, 0, s>>>(P, Q, num, alive, partials, fp);
    };
    if (fp.uses_aux) {
        if (fp.zern) { set_error("program: a Zernike surface and a grating fan in one program are not supported"); return PXF_ERR_UNSUPPORTED; }
        go(k_program<false, 3, false, true>, num);
    }
    else if (fp.zern) {
        // one ray per thread; PXF_ZERNPROG_MINB (tuning): register cap for 1 / 2 / 3 resident CTAs per SM.  Measured on
        // config 3's 14-op program at 5e7 rays: 9.40 / 6.28 / 6.01 ms (4: 6.28 ms); the same work as three launches
        // (transform, tracezern, 12-op tail) takes 7.0 ms
        static int zm = -1;
        if (zm < 0) { const char *e = getenv("PXF_ZERNPROG_MINB"); zm = e ? atoi(e) : 3; }
        if (zm == 1) go(k_program<false, 1, true>, num);
        else if (zm == 2) go(k_program<false, 2, true>, num);
        else go(k_program<false, 3, true>, num);
    }
    else if (aligned && variant == 0) go(k_program<true, 1>, (num + 1) >> 1);
    else if (variant == 3) go(k_program<false, 3>, num);
    else if (variant == 1) go(k_program<false, 1>, num);
    else go(k_program<false, 4>, num);
    count_launch();
    note_kernel(fp.zern ? "k_program<zernike> (interpreter)" : (fp.uses_aux ? "k_program<aux> (interpreter)" : "k_program (interpreter)"));
    return check_launch("k_program");
}

}  // namespace pxf

using namespace pxf;
// pxf_analysis.cu
namespace pxf { int sums_finalize(const double *partial, int nblocks, int ns, double *out_dev, cudaStream_t s); }

extern "C" int pxf_trace_program(double *const rays[10], int64_t num, const pxf_op *ops, int32_t nops,
                                 uint8_t *alive, pxf_stream_t stream)
{
    FusedProgram fp;
    int rc = build_program(fp, ops, nops);
    if (rc) return rc;
    return launch_program(rays, num, fp, alive, reinterpret_cast<cudaStream_t>(stream), nullptr, nullptr, nullptr);
}

extern "C" int pxf_trace_program_to(double *const rays_in[10], double *const rays_out[10], int64_t num,
                                    const pxf_op *ops, int32_t nops, uint8_t *alive, pxf_stream_t stream)
{
    if (!rays_out) { set_error("pxf_trace_program_to: null output table"); return PXF_ERR_INVALID; }
    FusedProgram fp;
    int rc = build_program(fp, ops, nops);
    if (rc) return rc;
    // out of place: rows that are read but never written must still appear in the output
    // bundle, so every row the program touches is stored
    fp.store_mask |= fp.load_mask;
    return launch_program(rays_in, num, fp, alive, reinterpret_cast<cudaStream_t>(stream), rays_out, nullptr, nullptr);
}

extern "C" int pxf_trace_program_aux(double *const rays_in[10], double *const rays_out[10], int64_t num,
                                     const pxf_op *ops, int32_t nops, uint8_t *alive, const pxf_program_aux *aux,
                                     double *sums_dev, void *scratch, pxf_stream_t stream)
{
    FusedProgram fp;
    int rc = build_program(fp, ops, nops, aux);
    if (rc) return rc;
    if (rays_out) fp.store_mask |= fp.load_mask;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (!sums_dev) return launch_program(rays_in, num, fp, alive, s, rays_out, nullptr, nullptr);
    if (!scratch) { set_error("pxf_trace_program_aux: sums need scratch"); return PXF_ERR_INVALID; }
    fp.load_mask |= (R_X | R_Y) & ~fp.store_mask;
    if (num == 0) {
        PXF_CUDA(cudaMemsetAsync(sums_dev, 0, 4 * sizeof(double), s));
        return PXF_OK;
    }
    int grid = 0;
    rc = launch_program(rays_in, num, fp, alive, s, rays_out, static_cast<double *>(scratch), &grid);
    if (rc) return rc;
    return sums_finalize(static_cast<const double *>(scratch), grid, 4, sums_dev, s);
}

extern "C" size_t pxf_zern_table_bytes(void) { return sizeof(ZernP); }

extern "C" int32_t pxf_zern_table_fill(const double *coeff, const int32_t *rorder, const int32_t *aorder, int32_t arrsize,
                                       double rad, int32_t opd, double nr, void *table_host)
{
    if (!coeff || !rorder || !aorder || arrsize <= 0 || !table_host) { set_error("pxf_zern_table_fill: bad argument"); return -1; }
    ZernP z;
    const int n = make_zern(z, coeff, rorder, aorder, arrsize, rad, opd != 0, nr);
    if (n < 0) { set_error("pxf_zern_table_fill: invalid Zernike table"); return -1; }
    memcpy(table_host, &z, sizeof(z));
    return n;
}

// ---- segmented programs -------------------------------------------------------------------------
static size_t seg_align(size_t v) { return (v + 15) & ~(size_t)15; }
static size_t seg_ops_offset(int nseg) { return seg_align(sizeof(SegHeader)) + seg_align((size_t)(nseg + 1) * 8); }

extern "C" size_t pxf_segmented_table_bytes(int32_t nops, int32_t nseg)
{
    if (nops < 1 || nseg < 1) return 0;
    size_t packs = seg_chain_bytes(nseg), jp = (size_t)nseg * jit_seg_pack_bytes(nops);
    return seg_align(seg_ops_offset(nseg) + (size_t)nseg * nops * sizeof(FusedOp)) + (packs > jp ? packs : jp);
}
static size_t seg_chain_offset(int nops, int nseg) { return seg_align(seg_ops_offset(nseg) + (size_t)nseg * nops * sizeof(FusedOp)); }

extern "C" int pxf_segmented_table_fill(const pxf_op *ops, int32_t nops, int32_t nseg, const int64_t *seg_start,
                                        void *table_host)
{
    if (!ops || !seg_start || !table_host || nops < 1 || nops > PXF_MAX_OPS || nseg < 1) {
        set_error("pxf_segmented_table_fill: bad argument");
        return PXF_ERR_INVALID;
    }
    if (seg_start[0] != 0) { set_error("segmented program: seg_start[0] must be 0"); return PXF_ERR_INVALID; }
    for (int sgm = 0; sgm < nseg; sgm++)
        if (seg_start[sgm + 1] < seg_start[sgm]) { set_error("segmented program: seg_start must be non-decreasing"); return PXF_ERR_INVALID; }
    char *base = static_cast<char *>(table_host);
    SegHeader *h = reinterpret_cast<SegHeader *>(base);
    memset(h, 0, seg_align(sizeof(SegHeader)));
    memcpy(base + seg_align(sizeof(SegHeader)), seg_start, (size_t)(nseg + 1) * 8);
    FusedOp *dst = reinterpret_cast<FusedOp *>(base + seg_ops_offset(nseg));
    h->nops = nops; h->nseg = nseg;
    for (int sgm = 0; sgm < nseg; sgm++) {
        FusedProgram fp;
        int rc = build_program(fp, ops + (size_t)sgm * nops, nops);
        if (rc) return rc;
        if (fp.zern || fp.uses_aux) { set_error("segmented program: PXF_OP_ZERNSURF / PXF_OP_GRATFAN are not supported"); return PXF_ERR_UNSUPPORTED; }
        if (sgm > 0)
            for (int k = 0; k < nops; k++)
                if (fp.ops[k].code != dst[k].code || fp.ops[k].row != dst[k].row) {   // dst[0..nops) = segment 0
                    set_error("segmented program: every segment must run the same opcode sequence (segment %d, op %d)", sgm, k);
                    return PXF_ERR_INVALID;
                }
        h->load_mask |= fp.load_mask;
        h->store_mask |= fp.store_mask;
        memcpy(dst + (size_t)sgm * nops, fp.ops, (size_t)nops * sizeof(FusedOp));
    }
    // a statically specialised kernel for the canonical chains: one parameter pack per segment behind the op table
    h->chain_id = seg_chain_fill(dst, nops, nseg, base + seg_chain_offset(nops, nseg));
    bool vig = false;
    for (int k = 0; k < nops; k++)
        if (dst[k].code == PXF_OP_VIGNETTE_MAG || dst[k].code == PXF_OP_VIGNETTE_BOX || dst[k].code == PXF_OP_VIGNETTE_ABS ||
            dst[k].code == PXF_OP_VIGNETTE_RHOGT) vig = true;
    if (h->chain_id == 0 || vig) {
        // any other op list: parameter packs for a kernel specialised at run time (pxf_jit.cu)
        const size_t pb = jit_seg_fill(dst, nops, nseg, base + seg_chain_offset(nops, nseg));
        if (pb) { h->chain_id = 2; h->jit_pack_bytes = (int)pb; }
        else if (vig) h->chain_id = 0;
    }
    return PXF_OK;
}

extern "C" int pxf_trace_program_segmented(double *const rays_in[10], double *const rays_out[10], int64_t num,
                                           const void *table_host, const void *table_dev, uint8_t *alive,
                                           pxf_stream_t stream)
{
    if (!rays_in || !table_host || !table_dev || num < 0) { set_error("pxf_trace_program_segmented: bad argument"); return PXF_ERR_INVALID; }
    if (sm_count() <= 0) { set_error("no CUDA device available (libpxf has no CPU fallback)"); return PXF_ERR_CUDA; }
    const SegHeader *h = static_cast<const SegHeader *>(table_host);
    const int nops = h->nops, nseg = h->nseg;
    const int64_t *hs = reinterpret_cast<const int64_t *>(static_cast<const char *>(table_host) + seg_align(sizeof(SegHeader)));
    if (nops < 1 || nops > PXF_MAX_OPS || nseg < 1 || hs[nseg] != num) {
        set_error("segmented program: table does not describe a bundle of %lld rays", (long long)num);
        return PXF_ERR_INVALID;
    }
    if (num == 0) return PXF_OK;
    unsigned LM = h->load_mask, SM = h->store_mask;
    if (rays_out) SM |= LM;
    RowPtrs P, Q;
    for (int k = 0; k < 10; k++) {
        P.p[k] = rays_in[k];
        Q.p[k] = rays_out ? rays_out[k] : rays_in[k];
        if ((LM & (1u << k)) && !P.p[k]) { set_error("segmented program: null input row pointer (row %d)", k); return PXF_ERR_INVALID; }
        if ((SM & (1u << k)) && !Q.p[k]) { set_error("segmented program: null output row pointer (row %d)", k); return PXF_ERR_INVALID; }
    }
    const FusedOp *hops = reinterpret_cast<const FusedOp *>(static_cast<const char *>(table_host) + seg_ops_offset(nseg));
    bool vig = false;
    for (int k = 0; k < nops; k++)
        if (hops[k].code == PXF_OP_VIGNETTE_MAG || hops[k].code == PXF_OP_VIGNETTE_BOX || hops[k].code == PXF_OP_VIGNETTE_ABS ||
            hops[k].code == PXF_OP_VIGNETTE_RHOGT) vig = true;
    if (vig && !alive) { set_error("program with a VIGNETTE op needs an alive array"); return PXF_ERR_INVALID; }
    const char *db = static_cast<const char *>(table_dev);
    const long long *dstart = reinterpret_cast<const long long *>(db + seg_align(sizeof(SegHeader)));
    const FusedOp *dops = reinterpret_cast<const FusedOp *>(db + seg_ops_offset(nseg));
    const size_t smem = (size_t)nops * sizeof(FusedOp);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (h->chain_id == 1 && !vig) {
        int rc = seg_chain_launch(h->chain_id, P, Q, num, alive, dstart, db + seg_chain_offset(nops, nseg), nseg, LM, SM, s);
        if (rc != PXF_ERR_UNSUPPORTED) return rc;
    }
    if (h->chain_id == 2) {
        int rc = jit_seg_launch(hops, nops, P, Q, num, alive, dstart, db + seg_chain_offset(nops, nseg), nseg, LM, SM, s);
        if (rc != PXF_ERR_UNSUPPORTED) return rc;
    }
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_program_seg<4>, PXF_BLOCK, smem) != cudaSuccess || nb <= 0) { cudaGetLastError(); nb = 2; }
    const int grid = grid_for(num, SEG_TILE, nb);
    k_program_seg<4><<<grid, PXF_BLOCK, smem, s>>>(P, Q, num, alive, dstart, dops, nops, nseg, LM, SM);
    count_launch();
    note_kernel("k_program_seg (interpreter)");
    return check_launch("k_program_seg");
}

extern "C" int pxf_trace_program_sums(double *const rays_in[10], double *const rays_out[10], int64_t num,
                                      const pxf_op *ops, int32_t nops, uint8_t *alive, double *sums_dev,
                                      void *scratch, pxf_stream_t stream)
{
    if (!sums_dev || !scratch) { set_error("pxf_trace_program_sums: null sums/scratch"); return PXF_ERR_INVALID; }
    FusedProgram fp;
    int rc = build_program(fp, ops, nops);
    if (rc) return rc;
    if (rays_out) fp.store_mask |= fp.load_mask;
    // the sums need the final x,y in registers: make sure they are loaded even if the program
    // itself never reads them
    fp.load_mask |= (R_X | R_Y) & ~fp.store_mask;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (num == 0) {
        PXF_CUDA(cudaMemsetAsync(sums_dev, 0, 4 * sizeof(double), s));
        return PXF_OK;
    }
    int grid = 0;
    rc = launch_program(rays_in, num, fp, alive, s, rays_out, static_cast<double *>(scratch), &grid);
    if (rc) return rc;
    return sums_finalize(static_cast<const double *>(scratch), grid, 4, sums_dev, s);
}
