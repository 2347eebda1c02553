// Statically specialised fused kernels for the reference's canonical op chains.
//
// The generic interpreter (pxf_fused.cu) pays for its generality: every scalar parameter is an
// indexed constant load (ADU pipe), the opcode switch costs branches and i-cache, and the
// register allocation is the maximum over all ops (pow, sincos, ...).  For the chains the
// reference's example scripts actually run -- Wolter-I pair to the focal plane
// (examples/axro/singlePassAlignment.py:246-269), Wolter-Schwarzschild pair
// (examples/axro/axialHeights.py:77-113), SPO pair (examples/arcus/cat.py:219-280) -- the op
// list is known at compile time: a variadic template composes the SAME per-op device functions
// (pxf_ray.cuh) into one straight-line kernel whose parameters are direct constant-bank operands.
// Results are bit-identical to the interpreter and to the per-routine kernels.
#include <stdlib.h>
#include "pxf_program.h"

namespace pxf {

struct NoP { int unused; };

// ---- op functors: P = folded parameter block (what build_program stores in FusedOp::q) ----
// apply: one ray; apply2: two rays of the same thread (default: one after the other; the Newton
// surfaces advance both rays through the same loop so their dependency chains interleave).
#define PXF_CHAIN_OP(NAME, CODEV, PTYPE, CALL)                                        \
    struct NAME {                                                                     \
        using P = PTYPE;                                                              \
        static constexpr int CODE = CODEV;                                            \
        PXF_DEV static bool apply(Ray &r, const P &p) { CALL; return true; }          \
        PXF_DEV static void apply2(Ray *rr, const P &p) { apply(rr[0], p); apply(rr[1], p); } \
    };
#define PXF_CHAIN_OP2(NAME, CODEV, PTYPE, CALL, CALL2)                                \
    struct NAME {                                                                     \
        using P = PTYPE;                                                              \
        static constexpr int CODE = CODEV;                                            \
        PXF_DEV static bool apply(Ray &r, const P &p) { CALL; return true; }          \
        PXF_DEV static void apply2(Ray *rr, const P &p) { CALL2; }                    \
    };
PXF_CHAIN_OP(CTransform, PXF_OP_TRANSFORM, TransformP, op_transform(r, p))
PXF_CHAIN_OP(CITransform, PXF_OP_ITRANSFORM, TransformP, op_itransform(r, p))
PXF_CHAIN_OP(CReflect, PXF_OP_REFLECT, NoP, (void)p; op_reflect(r))
PXF_CHAIN_OP(CFlat, PXF_OP_FLAT, NoP, (void)p; op_flat(r, false, 0.))
PXF_CHAIN_OP2(CWolterPrimary, PXF_OP_WOLTERPRIMARY, WolterP, op_wolterprimary(r, p), op_wolterprimary_w<2>(rr, p))
PXF_CHAIN_OP2(CWolterSecondary, PXF_OP_WOLTERSECONDARY, WolterP, op_woltersecondary(r, p), op_woltersecondary_w<2>(rr, p))
PXF_CHAIN_OP(CWsPrimary, PXF_OP_WSPRIMARY, WSP, op_wsprimary(r, p))
PXF_CHAIN_OP(CWsSecondary, PXF_OP_WSSECONDARY, WSP, op_wssecondary(r, p))
PXF_CHAIN_OP(CSpoCone, PXF_OP_SPOCONE, SpoP, op_spocone(r, p))
struct KickP { double dl, dm, sn; };
PXF_CHAIN_OP(CKick, PXF_OP_KICK, KickP,
             r.l = r.l + p.dl; r.m = r.m + p.dm; r.n = p.sn * sqrt(1. - sq(r.l) - sq(r.m)))

// ---- parameter pack and straight-line composition ----
template <class... Ops> struct ChainP;
template <> struct ChainP<> { int unused; };
template <class Op, class... Rest> struct ChainP<Op, Rest...> {
    typename Op::P head;
    ChainP<Rest...> tail;
};

template <class... Ops> struct Chain;
template <> struct Chain<> {
    static constexpr int N = 0;
    PXF_DEV static bool run(Ray &, const ChainP<> &) { return true; }
    PXF_DEV static void run2(Ray *, const ChainP<> &) {}
    static void fill(ChainP<> &, const FusedOp *) {}
    static bool match(const FusedOp *, int n) { return n == 0; }
};
template <class Op, class... Rest> struct Chain<Op, Rest...> {
    static constexpr int N = 1 + sizeof...(Rest);
    PXF_DEV static bool run(Ray &r, const ChainP<Op, Rest...> &p)
    {
        if (!Op::apply(r, p.head)) return false;
        return Chain<Rest...>::run(r, p.tail);
    }
    PXF_DEV static void run2(Ray *rr, const ChainP<Op, Rest...> &p)      // chains carry no vignette ops
    {
        Op::apply2(rr, p.head);
        Chain<Rest...>::run2(rr, p.tail);
    }
    static void fill(ChainP<Op, Rest...> &cp, const FusedOp *ops)
    {
        static_assert(sizeof(typename Op::P) <= sizeof(ops->q), "parameter block too large");
        memcpy(&cp.head, ops->q, sizeof(typename Op::P));
        Chain<Rest...>::fill(cp.tail, ops + 1);
    }
    static bool match(const FusedOp *ops, int n)
    {
        return n >= 1 && ops->code == Op::CODE && Chain<Rest...>::match(ops + 1, n - 1);
    }
};

// ---- kernel -------------------------------------------------------------------------------
// MODE 1: one ray per thread-iteration (8-byte accesses, also the path for unaligned rows)
// MODE 2: two rays per thread-iteration (double2 accesses), traced one after the other
// MODE 3: two rays per thread-iteration advanced together through the Newton loops
// PF    : software prefetch -- the loads of the NEXT iteration are issued before the current
//         rays are traced, so every warp keeps ~1.5-3 KB of loads in flight during its compute
//         phase.  Without it the kernel is limited by bytes in flight: a warp is either waiting
//         for its loads or computing, and with ~70 % of the time spent computing too few loads
//         are outstanding to keep HBM busy (measured: 4.1 ms vs 2.6 ms for the same traffic
//         with trivial compute, profiles/r01_notes.md).
// LMc/SMc: row masks known at compile time (0 = use the runtime masks LM/SM), so that rows the
//         chain never loads cost no registers in the prefetch buffer.
// MINB  : resident CTAs per SM the register allocation must allow.
template <unsigned M>
PXF_DEV void cload1(Ray &r, const RowPtrs &P, unsigned LM, int64_t i)
{
    if (M == 0) { fload1(r, P, LM, i); return; }
    r.opd = (M & R_OPD) ? P.p[0][i] : 0.;
    r.x = (M & R_X) ? P.p[1][i] : 0.;  r.y = (M & R_Y) ? P.p[2][i] : 0.;  r.z = (M & R_Z) ? P.p[3][i] : 0.;
    r.l = (M & R_L) ? P.p[4][i] : 0.;  r.m = (M & R_M) ? P.p[5][i] : 0.;  r.n = (M & R_N) ? P.p[6][i] : 0.;
    r.ux = (M & R_UX) ? P.p[7][i] : 0.; r.uy = (M & R_UY) ? P.p[8][i] : 0.; r.uz = (M & R_UZ) ? P.p[9][i] : 0.;
}
#define CLD2(bit, k, f)                                                               \
    if (M & bit) { double2 v = *reinterpret_cast<const double2 *>(P.p[k] + i); a.f = v.x; b.f = v.y; } \
    else { a.f = 0.; b.f = 0.; }
template <unsigned M>
PXF_DEV void cload2(Ray &a, Ray &b, const RowPtrs &P, unsigned LM, int64_t i)
{
    if (M == 0) { fload2(a, b, P, LM, i); return; }
    CLD2(R_OPD, 0, opd) CLD2(R_X, 1, x) CLD2(R_Y, 2, y) CLD2(R_Z, 3, z) CLD2(R_L, 4, l)
    CLD2(R_M, 5, m) CLD2(R_N, 6, n) CLD2(R_UX, 7, ux) CLD2(R_UY, 8, uy) CLD2(R_UZ, 9, uz)
}
template <unsigned M>
PXF_DEV void cstore1(const Ray &r, const RowPtrs &P, unsigned SM, int64_t i)
{
    fstore1(r, P, M ? M : SM, i);
}
template <unsigned M>
PXF_DEV void cstore2(const Ray &a, const Ray &b, const RowPtrs &P, unsigned SM, int64_t i)
{
    fstore2(a, b, P, M ? M : SM, i);
}

template <class C, class CP, int MODE, bool PF, int MINB, unsigned LMc, unsigned SMc>
__global__ void __launch_bounds__(PXF_BLOCK, MINB)
k_chain(const RowPtrs P, const RowPtrs Q, const int64_t num, uint8_t *__restrict__ alive,
        double *__restrict__ partials, const unsigned LM, const unsigned SM, const __grid_constant__ CP prm)
{
    double cnt = 0., sx = 0., sy = 0.;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    if (MODE >= 2) {
        const int64_t npair = num >> 1;
        int64_t q = tid;
        Ray cur[2], nxt[2];
        if (PF && q < npair) cload2<LMc>(cur[0], cur[1], P, LM, q << 1);
        while (q < npair) {
            const int64_t qn = q + nthr;
            if (PF) {
                if (qn < npair) cload2<LMc>(nxt[0], nxt[1], P, LM, qn << 1);
            } else {
                cload2<LMc>(cur[0], cur[1], P, LM, q << 1);
            }
            if (MODE == 3) {
                C::run2(cur, prm);
            } else {
                C::run(cur[0], prm);
                C::run(cur[1], prm);
            }
            const int64_t i = q << 1;
            cstore2<SMc>(cur[0], cur[1], Q, SM, i);
            if (alive) { alive[i] = 1; alive[i + 1] = 1; }
            cnt += 2.; sx += cur[0].x; sx += cur[1].x; sy += cur[0].y; sy += cur[1].y;
            if (PF) { cur[0] = nxt[0]; cur[1] = nxt[1]; }
            q = qn;
        }
        if ((num & 1) && tid == 0) {
            const int64_t i = num - 1;
            Ray a;
            cload1<LMc>(a, P, LM, i);
            const bool ka = C::run(a, prm);
            cstore1<SMc>(a, Q, SM, i);
            if (alive) alive[i] = ka ? 1 : 0;
            if (ka) { cnt += 1.; sx += a.x; sy += a.y; }
        }
    } else {
        int64_t i = tid;
        Ray cur, nxt;
        if (PF && i < num) cload1<LMc>(cur, P, LM, i);
        while (i < num) {
            const int64_t in = i + nthr;
            if (PF) {
                if (in < num) cload1<LMc>(nxt, P, LM, in);
            } else {
                cload1<LMc>(cur, P, LM, i);
            }
            const bool ka = C::run(cur, prm);
            cstore1<SMc>(cur, Q, SM, i);
            if (alive) alive[i] = ka ? 1 : 0;
            if (ka) { cnt += 1.; sx += cur.x; sy += cur.y; }
            if (PF) cur = nxt;
            i = in;
        }
    }
    if (partials) centroid_block_reduce(cnt, sx, sy, partials);
}

template <class C, class CP, int MODE, bool PF, int MINB, unsigned LMc, unsigned SMc>
static int launch_variant(const RowPtrs &P, const RowPtrs &Q, int64_t num, uint8_t *alive, unsigned LM, unsigned SM,
                          const CP &cp, cudaStream_t s, double *partials, int *grid_out)
{
    static int ctas = 0;
    auto kern = k_chain<C, CP, MODE, PF, MINB, LMc, SMc>;
    if (ctas == 0) {
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, PXF_BLOCK, 0) != cudaSuccess || nb <= 0) {
            cudaGetLastError();
            nb = MINB;
        }
        ctas = nb;
    }
    const int64_t items = MODE >= 2 ? ((num + 1) >> 1) : num;
    const int grid = grid_for(items, PXF_BLOCK, ctas);
    if (grid_out) *grid_out = grid;
    kern<<<grid, PXF_BLOCK, 0, s>>>(P, Q, num, alive, partials, LM, SM, cp);
    count_launch();
    return check_launch("k_chain");
}

// PXF_CHAIN_VARIANT=<mode><minb><pf> (e.g. 231 = two rays in sequence, 3 CTAs/SM, prefetch)
// overrides the tuned default of the tunable chain; for tuning only.
static int variant_override()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("PXF_CHAIN_VARIANT");
        v = e ? atoi(e) : 0;
    }
    return v;
}

// LMc/SMc: the row masks this chain has when traced in place or out of place (checked against
// build_program's liveness result at run time; any other masks take the runtime-mask kernel).
template <bool TUNABLE, unsigned LMc, unsigned SMc, class... Ops>
static int try_chain(const RowPtrs &P, const RowPtrs &Q, int64_t num, const FusedProgram &fp, uint8_t *alive,
                     bool aligned, cudaStream_t s, double *partials, int *grid_out)
{
    using C = Chain<Ops...>;
    using CP = ChainP<Ops...>;
    if (fp.nops != C::N || !C::match(fp.ops, fp.nops)) return PXF_ERR_UNSUPPORTED;
    CP cp;
    memset(&cp, 0, sizeof(cp));
    C::fill(cp, fp.ops);
    const unsigned LM = fp.load_mask, SM = fp.store_mask;
    const bool stat = LMc != 0 && LM == LMc && SM == SMc;
#define PXF_LV(MODE, PF, MINB)                                                                              \
    (stat ? launch_variant<C, CP, MODE, PF, MINB, LMc, SMc>(P, Q, num, alive, LM, SM, cp, s, partials, grid_out) \
          : launch_variant<C, CP, MODE, false, MINB, 0u, 0u>(P, Q, num, alive, LM, SM, cp, s, partials, grid_out))
    if (!aligned) return PXF_LV(1, true, 3);
    if constexpr (TUNABLE) {
        switch (variant_override()) {
            case 130: return PXF_LV(1, false, 3);
            case 131: return PXF_LV(1, true, 3);
            case 141: return PXF_LV(1, true, 4);
            case 220: return PXF_LV(2, false, 2);
            case 221: return PXF_LV(2, true, 2);
            case 230: return PXF_LV(2, false, 3);
            case 231: return PXF_LV(2, true, 3);
            case 320: return PXF_LV(3, false, 2);
            case 321: return PXF_LV(3, true, 2);
            case 331: return PXF_LV(3, true, 3);
            default: break;
        }
    }
    return PXF_LV(2, false, 3);      // tuned on B200 (profiles/r01_notes.md): two rays in sequence, 3 CTAs/SM, no prefetch
#undef PXF_LV
}

int launch_chain(const RowPtrs &P, const RowPtrs &Q, int64_t num, const FusedProgram &fp, uint8_t *alive,
                 bool aligned, cudaStream_t s, double *partials, int *grid_out)
{
    static int disabled = -1;
    if (disabled < 0) {
        const char *e = getenv("PXF_NO_SPECIALIZE");     // force the generic interpreter (tests, A/B timing)
        disabled = (e && e[0] == '1') ? 1 : 0;
    }
    if (disabled || fp.has_vignette) return PXF_ERR_UNSUPPORTED;
    int rc;
    // Wolter-I pair to the focal plane (BASELINE config 1)
    rc = try_chain<true, (R_POS | R_DIR), R_NINE, CTransform, CWolterPrimary, CReflect, CWolterSecondary, CReflect, CFlat>(P, Q, num, fp, alive, aligned, s, partials, grid_out);
    if (rc != PXF_ERR_UNSUPPORTED) return rc;
    rc = try_chain<false, 0u, 0u, CWolterPrimary, CReflect, CWolterSecondary, CReflect, CFlat>(P, Q, num, fp, alive, aligned, s, partials, grid_out);
    if (rc != PXF_ERR_UNSUPPORTED) return rc;
    rc = try_chain<false, 0u, 0u, CTransform, CWolterPrimary, CReflect, CWolterSecondary, CReflect>(P, Q, num, fp, alive, aligned, s, partials, grid_out);
    if (rc != PXF_ERR_UNSUPPORTED) return rc;
    // Wolter-Schwarzschild pair with the field-angle kick (BASELINE config 2)
    rc = try_chain<false, 0u, 0u, CTransform, CWsPrimary, CKick, CReflect, CWsSecondary, CReflect>(P, Q, num, fp, alive, aligned, s, partials, grid_out);
    if (rc != PXF_ERR_UNSUPPORTED) return rc;
    // SPO primary/secondary pair (BASELINE config 4, per-shell part)
    rc = try_chain<false, 0u, 0u, CTransform, CSpoCone, CReflect, CSpoCone, CReflect, CTransform>(P, Q, num, fp, alive, aligned, s, partials, grid_out);
    if (rc != PXF_ERR_UNSUPPORTED) return rc;
    // focus step: move the plane and trace to it (surfaces.focus, surfaces.py:502-510)
    rc = try_chain<false, 0u, 0u, CTransform, CFlat>(P, Q, num, fp, alive, aligned, s, partials, grid_out);
    return rc;
}

}  // namespace pxf
