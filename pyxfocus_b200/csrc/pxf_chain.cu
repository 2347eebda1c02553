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
#define PXF_CHAIN_OP(NAME, CODEV, PTYPE, CALL)                                        \
    struct NAME {                                                                     \
        using P = PTYPE;                                                              \
        static constexpr int CODE = CODEV;                                            \
        PXF_DEV static bool apply(Ray &r, const P &p) { CALL; return true; }          \
    };
PXF_CHAIN_OP(CTransform, PXF_OP_TRANSFORM, TransformP, op_transform(r, p))
PXF_CHAIN_OP(CITransform, PXF_OP_ITRANSFORM, TransformP, op_itransform(r, p))
PXF_CHAIN_OP(CReflect, PXF_OP_REFLECT, NoP, (void)p; op_reflect(r))
PXF_CHAIN_OP(CFlat, PXF_OP_FLAT, NoP, (void)p; op_flat(r, false, 0.))
PXF_CHAIN_OP(CWolterPrimary, PXF_OP_WOLTERPRIMARY, WolterP, op_wolterprimary(r, p))
PXF_CHAIN_OP(CWolterSecondary, PXF_OP_WOLTERSECONDARY, WolterP, op_woltersecondary(r, p))
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

// RPT rays per thread per iteration (2 = double2 accesses), MINB resident CTAs per SM the
// register allocation must allow.
template <class C, class CP, int RPT, int MINB>
__global__ void __launch_bounds__(PXF_BLOCK, MINB)
k_chain(const RowPtrs P, const RowPtrs Q, const int64_t num, uint8_t *__restrict__ alive,
        const unsigned LM, const unsigned SM, const __grid_constant__ CP prm)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    if (RPT == 2) {
        const int64_t npair = num >> 1;
        for (int64_t q = tid; q < npair; q += nthr) {
            const int64_t i = q << 1;
            Ray a, b;
            fload2(a, b, P, LM, i);
            const bool ka = C::run(a, prm);
            const bool kb = C::run(b, prm);
            fstore2(a, b, Q, SM, i);
            if (alive) { alive[i] = ka ? 1 : 0; alive[i + 1] = kb ? 1 : 0; }
        }
        if ((num & 1) && tid == 0) {
            const int64_t i = num - 1;
            Ray a;
            fload1(a, P, LM, i);
            const bool ka = C::run(a, prm);
            fstore1(a, Q, SM, i);
            if (alive) alive[i] = ka ? 1 : 0;
        }
    } else {
        for (int64_t i = tid; i < num; i += nthr) {
            Ray a;
            fload1(a, P, LM, i);
            const bool ka = C::run(a, prm);
            fstore1(a, Q, SM, i);
            if (alive) alive[i] = ka ? 1 : 0;
        }
    }
}

template <class C, class CP, int RPT, int MINB>
static int launch_variant(const RowPtrs &P, const RowPtrs &Q, int64_t num, uint8_t *alive, unsigned LM, unsigned SM,
                          const CP &cp, cudaStream_t s)
{
    static int ctas = 0;
    if (ctas == 0) {
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_chain<C, CP, RPT, MINB>, PXF_BLOCK, 0) != cudaSuccess ||
            nb <= 0) {
            cudaGetLastError();
            nb = MINB;
        }
        ctas = nb;
    }
    const int64_t items = RPT == 2 ? ((num + 1) >> 1) : num;
    const int grid = grid_for(items, PXF_BLOCK, ctas);
    k_chain<C, CP, RPT, MINB><<<grid, PXF_BLOCK, 0, s>>>(P, Q, num, alive, LM, SM, cp);
    count_launch();
    return check_launch("k_chain");
}

// PXF_CHAIN_VARIANT=<rpt><minb> (e.g. 14, 23) overrides the tuned default; for tuning only.
static int variant_override()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("PXF_CHAIN_VARIANT");
        v = e ? atoi(e) : 0;
    }
    return v;
}

template <bool TUNABLE, class... Ops>
static int try_chain(const RowPtrs &P, const RowPtrs &Q, int64_t num, const FusedProgram &fp, uint8_t *alive,
                     bool aligned, cudaStream_t s)
{
    using C = Chain<Ops...>;
    using CP = ChainP<Ops...>;
    if (fp.nops != C::N || !C::match(fp.ops, fp.nops)) return PXF_ERR_UNSUPPORTED;
    CP cp;
    memset(&cp, 0, sizeof(cp));
    C::fill(cp, fp.ops);
    const unsigned LM = fp.load_mask, SM = fp.store_mask;
    if (!aligned) return launch_variant<C, CP, 1, 3>(P, Q, num, alive, LM, SM, cp, s);
    if constexpr (TUNABLE) {
        switch (variant_override()) {
            case 13: return launch_variant<C, CP, 1, 3>(P, Q, num, alive, LM, SM, cp, s);
            case 14: return launch_variant<C, CP, 1, 4>(P, Q, num, alive, LM, SM, cp, s);
            case 22: return launch_variant<C, CP, 2, 2>(P, Q, num, alive, LM, SM, cp, s);
            case 23: return launch_variant<C, CP, 2, 3>(P, Q, num, alive, LM, SM, cp, s);
            default: break;
        }
    }
    return launch_variant<C, CP, 2, 2>(P, Q, num, alive, LM, SM, cp, s);
}

int launch_chain(const RowPtrs &P, const RowPtrs &Q, int64_t num, const FusedProgram &fp, uint8_t *alive,
                 bool aligned, cudaStream_t s)
{
    static int disabled = -1;
    if (disabled < 0) {
        const char *e = getenv("PXF_NO_SPECIALIZE");     // force the generic interpreter (tests, A/B timing)
        disabled = (e && e[0] == '1') ? 1 : 0;
    }
    if (disabled || fp.has_vignette) return PXF_ERR_UNSUPPORTED;
    int rc;
    // Wolter-I pair to the focal plane (BASELINE config 1)
    rc = try_chain<true, CTransform, CWolterPrimary, CReflect, CWolterSecondary, CReflect, CFlat>(P, Q, num, fp, alive, aligned, s);
    if (rc != PXF_ERR_UNSUPPORTED) return rc;
    rc = try_chain<false, CWolterPrimary, CReflect, CWolterSecondary, CReflect, CFlat>(P, Q, num, fp, alive, aligned, s);
    if (rc != PXF_ERR_UNSUPPORTED) return rc;
    rc = try_chain<false, CTransform, CWolterPrimary, CReflect, CWolterSecondary, CReflect>(P, Q, num, fp, alive, aligned, s);
    if (rc != PXF_ERR_UNSUPPORTED) return rc;
    // Wolter-Schwarzschild pair with the field-angle kick (BASELINE config 2)
    rc = try_chain<false, CTransform, CWsPrimary, CKick, CReflect, CWsSecondary, CReflect>(P, Q, num, fp, alive, aligned, s);
    if (rc != PXF_ERR_UNSUPPORTED) return rc;
    // SPO primary/secondary pair (BASELINE config 4, per-shell part)
    rc = try_chain<false, CTransform, CSpoCone, CReflect, CSpoCone, CReflect, CTransform>(P, Q, num, fp, alive, aligned, s);
    if (rc != PXF_ERR_UNSUPPORTED) return rc;
    // focus step: move the plane and trace to it (surfaces.focus, surfaces.py:502-510)
    rc = try_chain<false, CTransform, CFlat>(P, Q, num, fp, alive, aligned, s);
    return rc;
}

}  // namespace pxf
