// Fused-program representation shared by pxf_fused.cu (kernel + builder) and pxf_host.cu
// (host-buffer streaming entry point).  Not part of the ABI.
#pragma once
#include "pxf_internal.h"
#include "pxf_params.h"

namespace pxf {

#define FOP_PARAM_DOUBLES 28   // sizeof(WSP)/8 is the largest folded parameter block

struct FusedOp {
    int code;
    int row;                 // VIGNETTE_BOX / ABS: bundle row index
    double q[FOP_PARAM_DOUBLES];
};
struct FusedProgram {
    int nops;
    unsigned load_mask, store_mask;
    int has_vignette;
    FusedOp ops[PXF_MAX_OPS];
};

static_assert(sizeof(WSP) <= FOP_PARAM_DOUBLES * 8, "WSP does not fit a fused op slot");
static_assert(sizeof(ConicP) <= FOP_PARAM_DOUBLES * 8, "ConicP does not fit");
static_assert(sizeof(WolterSineP) <= FOP_PARAM_DOUBLES * 8, "WolterSineP does not fit");

// pxf_fused.cu
int build_program(FusedProgram &fp, const pxf_op *ops, int nops);
int launch_program(double *const rays[10], int64_t num, const FusedProgram &fp, uint8_t *alive, cudaStream_t s,
                   double *const rays_out[10] = nullptr);

}  // namespace pxf
