// stream_bwd.cuh - argument blocks of the streaming backward stages (block_stream_bwd.cu) and of the edge-weight
// gradient kernel (block_gw.cu), shared with the entry point that sequences them (block_bwd.cu).
#pragma once
#include "common.cuh"

enum { BW_X3 = 0, BW_X2A = 1, BW_X2B = 2, BW_X1 = 3, BW_BA = 4 };

struct StreamBwdArgs {
    glrgtv_shape s;
    glrgtv_block_params p;
    glrgtv_block_grads gr;
    const float* z;      // x2 | x1 | x1 | bA | y
    const float* src0;   // gout | gout | gout | gx1 | gbA
    const float* src1;   // X2A, X2B: gx2
    const float* op0;    // X3: r1 | X2A: r1 | X2B: gx1 (read-modify-write) | BA: gout
    const float* op1;    // X3: bB | BA: gx2
    const float* op2;    // X3: x (skip)
    const float *wT0, *wL0, *wT1, *wL1, *cT0, *cT1;
    float* out;
    int nch, band_rows, n_bands, n_strips;
};

struct GwArgs {
    glrgtv_shape s;             // FINE geometry
    glrgtv_block_params p;
    float* ggamma0; float* ggamma1;
    const float* z;
    const float* src0;
    const float* src1;
    const float *wT0, *wT1;
    float *gwT0, *gwL0, *gwT1, *gwL1;
    int assign;
};

template <int MODE> int glr_stream_bwd_stage(StreamBwdArgs a, int slot, void* stream);
template <int MODE> int glr_gw_stage(const GwArgs& a, int slot, void* stream);
// streaming form (block_gw_stream.cu); GLRGTV_ERR_UNSUPPORTED when the shape is outside its range
template <int MODE> int glr_gw_stream_stage(const GwArgs& a, int slot, void* stream);
