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

// ---- second-generation backward stages (bw2.cu)
struct B2Args {
    glrgtv_shape s;             // FULL-resolution geometry of the block
    glrgtv_block_params p;
    glrgtv_block_grads gr;
    const float* z;             // x2 | x1 | x1 | bA | y          [B,C,H,W]
    const float* src;           // gout | gA | gB | gx1 | gbA     [B,C,H,W]
    const float* op0;           // fine: X3 r1 | X2A r1 | X2B gx1 | BA gB
    const float* op1;           // fine: X3 bB | X2A gx2 | BA gout
    const float* op2;           // fine: X3 x (skip)
    const float* vc_in;         // fine: the half-resolution launch's result [B,C,H/2,W/2]
    float* vc_out;              // coarse
    const float* wT;            // this level: GTV weights [B,G,4,LH,LW]
    const float* wL;            // this level: GLR weights [B,G,4,LH,LW]
    float* gwT;                 // this level: edge-weight gradients [B,G,4,LH,LW], ACCUMULATED with red.global.add by every stage
    float* gwL;
    float* out0; float* out1; float* out2;
    int lg;                     // log2 of the floats of one shared-memory row (2 x lanes per row, a power of two)
    int nch, n_parts;           // channels per CTA, CTAs per graph
    int band_rows, n_bands;
};

// one backward stage: the half-resolution launch (level-1 operands, writes vc), then the full-resolution launch of `a`
template <int MODE>
int glr_bw2_stage(B2Args a, const float* wT1, const float* wL1, float* gwT1, float* gwL1, float* vc, int slot, void* stream);
