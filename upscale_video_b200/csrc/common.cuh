// common.cuh -- shared host/device definitions for libb2sr (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2sr {

// One "plane" = one tile of the reference's tiling (reference upscale_processing.py:398-434): the halo'd rectangle
// [fy0, fy0+Ht) x [fx0, fx0+Wt) of frame `frame` is run through the network on its own (zero padding at ITS
// borders), and only the core rectangle [cy0,cy1) x [cx0,cx1) (frame coordinates) is written out.
struct PlaneDev {
    int32_t frame;
    int32_t fy0, fx0;
    int32_t Ht, Wt;
    int32_t cy0, cy1, cx0, cx1;
    int32_t gplane;   // index of the plane inside its size group (TMA coordinate 3)
    int64_t pix_off;  // first pixel of this plane in the per-layer activation buffers
};

// Work item of the tcgen05 convolution kernel: a band of `w` (<= 128) columns x `rows` rows of one plane.
// w <= 0 marks a band that does not exist in this plane (pipelined mode only: the CTA just advances its counters).
struct TcItem {
    int32_t map;    // index into the tensor-map array (plane size group)
    int32_t plane;  // plane index inside its group (TMA coordinate 3 of the per-layer buffers)
    int32_t x0, y0;
    int32_t rows, w;
    int32_t Ht, Wt;
    int64_t pix_off;  // as PlaneDev::pix_off
    int32_t frame, fy0, fx0;
    int32_t cy0, cy1, cx0, cx1;
    int32_t grow0;  // pipelined mode: global row index of this plane's row 0 (rows of all planes of a pass, in order)
};

// Parameters of one convolution layer.  The layer-by-layer kernel takes one by value; the pipelined kernel reads
// an array of them (one per layer) from global memory.
struct TcParams {
    const CUtensorMap* maps;  // device array of input tensor maps, indexed by map_base + TcItem::map
    int32_t map_base;
    const TcItem* items;
    const int32_t* item_first;  // layer mode: CTA k processes items [item_first[k], item_first[k+1]);
                                // pipelined mode: band b processes items [item_first[b], item_first[b+1])
    int32_t n_items;
    const uint8_t* wimg;  // pre-swizzled shared-memory image of this layer's weights: [kx][(2-ky)*NOUT + o][CPIX halfs]
    const float* bias;    // [NOUT]
    const float* slope;   // [NOUT] (PReLU epilogue)
    float acc_scale;      // v = acc * acc_scale + bias  (1/255 for the first layer fed with raw 0..255 pixels)
    void* out;            // PReLU epilogue: __half activations; shuffle epilogue: frames (u8 or float), packed
    const uint8_t* frames_in;  // shuffle epilogue: packed u8 input frames (residual branch)
    int32_t frame_h, frame_w;  // input frame size
    int32_t scale;             // pixel-shuffle factor
    // ---- pipelined mode (one CTA per (layer, band), activations in L2-resident row rings) ----
    int32_t ring_in, ring_out;  // input / output of this layer is a row ring (not a full per-plane buffer)
    int32_t RR;                 // rows per ring
    int32_t Wmax;               // ring row pitch in pixels
    int32_t nb;                 // band CTAs per layer
    // counters, one per band, B2SR_FLAG_STRIDE words apart:
    uint32_t* done_in;          // rows published by the previous layer's band CTAs
    uint32_t* done_out;         // rows this layer's band CTAs have published
    uint32_t* cons_self;        // input rows this layer's band CTAs have pulled into shared memory
    uint32_t* cons_next;        // the same counters of the next layer (back-pressure on this layer's ring)
    long long* dbg;             // pipelined mode, optional: this CTA's 8 stall-accounting words
    int32_t direct_in;          // pipelined mode, first layer: read the packed u8 frames directly (no fp16 input planes in HBM)
    const uint8_t* frames_end;  // one past the last byte of frames_in (bounds of the aligned word copies)
};

#define B2SR_PIPE_MAX_LAYERS 20
#define B2SR_FLAG_STRIDE 32  // uint32 words between the counters of neighbouring bands: one 128-byte line each, so that
                             // pollers of different CTAs do not hammer the same L2 sector
struct PipeParams {  // passed by value (kernel parameter space)
    TcParams layers[B2SR_PIPE_MAX_LAYERS];
    int32_t n_layers;
    int32_t nb;
    long long* dbg;  // optional [n_layers * nb][8] stall accounting (B2SR_OPT_PIPE_DEBUG)
};

#define B2SR_SMEM_LIMIT (227 * 1024)

}  // namespace b2sr
