// common.cuh -- shared host/device definitions for libb2sr (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2sr {

// One "plane" = one tile of the reference's tiling (reference upscale_processing.py:398-434): the halo'd
// rectangle [fy0, fy0+Ht) x [fx0, fx0+Wt) of frame `frame` is run through the network on its own (zero padding
// at ITS borders), and only the core rectangle [cy0,cy1) x [cx0,cx1) (frame coordinates) is written out.
struct PlaneDev {
    int32_t frame;
    int32_t fy0, fx0;
    int32_t Ht, Wt;
    int32_t cy0, cy1, cx0, cx1;
    int32_t gplane;  // index of the plane inside its size group (TMA coordinate 3)
    int64_t pix_off;  // first pixel of this plane in the activation buffers
};

// Work item of the tcgen05 convolution kernel: a band of `w` (<= 128) columns x `rows` rows of one plane.
struct TcItem {
    int32_t map;    // index into the tensor-map array (plane group; the launch adds the ping/pong offset)
    int32_t plane;  // plane index inside its group (TMA coordinate 3)
    int32_t x0, y0;
    int32_t rows, w;
    int32_t Ht, Wt;
    int64_t pix_off;  // as PlaneDev::pix_off
    int32_t frame, fy0, fx0;
    int32_t cy0, cy1, cx0, cx1;
    int32_t pad_;
};

enum { EPI_PRELU = 0, EPI_SHUFFLE_U8 = 1, EPI_SHUFFLE_F32 = 2 };

struct TcParams {
    const CUtensorMap* maps;  // device array of activation tensor maps
    int32_t map_base;         // added to TcItem::map (selects ping or pong buffer)
    const TcItem* items;
    const int32_t* item_first;  // CTA k processes items [item_first[k], item_first[k+1])
    int32_t n_items;
    const uint8_t* wimg;  // pre-swizzled shared-memory image of this layer's weights: [tap][NOUT rows][CPIX halfs]
    const float* bias;    // [NOUT]
    const float* slope;   // [NOUT] (EPI_PRELU)
    float acc_scale;      // v = acc * acc_scale + bias  (1/255 for the first layer fed with raw 0..255 pixels)
    void* out;            // EPI_PRELU: __half activations; EPI_SHUFFLE_*: frames (u8 or float), packed
    const uint8_t* frames_in;  // EPI_SHUFFLE_*: packed u8 input frames (residual branch)
    int32_t frame_h, frame_w;  // input frame size
    int32_t scale;             // pixel-shuffle factor
    int32_t R;                 // shared-memory ring rows
    int32_t desc_mode;         // bring-up: 0 = base_offset 0 (absolute-address swizzle), 1 = base_offset from address
};

#define B2SR_SMEM_LIMIT (227 * 1024)

}  // namespace b2sr
