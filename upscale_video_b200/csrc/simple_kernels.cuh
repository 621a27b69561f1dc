// simple_kernels.cuh -- the non-tensor-core kernels of libb2sr (sm_100a).
//
//  * prep_kernel: the reference's `Mat.from_pixels(PIXEL_BGR)` + tile slicing (reference
//    upscale/upscale_processing.py:430-442, :265-270) -- u8 HWC frame pixels gathered into the halo'd planes as
//    fp16 NHWC with 16 channels (3 real, 13 zero).  The `* 1/255` of substract_mean_normalize (:443-445) is
//    folded into the first convolution's epilogue (TcParams::acc_scale): 0..255 is exact in fp16.
//  * simple_conv_kernel / simple_shuffle_kernel: a plain CUDA-core implementation of the same layers
//    (B2SR_OPT_IMPL = 1).  It exists for bring-up and as an on-device cross-check of the tcgen05 path (same
//    fp16 storage, fp32 accumulation); it is a GPU code path, not a CPU fallback.
#pragma once
#include "common.cuh"

namespace b2sr {

__global__ void prep_kernel(const uint8_t* __restrict__ frames, int fh, int fw, const PlaneDev* __restrict__ planes,
                            __half* __restrict__ in16) {
    const PlaneDev P = planes[blockIdx.y];
    const int npx = P.Ht * P.Wt;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < npx; idx += gridDim.x * blockDim.x) {
        const int y = idx / P.Wt, x = idx - y * P.Wt;
        const uint8_t* src = frames + ((size_t)((size_t)P.frame * fh + P.fy0 + y) * fw + P.fx0 + x) * 3;
        const __half2 h01 = __floats2half2_rn((float)src[0], (float)src[1]);
        const __half2 h2 = __floats2half2_rn((float)src[2], 0.f);
        uint4 a;
        a.x = *reinterpret_cast<const uint32_t*>(&h01);
        a.y = *reinterpret_cast<const uint32_t*>(&h2);
        a.z = 0u;
        a.w = 0u;
        uint4* dst = reinterpret_cast<uint4*>(in16 + (size_t)(P.pix_off + idx) * 16);
        dst[0] = a;
        dst[1] = make_uint4(0u, 0u, 0u, 0u);
    }
}

// in: [px][CINP] fp16, w: [9][CINP][NOUTP] fp16, out: [px][NOUTP] fp16 (bias + PReLU) or fp32 (bias only, LAST).
// One thread = one pixel x 8 output channels.  Zero padding at the plane border.
template <bool LAST>
__global__ void simple_conv_kernel(const __half* __restrict__ in, int CINP, const __half* __restrict__ w, int NOUTP,
                                   const float* __restrict__ bias, const float* __restrict__ slope, float acc_scale,
                                   const PlaneDev* __restrict__ planes, void* __restrict__ out) {
    const PlaneDev P = planes[blockIdx.y];
    const int G = NOUTP / 8;
    const int total = P.Ht * P.Wt * G;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int g = idx % G, p = idx / G;
        const int y = p / P.Wt, x = p - y * P.Wt;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = y + ky - 1;
            if (iy < 0 || iy >= P.Ht) continue;
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = x + kx - 1;
                if (ix < 0 || ix >= P.Wt) continue;
                const __half* ip = in + (size_t)(P.pix_off + (int64_t)iy * P.Wt + ix) * CINP;
                const __half* wp = w + (size_t)((ky * 3 + kx) * CINP) * NOUTP + g * 8;
                for (int c = 0; c < CINP; ++c) {
                    const float a = __half2float(ip[c]);
                    const uint4 wv = *reinterpret_cast<const uint4*>(wp + (size_t)c * NOUTP);
                    const __half2* wh = reinterpret_cast<const __half2*>(&wv);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 f = __half22float2(wh[j]);
                        acc[2 * j] = fmaf(a, f.x, acc[2 * j]);
                        acc[2 * j + 1] = fmaf(a, f.y, acc[2 * j + 1]);
                    }
                }
            }
        }
        const size_t o = (size_t)(P.pix_off + p) * NOUTP + g * 8;
        if (LAST) {
            float* dst = reinterpret_cast<float*>(out) + o;
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = fmaf(acc[j], acc_scale, bias[g * 8 + j]);
        } else {
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                float v0 = fmaf(acc[j], acc_scale, bias[g * 8 + j]);
                float v1 = fmaf(acc[j + 1], acc_scale, bias[g * 8 + j + 1]);
                v0 = v0 < 0.f ? v0 * slope[g * 8 + j] : v0;
                v1 = v1 < 0.f ? v1 * slope[g * 8 + j + 1] : v1;
                const __half2 h = __floats2half2_rn(v0, v1);
                pk[j >> 1] = *reinterpret_cast<const uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(out) + o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    }
}

// PixelShuffle(S) + nearest-upsampled input + `* 255` (+ cv2.imwrite rounding) for the simple path
// (reference models/2x_Compact_Pretrain.param:40-42, upscale_processing.py:462-477, :519).
template <bool F32OUT>
__global__ void simple_shuffle_kernel(const float* __restrict__ lastf, int NOUTP, const PlaneDev* __restrict__ planes,
                                      const uint8_t* __restrict__ frames, int fh, int fw, int S, void* __restrict__ out) {
    const PlaneDev P = planes[blockIdx.y];
    const int npx = P.Ht * P.Wt;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < npx; idx += gridDim.x * blockDim.x) {
        const int y = idx / P.Wt, x = idx - y * P.Wt;
        const int fy = P.fy0 + y, fx = P.fx0 + x;
        if (fy < P.cy0 || fy >= P.cy1 || fx < P.cx0 || fx >= P.cx1) continue;
        const uint8_t* px = frames + ((size_t)((size_t)P.frame * fh + fy) * fw + fx) * 3;
        const float* a = lastf + (size_t)(P.pix_off + idx) * NOUTP;
        const size_t OW = (size_t)fw * S;
        for (int ch = 0; ch < 3; ++ch) {
            const float xin = (float)px[ch] * (1.f / 255.f);
            for (int dy = 0; dy < S; ++dy)
                for (int dx = 0; dx < S; ++dx) {
                    const float v = (a[ch * S * S + dy * S + dx] + xin) * 255.f;
                    const size_t o = (((size_t)P.frame * fh * S + (size_t)fy * S + dy) * OW + (size_t)fx * S + dx) * 3 + ch;
                    if (F32OUT) {
                        reinterpret_cast<float*>(out)[o] = v;
                    } else {
                        const int iv = __float2int_rn(v);
                        reinterpret_cast<uint8_t*>(out)[o] = (uint8_t)min(max(iv, 0), 255);
                    }
                }
        }
    }
}

// debug: fp16 [px][CP] -> float [px][C] (first C channels)
__global__ void unpack_act_kernel(const __half* __restrict__ act, int CP, int C, size_t npx, float* __restrict__ out) {
    const size_t total = npx * C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t p = i / C;
        const int c = (int)(i - p * C);
        out[i] = __half2float(act[p * CP + c]);
    }
}

}  // namespace b2sr
