// graph_exec.cuh -- generic CUDA-core executor for ncnn graphs the tcgen05 Compact kernels do not cover.
//
// Today that is `4x_Valar_v1` (reference models/4x_Valar_v1.param:1-1208): a 23-block RRDB network -- 420 convolutions
// with 64..192 input channels fed by Concat, LeakyReLU fused into the convolution (ncnn activation_type 2), 1x1
// shortcut convolutions, Eltwise sums with coefficients (0.2, 1.0), two nearest x2 Interp layers, no pixel shuffle.
// Every layer type the reference's four models use is implemented here (so the Compact models can be run through
// this path too, as one more on-device cross-check): Convolution k=1/3 (+bias, +LeakyReLU), PReLU, PixelShuffle,
// Interp nearest, BinaryOp add / Eltwise sum with coefficients, Concat.  Split is resolved to aliases on the host.
//
// Activations are fp32 NHWC, accumulation is fp32: this is a correctness-first GPU path (plain FMA, no tensor
// cores); a tcgen05 RRDB kernel family (concat-K accumulation into one TMEM tile) is future work.
#pragma once
#include <mma.h>

#include "common.cuh"

namespace b2sr {

enum { GOP_CONV = 1, GOP_PRELU = 2, GOP_PIXELSHUFFLE = 3, GOP_NEAREST = 4, GOP_ADD = 5, GOP_CONCAT = 6 };

// u8 frame rectangle -> float plane, `x * float32(1/255.0)` (reference upscale_processing.py:437-445)
__global__ void g_input_kernel(const uint8_t* __restrict__ frames, int fh, int fw, PlaneDev P, float* __restrict__ out) {
    const int n = P.Ht * P.Wt * 3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = i % 3, p = i / 3, y = p / P.Wt, x = p - y * P.Wt;
        out[i] = (float)frames[((size_t)((size_t)P.frame * fh + P.fy0 + y) * fw + P.fx0 + x) * 3 + c] * (1.f / 255.f);
    }
}

// in [H][W][cin], w [K*K][cin][coutp] (coutp = cout rounded up to 4), out [H][W][cout]; zero padding K/2.
// One thread = one pixel x 4 output channels.
template <int K>
__global__ void g_conv_kernel(const float* __restrict__ in, int ldin, int H, int W, int cin, const float* __restrict__ w,
                              const float* __restrict__ bias, int cout, int coutp, int act, float slope, float* __restrict__ out,
                              int ldout) {
    const int Q = coutp / 4;
    const long long total = (long long)H * W * Q;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(idx % Q);
        const long long p = idx / Q;
        const int y = (int)(p / W), x = (int)(p - (long long)y * W);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
            const int iy = y + ky - K / 2;
            if (iy < 0 || iy >= H) continue;
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
                const int ix = x + kx - K / 2;
                if (ix < 0 || ix >= W) continue;
                const float* ip = in + ((size_t)iy * W + ix) * ldin;
                const float* wp = w + (size_t)((ky * K + kx) * cin) * coutp + q * 4;
                for (int c = 0; c < cin; ++c) {
                    const float a = ip[c];
                    const float4 wv = *reinterpret_cast<const float4*>(wp + (size_t)c * coutp);
                    acc.x = fmaf(a, wv.x, acc.x);
                    acc.y = fmaf(a, wv.y, acc.y);
                    acc.z = fmaf(a, wv.z, acc.z);
                    acc.w = fmaf(a, wv.w, acc.w);
                }
            }
        }
        float v[4] = {acc.x, acc.y, acc.z, acc.w};
        float* op = out + (size_t)p * ldout;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = q * 4 + j;
            if (o < cout) {
                float t = v[j] + (bias ? bias[o] : 0.f);
                if (act == 2) t = t > 0.f ? t : t * slope;  // ncnn activation_type 2: LeakyReLU
                op[o] = t;
            }
        }
    }
}

// Tensor-core version of g_conv_kernel for Cin % 16 == 0 and Cout % 16 == 0 (every RRDB convolution of Valar):
// legacy warp-level MMA (wmma m16n16k16, fp16 operands, fp32 accumulate -- HMMA, not tcgen05), activations rounded to
// fp16 when the halo'd input tile is staged in shared memory.  CTA = 8 rows x 32 columns of output pixels x all Cout;
// warp w owns row w of the tile (two 16-pixel segments x Cout/16 accumulator fragments).
constexpr int GW_TY = 8, GW_TX = 32, GW_THREADS = 256, GW_PAD = 16;  // GW_PAD halfs of padding per pixel: conflict-free fragment rows
template <int K, int NF /*Cout / 16*/>
__global__ void __launch_bounds__(GW_THREADS) g_conv_wmma_kernel(const float* __restrict__ in, int ldin, int H, int W, int cin,
                                                                  const __half* __restrict__ w /*[K*K][cin][NF*16]*/,
                                                                  const float* __restrict__ bias, int act, float slope,
                                                                  float* __restrict__ out, int ldout) {
    using namespace nvcuda;
    extern __shared__ __align__(128) uint8_t gsm[];
    constexpr int HALO = K / 2, SY = GW_TY + 2 * HALO, SX = GW_TX + 2 * HALO, COUT = NF * 16;
    const int ps = cin + GW_PAD;  // pixel stride in halfs (multiple of 16 -> 32-byte aligned fragment pointers)
    __half* tile = reinterpret_cast<__half*>(gsm);
    const int tiles_x = (W + GW_TX - 1) / GW_TX;
    const int ty0 = (blockIdx.x / tiles_x) * GW_TY, tx0 = (blockIdx.x % tiles_x) * GW_TX;
    // stage the input tile (zero outside the image = the convolution's zero padding), fp32 -> fp16
    const int c4 = cin / 4;
    for (int i = threadIdx.x; i < SY * SX * c4; i += GW_THREADS) {
        const int c = (i % c4) * 4, p = i / c4, sx = p % SX, sy = p / SX;
        const int y = ty0 + sy - HALO, x = tx0 + sx - HALO;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (y >= 0 && y < H && x >= 0 && x < W) v = *reinterpret_cast<const float4*>(in + ((size_t)y * W + x) * ldin + c);
        __half2* d = reinterpret_cast<__half2*>(tile + (size_t)p * ps + c);
        d[0] = __floats2half2_rn(v.x, v.y);
        d[1] = __floats2half2_rn(v.z, v.w);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[2][NF];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int n = 0; n < NF; ++n) wmma::fill_fragment(acc[s][n], 0.f);
    for (int ky = 0; ky < K; ++ky)
        for (int kx = 0; kx < K; ++kx) {
            const __half* wt = w + (size_t)((ky * K + kx) * cin) * COUT;
            const __half* a0 = tile + (size_t)((warp + ky) * SX + kx) * ps;
            for (int c = 0; c < cin; c += 16) {
                wmma::fragment<wmma::matrix_a, 16, 16, 16, __half, wmma::row_major> fa[2];
                wmma::load_matrix_sync(fa[0], a0 + c, ps);
                wmma::load_matrix_sync(fa[1], a0 + (size_t)16 * ps + c, ps);
#pragma unroll
                for (int n = 0; n < NF; ++n) {
                    wmma::fragment<wmma::matrix_b, 16, 16, 16, __half, wmma::row_major> fb;
                    wmma::load_matrix_sync(fb, wt + (size_t)c * COUT + n * 16, COUT);
                    wmma::mma_sync(acc[0][n], fa[0], fb, acc[0][n]);
                    wmma::mma_sync(acc[1][n], fa[1], fb, acc[1][n]);
                }
            }
        }
    __syncthreads();  // everyone is done reading the input tile: reuse it as fp32 output staging [TY][TX][COUT]
    float* stg = reinterpret_cast<float*>(gsm);
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int n = 0; n < NF; ++n)
            wmma::store_matrix_sync(stg + (size_t)(warp * GW_TX + s * 16) * COUT + n * 16, acc[s][n], COUT, wmma::mem_row_major);
    __syncthreads();
    for (int i = threadIdx.x; i < GW_TY * GW_TX * (COUT / 4); i += GW_THREADS) {
        const int c = (i % (COUT / 4)) * 4, p = i / (COUT / 4), x = tx0 + p % GW_TX, y = ty0 + p / GW_TX;
        if (y >= H || x >= W) continue;
        float4 v = *reinterpret_cast<const float4*>(stg + (size_t)p * COUT + c);
        float t[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            t[j] += bias ? bias[c + j] : 0.f;
            if (act == 2) t[j] = t[j] > 0.f ? t[j] : t[j] * slope;
        }
        *reinterpret_cast<float4*>(out + ((size_t)y * W + x) * ldout + c) = make_float4(t[0], t[1], t[2], t[3]);
    }
}

__global__ void g_prelu_kernel(const float* __restrict__ in, int ldin, size_t n, int C, const float* __restrict__ slope,
                               float* __restrict__ out, int ldout) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t p = i / C;
        const int c = (int)(i - p * C);
        const float v = in[p * ldin + c];
        out[p * ldout + c] = v < 0.f ? v * slope[c] : v;
    }
}

// in [H][W][C*r*r] -> out [H*r][W*r][C], ncnn PixelShuffle mode 0
__global__ void g_pixelshuffle_kernel(const float* __restrict__ in, int ldin, int H, int W, int C, int r, float* __restrict__ out,
                                      int ldout) {
    const size_t n = (size_t)H * r * W * r * C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const size_t p = i / C;
        const int ox = (int)(p % ((size_t)W * r)), oy = (int)(p / ((size_t)W * r));
        const int x = ox / r, dx = ox - x * r, y = oy / r, dy = oy - y * r;
        out[p * ldout + c] = in[((size_t)y * W + x) * ldin + (size_t)c * r * r + dy * r + dx];
    }
}

// nearest resize by integer factor r: in_y = min(int(y * (1/r)), H-1) (ncnn Interp resize_type 1)
__global__ void g_nearest_kernel(const float* __restrict__ in, int ldin, int H, int W, int C, int r, float* __restrict__ out, int ldout) {
    const size_t n = (size_t)H * r * W * r * C;
    const float inv = 1.f / (float)r;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const size_t p = i / C;
        const int ox = (int)(p % ((size_t)W * r)), oy = (int)(p / ((size_t)W * r));
        const int x = min((int)(ox * inv), W - 1), y = min((int)(oy * inv), H - 1);
        out[p * ldout + c] = in[((size_t)y * W + x) * ldin + c];
    }
}

// out = a * ca + b * cb  (BinaryOp add: ca = cb = 1; Eltwise SUM with coefficients)
__global__ void g_add_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb, size_t n, int C, float ca,
                             float cb, int plain, float* __restrict__ out, int ldout) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t p = i / C;
        const int c = (int)(i - p * C);
        const float x = a[p * lda + c], y = b[p * ldb + c];
        out[p * ldout + c] = plain ? x + y : x * ca + y * cb;
    }
}

// copy input [px][C] into channels [off, off+C) of out [px][Ctot]
__global__ void g_concat_kernel(const float* __restrict__ in, int ldin, size_t px, int C, int ldout, int off, float* __restrict__ out) {
    const size_t n = px * C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t p = i / C;
        out[p * ldout + off + (i - p * C)] = in[p * ldin + (i - p * C)];
    }
}

// network output plane [Ht*S][Wt*S][3] -> `* 255`, crop to the tile's core, store as f32 or as cv2.imwrite would (u8)
template <bool F32OUT>
__global__ void g_output_kernel(const float* __restrict__ net, int ldnet, PlaneDev P, int S, int fh, int fw, void* __restrict__ out) {
    const int ch = (P.cy1 - P.cy0) * S, cw = (P.cx1 - P.cx0) * S;
    const int n = ch * cw * 3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = i % 3, p = i / 3, y = p / cw, x = p - y * cw;
        const int py = (P.cy0 - P.fy0) * S + y, px = (P.cx0 - P.fx0) * S + x;  // inside the plane's output
        const float v = net[((size_t)py * (P.Wt * S) + px) * ldnet + c] * 255.f;
        const size_t o = (((size_t)P.frame * fh * S + (size_t)P.cy0 * S + y) * ((size_t)fw * S) + (size_t)P.cx0 * S + x) * 3 + c;
        if (F32OUT) {
            reinterpret_cast<float*>(out)[o] = v;
        } else {
            const int iv = __float2int_rn(v);
            reinterpret_cast<uint8_t*>(out)[o] = (uint8_t)min(max(iv, 0), 255);
        }
    }
}

}  // namespace b2sr
