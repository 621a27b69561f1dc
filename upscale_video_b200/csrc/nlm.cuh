// nlm.cuh -- the `-m n=<level>` denoise pass of the reference on sm_100a.
//
// reference upscale/upscale_processing.py:350-362 (`apply_denoise`):
//     cv2.fastNlMeansDenoisingColored(img, None, denoise, denoise, 5, 9)
// i.e. BGR -> Lab (linear, 8-bit fixed point), non-local means on the L plane (h) and on the (a, b) plane pair
// (hColor) with a 5x5 template and a 9x9 search window over a 6-px BORDER_REFLECT_101 frame, Lab -> BGR.
// Everything is integer arithmetic (OpenCV's fixed-point weight table, unsigned 32-bit accumulators), so the
// kernel is bit-exact against cv2 (oracle/nlmeans.py states the algorithm and is pinned against cv2 itself).
//
// Mapping.  One warp owns a tile of 28 columns x 16 rows and never talks to another warp (no __syncthreads):
//   * the warp stages the Lab values (packed L | a<<8 | b<<16) of its tile plus the 6-px frame in shared memory,
//     converting from BGR on the way in and resolving the reflect-101 border there;
//   * lane l stands on image column x0 + l - 2 (lanes 2..29 produce output, lanes 0,1,30,31 are template halo);
//   * for each of the 81 search offsets the warp walks its 20 template rows top to bottom: per row one packed
//     load of the pixel and one of its shifted partner, squared differences with __vabsdiffu4 + dp4a, the
//     vertical 5-sum as a running sum per column in registers and, for the 16 rows that complete a template, the
//     horizontal 5-sum with four warp shuffles per plane (four for both planes together at the usual small levels,
//     see PACKED below) -- the 25-tap template distance costs 2 shared-memory loads and 4 or 8 shuffles per pixel
//     and offset;
//   * weights come from the (truncated) fixed-point table in global memory (L1-resident: a few hundred bytes at
//     the reference's typical level 3); 5 accumulators per output row live in registers (80 per lane).
// Bound: integer ALU / shuffle issue, not HBM (6 B of algorithmic traffic per pixel against ~4 k instructions).
#pragma once
#include <stdint.h>

namespace b2sr {

constexpr int NLM_TW = 28;           // output columns per warp: 32 lanes minus the 2 + 2 template halo lanes
constexpr int NLM_TH = 16;           // output rows per warp (default; the kernel takes it as a template parameter)
constexpr int NLM_BORDER = 6;        // search 9/2 + template 5/2
constexpr int NLM_SW = NLM_TW + 2 * NLM_BORDER;  // staged columns (40)
constexpr int NLM_WARPS = 4;
constexpr int NLM_BIN_SHIFT = 5;     // template distance -> table index: sum >> 5 (32 = next power of two of 25)

struct NlmParams {
    const uint8_t* in;    // n frames, 3 interleaved u8 channels
    uint8_t* out;
    long long in_frame_stride, out_frame_stride;  // bytes between frames
    int32_t in_stride, out_stride;                // bytes between rows
    int32_t H, W;
    int32_t tiles_x, tiles_per_frame;
    long long n_tiles;
    int32_t fwd[9];             // linear RGB -> XYZ / white point, 12 fractional bits, rows X Y Z, columns R G B
    int32_t inv[9];             // XYZ -> linear RGB * white point, rows R G B
    const uint16_t* cbrt_tab;   // [3072] cube-root table, 15 fractional bits
    const int32_t* l2y;         // [256] L -> y   (14 fractional bits)
    const int32_t* l2fy;        // [256] L -> f(y)
    const int32_t* w_l;         // weight table of the L plane, first n_l entries (the rest are zero)
    const int32_t* w_ab;        // weight table of the (a, b) plane pair
    uint32_t n_l, n_ab;
};

__device__ __forceinline__ int nlm_reflect101(int p, int len) {
    if (len == 1) return 0;
    while ((unsigned)p >= (unsigned)len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}

__device__ __forceinline__ int nlm_clamp255(int v) { return min(max(v, 0), 255); }

// cvtColor COLOR_LBGR2Lab, 8-bit
__device__ __forceinline__ uint32_t nlm_bgr2lab(const NlmParams& P, int b, int g, int r) {
    b <<= 3;
    g <<= 3;
    r <<= 3;
    const int fx = __ldg(P.cbrt_tab + ((r * P.fwd[0] + g * P.fwd[1] + b * P.fwd[2] + 2048) >> 12));
    const int fy = __ldg(P.cbrt_tab + ((r * P.fwd[3] + g * P.fwd[4] + b * P.fwd[5] + 2048) >> 12));
    const int fz = __ldg(P.cbrt_tab + ((r * P.fwd[6] + g * P.fwd[7] + b * P.fwd[8] + 2048) >> 12));
    const int lscale = (116 * 255 + 50) / 100;
    const int lshift = -((16 * 255 * (1 << 15) + 50) / 100);
    const int L = nlm_clamp255((lscale * fy + lshift + (1 << 14)) >> 15);
    const int A = nlm_clamp255((500 * (fx - fy) + 128 * (1 << 15) + (1 << 14)) >> 15);
    const int B = nlm_clamp255((200 * (fy - fz) + 128 * (1 << 15) + (1 << 14)) >> 15);
    return (uint32_t)L | ((uint32_t)A << 8) | ((uint32_t)B << 16);
}

// f(t) -> t for the X and Z axes (OpenCV's abToXZ table, computed instead of stored: 14 fractional bits)
__device__ __forceinline__ int nlm_f2xz(int t) {
    const int base = 1 << 14;
    return t <= 3390 ? (t * 108) / 841 - ((base * 16 / 116) * 108) / 841 : ((t * t) / base * t) / base;
}

// cvtColor COLOR_Lab2LBGR, 8-bit; returns b | g<<8 | r<<16
__device__ __forceinline__ uint32_t nlm_lab2bgr(const NlmParams& P, int L, int A, int B) {
    const int base = 1 << 14;
    const int y = __ldg(P.l2y + L), ify = __ldg(P.l2fy + L);
    const int adiv = ((5 * A * 53687 + (1 << 7)) >> 13) - 128 * base / 500;
    const int bdiv = ((B * 41943 + (1 << 4)) >> 9) - 128 * base / 200 + 1;
    const int x = nlm_f2xz(ify + adiv), z = nlm_f2xz(ify - bdiv);
    uint32_t px = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {  // c = 0: B (matrix row 2), 1: G, 2: R
        const int row = 2 - c;
        int v = (P.inv[row * 3] * x + P.inv[row * 3 + 1] * y + P.inv[row * 3 + 2] * z + (1 << 13)) >> 14;
        v = min(max(v, 0), 4095);
        px |= (uint32_t)((v * 255) >> 12) << (8 * c);
    }
    return px;
}

// PACKED (levels whose weight tables are at most NLM_PACK_MAX_TABLE entries long, i.e. h <= 6 for the colour planes):
// both column sums are clamped to NLM_PACK_CLAMP and travel through the warp shuffles as two 16-bit fields of
// one word -- 4 shuffles per row instead of 8.  Exact: a clamped term alone pushes the template sum to >= NLM_PACK_CLAMP,
// whose table index (>> 5) is >= NLM_PACK_MAX_TABLE, where the true sum's weight is zero as well; 5 * NLM_PACK_CLAMP
// still fits 16 bits, so the horizontal sums never carry into the neighbouring field.
constexpr uint32_t NLM_PACK_CLAMP = 13107;                               // 5 * 13107 = 65535
constexpr uint32_t NLM_PACK_MAX_TABLE = NLM_PACK_CLAMP >> NLM_BIN_SHIFT;  // 409

template <bool PACKED, int TH>
__global__ void __launch_bounds__(NLM_WARPS * 32) nlm_kernel(const NlmParams P) {
    constexpr int NLM_SH = TH + 2 * NLM_BORDER;  // staged rows (28 for TH = 16)
    __shared__ uint32_t s_lab[NLM_WARPS][NLM_SH][NLM_SW];
    // PACKED: both weight tables (<= 409 entries each, zero-padded to 410) live in shared memory and are indexed with a
    // clamped index -- min + LDS instead of compare + pointer arithmetic + predicated global load + select
    __shared__ uint32_t s_w[PACKED ? 2 : 1][PACKED ? NLM_PACK_MAX_TABLE + 1 : 1];
    if (PACKED) {
        for (uint32_t i = threadIdx.x; i <= NLM_PACK_MAX_TABLE; i += NLM_WARPS * 32) {
            s_w[0][i] = i < P.n_l ? (uint32_t)__ldg(P.w_l + i) : 0u;
            s_w[1][i] = i < P.n_ab ? (uint32_t)__ldg(P.w_ab + i) : 0u;
        }
        __syncthreads();
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t = (long long)blockIdx.x * NLM_WARPS + warp;
    if (t >= P.n_tiles) return;  // warps are independent: no block-wide barrier below
    const int frame = (int)(t / P.tiles_per_frame);
    const int rem = (int)(t % P.tiles_per_frame);
    const int x0 = (rem % P.tiles_x) * NLM_TW, y0 = (rem / P.tiles_x) * TH;
    const uint8_t* src = P.in + frame * P.in_frame_stride;

    // stage: BGR -> Lab, reflect-101 at the image border
    for (int i = lane; i < NLM_SH * NLM_SW; i += 32) {
        const int sy = i / NLM_SW, sx = i - sy * NLM_SW;
        const int gy = nlm_reflect101(y0 + sy - NLM_BORDER, P.H), gx = nlm_reflect101(x0 + sx - NLM_BORDER, P.W);
        const uint8_t* p = src + (long long)gy * P.in_stride + gx * 3;
        s_lab[warp][sy][sx] = nlm_bgr2lab(P, __ldg(p), __ldg(p + 1), __ldg(p + 2));
    }
    __syncwarp();

    uint32_t est_l[TH], est_a[TH], est_b[TH], ws_l[TH], ws_ab[TH];
#pragma unroll
    for (int o = 0; o < TH; ++o) est_l[o] = est_a[o] = est_b[o] = ws_l[o] = ws_ab[o] = 0;

    const int cx = lane + NLM_BORDER - 2;  // staged column of this lane's pixel (image column x0 + lane - 2)
    const uint32_t* own = &s_lab[warp][NLM_BORDER - 2][cx];  // first template row of output row 0
#pragma unroll 1
    for (int dy = -4; dy <= 4; ++dy) {
#pragma unroll 1
        for (int dx = -4; dx <= 4; ++dx) {
            const uint32_t* other = own + dy * NLM_SW + dx;
            uint32_t v_l = 0, v_c = 0, q1 = 0, q2 = 0;
            uint32_t dl_hist[TH + 4], dc_hist[TH + 4];  // compile-time indexed: five of each are live
#pragma unroll
            for (int r = 0; r < TH + 4; ++r) {
                const uint32_t p = own[r * NLM_SW], q = other[r * NLM_SW];
                const uint32_t ad = __vabsdiffu4(p, q);
                const uint32_t l = ad & 0xffu;
                const uint32_t d_l = l * l;                       // dL^2
                const uint32_t d_c = __dp4a(ad, ad, 0u) - d_l;    // da^2 + db^2 (byte 3 is zero)
                // vertical 5-sum first, as a running sum per column ...
                dl_hist[r] = d_l, dc_hist[r] = d_c;
                v_l += d_l, v_c += d_c;
                if (r >= 5) v_l -= dl_hist[r - 5], v_c -= dc_hist[r - 5];
                if (r >= 4) {
                    // ... then, once rows r-4..r are in, the horizontal 5-sum across lanes: output row o = r - 4, whose
                    // partner pixel was loaded at r - 2
                    const int o = r - 4;
                    uint32_t k_l, k_c, w_l, w_c;
                    if (PACKED) {
                        const uint32_t v = min(v_l, NLM_PACK_CLAMP) | (min(v_c, NLM_PACK_CLAMP) << 16);
                        const uint32_t s = v + __shfl_up_sync(0xffffffffu, v, 1) + __shfl_up_sync(0xffffffffu, v, 2) +
                                           __shfl_down_sync(0xffffffffu, v, 1) + __shfl_down_sync(0xffffffffu, v, 2);
                        k_l = (s & 0xffffu) >> NLM_BIN_SHIFT;
                        k_c = s >> (16 + NLM_BIN_SHIFT);
                        w_l = s_w[0][min(k_l, NLM_PACK_MAX_TABLE)];
                        w_c = s_w[1][min(k_c, NLM_PACK_MAX_TABLE)];
                    } else {
                        k_l = (v_l + __shfl_up_sync(0xffffffffu, v_l, 1) + __shfl_up_sync(0xffffffffu, v_l, 2) +
                               __shfl_down_sync(0xffffffffu, v_l, 1) + __shfl_down_sync(0xffffffffu, v_l, 2)) >> NLM_BIN_SHIFT;
                        k_c = (v_c + __shfl_up_sync(0xffffffffu, v_c, 1) + __shfl_up_sync(0xffffffffu, v_c, 2) +
                               __shfl_down_sync(0xffffffffu, v_c, 1) + __shfl_down_sync(0xffffffffu, v_c, 2)) >> NLM_BIN_SHIFT;
                        w_l = k_l < P.n_l ? (uint32_t)__ldg(P.w_l + k_l) : 0u;
                        w_c = k_c < P.n_ab ? (uint32_t)__ldg(P.w_ab + k_c) : 0u;
                    }
                    est_l[o] += w_l * (q2 & 0xffu);
                    est_a[o] += w_c * ((q2 >> 8) & 0xffu);
                    est_b[o] += w_c * ((q2 >> 16) & 0xffu);
                    ws_l[o] += w_l;
                    ws_ab[o] += w_c;
                }
                q2 = q1, q1 = q;
            }
        }
    }

    const int x = x0 + lane - 2;
    if (lane >= 2 && lane < 2 + NLM_TW && x < P.W) {
        uint8_t* dst = P.out + frame * P.out_frame_stride + x * 3;
#pragma unroll
        for (int o = 0; o < TH; ++o) {
            const int y = y0 + o;
            if (y < P.H) {
                const uint32_t L = (est_l[o] + ws_l[o] / 2) / ws_l[o];
                const uint32_t A = (est_a[o] + ws_ab[o] / 2) / ws_ab[o];
                const uint32_t B = (est_b[o] + ws_ab[o] / 2) / ws_ab[o];
                const uint32_t px = nlm_lab2bgr(P, (int)L, (int)A, (int)B);
                uint8_t* d = dst + (long long)y * P.out_stride;
                d[0] = (uint8_t)px;
                d[1] = (uint8_t)(px >> 8);
                d[2] = (uint8_t)(px >> 16);
            }
        }
    }
}

}  // namespace b2sr
