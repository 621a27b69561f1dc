/*
 * oracle.c -- CPU restatement of the arithmetic on the reference's frame-upscale path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under upscale_video_b200/ may link, import or execute this file;
 * it is used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY UNPINNED: the reference (davlee1972/upscale_video) delegates all arithmetic to the third-party
 * `ncnn_vulkan` wheel (version unpinned, reference README.md:29; call sites
 * upscale/upscale_processing.py:265-281 and :437-453), which is not in /root/reference, not installable
 * offline, and the reference ships no golden outputs or known-answer tests.
 * (What CAN be pinned is: tests/test_architecture_cross_check.py shows that this restatement equals the published PyTorch
 * architectures the model files were exported from -- SRVGGNetCompact, ESRGAN-plus RRDB_Net -- to < 1e-8; and the Python side's
 * restatement of the reference's glue reproduces, bit for bit, files written by the reference's own upscale_processing.py run with
 * an ncnn stand-in -- tools/make_ref_glue_goldens.py, tests/golden/ref_glue.npz.)
 * This file therefore restates ncnn's *published layer definitions* for exactly the layers the reference's
 * model files use (reference models/2x_Compact_Pretrain.param:3-42, 1x_HurrDeblur...param:3-26,
 * 4x_Valar_v1.param:3-1208):
 *
 *   Convolution  (param 0=out-ch 1=kernel 4=pad 5=bias 9=activation; 9=2 -> LeakyReLU, slope in -23310)
 *                stride 1, dilation 1, zero padding, weights OIHW           -> oracle_conv_{f32,f64}
 *   PReLU        per-channel slope: y = x < 0 ? x * slope[c] : x            -> fused into conv (act=3) or oracle_prelu
 *   PixelShuffle mode 0: out[c][y*r+dy][x*r+dx] = in[c*r*r + dy*r + dx][y][x]-> oracle_pixelshuffle
 *   Interp       resize_type 1 (nearest): in_y = min((int)(y * (1/scale)), h-1) -> oracle_nearest
 *   BinaryOp add / Eltwise sum with coefficients / Concat(axis=channel)     -> numpy in oracle.py
 *
 * Two precisions of the same code are built: *_f32 (what ncnn's CPU path computes in) and *_f64 (the
 * reference value the goldens are frozen from).  Activations are HWC (channel-last) arrays of the same
 * type as the accumulator.
 *
 * Build: see oracle/Makefile  (gcc -O3 -fopenmp -shared -fPIC).
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define ACT_NONE 0
#define ACT_LEAKY 2 /* ncnn activation_type 2: slope = activation_params[0] */
#define ACT_PRELU 3 /* not an ncnn conv activation: conv followed by a PReLU layer, fused here */

#define CB 16 /* output-channel block */
#define PX 6  /* output pixels per register tile: PX*CB accumulators stay in vector registers */

#define DEFINE_CONV(NAME, REAL)                                                                        \
    /* in: [H][W][Cin], w: [Cout][Cin][k][k] (ncnn order), bias: [Cout] or NULL, out: [H][W][Cout] */  \
    int NAME(const REAL *in, int H, int W, int Cin, const float *w_oihw, const float *bias, int Cout,  \
             int k, int pad, int act, const float *slope, REAL *out)                                   \
    {                                                                                                  \
        if ((k != 1 && k != 3) || pad != k / 2) return -1;                                             \
        /* repack weights to [ky][kx][cin][cout_padded] so the inner loop is unit-stride over cout */  \
        int CoutP = (Cout + CB - 1) / CB * CB;                                                         \
        size_t nw = (size_t)CoutP * Cin * k * k;                                                       \
        REAL *wr = (REAL *)calloc(nw, sizeof(REAL));                                                   \
        /* one zero pixel so out-of-image taps need no branch in the inner loop */                     \
        REAL *zero = (REAL *)calloc((size_t)Cin, sizeof(REAL));                                        \
        if (!wr || !zero) { free(wr); free(zero); return -2; }                                         \
        for (int o = 0; o < Cout; ++o)                                                                 \
            for (int i = 0; i < Cin; ++i)                                                              \
                for (int t = 0; t < k * k; ++t)                                                        \
                    wr[((size_t)t * Cin + i) * CoutP + o] = (REAL)w_oihw[((size_t)o * Cin + i) * k * k + t]; \
        _Pragma("omp parallel for schedule(dynamic, 2)")                                               \
        for (int y = 0; y < H; ++y) {                                                                  \
            for (int x0 = 0; x0 < W; x0 += PX) {                                                       \
                int npx = W - x0 < PX ? W - x0 : PX;                                                   \
                for (int ob = 0; ob < CoutP; ob += CB) {                                               \
                    REAL acc[PX][CB];                                                                  \
                    for (int p = 0; p < PX; ++p)                                                       \
                        for (int j = 0; j < CB; ++j) acc[p][j] = (REAL)0;                              \
                    for (int ky = 0; ky < k; ++ky) {                                                   \
                        int iy = y + ky - pad;                                                         \
                        if (iy < 0 || iy >= H) continue;                                               \
                        for (int kx = 0; kx < k; ++kx) {                                               \
                            const REAL *ipx[PX];                                                       \
                            for (int p = 0; p < PX; ++p) {                                             \
                                int ix = x0 + p + kx - pad;                                            \
                                ipx[p] = (p < npx && ix >= 0 && ix < W) ? in + ((size_t)iy * W + ix) * Cin : zero; \
                            }                                                                          \
                            const REAL *wt = wr + ((size_t)(ky * k + kx) * Cin) * CoutP + ob;          \
                            for (int i = 0; i < Cin; ++i) {                                            \
                                const REAL *wv = wt + (size_t)i * CoutP;                               \
                                for (int p = 0; p < PX; ++p) {                                         \
                                    REAL a = ipx[p][i];                                                \
                                    _Pragma("omp simd")                                                \
                                    for (int j = 0; j < CB; ++j) acc[p][j] += a * wv[j];               \
                                }                                                                      \
                            }                                                                          \
                        }                                                                              \
                    }                                                                                  \
                    int nb = Cout - ob < CB ? Cout - ob : CB;                                          \
                    for (int p = 0; p < npx; ++p) {                                                    \
                        REAL *o_px = out + ((size_t)y * W + x0 + p) * Cout;                            \
                        for (int j = 0; j < nb; ++j) {                                                 \
                            REAL v = acc[p][j] + (bias ? (REAL)bias[ob + j] : (REAL)0);                \
                            if (act == ACT_LEAKY) v = v > 0 ? v : v * (REAL)slope[0];                  \
                            else if (act == ACT_PRELU) v = v < 0 ? v * (REAL)slope[ob + j] : v;        \
                            o_px[ob + j] = v;                                                          \
                        }                                                                              \
                    }                                                                                  \
                }                                                                                      \
            }                                                                                          \
        }                                                                                              \
        free(wr);                                                                                      \
        free(zero);                                                                                    \
        return 0;                                                                                      \
    }

DEFINE_CONV(oracle_conv_f32, float)
DEFINE_CONV(oracle_conv_f64, double)

#define DEFINE_MISC(SUF, REAL)                                                                         \
    void oracle_prelu_##SUF(REAL *x, size_t npix, int C, const float *slope)                           \
    {                                                                                                  \
        for (size_t p = 0; p < npix; ++p)                                                              \
            for (int c = 0; c < C; ++c) {                                                              \
                REAL v = x[p * C + c];                                                                 \
                x[p * C + c] = v < 0 ? v * (REAL)slope[c] : v;                                         \
            }                                                                                          \
    }                                                                                                  \
    /* in [H][W][C*r*r] -> out [H*r][W*r][C] */                                                        \
    void oracle_pixelshuffle_##SUF(const REAL *in, int H, int W, int C, int r, REAL *out)              \
    {                                                                                                  \
        for (int y = 0; y < H; ++y)                                                                    \
            for (int x = 0; x < W; ++x)                                                                \
                for (int c = 0; c < C; ++c)                                                            \
                    for (int dy = 0; dy < r; ++dy)                                                     \
                        for (int dx = 0; dx < r; ++dx)                                                 \
                            out[(((size_t)(y * r + dy)) * (W * r) + (x * r + dx)) * C + c] =           \
                                in[((size_t)y * W + x) * (C * r * r) + c * r * r + dy * r + dx];       \
    }                                                                                                  \
    /* nearest resize by (sy, sx): in [H][W][C] -> out [OH][OW][C], ncnn Interp resize_type=1 */        \
    void oracle_nearest_##SUF(const REAL *in, int H, int W, int C, float sy, float sx, int OH, int OW, \
                              REAL *out)                                                               \
    {                                                                                                  \
        float hs = 1.f / sy, ws = 1.f / sx; /* ncnn: hs = output_height ? h/outh : 1.f/height_scale */ \
        for (int y = 0; y < OH; ++y) {                                                                 \
            int iy = (int)(y * hs);                                                                    \
            if (iy > H - 1) iy = H - 1;                                                                \
            for (int x = 0; x < OW; ++x) {                                                             \
                int ix = (int)(x * ws);                                                                \
                if (ix > W - 1) ix = W - 1;                                                            \
                memcpy(out + ((size_t)y * OW + x) * C, in + ((size_t)iy * W + ix) * C, C * sizeof(REAL)); \
            }                                                                                          \
        }                                                                                              \
    }

DEFINE_MISC(f32, float)
DEFINE_MISC(f64, double)

/* cv2.imwrite on a float/double image == saturate_cast<uchar>: round-half-to-even, then clamp to [0,255]
 * (reference upscale_processing.py:288 and :519 hand float arrays to cv2.imwrite). */
void oracle_saturate_u8_f64(const double *in, size_t n, uint8_t *out)
{
    for (size_t i = 0; i < n; ++i) {
        double r = nearbyint(in[i]); /* default rounding mode: to nearest, ties to even */
        out[i] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
    }
}

int oracle_abi_version(void) { return 1; }
