/*
 * b2sr.h -- C ABI of the B200 frame super-resolution engine (libb2sr.so).
 *
 * This is the drop-in boundary for the reference's per-frame worker loop.  The reference
 * (davlee1972/upscale_video) has no FFI of its own: its worker functions call the third-party ncnn_vulkan
 * object API.  Each entry point below replaces one group of those calls; file:line citations are into
 * /root/reference/.
 *
 *   reference call site                                              replaced by
 *   ---------------------------------------------------------------  ---------------------------------
 *   ncnn.get_gpu_count / get_default_gpu_index / get_gpu_info         b2sr_device_count / b2sr_device_name
 *     (test_gpus.py:47-67)
 *   ncnn.Net(); opt.use_vulkan_compute; set_vulkan_device;            b2sr_create
 *     load_param; load_model   (upscale/upscale_processing.py:65-71)
 *   Mat.from_pixels + substract_mean_normalize + create_extractor +   b2sr_run_u8  (whole upscale_image
 *     input + extract + np.array + *255 + canvas scatter + imwrite      tile loop :499-519, or apply_model
 *     rounding   (upscale_processing.py:437-477, :487-519, :263-288)    :263-288 with tile = 0)
 *   the same, up to `output_tile * 255` (float, unrounded)            b2sr_run_f32 (process_tile :437-462)
 *     (upscale_processing.py:437-462)
 *   pool of per-frame tasks on device-resident frames (no reference   b2sr_run_batch_device
 *     equivalent: the reference moves pixels through PNG files)
 *   ncnn.destroy_gpu_instance (upscale_processing.py:292, :458)       b2sr_destroy
 *   exceptions caught at :289 / :454                                   negative return + b2sr_last_error
 *
 * Conventions: plain C types only; every function returning int returns 0 on success and a negative
 * B2SR_E_* code on failure, with a thread-local message available from b2sr_last_error().  A context is
 * bound to one CUDA device, owns its stream, weights, scratch and TMA descriptors, and is not thread-safe
 * (one context per worker process, like the reference's process-global `net`, upscale_processing.py:22,57).
 * There is no CPU fallback: creating a context without a usable sm_100 device fails.
 *
 * Pixel format everywhere: 8-bit, 3 interleaved channels in the order the caller has them (the reference
 * feeds cv2's BGR unswapped as PIXEL_BGR, :265-270); the network is applied channel-for-channel.
 */
#ifndef B2SR_H
#define B2SR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2SR_ABI_VERSION 1

/* error codes */
#define B2SR_OK 0
#define B2SR_E_INVALID (-1)     /* bad argument / unsupported network description */
#define B2SR_E_CUDA (-2)        /* CUDA runtime or driver error (message has the detail) */
#define B2SR_E_NODEVICE (-3)    /* no CUDA device / device is not sm_100 */
#define B2SR_E_NOMEM (-4)       /* host or device allocation failed */
#define B2SR_E_UNSUPPORTED (-5) /* valid request this build cannot serve */

/* network families */
#define B2SR_FAMILY_COMPACT 1 /* SRVGGNetCompact: conv(cin->nf)+PReLU, n_mid x [conv(nf->nf)+PReLU], conv(nf->cin*scale^2),
                                 PixelShuffle(scale) + nearest-upsample(input)  (models/2x_Compact_Pretrain.param:3-42) */

#define B2SR_FAMILY_GRAPH 2   /* any other graph of the layer types the reference's models use (today: the RRDB 4x_Valar_v1,
                                 models/4x_Valar_v1.param:1-1208), run op by op on CUDA cores: see b2sr_create_graph */

/* where image buffers live */
#define B2SR_MEM_HOST 0
#define B2SR_MEM_DEVICE 1

/* b2sr_set_option keys */
#define B2SR_OPT_IMPL 1        /* 0 = auto (pipelined tcgen05 schedule when layers x bands fits the SM count, else layer by
                                  layer), 1 = plain CUDA-core kernels, 2 = tcgen05 layer by layer, 3 = tcgen05 pipelined (error
                                  if it does not fit) */
#define B2SR_OPT_PROFILE 2     /* 1 = bracket every kernel launch with CUDA events (see b2sr_get_stat) */
#define B2SR_OPT_MAX_BATCH 3   /* frames per internal pass of b2sr_run_batch_device (0 = choose from free memory) */
#define B2SR_OPT_RING_ROWS 4   /* pipelined schedule: rows per inter-layer activation ring (0 = auto: all rings together ~36 MB, L2-resident) */
#define B2SR_OPT_PIPE_DEBUG 5  /* pipelined schedule: print per-layer stall accounting to stderr after every launch (synchronises) */
#define B2SR_OPT_SM_LIMIT 6    /* SMs the persistent (pipelined) grids may count on, 0 = all of the device.  The persistent kernels are
                                  launched cooperatively, so a grid the driver cannot make fully co-resident (MPS clients, a second
                                  context's persistent kernel) is refused and the pass runs layer by layer instead; this option
                                  states the limit up front (e.g. CUDA_MPS_ACTIVE_THREAD_PERCENTAGE) */

#define B2SR_OPT_SEG_PIPE 7    /* fused family: 1 = runs of convolutions as persistent segment launches (an RRDB of 4x_Valar_v1 = one
                                  cooperative launch, dense-block buffers in L2-resident rings: 8.8x less DRAM traffic, bit-identical
                                  results, but measured slower on B200 -- DESIGN.md section 3), 0 (default) = one launch per convolution */
#define B2SR_OPT_PAIR2 8       /* fused family: 1 = dense-block convolutions in the CTA-pair form (2-CTA clusters over band pairs,
                                  tcgen05.mma.cta_group::2, M = 256: one instruction stream per two bands, half of the stacked weights per
                                  CTA; same results for the 32-channel convolutions, <= 1 LSB apart for the 64-channel ones whose
                                  accumulator ring is shorter; measured 33.4 vs 28.3 ms per 540p frame on B200 -- DESIGN.md section 3),
                                  0 (default) = one CTA per band */

#define B2SR_OPT_ABLATE 9      /* fused family, MEASUREMENT ONLY (results are garbage): bit 0 = no MMAs, bit 1 = epilogues drain TMEM only,
                                  bit 2 = no row loads, bit 3 = epilogues without their global loads / stores (energy accounting) */

/* b2sr_get_stat keys */
#define B2SR_STAT_LAUNCHES 1       /* kernels launched by this context since creation / last reset */
#define B2SR_STAT_TC_LAUNCHES 2    /* ... of which tcgen05 convolution kernels */
#define B2SR_STAT_TC_MID_MS 3      /* profile mode: summed device time of nf->nf tcgen05 conv launches, ms */
#define B2SR_STAT_TC_MID_COUNT 4   /* profile mode: number of launches summed in B2SR_STAT_TC_MID_MS */
#define B2SR_STAT_ALL_MS 5         /* profile mode: summed device time of every launch, ms */
#define B2SR_STAT_TC_MID_PIXELS 6  /* profile mode: output pixels (exact, no halo/padding) those launches produced */
#define B2SR_STAT_PIPE_LAUNCHES 7  /* pipelined whole-network kernels launched */
#define B2SR_STAT_PIPE_MS 8        /* profile mode: summed device time of the pipelined launches, ms */
#define B2SR_STAT_HMMA_LAUNCHES 9  /* generic graph engine: convolutions launched on the warp-level MMA (wmma) kernel */
#define B2SR_STAT_PIPE_FALLBACKS 10 /* passes that ran layer by layer because the persistent grid did not fit / was refused */

typedef struct b2sr_ctx b2sr_ctx;

typedef struct b2sr_net_desc {
    int32_t family;    /* B2SR_FAMILY_* */
    int32_t cin;       /* image channels (3) */
    int32_t nf;        /* feature channels (64, 24) */
    int32_t n_mid;     /* number of nf->nf convolutions (16, 8) */
    int32_t scale;     /* pixel-shuffle factor 1, 2 or 4 */
    int32_t reserved[11];
} b2sr_net_desc;

int b2sr_abi_version(void);

/* Device enumeration (test_gpus.py:47-67). */
int b2sr_device_count(void);
int b2sr_default_device(void);
int b2sr_device_name(int device, char *buf, int buflen);

/*
 * Create an engine on `device`.  `weights` is a host blob of fp32 values: for every convolution in graph
 * order  weight[out][in][3][3] | bias[out] | slope[out]  (slope omitted after the last convolution).
 * All weights must be exactly representable in fp16 (true for every Compact model the reference ships);
 * otherwise B2SR_E_UNSUPPORTED.
 */
int b2sr_create(b2sr_ctx **out, int device, const void *weights, size_t nbytes, const b2sr_net_desc *desc);
void b2sr_destroy(b2sr_ctx *ctx);

/*
 * Generic graph engine (B2SR_FAMILY_GRAPH): the ncnn graph is handed over as a flat list of operations over numbered
 * activation slots (Split layers resolved to aliases and slots reused by the caller).  Replaces the same ncnn calls
 * as b2sr_create (upscale/upscale_processing.py:65-71) for models/4x_Valar_v1.param, which is not an SRVGGNetCompact.
 * fp32 activations; convolutions with cin % 16 == 0 and cout 32/64 run on warp-level MMA (fp16 operands, fp32 accumulate), the rest on
 * fp32 CUDA cores (B2SR_OPT_IMPL = 1 forces fp32 everywhere); every b2sr_run_* entry point works on such a context.
 */
#define B2SR_OP_CONV 1          /* Convolution k = 1 or 3, pad k/2, optional bias, act 0 = none / 2 = LeakyReLU(slope) */
#define B2SR_OP_PRELU 2         /* per-channel slopes at w_off */
#define B2SR_OP_PIXELSHUFFLE 3  /* factor r, ncnn mode 0 */
#define B2SR_OP_NEAREST 4       /* Interp resize_type 1, integer factor r */
#define B2SR_OP_ADD 5           /* out = coef[0] * in[0] + coef[1] * in[1]  (BinaryOp add: plain = 1) */
#define B2SR_OP_CONCAT 6        /* channel concatenation of in[0..nin) */
typedef struct b2sr_graph_op {
    int32_t type;
    int32_t nin;
    int32_t in[6]; /* input slots */
    int32_t out;   /* output slot */
    int32_t cin;   /* CONV: input channels */
    int32_t cout;  /* CONV: output channels */
    int32_t k;     /* CONV: kernel size */
    int32_t act;
    float slope;
    float coef[2];
    int32_t plain; /* ADD: 1 = a + b exactly (BinaryOp), 0 = with coefficients (Eltwise) */
    int32_t r;
    int64_t w_off; /* offsets in floats into the weight blob; -1 = absent.  CONV weights are OIHW */
    int64_t b_off;
    /* Shapes and views.  Every operand is `c` channels starting at channel `off` of a slot whose pixels are `ld` floats
     * apart (ld = 0: dense, ld = c).  Views let a Concat of values that are only ever concatenated as prefixes of one
     * another (the RRDB pattern [x], [x,x1], [x,x1,x2] ...) cost nothing: the producers write straight into their
     * channel slice of one wide slot and the "concatenated" tensor is a view of its first channels. */
    int32_t in_c[6], in_off[6], in_ld[6];
    int32_t out_c, out_off, out_ld;
    int32_t in_res, out_res; /* resolution factor (1, 2, 4 ...) of inputs / output relative to the network input */
    int32_t reserved;
} b2sr_graph_op;
int b2sr_create_graph(b2sr_ctx **out, int device, const b2sr_graph_op *ops, int n_ops, int n_slots, int in_slot, int out_slot,
                      int scale, const void *weights, size_t nbytes);

/*
 * Fused tcgen05 graph engine (B2SR_FAMILY_FUSED): the same ncnn calls as b2sr_create_graph, for graphs that lower to
 * "convolution + bias + LeakyReLU + up to two residual terms" operations over channel-last buffers -- the RRDB network
 * models/4x_Valar_v1.param:1-1208, whose Concat / Split / BinaryOp / Eltwise layers all disappear into views and
 * epilogues.  Every convolution runs on the tcgen05 kernel of csrc/tc_gconv.cuh (fp16 operands, fp32 accumulate):
 *   v = act(conv_k(in_buf[:, in_off : in_off + cin]) + bias);   v = v * coef_v[i] + res_i * coef_r[i]  (i < nres)
 *   out16_buf[:, out16_off : +cout] = fp16(v)   and / or   out32_buf[:, out32_off : +cout] = v
 * Buffers are numbered; `channels` is the pixel stride of a buffer (a 192-channel buffer holds a whole dense block,
 * convolutions read prefixes of it and write 32-channel slices), `dtype` 2 = fp16 (read by convolutions and NEAREST),
 * 4 = fp32 (read as residual), `res` the resolution factor relative to the network input.  in_buf = -1 is the input
 * image (cin = 3).  The op with final = 1 (cout = 3) produces the network output: * 255, crop to the tile core,
 * round-half-even + saturate (u8) or unrounded (f32), like b2sr_run_u8 / b2sr_run_f32 of the other families.
 */
#define B2SR_FAMILY_FUSED 3
#define B2SR_FOP_CONV 1    /* k = 1 or 3, pad k/2, cin <= 192 (multiple of 16, or 3 for the input image), cout 3 / 32 / 64 */
#define B2SR_FOP_NEAREST 2 /* out16 (res) = nearest-neighbour x r of in (res / r) */
typedef struct b2sr_fused_buf {
    int32_t channels, dtype, res, reserved;
} b2sr_fused_buf;
typedef struct b2sr_fused_op {
    int32_t type;
    int32_t res; /* resolution factor of the pixels this op writes */
    int32_t in_buf, in_off, cin;
    int32_t k, cout, act;
    float slope;
    int32_t nres;
    int64_t w_off, b_off; /* offsets in floats into the weight blob (OIHW weights, bias); b_off = -1: no bias */
    int32_t res_buf[2], res_off[2];
    float coef_v[2], coef_r[2];
    int32_t out16_buf, out16_off; /* -1 = not stored in this form */
    int32_t out32_buf, out32_off;
    int32_t r;     /* NEAREST: integer factor */
    int32_t final; /* 1 = network output */
    /* Optional fused 1x1 shortcut (the RRDB pattern "x2 = lrelu(conv3x3([x, x1])) + conv1x1(x)", models/4x_Valar_v1.param:9-12):
     * a bias-less 1x1 convolution over the first sc_cin (<= 64) channels of this op's own input view, accumulated beside
     * the 3x3 convolution and combined before the residual terms:  v = act(conv + bias);  v = v * sc_coef_v + s * sc_coef_r.
     * sc_cin = 0: none.  Requires cout = 32, k = 3. */
    int32_t sc_cin;
    float sc_coef_v, sc_coef_r;
    int32_t reserved;
    int64_t sc_w_off; /* [cout][sc_cin] weights in the blob */
} b2sr_fused_op;
int b2sr_create_fused(b2sr_ctx **out, int device, const b2sr_fused_op *ops, int n_ops, const b2sr_fused_buf *bufs, int n_bufs,
                      int scale, const void *weights, size_t nbytes);
/* Host-only planning aid (no device needed): how b2sr_create_fused would cut the program into persistent segments -- runs of
 * consecutive convolutions executed as ONE cooperative launch each, CTA = (stage, band), with every buffer slice that is
 * written and read inside the run kept in an L2-resident row ring -- on a device with `sms` SMs.  out[0] = number of
 * segments; per segment {op_begin, op_end, n_stages, n_ring_instances}, n_stages records of 13 ints {op, half, variant,
 * in_inst, grp_ring[3], out16_inst, out32_inst, res_inst[2], gate_op, bp_op}, n_ring_instances records {buffer,
 * last_reader}.  Returns the number of ints of the description (at most `cap` are written). */
int b2sr_fused_describe_segments(const b2sr_fused_op *ops, int n_ops, const b2sr_fused_buf *bufs, int n_bufs, int sms,
                                 int32_t *out, int cap);
/* Bring-up aid (fused family): run ops [0, upto] on one untiled u8 image and return buffer `buf` (all its channels,
 * converted to float) as (h * res) x (w * res) x channels floats on the host. */
int b2sr_debug_fused(b2sr_ctx *ctx, const uint8_t *in, int h, int w, int upto, int buf, float *out);

/*
 * One frame, u8 in -> u8 out ((h*scale) x (w*scale) x 3, round-half-even + saturate like cv2.imwrite).
 * tile/halo = 960/10 reproduces the reference's tiling (zero padding at tile borders, upscale_processing.py
 * :409-434); tile = 0 runs the frame as one piece (apply_model).  Strides are in bytes.  Synchronous.
 */
int b2sr_run_u8(b2sr_ctx *ctx, const uint8_t *in, int h, int w, int in_stride, uint8_t *out, int out_stride,
                int tile, int halo, int memspace);

/* Same, but the output is the float image the reference holds after `* 255` (unrounded), 3 floats/pixel. */
int b2sr_run_f32(b2sr_ctx *ctx, const uint8_t *in, int h, int w, int in_stride, float *out, int out_stride,
                 int tile, int halo, int memspace);

/*
 * n device-resident frames, packed (no row padding): in  n x h x w x 3, out  n x (h*scale) x (w*scale) x 3.
 * Asynchronous on the context's stream unless `sync` is non-zero.
 */
int b2sr_run_batch_device(b2sr_ctx *ctx, const uint8_t *d_in, uint8_t *d_out, int n, int h, int w, int tile,
                          int halo, int sync);

/* n host frames (pinned or pageable) through a double-buffered H2D -> run -> D2H pipeline; synchronous. */
int b2sr_run_batch_host(b2sr_ctx *ctx, const uint8_t *h_in, uint8_t *h_out, int n, int h, int w, int tile,
                        int halo);

/*
 * The same pipeline without the host waiting for it: the call returns once the copies and launches are enqueued, and
 * `b2sr_wait_batch(ctx, *ticket)` returns once this submission's last output frame is in `h_out`.  Submissions of one context
 * complete in order, and the first H2D copy of one runs under the network of the previous one (equal chunks, no tapered
 * ends), so a caller that keeps two submissions in flight on two pairs of PINNED buffers sees the device-resident rate
 * (pageable memory makes the copies, and therefore the call, synchronous).  `h_in` / `h_out` must stay valid and
 * untouched until the wait returns.  Any other call on the context first waits for all pending submissions.  This is the
 * reference's `pool.apply_async(upscale_image, ...)` + callback (upscale_processing.py:586-598) for a caller that holds
 * raw frames instead of PNG names.
 */
int b2sr_submit_batch_host(b2sr_ctx *ctx, const uint8_t *h_in, uint8_t *h_out, int n, int h, int w, int tile,
                           int halo, uint64_t *ticket);
int b2sr_wait_batch(b2sr_ctx *ctx, uint64_t ticket);

/* Bring-up aid: activations after convolution `layer` (0-based, post-PReLU) of a single untiled u8 image,
 * as float h x w x nf on the host. */
int b2sr_debug_layer(b2sr_ctx *ctx, const uint8_t *in, int h, int w, int layer, float *out);

int b2sr_set_option(b2sr_ctx *ctx, int key, int64_t value);
int b2sr_get_stat(b2sr_ctx *ctx, int key, double *value);
int b2sr_reset_stats(b2sr_ctx *ctx);
int b2sr_synchronize(b2sr_ctx *ctx);
/* The CUDA stream (cudaStream_t) the context launches on, for callers that time with their own events. */
void *b2sr_stream(b2sr_ctx *ctx);

const char *b2sr_last_error(void);

/*
 * Start-up weight broadcast (the reference has every worker read the model files itself,
 * upscale/upscale_processing.py:70-71).  Every rank creates its context from the same network description -- rank `root`
 * with the real parameters, the others with any blob of the right size (zeros pass the fp16-exactness check) -- then
 * b2sr_bcast_weights makes the device-side parameter buffers of all ranks identical with one grouped ncclBroadcast on the
 * context's stream.  `nccl_comm` is an ncclComm_t: the application's own, or one made with the helpers below (rank 0
 * calls b2sr_nccl_unique_id and ships the 128 bytes to the other ranks by any means; everybody calls b2sr_nccl_comm_init).
 * NCCL is resolved at run time (dlopen of the copy already in the process, else libnccl.so.2); B2SR_E_UNSUPPORTED if absent.
 * No collective is used after start-up: frames shard with no data-path communication.
 */
int b2sr_bcast_weights(b2sr_ctx *ctx, void *nccl_comm, int root);
int b2sr_nccl_unique_id(void *id128);
int b2sr_nccl_comm_init(void **comm, int n_ranks, int rank, const void *id128, int device);
int b2sr_nccl_comm_destroy(void *comm);

/*
 * Denoise pass (`-m n=<level>`).  Replaces the one library call of the reference's denoise worker,
 *     cv2.fastNlMeansDenoisingColored(img, None, denoise, denoise, 5, 9)      (upscale/upscale_processing.py:354)
 * -- BGR -> Lab, non-local means on the L plane (h_luma) and on the (a, b) plane pair (h_color) with a 5x5 template
 * and a 9x9 search window over a 6-px reflect-101 border, Lab -> BGR -- in OpenCV's own fixed-point arithmetic, so
 * results are bit-identical to cv2's CPU implementation.  Arguments keep cv2's order and meaning; only the window
 * sizes the reference passes (5, 9) are built, anything else is B2SR_E_UNSUPPORTED.  A context is bound to one device,
 * owns its stream and tables and is not thread-safe.  No CPU path: b2sr_nlm_create fails without an sm_100 device.
 */
typedef struct b2sr_nlm b2sr_nlm;
int b2sr_nlm_create(b2sr_nlm **out, int device);
void b2sr_nlm_destroy(b2sr_nlm *ctx);
/* One frame, 3 interleaved u8 channels in cv2's BGR order; strides in bytes; in == out is not allowed.  Synchronous. */
int b2sr_nlm_run_u8(b2sr_nlm *ctx, const uint8_t *in, int h, int w, int in_stride, uint8_t *out, int out_stride,
                    float h_luma, float h_color, int template_window, int search_window, int memspace);
/* n packed device-resident frames (n x h x w x 3), one launch; asynchronous on the context's stream unless `sync`. */
int b2sr_nlm_run_batch_device(b2sr_nlm *ctx, const uint8_t *d_in, uint8_t *d_out, int n, int h, int w, float h_luma,
                              float h_color, int sync);
/* n packed host frames (pinned or pageable) through a double-buffered H2D -> kernel -> D2H pipeline; synchronous. */
int b2sr_nlm_run_batch_host(b2sr_nlm *ctx, const uint8_t *h_in, uint8_t *h_out, int n, int h, int w, float h_luma,
                            float h_color);
int b2sr_nlm_synchronize(b2sr_nlm *ctx);
void *b2sr_nlm_stream(b2sr_nlm *ctx);    /* cudaStream_t of the context */
double b2sr_nlm_launches(b2sr_nlm *ctx); /* kernels launched by this context */
/* Host-only views of the tables the kernel uses (no device needed; tests compare them with the oracle's):
 * the fixed-point weight table of one plane (channels = 1: L, 2: a/b) -- returns its full length and copies
 * min(length, cap) entries -- and the Lab conversion tables (any pointer may be NULL). */
int b2sr_nlm_weight_table(float h, int channels, int32_t *out, int cap);
int b2sr_nlm_lab_tables(int32_t *fwd9, int32_t *inv9, int32_t *l2y256, int32_t *l2fy256, uint16_t *cbrt3072);

#ifdef __cplusplus
}
#endif
#endif /* B2SR_H */
