// engine.cu -- host side of libb2sr: the C ABI of include/b2sr.h over the sm_100a kernels in tc_conv.cuh /
// simple_kernels.cuh.
//
// What it replaces in the reference (davlee1972/upscale_video): the ncnn_vulkan object surface used by the
// frame-upscale worker loop -- Net/load_param/load_model (upscale/upscale_processing.py:65-71),
// Mat.from_pixels + substract_mean_normalize + Extractor.input/extract (:265-281, :437-453) -- together with the
// per-tile numpy glue of process_tile / upscale_image (:395-519): tiling with a 10-px halo, `* 255`, the crop
// into the canvas and cv2.imwrite's rounding all happen on the device here.
//
// There is no CPU fallback in this file: every entry point that computes needs an sm_100 device.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/b2sr.h"
#include "common.cuh"
#include "graph_exec.cuh"
#include "nlm.cuh"
#include "simple_kernels.cuh"
#include "tc_conv.cuh"
#include "tc_gconv.cuh"

using namespace b2sr;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                                        \
    do {                                                                                                      \
        cudaError_t e_ = (expr);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return fail(e_ == cudaErrorMemoryAllocation ? B2SR_E_NOMEM : B2SR_E_CUDA, "%s failed: %s (%s:%d)", \
                        #expr, cudaGetErrorString(e_), __FILE__, __LINE__);                                   \
    } while (0)

#define TRY(expr)               \
    do {                        \
        int rc_ = (expr);       \
        if (rc_ != 0) return rc_; \
    } while (0)

extern "C" const char* b2sr_last_error(void) { return g_err.c_str(); }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of the (function, device) pair, shared by every engine of the process:
// it is only ever raised, under a lock, so that engines on other threads that launch the same instance with another
// ring depth can never lower it under a launch in flight.
static int raise_dyn_smem(const void* fn, int bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, int> cur;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    int& v = cur[std::make_pair(fn, dev)];
    if (bytes > v) {
        CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        v = bytes;
    }
    return 0;
}

extern "C" int b2sr_abi_version(void) { return B2SR_ABI_VERSION; }

// ------------------------------------------------------------------------------------------------
// device enumeration (reference test_gpus.py:47-67)
// ------------------------------------------------------------------------------------------------
extern "C" int b2sr_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
extern "C" int b2sr_default_device(void) { return b2sr_device_count() > 0 ? 0 : -1; }
extern "C" int b2sr_device_name(int device, char* buf, int buflen) {
    cudaDeviceProp p;
    if (!buf || buflen <= 0) return fail(B2SR_E_INVALID, "b2sr_device_name: no buffer");
    CUDA_TRY(cudaGetDeviceProperties(&p, device));
    snprintf(buf, buflen, "%s (sm_%d%d, %d SMs, %.0f GiB)", p.name, p.major, p.minor, p.multiProcessorCount,
             p.totalGlobalMem / 1073741824.0);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct LayerDev {
    int cin = 0, cout = 0;    // real channels
    int cinp = 0, noutp = 0;  // padded channels (kernel layout)
    uint8_t* wimg = nullptr;  // swizzled shared-memory image [kx][(2-ky)*noutp + o][cinp] fp16 (taps stacked along N)
    __half* wplain = nullptr; // [9][cinp][noutp] fp16 (simple path)
    float* bias = nullptr;    // [noutp]
    float* slope = nullptr;   // [noutp] (absent for the last layer)
};

struct Group {  // planes of identical size share one TMA tensor map per activation buffer
    int Ht, Wt, count;
    int64_t pix_base;
};

struct ResItems {  // work items of the planes scaled to resolution factor `res` (fused graph family), cut into <= max_cta ranges
    int res = 1, max_cta = 0, pairs = 0;  // pairs: items are band PAIRS (256 columns), one range per 2-CTA cluster
    std::vector<TcItem> items;
    std::vector<int> first;
    TcItem* d_items = nullptr;
    int* d_first = nullptr;
    int n_cta = 0;
    double out_px = 0;
};

struct Plan {
    int n = 0, h = 0, w = 0, tile = 0, halo = 0;
    std::vector<PlaneDev> planes;
    std::vector<int> plane_group;  // planes[i] belongs to groups[plane_group[i]]
    std::vector<std::unique_ptr<ResItems>> res_items;
    CUtensorMap* d_fmaps = nullptr;  // fused family: [launch][group] input tensor maps, then [segment stage][group] ring maps
    uint64_t fmaps_gen = 0;          // buffer generation the maps were encoded for
    TcgParams* d_fstages = nullptr;  // fused family, pipelined segments: parameters of every stage
    uint64_t fstages_key = 0;        // (buffer generation, ring generation, ring rows) the stage table was built for
    uint32_t* d_fflags = nullptr;    // `done` counters of one segment launch: [stage][band], B2SR_FLAG_STRIDE words apart
    int f_rr = 0;                    // ring rows the ring maps were encoded for
    std::vector<Group> groups;
    std::vector<TcItem> items;
    std::vector<int> item_first;  // CTA k of a launch processes items [item_first[k], item_first[k+1])
    int n_cta = 0;
    TcItem* d_items = nullptr;
    int* d_item_first = nullptr;
    double out_px = 0;  // exact output pixels of all items
    int64_t total_px = 0;
    int max_plane_px = 0;
    PlaneDev* d_planes = nullptr;
    CUtensorMap* d_maps = nullptr;  // [3 buffers][groups]: 0 = in16, 1 = ping, 2 = pong
    void* bound_in16 = nullptr;     // buffers the maps were encoded for
    void* bound_ping = nullptr;
    void* bound_pong = nullptr;
    // pipelined mode: CTA (layer, band) streams band `band` of every plane
    int nb = 0;    // bands of the widest plane
    int Wmax = 0;  // widest plane
    int64_t rows_total = 0;
    std::vector<TcItem> pitems;
    std::vector<int> pband_first;
    TcItem* d_pitems = nullptr;
    int* d_pband_first = nullptr;
    CUtensorMap* d_pmaps = nullptr;  // [groups]: ring tensor maps (all rings share geometry; the base is per layer -> [layers-1][groups])
    uint32_t* d_flags = nullptr;     // done[L][nb] | cons[L][nb]
    void* bound_rings = nullptr;
    int bound_rr = 0;
    ~Plan() {
        for (auto& r : res_items) {
            if (r->d_items) cudaFree(r->d_items);
            if (r->d_first) cudaFree(r->d_first);
        }
        if (d_fmaps) cudaFree(d_fmaps);
        if (d_fstages) cudaFree(d_fstages);
        if (d_fflags) cudaFree(d_fflags);
        if (d_pitems) cudaFree(d_pitems);
        if (d_pband_first) cudaFree(d_pband_first);
        if (d_pmaps) cudaFree(d_pmaps);
        if (d_flags) cudaFree(d_flags);
        if (d_planes) cudaFree(d_planes);
        if (d_maps) cudaFree(d_maps);
        if (d_items) cudaFree(d_items);
        if (d_item_first) cudaFree(d_item_first);
    }
};

#define B2SR_DBG_WORDS (320 * 16)  // stall-accounting words: up to 320 CTAs (two per SM) x 16

struct ProfRec {
    cudaEvent_t a, b;
    int kind;  // 0 other, 1 tcgen05 mid conv
    double px;
};

struct GraphOp {  // b2sr_graph_op + device copies of its parameters
    b2sr_graph_op op;
    float* w = nullptr;    // CONV: repacked [k*k][cin][coutp]; PRELU: slopes
    __half* wh = nullptr;  // CONV with cin % 16 == 0 and cout in {32, 64}: fp16 [k*k][cin][cout] for the wmma kernel
    float* b = nullptr;
    int coutp = 0;
    int ld_in(int j) const { return op.in_ld[j] ? op.in_ld[j] : op.in_c[j]; }
    int ld_out() const { return op.out_ld ? op.out_ld : op.out_c; }
};

struct FusedLaunch {  // one tcgen05 launch of a fused convolution (a 192 -> 64 convolution is two 192 -> 32 launches)
    int op = 0;        // index into b2sr_ctx::fops
    int co0 = 0;       // first output channel of this launch
    int nco = 0;       // real output channels of this launch
    int NOUT = 0;      // padded output channels (16 / 32 / 64)
    int G = 0;         // channel groups of 64 in the input view
    int cinp = 0;      // input view channels, padded to 16
    int sc_ks = 0;     // fused 1x1 shortcut: K slabs (0 = none)
    int pair = 0;      // 1: both 32-channel halves of a 64-channel convolution in one launch of 2-CTA clusters (TMA multicast)
    int slots = 0;     // shared-memory ring slots
    uint8_t* wimg = nullptr;
    uint8_t* wimg_flip = nullptr;  // the same image with the ky blocks swapped (rows walked bottom-up)
    float* bias = nullptr;
    float* slope = nullptr;
};

// Pipelined execution of a run of fused convolutions (tcg_pipe_kernel): one persistent launch per segment.
struct RingInst {  // a buffer (or the part of it) that is written AND read inside a segment: lives in an L2-resident row ring
    int buf = -1;
    std::vector<std::pair<int, int>> written, touched;  // channel ranges [lo, hi)
    std::vector<int> writer_op;                          // op that wrote written[i]
    int last_reader = -1;                                // last op of the segment that reads it
    size_t offset = 0;                                   // byte offset inside the ring arena (set per plan)
};
struct FusedStage {  // one CTA row of the persistent grid: one (half of a) fused convolution
    int launch = 0, op = 0, half = 0, variant = 0;
    int in_inst = -1, grp_ring[3] = {0, 0, 0};
    int out16_inst = -1, out32_inst = -1, res_inst[2] = {-1, -1};
    int gate_op = -1, bp_op = -1;  // op whose stages' `done` gates the ring input rows / frees the output ring slots
};
struct FusedSegment {
    int op_begin = 0, op_end = 0;
    std::vector<FusedStage> stages;
    std::vector<RingInst> inst;
    std::vector<int> op_first_stage;  // stage index of op (op_begin + k)'s first stage; size = n ops + 1
    int stage_base = 0;               // index of this segment's first stage among all segments' stages
};

struct b2sr_ctx {
    int device = 0, sms = 0;
    int family = B2SR_FAMILY_COMPACT;
    // B2SR_FAMILY_FUSED
    std::vector<b2sr_fused_op> fops;
    std::vector<b2sr_fused_buf> fbufs;
    std::vector<FusedLaunch> flaunch;
    std::vector<FusedLaunch> flaunch2;  // CTA-pair (cta_group::2) form of the ops that have one: whole convolution, full weight image
    std::vector<int> fop_pair2;         // op -> index into flaunch2, or -1
    int ablate = 0;                     // measurement only (B2SR_ABLATE / B2SR_OPT_ABLATE): see TcgParams::ablate
    int pair2 = 0;                      // use the CTA-pair form where it exists (B2SR_PAIR2=1 / B2SR_OPT_PAIR2); measured slower: opt-in
    std::vector<int> fop_first;   // launches of op i: flaunch[fop_first[i] .. fop_first[i+1])
    std::vector<void*> fbuf_ptr;
    std::vector<size_t> fbuf_cap;  // bytes
    uint64_t fbuf_gen = 1;         // bumped whenever a fused buffer or in16 moves
    // a box reads 128 bytes (one channel group) per pixel out of a pixel stride of up to 384 bytes: promotion to 256-byte
    // L2 fetches would pull the neighbouring group in with it (B2SR_L2_PROMO=256 restores that for measurements)
    CUtensorMapL2promotion l2_promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    int flip_rows = 0;             // experiment (B2SR_FLIP=1): alternate the row direction of consecutive fused launches
    int pair_halves = 1;           // launch split convolutions as 2-CTA clusters sharing their input (B2SR_PAIR=0: two launches)
    int pair_clusters = 0;         // clusters of two 227 KB CTAs the device can hold at once (measured at the first paired launch)
    int pdl = 1;                   // programmatic dependent launch of the fused convolution kernels (B2SR_PDL=0 disables)
    std::vector<FusedSegment> fsegs;  // runs of convolutions executed as one persistent launch each (an RRDB of 4x_Valar_v1)
    int fseg_stages = 0;              // stages of all segments
    int seg_pipe = 0;                 // opt-in (B2SR_SEG_PIPE=1 / B2SR_OPT_SEG_PIPE): measured slower than one launch per convolution, see DESIGN.md
    uint8_t* frings = nullptr;        // ring arena of the pipelined segments
    size_t cap_frings = 0;
    uint64_t fring_gen = 1;
    size_t l2_persist = 0;         // bytes of L2 set aside for persisting accesses (0 = feature off)
    size_t l2_window_max = 0;
    std::vector<GraphOp> gops;  // B2SR_FAMILY_GRAPH
    int g_slots = 0, g_in = 0, g_out = 0;
    std::vector<float*> slot_buf;
    std::vector<size_t> slot_cap;
    cudaStream_t stream = nullptr, copy_in = nullptr, copy_out = nullptr;
    // host-batch pipeline (b2sr_run_batch_host / b2sr_submit_batch_host): events that order the two staging slots across chunks AND
    // across calls, the chunk counter that alternates the slots, and one event per recent submission (its last D2H)
    cudaEvent_t hb_in_ready[2] = {nullptr, nullptr}, hb_compute_done[2] = {nullptr, nullptr}, hb_out_done[2] = {nullptr, nullptr};
    cudaEvent_t hb_ticket[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool hb_used[2] = {false, false};
    unsigned hb_seq = 0;
    uint64_t hb_next_ticket = 1;
    bool hb_pending = false;  // submissions that nobody has waited for may still be using d_in / d_out
    b2sr_net_desc desc{};
    int CF = 0, NL = 0, cout_last = 0;
    std::vector<LayerDev> layers;
    // scratch
    int64_t cap_px = 0;
    __half *in16 = nullptr, *ping = nullptr, *pong = nullptr;
    float* lastf = nullptr;
    int64_t cap_lastf = 0;
    int64_t cap_pp = 0;          // capacity (pixels) of ping/pong, allocated only for the layer-by-layer schedules
    __half* rings = nullptr;     // pipelined mode: (layers-1) rings of ring_rows x Wmax pixels
    size_t cap_rings = 0;
    int ring_rows = 0;  // 0 = sized so that all rings together stay L2-resident
    int pipe_debug = 0;
    int sm_limit = 0;          // B2SR_OPT_SM_LIMIT: treat the device as having this many SMs free for persistent grids (0 = all)
    int pipe_unavailable = 0;  // a cooperative launch was refused (SMs taken by MPS / another context): stay on the layer schedule
    double n_pipe_fallback = 0;
    long long* d_dbg = nullptr;
    uint8_t *d_in = nullptr, *d_out = nullptr;  // staging for host-memory calls
    size_t cap_in = 0, cap_out = 0;
    uint8_t *d_in2 = nullptr, *d_out2 = nullptr;  // second set for the double-buffered host pipeline
    size_t cap_in2 = 0, cap_out2 = 0;
    std::vector<std::unique_ptr<Plan>> plans;
    // options
    int impl = 0, profile = 0, max_batch = 0;
    // stats
    double n_launch = 0, n_tc = 0, n_pipe = 0, n_hmma = 0;
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> ev_pool;
};

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_tmapEncodeTiled g_encode = nullptr;

static int get_encode() {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) return fail(B2SR_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    g_encode = (PFN_tmapEncodeTiled)fn;
    return 0;
}

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
static inline int pad_channels(int c) { return c <= 32 ? 32 : 64; }
// canonical shared-memory swizzles (address based): 128B -> Swizzle<3,4,3>, 64B -> <2,4,3>, 32B -> <1,4,3>
static inline uint32_t swizzle_addr(uint32_t a, int row_bytes) {
    const uint32_t mask = row_bytes == 128 ? 7u : (row_bytes == 64 ? 3u : 1u);
    return a ^ (((a >> 7) & mask) << 4);
}

// This library holds sm_100a code only (arch-specific, not forward compatible): any other part -- including sm_101 / sm_103,
// which report major 10 too -- would fail at the first launch with "no kernel image"; refuse it up front instead.
static int check_device(int device, cudaDeviceProp* prop) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(B2SR_E_NODEVICE, "no CUDA device is visible (this library has no CPU path)");
    }
    if (device < 0 || device >= ndev) return fail(B2SR_E_NODEVICE, "device %d out of range (%d visible)", device, ndev);
    CUDA_TRY(cudaGetDeviceProperties(prop, device));
    if (prop->major != 10 || prop->minor != 0)
        return fail(B2SR_E_NODEVICE, "device %d (%s) is sm_%d%d; this library contains sm_100a code only", device, prop->name,
                    prop->major, prop->minor);
    CUDA_TRY(cudaSetDevice(device));
    cudaFuncAttributes fa;  // the image must actually load on this part
    if (cudaFuncGetAttributes(&fa, prep_kernel) != cudaSuccess) {
        const char* why = cudaGetErrorString(cudaGetLastError());
        return fail(B2SR_E_NODEVICE, "device %d (%s): the sm_100a kernels of this library do not load (%s)", device, prop->name, why);
    }
    return 0;
}

static bool fp16_exact(float v) { return __half2float(__float2half_rn(v)) == v; }

static void free_layers(b2sr_ctx* c) {
    for (auto& L : c->layers) {
        if (L.wimg) cudaFree(L.wimg);
        if (L.wplain) cudaFree(L.wplain);
        if (L.bias) cudaFree(L.bias);
        if (L.slope) cudaFree(L.slope);
    }
    c->layers.clear();
}

extern "C" void b2sr_destroy(b2sr_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->copy_in) cudaStreamSynchronize(c->copy_in);  // (pending b2sr_submit_batch_host work: copies into / out of the caller's buffers)
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copy_out) cudaStreamSynchronize(c->copy_out);
    c->plans.clear();
    free_layers(c);
    for (auto& g : c->gops) {
        if (g.w) cudaFree(g.w);
        if (g.wh) cudaFree(g.wh);
        if (g.b) cudaFree(g.b);
    }
    for (float* p : c->slot_buf)
        if (p) cudaFree(p);
    for (auto* vec : {&c->flaunch, &c->flaunch2})
        for (auto& L : *vec)
            for (void* p : {(void*)L.wimg, (void*)L.wimg_flip, (void*)L.bias, (void*)L.slope})
                if (p) cudaFree(p);
    for (void* p : c->fbuf_ptr)
        if (p) cudaFree(p);
    for (void* p : {(void*)c->in16, (void*)c->ping, (void*)c->pong, (void*)c->lastf, (void*)c->d_in, (void*)c->d_out,
                    (void*)c->d_in2, (void*)c->d_out2, (void*)c->rings, (void*)c->d_dbg, (void*)c->frings})
        if (p) cudaFree(p);
    for (auto& r : c->prof) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    for (cudaEvent_t* ev : {c->hb_in_ready, c->hb_compute_done, c->hb_out_done})
        for (int i = 0; i < 2; ++i)
            if (ev[i]) cudaEventDestroy(ev[i]);
    for (cudaEvent_t e : c->hb_ticket)
        if (e) cudaEventDestroy(e);
    if (c->copy_in) cudaStreamDestroy(c->copy_in);
    if (c->copy_out) cudaStreamDestroy(c->copy_out);
    delete c;
}

static int upload_layer(b2sr_ctx* c, const float* w, const float* bias, const float* slope, int cin, int cout,
                        int cinp, int noutp) {
    LayerDev L;
    L.cin = cin;
    L.cout = cout;
    L.cinp = cinp;
    L.noutp = noutp;
    const int PB = cinp * 2;
    std::vector<uint8_t> img((size_t)9 * noutp * PB, 0);
    std::vector<__half> plain((size_t)9 * cinp * noutp, __float2half(0.f));
    for (int o = 0; o < cout; ++o)
        for (int i = 0; i < cin; ++i)
            for (int t = 0; t < 9; ++t) {
                const float v = w[((size_t)o * cin + i) * 9 + t];
                if (!fp16_exact(v))
                    return fail(B2SR_E_UNSUPPORTED, "weight %g (conv out %d in %d tap %d) is not exactly representable in fp16", v,
                                o, i, t);
                const __half hv = __float2half_rn(v);
                // stacked-tap image: tile kx (1024-aligned) holds rows [W(ky=2) | W(ky=1) | W(ky=0)], noutp rows each;
                // element (row, channel i) sits at the swizzled address inside its tile
                const int ky = t / 3, kx = t % 3;
                const uint32_t a = swizzle_addr((uint32_t)(((2 - ky) * noutp + o) * PB + i * 2), PB);
                memcpy(&img[(size_t)kx * 3 * noutp * PB + a], &hv, 2);
                plain[((size_t)t * cinp + i) * noutp + o] = hv;
            }
    std::vector<float> b(noutp, 0.f), s(noutp, 0.f);
    for (int o = 0; o < cout; ++o) {
        b[o] = bias[o];
        if (slope) s[o] = slope[o];
    }
    CUDA_TRY(cudaMalloc(&L.wimg, img.size()));
    CUDA_TRY(cudaMemcpy(L.wimg, img.data(), img.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&L.wplain, plain.size() * 2));
    CUDA_TRY(cudaMemcpy(L.wplain, plain.data(), plain.size() * 2, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&L.bias, noutp * 4));
    CUDA_TRY(cudaMemcpy(L.bias, b.data(), noutp * 4, cudaMemcpyHostToDevice));
    if (slope) {
        CUDA_TRY(cudaMalloc(&L.slope, noutp * 4));
        CUDA_TRY(cudaMemcpy(L.slope, s.data(), noutp * 4, cudaMemcpyHostToDevice));
    }
    c->layers.push_back(L);
    return 0;
}

extern "C" int b2sr_create(b2sr_ctx** out, int device, const void* weights, size_t nbytes, const b2sr_net_desc* d) {
    if (!out || !weights || !d) return fail(B2SR_E_INVALID, "b2sr_create: null argument");
    *out = nullptr;
    if (d->family != B2SR_FAMILY_COMPACT) return fail(B2SR_E_UNSUPPORTED, "network family %d is not supported", d->family);
    if (d->cin != 3) return fail(B2SR_E_UNSUPPORTED, "cin = %d (only 3-channel images are supported)", d->cin);
    if (d->nf < 8 || d->nf > 64 || d->nf % 8) return fail(B2SR_E_UNSUPPORTED, "nf = %d (need a multiple of 8, <= 64)", d->nf);
    if (d->scale != 1 && d->scale != 2 && d->scale != 4) return fail(B2SR_E_UNSUPPORTED, "scale = %d (need 1, 2 or 4)", d->scale);
    if (d->n_mid < 1 || d->n_mid > 64) return fail(B2SR_E_INVALID, "n_mid = %d", d->n_mid);
    const int cin = d->cin, nf = d->nf, cl = cin * d->scale * d->scale;
    const size_t need = ((size_t)nf * cin * 9 + 2 * nf + (size_t)d->n_mid * ((size_t)nf * nf * 9 + 2 * nf) + (size_t)cl * nf * 9 + cl) * 4;
    if (nbytes != need) return fail(B2SR_E_INVALID, "weight blob is %zu bytes, network description needs %zu", nbytes, need);

    cudaDeviceProp prop;
    TRY(check_device(device, &prop));
    TRY(get_encode());

    b2sr_ctx* c = new b2sr_ctx();
    c->device = device;
    c->sms = prop.multiProcessorCount;
    c->desc = *d;
    c->CF = pad_channels(nf);
    c->cout_last = cl;
    c->NL = round_up(cl, 16);
    int rc = 0;
    do {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking) != cudaSuccess) {
            rc = fail(B2SR_E_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        const float* p = (const float*)weights;
        rc = upload_layer(c, p, p + (size_t)nf * cin * 9, p + (size_t)nf * cin * 9 + nf, cin, nf, 16, c->CF);
        if (rc) break;
        p += (size_t)nf * cin * 9 + 2 * nf;
        for (int i = 0; i < d->n_mid && !rc; ++i) {
            rc = upload_layer(c, p, p + (size_t)nf * nf * 9, p + (size_t)nf * nf * 9 + nf, nf, nf, c->CF, c->CF);
            p += (size_t)nf * nf * 9 + 2 * nf;
        }
        if (rc) break;
        rc = upload_layer(c, p, p + (size_t)cl * nf * 9, nullptr, nf, cl, c->CF, c->NL);
    } while (0);
    if (rc) {
        std::string keep = g_err;
        b2sr_destroy(c);
        g_err = keep;
        return rc;
    }
    *out = c;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// planning: planes (reference tiles), launch classes, work items, tensor maps
// ------------------------------------------------------------------------------------------------
static int build_plan(b2sr_ctx* c, int n, int h, int w, int tile, int halo, Plan** out) {
    for (size_t i = 0; i < c->plans.size(); ++i) {
        Plan* p = c->plans[i].get();
        if (p->n == n && p->h == h && p->w == w && p->tile == tile && p->halo == halo) {
            if (i + 1 != c->plans.size()) std::rotate(c->plans.begin() + i, c->plans.begin() + i + 1, c->plans.end());  // LRU: most recent last
            *out = p;
            return 0;
        }
    }
    if (c->plans.size() > 8) {
        cudaStreamSynchronize(c->stream);  // the evicted plan's device tables may still be in use by launches in flight
        c->plans.erase(c->plans.begin());
    }
    std::unique_ptr<Plan> P(new Plan());
    P->n = n, P->h = h, P->w = w, P->tile = tile, P->halo = halo;
    // reference process_tile :398-427 (tile = 0: the whole frame is one plane, apply_model :263-281)
    struct Rect {
        int iy0, iy1, ix0, ix1, cy0, cy1, cx0, cx1;
    };
    std::vector<Rect> rects;
    if (tile <= 0) {
        rects.push_back({0, h, 0, w, 0, h, 0, w});
    } else {
        const int ty = (h + tile - 1) / tile, tx = (w + tile - 1) / tile;
        for (int y = 0; y < ty; ++y)
            for (int x = 0; x < tx; ++x) {
                const int sy = y * tile, ey = std::min(sy + tile, h), sx = x * tile, ex = std::min(sx + tile, w);
                const int by0 = sy >= halo ? -halo : 0, by1 = ey <= h - halo ? halo : 0;
                const int bx0 = sx >= halo ? -halo : 0, bx1 = ex <= w - halo ? halo : 0;
                rects.push_back({sy + by0, ey + by1, sx + bx0, ex + bx1, sy, ey, sx, ex});
            }
    }
    // groups by plane size
    std::map<std::pair<int, int>, int> gidx;
    for (auto& r : rects) {
        auto key = std::make_pair(r.iy1 - r.iy0, r.ix1 - r.ix0);
        if (!gidx.count(key)) {
            gidx[key] = (int)P->groups.size();
            P->groups.push_back({key.first, key.second, 0, 0});
        }
        P->groups[gidx[key]].count += n;
    }
    int64_t base = 0;
    for (auto& g : P->groups) {
        g.pix_base = base;
        base += (int64_t)g.count * g.Ht * g.Wt;
        P->max_plane_px = std::max(P->max_plane_px, g.Ht * g.Wt);
    }
    P->total_px = base;
    // planes; plane index inside its group = frame-major
    std::vector<int> fill(P->groups.size(), 0);
    std::vector<int>& plane_group = P->plane_group;
    int64_t band_rows = 0;  // one unit of tensor work = one row of one 128-column band (M = 128 whatever the band's width)
    for (int f = 0; f < n; ++f)
        for (auto& r : rects) {
            const int Ht = r.iy1 - r.iy0, Wt = r.ix1 - r.ix0;
            const int gi = gidx[std::make_pair(Ht, Wt)];
            Group& g = P->groups[gi];
            const int pl = fill[gi]++;
            PlaneDev pd{};
            pd.frame = f, pd.fy0 = r.iy0, pd.fx0 = r.ix0, pd.Ht = Ht, pd.Wt = Wt;
            pd.cy0 = r.cy0, pd.cy1 = r.cy1, pd.cx0 = r.cx0, pd.cx1 = r.cx1;
            pd.pix_off = g.pix_base + (int64_t)pl * Ht * Wt;
            pd.gplane = pl;
            P->planes.push_back(pd);
            plane_group.push_back(gi);
            band_rows += (int64_t)Ht * ((Wt + TC_BW - 1) / TC_BW);
        }
    // items: the linear sequence (plane, band, row) is cut into one contiguous, equally long range per CTA, so every
    // CTA of a launch gets the same number of band rows (+-1) in as few pieces as possible
    const int ncta = (int)std::min<int64_t>(c->sms, band_rows);
    P->item_first.assign(1, 0);
    int64_t pos = 0;  // band rows emitted so far
    int cta = 0;
    auto cut_at = [&](int k) { return band_rows * k / ncta; };
    for (size_t pi = 0; pi < P->planes.size(); ++pi) {
        const PlaneDev& pd = P->planes[pi];
        for (int x0 = 0; x0 < pd.Wt; x0 += TC_BW) {
            int y0 = 0;
            while (y0 < pd.Ht) {
                const int64_t room = cut_at(cta + 1) - pos;  // rows left in this CTA's range (>= 1)
                const int rows = (int)std::min<int64_t>(pd.Ht - y0, room);
                TcItem it{};
                it.map = plane_group[pi], it.plane = pd.gplane, it.x0 = x0, it.y0 = y0;
                it.rows = rows, it.w = std::min(TC_BW, pd.Wt - x0);
                it.Ht = pd.Ht, it.Wt = pd.Wt, it.pix_off = pd.pix_off;
                it.frame = pd.frame, it.fy0 = pd.fy0, it.fx0 = pd.fx0;
                it.cy0 = pd.cy0, it.cy1 = pd.cy1, it.cx0 = pd.cx0, it.cx1 = pd.cx1;
                P->items.push_back(it);
                P->out_px += (double)it.rows * it.w;
                y0 += rows, pos += rows;
                if (pos == cut_at(cta + 1)) {
                    ++cta;
                    P->item_first.push_back((int)P->items.size());
                }
            }
        }
    }
    P->n_cta = ncta;
    // pipelined mode: per band, one item per plane (the whole band column); planes narrower than the band index get
    // a placeholder (w = 0) so that every band CTA walks the same global row sequence
    for (auto& g : P->groups) {
        P->nb = std::max(P->nb, (g.Wt + TC_BW - 1) / TC_BW);
        P->Wmax = std::max(P->Wmax, g.Wt);
    }
    P->pband_first.assign(1, 0);
    for (int b = 0; b < P->nb; ++b) {
        int64_t grow = 0;
        for (size_t pi = 0; pi < P->planes.size(); ++pi) {
            const PlaneDev& pd = P->planes[pi];
            TcItem it{};
            it.map = plane_group[pi], it.plane = pd.gplane, it.x0 = b * TC_BW, it.y0 = 0;
            it.rows = pd.Ht, it.w = std::max(0, std::min(TC_BW, pd.Wt - b * TC_BW));
            it.Ht = pd.Ht, it.Wt = pd.Wt, it.pix_off = pd.pix_off;
            it.frame = pd.frame, it.fy0 = pd.fy0, it.fx0 = pd.fx0;
            it.cy0 = pd.cy0, it.cy1 = pd.cy1, it.cx0 = pd.cx0, it.cx1 = pd.cx1;
            it.grow0 = (int32_t)grow;
            grow += pd.Ht;
            P->pitems.push_back(it);
        }
        P->rows_total = grow;
        P->pband_first.push_back((int)P->pitems.size());
    }
    CUDA_TRY(cudaMalloc(&P->d_pitems, P->pitems.size() * sizeof(TcItem)));
    CUDA_TRY(cudaMemcpy(P->d_pitems, P->pitems.data(), P->pitems.size() * sizeof(TcItem), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&P->d_pband_first, P->pband_first.size() * sizeof(int)));
    CUDA_TRY(cudaMemcpy(P->d_pband_first, P->pband_first.data(), P->pband_first.size() * sizeof(int), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&P->d_pmaps, (size_t)B2SR_PIPE_MAX_LAYERS * P->groups.size() * sizeof(CUtensorMap)));
    CUDA_TRY(cudaMalloc(&P->d_flags, (size_t)2 * B2SR_PIPE_MAX_LAYERS * P->nb * B2SR_FLAG_STRIDE * sizeof(uint32_t)));
    CUDA_TRY(cudaMalloc(&P->d_planes, P->planes.size() * sizeof(PlaneDev)));
    CUDA_TRY(cudaMemcpy(P->d_planes, P->planes.data(), P->planes.size() * sizeof(PlaneDev), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&P->d_items, P->items.size() * sizeof(TcItem)));
    CUDA_TRY(cudaMemcpy(P->d_items, P->items.data(), P->items.size() * sizeof(TcItem), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&P->d_item_first, P->item_first.size() * sizeof(int)));
    CUDA_TRY(cudaMemcpy(P->d_item_first, P->item_first.data(), P->item_first.size() * sizeof(int), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&P->d_maps, 3 * P->groups.size() * sizeof(CUtensorMap)));
    *out = P.get();
    c->plans.push_back(std::move(P));
    return 0;
}

static int ensure_scratch(b2sr_ctx* c, int64_t px, bool pingpong) {
    if (px > c->cap_px) {
        cudaStreamSynchronize(c->stream);
        if (c->in16) cudaFree(c->in16);
        c->in16 = nullptr;
        c->cap_px = 0;
        const int64_t cap = px + px / 16 + 1024;
        CUDA_TRY(cudaMalloc(&c->in16, (size_t)cap * 16 * 2));
        c->cap_px = cap;
        c->fbuf_gen += 1;
    }
    if (pingpong && px > c->cap_pp) {
        cudaStreamSynchronize(c->stream);
        for (void* p : {(void*)c->ping, (void*)c->pong})
            if (p) cudaFree(p);
        c->ping = c->pong = nullptr;
        c->cap_pp = 0;
        const int64_t cap = px + px / 16 + 1024;
        CUDA_TRY(cudaMalloc(&c->ping, (size_t)cap * c->CF * 2));
        CUDA_TRY(cudaMalloc(&c->pong, (size_t)cap * c->CF * 2));
        c->cap_pp = cap;
    }
    return 0;
}

static CUtensorMapSwizzle swizzle_for(int C) {
    return C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (C == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

static int encode_maps(b2sr_ctx* c, Plan* P) {
    if (P->bound_in16 == c->in16 && P->bound_ping == c->ping && P->bound_pong == c->pong) return 0;
    const size_t G = P->groups.size();
    std::vector<CUtensorMap> maps(3 * G);
    for (int b = 0; b < 3; ++b) {
        const int C = b == 0 ? 16 : c->CF;
        __half* basep = b == 0 ? c->in16 : (b == 1 ? c->ping : c->pong);
        if (!basep) continue;  // ping/pong exist only once a layer-by-layer schedule has run
        const CUtensorMapSwizzle sw = swizzle_for(C);
        for (size_t g = 0; g < G; ++g) {
            const Group& gr = P->groups[g];
            cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)gr.Wt, (cuuint64_t)gr.Ht, (cuuint64_t)gr.count};
            cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)gr.Wt * C * 2, (cuuint64_t)gr.Ht * gr.Wt * C * 2};
            cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)TC_PITCH, 1, 1};
            cuuint32_t es[4] = {1, 1, 1, 1};
            CUresult r = g_encode(&maps[b * G + g], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, basep + (size_t)gr.pix_base * C, dims,
                                  strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS)
                return fail(B2SR_E_CUDA, "cuTensorMapEncodeTiled failed (%d) for buffer %d group %zu (%dx%dx%d)", (int)r, b, g, gr.Ht,
                            gr.Wt, gr.count);
        }
    }
    CUDA_TRY(cudaMemcpyAsync(P->d_maps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    P->bound_in16 = c->in16, P->bound_ping = c->ping, P->bound_pong = c->pong;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// launches
// ------------------------------------------------------------------------------------------------
static int prof_begin(b2sr_ctx* c, int kind, double px) {
    if (!c->profile) return 0;
    ProfRec r;
    r.kind = kind, r.px = px;
    for (cudaEvent_t* e : {&r.a, &r.b}) {
        if (!c->ev_pool.empty()) {
            *e = c->ev_pool.back();
            c->ev_pool.pop_back();
        } else {
            CUDA_TRY(cudaEventCreate(e));
        }
    }
    CUDA_TRY(cudaEventRecord(r.a, c->stream));
    c->prof.push_back(r);
    return 0;
}
static int prof_end(b2sr_ctx* c) {
    if (!c->profile) return 0;
    CUDA_TRY(cudaEventRecord(c->prof.back().b, c->stream));
    return 0;
}

template <int CPIX, int NOUT, int SHUF, bool F32OUT>
static int launch_tc(b2sr_ctx* c, const Plan* plan, TcParams P) {
    using Cfg = TcCfg<CPIX, NOUT, SHUF>;
    static_assert(Cfg::ring_rows() >= 3, "shared-memory ring too small");
    P.items = plan->d_items, P.n_items = (int)plan->items.size();
    P.item_first = plan->d_item_first;
    const int smem = Cfg::smem_bytes();
    auto kern = tc_conv_kernel<CPIX, NOUT, SHUF, F32OUT>;
    TRY(raise_dyn_smem((const void*)kern, smem));
    const int grid = plan->n_cta;
    kern<<<grid, TC_THREADS, smem, c->stream>>>(P);
    CUDA_TRY(cudaGetLastError());
    c->n_launch += 1, c->n_tc += 1;
    return 0;
}

static int tc_layer(b2sr_ctx* c, Plan* P, int li, int in_buf, void* out, const uint8_t* frames, int fh, int fw, bool f32out) {
    const LayerDev& L = c->layers[li];
    const bool first = li == 0, last = li == (int)c->layers.size() - 1;
    {
        const Plan* lc = P;
        TcParams p{};
        p.maps = P->d_maps;
        p.map_base = in_buf * (int)P->groups.size();
        p.wimg = L.wimg, p.bias = L.bias, p.slope = L.slope;
        p.acc_scale = first ? (1.f / 255.f) : 1.f;
        p.out = out;
        p.frames_in = frames, p.frame_h = fh, p.frame_w = fw, p.scale = c->desc.scale;
        if (c->pipe_debug) {
            if (!c->d_dbg) CUDA_TRY(cudaMalloc(&c->d_dbg, B2SR_DBG_WORDS * sizeof(long long)));
            CUDA_TRY(cudaMemsetAsync(c->d_dbg, 0, B2SR_DBG_WORDS * sizeof(long long), c->stream));
            p.dbg = c->d_dbg;
        }
        TRY(prof_begin(c, (!first && !last) ? 1 : 0, P->out_px));
        int rc = B2SR_E_UNSUPPORTED;
        const int CF = c->CF, S = c->desc.scale;
        if (first) {
            rc = CF == 64 ? launch_tc<16, 64, 0, false>(c, lc, p) : launch_tc<16, 32, 0, false>(c, lc, p);
        } else if (!last) {
            rc = CF == 64 ? launch_tc<64, 64, 0, false>(c, lc, p) : launch_tc<32, 32, 0, false>(c, lc, p);
        } else {
#define LAST_CASE(cf, nl, s)                                                                            \
    if (CF == cf && c->NL == nl && S == s)                                                              \
        rc = f32out ? launch_tc<cf, nl, s, true>(c, lc, p) : launch_tc<cf, nl, s, false>(c, lc, p);
            LAST_CASE(64, 16, 1)
            LAST_CASE(64, 16, 2)
            LAST_CASE(64, 48, 4)
            LAST_CASE(32, 16, 1)
            LAST_CASE(32, 16, 2)
            LAST_CASE(32, 48, 4)
#undef LAST_CASE
            if (rc == B2SR_E_UNSUPPORTED) fail(rc, "no tcgen05 kernel for nf-pad %d, last-pad %d, scale %d", CF, c->NL, S);
        }
        TRY(rc);
        TRY(prof_end(c));
        if (c->pipe_debug) {  // stall accounting of this layer's launch (synchronises)
            std::vector<long long> h(B2SR_DBG_WORDS);
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            CUDA_TRY(cudaMemcpy(h.data(), c->d_dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
            double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int k = 0; k < P->n_cta; ++k)
                for (int j = 0; j < 8; ++j) v[j] += (double)h[(size_t)k * 8 + j] / P->n_cta;
            fprintf(stderr, "b2sr layer %2d (%d CTAs), kcycles mean: producer total %7.0f (wait slot %6.0f) | issuer: wait data %6.0f, wait tmem %6.0f | epilogue w2: wait tfull %6.0f\n",
                    li, P->n_cta, v[0] / 1e3, v[6] / 1e3, v[3] / 1e3, v[4] / 1e3, v[5] / 1e3);
        }
    }
    return 0;
}

static int simple_layer(b2sr_ctx* c, Plan* P, int li, const __half* in, void* out) {
    const LayerDev& L = c->layers[li];
    const bool first = li == 0, last = li == (int)c->layers.size() - 1;
    const int total = P->max_plane_px * (L.noutp / 8);
    dim3 grid((unsigned)std::min(4096, (total + 255) / 256), (unsigned)P->planes.size());
    TRY(prof_begin(c, 0, 0));
    if (last)
        simple_conv_kernel<true><<<grid, 256, 0, c->stream>>>(in, L.cinp, L.wplain, L.noutp, L.bias, L.slope,
                                                              first ? 1.f / 255.f : 1.f, P->d_planes, out);
    else
        simple_conv_kernel<false><<<grid, 256, 0, c->stream>>>(in, L.cinp, L.wplain, L.noutp, L.bias, L.slope,
                                                               first ? 1.f / 255.f : 1.f, P->d_planes, out);
    CUDA_TRY(cudaGetLastError());
    TRY(prof_end(c));
    c->n_launch += 1;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// pipelined schedule: one persistent launch, CTA = (layer, band), activations in L2-resident row rings
// ------------------------------------------------------------------------------------------------
// The CTAs of the persistent kernel spin on each other's counters, so the whole grid must be co-resident.  Three guards:
// (1) layers x bands must not exceed the SMs this context may count on (B2SR_OPT_SM_LIMIT lowers that, e.g. under MPS with
// an active-thread percentage); (2) the launch is cooperative, so the driver itself refuses a grid it cannot make fully
// resident (cudaErrorCooperativeLaunchTooLarge) instead of starting part of it; (3) a refusal is not an error: the pass
// falls back to the layer-by-layer schedule (same kernels, same results) and the context stops trying.
static int usable_sms(const b2sr_ctx* c) { return c->sm_limit > 0 ? std::min(c->sm_limit, c->sms) : c->sms; }

static bool pipe_fits(const b2sr_ctx* c, const Plan* P) {
    const int L = (int)c->layers.size();
    return !c->pipe_unavailable && L <= B2SR_PIPE_MAX_LAYERS && P->nb >= 1 && (int64_t)L * P->nb <= usable_sms(c);
}

// Which schedule `impl = 0` (auto) picks.  Measured on B200 (tools/hurr_check.py, 8 frames per pass): for the nf = 24 network
// (1x_HurrDeblur) one launch per layer is 2.7x FASTER than the persistent grid -- 0.164 vs 0.445 ms per 540p frame: its layers
// are issue-bound (~1 000 cycles per band row for 288 cycles of MMAs), the persistent grid only occupies layers x bands = 80 SMs,
// and the 64-byte-per-pixel activations cost little HBM time -- so the persistent schedule is the default for 64-channel
// networks only (there it removes 9 GB of activation traffic per frame).
static bool auto_pipe(const b2sr_ctx* c) { return c->impl == 3 || (c->impl == 0 && c->CF > 32); }

#define B2SR_PIPE_REFUSED 1  // (positive: not an error code of the ABI)

template <int CF, int NL, int S, bool F32OUT>
static int launch_pipe(b2sr_ctx* c, const PipeParams& Q) {
    const int smem = TcPipeCfg<CF, NL, S>::smem_bytes();
    auto kern = tc_pipe_kernel<CF, NL, S, F32OUT>;
    TRY(raise_dyn_smem((const void*)kern, smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(Q.n_layers * Q.nb)), cfg.blockDim = dim3(TC_THREADS), cfg.dynamicSmemBytes = (size_t)smem, cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    // (B2SR_COOP=0: plain launch -- a profiling aid only: `ncu --set full` replays a kernel many times and does not collect
    // cooperative launches in multi-pass mode; without the attribute co-residency is the caller's responsibility again)
    static const bool coop = !(getenv("B2SR_COOP") && atoi(getenv("B2SR_COOP")) == 0);
    cfg.attrs = attr, cfg.numAttrs = coop ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, Q);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorLaunchOutOfResources) {
        cudaGetLastError();
        return B2SR_PIPE_REFUSED;
    }
    if (e != cudaSuccess) return fail(B2SR_E_CUDA, "persistent kernel launch failed: %s", cudaGetErrorString(e));
    return 0;
}

static int run_pipe(b2sr_ctx* c, Plan* P, const uint8_t* d_frames, void* d_out, bool f32out) {
    const int L = (int)c->layers.size(), G = (int)P->groups.size(), nb = P->nb, CF = c->CF;
    TRY(ensure_scratch(c, P->total_px, false));
    TRY(encode_maps(c, P));
    // Measured on B200 (profiles/): with 17 rings of 970 x 128 B rows, 16 rows per ring (34 MB) keep every ring write in
    // L2 (DRAM writes = the output frames only), 24 rows start to spill, 32 rows write 9 GB per 4 frames back to HBM.
    int RR = c->ring_rows;
    if (RR <= 0) {
        const double row_bytes = (double)(L - 1) * P->Wmax * CF * 2;
        RR = (int)(36.0e6 / row_bytes) / 4 * 4;
        RR = std::max(8, std::min(64, RR));
    }
    const size_t ring_px = (size_t)RR * P->Wmax;
    const size_t need = (size_t)(L - 1) * ring_px * CF * 2;
    if (need > c->cap_rings) {
        cudaStreamSynchronize(c->stream);
        if (c->rings) cudaFree(c->rings);
        c->rings = nullptr, c->cap_rings = 0;
        CUDA_TRY(cudaMalloc(&c->rings, need));
        c->cap_rings = need;
    }
    if (P->bound_rings != c->rings || P->bound_rr != RR) {
        // ring l (output of layer l) as a {C, Wt, RR, 1} tensor per plane-size group: rows are Wmax pixels apart,
        // columns beyond the group's Wt and rows outside [0, RR) read as zeros
        std::vector<CUtensorMap> maps((size_t)(L - 1) * G);
        for (int l = 0; l + 1 < L; ++l)
            for (int g = 0; g < G; ++g) {
                const Group& gr = P->groups[g];
                cuuint64_t dims[4] = {(cuuint64_t)CF, (cuuint64_t)gr.Wt, (cuuint64_t)RR, 1};
                cuuint64_t strides[3] = {(cuuint64_t)CF * 2, (cuuint64_t)P->Wmax * CF * 2, (cuuint64_t)ring_px * CF * 2};
                cuuint32_t box[4] = {(cuuint32_t)CF, (cuuint32_t)TC_PITCH, 1, 1};
                cuuint32_t es[4] = {1, 1, 1, 1};
                CUresult r = g_encode(&maps[(size_t)l * G + g], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, c->rings + (size_t)l * ring_px * CF,
                                      dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(CF),
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) return fail(B2SR_E_CUDA, "cuTensorMapEncodeTiled failed (%d) for ring %d group %d", (int)r, l, g);
            }
        CUDA_TRY(cudaMemcpyAsync(P->d_pmaps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        P->bound_rings = c->rings, P->bound_rr = RR;
    }
    // The first layer's CTAs read the packed u8 frames themselves (tc_conv.cuh, frame-row producer); B2SR_DIRECT_IN=0 restores
    // the separate prep_kernel pass that expands the frames to 16-channel fp16 planes in HBM first.
    static const bool direct_in = !(getenv("B2SR_DIRECT_IN") && atoi(getenv("B2SR_DIRECT_IN")) == 0);
    if (!direct_in) {
        dim3 grid((unsigned)std::min(2048, (P->max_plane_px + 255) / 256), (unsigned)P->planes.size());
        TRY(prof_begin(c, 0, 0));
        prep_kernel<<<grid, 256, 0, c->stream>>>(d_frames, P->h, P->w, P->d_planes, c->in16);
        CUDA_TRY(cudaGetLastError());
        TRY(prof_end(c));
        c->n_launch += 1;
    }
    const size_t fl = (size_t)nb * B2SR_FLAG_STRIDE;  // counter words per layer
    CUDA_TRY(cudaMemsetAsync(P->d_flags, 0, 2 * B2SR_PIPE_MAX_LAYERS * fl * sizeof(uint32_t), c->stream));
    PipeParams Q{};
    Q.n_layers = L, Q.nb = nb;
    if (c->pipe_debug) {
        if (!c->d_dbg) CUDA_TRY(cudaMalloc(&c->d_dbg, B2SR_DBG_WORDS * sizeof(long long)));
        CUDA_TRY(cudaMemsetAsync(c->d_dbg, 0, B2SR_DBG_WORDS * sizeof(long long), c->stream));
        Q.dbg = c->d_dbg;
    }
    uint32_t* done = P->d_flags;
    uint32_t* cons = P->d_flags + B2SR_PIPE_MAX_LAYERS * fl;
    for (int l = 0; l < L; ++l) {
        const LayerDev& Ld = c->layers[l];
        TcParams& p = Q.layers[l];
        p.maps = l == 0 ? P->d_maps : P->d_pmaps;  // layer 0 reads the in16 planes, layer l > 0 reads ring l-1
        p.map_base = l == 0 ? 0 : (l - 1) * G;
        p.items = P->d_pitems, p.item_first = P->d_pband_first, p.n_items = (int)P->pitems.size();
        p.wimg = Ld.wimg, p.bias = Ld.bias, p.slope = Ld.slope;
        p.acc_scale = l == 0 ? (1.f / 255.f) : 1.f;
        p.out = l == L - 1 ? d_out : (void*)(c->rings + (size_t)l * ring_px * CF);
        p.frames_in = d_frames, p.frame_h = P->h, p.frame_w = P->w, p.scale = c->desc.scale;
        p.ring_in = l > 0, p.ring_out = l < L - 1;
        p.direct_in = l == 0 && direct_in;
        p.frames_end = d_frames + (size_t)P->n * P->h * P->w * 3;
        p.RR = RR, p.Wmax = P->Wmax, p.nb = nb;
        p.done_in = l > 0 ? done + (size_t)(l - 1) * fl : nullptr;
        p.done_out = done + (size_t)l * fl;
        p.cons_self = cons + (size_t)l * fl;
        p.cons_next = l < L - 1 ? cons + (size_t)(l + 1) * fl : nullptr;
    }
    TRY(prof_begin(c, 2, P->out_px));
    int rc = B2SR_E_UNSUPPORTED;
    const int S = c->desc.scale;
#define PIPE_CASE(cf, nl, s) \
    if (CF == cf && c->NL == nl && S == s) rc = f32out ? launch_pipe<cf, nl, s, true>(c, Q) : launch_pipe<cf, nl, s, false>(c, Q);
    PIPE_CASE(64, 16, 2)
    PIPE_CASE(64, 48, 4)
    PIPE_CASE(32, 16, 1)
    PIPE_CASE(64, 16, 1)
#undef PIPE_CASE
    if (rc == B2SR_E_UNSUPPORTED) return fail(rc, "no pipelined kernel for nf-pad %d, last-pad %d, scale %d", CF, c->NL, S);
    if (rc == B2SR_PIPE_REFUSED) {
        if (c->profile) {  // the record opened by prof_begin has no end event: drop it
            c->ev_pool.push_back(c->prof.back().a), c->ev_pool.push_back(c->prof.back().b);
            c->prof.pop_back();
        }
        return rc;
    }
    TRY(rc);
    TRY(prof_end(c));
    c->n_launch += 1, c->n_tc += 1, c->n_pipe += 1;
    if (c->pipe_debug) {
        std::vector<long long> h(B2SR_DBG_WORDS);
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        CUDA_TRY(cudaMemcpy(h.data(), c->d_dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        fprintf(stderr, "b2sr pipe timing, %% of CTA time, mean over %d bands: layer kcycles | producer: starved(done) wait-empty | mma: wait-full wait-tempty | epilogue w2: back-pressure wait-tfull\n", nb);
        for (int l = 0; l < L; ++l) {
            double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int k = 0; k < nb; ++k)
                for (int j = 0; j < 8; ++j) v[j] += (double)h[(l * nb + k) * 8 + j];
            const double t = std::max(v[0], 1.0);
            fprintf(stderr, "  L%02d: %8.0f | %5.1f %5.1f | %5.1f %5.1f | %5.1f %5.1f | publishes/band %6.0f, %6.0f cycles each\n", l, v[0] / nb / 1e3,
                    100 * v[1] / t, 100 * v[6] / t, 100 * v[3] / t, 100 * v[4] / t, 100 * v[2] / t, 100 * v[5] / t,
                    (double)(h[(l * nb) * 8 + 7] % 1000000LL), (double)(h[(l * nb) * 8 + 7] / 1000000LL));
        }
    }
    return 0;
}


// ------------------------------------------------------------------------------------------------
// generic graph engine (B2SR_FAMILY_GRAPH): 4x_Valar_v1 and anything else the Compact kernels do not cover
// ------------------------------------------------------------------------------------------------
extern "C" int b2sr_create_graph(b2sr_ctx** out, int device, const b2sr_graph_op* ops, int n_ops, int n_slots, int in_slot,
                                 int out_slot, int scale, const void* weights, size_t nbytes) {
    if (!out || !ops || !weights) return fail(B2SR_E_INVALID, "b2sr_create_graph: null argument");
    *out = nullptr;
    if (n_ops < 1 || n_slots < 2 || n_slots > 4096 || in_slot < 0 || in_slot >= n_slots || out_slot < 0 || out_slot >= n_slots)
        return fail(B2SR_E_INVALID, "b2sr_create_graph: bad graph description");
    if (scale != 1 && scale != 2 && scale != 4) return fail(B2SR_E_UNSUPPORTED, "scale = %d (need 1, 2 or 4)", scale);
    cudaDeviceProp prop;
    TRY(check_device(device, &prop));
    TRY(get_encode());
    const float* wb = (const float*)weights;
    const int64_t nfl = (int64_t)(nbytes / 4);
    b2sr_ctx* c = new b2sr_ctx();
    c->device = device, c->sms = prop.multiProcessorCount, c->family = B2SR_FAMILY_GRAPH;
    c->desc.family = B2SR_FAMILY_GRAPH, c->desc.cin = 3, c->desc.scale = scale;
    c->CF = 64, c->NL = 16;  // unused by this family (scratch sizing only)
    c->g_slots = n_slots, c->g_in = in_slot, c->g_out = out_slot;
    c->slot_buf.assign(n_slots, nullptr);
    c->slot_cap.assign(n_slots, 0);
    int rc = 0;
    int out_c = 0, out_res = 0;
    do {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking) != cudaSuccess) {
            rc = fail(B2SR_E_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        for (int i = 0; i < n_ops && !rc; ++i) {
            GraphOp g;
            g.op = ops[i];
            const b2sr_graph_op& o = g.op;
            if (o.nin < 1 || o.nin > 6 || o.out < 0 || o.out >= n_slots || o.out_c < 1 || o.out_off < 0 || o.in_res < 1 || o.out_res < 1 ||
                (o.out_ld && o.out_ld < o.out_off + o.out_c)) {
                rc = fail(B2SR_E_INVALID, "op %d: bad operands", i);
                break;
            }
            for (int j = 0; j < o.nin; ++j)
                if (o.in[j] < 0 || o.in[j] >= n_slots || o.in_c[j] < 1 || o.in_off[j] < 0 || (o.in_ld[j] && o.in_ld[j] < o.in_off[j] + o.in_c[j]))
                    rc = fail(B2SR_E_INVALID, "op %d: bad input %d", i, j);
            if (rc) break;
            switch (o.type) {
                case B2SR_OP_CONV: {
                    if ((o.k != 1 && o.k != 3) || o.cin != o.in_c[0] || o.cout != o.out_c || o.nin != 1 || o.out_res != o.in_res) {
                        rc = fail(B2SR_E_UNSUPPORTED, "op %d: convolution k=%d cin=%d (input has %d) cout=%d", i, o.k, o.cin, o.in_c[0], o.cout);
                        break;
                    }
                    const int64_t nw = (int64_t)o.cout * o.cin * o.k * o.k;
                    if (o.w_off < 0 || o.w_off + nw > nfl || (o.b_off >= 0 && o.b_off + o.cout > nfl)) {
                        rc = fail(B2SR_E_INVALID, "op %d: weights outside the blob", i);
                        break;
                    }
                    g.coutp = round_up(o.cout, 4);
                    std::vector<float> w((size_t)o.k * o.k * o.cin * g.coutp, 0.f);
                    for (int oc = 0; oc < o.cout; ++oc)
                        for (int ic = 0; ic < o.cin; ++ic)
                            for (int t = 0; t < o.k * o.k; ++t)
                                w[((size_t)t * o.cin + ic) * g.coutp + oc] = wb[o.w_off + ((int64_t)oc * o.cin + ic) * o.k * o.k + t];
                    if (cudaMalloc(&g.w, w.size() * 4) != cudaSuccess ||
                        cudaMemcpy(g.w, w.data(), w.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
                        rc = fail(B2SR_E_NOMEM, "op %d: weight upload failed", i);
                        break;
                    }
                    // tensor-core path: fp16 weights, 16-byte aligned strided views
                    if (o.cin % 16 == 0 && (o.cout == 32 || o.cout == 64) && g.ld_in(0) % 4 == 0 && o.in_off[0] % 4 == 0 && g.ld_out() % 4 == 0 &&
                        o.out_off % 4 == 0) {
                        bool exact = true;
                        std::vector<__half> hw((size_t)o.k * o.k * o.cin * o.cout);
                        for (int t = 0; t < o.k * o.k; ++t)
                            for (int ic = 0; ic < o.cin; ++ic)
                                for (int oc = 0; oc < o.cout; ++oc) {
                                    const float v = w[((size_t)t * o.cin + ic) * g.coutp + oc];
                                    exact = exact && fp16_exact(v);
                                    hw[((size_t)t * o.cin + ic) * o.cout + oc] = __float2half_rn(v);
                                }
                        if (exact) {  // (true for every model the reference ships: their weights are stored as fp16)
                            if (cudaMalloc(&g.wh, hw.size() * 2) != cudaSuccess ||
                                cudaMemcpy(g.wh, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) {
                                rc = fail(B2SR_E_NOMEM, "op %d: weight upload failed", i);
                                break;
                            }
                        }
                    }
                    if (o.b_off >= 0) {
                        if (cudaMalloc(&g.b, o.cout * 4) != cudaSuccess ||
                            cudaMemcpy(g.b, wb + o.b_off, o.cout * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
                            rc = fail(B2SR_E_NOMEM, "op %d: bias upload failed", i);
                            break;
                        }
                    }
                    break;
                }
                case B2SR_OP_PRELU:
                    if (o.w_off < 0 || o.w_off + o.in_c[0] > nfl || o.out_c != o.in_c[0] || o.out_res != o.in_res) {
                        rc = fail(B2SR_E_INVALID, "op %d: bad PReLU", i);
                        break;
                    }
                    if (cudaMalloc(&g.w, o.in_c[0] * 4) != cudaSuccess ||
                        cudaMemcpy(g.w, wb + o.w_off, o.in_c[0] * 4, cudaMemcpyHostToDevice) != cudaSuccess)
                        rc = fail(B2SR_E_NOMEM, "op %d: slope upload failed", i);
                    break;
                case B2SR_OP_PIXELSHUFFLE:
                    if (o.r < 1 || o.in_c[0] != o.out_c * o.r * o.r || o.out_res != o.in_res * o.r) rc = fail(B2SR_E_INVALID, "op %d: bad pixel shuffle", i);
                    break;
                case B2SR_OP_NEAREST:
                    if (o.r < 1 || o.r > 8 || o.out_c != o.in_c[0] || o.out_res != o.in_res * o.r) rc = fail(B2SR_E_INVALID, "op %d: bad resize", i);
                    break;
                case B2SR_OP_ADD:
                    if (o.nin != 2 || o.in_c[0] != o.in_c[1] || o.out_c != o.in_c[0] || o.out_res != o.in_res) rc = fail(B2SR_E_INVALID, "op %d: bad add", i);
                    break;
                case B2SR_OP_CONCAT: {
                    int tot = 0;
                    for (int j = 0; j < o.nin; ++j) tot += o.in_c[j];
                    if (tot != o.out_c || o.out_res != o.in_res) rc = fail(B2SR_E_INVALID, "op %d: bad concat", i);
                    break;
                }
                default:
                    rc = fail(B2SR_E_UNSUPPORTED, "op %d: unknown type %d", i, o.type);
            }
            if (rc) {
                if (g.w) cudaFree(g.w);
                if (g.wh) cudaFree(g.wh);
                if (g.b) cudaFree(g.b);
                break;
            }
            if (o.out == out_slot) out_c = o.out_c, out_res = o.out_res;
            c->gops.push_back(g);
        }
        if (!rc && (out_c != 3 || out_res != scale))
            rc = fail(B2SR_E_INVALID, "graph output has %d channels at x%d, expected 3 at x%d", out_c, out_res, scale);
        if (!rc && (c->gops.back().op.out != out_slot || c->gops.back().op.out_off != 0))
            rc = fail(B2SR_E_INVALID, "the last op must produce the output slot");
    } while (0);
    if (rc) {
        std::string keep = g_err;
        b2sr_destroy(c);
        g_err = keep;
        return rc;
    }
    *out = c;
    return 0;
}

static int run_graph(b2sr_ctx* c, Plan* P, const uint8_t* d_frames, void* d_out, bool f32out) {
    const int S = c->desc.scale;
    // slot capacities for the largest plane
    std::vector<size_t> need(c->g_slots, 0);
    const size_t mp = (size_t)P->max_plane_px;
    need[c->g_in] = mp * 3;
    for (auto& g : c->gops) need[g.op.out] = std::max(need[g.op.out], mp * g.op.out_res * g.op.out_res * g.ld_out());
    for (int s = 0; s < c->g_slots; ++s)
        if (need[s] > c->slot_cap[s]) {
            cudaStreamSynchronize(c->stream);
            if (c->slot_buf[s]) cudaFree(c->slot_buf[s]);
            c->slot_buf[s] = nullptr, c->slot_cap[s] = 0;
            CUDA_TRY(cudaMalloc(&c->slot_buf[s], need[s] * sizeof(float)));
            c->slot_cap[s] = need[s];
        }
    auto blocks = [](size_t n) { return (unsigned)std::min<size_t>(148 * 16, (n + 255) / 256); };
    for (const PlaneDev& pl : P->planes) {  // tiles run one after the other, like the reference's loop (:502-516)
        TRY(prof_begin(c, 0, 0));
        g_input_kernel<<<blocks((size_t)pl.Ht * pl.Wt * 3), 256, 0, c->stream>>>(d_frames, P->h, P->w, pl, c->slot_buf[c->g_in]);
        c->n_launch += 1;
        for (auto& g : c->gops) {
            const b2sr_graph_op& o = g.op;
            const int H = pl.Ht * o.in_res, W = pl.Wt * o.in_res;
            const size_t px = (size_t)H * W;
            const float* in0 = c->slot_buf[o.in[0]] + o.in_off[0];
            float* outp = c->slot_buf[o.out] + o.out_off;
            const int ld0 = g.ld_in(0), ldo = g.ld_out();
            switch (o.type) {
                case B2SR_OP_CONV: {
                    if (g.wh && c->impl != 1) {  // HMMA path (B2SR_OPT_IMPL = 1 forces the fp32 CUDA-core kernel)
                        const int halo = o.k / 2;
                        const size_t smem = std::max((size_t)(GW_TY + 2 * halo) * (GW_TX + 2 * halo) * (o.cin + GW_PAD) * 2,
                                                     (size_t)GW_TY * GW_TX * o.cout * 4);
                        const unsigned nbw = (unsigned)(((H + GW_TY - 1) / GW_TY) * ((W + GW_TX - 1) / GW_TX));
#define WMMA_CASE(kk, nf)                                                                                                   \
    if (o.k == kk && o.cout == nf * 16) {                                                                                   \
        TRY(raise_dyn_smem((const void*)g_conv_wmma_kernel<kk, nf>, (int)smem)); \
        g_conv_wmma_kernel<kk, nf><<<nbw, GW_THREADS, smem, c->stream>>>(in0, ld0, H, W, o.cin, g.wh, g.b, o.act, o.slope, outp, ldo); \
    }
                        WMMA_CASE(3, 2)
                        WMMA_CASE(3, 4)
                        WMMA_CASE(1, 2)
                        WMMA_CASE(1, 4)
#undef WMMA_CASE
                        c->n_hmma += 1;
                        break;
                    }
                    const unsigned nb = blocks(px * (g.coutp / 4));
                    if (o.k == 3)
                        g_conv_kernel<3><<<nb, 256, 0, c->stream>>>(in0, ld0, H, W, o.cin, g.w, g.b, o.cout, g.coutp, o.act, o.slope, outp, ldo);
                    else
                        g_conv_kernel<1><<<nb, 256, 0, c->stream>>>(in0, ld0, H, W, o.cin, g.w, g.b, o.cout, g.coutp, o.act, o.slope, outp, ldo);
                    break;
                }
                case B2SR_OP_PRELU:
                    g_prelu_kernel<<<blocks(px * o.in_c[0]), 256, 0, c->stream>>>(in0, ld0, px * o.in_c[0], o.in_c[0], g.w, outp, ldo);
                    break;
                case B2SR_OP_PIXELSHUFFLE:
                    g_pixelshuffle_kernel<<<blocks(px * o.in_c[0]), 256, 0, c->stream>>>(in0, ld0, H, W, o.out_c, o.r, outp, ldo);
                    break;
                case B2SR_OP_NEAREST:
                    g_nearest_kernel<<<blocks(px * o.r * o.r * o.in_c[0]), 256, 0, c->stream>>>(in0, ld0, H, W, o.in_c[0], o.r, outp, ldo);
                    break;
                case B2SR_OP_ADD:
                    g_add_kernel<<<blocks(px * o.in_c[0]), 256, 0, c->stream>>>(in0, ld0, c->slot_buf[o.in[1]] + o.in_off[1], g.ld_in(1),
                                                                               px * o.in_c[0], o.in_c[0], o.coef[0], o.coef[1], o.plain, outp, ldo);
                    break;
                case B2SR_OP_CONCAT: {
                    int off = 0;
                    for (int j = 0; j < o.nin; ++j) {
                        g_concat_kernel<<<blocks(px * o.in_c[j]), 256, 0, c->stream>>>(c->slot_buf[o.in[j]] + o.in_off[j], g.ld_in(j), px, o.in_c[j],
                                                                                      ldo, off, outp);
                        off += o.in_c[j];
                        c->n_launch += 1;
                    }
                    c->n_launch -= 1;
                    break;
                }
            }
            c->n_launch += 1;
        }
        const b2sr_graph_op& last = c->gops.back().op;
        const size_t on = (size_t)(pl.cy1 - pl.cy0) * S * (pl.cx1 - pl.cx0) * S * 3;
        if (f32out)
            g_output_kernel<true><<<blocks(on), 256, 0, c->stream>>>(c->slot_buf[c->g_out], c->gops.back().ld_out(), pl, S, P->h, P->w, d_out);
        else
            g_output_kernel<false><<<blocks(on), 256, 0, c->stream>>>(c->slot_buf[c->g_out], c->gops.back().ld_out(), pl, S, P->h, P->w, d_out);
        (void)last;
        c->n_launch += 1;
        CUDA_TRY(cudaGetLastError());
        TRY(prof_end(c));
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// fused tcgen05 graph engine (B2SR_FAMILY_FUSED): RRDB-style graphs (4x_Valar_v1) on csrc/tc_gconv.cuh
// ------------------------------------------------------------------------------------------------
static int fused_fit(int NOUT, bool final, int G, bool sc) {
    if (final) return TcgCfg<16, 1>::ring_fit(G);
    if (sc) return TcgCfg<32, 0, true>::ring_fit(G);
    return NOUT == 64 ? TcgCfg<64, 0>::ring_fit(G) : TcgCfg<32, 0>::ring_fit(G);
}

// Stacked, pre-swizzled weight image of output channels [co0, co0 + nco) of a convolution:
// [group g][kx][(2 - ky) * NOUT + o][64 channels], rows of 128 bytes in the 128-byte swizzle pattern.
static int upload_fused_launch(FusedLaunch& L, const b2sr_fused_op& o, const float* wb, bool with_flip) {
    const int K = o.k, PB = TCG_PB;
    const size_t main_bytes = (size_t)L.G * 9 * L.NOUT * PB;
    const int parts = L.pair ? 2 : 1;  // a paired launch carries both halves: [image of channels co0..][image of channels co0 + NOUT..]
    std::vector<uint8_t> img(parts * main_bytes + (L.sc_ks ? (size_t)L.NOUT * PB : 0), 0), flp(img.size(), 0);
    if (L.sc_ks)  // shortcut image behind the 3x3 tiles: [o][64 ch], 128-byte swizzled rows
        for (int oc = 0; oc < L.nco; ++oc)
            for (int ic = 0; ic < o.sc_cin; ++ic) {
                const float v = wb[o.sc_w_off + (int64_t)(L.co0 + oc) * o.sc_cin + ic];
                if (!fp16_exact(v)) return fail(B2SR_E_UNSUPPORTED, "shortcut weight %g is not exactly representable in fp16", v);
                const __half hv = __float2half_rn(v);
                const uint32_t a = swizzle_addr((uint32_t)(oc * PB + ic * 2), PB);
                memcpy(&img[main_bytes + a], &hv, 2);
                memcpy(&flp[main_bytes + a], &hv, 2);
            }
    for (int part = 0; part < parts; ++part)
    for (int oc = 0; oc < L.nco; ++oc)
        for (int ic = 0; ic < o.cin; ++ic)
            for (int t = 0; t < K * K; ++t) {
                const float v = wb[o.w_off + ((int64_t)(L.co0 + part * L.NOUT + oc) * o.cin + ic) * K * K + t];
                if (!fp16_exact(v))
                    return fail(B2SR_E_UNSUPPORTED, "weight %g (conv out %d in %d tap %d) is not exactly representable in fp16", v, L.co0 + oc, ic, t);
                const __half hv = __float2half_rn(v);
                const int ky = K == 3 ? t / 3 : 1, kx = K == 3 ? t % 3 : 1;
                const int g = ic / 64, c = ic % 64;
                const uint32_t a = swizzle_addr((uint32_t)(((2 - ky) * L.NOUT + oc) * PB + c * 2), PB);
                memcpy(&img[part * main_bytes + ((size_t)g * 3 + kx) * 3 * L.NOUT * PB + a], &hv, 2);
                const uint32_t af = swizzle_addr((uint32_t)((ky * L.NOUT + oc) * PB + c * 2), PB);  // bottom-up walk: ky <-> 2 - ky
                memcpy(&flp[part * main_bytes + ((size_t)g * 3 + kx) * 3 * L.NOUT * PB + af], &hv, 2);
            }
    std::vector<float> b(parts * L.NOUT, 0.f), s(parts * L.NOUT, 1.f);
    for (int part = 0; part < parts; ++part)
        for (int oc = 0; oc < L.nco; ++oc) {
            if (o.b_off >= 0) b[part * L.NOUT + oc] = wb[o.b_off + L.co0 + part * L.NOUT + oc];
            if (o.act == 2) s[part * L.NOUT + oc] = o.slope;
        }
    CUDA_TRY(cudaMalloc(&L.wimg, img.size()));
    CUDA_TRY(cudaMemcpy(L.wimg, img.data(), img.size(), cudaMemcpyHostToDevice));
    if (with_flip) {
        CUDA_TRY(cudaMalloc(&L.wimg_flip, flp.size()));
        CUDA_TRY(cudaMemcpy(L.wimg_flip, flp.data(), flp.size(), cudaMemcpyHostToDevice));
    }
    CUDA_TRY(cudaMalloc(&L.bias, b.size() * 4));
    CUDA_TRY(cudaMemcpy(L.bias, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&L.slope, s.size() * 4));
    CUDA_TRY(cudaMemcpy(L.slope, s.data(), s.size() * 4, cudaMemcpyHostToDevice));
    return 0;
}


// ------------------------------------------------------------------------------------------------
// pipelined segments of the fused program (tcg_pipe_kernel): analysis at create time
// ------------------------------------------------------------------------------------------------
typedef std::vector<std::pair<int, int>> Ranges;  // channel ranges [lo, hi)
static bool rng_overlaps(const Ranges& R, int lo, int hi) {
    for (auto& r : R)
        if (lo < r.second && r.first < hi) return true;
    return false;
}
static bool rng_covers(const Ranges& R, int lo, int hi) {
    int pos = lo;
    for (bool moved = true; pos < hi && moved;) {
        moved = false;
        for (auto& r : R)
            if (r.first <= pos && r.second > pos) pos = r.second, moved = true;
    }
    return pos >= hi;
}

static bool seg_candidate(const b2sr_ctx* c, int j) {
    const b2sr_fused_op& o = c->fops[j];
    if (o.type != B2SR_FOP_CONV || o.res != 1 || o.k != 3 || o.final || o.in_buf < 0 || o.cin > 192) return false;
    if (c->fop_first[j] == c->fop_first[j + 1]) return false;
    for (int li = c->fop_first[j]; li < c->fop_first[j + 1]; ++li)
        if (c->flaunch[li].NOUT != 32) return false;
    return true;
}
static int seg_op_stages(const b2sr_ctx* c, int j) {
    int n = 0;
    for (int li = c->fop_first[j]; li < c->fop_first[j + 1]; ++li) n += c->flaunch[li].pair ? 2 : 1;
    return n;
}

// Who reads the value op j writes into channels [lo, hi) of buffer `buf` before it is overwritten: ops inside the segment
// (<= seg_end) and / or after it.
static void value_readers(const b2sr_ctx* c, int j, int buf, int lo, int hi, int seg_end, bool* inside, bool* outside) {
    *inside = *outside = false;
    for (int k = j + 1; k < (int)c->fops.size(); ++k) {
        const b2sr_fused_op& o = c->fops[k];
        bool reads = o.in_buf == buf && o.in_off < hi && lo < o.in_off + o.cin;
        if (o.type == B2SR_FOP_CONV)
            for (int q = 0; q < o.nres; ++q) reads = reads || (o.res_buf[q] == buf && o.res_off[q] < hi && lo < o.res_off[q] + o.cout);
        if (reads) (k <= seg_end ? *inside : *outside) = true;
        const bool kills = (o.out16_buf == buf && o.out16_off <= lo && hi <= o.out16_off + o.cout) ||
                           (o.out32_buf == buf && o.out32_off <= lo && hi <= o.out32_off + o.cout);
        if (kills) break;
    }
}

// Decides, for the run of convolutions [ob, oe], which buffer slices become rings and how the stages gate one another.
// Returns false when the run does not have the shape the persistent kernel supports (it then runs launch by launch).
// The flow control uses ONE counter per (stage, band) and waits on one op only, which is sound when the ops form a chain:
//  * input gating waits for the op p* that wrote the newest ring slice of the input view; every other ring slice the stage
//    reads (input view or residual) must come from p* or from an op p* itself (transitively) reads through its ring inputs --
//    then "p* published row y + 1 on bands b-1..b+1" implies those rows exist too;
//  * a ring slot is reused once the LAST op that reads the ring has finished with the row; every other reader must be an op
//    that last reader (transitively) depends on.
static bool analyse_segment(const b2sr_ctx* c, int ob, int oe, FusedSegment* S) {
    const int nbuf = (int)c->fbufs.size();
    std::vector<int> cur(nbuf, -1);
    std::vector<Ranges> frame_read(nbuf), frame_written(nbuf), pre_touch(nbuf);
    std::map<int, std::vector<int>> dep;                 // op -> ops it (transitively) reads through ring conv inputs
    std::vector<std::vector<int>> inst_readers;          // per ring instance: reading ops
    auto in_dep = [&](int op, int q) { return std::find(dep[op].begin(), dep[op].end(), q) != dep[op].end(); };
    *S = FusedSegment();
    S->op_begin = ob, S->op_end = oe;
    for (int j = ob; j <= oe; ++j) {
        const b2sr_fused_op& o = c->fops[j];
        S->op_first_stage.push_back((int)S->stages.size());
        std::vector<int> prod_in, prod_res;
        auto classify = [&](int buf, int lo, int hi, int* inst, std::vector<int>* producers) -> int {  // 1 ring, 0 frame, -1 unsupported mix
            const int ci = cur[buf];
            if (ci >= 0 && rng_covers(S->inst[ci].written, lo, hi)) {
                *inst = ci;
                for (size_t w = 0; w < S->inst[ci].written.size(); ++w)
                    if (lo < S->inst[ci].written[w].second && S->inst[ci].written[w].first < hi) producers->push_back(S->inst[ci].writer_op[w]);
                S->inst[ci].touched.push_back({lo, hi});
                S->inst[ci].last_reader = j;
                inst_readers[ci].push_back(j);
                return 1;
            }
            if (ci >= 0 && rng_overlaps(S->inst[ci].written, lo, hi)) return -1;
            frame_read[buf].push_back({lo, hi});
            (ci >= 0 ? S->inst[ci].touched : pre_touch[buf]).push_back({lo, hi});
            return 0;
        };
        FusedStage proto;
        proto.op = j;
        const int G = (o.cin + 63) / 64;
        if (G > 3) return false;
        for (int g = 0; g < G; ++g) {
            int inst = -1;
            const int cls = classify(o.in_buf, o.in_off + 64 * g, o.in_off + std::min(64 * g + 64, o.cin), &inst, &prod_in);
            if (cls < 0) return false;
            proto.grp_ring[g] = cls;
            if (cls == 1) {
                if (proto.in_inst >= 0 && proto.in_inst != inst) return false;
                proto.in_inst = inst;
            }
        }
        for (int q = 0; q < o.nres; ++q) {
            int inst = -1;
            const int cls = classify(o.res_buf[q], o.res_off[q], o.res_off[q] + o.cout, &inst, &prod_res);
            if (cls < 0) return false;
            proto.res_inst[q] = cls == 1 ? inst : -1;
        }
        if (!prod_in.empty()) {
            const int pstar = *std::max_element(prod_in.begin(), prod_in.end());
            for (const std::vector<int>* P : {&prod_in, &prod_res})
                for (int q : *P)
                    if (q != pstar && !in_dep(pstar, q)) return false;
            proto.gate_op = pstar;
            std::vector<int>& d = dep[j];
            for (int q : prod_in) {
                d.push_back(q);
                d.insert(d.end(), dep[q].begin(), dep[q].end());
            }
            std::sort(d.begin(), d.end());
            d.erase(std::unique(d.begin(), d.end()), d.end());
        } else if (!prod_res.empty()) {
            return false;  // a ring residual without a ring input to gate on
        }
        // writes
        auto place = [&](int buf, int lo, int hi, int* inst_out) -> bool {
            bool inside = false, outside = false;
            value_readers(c, j, buf, lo, hi, oe, &inside, &outside);
            if (outside) {
                if (inside) return false;  // would need both a ring and a frame copy
                frame_written[buf].push_back({lo, hi});
                *inst_out = -1;
                return true;
            }
            int ci = cur[buf];
            if (ci < 0 || rng_overlaps(S->inst[ci].touched, lo, hi)) {
                RingInst R;
                R.buf = buf;
                if (ci < 0) R.touched = pre_touch[buf];
                if (rng_overlaps(R.touched, lo, hi)) R.touched.clear();  // (a fresh instance: the frame copy keeps serving the earlier readers)
                S->inst.push_back(R);
                inst_readers.push_back({});
                ci = cur[buf] = (int)S->inst.size() - 1;
            }
            S->inst[ci].written.push_back({lo, hi});
            S->inst[ci].touched.push_back({lo, hi});
            S->inst[ci].writer_op.push_back(j);
            *inst_out = ci;
            return true;
        };
        if (o.out16_buf >= 0 && !place(o.out16_buf, o.out16_off, o.out16_off + o.cout, &proto.out16_inst)) return false;
        if (o.out32_buf >= 0 && !place(o.out32_buf, o.out32_off, o.out32_off + o.cout, &proto.out32_inst)) return false;
        // epilogue variant (the same template instances the per-launch path picks)
        const int outs = (o.out16_buf >= 0 ? 1 : 0) | (o.out32_buf >= 0 ? 2 : 0);
        bool resf32 = true, resf16 = true;
        for (int q = 0; q < o.nres; ++q) {
            resf32 = resf32 && c->fbufs[o.res_buf[q]].dtype == 4;
            resf16 = resf16 && c->fbufs[o.res_buf[q]].dtype == 2;
        }
        if (o.sc_cin) proto.variant = B2SR_TCG_VARIANT_SC;
        else if (o.nres == 0 && outs == 1) proto.variant = B2SR_TCG_VARIANT_PLAIN;
        else if (o.nres == 1 && outs == 1 && resf16) proto.variant = B2SR_TCG_VARIANT_R16;
        else if (o.nres == 1 && outs == 3 && resf32) proto.variant = B2SR_TCG_VARIANT_R1_O3;
        else if (o.nres == 2 && outs == 3 && resf32) proto.variant = B2SR_TCG_VARIANT_R2_O3;
        else proto.variant = B2SR_TCG_VARIANT_GENERIC;
        for (int li = c->fop_first[j]; li < c->fop_first[j + 1]; ++li)
            for (int half = 0; half < (c->flaunch[li].pair ? 2 : 1); ++half) {
                FusedStage st = proto;
                st.launch = li, st.half = half;
                S->stages.push_back(st);
            }
    }
    S->op_first_stage.push_back((int)S->stages.size());
    for (int b = 0; b < nbuf; ++b)
        for (auto& w : frame_written[b])
            if (rng_overlaps(frame_read[b], w.first, w.second)) return false;  // a stage would read a frame slice another stage is writing
    if (S->inst.empty()) return false;  // nothing stays on chip: no point
    for (size_t i = 0; i < S->inst.size(); ++i) {
        const int rk = S->inst[i].last_reader;
        for (int q : inst_readers[i])
            if (q != rk && !in_dep(rk, q)) return false;
    }
    for (FusedStage& st : S->stages) {
        const int a = st.out16_inst >= 0 ? S->inst[st.out16_inst].last_reader : -1, b = st.out32_inst >= 0 ? S->inst[st.out32_inst].last_reader : -1;
        if (a >= 0 && b >= 0 && a != b) return false;
        st.bp_op = a >= 0 ? a : b;
    }
    return true;
}

static void build_segments(b2sr_ctx* c) {
    c->fsegs.clear();
    c->fseg_stages = 0;
    const char* e = getenv("B2SR_SEG_PIPE");
    if (e) c->seg_pipe = atoi(e) != 0;  // (B2SR_OPT_SEG_PIPE switches it at run time)
    const int n_ops = (int)c->fops.size(), max_stages = c->sms / 8;  // planes of the reference's tiling are <= 980 px = 8 bands wide
    for (int i = 0; i < n_ops;) {
        if (!seg_candidate(c, i)) {
            ++i;
            continue;
        }
        int j = i, stages = 0;
        while (j < n_ops && seg_candidate(c, j) && stages + seg_op_stages(c, j) <= max_stages) stages += seg_op_stages(c, j), ++j;
        FusedSegment S;
        if (j - i >= 2 && analyse_segment(c, i, j - 1, &S)) {
            S.stage_base = c->fseg_stages;
            c->fseg_stages += (int)S.stages.size();
            c->fsegs.push_back(S);
        }
        i = std::max(j, i + 1);
    }
}

// How every convolution of the fused program is launched (output-channel padding, halves, ring slots); with `wb` the
// weight images are uploaded as well (without: host-only planning, used by b2sr_fused_describe_segments).
static int plan_fused_launches(b2sr_ctx* c, const float* wb) {
    const int n_ops = (int)c->fops.size();
    int rc = 0;
    for (int i = 0; i < n_ops && !rc; ++i) {
        const b2sr_fused_op& o = c->fops[i];
        c->fop_first.push_back((int)c->flaunch.size());
        if (o.type != B2SR_FOP_CONV) continue;
        const int cinp = o.in_buf < 0 ? 16 : o.cin, G = (cinp + 63) / 64;
        int NOUT = o.final ? 16 : o.cout, parts = 1;
        // the stacked weights of all groups stay resident in shared memory beside >= 4 ring slots; a wide
        // convolution that does not fit (192 -> 64: 221 KB) is launched as two halves of 32 output channels
        const bool sc = o.sc_cin != 0;
        if (fused_fit(NOUT, o.final != 0, G, sc) < 4) {
            if (NOUT == 64 && fused_fit(32, false, G, false) >= 4) {
                NOUT = 32, parts = 2;
            } else {
                rc = fail(B2SR_E_UNSUPPORTED, "op %d: %d -> %d convolution does not fit shared memory", i, o.cin, o.cout);
                break;
            }
        }
        const bool pair = parts == 2 && c->pair_halves && o.cout == 2 * NOUT;
        if (pair) parts = 1;
        for (int part = 0; part < parts && !rc; ++part) {
            FusedLaunch L;
            L.pair = pair;
            L.op = i, L.co0 = part * NOUT, L.nco = std::min(o.cout - L.co0, NOUT), L.NOUT = NOUT, L.G = G, L.cinp = cinp;
            L.sc_ks = o.sc_cin / 16;
            L.slots = std::min(12, fused_fit(NOUT, o.final != 0, G, sc));
            if (wb) rc = upload_fused_launch(L, o, wb, c->flip_rows != 0);
            c->flaunch.push_back(L);  // (pushed even on failure so that b2sr_destroy frees what was allocated)
        }
    }
    c->fop_first.push_back((int)c->flaunch.size());
    // CTA-pair form: the whole convolution (all output channels) as one launch of 2-CTA clusters over band pairs; every CTA
    // keeps half of the stacked weights, so even the 192 -> 64 convolution (221 KB) fits
    c->fop_pair2.assign(n_ops, -1);
    for (int i = 0; i < n_ops && !rc; ++i) {
        const b2sr_fused_op& o = c->fops[i];
        if (o.type != B2SR_FOP_CONV || o.k != 3 || o.final || o.in_buf < 0 || (o.cout != 32 && o.cout != 64)) continue;
        const int G = (o.cin + 63) / 64;
        const bool sc = o.sc_cin != 0;
        const int fit = sc ? TcgCfg<32, 0, true>::ring_fit2(G) : (o.cout == 64 ? TcgCfg<64, 0>::ring_fit2(G) : TcgCfg<32, 0>::ring_fit2(G));
        if (fit < 4 || (sc && o.cout != 32)) continue;
        FusedLaunch L;
        L.op = i, L.co0 = 0, L.nco = o.cout, L.NOUT = o.cout, L.G = G, L.cinp = o.cin, L.pair = 0;
        L.sc_ks = o.sc_cin / 16;
        L.slots = std::min(12, fit);
        if (wb) rc = upload_fused_launch(L, o, wb, false);
        c->fop_pair2[i] = (int)c->flaunch2.size();
        c->flaunch2.push_back(L);
    }
    return rc;
}

// Host-only (no device needed): how the fused program would be cut into persistent segments on a device with `sms` SMs.
// out[0] = number of segments, then per segment: op_begin, op_end, n_stages, n_ring_instances, followed by n_stages
// records of 12 ints {op, half, variant, in_inst, grp_ring[0..2], out16_inst, out32_inst, res_inst[0..1], gate_op} and one
// more int per stage {bp_op}; then per ring instance {buffer, last_reader}.  Returns the number of ints needed (<= cap
// were written), or a negative error.
extern "C" int b2sr_fused_describe_segments(const b2sr_fused_op* ops, int n_ops, const b2sr_fused_buf* bufs, int n_bufs, int sms,
                                            int32_t* out, int cap) {
    if (!ops || !bufs || n_ops < 1 || n_bufs < 1 || sms < 1) return fail(B2SR_E_INVALID, "b2sr_fused_describe_segments: bad argument");
    b2sr_ctx c;
    c.sms = sms, c.family = B2SR_FAMILY_FUSED;
    c.fops.assign(ops, ops + n_ops);
    c.fbufs.assign(bufs, bufs + n_bufs);
    TRY(plan_fused_launches(&c, nullptr));
    build_segments(&c);
    std::vector<int32_t> v;
    v.push_back((int32_t)c.fsegs.size());
    for (const FusedSegment& S : c.fsegs) {
        v.insert(v.end(), {S.op_begin, S.op_end, (int32_t)S.stages.size(), (int32_t)S.inst.size()});
        for (const FusedStage& st : S.stages)
            v.insert(v.end(), {st.op, st.half, st.variant, st.in_inst, st.grp_ring[0], st.grp_ring[1], st.grp_ring[2], st.out16_inst,
                               st.out32_inst, st.res_inst[0], st.res_inst[1], st.gate_op, st.bp_op});
        for (const RingInst& R : S.inst) v.insert(v.end(), {R.buf, R.last_reader});
    }
    if (out)
        for (int i = 0; i < cap && i < (int)v.size(); ++i) out[i] = v[i];
    return (int)v.size();
}

extern "C" int b2sr_create_fused(b2sr_ctx** out, int device, const b2sr_fused_op* ops, int n_ops, const b2sr_fused_buf* bufs,
                                 int n_bufs, int scale, const void* weights, size_t nbytes) {
    if (!out || !ops || !bufs || !weights) return fail(B2SR_E_INVALID, "b2sr_create_fused: null argument");
    *out = nullptr;
    if (n_ops < 1 || n_ops > 100000 || n_bufs < 1 || n_bufs > 4096) return fail(B2SR_E_INVALID, "b2sr_create_fused: bad program size");
    if (scale != 1 && scale != 2 && scale != 4) return fail(B2SR_E_UNSUPPORTED, "scale = %d (need 1, 2 or 4)", scale);
    for (int i = 0; i < n_bufs; ++i)
        if ((bufs[i].dtype != 2 && bufs[i].dtype != 4) || bufs[i].channels < 8 || bufs[i].channels % 8 || bufs[i].res < 1 || bufs[i].res > 8)
            return fail(B2SR_E_INVALID, "buffer %d: %d channels, dtype %d, res %d", i, bufs[i].channels, bufs[i].dtype, bufs[i].res);
    const float* wb = (const float*)weights;
    const int64_t nfl = (int64_t)(nbytes / 4);
    // validate the program before touching the device
    for (int i = 0; i < n_ops; ++i) {
        const b2sr_fused_op& o = ops[i];
        auto view_ok = [&](int b, int off, int c, int res, int dtype /*0 = any*/) {
            return b >= 0 && b < n_bufs && off >= 0 && off % 8 == 0 && off + c <= bufs[b].channels && bufs[b].res == res &&
                   (dtype == 0 || bufs[b].dtype == dtype);
        };
        if (o.res < 1 || o.res > 8) return fail(B2SR_E_INVALID, "op %d: res %d", i, o.res);
        if (o.type == B2SR_FOP_NEAREST) {
            if (o.r < 2 || o.r > 4 || o.res % o.r || o.cin % 8 || o.cout != o.cin || !view_ok(o.in_buf, o.in_off, o.cin, o.res / o.r, 2) ||
                !view_ok(o.out16_buf, o.out16_off, o.cout, o.res, 2) || o.out32_buf >= 0 || o.nres || o.final)
                return fail(B2SR_E_INVALID, "op %d: bad nearest-neighbour resize", i);
            continue;
        }
        if (o.type != B2SR_FOP_CONV) return fail(B2SR_E_UNSUPPORTED, "op %d: unknown type %d", i, o.type);
        if ((o.k != 1 && o.k != 3) || (o.act != 0 && o.act != 2) || o.nres < 0 || o.nres > 2)
            return fail(B2SR_E_UNSUPPORTED, "op %d: convolution k=%d act=%d nres=%d", i, o.k, o.act, o.nres);
        if (o.in_buf < 0 ? (o.cin != 3 || o.res != 1) : (o.cin % 16 || o.cin > 192 || !view_ok(o.in_buf, o.in_off, o.cin, o.res, 2)))
            return fail(B2SR_E_UNSUPPORTED, "op %d: convolution input (buffer %d, offset %d, %d channels)", i, o.in_buf, o.in_off, o.cin);
        if (o.final ? (o.cout != 3 || o.nres || o.out16_buf >= 0 || o.out32_buf >= 0 || i != n_ops - 1 || o.res != scale)
                    : ((o.cout != 32 && o.cout != 64) || (o.out16_buf < 0 && o.out32_buf < 0)))
            return fail(B2SR_E_UNSUPPORTED, "op %d: convolution output (%d channels, final %d)", i, o.cout, o.final);
        if ((o.out16_buf >= 0 && !view_ok(o.out16_buf, o.out16_off, o.cout, o.res, 2)) ||
            (o.out32_buf >= 0 && !view_ok(o.out32_buf, o.out32_off, o.cout, o.res, 4)))
            return fail(B2SR_E_INVALID, "op %d: bad output view", i);
        for (int q = 0; q < o.nres; ++q)
            if (!view_ok(o.res_buf[q], o.res_off[q], o.cout, o.res, 0)) return fail(B2SR_E_INVALID, "op %d: bad residual %d", i, q);
        if (o.sc_cin) {
            if (o.sc_cin < 16 || o.sc_cin > 64 || o.sc_cin % 16 || o.sc_cin > o.cin || o.in_buf < 0 || o.cout != 32 || o.k != 3 || o.nres || o.out32_buf >= 0 ||
                o.out16_buf < 0 || o.final)
                return fail(B2SR_E_UNSUPPORTED, "op %d: fused 1x1 shortcut over %d channels (needs k = 3, cout = 32, fp16 output only, no residual terms)", i, o.sc_cin);
            if (o.sc_w_off < 0 || o.sc_w_off + (int64_t)o.cout * o.sc_cin > nfl) return fail(B2SR_E_INVALID, "op %d: shortcut weights outside the blob", i);
        }
        if (o.sc_cin) {
            if (o.sc_cin < 16 || o.sc_cin > 64 || o.sc_cin % 16 || o.sc_cin > o.cin || o.in_buf < 0 || o.cout != 32 || o.k != 3 || o.nres || o.out32_buf >= 0 ||
                o.out16_buf < 0 || o.final)
                return fail(B2SR_E_UNSUPPORTED, "op %d: fused 1x1 shortcut over %d channels (needs k = 3, cout = 32, fp16 output only, no residual terms)", i, o.sc_cin);
            if (o.sc_w_off < 0 || o.sc_w_off + (int64_t)o.cout * o.sc_cin > nfl) return fail(B2SR_E_INVALID, "op %d: shortcut weights outside the blob", i);
        }
        const int64_t nw = (int64_t)o.cout * o.cin * o.k * o.k;
        if (o.w_off < 0 || o.w_off + nw > nfl || (o.b_off >= 0 && o.b_off + o.cout > nfl)) return fail(B2SR_E_INVALID, "op %d: weights outside the blob", i);
    }
    if (!ops[n_ops - 1].final) return fail(B2SR_E_INVALID, "the last op must produce the network output");
    cudaDeviceProp prop;
    TRY(check_device(device, &prop));
    TRY(get_encode());
    b2sr_ctx* c = new b2sr_ctx();
    c->device = device, c->sms = prop.multiProcessorCount, c->family = B2SR_FAMILY_FUSED;
    c->desc.family = B2SR_FAMILY_FUSED, c->desc.cin = 3, c->desc.scale = scale;
    c->CF = 64, c->NL = 16;
    {
        // Experiment kept behind B2SR_L2_PERSIST=<MB> (default off): mark a fraction of each launch's input buffer as
        // L2-persisting so that the following launches of the dense block (which re-read it) hit L2.  Measured on
        // B200, 540p (199 MB per 192-channel buffer): 34.4 ms/frame without, 38.3 ms with 48 MB set aside, 59 ms with
        // the maximum -- the set-aside starves the normal L2 traffic (halo re-reads, weights, fp32 trunk), so it is off.
        const char* ce = getenv("B2SR_PAIR");
        if (ce && atoi(ce) == 0) c->pair_halves = 0;
        const char* c2 = getenv("B2SR_PAIR2");
        if (c2) c->pair2 = atoi(c2) != 0;
        const char* pe = getenv("B2SR_PDL");
        if (pe && atoi(pe) == 0) c->pdl = 0;
        const char* le = getenv("B2SR_L2_PROMO");
        if (le) c->l2_promo = atoi(le) == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : (atoi(le) == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : (atoi(le) == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B));
        const char* fe = getenv("B2SR_FLIP");
        if (fe && atoi(fe) != 0) c->flip_rows = 1;
        const char* e = getenv("B2SR_L2_PERSIST");
        if (e && atoi(e) > 0 && prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0) {
            const size_t want = (size_t)atoi(e) << 20;
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min(want, (size_t)prop.persistingL2CacheMaxSize)) == cudaSuccess) {
                c->l2_persist = std::min(want, (size_t)prop.persistingL2CacheMaxSize);
                c->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
            } else {
                cudaGetLastError();
            }
        }
    }
    c->fops.assign(ops, ops + n_ops);
    c->fbufs.assign(bufs, bufs + n_bufs);
    c->fbuf_ptr.assign(n_bufs, nullptr);
    c->fbuf_cap.assign(n_bufs, 0);
    int rc = 0;
    do {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking) != cudaSuccess) {
            rc = fail(B2SR_E_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        rc = plan_fused_launches(c, wb);
        if (!rc) build_segments(c);
    } while (0);
    if (rc) {
        std::string keep = g_err;
        b2sr_destroy(c);
        g_err = keep;
        return rc;
    }
    *out = c;
    return 0;
}

// work items of the plan's planes at resolution factor `res`: the linear sequence (plane, band, row) cut into one
// contiguous, equally long range per CTA (as build_plan does for res = 1)
static int fused_items(b2sr_ctx* c, Plan* P, int res, int max_cta, ResItems** out, int pairs = 0) {
    for (auto& r : P->res_items)
        if (r->res == res && r->max_cta == max_cta && r->pairs == pairs) {
            *out = r.get();
            return 0;
        }
    std::unique_ptr<ResItems> R(new ResItems());
    R->res = res, R->max_cta = max_cta, R->pairs = pairs;
    const int BW = pairs ? 2 * TC_BW : TC_BW;  // columns per item
    int64_t band_rows = 0;
    for (const PlaneDev& pd : P->planes) band_rows += (int64_t)pd.Ht * res * ((pd.Wt * res + BW - 1) / BW);
    const int ncta = (int)std::min<int64_t>(max_cta, band_rows);
    R->first.assign(1, 0);
    int64_t pos = 0;
    int cta = 0;
    auto cut_at = [&](int k) { return band_rows * k / ncta; };
    for (size_t pi = 0; pi < P->planes.size(); ++pi) {
        const PlaneDev& pd = P->planes[pi];
        const int Ht = pd.Ht * res, Wt = pd.Wt * res;
        for (int x0 = 0; x0 < Wt; x0 += BW) {
            int y0 = 0;
            while (y0 < Ht) {
                const int64_t room = cut_at(cta + 1) - pos;
                const int rows = (int)std::min<int64_t>(Ht - y0, room);
                TcItem it{};
                it.map = P->plane_group[pi], it.plane = pd.gplane, it.x0 = x0, it.y0 = y0;
                it.rows = rows, it.w = std::min(BW, Wt - x0);
                it.Ht = Ht, it.Wt = Wt, it.pix_off = pd.pix_off * res * res;
                it.frame = pd.frame, it.fy0 = pd.fy0 * res, it.fx0 = pd.fx0 * res;
                it.cy0 = pd.cy0 * res, it.cy1 = pd.cy1 * res, it.cx0 = pd.cx0 * res, it.cx1 = pd.cx1 * res;
                R->items.push_back(it);
                R->out_px += (double)it.rows * it.w;
                y0 += rows, pos += rows;
                if (pos == cut_at(cta + 1)) {
                    ++cta;
                    R->first.push_back((int)R->items.size());
                }
            }
        }
    }
    R->n_cta = ncta;
    CUDA_TRY(cudaMalloc(&R->d_items, R->items.size() * sizeof(TcItem)));
    CUDA_TRY(cudaMemcpy(R->d_items, R->items.data(), R->items.size() * sizeof(TcItem), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&R->d_first, R->first.size() * sizeof(int)));
    CUDA_TRY(cudaMemcpy(R->d_first, R->first.data(), R->first.size() * sizeof(int), cudaMemcpyHostToDevice));
    *out = R.get();
    P->res_items.push_back(std::move(R));
    return 0;
}

template <int NOUT, int MODE, bool F32OUT, int NRES, int OUTS, bool RF16 = false, bool SC = false>
static int launch_tcg(b2sr_ctx* c, const FusedLaunch& L, const ResItems* R, const TcgParams& p) {
    auto kern = tcg_conv_kernel<NOUT, MODE, F32OUT, NRES, OUTS, RF16, SC>;
    const int smem = TcgCfg<NOUT, MODE, SC>::smem_bytes(L.G, L.slots);
    TRY(raise_dyn_smem((const void*)kern, smem));
    // programmatic dependent launch: this launch's prologue (barrier init, TMEM allocation, weight load) overlaps the tail
    // of the previous launch of the stream; the kernel waits (griddepcontrol.wait) before it touches activation buffers
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(R->n_cta * (L.pair ? 2 : 1))), cfg.blockDim = dim3(TC_THREADS), cfg.dynamicSmemBytes = (size_t)smem, cfg.stream = c->stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (L.pair) {  // clusters of two CTAs: one TMA multicast feeds both halves of the convolution
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 2, attr[na].val.clusterDim.y = 1, attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (c->pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr, cfg.numAttrs = na;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, p));
    c->n_launch += 1, c->n_tc += 1;
    return 0;
}

// Parameters of launch `li` of the fused program that do not depend on how the work is cut into CTAs.
static void fused_fill_params(b2sr_ctx* c, Plan* P, int li, void* d_out, TcgParams& p) {
    const FusedLaunch& L = c->flaunch[li];
    const b2sr_fused_op& o = c->fops[L.op];
    const int G = (int)P->groups.size();
    p.maps = P->d_fmaps, p.map_base = li * G;
    p.wimg = L.wimg, p.bias = L.bias, p.slope = L.slope;
    p.acc_scale = o.in_buf < 0 ? (1.f / 255.f) : 1.f;
    p.groups = L.G, p.cin = L.cinp, p.k1 = o.k == 1, p.ring_slots = L.slots;
    p.sc_ks = L.sc_ks, p.sc_cv = o.sc_coef_v, p.sc_cr = o.sc_coef_r;
    static const int ablate_env = getenv("B2SR_ABLATE") ? atoi(getenv("B2SR_ABLATE")) : 0;  // measurement only: see TcgParams::ablate
    p.ablate = c->ablate ? c->ablate : ablate_env;
    p.pair = L.pair, p.pair_wbytes = L.G * 9 * L.NOUT * TCG_PB;
    p.nres = o.nres;
    for (int q = 0; q < o.nres; ++q) {
        const b2sr_fused_buf& B = c->fbufs[o.res_buf[q]];
        p.res_ptr[q] = (const uint8_t*)c->fbuf_ptr[o.res_buf[q]] + (size_t)(o.res_off[q] + L.co0) * B.dtype;
        p.res_ld[q] = B.channels, p.res_f32[q] = B.dtype == 4;
        p.coef_v[q] = o.coef_v[q], p.coef_r[q] = o.coef_r[q];
    }
    if (o.out16_buf >= 0) {
        p.out16 = (__half*)c->fbuf_ptr[o.out16_buf] + o.out16_off + L.co0;
        p.out16_ld = c->fbufs[o.out16_buf].channels;
    }
    if (o.out32_buf >= 0) {
        p.out32 = (float*)c->fbuf_ptr[o.out32_buf] + o.out32_off + L.co0;
        p.out32_ld = c->fbufs[o.out32_buf].channels;
    }
    p.frames_out = d_out, p.frame_h = P->h * o.res, p.frame_w = P->w * o.res;
}


// Parameters of the CTA-pair form of op i: the op's first ordinary launch supplies the input maps (same input view).
static void fused_fill_params2(b2sr_ctx* c, Plan* P, int i, const FusedLaunch& L, TcgParams& p) {
    fused_fill_params(c, P, c->fop_first[i], nullptr, p);  // maps, residuals, outputs (channel offset 0 of the op)
    p.wimg = L.wimg, p.bias = L.bias, p.slope = L.slope;
    p.groups = L.G, p.cin = L.cinp, p.ring_slots = L.slots, p.sc_ks = L.sc_ks;
    p.pair = 0, p.pair_wbytes = 0, p.flip = 0;
}

// Ring arena layout, ring tensor maps and the stage table of every pipelined segment, for the planes of P.
static int prepare_segments(b2sr_ctx* c, Plan* P) {
    const int G = (int)P->groups.size(), nb = P->nb;
    const int RR = c->ring_rows > 0 ? std::max(8, c->ring_rows) : 24;
    // Every ring: RR rows x Wmax pixels x the buffer's channels.  24 rows: an RRDB segment then holds 3 x 8.9 MB of dense
    // block + 2 x 6 MB of fp32 trunk (960-px planes) -- inside what the L2 keeps without spilling.
    size_t need = 0;
    for (FusedSegment& S : c->fsegs) {
        size_t off = 0;
        for (RingInst& R : S.inst) {
            const b2sr_fused_buf& B = c->fbufs[R.buf];
            R.offset = off;
            off += ((size_t)RR * P->Wmax * B.channels * B.dtype + 1023) / 1024 * 1024;
        }
        need = std::max(need, off);
    }
    if (need > c->cap_frings) {
        cudaStreamSynchronize(c->stream);
        if (c->frings) cudaFree(c->frings);
        c->frings = nullptr, c->cap_frings = 0;
        CUDA_TRY(cudaMalloc(&c->frings, need));
        c->cap_frings = need;
        c->fring_gen += 1;
    }
    if (c->pipe_debug && !c->d_dbg) CUDA_TRY(cudaMalloc(&c->d_dbg, B2SR_DBG_WORDS * sizeof(long long)));
    const uint64_t key = (c->fbuf_gen << 24) ^ (c->fring_gen << 8) ^ (uint64_t)RR ^ ((uint64_t)P->Wmax << 44) ^ ((uint64_t)(c->pipe_debug != 0) << 60);
    if (P->fstages_key == key && P->d_fstages) return 0;
    int seg_max = 0;
    for (const FusedSegment& S : c->fsegs) seg_max = std::max(seg_max, (int)S.stages.size());
    if (!P->d_fstages) CUDA_TRY(cudaMalloc(&P->d_fstages, (size_t)c->fseg_stages * sizeof(TcgParams)));
    if (!P->d_fflags) CUDA_TRY(cudaMalloc(&P->d_fflags, (size_t)seg_max * nb * B2SR_FLAG_STRIDE * sizeof(uint32_t)));
    std::vector<CUtensorMap> rmaps((size_t)c->fseg_stages * G);
    memset(rmaps.data(), 0, rmaps.size() * sizeof(CUtensorMap));
    std::vector<TcgParams> table(c->fseg_stages);
    const size_t fl = (size_t)nb * B2SR_FLAG_STRIDE;  // counter words per stage
    for (const FusedSegment& S : c->fsegs)
        for (size_t t = 0; t < S.stages.size(); ++t) {
            const FusedStage& st = S.stages[t];
            const FusedLaunch& L = c->flaunch[st.launch];
            const b2sr_fused_op& o = c->fops[st.op];
            const int sg = S.stage_base + (int)t;
            TcgParams& p = table[sg];
            memset(&p, 0, sizeof p);
            fused_fill_params(c, P, st.launch, nullptr, p);
            p.items = P->d_pitems, p.item_first = P->d_pband_first;
            p.pair = 0, p.flip = 0, p.dbg = c->pipe_debug ? c->d_dbg : nullptr;
            p.half = st.half, p.variant = st.variant;
            p.nogate = getenv("B2SR_SEG_NOGATE") && atoi(getenv("B2SR_SEG_NOGATE")) != 0;
            p.nb = nb, p.RR = RR, p.Wmax = P->Wmax;
            p.ring_map_base = (int)c->flaunch.size() * G + sg * G;
            if (st.in_inst >= 0) {
                const int Cb = c->fbufs[o.in_buf].channels;
                for (int g = 0; g < G; ++g) {
                    const Group& gr = P->groups[g];
                    cuuint64_t dims[4] = {(cuuint64_t)L.cinp, (cuuint64_t)gr.Wt, (cuuint64_t)RR, 1};
                    cuuint64_t strides[3] = {(cuuint64_t)Cb * 2, (cuuint64_t)P->Wmax * Cb * 2, (cuuint64_t)RR * P->Wmax * Cb * 2};
                    cuuint32_t box[4] = {64, (cuuint32_t)TC_PITCH, 1, 1};
                    cuuint32_t es[4] = {1, 1, 1, 1};
                    void* basep = (void*)((__half*)(c->frings + S.inst[st.in_inst].offset) + o.in_off);
                    CUresult e = g_encode(&rmaps[(size_t)sg * G + g], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, basep, dims, strides, box, es,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, c->l2_promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                    if (e != CUDA_SUCCESS) return fail(B2SR_E_CUDA, "cuTensorMapEncodeTiled failed (%d) for the ring of op %d group %d", (int)e, st.op, g);
                }
            }
            for (int g = 0; g < 3; ++g) p.grp_ring[g] = st.grp_ring[g];
            if (st.out16_inst >= 0) {
                p.out16 = (__half*)(c->frings + S.inst[st.out16_inst].offset) + o.out16_off + L.co0;
                p.out16_ring = 1;
            }
            if (st.out32_inst >= 0) {
                p.out32 = (float*)(c->frings + S.inst[st.out32_inst].offset) + o.out32_off + L.co0;
                p.out32_ring = 1;
            }
            for (int q = 0; q < o.nres; ++q)
                if (st.res_inst[q] >= 0) {
                    p.res_ptr[q] = c->frings + S.inst[st.res_inst[q]].offset + (size_t)(o.res_off[q] + L.co0) * c->fbufs[o.res_buf[q]].dtype;
                    p.res_ring[q] = 1;
                }
            auto op_stages = [&](int op, const uint32_t** dst) -> int {
                const int a = S.op_first_stage[op - S.op_begin], b = S.op_first_stage[op - S.op_begin + 1];
                if (b - a > 2) return -1;
                for (int k = a; k < b; ++k) dst[k - a] = P->d_fflags + (size_t)k * fl;
                return b - a;
            };
            if (st.gate_op >= 0 && (p.n_in = op_stages(st.gate_op, p.done_in)) < 0) return fail(B2SR_E_UNSUPPORTED, "segment stage waits on more than two stages");
            if (st.bp_op >= 0 && (p.n_bp = op_stages(st.bp_op, p.bp)) < 0) return fail(B2SR_E_UNSUPPORTED, "segment stage waits on more than two stages");
            p.done_out = P->d_fflags + t * fl;
        }
    CUDA_TRY(cudaMemcpyAsync(P->d_fmaps + c->flaunch.size() * (size_t)G, rmaps.data(), rmaps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(P->d_fstages, table.data(), table.size() * sizeof(TcgParams), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    P->fstages_key = key;
    return 0;
}

// One persistent cooperative launch for the convolutions of segment S (B2SR_PIPE_REFUSED: the grid cannot be co-resident).
static int launch_segment(b2sr_ctx* c, Plan* P, const FusedSegment& S) {
    const int nb = P->nb, ns = (int)S.stages.size();
    int smem = 0;
    for (const FusedStage& st : S.stages) {
        const FusedLaunch& L = c->flaunch[st.launch];
        smem = std::max(smem, L.sc_ks ? TcgCfg<32, 0, true>::smem_bytes(L.G, L.slots) : TcgCfg<32, 0>::smem_bytes(L.G, L.slots));
    }
    TRY(raise_dyn_smem((const void*)tcg_pipe_kernel, smem));
    CUDA_TRY(cudaMemsetAsync(P->d_fflags, 0, (size_t)ns * nb * B2SR_FLAG_STRIDE * sizeof(uint32_t), c->stream));
    TcgPipeParams Q{};
    Q.stages = P->d_fstages + S.stage_base, Q.n_stages = ns, Q.nb = nb;
    double px = 0;
    for (const TcItem& it : P->pitems) px += (double)it.rows * std::max(0, it.w);
    TRY(prof_begin(c, 3, px));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(ns * nb)), cfg.blockDim = dim3(TC_THREADS), cfg.dynamicSmemBytes = (size_t)smem, cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    static const bool coop = !(getenv("B2SR_COOP") && atoi(getenv("B2SR_COOP")) == 0);  // (profiling aid, see launch_pipe)
    cfg.attrs = attr, cfg.numAttrs = coop ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, tcg_pipe_kernel, Q);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorLaunchOutOfResources) {
        cudaGetLastError();
        if (c->profile) {
            c->ev_pool.push_back(c->prof.back().a), c->ev_pool.push_back(c->prof.back().b);
            c->prof.pop_back();
        }
        return B2SR_PIPE_REFUSED;
    }
    if (e != cudaSuccess) return fail(B2SR_E_CUDA, "persistent segment launch failed: %s", cudaGetErrorString(e));
    TRY(prof_end(c));
    c->n_launch += 1, c->n_tc += 1, c->n_pipe += 1;
    if (c->pipe_debug && S.stage_base == 0) {  // stall accounting of the first segment (synchronises)
        std::vector<long long> h(B2SR_DBG_WORDS);
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        CUDA_TRY(cudaMemcpy(h.data(), c->d_dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        fprintf(stderr, "b2sr segment ops %d..%d, %d stages x %d bands; per stage, mean over bands, kcycles: issuer total | issuer waits: data (full) tmem (tempty) | "
                "producer: gate wait, slot (empty) wait | epilogue warp 2: total, wait tfull, back-pressure wait, tmem ld/zero\n", S.op_begin, S.op_end, ns, nb);
        for (int t = 0; t < ns; ++t) {
            double v[16] = {0};
            for (int b = 0; b < nb; ++b)
                for (int j = 0; j < 16; ++j) v[j] += (double)h[(size_t)(t * nb + b) * 16 + j] / nb / 1e3;
            fprintf(stderr, "  stage %2d (op %3d half %d variant %d): %8.0f | %7.0f %7.0f | %7.0f %7.0f | %8.0f %7.0f %7.0f %7.0f\n", t, S.stages[t].op, S.stages[t].half,
                    S.stages[t].variant, v[0], v[1], v[2], v[11], v[4], v[6], v[5], v[12], v[10]);
        }
    }
    return 0;
}

template <int NOUT, int NRES, int OUTS, bool RF16 = false, bool SC = false>
static int launch_tcg2(b2sr_ctx* c, const FusedLaunch& L, const ResItems* R, const TcgParams& p) {
    auto kern = tcg_pair2_kernel<NOUT, NRES, OUTS, RF16, SC>;
    const int smem = TcgCfg<NOUT, 0, SC>::smem_bytes2(L.G, L.slots);
    TRY(raise_dyn_smem((const void*)kern, smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(R->n_cta * 2)), cfg.blockDim = dim3(TC_THREADS), cfg.dynamicSmemBytes = (size_t)smem, cfg.stream = c->stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2, attr[na].val.clusterDim.y = 1, attr[na].val.clusterDim.z = 1;
    ++na;
    if (c->pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr, cfg.numAttrs = na;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, p));
    c->n_launch += 1, c->n_tc += 1;
    return 0;
}

// Clusters of two 227 KB CTAs the device can hold at once (a GPC with an odd number of SMs loses one).
static int pair_cluster_count(b2sr_ctx* c) {
    if (c->pair_clusters) return c->pair_clusters;
    cudaLaunchConfig_t q{};
    q.gridDim = dim3((unsigned)c->sms / 2 * 2), q.blockDim = dim3(TC_THREADS);
    q.dynamicSmemBytes = (size_t)TcgCfg<32, 0>::smem_bytes(3, 5);
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = 2, qa[0].val.clusterDim.y = 1, qa[0].val.clusterDim.z = 1;
    q.attrs = qa, q.numAttrs = 1;
    auto kq = tcg_conv_kernel<32, 0, false, 1, 3>;
    if (raise_dyn_smem((const void*)kq, (int)q.dynamicSmemBytes) != 0) cudaGetLastError();
    int ncl = 0;
    if (cudaOccupancyMaxActiveClusters(&ncl, kq, &q) != cudaSuccess || ncl < 1) {
        cudaGetLastError();
        ncl = c->sms / 2 - 4;
    }
    c->pair_clusters = std::min(ncl, c->sms / 2);
    return c->pair_clusters;
}

// Runs ops [0, upto] (upto < 0: the whole program) for the planes of P.
static int run_fused(b2sr_ctx* c, Plan* P, const uint8_t* d_frames, void* d_out, bool f32out, int upto) {
    const int n_ops = (int)c->fops.size(), G = (int)P->groups.size();
    if (upto < 0 || upto >= n_ops) upto = n_ops - 1;
    TRY(ensure_scratch(c, P->total_px, false));
    for (size_t b = 0; b < c->fbufs.size(); ++b) {
        const b2sr_fused_buf& B = c->fbufs[b];
        const size_t need = ((size_t)P->total_px * B.res * B.res + 1024) * B.channels * B.dtype;
        if (need > c->fbuf_cap[b]) {
            cudaStreamSynchronize(c->stream);
            if (c->fbuf_ptr[b]) cudaFree(c->fbuf_ptr[b]);
            c->fbuf_ptr[b] = nullptr, c->fbuf_cap[b] = 0;
            CUDA_TRY(cudaMalloc(&c->fbuf_ptr[b], need));
            c->fbuf_cap[b] = need;
            c->fbuf_gen += 1;
        }
    }
    if (P->fmaps_gen != c->fbuf_gen) {
        // one input tensor map per (launch, plane-size group): {cin, Wt*res, Ht*res, planes} over the input view, box
        // {64 channels, 136 pixels}: channels beyond cin and pixels outside the plane read as zeros
        std::vector<CUtensorMap> maps(c->flaunch.size() * (size_t)G);
        for (size_t li = 0; li < c->flaunch.size(); ++li) {
            const FusedLaunch& L = c->flaunch[li];
            const b2sr_fused_op& o = c->fops[L.op];
            const int Cb = o.in_buf < 0 ? 16 : c->fbufs[o.in_buf].channels, r = o.res;
            const __half* bufp = o.in_buf < 0 ? c->in16 : (const __half*)c->fbuf_ptr[o.in_buf];
            for (int g = 0; g < G; ++g) {
                const Group& gr = P->groups[g];
                const cuuint64_t Wt = (cuuint64_t)gr.Wt * r, Ht = (cuuint64_t)gr.Ht * r;
                cuuint64_t dims[4] = {(cuuint64_t)L.cinp, Wt, Ht, (cuuint64_t)gr.count};
                cuuint64_t strides[3] = {(cuuint64_t)Cb * 2, Wt * Cb * 2, Ht * Wt * Cb * 2};
                cuuint32_t box[4] = {64, (cuuint32_t)TC_PITCH, 1, 1};
                cuuint32_t es[4] = {1, 1, 1, 1};
                void* basep = (void*)(bufp + ((size_t)gr.pix_base * r * r * Cb + (o.in_buf < 0 ? 0 : o.in_off)));
                CUresult e = g_encode(&maps[li * G + g], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, basep, dims, strides, box, es,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, c->l2_promo,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (e != CUDA_SUCCESS)
                    return fail(B2SR_E_CUDA, "cuTensorMapEncodeTiled failed (%d) for op %d group %d (%d ch of %d, %llux%llu)", (int)e, L.op, g,
                                L.cinp, Cb, (unsigned long long)Ht, (unsigned long long)Wt);
            }
        }
        if (!P->d_fmaps) CUDA_TRY(cudaMalloc(&P->d_fmaps, (maps.size() + (size_t)c->fseg_stages * G) * sizeof(CUtensorMap)));
        CUDA_TRY(cudaMemcpyAsync(P->d_fmaps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        P->fmaps_gen = c->fbuf_gen;
    }
    // whole program on the planes of the reference's tiling: runs of convolutions go as one persistent launch each
    int seg_max = 0;
    for (const FusedSegment& S : c->fsegs) seg_max = std::max(seg_max, (int)S.stages.size());
    bool use_segs = c->seg_pipe && !c->pipe_unavailable && !c->fsegs.empty() && !c->flip_rows &&
                    P->nb >= 1 && (int64_t)seg_max * P->nb <= usable_sms(c);
    if (use_segs) TRY(prepare_segments(c, P));
    size_t next_seg = 0;
    {
        dim3 grid((unsigned)std::min(2048, (P->max_plane_px + 255) / 256), (unsigned)P->planes.size());
        TRY(prof_begin(c, 0, 0));
        prep_kernel<<<grid, 256, 0, c->stream>>>(d_frames, P->h, P->w, P->d_planes, c->in16);
        CUDA_TRY(cudaGetLastError());
        TRY(prof_end(c));
        c->n_launch += 1;
    }
    for (int i = 0; i <= upto; ++i) {
        const b2sr_fused_op& o = c->fops[i];
        while (next_seg < c->fsegs.size() && c->fsegs[next_seg].op_begin < i) ++next_seg;
        if (use_segs && next_seg < c->fsegs.size() && c->fsegs[next_seg].op_begin == i && c->fsegs[next_seg].op_end <= upto) {
            const int rc = launch_segment(c, P, c->fsegs[next_seg]);
            if (rc == 0) {
                i = c->fsegs[next_seg].op_end;
                continue;
            }
            if (rc != B2SR_PIPE_REFUSED) return rc;
            c->pipe_unavailable = 1, use_segs = false;  // the driver cannot make the grid co-resident: launch by launch from here on
            c->n_pipe_fallback += 1;
            fprintf(stderr, "b2sr: device %d cannot hold a %zu x %d persistent grid right now; running the convolutions launch by launch\n",
                    c->device, c->fsegs[next_seg].stages.size(), P->nb);
        }
        if (o.type == B2SR_FOP_NEAREST) {
            const int ri = o.res / o.r;
            const size_t work = (size_t)P->max_plane_px * o.res * o.res * (o.cin / 8);
            dim3 grid((unsigned)std::min<size_t>(148 * 16, (work + 255) / 256), (unsigned)P->planes.size());
            TRY(prof_begin(c, 0, 0));
            tcg_nearest_kernel<<<grid, 256, 0, c->stream>>>((const __half*)c->fbuf_ptr[o.in_buf] + o.in_off, c->fbufs[o.in_buf].channels,
                                                            P->d_planes, ri, o.r, o.cin, (__half*)c->fbuf_ptr[o.out16_buf] + o.out16_off,
                                                            c->fbufs[o.out16_buf].channels);
            CUDA_TRY(cudaGetLastError());
            TRY(prof_end(c));
            c->n_launch += 1;
            continue;
        }
        // (B2SR_PAIR2_MASK, measurements: bit 0 = 32-channel convolutions, bit 1 = 64-channel ones, bit 2 = the fused-shortcut one)
        static const int p2mask = getenv("B2SR_PAIR2_MASK") ? atoi(getenv("B2SR_PAIR2_MASK")) : 7;
        const int p2kind = o.sc_cin ? 4 : (o.cout == 64 ? 2 : 1);
        if (c->pair2 && !c->flip_rows && c->fop_pair2[i] >= 0 && (p2mask & p2kind)) {
            // the CTA-pair form: ONE launch of 2-CTA clusters over band pairs for the whole convolution
            const FusedLaunch& L = c->flaunch2[c->fop_pair2[i]];
            ResItems* R = nullptr;
            TRY(fused_items(c, P, o.res, pair_cluster_count(c), &R, 1));
            TcgParams p{};
            fused_fill_params2(c, P, i, L, p);
            p.items = R->d_items, p.item_first = R->d_first;
            const bool dbg2 = c->pipe_debug && i < c->pipe_debug;
            if (dbg2) {
                if (!c->d_dbg) CUDA_TRY(cudaMalloc(&c->d_dbg, B2SR_DBG_WORDS * sizeof(long long)));
                CUDA_TRY(cudaMemsetAsync(c->d_dbg, 0, B2SR_DBG_WORDS * sizeof(long long), c->stream));
                p.dbg = c->d_dbg;
            }
            TRY(prof_begin(c, 1, R->out_px));
            const int outs = (o.out16_buf >= 0 ? 1 : 0) | (o.out32_buf >= 0 ? 2 : 0);
            bool resf32 = true, resf16 = true;
            for (int q = 0; q < o.nres; ++q) {
                resf32 = resf32 && c->fbufs[o.res_buf[q]].dtype == 4;
                resf16 = resf16 && c->fbufs[o.res_buf[q]].dtype == 2;
            }
            const int key = resf32 ? o.nres * 4 + outs : (resf16 ? 100 + o.nres * 4 + outs : -1);
            int rc;
            if (L.NOUT == 64) {
                switch (key) {
                    case 0 * 4 + 1: rc = launch_tcg2<64, 0, 1>(c, L, R, p); break;
                    case 0 * 4 + 3: rc = launch_tcg2<64, 0, 3>(c, L, R, p); break;
                    case 1 * 4 + 1: rc = launch_tcg2<64, 1, 1>(c, L, R, p); break;
                    case 1 * 4 + 3: rc = launch_tcg2<64, 1, 3>(c, L, R, p); break;
                    case 100 + 1 * 4 + 1: rc = launch_tcg2<64, 1, 1, true>(c, L, R, p); break;
                    default: rc = launch_tcg2<64, -1, 0>(c, L, R, p);
                }
            } else if (L.sc_ks) {
                rc = launch_tcg2<32, 0, 1, false, true>(c, L, R, p);
            } else {
                switch (key) {
                    case 0 * 4 + 1: rc = launch_tcg2<32, 0, 1>(c, L, R, p); break;
                    case 100 + 1 * 4 + 1: rc = launch_tcg2<32, 1, 1, true>(c, L, R, p); break;
                    default: rc = launch_tcg2<32, -1, 0>(c, L, R, p);
                }
            }
            TRY(rc);
            TRY(prof_end(c));
            if (dbg2) {
                std::vector<long long> h(B2SR_DBG_WORDS);
                CUDA_TRY(cudaStreamSynchronize(c->stream));
                CUDA_TRY(cudaMemcpy(h.data(), c->d_dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
                double v[2][16] = {{0}, {0}};
                for (int k = 0; k < 2 * R->n_cta; ++k)
                    for (int j = 0; j < 16; ++j) v[k & 1][j] += (double)h[(size_t)k * 16 + j] / R->n_cta / 1e3;
                fprintf(stderr, "b2sr pair launch op %3d (%3d->%2d, G %d, slots %d, %d clusters), kcycles: leader issuer %6.0f (waits: own row %5.0f, peer row %5.0f, own block %5.0f, peer block %5.0f) | "
                        "producer r0 %6.0f (slot wait %5.0f) r1 %6.0f (%5.0f) | epilogue w2 r0 %6.0f (tfull wait %5.0f, tmem %4.0f) r1 %6.0f (%5.0f, %4.0f)\n",
                        i, o.cin, o.cout, L.G, L.slots, R->n_cta, v[0][0], v[0][1], v[0][8], v[0][2], v[0][3], v[0][7], v[0][4], v[1][7], v[1][4], v[0][6], v[0][5], v[0][10],
                        v[1][6], v[1][5], v[1][10]);
            }
            continue;
        }
        for (int li = c->fop_first[i]; li < c->fop_first[i + 1]; ++li) {
            const FusedLaunch& L = c->flaunch[li];
            if (L.pair && !c->pair_clusters) {
                // how many 2-CTA clusters of this footprint fit the device at once (a GPC with an odd number of SMs loses one)
                cudaLaunchConfig_t q{};
                q.gridDim = dim3((unsigned)c->sms / 2 * 2), q.blockDim = dim3(TC_THREADS);
                q.dynamicSmemBytes = (size_t)TcgCfg<32, 0>::smem_bytes(L.G, L.slots);
                cudaLaunchAttribute qa[1];
                qa[0].id = cudaLaunchAttributeClusterDimension;
                qa[0].val.clusterDim.x = 2, qa[0].val.clusterDim.y = 1, qa[0].val.clusterDim.z = 1;
                q.attrs = qa, q.numAttrs = 1;
                auto kq = tcg_conv_kernel<32, 0, false, 1, 3>;
                TRY(raise_dyn_smem((const void*)kq, (int)q.dynamicSmemBytes));
                int ncl = 0;
                if (cudaOccupancyMaxActiveClusters(&ncl, kq, &q) != cudaSuccess || ncl < 1) {
                    cudaGetLastError();
                    ncl = c->sms / 2 - 4;
                }
                c->pair_clusters = std::min(ncl, c->sms / 2);
            }
            ResItems* R = nullptr;
            TRY(fused_items(c, P, o.res, L.pair ? c->pair_clusters : c->sms, &R));
            TcgParams p{};
            fused_fill_params(c, P, li, d_out, p);
            p.items = R->d_items, p.item_first = R->d_first;
            // Experiment, off by default (B2SR_FLIP=1): alternate launches walk their rows in opposite directions, so that what
            // the previous launch touched LAST is what this one touches FIRST and may still be in L2 (a 540p dense-block
            // buffer is 199 MB).  Measured on B200: no gain (540p 31.0 vs 30.8 ms, batches of 4: 30.7 vs 29.6 ms) -- all 148
            // CTAs sweep their own ranges at once, so no part of the buffer is markedly "more recent" than the rest.
            p.flip = c->flip_rows && (li & 1) && L.wimg_flip;
            if (p.flip) p.wimg = L.wimg_flip;
            if (c->l2_persist && o.in_buf >= 0) {
                const b2sr_fused_buf& B = c->fbufs[o.in_buf];
                const size_t bytes = (size_t)P->total_px * B.res * B.res * B.channels * B.dtype;
                cudaStreamAttrValue av{};
                av.accessPolicyWindow.base_ptr = c->fbuf_ptr[o.in_buf];
                av.accessPolicyWindow.num_bytes = std::min(bytes, c->l2_window_max);
                av.accessPolicyWindow.hitRatio = (float)std::min(1.0, 0.9 * (double)c->l2_persist / (double)av.accessPolicyWindow.num_bytes);
                av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                if (cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
            }
            const bool dbg = c->pipe_debug && (li < c->pipe_debug || li >= (int)c->flaunch.size() - 5);
            if (dbg) {
                if (!c->d_dbg) CUDA_TRY(cudaMalloc(&c->d_dbg, B2SR_DBG_WORDS * sizeof(long long)));
                CUDA_TRY(cudaMemsetAsync(c->d_dbg, 0, B2SR_DBG_WORDS * sizeof(long long), c->stream));
                p.dbg = c->d_dbg;
            }
            TRY(prof_begin(c, 1, R->out_px));
            int rc;
            // epilogue variant: specialised for the combinations RRDB graphs use (fp32 residuals), generic otherwise
            const int outs = (o.out16_buf >= 0 ? 1 : 0) | (o.out32_buf >= 0 ? 2 : 0);
            bool resf32 = true, resf16 = true;
            for (int q = 0; q < o.nres; ++q) {
                resf32 = resf32 && c->fbufs[o.res_buf[q]].dtype == 4;
                resf16 = resf16 && c->fbufs[o.res_buf[q]].dtype == 2;
            }
            const int key = resf32 ? o.nres * 4 + outs : (resf16 ? 100 + o.nres * 4 + outs : -1);
            if (o.final) {
                rc = f32out ? launch_tcg<16, 1, true, -1, 0>(c, L, R, p) : launch_tcg<16, 1, false, -1, 0>(c, L, R, p);
            } else if (L.NOUT == 64) {
                switch (key) {
                    case 0 * 4 + 1: rc = launch_tcg<64, 0, false, 0, 1>(c, L, R, p); break;  // tail convolutions
                    case 0 * 4 + 3: rc = launch_tcg<64, 0, false, 0, 3>(c, L, R, p); break;  // head convolution
                    case 1 * 4 + 1: rc = launch_tcg<64, 0, false, 1, 1>(c, L, R, p); break;  // trunk convolution + skip
                    case 100 + 1 * 4 + 1: rc = launch_tcg<64, 0, false, 1, 1, true>(c, L, R, p); break;
                    default: rc = launch_tcg<64, 0, false, -1, 0>(c, L, R, p);
                }
            } else if (L.sc_ks) {
                rc = launch_tcg<32, 0, false, 0, 1, false, true>(c, L, R, p);  // x2 = lrelu(conv3x3([x, x1])) + conv1x1(x)
            } else {
                switch (key) {
                    case 0 * 4 + 1: rc = launch_tcg<32, 0, false, 0, 1>(c, L, R, p); break;  // x1, x3 (and the 1x1 shortcut kept in fp16)
                    case 0 * 4 + 2: rc = launch_tcg<32, 0, false, 0, 2>(c, L, R, p); break;  // 1x1 shortcut kept in fp32
                    case 1 * 4 + 1: rc = launch_tcg<32, 0, false, 1, 1>(c, L, R, p); break;  // x4 = lrelu(conv) + x2 (fp32 x2)
                    case 1 * 4 + 3: rc = launch_tcg<32, 0, false, 1, 3>(c, L, R, p); break;  // dense-block output: 0.2 v + x
                    case 2 * 4 + 3: rc = launch_tcg<32, 0, false, 2, 3>(c, L, R, p); break;  // RRDB output
                    case 2 * 4 + 1: rc = launch_tcg<32, 0, false, 2, 1>(c, L, R, p); break;  // last RRDB output (no later residual use)
                    case 100 + 1 * 4 + 1: rc = launch_tcg<32, 0, false, 1, 1, true>(c, L, R, p); break;  // x2, x4 with fp16 residuals
                    case 100 + 2 * 4 + 1: rc = launch_tcg<32, 0, false, 2, 1, true>(c, L, R, p); break;
                    default: rc = launch_tcg<32, 0, false, -1, 0>(c, L, R, p);
                }
            }
            TRY(rc);
            TRY(prof_end(c));
            if (dbg) {
                std::vector<long long> h(B2SR_DBG_WORDS);
                CUDA_TRY(cudaStreamSynchronize(c->stream));
                CUDA_TRY(cudaMemcpy(h.data(), c->d_dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
                double v[16] = {0};
                for (int k = 0; k < R->n_cta; ++k)
                    for (int j = 0; j < 16; ++j) v[j] += (double)h[k * 16 + j] / R->n_cta;
                fprintf(stderr, "b2sr fused launch %3d (op %3d k%d %3d->%2d nres %d res %d, G %d slots %d): issuer %7.0f cyc (to first MMA %6.0f; wait full %6.0f, tempty %6.0f) | "
                        "producer %7.0f (wait empty %6.0f) | epilogue warp 2 %7.0f (wait tfull %6.0f, tmem ld/zero %6.0f) | issuer: mma issue %6.0f commits %6.0f\n", li, L.op, o.k, o.cin, L.nco, o.nres, o.res, L.G,
                        L.slots, v[0], v[3], v[1], v[2], v[7], v[4], v[6], v[5], v[10], v[8], v[9]);
            }
        }
    }
    return 0;
}

extern "C" int b2sr_debug_fused(b2sr_ctx* c, const uint8_t* in, int h, int w, int upto, int buf, float* out) {
    if (!c || !in || !out) return fail(B2SR_E_INVALID, "b2sr_debug_fused: null argument");
    if (c->family != B2SR_FAMILY_FUSED) return fail(B2SR_E_UNSUPPORTED, "b2sr_debug_fused: fused family only");
    if (h < 1 || w < 1 || h > 16384 || w > 16384) return fail(B2SR_E_INVALID, "image %dx%d", h, w);
    if (buf < 0 || buf >= (int)c->fbufs.size() || upto < 0 || upto >= (int)c->fops.size() - 1)
        return fail(B2SR_E_INVALID, "b2sr_debug_fused: buffer %d / op %d out of range", buf, upto);
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t in_bytes = (size_t)h * w * 3;
    if (in_bytes > c->cap_in) {
        if (c->d_in) cudaFree(c->d_in);
        c->d_in = nullptr, c->cap_in = 0;
        CUDA_TRY(cudaMalloc(&c->d_in, in_bytes));
        c->cap_in = in_bytes;
    }
    CUDA_TRY(cudaMemcpyAsync(c->d_in, in, in_bytes, cudaMemcpyHostToDevice, c->stream));
    Plan* P = nullptr;
    TRY(build_plan(c, 1, h, w, 0, 0, &P));
    TRY(run_fused(c, P, c->d_in, nullptr, false, upto));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const b2sr_fused_buf& B = c->fbufs[buf];
    const size_t n = (size_t)h * B.res * w * B.res * B.channels;
    if (B.dtype == 4) {
        CUDA_TRY(cudaMemcpy(out, c->fbuf_ptr[buf], n * 4, cudaMemcpyDeviceToHost));
    } else {
        std::vector<__half> tmp(n);
        CUDA_TRY(cudaMemcpy(tmp.data(), c->fbuf_ptr[buf], n * 2, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n; ++i) out[i] = __half2float(tmp[i]);
    }
    return 0;
}

static bool use_tc(const b2sr_ctx* c) { return c->impl != 1; }

// Runs layers [0, upto] for the planes of P; the final layer writes `out` (u8 or f32 frames).  Returns in *act the
// buffer that holds the activations of layer `upto` when upto is not the last layer.
static int run_plan(b2sr_ctx* c, Plan* P, const uint8_t* d_frames, void* d_out, bool f32out, int upto, const __half** act) {
    if (c->family == B2SR_FAMILY_GRAPH) return run_graph(c, P, d_frames, d_out, f32out);
    if (c->family == B2SR_FAMILY_FUSED) return run_fused(c, P, d_frames, d_out, f32out, -1);
    const int nl = (int)c->layers.size();
    if (upto < 0 || upto >= nl) upto = nl - 1;
    if (upto == nl - 1 && auto_pipe(c)) {  // whole network: persistent schedule when it fits the GPU (and pays: auto_pipe)
        if (pipe_fits(c, P)) {
            const int rc = run_pipe(c, P, d_frames, d_out, f32out);
            if (rc != B2SR_PIPE_REFUSED) return rc;
            // the driver could not make the grid co-resident (SMs held by MPS clients, another context's persistent kernel,
            // a MIG / green-context partition smaller than reported): same network layer by layer from here on
            c->pipe_unavailable = 1;
            fprintf(stderr, "b2sr: device %d cannot hold the %d x %d persistent grid right now; using the layer-by-layer schedule\n", c->device, nl, P->nb);
        }
        if (c->impl == 3)
            return fail(B2SR_E_UNSUPPORTED, "pipelined schedule needs layers x bands = %d x %d co-resident CTAs, device offers %d SMs%s", nl, P->nb,
                        usable_sms(c), c->pipe_unavailable ? " (a cooperative launch was refused)" : "");
        if ((int)c->layers.size() <= B2SR_PIPE_MAX_LAYERS) c->n_pipe_fallback += 1;
    }
    TRY(ensure_scratch(c, P->total_px, true));
    TRY(encode_maps(c, P));
    {
        dim3 grid((unsigned)std::min(2048, (P->max_plane_px + 255) / 256), (unsigned)P->planes.size());
        TRY(prof_begin(c, 0, 0));
        prep_kernel<<<grid, 256, 0, c->stream>>>(d_frames, P->h, P->w, P->d_planes, c->in16);
        CUDA_TRY(cudaGetLastError());
        TRY(prof_end(c));
        c->n_launch += 1;
    }
    int in_buf = 0;  // 0 in16, 1 ping, 2 pong
    for (int li = 0; li <= upto; ++li) {
        const bool last = li == nl - 1;
        const int out_buf = in_buf == 1 ? 2 : 1;
        const __half* inp = in_buf == 0 ? c->in16 : (in_buf == 1 ? c->ping : c->pong);
        __half* outp = out_buf == 1 ? c->ping : c->pong;
        if (use_tc(c)) {
            TRY(tc_layer(c, P, li, in_buf, last ? d_out : (void*)outp, d_frames, P->h, P->w, f32out));
        } else if (!last) {
            TRY(simple_layer(c, P, li, inp, outp));
        } else {
            const LayerDev& L = c->layers[li];
            if (c->cap_lastf < P->total_px) {
                cudaStreamSynchronize(c->stream);
                if (c->lastf) cudaFree(c->lastf);
                c->lastf = nullptr, c->cap_lastf = 0;
                CUDA_TRY(cudaMalloc(&c->lastf, (size_t)P->total_px * L.noutp * 4));
                c->cap_lastf = P->total_px;
            }
            TRY(simple_layer(c, P, li, inp, c->lastf));
            dim3 grid((unsigned)std::min(2048, (P->max_plane_px + 255) / 256), (unsigned)P->planes.size());
            TRY(prof_begin(c, 0, 0));
            if (f32out)
                simple_shuffle_kernel<true><<<grid, 256, 0, c->stream>>>(c->lastf, L.noutp, P->d_planes, d_frames, P->h, P->w,
                                                                        c->desc.scale, d_out);
            else
                simple_shuffle_kernel<false><<<grid, 256, 0, c->stream>>>(c->lastf, L.noutp, P->d_planes, d_frames, P->h, P->w,
                                                                         c->desc.scale, d_out);
            CUDA_TRY(cudaGetLastError());
            TRY(prof_end(c));
            c->n_launch += 1;
        }
        if (!last) {
            if (act) *act = outp;
            in_buf = out_buf;
        }
    }
    return 0;
}

static int check_geom(int n, int h, int w, int tile, int halo) {
    if (n < 1 || h < 1 || w < 1) return fail(B2SR_E_INVALID, "empty image (%d frames of %dx%d)", n, h, w);
    if (h > 16384 || w > 16384) return fail(B2SR_E_INVALID, "image %dx%d too large", h, w);
    if (tile < 0 || halo < 0 || (tile > 0 && halo >= tile)) return fail(B2SR_E_INVALID, "bad tiling: tile %d halo %d", tile, halo);
    return 0;
}

static int frames_per_pass(const b2sr_ctx* c, int h, int w, int tile) {
    if (c->max_batch > 0) return c->max_batch;
    if (c->family == B2SR_FAMILY_GRAPH) return 1;
    if (c->family == B2SR_FAMILY_FUSED)  // ~8 KB of activation buffers per input pixel; longer passes amortise the per-launch fill / drain
        return (int)std::max(1.0, std::min(8.0, floor(2.2e6 / ((double)h * w * 1.06))));
    const double px = (double)h * w * 1.06;
    // the pipelined schedule keeps only the 16-channel input planes per frame: long passes amortise its fill/drain
    const int tw = tile > 0 ? std::min(w, tile + 20) : w;
    const bool pipe = auto_pipe(c) && !c->pipe_unavailable && (int64_t)c->layers.size() * ((tw + TC_BW - 1) / TC_BW) <= usable_sms(c);
    if (pipe) return (int)std::max(1.0, std::min(32.0, floor(7.0e7 / px)));
    return (int)std::max(1.0, std::min(16.0, floor(9.0e6 / px)));
}

static int run_batch_dev(b2sr_ctx* c, const uint8_t* d_in, void* d_out, bool f32out, int n, int h, int w, int tile, int halo) {
    const int S = c->desc.scale;
    const int B = frames_per_pass(c, h, w, tile);
    const size_t in_frame = (size_t)h * w * 3, out_frame = (size_t)h * S * w * S * 3 * (f32out ? 4 : 1);
    for (int f = 0; f < n; f += B) {
        const int nb = std::min(B, n - f);
        Plan* P = nullptr;
        TRY(build_plan(c, nb, h, w, tile, halo, &P));
        TRY(run_plan(c, P, d_in + (size_t)f * in_frame, (uint8_t*)d_out + (size_t)f * out_frame, f32out, -1, nullptr));
    }
    return 0;
}

extern "C" int b2sr_run_batch_device(b2sr_ctx* c, const uint8_t* d_in, uint8_t* d_out, int n, int h, int w, int tile,
                                     int halo, int sync) {
    if (!c || !d_in || !d_out) return fail(B2SR_E_INVALID, "b2sr_run_batch_device: null argument");
    TRY(check_geom(n, h, w, tile, halo));
    CUDA_TRY(cudaSetDevice(c->device));
    TRY(run_batch_dev(c, d_in, d_out, false, n, h, w, tile, halo));
    if (sync) CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

static int grow(uint8_t** p, size_t* cap, size_t need) {
    if (need <= *cap) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr, *cap = 0;
    CUDA_TRY(cudaMalloc(p, need));
    *cap = need;
    return 0;
}

static int drain_host_pipeline(b2sr_ctx* c);
static int run_one(b2sr_ctx* c, const uint8_t* in, int h, int w, int in_stride, void* out, int out_stride, bool f32out,
                   int tile, int halo, int memspace) {
    if (!c || !in || !out) return fail(B2SR_E_INVALID, "b2sr_run: null argument");
    TRY(check_geom(1, h, w, tile, halo));
    TRY(drain_host_pipeline(c));  // (pending b2sr_submit_batch_host work shares this call's staging buffers)
    const int S = c->desc.scale;
    const size_t in_row = (size_t)w * 3, out_row = (size_t)w * S * 3 * (f32out ? 4 : 1);
    if (in_stride == 0) in_stride = (int)in_row;
    if (out_stride == 0) out_stride = (int)out_row;
    if ((size_t)in_stride < in_row || (size_t)out_stride < out_row) return fail(B2SR_E_INVALID, "stride smaller than a row");
    if (memspace != B2SR_MEM_HOST && memspace != B2SR_MEM_DEVICE) return fail(B2SR_E_INVALID, "memspace %d", memspace);
    CUDA_TRY(cudaSetDevice(c->device));
    const bool packed = (size_t)in_stride == in_row && (size_t)out_stride == out_row;
    const uint8_t* din = in;
    void* dout = out;
    if (memspace == B2SR_MEM_HOST || !packed) {
        TRY(grow(&c->d_in, &c->cap_in, in_row * h));
        TRY(grow(&c->d_out, &c->cap_out, out_row * h * S));
        CUDA_TRY(cudaMemcpy2DAsync(c->d_in, in_row, in, in_stride, in_row, h,
                                   memspace == B2SR_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, c->stream));
        din = c->d_in, dout = c->d_out;
    }
    TRY(run_batch_dev(c, din, dout, f32out, 1, h, w, tile, halo));
    if (dout != out)
        CUDA_TRY(cudaMemcpy2DAsync(out, out_stride, dout, out_row, out_row, (size_t)h * S,
                                   memspace == B2SR_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int b2sr_run_u8(b2sr_ctx* c, const uint8_t* in, int h, int w, int in_stride, uint8_t* out, int out_stride, int tile,
                           int halo, int memspace) {
    return run_one(c, in, h, w, in_stride, out, out_stride, false, tile, halo, memspace);
}
extern "C" int b2sr_run_f32(b2sr_ctx* c, const uint8_t* in, int h, int w, int in_stride, float* out, int out_stride, int tile,
                            int halo, int memspace) {
    return run_one(c, in, h, w, in_stride, out, out_stride, true, tile, halo, memspace);
}

// Host frames -> H2D -> network -> D2H, chunk by chunk, copies overlapped with compute on separate streams
// (double-buffered staging).  The reference moves every frame through PNG files instead (upscale_processing.py:487,:519).
// Waits for every submission of the host-batch pipeline (they use d_in / d_out and the copy streams).
static int drain_host_pipeline(b2sr_ctx* c) {
    if (!c->hb_pending) return 0;
    c->hb_pending = false;
    cudaError_t e1 = cudaStreamSynchronize(c->copy_in), e2 = cudaStreamSynchronize(c->stream), e3 = cudaStreamSynchronize(c->copy_out);
    for (cudaError_t e : {e1, e2, e3})
        if (e != cudaSuccess) return fail(B2SR_E_CUDA, "host-batch pipeline: %s", cudaGetErrorString(e));
    return 0;
}

// Enqueues n host frames as chunks on the three streams: H2D of chunk k+1 (copy_in), the network on chunk k (stream) and D2H of
// chunk k-1 (copy_out) run concurrently; two staging slots, ordered by events that persist across calls, so that the first copy
// of one call may run under the last compute of the previous one.  Nothing here waits on the host.
static int enqueue_batch_host(b2sr_ctx* c, const uint8_t* h_in, uint8_t* h_out, int n, int h, int w, int tile, int halo, bool taper) {
    const int S = c->desc.scale;
    // frames per chunk (B2SR_OPT_MAX_BATCH overrides): 4 in a synchronous call, 8 in a submission whose ends overlap with its
    // neighbours (measured at 1080p, tools/e2e_stream_sweep.py: 393 / 396 / 394 fps end to end with chunks of 4 / 8 / 16)
    const int B = std::min(frames_per_pass(c, h, w, tile), c->max_batch > 0 ? c->max_batch : (taper ? 4 : 8));
    const size_t in_frame = (size_t)h * w * 3, out_frame = in_frame * S * S;
    if (in_frame * B > c->cap_in || out_frame * B > c->cap_out || in_frame * B > c->cap_in2 || out_frame * B > c->cap_out2)
        TRY(drain_host_pipeline(c));  // the staging buffers are about to be replaced
    TRY(grow(&c->d_in, &c->cap_in, in_frame * B));
    TRY(grow(&c->d_out, &c->cap_out, out_frame * B));
    TRY(grow(&c->d_in2, &c->cap_in2, in_frame * B));
    TRY(grow(&c->d_out2, &c->cap_out2, out_frame * B));
    uint8_t* din[2] = {c->d_in, c->d_in2};
    uint8_t* dout[2] = {c->d_out, c->d_out2};
    for (int i = 0; i < 2; ++i) {
        if (!c->hb_in_ready[i]) CUDA_TRY(cudaEventCreateWithFlags(&c->hb_in_ready[i], cudaEventDisableTiming));
        if (!c->hb_compute_done[i]) CUDA_TRY(cudaEventCreateWithFlags(&c->hb_compute_done[i], cudaEventDisableTiming));
        if (!c->hb_out_done[i]) CUDA_TRY(cudaEventCreateWithFlags(&c->hb_out_done[i], cudaEventDisableTiming));
    }
    // Chunk schedule.  In a synchronous call the H2D of the first chunk and the D2H of the last one are not hidden behind
    // compute: both are one frame long there (tapered ends), the chunks in between hold B frames.  Submissions that overlap
    // with their neighbours use equal chunks.
    std::vector<int> chunks;
    {
        int left = n;
        const int tail = (taper && n >= 3) ? 1 : 0;
        if (tail) chunks.push_back(1), left -= 2;
        for (; left > 0; left -= B) chunks.push_back(std::min(B, left));
        if (tail) chunks.push_back(1);
    }
    c->hb_pending = true;
    int f = 0;
    for (size_t k = 0; k < chunks.size(); f += chunks[k], ++k) {
        const int nb = chunks[k], s = (int)(c->hb_seq++ & 1u);
        // staging slot s is free once the compute that read din[s] and the D2H that read dout[s] (two chunks ago, maybe in the
        // previous call) are done
        if (c->hb_used[s]) {
            CUDA_TRY(cudaStreamWaitEvent(c->copy_in, c->hb_compute_done[s], 0));
            CUDA_TRY(cudaStreamWaitEvent(c->stream, c->hb_out_done[s], 0));
        }
        c->hb_used[s] = true;
        CUDA_TRY(cudaMemcpyAsync(din[s], h_in + (size_t)f * in_frame, in_frame * nb, cudaMemcpyHostToDevice, c->copy_in));
        CUDA_TRY(cudaEventRecord(c->hb_in_ready[s], c->copy_in));
        CUDA_TRY(cudaStreamWaitEvent(c->stream, c->hb_in_ready[s], 0));
        TRY(run_batch_dev(c, din[s], dout[s], false, nb, h, w, tile, halo));
        CUDA_TRY(cudaEventRecord(c->hb_compute_done[s], c->stream));
        CUDA_TRY(cudaStreamWaitEvent(c->copy_out, c->hb_compute_done[s], 0));
        CUDA_TRY(cudaMemcpyAsync(h_out + (size_t)f * out_frame, dout[s], out_frame * nb, cudaMemcpyDeviceToHost, c->copy_out));
        CUDA_TRY(cudaEventRecord(c->hb_out_done[s], c->copy_out));
    }
    return 0;
}

extern "C" int b2sr_run_batch_host(b2sr_ctx* c, const uint8_t* h_in, uint8_t* h_out, int n, int h, int w, int tile, int halo) {
    if (!c || !h_in || !h_out) return fail(B2SR_E_INVALID, "b2sr_run_batch_host: null argument");
    TRY(check_geom(n, h, w, tile, halo));
    CUDA_TRY(cudaSetDevice(c->device));
    static const bool taper = !(getenv("B2SR_E2E_TAPER") && atoi(getenv("B2SR_E2E_TAPER")) == 0);  // B2SR_E2E_TAPER=0: equal chunks
    const int rc = enqueue_batch_host(c, h_in, h_out, n, h, w, tile, halo, taper);
    const int rd = drain_host_pipeline(c);  // (also after a failed enqueue: nothing of this call may still be running when it returns)
    return rc ? rc : rd;
}

extern "C" int b2sr_submit_batch_host(b2sr_ctx* c, const uint8_t* h_in, uint8_t* h_out, int n, int h, int w, int tile, int halo,
                                      uint64_t* ticket) {
    if (!c || !h_in || !h_out || !ticket) return fail(B2SR_E_INVALID, "b2sr_submit_batch_host: null argument");
    TRY(check_geom(n, h, w, tile, halo));
    CUDA_TRY(cudaSetDevice(c->device));
    const int rc = enqueue_batch_host(c, h_in, h_out, n, h, w, tile, halo, false);
    if (rc) {
        const std::string keep = b2sr_last_error();
        drain_host_pipeline(c);
        fail(rc, "%s", keep.c_str());
        return rc;
    }
    const uint64_t t = c->hb_next_ticket++;
    cudaEvent_t& ev = c->hb_ticket[t % 8];
    if (!ev) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(ev, c->copy_out));  // the D2H copies of a context complete in submission order
    *ticket = t;
    return 0;
}

extern "C" int b2sr_wait_batch(b2sr_ctx* c, uint64_t ticket) {
    if (!c) return fail(B2SR_E_INVALID, "b2sr_wait_batch: null context");
    if (ticket == 0 || ticket >= c->hb_next_ticket) return fail(B2SR_E_INVALID, "b2sr_wait_batch: unknown ticket %llu", (unsigned long long)ticket);
    CUDA_TRY(cudaSetDevice(c->device));
    // a ticket whose event slot has been reused by a later submission: that one's copies finish after this one's
    uint64_t t = ticket;
    while (t + 8 < c->hb_next_ticket) t += 8;
    CUDA_TRY(cudaEventSynchronize(c->hb_ticket[t % 8]));
    if (ticket + 1 == c->hb_next_ticket) c->hb_pending = false;  // the newest submission is complete: so is every earlier one
    return 0;
}

extern "C" int b2sr_debug_layer(b2sr_ctx* c, const uint8_t* in, int h, int w, int layer, float* out) {
    if (!c || !in || !out) return fail(B2SR_E_INVALID, "b2sr_debug_layer: null argument");
    TRY(check_geom(1, h, w, 0, 0));
    if (c->family != B2SR_FAMILY_COMPACT) return fail(B2SR_E_UNSUPPORTED, "b2sr_debug_layer: Compact family only");
    TRY(drain_host_pipeline(c));
    if (layer < 0 || layer >= (int)c->layers.size() - 1) return fail(B2SR_E_INVALID, "layer %d has no activation output", layer);
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t in_bytes = (size_t)h * w * 3;
    TRY(grow(&c->d_in, &c->cap_in, in_bytes));
    CUDA_TRY(cudaMemcpyAsync(c->d_in, in, in_bytes, cudaMemcpyHostToDevice, c->stream));
    Plan* P = nullptr;
    TRY(build_plan(c, 1, h, w, 0, 0, &P));
    const __half* act = nullptr;
    TRY(run_plan(c, P, c->d_in, nullptr, false, layer, &act));
    const int nf = c->desc.nf;
    const size_t npx = (size_t)h * w;
    TRY(grow(&c->d_out, &c->cap_out, npx * nf * 4));
    unpack_act_kernel<<<(unsigned)std::min<size_t>(4096, (npx * nf + 255) / 256), 256, 0, c->stream>>>(act, c->CF, nf, npx, (float*)c->d_out);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, c->d_out, npx * nf * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// options, stats
// ------------------------------------------------------------------------------------------------
extern "C" int b2sr_set_option(b2sr_ctx* c, int key, int64_t value) {
    if (!c) return fail(B2SR_E_INVALID, "null context");
    switch (key) {
        case B2SR_OPT_IMPL:
            if (value < 0 || value > 3) return fail(B2SR_E_INVALID, "impl %lld", (long long)value);
            c->impl = (int)value;
            return 0;
        case B2SR_OPT_PROFILE:
            c->profile = value != 0;
            return 0;
        case B2SR_OPT_MAX_BATCH:
            if (value < 0 || value > 4096) return fail(B2SR_E_INVALID, "max_batch %lld", (long long)value);
            c->max_batch = (int)value;
            return 0;
        case B2SR_OPT_PIPE_DEBUG:
            c->pipe_debug = (int)value;
            return 0;
        case B2SR_OPT_SM_LIMIT:
            if (value < 0 || value > 100000) return fail(B2SR_E_INVALID, "sm limit %lld", (long long)value);
            c->sm_limit = (int)value;
            c->pipe_unavailable = 0;  // a new limit is a new chance for the persistent schedule
            return 0;
        case B2SR_OPT_SEG_PIPE:
            c->seg_pipe = value != 0;
            return 0;
        case B2SR_OPT_PAIR2:
            c->pair2 = value != 0;
            return 0;
        case B2SR_OPT_ABLATE:
            c->ablate = (int)value;
            return 0;
        case B2SR_OPT_RING_ROWS:
            if (value != 0 && (value < 4 || value > 4096)) return fail(B2SR_E_INVALID, "ring rows %lld (need 0 = auto, or 4..4096)", (long long)value);
            c->ring_rows = (int)value;
            return 0;
    }
    return fail(B2SR_E_INVALID, "unknown option %d", key);
}

extern "C" int b2sr_reset_stats(b2sr_ctx* c) {
    if (!c) return fail(B2SR_E_INVALID, "null context");
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->n_launch = c->n_tc = c->n_pipe = c->n_hmma = c->n_pipe_fallback = 0;
    for (auto& r : c->prof) {
        c->ev_pool.push_back(r.a);
        c->ev_pool.push_back(r.b);
    }
    c->prof.clear();
    return 0;
}

extern "C" int b2sr_get_stat(b2sr_ctx* c, int key, double* value) {
    if (!c || !value) return fail(B2SR_E_INVALID, "null argument");
    switch (key) {
        case B2SR_STAT_LAUNCHES:
            *value = c->n_launch;
            return 0;
        case B2SR_STAT_TC_LAUNCHES:
            *value = c->n_tc;
            return 0;
        case B2SR_STAT_PIPE_LAUNCHES:
            *value = c->n_pipe;
            return 0;
        case B2SR_STAT_HMMA_LAUNCHES:
            *value = c->n_hmma;
            return 0;
        case B2SR_STAT_PIPE_FALLBACKS:
            *value = c->n_pipe_fallback;
            return 0;
        case B2SR_STAT_TC_MID_MS:
        case B2SR_STAT_TC_MID_COUNT:
        case B2SR_STAT_ALL_MS:
        case B2SR_STAT_PIPE_MS:
        case B2SR_STAT_TC_MID_PIXELS: {
            CUDA_TRY(cudaSetDevice(c->device));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            double mid_ms = 0, all_ms = 0, cnt = 0, px = 0, pipe_ms = 0;
            for (auto& r : c->prof) {
                float ms = 0;
                CUDA_TRY(cudaEventElapsedTime(&ms, r.a, r.b));
                all_ms += ms;
                if (r.kind == 1) mid_ms += ms, cnt += 1, px += r.px;
                if (r.kind == 2 || r.kind == 3) pipe_ms += ms;
            }
            *value = key == B2SR_STAT_TC_MID_MS ? mid_ms : key == B2SR_STAT_TC_MID_COUNT ? cnt : key == B2SR_STAT_ALL_MS ? all_ms
                     : key == B2SR_STAT_PIPE_MS ? pipe_ms : px;
            return 0;
        }
    }
    return fail(B2SR_E_INVALID, "unknown stat %d", key);
}

extern "C" int b2sr_synchronize(b2sr_ctx* c) {
    if (!c) return fail(B2SR_E_INVALID, "null context");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" void* b2sr_stream(b2sr_ctx* c) { return c ? (void*)c->stream : nullptr; }


// ------------------------------------------------------------------------------------------------
// start-up weight broadcast over NCCL: every rank creates its context from the same network
// description -- rank `root` with the real parameters, the others with anything of the right size (zeros) -- and one
// grouped ncclBroadcast per device-side parameter buffer makes them identical.  Nothing is communicated afterwards.
// NCCL is resolved at run time (the copy already loaded into the process -- e.g. torch's -- else libnccl.so.2), so the
// library itself has no link-time dependency on it.
// ------------------------------------------------------------------------------------------------
#include <dlfcn.h>
namespace {
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, const void*, int) = nullptr;  // ncclUniqueId is passed BY VALUE (128 bytes) in the real ABI: see nccl_comm_init below
    int (*CommDestroy)(void*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
struct NcclId {
    char bytes[128];
};
NcclApi g_nccl;
int load_nccl() {
    if (g_nccl.lib) return 0;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(B2SR_E_UNSUPPORTED, "NCCL is not available in this process (%s)", dlerror());
    g_nccl.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, const void*, int))dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
    g_nccl.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
    g_nccl.Broadcast = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclBroadcast");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.GroupStart || !g_nccl.GroupEnd || !g_nccl.Broadcast)
        return fail(B2SR_E_UNSUPPORTED, "libnccl.so.2 lacks an expected symbol");
    g_nccl.lib = h;
    return 0;
}
int nccl_fail(const char* what, int rc) {
    return fail(B2SR_E_CUDA, "%s failed: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "NCCL error");
}
}  // namespace

extern "C" int b2sr_nccl_unique_id(void* id128) {
    if (!id128) return fail(B2SR_E_INVALID, "b2sr_nccl_unique_id: null argument");
    TRY(load_nccl());
    const int rc = g_nccl.GetUniqueId(id128);
    return rc ? nccl_fail("ncclGetUniqueId", rc) : 0;
}

extern "C" int b2sr_nccl_comm_init(void** comm, int n_ranks, int rank, const void* id128, int device) {
    if (!comm || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(B2SR_E_INVALID, "b2sr_nccl_comm_init: bad argument");
    TRY(load_nccl());
    CUDA_TRY(cudaSetDevice(device));
    NcclId id;
    memcpy(id.bytes, id128, sizeof id.bytes);
    // ncclResult_t ncclCommInitRank(ncclComm_t* comm, int nranks, ncclUniqueId commId, int rank): the id travels by value
    auto init = (int (*)(void**, int, NcclId, int))(void*)g_nccl.CommInitRank;
    const int rc = init(comm, n_ranks, id, rank);
    return rc ? nccl_fail("ncclCommInitRank", rc) : 0;
}

extern "C" int b2sr_nccl_comm_destroy(void* comm) {
    if (!comm) return 0;
    TRY(load_nccl());
    const int rc = g_nccl.CommDestroy(comm);
    return rc ? nccl_fail("ncclCommDestroy", rc) : 0;
}

extern "C" int b2sr_bcast_weights(b2sr_ctx* c, void* nccl_comm, int root) {
    if (!c || !nccl_comm || root < 0) return fail(B2SR_E_INVALID, "b2sr_bcast_weights: bad argument");
    TRY(load_nccl());
    CUDA_TRY(cudaSetDevice(c->device));
    std::vector<std::pair<void*, size_t>> bufs;  // every device-side parameter buffer of the context, in creation order
    for (const LayerDev& L : c->layers) {
        const size_t PB = (size_t)L.cinp * 2;
        bufs.push_back({L.wimg, (size_t)9 * L.noutp * PB});
        bufs.push_back({L.wplain, (size_t)9 * L.cinp * L.noutp * 2});
        bufs.push_back({L.bias, (size_t)L.noutp * 4});
        if (L.slope) bufs.push_back({L.slope, (size_t)L.noutp * 4});
    }
    for (const FusedLaunch& L : c->flaunch) {
        const b2sr_fused_op& o = c->fops[L.op];
        const int parts = L.pair ? 2 : 1;
        const size_t img = (size_t)parts * L.G * 9 * L.NOUT * TCG_PB + (L.sc_ks ? (size_t)L.NOUT * TCG_PB : 0);
        (void)o;
        bufs.push_back({L.wimg, img});
        if (L.wimg_flip) bufs.push_back({L.wimg_flip, img});
        bufs.push_back({L.bias, (size_t)parts * L.NOUT * 4});
        bufs.push_back({L.slope, (size_t)parts * L.NOUT * 4});
    }
    for (const GraphOp& g : c->gops) {
        const b2sr_graph_op& o = g.op;
        if (o.type == B2SR_OP_CONV) {
            bufs.push_back({g.w, (size_t)o.k * o.k * o.cin * g.coutp * 4});
            if (g.wh) bufs.push_back({g.wh, (size_t)o.k * o.k * o.cin * o.cout * 2});
            if (g.b) bufs.push_back({g.b, (size_t)o.cout * 4});
        } else if (o.type == B2SR_OP_PRELU) {
            bufs.push_back({g.w, (size_t)o.in_c[0] * 4});
        }
    }
    int rc = g_nccl.GroupStart();
    if (rc) return nccl_fail("ncclGroupStart", rc);
    for (auto& b : bufs)
        if (b.first && b.second && !rc) rc = g_nccl.Broadcast(b.first, b.first, b.second, /*ncclUint8*/ 1, root, nccl_comm, c->stream);
    const int rc2 = g_nccl.GroupEnd();
    if (rc) return nccl_fail("ncclBroadcast", rc);
    if (rc2) return nccl_fail("ncclGroupEnd", rc2);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// denoise pass (reference apply_denoise, upscale/upscale_processing.py:350-362)
// ------------------------------------------------------------------------------------------------
#include "nlm_host.inl"
