// tc_gconv.cuh -- tcgen05 convolution for RRDB-style graphs (sm_100a): 3x3 / 1x1 convolutions whose input is a view
// of up to 192 channels of a wide channel-last fp16 buffer (the dense-block "concat" never materialises), with bias,
// LeakyReLU and up to two residual terms `v = v * cv + r * cr` folded into the epilogue.
//
// Replaces, for models/4x_Valar_v1.param (reference upscale/upscale_processing.py:437-453 `ex.extract` on that graph),
// the layer groups  Convolution [+ BinaryOp / Eltwise ...]  and makes Concat / Split free:
//   Conv_1 .. Conv_16 + Add_7 / Add_14 / Add_19 (:6-22), the RRDB tails Add_57 + Add_60 (:57-58), Conv_1381 + Add_1382,
//   the up-sampling tail Conv_1384 .. Conv_1391 (:1201-1208).
//
// Same contraction as tc_conv.cuh (row-stationary, the three ky taps stacked along N, accumulator blocks in a TMEM
// ring, one M = 128 tile = one band row), generalised along K: the input view is cut into G = ceil(cin / 64) channel
// groups; a shared-memory ring slot holds ONE group of one input row (136 pixels x 64 channels, 128-byte swizzle rows,
// written by one TMA box {64 ch, 136 px} whose channels beyond `cin` are zero-filled by the TMA unit); the issuer walks
// (row, group) pairs and accumulates all groups of a row into the same TMEM blocks, so a 192-channel input costs three
// ring slots per row and no concat copy.  The stacked weights of all groups stay resident in shared memory
// ([g][kx][3 * NOUT rows][64 ch], <= 110.6 KB).  A 192 -> 64 convolution (221 KB of weights) runs as clusters of two CTAs:
// both walk the same rows, rank r keeps the weights of output channels 32r .. 32r+31, and rank 0 feeds both rings with
// one multicast TMA load per (row, group).
//
// Accumulators: output row y lives in TMEM block y mod NB; the three-block window of an input row never wraps (two
// extension blocks stand for homes 0 and 1 and are added by the epilogue), and because homes follow plane rows a frame
// is computed bit-identically whatever the CTA ranges are.  An optional second set of NB blocks accumulates a fused
// 1x1 shortcut convolution over the first channels of the same input rows (x2 = lrelu(conv3x3([x, x1])) + conv1x1(x)).
//
// Launches use programmatic dependent launch: setup (barriers, TMEM, weights) overlaps the previous launch's tail,
// griddepcontrol.wait precedes the first access to activation buffers.
//
// Epilogue (two warp sets on alternate rows): tcgen05.ld -> + bias -> LeakyReLU -> residual terms read from fp32 (or
// fp16) buffers -> fp32 copy for later residual adds (the RRDB trunk stays unrounded) and / or fp16 copy into a
// channel slice of the next convolution's input buffer (swizzled staging, 16-byte coalesced stores); the network's
// last convolution writes `* 255`, cropped to the tile core, rounded half-to-even and saturated, as u8 frames.
#pragma once
#include "tc_conv.cuh"

namespace b2sr {

struct TcgParams {
    const CUtensorMap* maps;  // device array, indexed by map_base + TcItem::map (plane-size group)
    int32_t map_base;
    const TcItem* items;
    const int32_t* item_first;  // CTA k processes items [item_first[k], item_first[k+1])
    const uint8_t* wimg;        // [groups][kx][(2-ky)*NOUT + o][64 halfs], 128-byte swizzled rows
    const float* bias;          // [NOUT]
    const float* slope;         // [NOUT]: LeakyReLU slope, 1 = no activation
    float acc_scale;            // v = acc * acc_scale + bias (1/255 for the convolution fed with raw 0..255 pixels)
    int32_t groups;             // channel groups of 64 in the input view (1..3)
    int32_t cin;                // channels of the input view (multiple of 16)
    int32_t k1;                 // 1x1 convolution: only the centre column tap is issued (the others are zero)
    int32_t sc_ks;              // fused 1x1 shortcut over the first sc_ks * 16 channels of the input view (SC kernels)
    float sc_cv, sc_cr;         // v = act(conv + bias) * sc_cv + shortcut * sc_cr
    int32_t pair;               // launched as clusters of two CTAs that share their input rows (TMA multicast) and compute
                                // NOUT output channels each: CTA rank r uses weights / bias / slope / output / residual slice r
    int32_t pair_wbytes;        // bytes between the two weight images
    int32_t flip;               // rows are walked bottom-up (plane row = Ht - 1 - y; wimg has the ky blocks swapped to match)
    int32_t ring_slots;         // shared-memory ring slots (one (row, group) each)
    int32_t nres;               // residual terms
    const void* res_ptr[2];     // channel 0 of this launch's slice, pixel 0 of the buffer
    int32_t res_ld[2];          // elements per pixel
    int32_t res_f32[2];
    float coef_v[2], coef_r[2];
    __half* out16;              // fp16 output slice or nullptr
    int32_t out16_ld;
    float* out32;               // fp32 output slice or nullptr
    int32_t out32_ld;
    void* frames_out;           // final convolution: packed frames (u8 or float), frame_h x frame_w x 3
    int32_t frame_h, frame_w;
    long long* dbg;             // optional [CTA][8] stall accounting (B2SR_OPT_PIPE_DEBUG)
    // ---- pipelined mode (tcg_pipe_kernel): CTA = (stage, band) of a segment of consecutive convolutions; every buffer
    // slice that is written AND read inside the segment lives in an L2-resident ring of RR rows (pitch Wmax pixels, the
    // buffer's own channel layout) instead of a frame-sized HBM buffer; flow control is per band and row through one
    // monotonic `done` counter per (stage, band), B2SR_FLAG_STRIDE words apart.
    int32_t nb, RR, Wmax;
    int32_t ring_map_base;      // maps[ring_map_base + TcItem::map]: ring instance read by this stage's input view
    int32_t grp_ring[3];        // channel group g of the input view comes from the ring (else from the frame buffer)
    int32_t out16_ring, out32_ring, res_ring[2];  // which outputs / residual sources are rings
    int32_t n_in;               // stages whose `done` gates this stage's ring input rows (the op that wrote last): 0..2
    const uint32_t* done_in[2];
    uint32_t* done_out;         // rows this stage's band CTAs have written and released
    int32_t n_bp;               // stages whose `done` frees this stage's output ring slots (the last op that reads them): 0..2
    const uint32_t* bp[2];
    int32_t variant;            // which template instance runs this stage (tcg_pipe_kernel)
    int32_t nogate;             // measurement only (B2SR_SEG_NOGATE=1): ignore every counter -- stages free-run on stale data, results are garbage
    int32_t half;               // which 32-channel half of a 64-channel convolution this stage computes (weights / bias /
                                // output / residual slices are offset like a cluster rank's in the paired launch)
    int32_t ablate;             // measurement only (B2SR_ABLATE, energy accounting; results are garbage): bit 0 = no MMAs are issued
                                // (commits only), bit 1 = the epilogue drains the accumulators but neither reads residuals nor
                                // computes / stores anything, bit 2 = the producer arrives on the full barriers without loading rows,
                                // bit 3 = the epilogue does everything but its global loads and stores
};

// Programmatic dependent launch: a kernel launched with the programmatic-stream-serialization attribute may become resident
// while its predecessor drains; griddep_wait() blocks until the predecessor grid has completed and its memory is visible.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one TMA load delivered to the same shared-memory offset (and mbarrier) of every CTA in `mask`
__device__ __forceinline__ void tma_load_4d_multicast(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3,
                                                      uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "h"(mask)
        : "memory");
}
// arrive (once all MMAs issued so far have completed) on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_multicast(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
                 : "memory");
}

// ---- CTA-pair (cta_group::2) helpers: one instruction stream drives the tensor cores of two SMs (M = 256) ----
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// arrive (once all MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3)
                 : "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `target` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t target) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar), "r"(target)
        : "memory");
}
// wait on a local barrier that a thread of the PEER CTA arrives on (cluster-scope acquire)
__device__ __forceinline__ uint32_t mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, int who) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (clock64() - t0 > 6000000000LL) mbar_timeout(bar, parity, who);
    }
}

constexpr int TCG_PB = 128;                      // bytes per pixel of one channel group == one SW128 swizzle row
constexpr int TCG_SUBROWB = TC_PITCH * TCG_PB;   // one ring slot
constexpr int TCG_PIPE_WORDS = 8;  // 8-byte slots for the pipelined mode's shared words (progress, polled minima)
constexpr int TCG_PAIR_WORDS = TC_MAX_SLOTS + TC_NBLK + 2;  // CTA-pair mode: the peer's full / tempty / weights barriers in the leader
constexpr int TCG_BAR_WORDS = 2 * TC_MAX_SLOTS + 2 * TC_NBLK + 2 + TCG_PIPE_WORDS + TCG_PAIR_WORDS;

template <int NOUT, int MODE, bool SC = false>
struct TcgCfg {
    static constexpr int OB = NOUT * 2;
    static constexpr int STG = MODE == 0 ? TC_NSETS * 4 * 32 * OB : 0;
    static constexpr int MISC = 2 * NOUT * 4 + TCG_BAR_WORDS * 8 + 64;
    // Accumulator blocks: output row y lives in block y % NB ("home").  The window of an input row (the blocks of output
    // rows r-1, r, r+1) never wraps: it may run into two extension blocks NB, NB + 1 that stand for homes 0 and 1, and
    // the epilogue adds block NB + h to block h for rows with home h < 2.  (With a wrapping ring two of every NB rows
    // needed a second, small-N MMA per tap and slab; small-N MMAs cost as much as N = 96 ones.)
    // SC kernels keep a second set of NB blocks (the 1x1 shortcut's accumulators) behind the extension blocks.
    static constexpr int NB = (NOUT == 64 || SC) ? 6 : 8;
    static constexpr int SC_COL0 = (NB + 2) * NOUT;  // first TMEM column of the shortcut blocks
    static constexpr int USED = (NB + 2) * NOUT + (SC ? NB * NOUT : 0);
    static constexpr int TCOLS = USED <= 128 ? 128 : (USED <= 256 ? 256 : 512);
    static constexpr int SCB = SC ? NOUT * TCG_PB : 0;  // shortcut weight image [NOUT rows][64 ch]
    static constexpr uint32_t IDESC0 = (1u << 4) | ((uint32_t)(TC_TILE_M >> 4) << 24);
    static_assert(NOUT == 16 || NOUT == 32 || NOUT == 64, "NOUT");
    static_assert(USED <= 512 && NB % TC_NSETS == 0, "accumulator blocks exceed TMEM");
    static_assert(!SC || (NOUT == 32 && MODE == 0), "the fused shortcut exists for 32-channel activation launches");
    static constexpr int weight_bytes(int groups) { return groups * 9 * NOUT * TCG_PB + SCB; }
    static constexpr int ring_fit(int groups) { return (B2SR_SMEM_LIMIT - 1024 - weight_bytes(groups) - STG - MISC) / TCG_SUBROWB; }
    static constexpr int smem_bytes(int groups, int slots) { return 1024 + weight_bytes(groups) + slots * TCG_SUBROWB + STG + MISC; }
    // CTA-pair mode: every CTA keeps HALF of the stacked weights (its half of the B rows of every MMA)
    static constexpr int ring_fit2(int groups) { return (B2SR_SMEM_LIMIT - 1024 - weight_bytes(groups) / 2 - STG - MISC) / TCG_SUBROWB; }
    static constexpr int smem_bytes2(int groups, int slots) { return 1024 + weight_bytes(groups) / 2 + slots * TCG_SUBROWB + STG + MISC; }
};

// NRES / OUTS specialise the MODE 0 epilogue (a lone warp per scheduler runs it: its instruction count is its speed):
// NRES = number of residual terms, all read from fp32 buffers (RF16: all from fp16 buffers), or -1 = everything taken
// from the parameters at run time; OUTS = bit 0: fp16 copy, bit 1: fp32 copy, or 0 = decided at run time.  NRES = 0, OUTS = 1 is the
// plain bias + LeakyReLU -> fp16 epilogue.
template <int NOUT, int MODE /*0 = activation buffers, 1 = network output frames*/, bool F32OUT, int NRES, int OUTS, bool RF16,
          bool SC /*fused 1x1 shortcut*/, bool PIPE /*CTA = (stage, band) of a persistent segment launch*/,
          bool P2 = false /*CTA pair: the two CTAs of a cluster take neighbouring bands of the same rows, cta_group::2 MMAs*/>
__device__ __forceinline__ void tcg_body(const TcgParams& P, const int it_begin, const int it_end, const int band, const uint32_t rank,
                                         uint8_t* smem_raw) {
    using C = TcgCfg<NOUT, MODE, SC>;
    static_assert(!PIPE || MODE == 0, "pipelined stages write activation buffers");
    static_assert(!P2 || (!PIPE && MODE == 0), "the CTA-pair form exists for per-launch activation convolutions");
    // CTA-pair mode (P2).  Both CTAs walk the same rows; CTA r owns band 2 * pair + r: its own input rows (own TMA ring), its
    // own accumulators and epilogue warps, and HALF of the stacked weights (rows [r * N/2, (r+1) * N/2) of every tile).  The
    // leader (r = 0) issues tcgen05.mma.cta_group::2 (M = 256: both bands, N = 3 * NOUT) once both rings hold the row; its
    // commits arrive on the barriers of both CTAs.  One instruction stream per two bands halves the per-band issue cost of
    // these issue-bound layers, and half the weights per CTA leaves room for a deeper input ring.  Windows are ALWAYS full
    // (N = 3 * NOUT, the B split must not move): the two output rows above and below a CTA range are accumulated as
    // phantom rows and drained without being stored.
    const uint32_t prank = P2 ? cluster_ctarank() : 0u;
    [[maybe_unused]] const bool leader = prank == 0u;
    const uint8_t* wimg = P.wimg + (size_t)rank * P.pair_wbytes;
    const float* bias_g = P.bias + rank * NOUT;
    const float* slope_g = P.slope + rank * NOUT;
    __half* const out16 = P.out16 ? P.out16 + rank * NOUT : nullptr;
    float* const out32 = P.out32 ? P.out32 + rank * NOUT : nullptr;
    const void* res_ptr[2];
#pragma unroll
    for (int r = 0; r < 2; ++r)
        res_ptr[r] = reinterpret_cast<const uint8_t*>(P.res_ptr[r]) + (size_t)rank * NOUT * (P.res_f32[r] ? 4 : 2);
    const int G = P.groups, R = P.ring_slots;
    constexpr int NB = C::NB;
    const uint32_t WB = ((uint32_t)(G * 9 * NOUT * TCG_PB) + C::SCB) / (P2 ? 2u : 1u);  // stacked 3x3 weights of all groups [+ the shortcut image]
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (sbase - raw);

    const uint32_t w_s = sbase;
    const uint32_t ring_s = sbase + WB;
    const uint32_t stg_off = WB + (uint32_t)R * TCG_SUBROWB;
    const uint32_t fl_off = stg_off + C::STG;
    float* s_bias = reinterpret_cast<float*>(gbase + fl_off);
    float* s_slope = s_bias + NOUT;
    const uint32_t bar_off = fl_off + 2 * NOUT * 4;
    const uint32_t bar_s = sbase + bar_off;
    auto full_bar = [&](int s) { return bar_s + 8u * s; };
    auto empty_bar = [&](int s) { return bar_s + 8u * (TC_MAX_SLOTS + s); };
    auto tfull_bar = [&](int b) { return bar_s + 8u * (2 * TC_MAX_SLOTS + b); };
    auto tempty_bar = [&](int b) { return bar_s + 8u * (2 * TC_MAX_SLOTS + TC_NBLK + b); };
    const uint32_t w_bar = bar_s + 8u * (2 * TC_MAX_SLOTS + 2 * TC_NBLK);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(gbase + bar_off + 8 * (2 * TC_MAX_SLOTS + 2 * TC_NBLK + 1));
    // pipelined mode: progress words of the 8 epilogue warps (2 + CTA-local index of the last row stored), and the minima the
    // poller warp keeps fresh: rows available in the input ring (bands b-1..b+1 of the gating stages) and rows the readers of
    // this stage's output ring slots have finished
    volatile uint32_t* s_prog = reinterpret_cast<volatile uint32_t*>(gbase + bar_off + 8 * (2 * TC_MAX_SLOTS + 2 * TC_NBLK + 2));
    volatile uint32_t* s_avail = s_prog + 4 * TC_NSETS;
    volatile uint32_t* s_bpmin = s_prog + 4 * TC_NSETS + 1;
    uint32_t* s_finished = const_cast<uint32_t*>(s_prog) + 4 * TC_NSETS + 2;
    static_assert((4 * TC_NSETS + 3) * 4 <= TCG_PIPE_WORDS * 8, "pipelined-mode words do not fit their slots");
    constexpr int PAIR_BASE = 2 * TC_MAX_SLOTS + 2 * TC_NBLK + 2 + TCG_PIPE_WORDS;
    [[maybe_unused]] auto pfull_bar = [&](int s_) { return bar_s + 8u * (PAIR_BASE + s_); };                    // (leader) the peer's row has landed
    [[maybe_unused]] auto ptempty_bar = [&](int b_) { return bar_s + 8u * (PAIR_BASE + TC_MAX_SLOTS + b_); };     // (leader) the peer drained its block
    [[maybe_unused]] const uint32_t pw_bar = bar_s + 8u * (PAIR_BASE + TC_MAX_SLOTS + TC_NBLK);                  // (leader) the peer's weights are loaded
    const bool ring_in = PIPE && P.n_in > 0, ring_out = PIPE && (P.out16_ring || P.out32_ring);
    const uint32_t RR = PIPE ? (uint32_t)P.RR : 1u;
    const int nb_lo = band > 0 ? band - 1 : 0, nb_hi = PIPE && band + 1 < P.nb ? band + 1 : (PIPE ? P.nb - 1 : 0);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr bool PLAIN = NRES == 0 && OUTS == 1;
    static_assert(!SC || PLAIN, "the fused shortcut is implemented for the plain (fp16-only, no residual) epilogue");
    constexpr int NPRE = NRES >= 0 ? (NRES > 0 ? NRES : 1) : (NOUT <= 32 ? 2 : 1);  // residual terms prefetched into registers
    static_assert(NRES < 0 || NOUT * NRES <= (RF16 ? 128 : 64), "prefetched residuals do not fit the register budget");
    const int nres = NRES >= 0 ? NRES : P.nres;
    const bool has16 = OUTS ? (OUTS & 1) != 0 : out16 != nullptr, has32 = OUTS ? (OUTS & 2) != 0 : out32 != nullptr;

    if (threadIdx.x == 0) {
        for (int s = 0; s < R; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), P.pair ? 2 : 1);  // paired: a slot is free once BOTH CTAs' MMAs have read their copy
        }
        for (int b = 0; b < NB; ++b) {
            mbar_init(tfull_bar(b), 1);
            mbar_init(tempty_bar(b), 4);
        }
        mbar_init(w_bar, 1);
        if constexpr (P2) {
            for (int s_ = 0; s_ < R; ++s_) mbar_init(pfull_bar(s_), 1);
            for (int b_ = 0; b_ < NB; ++b_) mbar_init(ptempty_bar(b_), 4);
            mbar_init(pw_bar, 1);
        }
        if constexpr (PIPE) {
            for (int w = 0; w < 4 * TC_NSETS; ++w) s_prog[w] = (uint32_t)(w >> 2);
            *s_avail = 0u;
            *s_bpmin = 0u;
            *s_finished = 0u;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if constexpr (P2) {  // one warp in EACH CTA of the pair performs the pair allocation
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"((uint32_t)C::TCOLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                         "r"((uint32_t)C::TCOLS)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    if (warp >= 2) {
        for (int i = threadIdx.x - 64; i < NOUT; i += TC_THREADS - 64) {
            s_bias[i] = bias_g[i];
            s_slope[i] = slope_g[i];
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (P2 || (!PIPE && P.pair)) cluster_sync_all();  // the peer's barriers exist before anything (multicast data, commits, arrives) can reach them
    const uint32_t tmem_base = *s_tmem;
    // pair mode: this CTA's band of a work item (items describe the band PAIR: x0 = first column of band 2 * pair)
    auto my_item = [&](int it_) {
        TcItem I_ = P.items[it_];
        if constexpr (P2) {
            I_.x0 += (int)prank * TC_BW;
            I_.w = max(0, min(TC_BW, I_.Wt - I_.x0));  // (an odd band count leaves the last pair's second CTA without columns: it still runs the protocol)
        }
        return I_;
    };

    if (warp == 0) {
        // ======================= TMA producer =======================
        if (lane == 0) {
            mbar_expect_tx(w_bar, WB);
            const uint32_t chunk = 3u * NOUT * TCG_PB;  // one (group, kx) tile
            if constexpr (P2) {  // this CTA's half of the rows of every tile (the swizzle pattern repeats every 8 rows = 1 KB: halves keep it)
                for (int t = 0; t < 3 * G; ++t) bulk_g2s(w_s + t * (chunk / 2), wimg + (size_t)t * chunk + prank * (chunk / 2), chunk / 2, w_bar);
                if constexpr (SC) bulk_g2s(w_s + 3 * G * (chunk / 2), wimg + (size_t)3 * G * chunk + prank * (C::SCB / 2), C::SCB / 2, w_bar);
            } else {
                for (int t = 0; t < 3 * G; ++t) bulk_g2s(w_s + t * chunk, wimg + (size_t)t * chunk, chunk, w_bar);
                if constexpr (SC) bulk_g2s(w_s + 3 * G * chunk, wimg + (size_t)3 * G * chunk, C::SCB, w_bar);
            }
            int slot = 0;
            uint32_t phase = 0;
            long long w_empty = 0, w_gate = 0;
            const long long t_begin = clock64();
            if constexpr (!PIPE) griddep_wait();  // everything above (barriers, TMEM, weights: constants) overlapped the previous launch's tail
            uint32_t seen = 0u;  // last value read from s_avail
            for (int it = it_begin; it < it_end; ++it) {
                const TcItem I = my_item(it);
                if (PIPE && I.w <= 0) continue;
                const CUtensorMap* map = P.maps + (P.map_base + I.map);
                [[maybe_unused]] const CUtensorMap* rmap = PIPE ? P.maps + (P.ring_map_base + I.map) : nullptr;
                const int rows_in = I.rows + 2;
                for (int rho = 0; rho < rows_in; ++rho) {
                    int y = I.y0 - 1 + rho;  // rows outside the plane are zero-filled by TMA = the conv's zero padding
                    if (!PIPE && P.flip) y = I.Ht - 1 - y;
                    [[maybe_unused]] int ry = -1;  // ring row of this input row (any out-of-bounds coordinate reads as zeros)
                    if constexpr (PIPE) {
                        if (ring_in && y >= 0 && y < I.Ht) {
                            const uint32_t gy = (uint32_t)(I.grow0 + y);
                            ry = (int)(gy % RR);
                            if (seen < gy + 1u && !P.nogate) {  // rows 0..gy of bands b-1, b, b+1 of the gating stages written and released?
                                const long long t0 = clock64();
                                while ((seen = *s_avail) < gy + 1u) {
                                    __nanosleep(20);
                                    if (clock64() - t0 > 20000000000LL) flag_timeout(P.done_in[0] + band * B2SR_FLAG_STRIDE, gy + 1u, 10);
                                }
                                __threadfence_block();       // pairs with the poller's fence before it stored s_avail
                                fence_proxy_async_global();  // the rows were written through the generic proxy; TMA reads them
                                w_gate += clock64() - t0;
                            }
                        }
                    }
                    for (int g = 0; g < G; ++g) {
                        mbar_wait_clocked(empty_bar(slot), phase ^ 1u, 0, w_empty);
                        if (P.ablate & 4) {
                            mbar_arrive(full_bar(slot));
                            if (++slot == R) slot = 0, phase ^= 1u;
                            continue;
                        }
                        mbar_expect_tx(full_bar(slot), TCG_SUBROWB);
                        if (PIPE && P.grp_ring[g])
                            tma_load_4d(ring_s + slot * TCG_SUBROWB, rmap, full_bar(slot), g * 64, I.x0 - 1, ry, 0);
                        else if (PIPE || !P.pair)
                            tma_load_4d(ring_s + slot * TCG_SUBROWB, map, full_bar(slot), g * 64, I.x0 - 1, y, I.plane);
                        else if (rank == 0)  // one load from L2 / HBM fills this slot in both CTAs (each armed its own barrier)
                            tma_load_4d_multicast(ring_s + slot * TCG_SUBROWB, map, full_bar(slot), g * 64, I.x0 - 1, y, I.plane, (uint16_t)3);
                        if (++slot == R) {
                            slot = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
            // all input rows of this CTA are requested: the next launch of the stream may start taking the SMs that CTAs
            // of this grid free (it does its own prologue, then waits for this grid to complete before touching buffers)
            if constexpr (!PIPE) griddep_launch_dependents();
            if (ring_in) atomicAdd(s_finished, 1u);
            if (P.dbg) {
                P.dbg[blockIdx.x * 16 + 4] = w_empty;
                P.dbg[blockIdx.x * 16 + 7] = clock64() - t_begin;
                P.dbg[blockIdx.x * 16 + 11] = w_gate;
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
      if constexpr (P2) {
        // ---- CTA-pair mode.  Every input row is issued with its FULL window: output rows y-1, y, y+1 of input row y, N = 3 * NOUT.
        // Row sequence of a work item [y0, y0 + rows): input rows y0-1 .. y0+rows (rho = 0 .. rows+1), output rows
        // y0-2 .. y0+rows+1 (the first and last two are phantom rows: accumulated, drained, never stored).
        const uint32_t desc_hi = ((8u * TCG_PB) >> 4) | (1u << 14) | (2u << 29);  // SBO = 1024 B | version | SWIZZLE_128B
        const uint64_t hi64 = (uint64_t)desc_hi << 32;
        const uint32_t w_lo = (w_s >> 4) | (1u << 16);
        const uint32_t ring_lo = (ring_s >> 4) | (1u << 16);
        constexpr uint32_t KXB = (3 * NOUT * TCG_PB / 2) >> 4;  // descriptor units between this CTA's half tiles of kx, kx + 1
        constexpr uint32_t GRPB = 3 * KXB;                      // ... between channel groups
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)((3 * NOUT) >> 3) << 17) | ((256u >> 4) << 24);  // D f32, A = B = f16, M = 256, N = 3 NOUT
        const int itb = __shfl_sync(0xffffffffu, it_begin, 0), ite = __shfl_sync(0xffffffffu, it_end, 0);
        int slot = 0;
        uint32_t phase = 0;
        if (!leader) {
            // the peer: forwards "my row has landed" (and, first, "my weights are loaded") to the leader's barriers
            if (lane == 0) {
                mbar_wait(w_bar, 0, 1);
                mbar_arrive_remote(pw_bar, 0u);
                for (int it = itb; it < ite; ++it) {
                    const int rows = P.items[it].rows;
                    for (int rg = 0; rg < (rows + 2) * G; ++rg) {
                        mbar_wait(full_bar(slot), phase, 2);
                        mbar_arrive_remote(pfull_bar(slot), 0u);
                        if (++slot == R) slot = 0, phase ^= 1u;
                    }
                }
            }
        } else {
            mbar_wait(w_bar, 0, 1);
            mbar_wait_cluster(pw_bar, 0, 1);
            uint32_t tmask = 0;  // bit h: parity of the next use of accumulator block h
            long long w_te = 0, w_pte = 0, w_fu = 0, w_pfu = 0;
            const long long t_begin = clock64();
            auto take_block = [&](uint32_t h) {  // block h drained and zeroed by the epilogue warps of BOTH CTAs?
                const long long t0 = P.dbg ? clock64() : 0;
                mbar_wait(tempty_bar(h), (tmask >> h) & 1u, 3);
                const long long t1 = P.dbg ? clock64() : 0;
                mbar_wait_cluster(ptempty_bar(h), (tmask >> h) & 1u, 3);
                if (P.dbg) w_te += t1 - t0, w_pte += clock64() - t1;
                tmask ^= 1u << h;
            };
            for (int it = itb; it < ite; ++it) {
                const int rows = __shfl_sync(0xffffffffu, P.items[it].rows, 0);
                const uint32_t y0 = (uint32_t)__shfl_sync(0xffffffffu, P.items[it].y0, 0);
                for (int rho = 0; rho < rows + 2; ++rho) {
                    // window = homes of output rows y0+rho-2, y0+rho-1, y0+rho (it may run into the extension blocks NB, NB+1)
                    const uint32_t home0 = (y0 + (uint32_t)rho + 2u * NB - 2u) % NB;
                    if (rho == 0) {
                        take_block(home0);
                        take_block((home0 + 1u) % NB);
                    }
                    take_block((home0 + 2u) % NB);
                    tc_fence_after();
                    const uint32_t d = tmem_base + home0 * NOUT;
                    if (elect_one_sync()) {
                        int sl = slot;
                        uint32_t ph = phase;
                        for (int g = 0; g < G; ++g) {
                            const long long t0 = P.dbg ? clock64() : 0;
                            mbar_wait(full_bar(sl), ph, 2);
                            const long long t1 = P.dbg ? clock64() : 0;
                            mbar_wait_cluster(pfull_bar(sl), ph, 2);
                            if (P.dbg) w_fu += t1 - t0, w_pfu += clock64() - t1;
                            tc_fence_after();
                            const int ks = min(4, (P.cin - g * 64) >> 4);  // K = 16 slabs present in this group
                            const uint64_t a0 = hi64 | (uint64_t)(ring_lo + (uint32_t)sl * (TCG_SUBROWB >> 4));
                            const uint64_t b0 = hi64 | (uint64_t)(w_lo + (uint32_t)g * GRPB);
                            if (ks == 4) {
#pragma unroll
                                for (int m = 0; m < 12; ++m) {
                                    const int kx = m >> 2, k = m & 3;
                                    umma2_f16(d, a0 + (uint64_t)((kx * TCG_PB + k * 32) >> 4), b0 + (uint64_t)(kx * KXB + ((k * 32) >> 4)), idesc, 1u);
                                }
                            } else {
                                for (int kx = 0; kx < 3; ++kx)
                                    for (int k = 0; k < ks; ++k)
                                        umma2_f16(d, a0 + (uint64_t)((kx * TCG_PB + k * 32) >> 4), b0 + (uint64_t)(kx * KXB + ((k * 32) >> 4)), idesc, 1u);
                            }
                            if constexpr (SC) {
                                // input row rho is the centre row of output row rho - 1 (real rows only): its first channels through
                                // the 1x1 weights (column tap kx = 1) into that row's shortcut block
                                if (g == 0 && rho >= 1 && rho <= rows) {
                                    const uint32_t ds = tmem_base + C::SC_COL0 + ((y0 + (uint32_t)rho - 1u) % NB) * NOUT;
                                    const uint64_t bs = hi64 | (uint64_t)(w_lo + (uint32_t)G * GRPB);
                                    constexpr uint32_t ids = (1u << 4) | ((uint32_t)(NOUT >> 3) << 17) | ((256u >> 4) << 24);
                                    for (int k = 0; k < P.sc_ks; ++k)
                                        umma2_f16(ds, a0 + (uint64_t)((TCG_PB + k * 32) >> 4), bs + (uint64_t)((k * 32) >> 4), ids, 1u);
                                }
                            }
                            umma2_commit_both(empty_bar(sl));  // this (row, group) slot may be refilled in both CTAs
                            if (++sl == R) sl = 0, ph ^= 1u;
                        }
                        umma2_commit_both(tfull_bar(home0));  // output row y0+rho-2 is complete
                        if (rho == rows + 1) {                // the range is over: its last two (phantom) rows are as complete as they get
                            umma2_commit_both(tfull_bar((home0 + 1u) % NB));
                            umma2_commit_both(tfull_bar((home0 + 2u) % NB));
                        }
                    }
                    __syncwarp();
                    slot += G;
                    if (slot >= R) slot -= R, phase ^= 1u;
                }
            }
            if (P.dbg) {  // (the elected lane holds the full / pfull waits, every lane the block waits)
                if (lane == 0) {
                    P.dbg[blockIdx.x * 16 + 0] = clock64() - t_begin;
                    P.dbg[blockIdx.x * 16 + 2] = w_te;
                    P.dbg[blockIdx.x * 16 + 3] = w_pte;
                }
                if (w_fu || w_pfu) {
                    P.dbg[blockIdx.x * 16 + 1] = w_fu;
                    P.dbg[blockIdx.x * 16 + 8] = w_pfu;
                }
            }
        }
      } else {
        // The whole warp runs the (warp-uniform) control flow, one elected lane issues.  Work unit: one channel group
        // of one input row; all groups of a row accumulate into the same window of accumulator blocks.
        const uint32_t desc_hi = ((8u * TCG_PB) >> 4) | (1u << 14) | (2u << 29);  // SBO = 1024 B | version | SWIZZLE_128B
        const uint64_t hi64 = (uint64_t)desc_hi << 32;
        const uint32_t w_lo = (w_s >> 4) | (1u << 16);
        const uint32_t ring_lo = (ring_s >> 4) | (1u << 16);
        constexpr uint32_t KXB = (3 * NOUT * TCG_PB) >> 4;  // descriptor units between the stacked tiles of kx, kx + 1
        constexpr uint32_t BLKB = (NOUT * TCG_PB) >> 4;     // ... between the ky blocks inside one tile
        constexpr uint32_t GRPB = 3 * KXB;                  // ... between channel groups
        const int kx_lo = P.k1 ? 1 : 0, kx_hi = P.k1 ? 2 : 3;
        int slot = 0;
        uint32_t phase = 0;
        long long w_full = 0, w_tempty = 0, t_first = 0, t_mma = 0, t_commit = 0;
        const long long t_begin = clock64();
        mbar_wait(w_bar, 0, 1);
        uint32_t ok_tempty = mbar_test_wait(tempty_bar(0), 0);
        uint32_t ok_full = 0u;  // the first group of the next input row was already seen in shared memory (probed behind the MMAs)
        // Loop bounds that come from global memory are broadcast with a shuffle: the compiler then knows they are
        // warp-uniform and keeps the whole descriptor arithmetic in uniform registers (UTCHMMA takes uniform operands;
        // with per-thread values every MMA cost five R2UR moves).
        const int itb = __shfl_sync(0xffffffffu, it_begin, 0), ite = __shfl_sync(0xffffffffu, it_end, 0);
        uint32_t tmask = 0;  // bit h: parity of the next use of accumulator block h (tempty barrier phase)
        for (int it = itb; it < ite; ++it) {
            if (PIPE && __shfl_sync(0xffffffffu, P.items[it].w, 0) <= 0) continue;  // band absent from this plane
            const int rows = __shfl_sync(0xffffffffu, P.items[it].rows, 0);
            // Homes follow the plane row (y % NB), not the CTA's row count: which rows are summed from two blocks
            // then does not depend on how the launch was cut into CTA ranges, so a frame computes bit-identically
            // alone and inside a batch.
            const uint32_t y0 = (uint32_t)__shfl_sync(0xffffffffu, P.items[it].y0, 0);
            for (int rho = 0; rho < rows + 2; ++rho) {  // input row rho feeds output rows rho - ky, ky = 0..2
                const bool fresh = rho < rows;
                const int t0 = rho >= 2 ? rho - 2 : 0;
                const int t1 = fresh ? rho : rows - 1;
                const int cnt = t1 - t0 + 1;
                // The window of an input row starts at the home of output row rho - 2 and spans three blocks (it may reach
                // blocks NB, NB + 1); at the top of an item the rows above the item are simply left out (brow0 skips
                // their blocks), so every output row collects its three contributions in the same physical blocks
                // wherever the item boundaries fall.
                const uint32_t home0 = (y0 + (uint32_t)rho + (uint32_t)NB - 2u) % NB;
                const uint32_t hf = (y0 + (uint32_t)rho) % NB;    // home of the output row that starts with this input row
                const uint32_t brow0 = rho >= 2 ? 0u : (uint32_t)(2 - rho);
                const uint32_t idesc = C::IDESC0 | ((uint32_t)((cnt * NOUT) >> 3) << 17);
                const uint32_t d = tmem_base + (home0 + brow0) * NOUT;
                // the new output row's block drained and zeroed?  all channel groups of this input row in shared memory?
                if (fresh) {
                    if (!ok_tempty) mbar_wait_clocked(tempty_bar(hf), (tmask >> hf) & 1u, 3, w_tempty);
                    tmask ^= 1u << hf;
                }
                if (!ok_full) mbar_wait_clocked(full_bar(slot), phase, 2, w_full);  // (the elected lane waits for the later groups itself)
                tc_fence_after();
                if (P.dbg && t_first == 0) t_first = clock64() - t_begin;
                {  // probe for the next row of this item, latency hidden behind the MMAs
                    const uint32_t hn = (hf + 1u) % NB;
                    ok_tempty = rho + 1 < rows ? mbar_test_wait(tempty_bar(hn), (tmask >> hn) & 1u) : 0u;
                    int ns = slot + G;
                    uint32_t np = phase;
                    if (ns >= R) ns -= R, np ^= 1u;
                    ok_full = mbar_test_wait(full_bar(ns), np);  // a completed phase stays completed until this warp consumes it
                }
                const long long tq0 = P.dbg ? clock64() : 0;
                if (elect_one_sync()) {
                    int sl = slot;
                    uint32_t ph = phase;
                    for (int g = 0; g < G; ++g) {
                        if (g > 0) {  // group g of this row in shared memory?  (a single thread may wait on an mbarrier)
                            mbar_wait(full_bar(sl), ph, 2);
                            tc_fence_after();
                        }
                        const int ks = min(4, (P.cin - g * 64) >> 4);  // K = 16 slabs present in this group
                        const uint64_t a0 = hi64 | (uint64_t)(ring_lo + (uint32_t)sl * (TCG_SUBROWB >> 4));
                        const uint64_t b0 = hi64 | (uint64_t)(w_lo + (uint32_t)g * GRPB + brow0 * BLKB);
                        if (P.ablate & 1) {
                        } else if (ks == 4 && !P.k1) {  // the common case, straight-line: 12 MMAs, one 64-bit add per descriptor
#pragma unroll
                            for (int m = 0; m < 12; ++m) {
                                const int kx = m >> 2, k = m & 3;
                                umma_f16(d, a0 + (uint64_t)((kx * TCG_PB + k * 32) >> 4), b0 + (uint64_t)(kx * KXB + ((k * 32) >> 4)), idesc, 1u);
                            }
                        } else {
                            for (int kx = kx_lo; kx < kx_hi; ++kx)
                                for (int k = 0; k < ks; ++k)
                                    umma_f16(d, a0 + (uint64_t)((kx * TCG_PB + k * 32) >> 4), b0 + (uint64_t)(kx * KXB + ((k * 32) >> 4)), idesc, 1u);
                        }
                        if constexpr (SC) {
                            // input row rho is the centre row of output row rho - 1: its first channels through the 1x1 weights
                            // (column tap kx = 1) into that row's shortcut block
                            if (g == 0 && rho >= 1 && rho <= rows && !(P.ablate & 1)) {
                                const uint32_t ds = tmem_base + C::SC_COL0 + ((y0 + (uint32_t)rho - 1u) % NB) * NOUT;
                                const uint64_t bs = hi64 | (uint64_t)(w_lo + (uint32_t)G * GRPB);
                                constexpr uint32_t ids = C::IDESC0 | ((uint32_t)(NOUT >> 3) << 17);
                                for (int k = 0; k < P.sc_ks; ++k)
                                    umma_f16(ds, a0 + (uint64_t)((TCG_PB + k * 32) >> 4), bs + (uint64_t)((k * 32) >> 4), ids, 1u);
                            }
                        }
                        // this (row, group) slot may be refilled once its MMAs have completed (paired: in both CTAs)
                        if (!PIPE && P.pair) umma_commit_multicast(empty_bar(sl), (uint16_t)3); else umma_commit(empty_bar(sl));
                        if (++sl == R) sl = 0, ph ^= 1u;
                    }
                    if (rho >= 2) umma_commit(tfull_bar(home0));  // output row rho - 2 is complete (home0 is its home)
                    if (P.dbg) t_mma += clock64() - tq0;
                }
                __syncwarp();
                slot += G;
                if (slot >= R) slot -= R, phase ^= 1u;
            }
        }
        if (P.dbg && lane == 0) {
            P.dbg[blockIdx.x * 16 + 0] = clock64() - t_begin;
            P.dbg[blockIdx.x * 16 + 1] = w_full;
            P.dbg[blockIdx.x * 16 + 2] = w_tempty;
            P.dbg[blockIdx.x * 16 + 3] = t_first;
        }
        if (P.dbg && t_mma) {  // (the elected lane holds these)
            P.dbg[blockIdx.x * 16 + 8] = t_mma;
            P.dbg[blockIdx.x * 16 + 9] = t_commit;
        }
      }
    } else if (warp < 2 + 4 * TC_NSETS) {
        // ======================= epilogue =======================
        const int q = warp & 3;
        const uint32_t set = (warp - 2) >> 2;
        uint32_t tile_cnt = 0, emask = 0;
        for (uint32_t b = set; b < (uint32_t)NB; b += TC_NSETS) {  // hand every block of this set to the issuer zeroed
#pragma unroll
            for (int j = 0; j < NOUT; j += 16) {
                tmem_st16_zero(tmem_base + ((uint32_t)(q * 32) << 16) + b * NOUT + j);
                if (b < 2) tmem_st16_zero(tmem_base + ((uint32_t)(q * 32) << 16) + (NB + b) * NOUT + j);  // its extension block
                if constexpr (SC) tmem_st16_zero(tmem_base + ((uint32_t)(q * 32) << 16) + C::SC_COL0 + b * NOUT + j);
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (P2 && !leader) mbar_arrive_remote(ptempty_bar(b), 0u); else mbar_arrive(tempty_bar(b));
            }
        }
        const float scale_acc = P.acc_scale;
        const int c = q * 32 + lane;
        if constexpr (!PIPE) griddep_wait();  // residual buffers are read and output buffers written only after the previous launch has completed
        long long w_tfull = 0, t_tmem = 0, w_bp = 0;
        const long long t_begin = clock64();
        [[maybe_unused]] uint32_t bp_seen = 0u;  // last value read from s_bpmin
        [[maybe_unused]] uint32_t av_seen = 0u;  // last value read from s_avail (gate of ring residual prefetches)
        for (int it = it_begin; it < it_end; ++it) {
            const TcItem I = my_item(it);
            if (PIPE && I.w <= 0) continue;
            const bool valid = c < I.w;
            for (int t = P2 ? -2 : 0; t < I.rows + (P2 ? 2 : 0); ++t, ++tile_cnt) {
                const uint32_t buf = (uint32_t)(I.y0 + t + 2 * NB) % NB;  // home block of this output row
                const uint32_t par = (emask >> buf) & 1u;        // parity of this use of the block (both sets' rows count)
                emask ^= 1u << buf;
                if (tile_cnt % TC_NSETS != set) continue;
                if constexpr (P2) {
                    if (t < 0 || t >= I.rows) {
                        // phantom row (pair mode issues full windows only): wait until its partial sums are final, throw them
                        // away -- the block (and its extension block) must be zero for the next row that lives there
                        mbar_wait_clocked(tfull_bar(buf), par, 4, w_tfull);
                        tc_fence_after();
                        const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + buf * NOUT;
#pragma unroll
                        for (int j = 0; j < NOUT; j += 16) {
                            tmem_st16_zero(taddr0 + j);
                            if (buf < 2) tmem_st16_zero(taddr0 + (uint32_t)(NB * NOUT) + j);
                        }
                        tmem_wait_st();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (!leader) mbar_arrive_remote(ptempty_bar(buf), 0u); else mbar_arrive(tempty_bar(buf));
                        }
                        continue;
                    }
                }
                // residual terms of this thread's pixel are fetched BEFORE the wait for the accumulator, so their
                // latency (measured: ~700 cycles per dependent chunk, 4-8 chunks per row) hides behind the MMAs
                [[maybe_unused]] long long pix = -1;
                [[maybe_unused]] uint4 rraw[PLAIN ? 1 : NPRE][PLAIN ? 1 : NOUT / 4];
                const int yy = (!PIPE && P.flip) ? I.Ht - 1 - (I.y0 + t) : I.y0 + t;  // plane row of this output row
                if constexpr (MODE == 0) pix = valid ? (long long)I.pix_off + (long long)yy * I.Wt + I.x0 + c : -1;
                // pipelined mode: a buffer that lives in a ring is addressed by (global row mod RR, column); which of the
                // residual sources / outputs are rings is a property of the stage
                [[maybe_unused]] const uint32_t grow = PIPE ? (uint32_t)(I.grow0 + yy) : 0u;
                [[maybe_unused]] const long long rpix = (PIPE && valid) ? (long long)(grow % RR) * P.Wmax + I.x0 + c : -1;
                const long long pix16 = (PIPE && P.out16_ring) ? rpix : pix, pix32 = (PIPE && P.out32_ring) ? rpix : pix;
                const long long pixr[2] = {(PIPE && P.res_ring[0]) ? rpix : pix, (PIPE && P.res_ring[1]) ? rpix : pix};
                if constexpr (PIPE && MODE == 0 && !PLAIN) {
                    // A residual that lives in a ring is prefetched long before this row's accumulator is complete, so the
                    // prefetch needs its own gate: the row was written by the gating op p* itself (published once done >= grow + 1)
                    // or by an op p* reads through its ring inputs (published before p* could finish row grow - 1).
                    if ((P.res_ring[0] || P.res_ring[1]) && av_seen < grow + 1u && !P.nogate) {
                        const long long t0 = clock64();
                        while ((av_seen = *s_avail) < grow + 1u) {
                            __nanosleep(20);
                            if (clock64() - t0 > 20000000000LL) flag_timeout(P.done_in[0] + band * B2SR_FLAG_STRIDE, grow + 1u, 25);
                        }
                        __threadfence_block();
                    }
                }
                if constexpr (MODE == 0 && !PLAIN) {
#pragma unroll
                    for (int r = 0; r < NPRE; ++r) {
                        if (r >= nres || pix < 0 || (P.ablate & 10)) continue;
                        // (pipelined mode reads residuals with ld.global.cg: a ring slot is rewritten by another SM during
                        // the launch, so a stale L1 line must never be hit)
                        if (NRES >= 0 ? !RF16 : P.res_f32[r] != 0) {
                            const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(res_ptr[r]) + pixr[r] * P.res_ld[r]);
#pragma unroll
                            for (int j = 0; j < NOUT / 4; ++j) rraw[r][j] = PIPE ? __ldcg(rp + j) : rp[j];
                        } else {
                            const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(res_ptr[r]) + pixr[r] * P.res_ld[r]);
#pragma unroll
                            for (int j = 0; j < NOUT / 8; ++j) rraw[r][j] = PIPE ? __ldcg(rp + j) : rp[j];
                        }
                    }
                }
                mbar_wait_clocked(tfull_bar(buf), par, 4, w_tfull);
                tc_fence_after();
                const long long tq0 = P.dbg ? clock64() : 0;
                uint32_t acc[NOUT];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * NOUT;
#pragma unroll
                for (int j = 0; j < NOUT; j += 16) tmem_ld16(taddr + j, acc + j);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < NOUT; j += 16) tmem_st16_zero(taddr + j);
                if (buf < 2) {  // homes 0 and 1: part of the sum sits in the extension block (windows that started at NB-2, NB-1)
                    const uint32_t xaddr = taddr + (uint32_t)(NB * NOUT);
#pragma unroll
                    for (int j0 = 0; j0 < NOUT; j0 += 16) {
                        uint32_t ext[16];
                        tmem_ld16(xaddr + j0, ext);
                        tmem_wait_ld();
#pragma unroll
                        for (int e = 0; e < 16; ++e) acc[j0 + e] = __float_as_uint(__uint_as_float(acc[j0 + e]) + __uint_as_float(ext[e]));
                        tmem_st16_zero(xaddr + j0);
                    }
                }
                [[maybe_unused]] uint32_t sacc[SC ? NOUT : 1];
                if constexpr (SC) {
                    const uint32_t saddr = tmem_base + ((uint32_t)(q * 32) << 16) + C::SC_COL0 + buf * NOUT;
#pragma unroll
                    for (int j = 0; j < NOUT; j += 16) tmem_ld16(saddr + j, sacc + j);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < NOUT; j += 16) tmem_st16_zero(saddr + j);
                }
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (P2 && !leader) mbar_arrive_remote(ptempty_bar(buf), 0u); else mbar_arrive(tempty_bar(buf));
                }
                if (P.dbg) t_tmem += clock64() - tq0;
                if (!PIPE && (P.ablate & 2)) continue;
                if constexpr (PIPE) {
                    // The ring slot of this row still holds row grow - RR.  Input row i is read by output rows i-1, i, i+1 of
                    // the stages that read this ring (and by their epilogues as a residual of row i): all of that is over
                    // once those stages have FINISHED rows 0..i+1, i.e. done >= i + 2, on bands b-1..b+1.
                    if (ring_out && grow + 2u > RR && bp_seen < grow + 2u - RR && !P.nogate) {
                        const long long t0 = clock64();
                        while ((bp_seen = *s_bpmin) < grow + 2u - RR) {
                            __nanosleep(20);
                            if (clock64() - t0 > 20000000000LL) flag_timeout(P.bp[0] + band * B2SR_FLAG_STRIDE, grow + 2u - RR, 20);
                        }
                        __threadfence_block();  // pairs with the poller's fence before it stored s_bpmin
                        w_bp += clock64() - t0;
                    }
                    __syncwarp();
                }

                if constexpr (MODE == 0 && PLAIN) {
                    // bias + LeakyReLU -> fp16, via swizzled per-warp staging, then 128-bit coalesced global stores
                    constexpr int CH = C::OB / 16;
                    uint4* stg = reinterpret_cast<uint4*>(gbase + stg_off + (warp - 2) * (32 * C::OB));
                    const float4* sb4 = reinterpret_cast<const float4*>(s_bias);
                    const float4* ss4 = reinterpret_cast<const float4*>(s_slope);
#pragma unroll
                    for (int j = 0; j < NOUT; j += 8) {
                        const float4 b0 = sb4[j >> 2], b1 = sb4[(j >> 2) + 1], l0 = ss4[j >> 2], l1 = ss4[(j >> 2) + 1];
                        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                        const float sl[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
                        uint32_t pk[4];
#pragma unroll
                        for (int e = 0; e < 8; e += 2) {
                            float v0 = fmaf(__uint_as_float(acc[j + e]), scale_acc, bb[e]);
                            float v1 = fmaf(__uint_as_float(acc[j + e + 1]), scale_acc, bb[e + 1]);
                            v0 = v0 < 0.f ? v0 * sl[e] : v0;
                            v1 = v1 < 0.f ? v1 * sl[e + 1] : v1;
                            if constexpr (SC) {
                                v0 = v0 * P.sc_cv + __uint_as_float(sacc[j + e]) * P.sc_cr;
                                v1 = v1 * P.sc_cv + __uint_as_float(sacc[j + e + 1]) * P.sc_cr;
                            }
                            __half2 h = __floats2half2_rn(v0, v1);
                            pk[e >> 1] = *reinterpret_cast<uint32_t*>(&h);
                        }
                        const int qi = lane * CH + (j >> 3);
                        stg[qi ^ ((qi >> 3) & 7)] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                    __syncwarp();
                    uint8_t* outp = reinterpret_cast<uint8_t*>(out16);
#pragma unroll
                    for (int i = 0; i < CH; ++i) {
                        const int qi = i * 32 + lane;
                        const uint4 v4 = stg[qi ^ ((qi >> 3) & 7)];
                        const long long o = __shfl_sync(0xffffffffu, pix16, qi / CH);
                        if (o >= 0 && !(P.ablate & 8)) *reinterpret_cast<uint4*>(outp + (size_t)o * P.out16_ld * 2 + (qi % CH) * 16) = v4;
                    }
                    __syncwarp();
                } else if constexpr (MODE == 0) {
                    constexpr int CH = C::OB / 16;
                    uint4* stg = reinterpret_cast<uint4*>(gbase + stg_off + (warp - 2) * (32 * C::OB));
                    const float4* sb4 = reinterpret_cast<const float4*>(s_bias);
                    const float4* ss4 = reinterpret_cast<const float4*>(s_slope);
                    // v = act(acc * scale + bias), then the residual terms; the results replace acc[] (as float bits)
#pragma unroll
                    for (int j = 0; j < NOUT; j += 4) {
                        const float4 b4 = sb4[j >> 2], l4 = ss4[j >> 2];
                        const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, sl[4] = {l4.x, l4.y, l4.z, l4.w};
                        float v[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float a = fmaf(__uint_as_float(acc[j + e]), scale_acc, bb[e]);
                            v[e] = a < 0.f ? a * sl[e] : a;
                        }
#pragma unroll
                        for (int r = 0; r < 2; ++r) {
                            if (r >= nres) continue;
                            float rv[4] = {0.f, 0.f, 0.f, 0.f};
                            if (pix >= 0) {
                                if (r < NPRE) {
                                    if (NRES >= 0 ? !RF16 : P.res_f32[r] != 0) {
                                        const uint4 u = rraw[r][j >> 2];
                                        rv[0] = __uint_as_float(u.x), rv[1] = __uint_as_float(u.y), rv[2] = __uint_as_float(u.z), rv[3] = __uint_as_float(u.w);
                                    } else {
                                        const uint4 u = rraw[r][j >> 3];
                                        const uint32_t w0 = (j & 4) ? u.z : u.x, w1 = (j & 4) ? u.w : u.y;
                                        const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&w0)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&w1));
                                        rv[0] = f0.x, rv[1] = f0.y, rv[2] = f1.x, rv[3] = f1.y;
                                    }
                                } else if (P.res_f32[r]) {  // (a second residual of a 64-channel launch: not prefetched)
                                    const float4 u = __ldcg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(res_ptr[r]) + pixr[r] * P.res_ld[r] + j));
                                    rv[0] = u.x, rv[1] = u.y, rv[2] = u.z, rv[3] = u.w;
                                } else {
                                    const uint2 u = __ldcg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(res_ptr[r]) + pixr[r] * P.res_ld[r] + j));
                                    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
                                    rv[0] = f0.x, rv[1] = f0.y, rv[2] = f1.x, rv[3] = f1.y;
                                }
                            }
                            const float cv = P.coef_v[r], cr = P.coef_r[r];
#pragma unroll
                            for (int e = 0; e < 4; ++e) v[e] = v[e] * cv + rv[e] * cr;
                        }
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[j + e] = __float_as_uint(v[e]);
                    }
                    // Stores go through this warp's swizzled staging tile (32 pixels x OB bytes) so that every store
                    // instruction writes whole 16-byte chunks of consecutive pixels: the fp32 copy in two passes of
                    // NOUT / 2 channels (per-thread float4 stores to 32 different lines per instruction were measured to
                    // cost ~1700 cycles per row), then the fp16 copy in one.
                    if (has32) {
                        uint8_t* outp = reinterpret_cast<uint8_t*>(out32);
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            __syncwarp();
#pragma unroll
                            for (int i = 0; i < CH; ++i) {
                                const int qi = lane * CH + i, j = half * (NOUT / 2) + i * 4;
                                stg[qi ^ ((qi >> 3) & 7)] = make_uint4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
                            }
                            __syncwarp();
#pragma unroll
                            for (int i = 0; i < CH; ++i) {
                                const int qi = i * 32 + lane;
                                const uint4 v4 = stg[qi ^ ((qi >> 3) & 7)];
                                const long long o = __shfl_sync(0xffffffffu, pix32, qi / CH);
                                if (o >= 0 && !(P.ablate & 8)) *reinterpret_cast<uint4*>(outp + ((size_t)o * P.out32_ld + half * (NOUT / 2)) * 4 + (qi % CH) * 16) = v4;
                            }
                        }
                    }
                    if (has16) {
                        uint8_t* outp = reinterpret_cast<uint8_t*>(out16);
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < CH; ++i) {
                            const int qi = lane * CH + i, j = i * 8;
                            uint32_t pk[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                __half2 h = __floats2half2_rn(__uint_as_float(acc[j + 2 * e]), __uint_as_float(acc[j + 2 * e + 1]));
                                pk[e] = *reinterpret_cast<uint32_t*>(&h);
                            }
                            stg[qi ^ ((qi >> 3) & 7)] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        }
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < CH; ++i) {
                            const int qi = i * 32 + lane;
                            const uint4 v4 = stg[qi ^ ((qi >> 3) & 7)];
                            const long long o = __shfl_sync(0xffffffffu, pix16, qi / CH);
                            if (o >= 0 && !(P.ablate & 8)) *reinterpret_cast<uint4*>(outp + (size_t)o * P.out16_ld * 2 + (qi % CH) * 16) = v4;
                        }
                    }
                    __syncwarp();
                } else {
                    // network output: (acc + bias) * 255, cropped to the tile core, as cv2.imwrite would store it
                    const int fy = I.fy0 + yy, fx = I.fx0 + I.x0 + c;
                    if (valid && fy >= I.cy0 && fy < I.cy1 && fx >= I.cx0 && fx < I.cx1) {
                        const size_t o = (((size_t)I.frame * P.frame_h + fy) * P.frame_w + fx) * 3;
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {
                            float a = fmaf(__uint_as_float(acc[ch]), scale_acc, s_bias[ch]);
                            a = a < 0.f ? a * s_slope[ch] : a;
                            a *= 255.f;
                            if constexpr (F32OUT) {
                                reinterpret_cast<float*>(P.frames_out)[o + ch] = a;
                            } else {
                                const int iv = __float2int_rn(a);  // round half to even, like cv2's saturate_cast
                                reinterpret_cast<uint8_t*>(P.frames_out)[o + ch] = (uint8_t)min(max(iv, 0), 255);
                            }
                        }
                    }
                }
                if constexpr (PIPE) {
                    // report the row to the publisher with CTA-scope ordering only; its single GPU-scope release covers every
                    // warp's stores by cumulativity
                    __syncwarp();
                    if (lane == 0) {
                        __threadfence_block();
                        s_prog[warp - 2] = tile_cnt + 2u;
                    }
                    __syncwarp();
                }
            }
        }
        if constexpr (PIPE) if (lane == 0) atomicAdd(s_finished, 1u);
        if (P.dbg && warp == 2 && lane == 0) {
            P.dbg[blockIdx.x * 16 + 5] = w_tfull;
            P.dbg[blockIdx.x * 16 + 6] = clock64() - t_begin;
            P.dbg[blockIdx.x * 16 + 10] = t_tmem;
            P.dbg[blockIdx.x * 16 + 12] = w_bp;
        }
    } else if constexpr (PIPE) {
      if (warp == 2 + 4 * TC_NSETS) {
        // ======================= publisher (pipelined mode) =======================
        // Turns the epilogue warps' progress words into one monotonic "rows of this (stage, band) written and released"
        // counter in global memory; it polls, so a slow release only batches several rows into one update.
        if (lane == 0) {
            static_assert(TC_NSETS == 2, "the progress-word arithmetic below assumes two epilogue sets");
            uint32_t n_local = 0;
            for (int it = it_begin; it < it_end; ++it)
                if (P.items[it].w > 0) n_local += (uint32_t)P.items[it].rows;
            const uint32_t g_end = it_end > it_begin ? (uint32_t)(P.items[it_end - 1].grow0 + P.items[it_end - 1].Ht) : 0u;
            uint32_t cnt = 0, t = 0, published = 0;
            int it = it_begin;
            const long long t_start = clock64();
            for (;;) {
                uint32_t m = n_local;  // rows [0, m) are complete: every warp is past them
#pragma unroll
                for (int w = 0; w < 4 * TC_NSETS; ++w) {
                    const uint32_t v = s_prog[w];
                    m = v < m ? v : m;
                }
                uint32_t adv = m - cnt;
                while (it < it_end) {  // move (it, t) forward by `adv` rows, stepping over absent bands and finished items
                    const TcItem* Ip = P.items + it;
                    const uint32_t rows = Ip->w > 0 ? (uint32_t)Ip->rows : 0u;
                    if (t + adv < rows) {
                        t += adv;
                        adv = 0;
                        break;
                    }
                    adv -= rows - t;
                    ++it;
                    t = 0;
                }
                cnt = m;
                const uint32_t g = it < it_end ? (uint32_t)(P.items[it].grow0 + P.items[it].y0) + t : g_end;
                if (g > published) {
                    __threadfence_block();  // acquire side of the progress words ...
                    st_release_gpu(P.done_out + band * B2SR_FLAG_STRIDE, g);  // ... then one GPU-scope release for all eight warps' stores
                    published = g;
                }
                if (it >= it_end) break;
                __nanosleep(40);
                if (clock64() - t_start > 20000000000LL) flag_timeout(P.done_out + band * B2SR_FLAG_STRIDE, g_end, 30);
            }
        }
      } else if (warp == 3 + 4 * TC_NSETS) {
        // ======================= counter poller (pipelined mode) =======================
        // Keeps shared-memory minima of the neighbours' global counters fresh (acquire loads, off every critical path).
        if ((ring_in || ring_out) && lane == 0) {
            const uint32_t need_fin = (ring_in ? 1u : 0u) + (ring_out ? 4u * TC_NSETS : 0u);
            const long long t_start = clock64();
            while (*reinterpret_cast<volatile uint32_t*>(s_finished) < need_fin) {
                // all counters are read with independent relaxed loads (one L2 round trip for the lot), then ONE acquire
                // fence orders them before the shared-memory words the other warps act on
                uint32_t ma = 0xffffffffu, mb = 0xffffffffu;
                if (ring_in) {
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        if (k < P.n_in) {
                            const uint32_t v0 = ld_relaxed_gpu(P.done_in[k] + nb_lo * B2SR_FLAG_STRIDE), v1 = ld_relaxed_gpu(P.done_in[k] + band * B2SR_FLAG_STRIDE),
                                           v2 = ld_relaxed_gpu(P.done_in[k] + nb_hi * B2SR_FLAG_STRIDE);
                            const uint32_t m = v0 < v1 ? (v0 < v2 ? v0 : v2) : (v1 < v2 ? v1 : v2);
                            ma = m < ma ? m : ma;
                        }
                }
                if (ring_out) {
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        if (k < P.n_bp) {
                            const uint32_t v0 = ld_relaxed_gpu(P.bp[k] + nb_lo * B2SR_FLAG_STRIDE), v1 = ld_relaxed_gpu(P.bp[k] + band * B2SR_FLAG_STRIDE),
                                           v2 = ld_relaxed_gpu(P.bp[k] + nb_hi * B2SR_FLAG_STRIDE);
                            const uint32_t m = v0 < v1 ? (v0 < v2 ? v0 : v2) : (v1 < v2 ? v1 : v2);
                            mb = m < mb ? m : mb;
                        }
                }
                fence_acq_rel_gpu();
                __threadfence_block();
                if (ring_in) *s_avail = ma;
                if (ring_out) *s_bpmin = mb;
                __nanosleep(64);
                if (clock64() - t_start > 40000000000LL) flag_timeout(ring_in ? P.done_in[0] : P.bp[0], 0xffffffffu, 40);
            }
        }
      }
    }

    tc_fence_before();
    __syncthreads();
    if (P2 || (!PIPE && P.pair)) cluster_sync_all();  // the peer's last commits / arrives / multicast writes target this CTA's shared memory: stay until it is done
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        if constexpr (P2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TCOLS) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TCOLS)
                         : "memory");
    }
}

// one launch per convolution (or per 2-CTA-cluster pair of its two halves)
template <int NOUT, int MODE, bool F32OUT, int NRES, int OUTS, bool RF16 = false, bool SC = false>
__global__ void __launch_bounds__(TC_THREADS, 1) tcg_conv_kernel(const __grid_constant__ TcgParams P) {
    extern __shared__ uint8_t smem_raw[];
    // paired launch: the two CTAs of a cluster walk the same row range; rank r computes output-channel half r
    const uint32_t rank = P.pair ? cluster_ctarank() : 0u;
    const int range = P.pair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    tcg_body<NOUT, MODE, F32OUT, NRES, OUTS, RF16, SC, false>(P, P.item_first[range], P.item_first[range + 1], 0, rank, smem_raw);
}

// CTA-pair form (cta_group::2): clusters of two CTAs = two neighbouring 128-column bands over the same rows (work items describe band pairs)
template <int NOUT, int NRES, int OUTS, bool RF16 = false, bool SC = false>
__global__ void __launch_bounds__(TC_THREADS, 1) tcg_pair2_kernel(const __grid_constant__ TcgParams P) {
    extern __shared__ uint8_t smem_raw[];
    const int range = (int)(blockIdx.x >> 1);
    tcg_body<NOUT, 0, false, NRES, OUTS, RF16, SC, false, true>(P, P.item_first[range], P.item_first[range + 1], 0, 0u, smem_raw);
}

// ------------------------------------------------------------------------------------------------
// pipelined mode: ONE persistent (cooperative) launch for a segment of consecutive convolutions -- an RRDB of
// 4x_Valar_v1 is 15 convolutions = 18 stages (each 192 -> 64 convolution is two 32-channel stages) x 8 bands = 144 CTAs.
// CTA (s, b) keeps stage s's weights resident and streams band b of every plane of the pass top to bottom.  The
// dense-block buffers [x | x1 | x2 | x3 | x4] and the fp32 trunk copies between the blocks never leave L2: each is a
// ring of RR rows; a stage pulls input rows with TMA as soon as the stage that wrote the newest slice of its input view
// has published them (`done` counters of bands b-1..b+1), and overwrites a ring slot once the last readers of that
// ring have finished with the row it held.  Only the segment's input (x, fp32 x) and output go through HBM.
// Accumulator homes follow plane rows exactly as in the per-launch kernel, so results are bit-identical to it.
// ------------------------------------------------------------------------------------------------
#define B2SR_TCG_VARIANT_PLAIN 0   // x1, x3: lrelu(conv + b) -> fp16
#define B2SR_TCG_VARIANT_SC 1      // x2: lrelu(conv3x3 + b) + conv1x1(x) -> fp16
#define B2SR_TCG_VARIANT_R16 2     // x4: lrelu(conv + b) + x2 (fp16 residual) -> fp16
#define B2SR_TCG_VARIANT_R1_O3 3   // dense-block output half: 0.2 v + x (fp32 residual) -> fp16 + fp32
#define B2SR_TCG_VARIANT_R2_O3 4   // RRDB output half: two fp32 residual terms -> fp16 + fp32
#define B2SR_TCG_VARIANT_GENERIC 5 // anything else with 32 output channels (run-time residual / output forms)

struct TcgPipeParams {
    const TcgParams* stages;  // device array [n_stages]
    int32_t n_stages, nb;
};

__global__ void __launch_bounds__(TC_THREADS, 1) tcg_pipe_kernel(const __grid_constant__ TcgPipeParams Q) {
    extern __shared__ uint8_t smem_raw[];
    // The stage's parameters live in SHARED memory: the per-launch kernel reads them from the constant bank (kernel
    // parameter), a local-memory copy here turned every P.field of the epilogue into a local load that misses the small L1
    // the streaming loads and stores keep flushing -- measured 10 000 - 14 000 cycles per row and warp in the residual
    // epilogues, 6 x what they take with the parameters in shared memory.
    __shared__ TcgParams sP;
    const int stage = (int)blockIdx.x / Q.nb, band = (int)blockIdx.x % Q.nb;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(Q.stages + stage);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&sP);
        for (int i = threadIdx.x; i < (int)(sizeof(TcgParams) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const TcgParams& P = sP;
    const int it_begin = P.item_first[band], it_end = P.item_first[band + 1];
    switch (P.variant) {
        case B2SR_TCG_VARIANT_PLAIN: tcg_body<32, 0, false, 0, 1, false, false, true>(P, it_begin, it_end, band, (uint32_t)P.half, smem_raw); break;
        case B2SR_TCG_VARIANT_SC: tcg_body<32, 0, false, 0, 1, false, true, true>(P, it_begin, it_end, band, (uint32_t)P.half, smem_raw); break;
        case B2SR_TCG_VARIANT_R16: tcg_body<32, 0, false, 1, 1, true, false, true>(P, it_begin, it_end, band, (uint32_t)P.half, smem_raw); break;
        case B2SR_TCG_VARIANT_R1_O3: tcg_body<32, 0, false, 1, 3, false, false, true>(P, it_begin, it_end, band, (uint32_t)P.half, smem_raw); break;
        case B2SR_TCG_VARIANT_R2_O3: tcg_body<32, 0, false, 2, 3, false, false, true>(P, it_begin, it_end, band, (uint32_t)P.half, smem_raw); break;
        default: tcg_body<32, 0, false, -1, 0, false, false, true>(P, it_begin, it_end, band, (uint32_t)P.half, smem_raw); break;
    }
}

// nearest-neighbour x r of a channel-last fp16 buffer (ncnn Interp resize_type 1 with integer scales: source index
// = floor(dst / r)); reference models/4x_Valar_v1.param:1203,1205.  One thread = 8 channels of one output pixel.
__global__ void tcg_nearest_kernel(const __half* __restrict__ in, int ldin, const PlaneDev* __restrict__ planes, int res_in, int r, int C,
                                   __half* __restrict__ out, int ldout) {
    const PlaneDev P = planes[blockIdx.y];
    const int Wi = P.Wt * res_in, Wo = Wi * r, Ho = P.Ht * res_in * r, C8 = C / 8;
    const long long in_base = (long long)P.pix_off * res_in * res_in, out_base = in_base * r * r;
    const long long total = (long long)Ho * Wo * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % C8);
        const long long p = i / C8;
        const int x = (int)(p % Wo), y = (int)(p / Wo);
        const uint4 v = *reinterpret_cast<const uint4*>(in + (in_base + (long long)(y / r) * Wi + x / r) * ldin + ch * 8);
        *reinterpret_cast<uint4*>(out + (out_base + p) * ldout + ch * 8) = v;
    }
}

}  // namespace b2sr
