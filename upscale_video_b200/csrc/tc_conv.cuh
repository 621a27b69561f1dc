// tc_conv.cuh -- 3x3 convolution as an im2col-free tcgen05 contraction (sm_100a).
//
// Replaces the Convolution(+PReLU) / PixelShuffle / Interp / BinaryOp layers that ncnn_vulkan executes for the
// reference inside `ex.extract` (reference upscale/upscale_processing.py:278-281, :450-453; graph
// models/2x_Compact_Pretrain.param:5-42).
//
// Layout.  Activations are channel-last fp16, CPIX channels per pixel (CPIX*2 = 32/64/128 bytes = exactly one
// swizzle row of the 32B/64B/128B shared-memory swizzle).  A work item is a band of `w` <= 128 columns x `rows`
// rows of a plane.  The producer warp streams the band's input rows (columns x0-1 .. x0+TC_PITCH-2, rows
// y0-1 .. y0+rows; TMA zero-fills outside the plane = the convolution's zero padding at the reference's tile
// border) into a ring of R shared-memory rows of TC_PITCH pixels.  One M=128 accumulator tile = one band row.
// Because a pixel is one swizzle row, the A operand of column tap kx for an input row is the 128 consecutive
// pixels of that ring row starting at pixel kx -- shifted UMMA descriptors over the same bytes, no im2col, no data
// replication.
//
// Row-stationary, tap-stacked contraction: an input row r feeds output rows r+1, r, r-1 through taps ky = 0, 1, 2.
// The weights of the three ky taps of one kx are stacked along N ([W(2,kx); W(1,kx); W(0,kx)], N = 3*NOUT), so ONE
// tcgen05.mma per (kx, 16-channel slab) updates three output rows at once: accumulator blocks of NOUT columns sit
// in a ring of TC_NBLK blocks in TMEM, output row t in block t mod TC_NBLK, and the N = 3*NOUT window slides by one
// block per input row.  Each input row is read from shared memory by 3*CPIX/16 MMAs (12 for 64 channels) instead of
// 36 -- the A-operand traffic, which bounded the N = 64 formulation at the 128 B/clk shared-memory port, drops 3x.
// Every MMA accumulates (the issue queue is shallow, so the issuer's per-row work is kept minimal): the epilogue
// warp that drains a block writes zeros back (tcgen05.st) before handing it to the issuer again.  A window that
// wraps around the ring is issued as two MMAs (N = 2*NOUT + NOUT).
// The stacked weights stay resident in shared memory for the life of the CTA (pre-swizzled on the host).
//
// Roles (384 threads): warp 0 = TMA producer; warp 1 = TMEM owner + tcgen05.mma issuer (the whole warp runs the
// control flow so descriptors live in uniform registers, one elected lane issues); warp 10 = publisher and
// warp 11 = counter poller (pipelined mode); warps 2..5 and 6..9 = two epilogue sets that take alternate rows (tcgen05.ld -> bias/PReLU -> fp16 -> swizzled staging -> 128-bit coalesced
// stores, or pixel-shuffle + nearest-upsampled residual + x255 + round-half-even/saturate for the last layer).
//
// Two schedules share this body:
//  * layer mode (tc_conv_kernel): one launch per layer over full per-plane activation buffers in HBM; the linear
//    sequence (plane, band, row) is cut into one contiguous equal range per CTA.
//  * pipelined mode (tc_pipe_kernel): ONE persistent launch for the whole network.  CTA (l, b) keeps layer l's
//    weights resident and streams band b of every plane, top to bottom; layer l writes its output rows into a ring
//    of RR rows that stays L2-resident, layer l+1's CTAs pull them as soon as they are published.  Flow control is
//    per band and per row through counters in global memory: `done` (rows written and fenced; the consumer's TMA
//    producer acquires the counters of bands b-1, b, b+1 before loading a row) and `cons` (rows pulled into shared
//    memory; the producer's epilogue waits for row g-RR to be consumed before overwriting its ring slot).  No
//    activation ever goes to HBM, no layer is ever relaunched, weights are loaded once per CTA.
#pragma once
#include <stdio.h>

#include "common.cuh"

namespace b2sr {

constexpr int TC_NSETS = 2;                       // epilogue warp sets (output rows alternate between them)
constexpr int TC_NBLK = 8;                        // accumulator blocks (output rows in flight) in the TMEM ring
constexpr int TC_THREADS = 128 + 128 * TC_NSETS;   // producer, MMA, publisher and poller warps + 4 epilogue warps per set
constexpr int TC_NROWBAR = 16;                    // 8-byte slots reserved for the epilogue warps' progress words (pipelined mode)
constexpr int TC_MAX_SLOTS = 64;
constexpr int TC_TILE_M = 128;

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ uint32_t mbar_test_wait(uint32_t bar, uint32_t parity) {  // never suspends
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
// Bounded wait: a protocol bug must trap (and fail the launch) rather than hang the GPU box.
__device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity, int who) {
    printf("b2sr: mbarrier timeout block %d thread %d bar 0x%x parity %u who %d\n", (int)blockIdx.x, (int)threadIdx.x,
           bar, parity, who);
    __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int who) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 6000000000LL) mbar_timeout(bar, parity, who);
    }
}
__device__ __forceinline__ void mbar_wait_timed(uint32_t bar, uint32_t parity, int who, long long& waited) {
    if (mbar_test_wait(bar, parity)) return;  // (test_wait never suspends: a true result costs nothing and counts nothing)
    const long long t0 = clock64();
    mbar_wait(bar, parity, who);
    waited += clock64() - t0;
}
// try_wait may itself suspend the thread for a while before it returns true, which mbar_wait_timed does not see:
// this variant brackets the whole wait (stall accounting of the fused graph kernel)
__device__ __forceinline__ void mbar_wait_clocked(uint32_t bar, uint32_t parity, int who, long long& waited) {
    const long long t0 = clock64();
    mbar_wait(bar, parity, who);
    waited += clock64() - t0;
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_gpu(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __noinline__ void flag_timeout(const uint32_t* p, uint32_t need, int who) {
    printf("b2sr: flag timeout block %d thread %d need %u have %u who %d\n", (int)blockIdx.x, (int)threadIdx.x, need,
           ld_acquire_gpu(p), who);
    __trap();
}
__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Inter-CTA hand-off of a ring row (pipelined mode), as a chain the PTX memory model covers end to end:
//   writer CTA:  epilogue warps st.global (generic proxy) -> fence.acq_rel.cta + progress word -> publisher: st.release.gpu done[]
//   reader CTA:  poller warp ld.relaxed.gpu done[] x3 (independent) + ONE fence.acq_rel.gpu -> fence.acq_rel.cta + st.shared s_avail
//                -> producer lane ld.shared s_avail
//                -> fence.acq_rel.cta -> fence.proxy.async (generic -> async proxy, once per advance of `seen`, not per row)
//                -> cp.async.bulk.tensor (TMA, async proxy) of the row.
// The acquire sits in the poller warp, off every critical path.  (A fence.acq_rel.gpu per row in the TMA-issuing thread
// was measured to serialise the row pipeline -- each fence waits for the loads in flight: 85 fps instead of ~400.)
// The reverse edge (ring slot reuse): the issuer stores cons[] only after the row's mbarrier phase completed, i.e. after
// the TMA read of the slot has finished; the writer's poller acquires cons[] before its epilogue overwrites the slot.
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ uint32_t elect_one_sync() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred;
}
__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {  // 32 lanes x 16 columns of zeros
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// compile-time geometry shared by host and device
// ------------------------------------------------------------------------------------------------
constexpr int TC_BW = 128;     // output columns per band = UMMA M
constexpr int TC_PITCH = 136;  // pixels per ring row: 130 needed (+-1 halo); 136 keeps every row swizzle-atom aligned

template <int CPIX, int NOUT, int SHUF>
struct TcCfg {
    static constexpr int PB = CPIX * 2;            // bytes per input pixel == swizzle row
    static constexpr int KSLABS = CPIX / 16;       // K=16 MMAs per filter tap
    static constexpr int WB = 9 * NOUT * PB;       // weight image bytes
    static constexpr int ROWB = TC_PITCH * PB;     // bytes per ring row
    static constexpr int OB = NOUT * 2;            // bytes per output pixel (PReLU epilogue)
    static constexpr int CH = OB / 16;             // 16-byte chunks per output pixel
    static constexpr int STG = SHUF == 0 ? TC_NSETS * 4 * 32 * OB : 0;
    static constexpr int TCOLS = TC_NBLK * NOUT <= 128 ? 128 : (TC_NBLK * NOUT <= 256 ? 256 : 512);
    static constexpr uint32_t LAYOUT = CPIX == 64 ? 2u : (CPIX == 32 ? 4u : 6u);  // UMMA LayoutType: SW128 / SW64 / SW32
    // instruction descriptor without N: D = f32, A = B = f16, both K-major, M = 128; N is added per MMA
    static constexpr uint32_t IDESC0 = (1u << 4) | ((uint32_t)(TC_TILE_M >> 4) << 24);
    static constexpr int MISC = 2 * NOUT * 4 + (2 * TC_MAX_SLOTS + 2 * TC_NBLK + TC_NROWBAR + 4) * 8 + 64;
    static_assert(TC_NBLK * NOUT <= 512, "accumulator ring exceeds TMEM");
    static_assert(CPIX == 16 || CPIX == 32 || CPIX == 64, "one pixel must be one swizzle row");
    static_assert(NOUT % 16 == 0 && NOUT >= 16 && NOUT <= 64, "UMMA M=128 needs N % 16 == 0");
    static_assert((NOUT * PB) % 1024 == 0, "per-tap weight tile must keep 1024-byte (swizzle atom) alignment");
    static_assert(ROWB % (8 * PB) == 0, "ring rows must start on a swizzle atom");
    // ring rows that fit beside the weights
    static constexpr int RING_FIT = (B2SR_SMEM_LIMIT - 1024 - WB - STG - MISC) / ROWB;
    static constexpr int RING = RING_FIT > 16 ? 16 : RING_FIT;
    static constexpr int SMEM = 1024 + WB + RING * ROWB + STG + MISC;
    static constexpr int ring_rows() { return RING; }
    static constexpr int smem_bytes() { return SMEM; }
};

// ------------------------------------------------------------------------------------------------
// the CTA body (shared by both schedules)
// ------------------------------------------------------------------------------------------------
template <int CPIX, int NOUT, int SHUF /*0 = PReLU->fp16, else pixel-shuffle factor*/, bool F32OUT, bool PIPE>
__device__ __forceinline__ void tc_conv_body(const TcParams& P, const int it_begin, const int it_end, const int band,
                                             uint8_t* smem_raw) {
    using C = TcCfg<CPIX, NOUT, SHUF>;
    constexpr int PB = C::PB;
    constexpr uint32_t ROWB = C::ROWB;
    constexpr int R = C::RING;
    static_assert(R >= 3, "shared-memory ring too small");
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (sbase - raw);

    const uint32_t w_s = sbase;
    const uint32_t ring_s = sbase + C::WB;
    const uint32_t stg_off = C::WB + (uint32_t)R * ROWB;
    const uint32_t fl_off = stg_off + C::STG;
    float* s_bias = reinterpret_cast<float*>(gbase + fl_off);
    float* s_slope = s_bias + NOUT;
    const uint32_t bar_off = fl_off + 2 * NOUT * 4;
    const uint32_t bar_s = sbase + bar_off;  // 8-byte aligned: all terms are multiples of 8
    auto full_bar = [&](int s) { return bar_s + 8u * s; };
    auto empty_bar = [&](int s) { return bar_s + 8u * (TC_MAX_SLOTS + s); };
    auto tfull_bar = [&](int b) { return bar_s + 8u * (2 * TC_MAX_SLOTS + b); };
    auto tempty_bar = [&](int b) { return bar_s + 8u * (2 * TC_MAX_SLOTS + TC_NBLK + b); };
    // progress words of the 4*TC_NSETS epilogue warps (pipelined mode): prog[w] = 2 + CTA-local index of the last row
    // warp w has stored and fenced (w's set owns every TC_NSETS-th row); starts at the set index
    volatile uint32_t* s_prog = reinterpret_cast<volatile uint32_t*>(gbase + bar_off + 8 * (2 * TC_MAX_SLOTS + 2 * TC_NBLK));
    static_assert((4 * TC_NSETS + 3) * 4 <= TC_NROWBAR * 8, "progress words do not fit their slots");
    // words maintained by the poller warp so that no critical thread ever waits on an L2 round trip:
    volatile uint32_t* s_avail = s_prog + 4 * TC_NSETS;        // min over bands b-1..b+1 of the previous layer's `done`
    volatile uint32_t* s_consmin = s_prog + 4 * TC_NSETS + 1;  // min over bands b-1..b+1 of the next layer's `cons`
    uint32_t* s_finished = const_cast<uint32_t*>(s_prog) + 4 * TC_NSETS + 2;  // roles that no longer need the poller
    const uint32_t w_bar = bar_s + 8u * (2 * TC_MAX_SLOTS + 2 * TC_NBLK + TC_NROWBAR);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(gbase + bar_off + 8 * (2 * TC_MAX_SLOTS + 2 * TC_NBLK + TC_NROWBAR + 1));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool ring_in = PIPE && P.ring_in, ring_out = PIPE && P.ring_out;
    const uint32_t RR = PIPE ? (uint32_t)P.RR : 1u;
    // neighbours whose rows overlap this band's 130-pixel input window / whose input windows overlap this band
    const int nb_lo = band > 0 ? band - 1 : 0, nb_hi = band + 1 < P.nb ? band + 1 : P.nb - 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < R; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < TC_NBLK; ++b) {
            mbar_init(tfull_bar(b), 1);
            mbar_init(tempty_bar(b), 4);
        }
        for (int w = 0; w < 4 * TC_NSETS; ++w) s_prog[w] = (uint32_t)(w >> 2);
        *s_avail = 0u;
        *s_consmin = 0u;
        *s_finished = 0u;
        mbar_init(w_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                     "r"((uint32_t)C::TCOLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp >= 2) {
        for (int i = threadIdx.x - 64; i < NOUT; i += TC_THREADS - 64) {
            s_bias[i] = P.bias[i];
            s_slope[i] = (SHUF == 0) ? P.slope[i] : 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if constexpr (PIPE && CPIX == 16) {
      if (warp == 0 && P.direct_in) {
        // ======================= frame-row producer (first layer of the pipelined schedule) =======================
        // The network input is read straight from the packed u8 frames: the reference's Mat.from_pixels(PIXEL_BGR) + tile
        // slicing (upscale_processing.py:430-442) happen on the way into shared memory, with no fp16 copy of the frames in
        // HBM and no separate pass over them.  The band's 136-pixel window of one frame row (408 bytes) is copied as aligned
        // 4-byte words with cp.async into a small staging ring, FR_DEPTH rows ahead (hides the HBM latency without holding
        // registers); the warp then expands the row to the 16-channel fp16 swizzle-32B ring row the MMA descriptors expect:
        // channels 0-2 = the pixel's bytes (0..255 is exact in fp16; * 1/255 is folded into this layer's epilogue), 3-15 zero,
        // pixels outside the plane (= the reference tile) zero = the convolution's zero padding.
        constexpr int FR_DEPTH = 6, FR_SLOTS = 8, FR_BYTES = 512;  // staging: 8 rows of 128 words
        static_assert(PB == 32, "one 16-channel fp16 pixel is 32 bytes");
        static_assert(TC_PITCH * 3 + 8 <= FR_BYTES, "staging row too small");
        const uint32_t fr_off = (uint32_t)((C::WB + R * ROWB + C::STG + C::MISC + 15) & ~15);
        const uint32_t fr_s = sbase + fr_off;
        if (lane == 0) {
            mbar_expect_tx(w_bar, C::WB);
            for (int t = 0; t < 9; ++t) bulk_g2s(w_s + t * (NOUT * PB), P.wimg + (size_t)t * (NOUT * PB), NOUT * PB, w_bar);
        }
        for (int i = lane; i < R * TC_PITCH; i += 32) {  // the all-zero half (channels 8..15) of every pixel of every slot, once
            const uint32_t a = ring_s + (uint32_t)(i / TC_PITCH) * ROWB + (uint32_t)(i % TC_PITCH) * 32u + 16u;
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a ^ (((a >> 7) & 1u) << 4)), "r"(0u) : "memory");
        }
        // (item, input row) iterator of the rows being REQUESTED, FR_DEPTH ahead of the rows being expanded.  Everything that
        // depends on the item only is kept in registers (the item table is in global memory, and this warp's L1 is busy).
        int it_f = it_begin - 1, rho_f = 0, n_req = 0, f_rows_in = 0, f_y_first = 0, f_Ht = 0;
        uintptr_t f_w0 = 0, f_b0 = 0, f_b1 = 0;  // row 0 of the plane: window start, first / one-past-last needed byte
        const uintptr_t f_lo = reinterpret_cast<uintptr_t>(P.frames_in), f_hi = reinterpret_cast<uintptr_t>(P.frames_end);
        const size_t f_pitch = (size_t)P.frame_w * 3;
        auto next_item = [&]() {
            for (++it_f; it_f < it_end; ++it_f) {
                const TcItem I = P.items[it_f];
                if (I.w <= 0) continue;
                const uint8_t* rowp = P.frames_in + ((size_t)((size_t)I.frame * P.frame_h + I.fy0) * P.frame_w + I.fx0) * 3;
                f_w0 = reinterpret_cast<uintptr_t>(rowp) + (uintptr_t)((ptrdiff_t)(I.x0 - 1) * 3);
                f_b0 = reinterpret_cast<uintptr_t>(rowp + (size_t)max(I.x0 - 1, 0) * 3);
                f_b1 = reinterpret_cast<uintptr_t>(rowp + (size_t)min(I.x0 - 1 + TC_PITCH, I.Wt) * 3);
                f_rows_in = I.rows + 2, f_y_first = I.y0 - 1, f_Ht = I.Ht;
                break;
            }
            rho_f = 0;
        };
        next_item();
        auto request = [&]() {  // cp.async the next row's words into staging slot n_req % FR_SLOTS; always commits one group
            if (it_f < it_end) {
                const int y = f_y_first + rho_f;
                if (y >= 0 && y < f_Ht) {
                    const uintptr_t off = (uintptr_t)y * f_pitch;
                    const uintptr_t base = (f_w0 + off) & ~(uintptr_t)3;  // aligned word holding the window's first byte
                    const uintptr_t b0 = f_b0 + off, b1 = f_b1 + off;
                    const uint32_t dst = fr_s + (uint32_t)(n_req % FR_SLOTS) * FR_BYTES;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int j = lane + 32 * k;  // word j of the staging row
                        const uintptr_t wa = base + (uintptr_t)j * 4;
                        if (wa + 4 > b0 && wa < b1) {  // the word holds at least one needed byte
                            if (wa >= f_lo && wa + 4 <= f_hi) {
                                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + (uint32_t)j * 4), "l"(wa) : "memory");
                            } else {  // the word straddles an end of the frame buffer: its in-range bytes one by one
                                for (int e = 0; e < 4; ++e)
                                    if (wa + e >= b0 && wa + e < b1)
                                        asm volatile("st.shared.u8 [%0], %1;" ::"r"(dst + (uint32_t)j * 4 + e), "r"((uint32_t)*reinterpret_cast<const uint8_t*>(wa + e)) : "memory");
                            }
                        }
                    }
                }
                if (++rho_f == f_rows_in) next_item();
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            ++n_req;
        };
        for (int i = 0; i < FR_DEPTH; ++i) request();
        int slot = 0, n_done = 0;
        uint32_t phase = 0;
        long long w_empty = 0, w_data = 0, t_expand = 0, t_fence = 0, t_req = 0;
        const long long t_begin = clock64();
        for (int it = it_begin; it < it_end; ++it) {
            const TcItem* Ip = P.items + it;
            if (Ip->w <= 0) continue;
            const int rows_in = Ip->rows + 2, x_first = Ip->x0 - 1, Wt = Ip->Wt, Ht = Ip->Ht, y_first = Ip->y0 - 1;
            const uint8_t* row0 = P.frames_in + ((size_t)((size_t)Ip->frame * P.frame_h + Ip->fy0) * P.frame_w + Ip->fx0) * 3 + (ptrdiff_t)x_first * 3;
            for (int rho = 0; rho < rows_in; ++rho, ++n_done) {
                const int y = y_first + rho;
                const long long tw0 = P.dbg ? clock64() : 0;
                asm volatile("cp.async.wait_group %0;" ::"n"(FR_DEPTH - 1) : "memory");  // this row's words have landed (own copies)
                const long long tw1 = P.dbg ? clock64() : 0;
                if (lane == 0) mbar_wait(empty_bar(slot), phase ^ 1u, 0);
                __syncwarp();  // ... and everybody else's
                const long long tw2 = P.dbg ? clock64() : 0;
                if (P.dbg) w_data += tw1 - tw0, w_empty += tw2 - tw1;
                // byte offset of the window inside its staging row = misalignment of the window's first byte
                const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(row0 + (size_t)y * P.frame_w * 3) & 3u);
                const uint32_t src = fr_s + (uint32_t)(n_done % FR_SLOTS) * FR_BYTES + mis;
                const bool row_in = y >= 0 && y < Ht;
#pragma unroll
                for (int k = 0; k < (TC_PITCH + 31) / 32; ++k) {
                    const int pp = lane + 32 * k;
                    if (pp < TC_PITCH) {
                        const int x = x_first + pp;
                        uint32_t c0 = 0u, c1 = 0u, c2 = 0u;
                        if (row_in && x >= 0 && x < Wt) {
                            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(c0) : "r"(src + (uint32_t)pp * 3));
                            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(c1) : "r"(src + (uint32_t)pp * 3 + 1));
                            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(c2) : "r"(src + (uint32_t)pp * 3 + 2));
                        }
                        const __half2 h01 = __floats2half2_rn((float)c0, (float)c1);
                        const __half2 h2 = __floats2half2_rn((float)c2, 0.f);
                        const uint32_t a = ring_s + (uint32_t)slot * ROWB + (uint32_t)pp * 32u;
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %3};" ::"r"(a ^ (((a >> 7) & 1u) << 4)),
                                     "r"(*reinterpret_cast<const uint32_t*>(&h01)), "r"(*reinterpret_cast<const uint32_t*>(&h2)), "r"(0u)
                                     : "memory");
                    }
                }
                const long long tw3 = P.dbg ? clock64() : 0;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> the tensor core's (async proxy) reads
                __syncwarp();  // every lane's stores are fenced, and everybody is done reading this staging slot
                if (lane == 0) mbar_arrive(full_bar(slot));
                const long long tw4 = P.dbg ? clock64() : 0;
                request();
                if (P.dbg) t_expand += tw3 - tw2, t_fence += tw4 - tw3, t_req += clock64() - tw4;
                if (++slot == R) {
                    slot = 0;
                    phase ^= 1u;
                }
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (P.dbg && lane == 0) {
            P.dbg[0] = clock64() - t_begin;
            P.dbg[1] = w_data;  // (printed in the "starved" column: waiting for the frame bytes)
            P.dbg[6] = w_empty;
            (void)t_expand, (void)t_fence, (void)t_req;  // measured on B200: expand 550, fence + arrive 68, request 680 cycles per row
        }
      }
    }
    if (warp == 0) {
        // ======================= TMA producer =======================
        if (lane == 0 && !(PIPE && CPIX == 16 && P.direct_in)) {
            mbar_expect_tx(w_bar, C::WB);
            for (int t = 0; t < 9; ++t)  // 9 chunks of the [kx][3*NOUT rows][CPIX] image
                bulk_g2s(w_s + t * (NOUT * PB), P.wimg + (size_t)t * (NOUT * PB), NOUT * PB, w_bar);
            int slot = 0;
            uint32_t phase = 0;
            uint32_t seen = 0u;  // last value read from s_avail
            long long waited = 0, w_empty = 0;
            const long long t_begin = clock64();
            for (int it = it_begin; it < it_end; ++it) {
                const TcItem I = P.items[it];
                if (PIPE && I.w <= 0) continue;
                const CUtensorMap* map = P.maps + (P.map_base + I.map);
                const int rows_in = I.rows + 2;
                for (int rho = 0; rho < rows_in; ++rho) {
                    const int y = I.y0 - 1 + rho;  // plane row of this input row
                    int cy = y, cp = I.plane;
                    if (ring_in) {
                        cp = 0;
                        if (y >= 0 && y < I.Ht) {
                            const uint32_t g = (uint32_t)(I.grow0 + y);
                            cy = (int)(g % RR);
                            if (seen < g + 1u) {  // rows 0..g of bands b-1, b, b+1 of the previous layer written and released?
                                const long long t0 = clock64();
                                while ((seen = *s_avail) < g + 1u) {
                                    __nanosleep(20);
                                    if (clock64() - t0 > 20000000000LL) flag_timeout(P.done_in + band * B2SR_FLAG_STRIDE, g + 1u, 10);
                                }
                                __threadfence_block();       // pairs with the poller's fence before it stored s_avail
                                fence_proxy_async_global();  // rows up to `seen` were written through the generic proxy; TMA reads them
                                waited += clock64() - t0;
                            }
                        } else {
                            cy = -1;  // outside the plane: any out-of-bounds coordinate reads as zeros
                        }
                    }
                    mbar_wait_timed(empty_bar(slot), phase ^ 1u, 0, w_empty);
                    mbar_expect_tx(full_bar(slot), ROWB);
                    tma_load_4d(ring_s + slot * ROWB, map, full_bar(slot), 0, I.x0 - 1, cy, cp);
                    if (++slot == R) {
                        slot = 0;
                        phase ^= 1u;
                    }
                }
            }
            if (ring_in) atomicAdd(s_finished, 1u);
            if (P.dbg) {
                P.dbg[0] = clock64() - t_begin;
                P.dbg[1] = waited;
                P.dbg[6] = w_empty;
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
        // Every lane runs the (warp-uniform) control flow; one elected lane issues the tcgen05 instructions.
        // The issue queue is shallow, so everything between two rows' MMAs is kept short: barrier states for the
        // NEXT row are probed (non-blocking) before this row's MMAs are issued and their latency hides behind them.
        static_assert((TC_NBLK & (TC_NBLK - 1)) == 0, "block ring size must be a power of two");
        const uint32_t desc_hi = ((8u * PB) >> 4) | (1u << 14) | (C::LAYOUT << 29);  // SBO | version | swizzle mode
        const uint64_t hi64 = (uint64_t)desc_hi << 32;
        const uint32_t w_lo = (w_s >> 4) | (1u << 16);  // start address | LBO(unused)=1
        const uint32_t ring_lo = (ring_s >> 4) | (1u << 16);
        constexpr uint32_t KXB = (3 * NOUT * PB) >> 4;  // descriptor units between the stacked weight tiles of kx, kx+1
        constexpr uint32_t BLKB = (NOUT * PB) >> 4;     // ... between the ky blocks inside one tile
        constexpr int NM = 3 * C::KSLABS;               // MMAs per input row
        int slot = 0;  // ring slot of the current input row
        uint32_t phase = 0;
        uint32_t g0 = 0;      // CTA-local index of the current item's output row 0 (block ring / barrier phases)
        uint32_t gfresh = 0;  // CTA-local index of the next output row to be started (== g0 + rho while rho < rows)
        long long w_full = 0, w_tempty = 0;
        mbar_wait(w_bar, 0, 1);
        uint32_t ok_full = mbar_test_wait(full_bar(0), 0);
        uint32_t ok_tempty = mbar_test_wait(tempty_bar(0), 0);
        for (int it = it_begin; it < it_end; ++it) {
            const TcItem* Ip = P.items + it;
            const int rows = Ip->rows;
            if (PIPE && Ip->w <= 0) {  // band absent from this plane: nothing to pull, just move the counter on
                if (ring_in && lane == 0) st_relaxed_gpu(P.cons_self + band * B2SR_FLAG_STRIDE, (uint32_t)(Ip->grow0 + Ip->Ht));
                continue;
            }
            const int y_first = Ip->y0 - 1, plane_h = Ip->Ht;
            const uint32_t grow0 = (uint32_t)Ip->grow0;
            for (int rho = 0; rho < rows + 2; ++rho) {  // input row rho feeds output rows rho - ky, ky = 0..2
                const bool fresh = rho < rows;  // output row `rho` receives its first contribution (ky = 0)
                if (!ok_full) mbar_wait_timed(full_bar(slot), phase, 2, w_full);
                // block of the new output row: drained and zeroed by its epilogue set? (use u of a block completes phase u)
                if (fresh && !ok_tempty) mbar_wait_timed(tempty_bar(gfresh & (TC_NBLK - 1)), (gfresh / TC_NBLK) & 1u, 3, w_tempty);
                tc_fence_after();
                if (ring_in) {  // the row is in shared memory now: its ring slot in L2 may be overwritten
                    const int y = y_first + rho;
                    if (y >= 0 && y < plane_h && lane == 0) st_relaxed_gpu(P.cons_self + band * B2SR_FLAG_STRIDE, grow0 + (uint32_t)y + 1u);
                }
                // probes for the next input row
                const int nslot = slot + 1 == R ? 0 : slot + 1;
                const uint32_t nphase = slot + 1 == R ? phase ^ 1u : phase;
                const uint32_t gn = gfresh + (fresh ? 1u : 0u);
                ok_full = mbar_test_wait(full_bar(nslot), nphase);
                ok_tempty = mbar_test_wait(tempty_bar(gn & (TC_NBLK - 1)), (gn / TC_NBLK) & 1u);

                const int t0 = rho >= 2 ? rho - 2 : 0;  // output rows touched: t0 .. t1 (clipped to the item)
                const int t1 = fresh ? rho : rows - 1;
                const int cnt = t1 - t0 + 1;
                const uint32_t blk0 = (g0 + (uint32_t)t0) & (TC_NBLK - 1);  // accumulator block of output row t0
                const int wrap = (int)(TC_NBLK - blk0);
                const int n1 = cnt < wrap ? cnt : wrap;                      // blocks before the ring wraps
                const uint32_t brow0 = rho >= 2 ? 0u : (uint32_t)(2 - rho);  // t0's tap block in [W(2) | W(1) | W(0)]
                if (elect_one_sync()) {
                    const uint32_t a_lo = ring_lo + (uint32_t)slot * (ROWB >> 4);
                    const uint32_t b_lo = w_lo + brow0 * BLKB;
                    const uint32_t id1 = C::IDESC0 | ((uint32_t)((n1 * NOUT) >> 3) << 17);
                    const uint32_t d1 = tmem_base + blk0 * NOUT;
                    // base_offset stays 0 although kx shifts the A start inside a swizzle atom: the hardware swizzles
                    // on absolute shared-memory address bits, like TMA did when it wrote the row
                    if (n1 == cnt) {
#pragma unroll
                        for (int m = 0; m < NM; ++m) {
                            const int kx = m / C::KSLABS, k = m % C::KSLABS;
                            const uint32_t ao = (uint32_t)((kx * PB + k * 32) >> 4), bo = (uint32_t)(kx * KXB + ((k * 32) >> 4));
                            umma_f16(d1, hi64 | (a_lo + ao), hi64 | (b_lo + bo), id1, 1u);
                        }
                    } else {  // the window wraps around the block ring: two MMAs per (kx, slab)
                        const uint32_t id2 = C::IDESC0 | ((uint32_t)(((cnt - n1) * NOUT) >> 3) << 17);
                        const uint32_t b_lo2 = b_lo + (uint32_t)n1 * BLKB;
#pragma unroll
                        for (int m = 0; m < NM; ++m) {
                            const int kx = m / C::KSLABS, k = m % C::KSLABS;
                            const uint32_t ao = (uint32_t)((kx * PB + k * 32) >> 4), bo = (uint32_t)(kx * KXB + ((k * 32) >> 4));
                            umma_f16(d1, hi64 | (a_lo + ao), hi64 | (b_lo + bo), id1, 1u);
                            umma_f16(tmem_base, hi64 | (a_lo + ao), hi64 | (b_lo2 + bo), id2, 1u);
                        }
                    }
                    umma_commit(empty_bar(slot));                                                     // input row consumed
                    if (rho >= 2) umma_commit(tfull_bar((g0 + (uint32_t)rho - 2u) & (TC_NBLK - 1)));  // output row rho-2 done
                }
                __syncwarp();
                slot = nslot;
                phase = nphase;
                gfresh = gn;
            }
            g0 += (uint32_t)rows;
        }
        if (P.dbg && lane == 0) {
            P.dbg[3] = w_full;
            P.dbg[4] = w_tempty;
        }
    } else if (warp == 2 + 4 * TC_NSETS) {
        // ======================= publisher (pipelined mode) =======================
        // Turns the epilogue warps' progress words into one monotonic "rows of this band written and fenced" counter
        // in global memory.  It polls, so a slow release store only batches several rows into one update.
        if (ring_out && lane == 0) {
            static_assert(TC_NSETS == 2, "the progress-word arithmetic below assumes two epilogue sets");
            uint32_t n_local = 0;  // output rows this CTA produces
            for (int it = it_begin; it < it_end; ++it)
                if (P.items[it].w > 0) n_local += (uint32_t)P.items[it].rows;
            const uint32_t g_end = it_end > it_begin ? (uint32_t)(P.items[it_end - 1].grow0 + P.items[it_end - 1].Ht) : 0u;
            uint32_t cnt = 0;  // CTA-local rows accounted for so far
            int it = it_begin;
            uint32_t t = 0;    // rows of item `it` accounted for
            uint32_t published = 0;
            long long n_pub = 0, t_pub = 0;
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
                    const long long tp = clock64();
                    __threadfence_block();  // acquire side of the progress words ...
                    st_release_gpu(P.done_out + band * B2SR_FLAG_STRIDE, g);  // ... then one GPU-scope release for all eight warps' stores
                    published = g;
                    n_pub += 1;
                    t_pub += clock64() - tp;
                }
                if (it >= it_end) break;
                __nanosleep(100);
                if (clock64() - t_start > 20000000000LL) flag_timeout(P.done_out + band * B2SR_FLAG_STRIDE, g_end, 30);
            }
            if (P.dbg) P.dbg[7] = n_pub > 0 ? (t_pub / n_pub) * 1000000LL + n_pub : 0;  // mean cycles per publish * 1e6 + count
        }
    } else if (warp == 3 + 4 * TC_NSETS) {
        // ======================= counter poller (pipelined mode) =======================
        // Keeps shared-memory copies of the neighbours' global counters fresh, so that the TMA producer and the
        // epilogue warps test a shared-memory word instead of paying L2 round trips on their critical paths.
        if constexpr (PIPE) if ((ring_in || ring_out) && lane == 0) {
            const uint32_t need_fin = (ring_in ? 1u : 0u) + (ring_out ? 4u * TC_NSETS : 0u);
            const uint32_t* d0 = ring_in ? P.done_in + nb_lo * B2SR_FLAG_STRIDE : nullptr;
            const uint32_t* d1 = ring_in ? P.done_in + band * B2SR_FLAG_STRIDE : nullptr;
            const uint32_t* d2 = ring_in ? P.done_in + nb_hi * B2SR_FLAG_STRIDE : nullptr;
            const uint32_t* c0 = ring_out ? P.cons_next + nb_lo * B2SR_FLAG_STRIDE : nullptr;
            const uint32_t* c1 = ring_out ? P.cons_next + band * B2SR_FLAG_STRIDE : nullptr;
            const uint32_t* c2 = ring_out ? P.cons_next + nb_hi * B2SR_FLAG_STRIDE : nullptr;
            const long long t_start = clock64();
            while (*reinterpret_cast<volatile uint32_t*>(s_finished) < need_fin) {
                // independent relaxed loads (one L2 round trip for all six), then ONE acquire fence before the shared words
                uint32_t ma = 0u, mc = 0u;
                if (ring_in) {
                    const uint32_t a = ld_relaxed_gpu(d0), b = ld_relaxed_gpu(d1), c = ld_relaxed_gpu(d2);
                    ma = a < b ? (a < c ? a : c) : (b < c ? b : c);
                }
                if (ring_out) {
                    const uint32_t a = ld_relaxed_gpu(c0), b = ld_relaxed_gpu(c1), c = ld_relaxed_gpu(c2);
                    mc = a < b ? (a < c ? a : c) : (b < c ? b : c);
                }
                fence_acq_rel_gpu();
                __threadfence_block();
                if (ring_in) *s_avail = ma;
                if (ring_out) *s_consmin = mc;
                __nanosleep(200);
                if (clock64() - t_start > 40000000000LL) flag_timeout(ring_in ? d1 : c1, 0xffffffffu, 40);
            }
        }
    } else {
        // ======================= epilogue =======================
        const int q = warp & 3;                // TMEM lane quadrant this warp may read
        const uint32_t set = (warp - 2) >> 2;  // this warp's epilogue set (takes output rows g with g % TC_NSETS == set)
        uint32_t tile_cnt = 0;
        // hand every accumulator block of this set to the issuer zeroed (all MMAs accumulate)
        for (uint32_t b = set; b < TC_NBLK; b += TC_NSETS) {
#pragma unroll
            for (int j = 0; j < NOUT; j += 16) tmem_st16_zero(tmem_base + ((uint32_t)(q * 32) << 16) + b * NOUT + j);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(b));
        }
        const float scale_acc = P.acc_scale;
        const int c = q * 32 + lane;  // column inside the band
        uint32_t cons_seen = 0u;  // last value read from s_consmin
        long long waited = 0, w_tfull = 0;
        for (int it = it_begin; it < it_end; ++it) {
            const TcItem I = P.items[it];
            if (PIPE && I.w <= 0) continue;
            const bool valid = c < I.w;
            for (int t = 0; t < I.rows; ++t, ++tile_cnt) {
                if (tile_cnt % TC_NSETS != set) continue;
                const uint32_t buf = tile_cnt % TC_NBLK;
                mbar_wait_timed(tfull_bar(buf), (tile_cnt / TC_NBLK) & 1u, 4, w_tfull);
                tc_fence_after();
                uint32_t acc[NOUT];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * NOUT;
#pragma unroll
                for (int j = 0; j < NOUT; j += 16) tmem_ld16(taddr + j, acc + j);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < NOUT; j += 16) tmem_st16_zero(taddr + j);  // the block's next row accumulates from zero
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar(buf));

                if constexpr (SHUF == 0) {
                    // bias + PReLU -> fp16, via swizzled per-warp staging, then 128-bit coalesced global stores
                    constexpr int CH = C::CH;
                    uint4* stg = reinterpret_cast<uint4*>(gbase + stg_off + (warp - 2) * (32 * C::OB));
                    int off;
                    uint8_t* outp;
                    if (ring_out) {
                        const uint32_t g = (uint32_t)(I.grow0 + I.y0 + t);
                        if (g >= RR) {  // ring slot still holds row g - RR: have bands b-1, b, b+1 of the next layer pulled it?
                            if (cons_seen < g - RR + 1u) {
                                const long long t0 = clock64();
                                while ((cons_seen = *s_consmin) < g - RR + 1u) {
                                    __nanosleep(20);
                                    if (clock64() - t0 > 20000000000LL) flag_timeout(P.cons_next + band * B2SR_FLAG_STRIDE, g - RR + 1u, 20);
                                }
                                __threadfence_block();  // pairs with the poller's fence before it stored s_consmin
                                waited += clock64() - t0;
                            }
                            __syncwarp();
                        }
                        off = valid ? (int)((g % RR) * (uint32_t)P.Wmax) + I.x0 + c : -1;
                        outp = reinterpret_cast<uint8_t*>(P.out);
                    } else {
                        off = valid ? ((I.y0 + t) * I.Wt + I.x0 + c) : -1;
                        outp = reinterpret_cast<uint8_t*>(P.out) + (size_t)I.pix_off * C::OB;
                    }
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
                            __half2 h = __floats2half2_rn(v0, v1);
                            pk[e >> 1] = *reinterpret_cast<uint32_t*>(&h);
                        }
                        const int qi = lane * CH + (j >> 3);
                        stg[qi ^ ((qi >> 3) & 7)] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < CH; ++i) {
                        const int qi = i * 32 + lane;
                        const uint4 v = stg[qi ^ ((qi >> 3) & 7)];
                        const int o = __shfl_sync(0xffffffffu, off, qi / CH);
                        if (o >= 0) *reinterpret_cast<uint4*>(outp + (size_t)o * C::OB + (qi % CH) * 16) = v;
                    }
                    if (ring_out) {
                        // Report the row to the publisher with CTA-scope ordering only; the publisher's single GPU-scope
                        // release covers every warp's stores by cumulativity (the __syncthreads + one __threadfence idiom).
                        __syncwarp();
                        if (lane == 0) {
                            __threadfence_block();
                            s_prog[warp - 2] = tile_cnt + 2u;
                        }
                    }
                    __syncwarp();
                } else {
                    // last layer: pixel shuffle + nearest-upsampled input residual + x255 (+ round/saturate)
                    constexpr int S = SHUF;
                    const int fy = I.fy0 + I.y0 + t, fx = I.fx0 + I.x0 + c;
                    if (valid && fy >= I.cy0 && fy < I.cy1 && fx >= I.cx0 && fx < I.cx1) {
                        const uint8_t* px = P.frames_in + ((size_t)((size_t)I.frame * P.frame_h + fy) * P.frame_w + fx) * 3;
                        float xin[3];
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) xin[ch] = (float)px[ch] * (1.f / 255.f);
                        const size_t OW = (size_t)P.frame_w * S;
#pragma unroll
                        for (int dy = 0; dy < S; ++dy) {
                            float v[S * 3];
#pragma unroll
                            for (int dx = 0; dx < S; ++dx)
#pragma unroll
                                for (int ch = 0; ch < 3; ++ch) {
                                    const int n = ch * S * S + dy * S + dx;
                                    float t0 = fmaf(__uint_as_float(acc[n]), scale_acc, s_bias[n]);
                                    t0 = t0 + xin[ch];
                                    v[dx * 3 + ch] = t0 * 255.f;
                                }
                            const size_t o = (((size_t)I.frame * P.frame_h * S + (size_t)fy * S + dy) * OW + (size_t)fx * S) * 3;
                            if constexpr (F32OUT) {
                                float* dst = reinterpret_cast<float*>(P.out) + o;
#pragma unroll
                                for (int e = 0; e < S * 3; ++e) dst[e] = v[e];
                            } else {
                                uint8_t b[S * 3];
#pragma unroll
                                for (int e = 0; e < S * 3; ++e) {
                                    int iv = __float2int_rn(v[e]);  // round half to even, like cv2's saturate_cast
                                    b[e] = (uint8_t)min(max(iv, 0), 255);
                                }
                                uint8_t* dst = reinterpret_cast<uint8_t*>(P.out) + o;
                                if constexpr (S == 4) {
#pragma unroll
                                    for (int e = 0; e < 3; ++e)
                                        reinterpret_cast<uint32_t*>(dst)[e] = (uint32_t)b[4 * e] | ((uint32_t)b[4 * e + 1] << 8) |
                                                                              ((uint32_t)b[4 * e + 2] << 16) |
                                                                              ((uint32_t)b[4 * e + 3] << 24);
                                } else if constexpr (S == 2) {
#pragma unroll
                                    for (int e = 0; e < 3; ++e)
                                        reinterpret_cast<uint16_t*>(dst)[e] = (uint16_t)(b[2 * e] | (b[2 * e + 1] << 8));
                                } else {
#pragma unroll
                                    for (int e = 0; e < 3; ++e) dst[e] = b[e];
                                }
                            }
                        }
                    }
                }
            }
        }
        if (ring_out && lane == 0) atomicAdd(s_finished, 1u);
        if (P.dbg && warp == 2 && lane == 0) {
            P.dbg[2] = waited;
            P.dbg[5] = w_tfull;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TCOLS)
                     : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// layer mode: one launch per convolution
// ------------------------------------------------------------------------------------------------
template <int CPIX, int NOUT, int SHUF, bool F32OUT>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_conv_kernel(const __grid_constant__ TcParams P) {
    extern __shared__ uint8_t smem_raw[];
    if (P.dbg) {  // optional stall accounting (B2SR_OPT_PIPE_DEBUG): this CTA's 8 words
        TcParams Q = P;
        Q.dbg = P.dbg + (size_t)blockIdx.x * 8;
        tc_conv_body<CPIX, NOUT, SHUF, F32OUT, false>(Q, Q.item_first[blockIdx.x], Q.item_first[blockIdx.x + 1], 0, smem_raw);
        return;
    }
    tc_conv_body<CPIX, NOUT, SHUF, F32OUT, false>(P, P.item_first[blockIdx.x], P.item_first[blockIdx.x + 1], 0, smem_raw);
}

// ------------------------------------------------------------------------------------------------
// pipelined mode: one persistent launch for the whole network, CTA = (layer, band)
// ------------------------------------------------------------------------------------------------
template <int CF /*padded feature channels*/, int NL /*padded last-layer channels*/, int S /*scale*/, bool F32OUT>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_pipe_kernel(const __grid_constant__ PipeParams Q) {
    extern __shared__ uint8_t smem_raw[];
    const int layer = (int)blockIdx.x / Q.nb, band = (int)blockIdx.x % Q.nb;
    TcParams P = Q.layers[layer];
    P.dbg = Q.dbg ? Q.dbg + (size_t)(layer * Q.nb + band) * 8 : nullptr;
    const int it_begin = P.item_first[band], it_end = P.item_first[band + 1];
    if (layer == 0)
        tc_conv_body<16, CF, 0, false, true>(P, it_begin, it_end, band, smem_raw);
    else if (layer == Q.n_layers - 1)
        tc_conv_body<CF, NL, S, F32OUT, true>(P, it_begin, it_end, band, smem_raw);
    else
        tc_conv_body<CF, CF, 0, false, true>(P, it_begin, it_end, band, smem_raw);
}

template <int CF, int NL, int S>
struct TcPipeCfg {
    static constexpr int a_ = TcCfg<16, CF, 0>::smem_bytes(), b_ = TcCfg<CF, CF, 0>::smem_bytes(), c_ = TcCfg<CF, NL, S>::smem_bytes();
    static constexpr int a2_ = a_ + 8 * 512 + 16;  // + the first layer's frame-row staging ring (direct u8 input)
    static constexpr int smem_bytes() { return a2_ > b_ ? (a2_ > c_ ? a2_ : c_) : (b_ > c_ ? b_ : c_); }
};

}  // namespace b2sr
