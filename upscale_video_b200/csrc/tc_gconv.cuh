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
// ([g][kx][3 * NOUT rows][64 ch], <= 110.6 KB: a 192 -> 64 convolution is launched as two 192 -> 32 halves).
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
};

constexpr int TCG_PB = 128;                      // bytes per pixel of one channel group == one SW128 swizzle row
constexpr int TCG_SUBROWB = TC_PITCH * TCG_PB;   // one ring slot
constexpr int TCG_BAR_WORDS = 2 * TC_MAX_SLOTS + 2 * TC_NBLK + 2;

template <int NOUT, int MODE>
struct TcgCfg {
    static constexpr int OB = NOUT * 2;
    static constexpr int STG = MODE == 0 ? TC_NSETS * 4 * 32 * OB : 0;
    static constexpr int MISC = 2 * NOUT * 4 + TCG_BAR_WORDS * 8 + 64;
    static constexpr int TCOLS = TC_NBLK * NOUT <= 128 ? 128 : (TC_NBLK * NOUT <= 256 ? 256 : 512);
    static constexpr uint32_t IDESC0 = (1u << 4) | ((uint32_t)(TC_TILE_M >> 4) << 24);
    static_assert(NOUT == 16 || NOUT == 32 || NOUT == 64, "NOUT");
    static constexpr int weight_bytes(int groups) { return groups * 9 * NOUT * TCG_PB; }
    static constexpr int ring_fit(int groups) { return (B2SR_SMEM_LIMIT - 1024 - weight_bytes(groups) - STG - MISC) / TCG_SUBROWB; }
    static constexpr int smem_bytes(int groups, int slots) { return 1024 + weight_bytes(groups) + slots * TCG_SUBROWB + STG + MISC; }
};

template <int NOUT, int MODE /*0 = activation buffers, 1 = network output frames*/, bool F32OUT>
__global__ void __launch_bounds__(TC_THREADS, 1) tcg_conv_kernel(const __grid_constant__ TcgParams P) {
    using C = TcgCfg<NOUT, MODE>;
    extern __shared__ uint8_t smem_raw[];
    const int it_begin = P.item_first[blockIdx.x], it_end = P.item_first[blockIdx.x + 1];
    const int G = P.groups, R = P.ring_slots;
    const uint32_t WB = (uint32_t)(G * 9 * NOUT * TCG_PB);
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

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < R; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < TC_NBLK; ++b) {
            mbar_init(tfull_bar(b), 1);
            mbar_init(tempty_bar(b), 4);
        }
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
            s_slope[i] = P.slope[i];
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 0) {
        // ======================= TMA producer =======================
        if (lane == 0) {
            mbar_expect_tx(w_bar, WB);
            const uint32_t chunk = 3u * NOUT * TCG_PB;  // one (group, kx) tile
            for (int t = 0; t < 3 * G; ++t) bulk_g2s(w_s + t * chunk, P.wimg + (size_t)t * chunk, chunk, w_bar);
            int slot = 0;
            uint32_t phase = 0;
            for (int it = it_begin; it < it_end; ++it) {
                const TcItem I = P.items[it];
                const CUtensorMap* map = P.maps + (P.map_base + I.map);
                const int rows_in = I.rows + 2;
                for (int rho = 0; rho < rows_in; ++rho) {
                    const int y = I.y0 - 1 + rho;  // rows outside the plane are zero-filled by TMA = the conv's zero padding
                    for (int g = 0; g < G; ++g) {
                        mbar_wait(empty_bar(slot), phase ^ 1u, 0);
                        mbar_expect_tx(full_bar(slot), TCG_SUBROWB);
                        tma_load_4d(ring_s + slot * TCG_SUBROWB, map, full_bar(slot), g * 64, I.x0 - 1, y, I.plane);
                        if (++slot == R) {
                            slot = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
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
        uint32_t g0 = 0;      // CTA-local index of the current item's output row 0
        uint32_t gfresh = 0;  // CTA-local index of the next output row to be started
        mbar_wait(w_bar, 0, 1);
        uint32_t ok_full = mbar_test_wait(full_bar(0), 0);
        uint32_t ok_tempty = mbar_test_wait(tempty_bar(0), 0);
        for (int it = it_begin; it < it_end; ++it) {
            const int rows = P.items[it].rows;
            for (int rho = 0; rho < rows + 2; ++rho) {  // input row rho feeds output rows rho - ky, ky = 0..2
                const bool fresh = rho < rows;
                const uint32_t gn = gfresh + (fresh ? 1u : 0u);
                const int t0 = rho >= 2 ? rho - 2 : 0;
                const int t1 = fresh ? rho : rows - 1;
                const int cnt = t1 - t0 + 1;
                const uint32_t blk0 = (g0 + (uint32_t)t0) & (TC_NBLK - 1);
                const int wrap = (int)(TC_NBLK - blk0);
                const int n1 = cnt < wrap ? cnt : wrap;
                const uint32_t brow0 = rho >= 2 ? 0u : (uint32_t)(2 - rho);
                const uint32_t id1 = C::IDESC0 | ((uint32_t)((n1 * NOUT) >> 3) << 17);
                const uint32_t id2 = C::IDESC0 | ((uint32_t)((((cnt - n1) > 0 ? (cnt - n1) : 1) * NOUT) >> 3) << 17);
                const uint32_t d1 = tmem_base + blk0 * NOUT;
                for (int g = 0; g < G; ++g) {
                    if (!ok_full) mbar_wait(full_bar(slot), phase, 2);
                    if (g == 0 && fresh && !ok_tempty) mbar_wait(tempty_bar(gfresh & (TC_NBLK - 1)), (gfresh / TC_NBLK) & 1u, 3);
                    tc_fence_after();
                    const int nslot = slot + 1 == R ? 0 : slot + 1;
                    const uint32_t nphase = slot + 1 == R ? phase ^ 1u : phase;
                    ok_full = mbar_test_wait(full_bar(nslot), nphase);
                    if (g == G - 1) ok_tempty = mbar_test_wait(tempty_bar(gn & (TC_NBLK - 1)), (gn / TC_NBLK) & 1u);
                    const int ks = min(4, (P.cin - g * 64) >> 4);  // K = 16 slabs present in this group
                    if (elect_one_sync()) {
                        const uint32_t a_lo = ring_lo + (uint32_t)slot * (TCG_SUBROWB >> 4);
                        const uint32_t b_lo = w_lo + (uint32_t)g * GRPB + brow0 * BLKB;
                        const uint32_t b_lo2 = b_lo + (uint32_t)n1 * BLKB;
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            if (kx < kx_lo || kx >= kx_hi) continue;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (k >= ks) continue;
                                const uint32_t ao = (uint32_t)((kx * TCG_PB + k * 32) >> 4), bo = (uint32_t)(kx * KXB + ((k * 32) >> 4));
                                umma_f16(d1, hi64 | (a_lo + ao), hi64 | (b_lo + bo), id1, 1u);
                                if (n1 != cnt) umma_f16(tmem_base, hi64 | (a_lo + ao), hi64 | (b_lo2 + bo), id2, 1u);
                            }
                        }
                        umma_commit(empty_bar(slot));
                        if (g == G - 1 && rho >= 2) umma_commit(tfull_bar((g0 + (uint32_t)rho - 2u) & (TC_NBLK - 1)));
                    }
                    __syncwarp();
                    slot = nslot;
                    phase = nphase;
                }
                gfresh = gn;
            }
            g0 += (uint32_t)rows;
        }
    } else if (warp < 2 + 4 * TC_NSETS) {
        // ======================= epilogue =======================
        const int q = warp & 3;
        const uint32_t set = (warp - 2) >> 2;
        uint32_t tile_cnt = 0;
        for (uint32_t b = set; b < TC_NBLK; b += TC_NSETS) {
#pragma unroll
            for (int j = 0; j < NOUT; j += 16) tmem_st16_zero(tmem_base + ((uint32_t)(q * 32) << 16) + b * NOUT + j);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(b));
        }
        const float scale_acc = P.acc_scale;
        const int c = q * 32 + lane;
        for (int it = it_begin; it < it_end; ++it) {
            const TcItem I = P.items[it];
            const bool valid = c < I.w;
            for (int t = 0; t < I.rows; ++t, ++tile_cnt) {
                if (tile_cnt % TC_NSETS != set) continue;
                const uint32_t buf = tile_cnt % TC_NBLK;
                mbar_wait(tfull_bar(buf), (tile_cnt / TC_NBLK) & 1u, 4);
                tc_fence_after();
                uint32_t acc[NOUT];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * NOUT;
#pragma unroll
                for (int j = 0; j < NOUT; j += 16) tmem_ld16(taddr + j, acc + j);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < NOUT; j += 16) tmem_st16_zero(taddr + j);
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar(buf));

                if constexpr (MODE == 0) {
                    constexpr int CH = C::OB / 16;
                    uint4* stg = reinterpret_cast<uint4*>(gbase + stg_off + (warp - 2) * (32 * C::OB));
                    const long long pix = valid ? (long long)I.pix_off + (long long)(I.y0 + t) * I.Wt + I.x0 + c : -1;
                    const float4* sb4 = reinterpret_cast<const float4*>(s_bias);
                    const float4* ss4 = reinterpret_cast<const float4*>(s_slope);
#pragma unroll
                    for (int j = 0; j < NOUT; j += 8) {
                        const float4 b0 = sb4[j >> 2], b1 = sb4[(j >> 2) + 1], l0 = ss4[j >> 2], l1 = ss4[(j >> 2) + 1];
                        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                        const float sl[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
                        float v[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float a = fmaf(__uint_as_float(acc[j + e]), scale_acc, bb[e]);
                            v[e] = a < 0.f ? a * sl[e] : a;
                        }
#pragma unroll
                        for (int r = 0; r < 2; ++r) {
                            if (r >= P.nres) continue;
                            float rv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                            if (pix >= 0) {
                                if (P.res_f32[r]) {
                                    const float4* rp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(P.res_ptr[r]) + pix * P.res_ld[r] + j);
                                    const float4 r0 = rp[0], r1 = rp[1];
                                    rv[0] = r0.x, rv[1] = r0.y, rv[2] = r0.z, rv[3] = r0.w, rv[4] = r1.x, rv[5] = r1.y, rv[6] = r1.z, rv[7] = r1.w;
                                } else {
                                    const uint4 rr = *reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(P.res_ptr[r]) + pix * P.res_ld[r] + j);
                                    const __half2* h2 = reinterpret_cast<const __half2*>(&rr);
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        const float2 f = __half22float2(h2[e]);
                                        rv[2 * e] = f.x, rv[2 * e + 1] = f.y;
                                    }
                                }
                            }
                            const float cv = P.coef_v[r], cr = P.coef_r[r];
#pragma unroll
                            for (int e = 0; e < 8; ++e) v[e] = v[e] * cv + rv[e] * cr;
                        }
                        if (P.out32 && pix >= 0) {
                            float4* op = reinterpret_cast<float4*>(P.out32 + pix * P.out32_ld + j);
                            op[0] = make_float4(v[0], v[1], v[2], v[3]);
                            op[1] = make_float4(v[4], v[5], v[6], v[7]);
                        }
                        uint32_t pk[4];
#pragma unroll
                        for (int e = 0; e < 8; e += 2) {
                            __half2 h = __floats2half2_rn(v[e], v[e + 1]);
                            pk[e >> 1] = *reinterpret_cast<uint32_t*>(&h);
                        }
                        const int qi = lane * CH + (j >> 3);
                        stg[qi ^ ((qi >> 3) & 7)] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                    if (P.out16) {
                        __syncwarp();
                        uint8_t* outp = reinterpret_cast<uint8_t*>(P.out16);
#pragma unroll
                        for (int i = 0; i < CH; ++i) {
                            const int qi = i * 32 + lane;
                            const uint4 v4 = stg[qi ^ ((qi >> 3) & 7)];
                            const long long o = __shfl_sync(0xffffffffu, pix, qi / CH);
                            if (o >= 0) *reinterpret_cast<uint4*>(outp + (size_t)o * P.out16_ld * 2 + (qi % CH) * 16) = v4;
                        }
                        __syncwarp();
                    }
                } else {
                    // network output: (acc + bias) * 255, cropped to the tile core, as cv2.imwrite would store it
                    const int fy = I.fy0 + I.y0 + t, fx = I.fx0 + I.x0 + c;
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
            }
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
