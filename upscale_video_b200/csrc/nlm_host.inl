// nlm_host.inl -- host side of the denoise pass (included at the end of engine.cu; uses its fail()/CUDA_TRY/grow).
// Table construction follows OpenCV's published fixed-point definitions (the reference's only call on this path is
// cv2.fastNlMeansDenoisingColored, upscale/upscale_processing.py:354); tests/test_nlmeans_oracle.py checks every table
// built here against oracle/nlmeans.py, which is itself pinned against cv2.

#include <cmath>

namespace {

struct NlmLabTables {
    int32_t fwd[9], inv[9];
    int32_t l2y[256], l2fy[256];
    uint16_t cbrt_tab[3072];
};

// volatile stores force every intermediate to be rounded to float, whatever the host compiler contracts
static inline float f32(float v) {
    volatile float t = v;
    return t;
}

static void nlm_build_lab_tables(NlmLabTables* T) {
    static const float m[9] = {0.412453f, 0.357580f, 0.180423f, 0.212671f, 0.715160f, 0.072169f, 0.019334f, 0.119193f, 0.950227f};
    static const float x2r[9] = {3.240479f, -1.53715f, -0.498535f, -0.969256f, 1.875991f, 0.041556f, 0.055648f, -0.204043f, 1.057311f};
    static const float wp[3] = {0.950456f, 1.0f, 1.088754f};
    for (int i = 0; i < 3; ++i) {
        const float s = f32(4096.0f / wp[i]);
        for (int j = 0; j < 3; ++j) T->fwd[i * 3 + j] = (int32_t)lrintf(f32(m[i * 3 + j] * s));
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T->inv[i * 3 + j] = (int32_t)lrint((4096.0 * (double)x2r[i * 3 + j]) * (double)wp[j]);
    const float lth = f32(216.0f / 24389.0f), lsc = f32(841.0f / 108.0f), lb = f32(16.0f / 116.0f);
    for (int i = 0; i < 3072; ++i) {
        const float x = f32((float)i / 2040.0f);
        const float v = x < lth ? f32(f32(x * lsc) + lb) : (float)cbrt((double)x);
        T->cbrt_tab[i] = (uint16_t)lrintf(f32(32768.0f * v));
    }
    // an exact .5 tie after the float rounding of the cube root; OpenCV's soft-float cube root lands one ulp lower
    // (the only entry where a correctly rounded cube root and cv2 disagree -- found by the exhaustive 2^24 comparison)
    T->cbrt_tab[324] = 17745;
    const int base = 1 << 14;
    for (int k = 0; k < 256; ++k) {
        if (k <= 20) {
            T->l2y[k] = (int32_t)lrintf(f32((float)(k * base * 20 * 9) / (float)(17 * 24389)));
            T->l2fy[k] = (int32_t)lrintf(f32((float)base * f32(f32(16.0f / 116.0f) + f32((float)(k * 5) / (float)(3 * 17 * 29)))));
        } else {
            const float fy = f32(f32((float)(k * 100 * base) / (float)(255 * 116)) + f32((float)(16 * base) / 116.0f));
            T->l2fy[k] = (int32_t)lrintf(fy);
            T->l2y[k] = (int32_t)lrintf(f32(f32(f32(fy * fy) * fy) / (float)((double)base * base)));
        }
    }
}

// almost_dist2weight of OpenCV's FastNlMeansDenoisingInvoker<_, int, unsigned, DistSquared, int>
static std::vector<int32_t> nlm_weight_table(float h, int channels) {
    const int fixed_point_mult = 2147483647 / (9 * 9 * 255);
    const double mult = (double)(1 << NLM_BIN_SHIFT) / 25.0;
    const int n = (int)(255.0 * 255.0 * channels / mult + 1);
    const float den = f32(f32(h * h) * (float)channels);
    std::vector<int32_t> tab((size_t)n);
    for (int k = 0; k < n; ++k) {
        double w = std::exp(-(k * mult) / (double)den);
        if (std::isnan(w)) w = 1.0;
        int32_t weight = (int32_t)lrint(fixed_point_mult * w);
        if (weight < 0.001 * fixed_point_mult) weight = 0;
        tab[(size_t)k] = weight;
    }
    return tab;
}

}  // namespace

struct b2sr_nlm {
    int device = 0;
    cudaStream_t stream = nullptr, copy_in = nullptr, copy_out = nullptr;
    NlmLabTables host_tabs;
    uint8_t* d_tabs = nullptr;  // l2y | l2fy | cbrt
    int32_t *d_wl = nullptr, *d_wab = nullptr;
    uint32_t n_l = 0, n_ab = 0;
    float h_l = -1.f, h_ab = -1.f;  // levels the resident weight tables were built for
    uint8_t *d_in = nullptr, *d_out = nullptr, *d_in2 = nullptr, *d_out2 = nullptr;
    size_t cap_in = 0, cap_out = 0, cap_in2 = 0, cap_out2 = 0;
    double n_launch = 0;
};

extern "C" int b2sr_nlm_weight_table(float h, int channels, int32_t* out, int cap) {
    if (!(h > 0.f) || (channels != 1 && channels != 2)) return fail(B2SR_E_INVALID, "b2sr_nlm_weight_table: h %g channels %d", h, channels);
    std::vector<int32_t> t = nlm_weight_table(h, channels);
    if (out) memcpy(out, t.data(), sizeof(int32_t) * std::min<size_t>(t.size(), (size_t)std::max(cap, 0)));
    return (int)t.size();
}

extern "C" int b2sr_nlm_lab_tables(int32_t* fwd9, int32_t* inv9, int32_t* l2y256, int32_t* l2fy256, uint16_t* cbrt3072) {
    NlmLabTables T;
    nlm_build_lab_tables(&T);
    if (fwd9) memcpy(fwd9, T.fwd, sizeof T.fwd);
    if (inv9) memcpy(inv9, T.inv, sizeof T.inv);
    if (l2y256) memcpy(l2y256, T.l2y, sizeof T.l2y);
    if (l2fy256) memcpy(l2fy256, T.l2fy, sizeof T.l2fy);
    if (cbrt3072) memcpy(cbrt3072, T.cbrt_tab, sizeof T.cbrt_tab);
    return 0;
}

extern "C" void b2sr_nlm_destroy(b2sr_nlm* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (void* p : {(void*)c->d_tabs, (void*)c->d_wl, (void*)c->d_wab, (void*)c->d_in, (void*)c->d_out, (void*)c->d_in2, (void*)c->d_out2})
        if (p) cudaFree(p);
    for (cudaStream_t s : {c->stream, c->copy_in, c->copy_out})
        if (s) cudaStreamDestroy(s);
    delete c;
}

extern "C" int b2sr_nlm_create(b2sr_nlm** out, int device) {
    if (!out) return fail(B2SR_E_INVALID, "b2sr_nlm_create: null argument");
    *out = nullptr;
    cudaDeviceProp prop;
    TRY(check_device(device, &prop));
    b2sr_nlm* c = new b2sr_nlm();
    c->device = device;
    nlm_build_lab_tables(&c->host_tabs);
    auto body = [&]() -> int {
        CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking));
        const size_t nb = sizeof c->host_tabs.l2y + sizeof c->host_tabs.l2fy + sizeof c->host_tabs.cbrt_tab;
        CUDA_TRY(cudaMalloc(&c->d_tabs, nb));
        CUDA_TRY(cudaMemcpyAsync(c->d_tabs, c->host_tabs.l2y, sizeof c->host_tabs.l2y, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(c->d_tabs + 1024, c->host_tabs.l2fy, sizeof c->host_tabs.l2fy, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(c->d_tabs + 2048, c->host_tabs.cbrt_tab, sizeof c->host_tabs.cbrt_tab, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        return 0;
    };
    const int rc = body();
    if (rc) {
        b2sr_nlm_destroy(c);
        return rc;
    }
    *out = c;
    return 0;
}

// (re)build the device weight tables when the levels change; only the non-zero head of each table is uploaded
static int nlm_set_levels(b2sr_nlm* c, float h_l, float h_ab) {
    if (!(h_l > 0.f) || !(h_ab > 0.f) || h_l > 1000.f || h_ab > 1000.f)
        return fail(B2SR_E_INVALID, "denoise levels h = %g, hColor = %g (need 0 < h <= 1000)", h_l, h_ab);
    if (h_l == c->h_l && h_ab == c->h_ab) return 0;
    CUDA_TRY(cudaStreamSynchronize(c->stream));  // a previous launch may still read the old tables
    for (int plane = 0; plane < 2; ++plane) {
        std::vector<int32_t> t = nlm_weight_table(plane ? h_ab : h_l, plane ? 2 : 1);
        size_t n = t.size();
        while (n > 1 && t[n - 1] == 0) --n;
        int32_t** d = plane ? &c->d_wab : &c->d_wl;
        if (*d) cudaFree(*d);
        *d = nullptr;
        CUDA_TRY(cudaMalloc(d, n * sizeof(int32_t)));
        CUDA_TRY(cudaMemcpy(*d, t.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice));
        (plane ? c->n_ab : c->n_l) = (uint32_t)n;
    }
    c->h_l = h_l, c->h_ab = h_ab;
    return 0;
}

static int nlm_check(int n, int h, int w, int template_window, int search_window) {
    if (n < 1 || h < 1 || w < 1) return fail(B2SR_E_INVALID, "empty image (%d frames of %dx%d)", n, h, w);
    if (h > 16384 || w > 16384) return fail(B2SR_E_INVALID, "image %dx%d too large", h, w);
    if (template_window != 5 || search_window != 9)
        return fail(B2SR_E_UNSUPPORTED, "templateWindowSize %d / searchWindowSize %d: only the reference's 5 / 9 is built",
                    template_window, search_window);
    return 0;
}

static int nlm_launch(b2sr_nlm* c, const uint8_t* d_in, long long in_frame_stride, int in_stride, uint8_t* d_out,
                      long long out_frame_stride, int out_stride, int n, int h, int w) {
    NlmParams P;
    P.in = d_in, P.out = d_out;
    P.in_frame_stride = in_frame_stride, P.out_frame_stride = out_frame_stride;
    P.in_stride = in_stride, P.out_stride = out_stride;
    P.H = h, P.W = w;
    P.tiles_x = (w + NLM_TW - 1) / NLM_TW;
    // Rows per warp tile (B2SR_NLM_TH overrides): fewer rows = more template-halo rework (20/16 -> 16/12 -> 12/8 rows walked per
    // output row) but fewer accumulators per lane (80 -> 60 -> 40) and more resident warps.  Measured on B200, 16 frames of
    // 1080p per launch (profiles/r02i_checks.log): packed variant (levels <= 6) 16 rows 3 319 / 12 rows 3 388 / 8 rows 3 235
    // frames/s (within run-to-run noise: the bench measured 2 824 with 12 rows against 3 015 - 3 078 with 16 on other boxes);
    // generic variant (level 10) 2 608 / 2 500 / 2 902 -- so 16 rows stay for the packed kernel, 8 for the generic one.
    static const int th_env = getenv("B2SR_NLM_TH") ? atoi(getenv("B2SR_NLM_TH")) : 0;
    static const bool allow_packed = !(getenv("B2SR_NLM_PACKED") && atoi(getenv("B2SR_NLM_PACKED")) == 0);
    const bool packed = allow_packed && c->n_l <= NLM_PACK_MAX_TABLE && c->n_ab <= NLM_PACK_MAX_TABLE;
    const int TH = (th_env == 8 || th_env == 12 || th_env == 16) ? th_env : (packed ? 16 : 8);
    P.tiles_per_frame = P.tiles_x * ((h + TH - 1) / TH);
    P.n_tiles = (long long)P.tiles_per_frame * n;
    memcpy(P.fwd, c->host_tabs.fwd, sizeof P.fwd);
    memcpy(P.inv, c->host_tabs.inv, sizeof P.inv);
    P.l2y = (const int32_t*)c->d_tabs;
    P.l2fy = (const int32_t*)(c->d_tabs + 1024);
    P.cbrt_tab = (const uint16_t*)(c->d_tabs + 2048);
    P.w_l = c->d_wl, P.w_ab = c->d_wab;
    P.n_l = c->n_l, P.n_ab = c->n_ab;
    const long long blocks = (P.n_tiles + NLM_WARPS - 1) / NLM_WARPS;
    if (blocks > 0x7fffffffLL) return fail(B2SR_E_INVALID, "too many tiles (%lld)", P.n_tiles);
    const dim3 grid((unsigned)blocks), block(NLM_WARPS * 32);
#define NLM_LAUNCH(PK, T) nlm_kernel<PK, T><<<grid, block, 0, c->stream>>>(P)
    if (TH == 8) packed ? NLM_LAUNCH(true, 8) : NLM_LAUNCH(false, 8);
    else if (TH == 12) packed ? NLM_LAUNCH(true, 12) : NLM_LAUNCH(false, 12);
    else packed ? NLM_LAUNCH(true, 16) : NLM_LAUNCH(false, 16);
#undef NLM_LAUNCH
    CUDA_TRY(cudaGetLastError());
    c->n_launch += 1;
    return 0;
}

extern "C" int b2sr_nlm_run_u8(b2sr_nlm* c, const uint8_t* in, int h, int w, int in_stride, uint8_t* out, int out_stride,
                               float h_luma, float h_color, int template_window, int search_window, int memspace) {
    if (!c || !in || !out) return fail(B2SR_E_INVALID, "b2sr_nlm_run_u8: null argument");
    TRY(nlm_check(1, h, w, template_window, search_window));
    if (in_stride < w * 3 || out_stride < w * 3) return fail(B2SR_E_INVALID, "row stride smaller than a row");
    {   // every output pixel reads a 13 x 13 neighbourhood of the input: the two images must not share memory
        const uint8_t *a0 = in, *a1 = in + (size_t)(h - 1) * in_stride + (size_t)w * 3;
        const uint8_t *b0 = out, *b1 = out + (size_t)(h - 1) * out_stride + (size_t)w * 3;
        if (a0 < b1 && b0 < a1) return fail(B2SR_E_INVALID, "b2sr_nlm_run_u8: input and output overlap (in-place is not supported)");
    }
    CUDA_TRY(cudaSetDevice(c->device));
    TRY(nlm_set_levels(c, h_luma, h_color));
    if (memspace == B2SR_MEM_DEVICE) {
        TRY(nlm_launch(c, in, 0, in_stride, out, 0, out_stride, 1, h, w));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        return 0;
    }
    if (memspace != B2SR_MEM_HOST) return fail(B2SR_E_INVALID, "memspace %d", memspace);
    const size_t row = (size_t)w * 3, bytes = row * h;
    TRY(grow(&c->d_in, &c->cap_in, bytes));
    TRY(grow(&c->d_out, &c->cap_out, bytes));
    CUDA_TRY(cudaMemcpy2DAsync(c->d_in, row, in, (size_t)in_stride, row, (size_t)h, cudaMemcpyHostToDevice, c->stream));
    TRY(nlm_launch(c, c->d_in, 0, (int)row, c->d_out, 0, (int)row, 1, h, w));
    CUDA_TRY(cudaMemcpy2DAsync(out, (size_t)out_stride, c->d_out, row, row, (size_t)h, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int b2sr_nlm_run_batch_device(b2sr_nlm* c, const uint8_t* d_in, uint8_t* d_out, int n, int h, int w, float h_luma,
                                         float h_color, int sync) {
    if (!c || !d_in || !d_out) return fail(B2SR_E_INVALID, "b2sr_nlm_run_batch_device: null argument");
    TRY(nlm_check(n, h, w, 5, 9));
    CUDA_TRY(cudaSetDevice(c->device));
    TRY(nlm_set_levels(c, h_luma, h_color));
    const long long frame = (long long)h * w * 3;
    TRY(nlm_launch(c, d_in, frame, w * 3, d_out, frame, w * 3, n, h, w));
    if (sync) CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

// n host frames through a double-buffered H2D -> kernel -> D2H pipeline (three streams); synchronous.
extern "C" int b2sr_nlm_run_batch_host(b2sr_nlm* c, const uint8_t* h_in, uint8_t* h_out, int n, int h, int w, float h_luma,
                                       float h_color) {
    if (!c || !h_in || !h_out) return fail(B2SR_E_INVALID, "b2sr_nlm_run_batch_host: null argument");
    TRY(nlm_check(n, h, w, 5, 9));
    CUDA_TRY(cudaSetDevice(c->device));
    TRY(nlm_set_levels(c, h_luma, h_color));
    const size_t frame = (size_t)h * w * 3;
    // small chunks: the kernel needs ~0.3 ms per 1080p frame, about what PCIe needs per direction, so the three streams only
    // overlap well when a chunk is a frame or two (one 1080p frame already fills the GPU twice over: 4692 warp tiles)
    static const int chunk_env = getenv("B2SR_NLM_CHUNK") ? atoi(getenv("B2SR_NLM_CHUNK")) : 0;
    const int B = chunk_env > 0 ? std::min(chunk_env, 64)
                                : (int)std::max<size_t>(1, std::min<size_t>(2, ((size_t)16 << 20) / frame));
    TRY(grow(&c->d_in, &c->cap_in, frame * B));
    TRY(grow(&c->d_out, &c->cap_out, frame * B));
    TRY(grow(&c->d_in2, &c->cap_in2, frame * B));
    TRY(grow(&c->d_out2, &c->cap_out2, frame * B));
    uint8_t* din[2] = {c->d_in, c->d_in2};
    uint8_t* dout[2] = {c->d_out, c->d_out2};
    cudaEvent_t in_ready[2], compute_done[2], out_done[2];
    for (int i = 0; i < 2; ++i) {
        CUDA_TRY(cudaEventCreateWithFlags(&in_ready[i], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&compute_done[i], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&out_done[i], cudaEventDisableTiming));
    }
    int rc = 0, k = 0;
    for (int f = 0; f < n && !rc; f += B, ++k) {
        const int nb = std::min(B, n - f), s = k & 1;
        if (k >= 2) {  // slot s is free once chunk k-2's kernel has read din[s] and its D2H has read dout[s]
            cudaStreamWaitEvent(c->copy_in, compute_done[s], 0);
            cudaStreamWaitEvent(c->stream, out_done[s], 0);
        }
        cudaMemcpyAsync(din[s], h_in + (size_t)f * frame, frame * nb, cudaMemcpyHostToDevice, c->copy_in);
        cudaEventRecord(in_ready[s], c->copy_in);
        cudaStreamWaitEvent(c->stream, in_ready[s], 0);
        rc = nlm_launch(c, din[s], (long long)frame, w * 3, dout[s], (long long)frame, w * 3, nb, h, w);
        if (rc) break;
        cudaEventRecord(compute_done[s], c->stream);
        cudaStreamWaitEvent(c->copy_out, compute_done[s], 0);
        cudaMemcpyAsync(h_out + (size_t)f * frame, dout[s], frame * nb, cudaMemcpyDeviceToHost, c->copy_out);
        cudaEventRecord(out_done[s], c->copy_out);
    }
    cudaError_t e1 = cudaStreamSynchronize(c->copy_in), e2 = cudaStreamSynchronize(c->stream), e3 = cudaStreamSynchronize(c->copy_out);
    for (int i = 0; i < 2; ++i) {
        cudaEventDestroy(in_ready[i]);
        cudaEventDestroy(compute_done[i]);
        cudaEventDestroy(out_done[i]);
    }
    if (rc) return rc;
    for (cudaError_t e : {e1, e2, e3})
        if (e != cudaSuccess) return fail(B2SR_E_CUDA, "b2sr_nlm_run_batch_host: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int b2sr_nlm_synchronize(b2sr_nlm* c) {
    if (!c) return fail(B2SR_E_INVALID, "null context");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" void* b2sr_nlm_stream(b2sr_nlm* c) { return c ? (void*)c->stream : nullptr; }
extern "C" double b2sr_nlm_launches(b2sr_nlm* c) { return c ? c->n_launch : 0.0; }
