"""Lane-by-lane numpy emulation of ``nlm_kernel`` (upscale_video_b200/csrc/nlm.cuh).

Test infrastructure: it follows the kernel's index arithmetic statement by statement -- tile origin, the staged
40 x 28 Lab window with the reflect-101 border, lane l on image column x0 + l - 2, shuffles as lane shifts that
return the lane's own value at the warp edge, the running 5-row column sums and the two-rows-late partner pixel -- so
that the CPU suite can check the warp algorithm against the oracle without a GPU.  The Lab conversions and the
weight tables are taken from the product library through its host-only C-ABI entry points.
"""
import numpy as np

TW, TH, BORDER = 28, 16, 6
SW, SH = TW + 2 * BORDER, TH + 2 * BORDER


def reflect101(p, n):
    if n == 1:
        return 0
    while p < 0 or p >= n:
        p = -p if p < 0 else 2 * n - 2 - p
    return p


def shfl_up(v, d):
    out = v.copy()
    out[d:] = v[:-d]
    return out


def shfl_down(v, d):
    out = v.copy()
    out[:-d] = v[d:]
    return out


PACK_CLAMP = 13107
PACK_MAX_TABLE = PACK_CLAMP >> 5


def run(lab, w_l, w_ab, packed=None, th=TH):
    """lab: H x W x 3 uint8 (already converted) -> denoised Lab, H x W x 3 uint8."""
    H, W, _ = lab.shape
    TH, SH = th, th + 2 * BORDER  # noqa: N806 -- rows per warp tile is a template parameter of the kernel (B2SR_NLM_TH)
    if packed is None:  # the host's choice (nlm_launch in csrc/nlm_host.inl)
        packed = len(w_l) <= PACK_MAX_TABLE and len(w_ab) <= PACK_MAX_TABLE
    out = np.zeros_like(lab)
    lane = np.arange(32)
    for y0 in range(0, H, TH):
        for x0 in range(0, W, TW):
            s = np.zeros((SH, SW, 3), dtype=np.int64)
            for sy in range(SH):
                gy = reflect101(y0 + sy - BORDER, H)
                for sx in range(SW):
                    s[sy, sx] = lab[gy, reflect101(x0 + sx - BORDER, W)]
            est = np.zeros((TH, 32, 3), dtype=np.uint64)
            ws = np.zeros((TH, 32, 2), dtype=np.uint64)
            cx = lane + BORDER - 2
            for dy in range(-4, 5):
                for dx in range(-4, 5):
                    v_l = np.zeros(32, dtype=np.int64)
                    v_c = np.zeros(32, dtype=np.int64)
                    dl_hist, dc_hist = [], []
                    q1 = q2 = np.zeros((32, 3), dtype=np.int64)
                    for r in range(TH + 4):
                        p = s[BORDER - 2 + r, cx]
                        q = s[BORDER - 2 + r + dy, cx + dx]
                        ad = np.abs(p - q)
                        d_l = ad[:, 0] ** 2
                        d_c = ad[:, 1] ** 2 + ad[:, 2] ** 2
                        dl_hist.append(d_l)
                        dc_hist.append(d_c)
                        v_l = v_l + d_l  # vertical running sum per column first
                        v_c = v_c + d_c
                        if r >= 5:
                            v_l = v_l - dl_hist[r - 5]
                            v_c = v_c - dc_hist[r - 5]
                        if r >= 4:  # then the horizontal 5-sum across lanes
                            o = r - 4
                            if packed:  # two clamped 16-bit fields in one 32-bit word through the shuffles
                                v = (np.minimum(v_l, PACK_CLAMP) | (np.minimum(v_c, PACK_CLAMP) << 16)).astype(np.uint32)
                                sv = v + shfl_up(v, 1) + shfl_up(v, 2) + shfl_down(v, 1) + shfl_down(v, 2)  # uint32: wraps like the GPU
                                k_l, k_c = ((sv & 0xffff) >> 5).astype(np.int64), (sv >> 21).astype(np.int64)
                            else:
                                k_l = (v_l + shfl_up(v_l, 1) + shfl_up(v_l, 2) + shfl_down(v_l, 1) + shfl_down(v_l, 2)) >> 5
                                k_c = (v_c + shfl_up(v_c, 1) + shfl_up(v_c, 2) + shfl_down(v_c, 1) + shfl_down(v_c, 2)) >> 5
                            wl = np.where(k_l < len(w_l), w_l[np.minimum(k_l, len(w_l) - 1)], 0).astype(np.uint64)
                            wc = np.where(k_c < len(w_ab), w_ab[np.minimum(k_c, len(w_ab) - 1)], 0).astype(np.uint64)
                            est[o, :, 0] += wl * q2[:, 0].astype(np.uint64)
                            est[o, :, 1] += wc * q2[:, 1].astype(np.uint64)
                            est[o, :, 2] += wc * q2[:, 2].astype(np.uint64)
                            ws[o, :, 0] += wl
                            ws[o, :, 1] += wc
                        q2, q1 = q1, q
            assert est.max() < 2 ** 32 and ws.max() < 2 ** 32  # the kernel's accumulators are 32-bit
            for l in range(2, 2 + TW):
                x = x0 + l - 2
                if x >= W:
                    continue
                for o in range(TH):
                    y = y0 + o
                    if y >= H:
                        continue
                    out[y, x, 0] = (est[o, l, 0] + ws[o, l, 0] // 2) // ws[o, l, 0]
                    out[y, x, 1] = (est[o, l, 1] + ws[o, l, 1] // 2) // ws[o, l, 1]
                    out[y, x, 2] = (est[o, l, 2] + ws[o, l, 1] // 2) // ws[o, l, 1]
    return out
