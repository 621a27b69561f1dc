"""numpy interpreter of a ``ncnn_model.FusedProgram`` (the op list of ``b2sr_create_fused``) -- TEST INFRASTRUCTURE.

Two modes:
* ``exact=True``: float64 everywhere, no rounding -- must reproduce the oracle's graph interpreter bit for bit, which
  pins the lowering (fusion of adds into convolutions, deferred scheduling, buffer recycling, channel-slice views).
* ``exact=False``: the device arithmetic -- fp32 math, values stored in fp16 buffers rounded to fp16, raw 0..255
  input scaled by 1/255 after the first convolution's accumulation -- which bounds what the tcgen05 kernels may differ
  from the oracle by, and gives per-op expected buffers for GPU bring-up (``taps``).
"""
import numpy as np

from oracle import oracle
from upscale_video_b200 import ncnn_model as M


def run_fused(prog: M.FusedProgram, x, exact=True, taps=None, upto=None):
    """x: HWC image in [0, 1] (exact) or raw u8 (device mode).  Returns the network output HWC (unscaled, i.e. before
    the ``* 255``).  ``taps``: list that receives (op index, out16 array or None, out32 array or None) per op."""
    real = np.float64 if exact else np.float32
    prec = "f64" if exact else "f32"
    H, W, _ = x.shape
    bufs = [None] * len(prog.bufs)

    def buf(i):
        b = prog.bufs[i]
        if bufs[i] is None:
            bufs[i] = np.full((H * b["res"], W * b["res"], b["channels"]), np.nan, real)  # poison: unwritten slices show up
        return bufs[i]

    out = None
    for idx, o in enumerate(prog.ops):
        if o["type"] == M.FOP_NEAREST:
            src = buf(o["in_buf"])[:, :, o["in_off"]:o["in_off"] + o["cin"]]
            v = oracle.nearest(np.ascontiguousarray(src), float(o["r"]), float(o["r"]), prec)
        else:
            if o["in_buf"] < 0:
                src = np.asarray(x, real)
            else:
                src = buf(o["in_buf"])[:, :, o["in_off"]:o["in_off"] + o["cin"]]
            assert not np.isnan(src).any(), "op %d reads an unwritten slice" % idx
            w = prog.weights[o["w_off"]:o["w_off"] + o["cout"] * o["cin"] * o["k"] * o["k"]].reshape(o["cout"], o["cin"], o["k"], o["k"])
            b = prog.weights[o["b_off"]:o["b_off"] + o["cout"]] if o["b_off"] >= 0 else None
            if o["in_buf"] < 0 and not exact:
                v = oracle.conv(src, w, None, o["k"] // 2, 0, None, prec) * np.float32(1.0 / 255.0)
                if b is not None:
                    v = v + b.astype(real)
                if o["act"] == 2:
                    v = np.where(v < 0, v * real(np.float32(o["slope"])), v)
            else:
                v = oracle.conv(np.ascontiguousarray(src), w, b, o["k"] // 2, o["act"],
                                np.asarray([o["slope"]], np.float32) if o["act"] == 2 else None, prec)
            if o.get("sc_cin", 0):
                ws = prog.weights[o["sc_w_off"]:o["sc_w_off"] + o["cout"] * o["sc_cin"]].reshape(o["cout"], o["sc_cin"], 1, 1)
                sc = oracle.conv(np.ascontiguousarray(src[:, :, :o["sc_cin"]]), ws, None, 0, 0, None, prec)
                v = v * real(np.float32(o["sc_coef_v"])) + sc * real(np.float32(o["sc_coef_r"]))
            for q in range(o["nres"]):
                r = buf(o["res_buf"][q])[:, :, o["res_off"][q]:o["res_off"][q] + o["cout"]]
                assert not np.isnan(r).any(), "op %d residual %d reads an unwritten slice" % (idx, q)
                v = v * real(np.float32(o["coef_v"][q])) + r * real(np.float32(o["coef_r"][q]))
            v = v.astype(real)
        t16 = t32 = None
        if o["out16_buf"] >= 0:
            t16 = v if exact else v.astype(np.float16).astype(np.float32)
            buf(o["out16_buf"])[:, :, o["out16_off"]:o["out16_off"] + o["cout"]] = t16
        if o["out32_buf"] >= 0:
            t32 = v
            buf(o["out32_buf"])[:, :, o["out32_off"]:o["out32_off"] + o["cout"]] = t32
        if taps is not None:
            taps.append((idx, t16, t32, v if o["final"] else None))
        if o["final"]:
            out = v
        if upto is not None and idx >= upto:
            break
    return out
