#!/usr/bin/env python
"""GPU bring-up report (run by hand under gpurun; not collected by pytest).

    python tests/bringup_gpu.py            # runs every mode in its own subprocess (a device trap kills the context)
    python tests/bringup_gpu.py MODE [MODEL H W]   # one mode: simple | layer | pipe

For one small frame it compares every layer's activations (b2sr_debug_layer) and the final u8 image of the CUDA
engine against the CPU oracle, for the CUDA-core path and the two tcgen05 schedules (layer by layer, pipelined).
"""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def stats(name, got, ref):
    d = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    scale = np.abs(ref).max() + 1e-9
    bad = np.argwhere(d > 0.02 * scale + 0.02)
    msg = "%-22s max|d|=%9.5f  mean|d|=%9.6f  ref|max|=%8.3f  bad=%d/%d" % (name, d.max(), d.mean(), scale, len(bad), d.size)
    if len(bad):
        ys, xs, cs = bad[:, 0], bad[:, 1], bad[:, 2]
        msg += "  bad y[%d..%d] x[%d..%d] c[%d..%d] first=%s got=%.4f ref=%.4f" % (
            ys.min(), ys.max(), xs.min(), xs.max(), cs.min(), cs.max(), tuple(bad[0]), got[tuple(bad[0])], ref[tuple(bad[0])])
    print(msg, flush=True)
    return len(bad) == 0


def run_mode(mode, model="2x_Compact_Pretrain", shape=(40, 56), seed=1):
    from oracle import oracle
    from upscale_video_b200 import engine as E
    from upscale_video_b200 import ncnn_model

    mdir = ncnn_model.packaged_model_dir()
    layers = oracle.read_model(mdir, model)
    g = ncnn_model.load_model(mdir, model)
    rng = np.random.default_rng(seed)
    # smooth-ish natural-like content + noise so outputs are not saturated
    h, w = shape
    base = np.linspace(30, 220, w)[None, :, None] * np.ones((h, 1, 3)) + rng.normal(0, 12, (h, w, 3))
    img = np.clip(base, 0, 255).astype(np.uint8)
    taps = {}
    y_ref = oracle.run_graph(layers, oracle.from_pixels_normalize(img), "f32", taps=taps)
    prelu_names = [L["tops"][0] for L in layers if L["type"] == "PReLU"]
    eng = E.Engine(g, 0)
    scale = eng.scale
    print("== mode %s model %s image %s device %s" % (mode, model, img.shape, E.device_name(0)), flush=True)
    eng.set_option(E.OPT_IMPL, {"simple": E.IMPL_SIMPLE, "layer": E.IMPL_TCGEN05, "pipe": E.IMPL_PIPELINED}[mode])
    ok = True
    for li, name in enumerate(prelu_names if mode != "pipe" else []):
        t0 = time.time()
        got = eng.debug_layer(img, li)
        good = stats("layer %2d (%s)" % (li, name), got, taps[name])
        ok &= good
        if not good and li >= 2:
            break
    out = eng.run_u8(img, tile=0, halo=0)
    ref = oracle.apply_model_array(layers, img, "f64") if scale == 1 else oracle.upscale_image_array(layers, img, scale, "f64", tile_size=10**6)
    d = np.abs(out.astype(int) - ref.astype(int))
    print("final u8: max|d|=%d  frac(d>=1)=%.4f frac(d>1)=%.6f" % (d.max(), (d >= 1).mean(), (d > 1).mean()), flush=True)
    ok &= d.max() <= 1
    print("== mode %s %s" % (mode, "OK" if ok else "MISMATCH"), flush=True)
    return ok


if __name__ == "__main__":
    if len(sys.argv) > 1:
        m = sys.argv[1]
        model = sys.argv[2] if len(sys.argv) > 2 else "2x_Compact_Pretrain"
        shape = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (40, 56)
        sys.exit(0 if run_mode(m, model, shape) else 1)
    rc = 0
    for m in ("simple", "layer", "pipe"):
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), m], timeout=300)
            print("-- mode %s exit %d" % (m, r.returncode), flush=True)
            rc |= r.returncode != 0
        except subprocess.TimeoutExpired:
            print("-- mode %s TIMEOUT" % m, flush=True)
            rc = 1
    sys.exit(rc)
