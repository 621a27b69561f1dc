"""GPU bring-up of the fused tcgen05 graph engine (run under gpurun, not a pytest file):

    python tests/bringup_fused.py [--h 40 --w 300] [--full]

Runs 4x_Valar_v1 op by op (b2sr_debug_fused) against the numpy emulation of the fused program in device arithmetic
(tests/fused_emulator.py) and prints the first ops whose output buffers disagree, then the end-to-end parity against
the oracle and a 540p timing.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fused_emulator as FE  # noqa: E402
from oracle import oracle  # noqa: E402
from upscale_video_b200 import engine as E, ncnn_model as M  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--h", type=int, default=40)
ap.add_argument("--w", type=int, default=300)
ap.add_argument("--full", action="store_true", help="check every op instead of a sample")
ap.add_argument("--skip-ops", action="store_true")
a = ap.parse_args()

md = M.packaged_model_dir()
eng = E.Engine.from_files(md, "4x_Valar_v1", 0)
assert eng.fused, "fused engine not selected"
prog = eng.program
rng = np.random.default_rng(3)
base = np.linspace(30, 220, a.w)[None, :, None] * np.ones((a.h, 1, 3)) + rng.normal(0, 12, (a.h, a.w, 3))
img = np.clip(base, 0, 255).astype(np.uint8)

if not a.skip_ops:
    taps = []
    t0 = time.time()
    FE.run_fused(prog, img, exact=False, taps=taps)
    print("emulation %.1f s" % (time.time() - t0), flush=True)
    n = len(prog.ops)
    sample = list(range(n - 1)) if a.full else sorted(set(list(range(0, 22)) + list(range(22, n - 8, 37)) + list(range(n - 8, n - 1))))
    bad = 0
    for i in sample:
        o = prog.ops[i]
        _, t16, t32, _ = taps[i]
        msgs = []
        for name, exp, b, off in (("out16", t16, o["out16_buf"], o["out16_off"]), ("out32", t32, o["out32_buf"], o["out32_off"])):
            if exp is None:
                continue
            got = eng.debug_fused(img, i, b)[:, :, off:off + o["cout"]]
            err = np.abs(got - exp)
            tol = 4e-3 * max(1e-3, float(np.abs(exp).max()))
            if not np.isfinite(got).all() or err.max() > tol:
                y, x, ch = np.unravel_index(np.argmax(np.where(np.isfinite(err), err, np.inf)), err.shape)
                msgs.append("%s buf %d: max err %.4g (scale %.3g) at y=%d x=%d c=%d got %.5g exp %.5g, bad px %.2f%%, bad rows %s cols %s ch %s" % (
                    name, b, err.max(), np.abs(exp).max(), y, x, ch, got[y, x, ch], exp[y, x, ch], 100 * (err > tol).any(axis=2).mean(),
                    np.unique(np.nonzero(err > tol)[0])[:8], np.unique(np.nonzero(err > tol)[1])[:8], np.unique(np.nonzero(err > tol)[2])[:8]))
        if msgs:
            bad += 1
            print("op %d (k=%d cin=%d cout=%d nres=%d res=%d type=%d): MISMATCH" % (i, o["k"], o["cin"], o["cout"], o["nres"], o["res"], o["type"]))
            for m in msgs:
                print("   ", m)
            if bad >= 6:
                break
    print("per-op check: %d of %d sampled ops disagree" % (bad, len(sample)), flush=True)

ref = oracle.upscale_image_array(oracle.read_model(md, "4x_Valar_v1"), img, 4, "f32", tile_size=960, halo=10)
out = eng.run_u8(img)
d = np.abs(out.astype(int) - ref.astype(int))
print("end to end %dx%d: max |diff| %d LSB, mismatching %.3f%%, tcgen05 launches %d" % (a.h, a.w, d.max(), 100 * (d > 0).mean(), eng.stat(E.STAT_TC_LAUNCHES)), flush=True)
g = np.load(os.path.join(ROOT, "tests", "golden", "valar4x_crop.npz"))
d = np.abs(eng.run_u8(g["x"]).astype(int) - g["y"].astype(int))
print("golden crop: max |diff| %d LSB, mismatching %.3f%%" % (d.max(), 100 * (d > 0).mean()), flush=True)
img2 = np.clip(np.linspace(30, 220, 980)[None, :, None] * np.ones((20, 1, 3)) + rng.normal(0, 12, (20, 980, 3)), 0, 255).astype(np.uint8)
ref2 = oracle.upscale_image_array(oracle.read_model(md, "4x_Valar_v1"), img2, 4, "f32")
d = np.abs(eng.run_u8(img2).astype(int) - ref2.astype(int))
print("two tiles 20x980: max |diff| %d LSB, mismatching %.3f%%" % (d.max(), 100 * (d > 0).mean()), flush=True)

big = rng.integers(0, 256, (540, 960, 3), dtype=np.uint8)
eng.run_u8(big)
ts = []
for _ in range(3):
    t0 = time.perf_counter()
    eng.run_u8(big)
    ts.append(time.perf_counter() - t0)
eng.set_option(E.OPT_PROFILE, 1)
eng.reset_stats()
eng.run_u8(big)
print("540p: best %.1f ms/frame host-to-host; device time of all launches %.1f ms (%.0f TFLOP/s)" % (
    min(ts) * 1e3, eng.stat(E.STAT_ALL_MS), 18.73 / max(eng.stat(E.STAT_ALL_MS), 1e-3) * 1e3), flush=True)
