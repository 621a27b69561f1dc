"""Time the generic graph engine on 4x_Valar_v1 (run under gpurun; also the ncu target for its kernels).

    python tools/valar_pass.py [--h 540 --w 960 --reps 3 --impl 0]
"""
import argparse
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from upscale_video_b200 import engine as E, ncnn_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--h", type=int, default=540)
ap.add_argument("--w", type=int, default=960)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--impl", type=int, default=0)
a = ap.parse_args()
eng = E.Engine.from_files(ncnn_model.packaged_model_dir(), "4x_Valar_v1", 0)
eng.set_option(E.OPT_IMPL, a.impl)
img = np.random.default_rng(0).integers(0, 256, (a.h, a.w, 3), dtype=np.uint8)
eng.run_u8(img)
ts = []
for _ in range(a.reps):
    t0 = time.perf_counter()
    eng.run_u8(img)
    ts.append(time.perf_counter() - t0)
macs = sum(int(np.prod(l.weights["weight"].shape)) for l in ncnn_model.load_model(ncnn_model.packaged_model_dir(), "4x_Valar_v1").convs())
print("valar %dx%d: best %.1f ms/frame, launches %d (hmma %d)" % (a.h, a.w, min(ts) * 1e3, eng.stat(E.STAT_LAUNCHES), eng.stat(E.STAT_HMMA_LAUNCHES)))
