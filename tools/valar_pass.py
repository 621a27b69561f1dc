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
ap.add_argument("--batch", type=int, default=0, help="also time N device-resident frames per call (b2sr_run_batch_device)")
ap.add_argument("--chain", type=int, default=-1, help="compile_fused fp32_chain override")
ap.add_argument("--debug", type=int, default=0, help="print stall accounting of the first N fused launches (and the last 5)")
a = ap.parse_args()
if a.chain >= 0:
    _cf = ncnn_model.compile_fused
    ncnn_model.compile_fused = lambda g, *x, **k: _cf(g, *x, fp32_chain=a.chain, **k)
eng = E.Engine.from_files(ncnn_model.packaged_model_dir(), "4x_Valar_v1", 0)
eng.set_option(E.OPT_IMPL, a.impl)
img = np.random.default_rng(0).integers(0, 256, (a.h, a.w, 3), dtype=np.uint8)
eng.run_u8(img)
if a.debug:
    eng.set_option(E.OPT_PIPE_DEBUG, a.debug)
    eng.run_u8(img)
    eng.set_option(E.OPT_PIPE_DEBUG, 0)
ts = []
for _ in range(a.reps):
    t0 = time.perf_counter()
    eng.run_u8(img)
    ts.append(time.perf_counter() - t0)
macs = sum(int(np.prod(l.weights["weight"].shape)) for l in ncnn_model.load_model(ncnn_model.packaged_model_dir(), "4x_Valar_v1").convs())
print("valar %dx%d: best %.1f ms/frame, launches %d (hmma %d)" % (a.h, a.w, min(ts or [float("nan")]) * 1e3, eng.stat(E.STAT_LAUNCHES), eng.stat(E.STAT_HMMA_LAUNCHES)))

if a.batch:
    import torch
    frames = torch.from_numpy(np.random.default_rng(1).integers(0, 256, (a.batch, a.h, a.w, 3), dtype=np.uint8)).cuda()
    out = torch.empty((a.batch, a.h * 4, a.w * 4, 3), dtype=torch.uint8, device="cuda")
    eng.run_batch_device(frames, out, a.batch, a.h, a.w, sync=True)
    best = 1e9
    for _ in range(max(1, a.reps)):
        t0 = time.perf_counter()
        eng.run_batch_device(frames, out, a.batch, a.h, a.w, sync=True)
        best = min(best, time.perf_counter() - t0)
    for i in (0, a.batch - 1):
        one = eng.run_u8(frames[i].cpu().numpy())
        d = np.abs(out[i].cpu().numpy().astype(int) - one.astype(int))
        print("frame %d: batch vs single: max |diff| %d, differing values %d of %d (rows %s)" % (i, d.max(), (d > 0).sum(), d.size, np.unique(np.nonzero(d)[0])[:12]))
    print("valar batch of %d: %.1f ms/frame (%.0f TFLOP/s on 18.73 TFLOP per 540p frame)" % (a.batch, best / a.batch * 1e3, 18.73 * a.h * a.w / (540 * 960) / (best / a.batch) ))
