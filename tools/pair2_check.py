"""GPU: CTA-pair (cta_group::2) form of the RRDB convolutions vs one CTA per band: equality report + timing.

    python tools/pair2_check.py [--h 540 --w 960 --n 4]"""
import argparse
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402

from upscale_video_b200 import engine as E, ncnn_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--h", type=int, default=540)
ap.add_argument("--w", type=int, default=960)
ap.add_argument("--n", type=int, default=4)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
eng = E.Engine.from_files(ncnn_model.packaged_model_dir(), "4x_Valar_v1", 0)
rng = np.random.default_rng(0)
frames = torch.from_numpy(rng.integers(0, 256, (a.n, a.h, a.w, 3), dtype=np.uint8)).cuda()
outs = {}
for mode in (1, 0):
    eng.set_option(E.OPT_PAIR2, mode)
    out = torch.zeros((a.n, a.h * 4, a.w * 4, 3), dtype=torch.uint8, device="cuda")
    eng.run_batch_device(frames, out, a.n, a.h, a.w, sync=True)
    best = 1e9
    for _ in range(a.reps):
        t0 = time.perf_counter()
        eng.run_batch_device(frames, out, a.n, a.h, a.w, sync=True)
        best = min(best, time.perf_counter() - t0)
    outs[mode] = out.cpu().numpy()
    print("mode %s: %.2f ms/frame" % ("cta pairs" if mode else "one CTA per band", best / a.n * 1e3), flush=True)
for f in range(a.n):
    d = np.abs(outs[1][f].astype(int) - outs[0][f].astype(int))
    bad = np.argwhere(d.max(axis=2) > 0)
    if len(bad) == 0:
        print("frame %d: identical" % f)
        continue
    ys, xs = bad[:, 0] // 4, bad[:, 1] // 4
    print("frame %d: max |diff| %d, %.2f%% of values; LR rows %d..%d, LR cols %d..%d; bad rows per 128-col band: %s" % (
        f, d.max(), 100 * (d > 0).mean(), ys.min(), ys.max(), xs.min(), xs.max(),
        {int(b): int(len(np.unique(ys[xs // 128 == b]))) for b in np.unique(xs // 128)}))
