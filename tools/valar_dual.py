"""Experiment: two engine contexts (two CUDA streams) on one GPU running independent 4x_Valar_v1 batches concurrently, so
that the CTAs of one context's launch fill the SMs the other context's launch is draining.

    python tools/valar_dual.py [--batch 4 --reps 3 --contexts 2]
"""
import argparse
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from upscale_video_b200 import engine as E, ncnn_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--contexts", type=int, default=2)
a = ap.parse_args()
h, w = 540, 960
engs = [E.Engine.from_files(ncnn_model.packaged_model_dir(), "4x_Valar_v1", 0) for _ in range(a.contexts)]
ins = [torch.randint(0, 256, (a.batch, h, w, 3), dtype=torch.uint8, device="cuda") for _ in engs]
outs = [torch.empty((a.batch, h * 4, w * 4, 3), dtype=torch.uint8, device="cuda") for _ in engs]
for e, i, o in zip(engs, ins, outs):
    e.run_batch_device(i, o, a.batch, h, w, sync=True)


def work(k):
    for _ in range(a.reps):
        engs[k].run_batch_device(ins[k], outs[k], a.batch, h, w, sync=True)


for n in (1, a.contexts):
    ts = [threading.Thread(target=work, args=(k,)) for k in range(n)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    print("%d context(s): %.1f ms/frame aggregate (%.0f TFLOP/s)" % (n, dt / (n * a.reps * a.batch) * 1e3, 18.73 * n * a.reps * a.batch / dt))
ref = engs[0].run_u8(ins[1][0].cpu().numpy())
assert np.array_equal(outs[1][0].cpu().numpy(), ref), "contexts disagree"
