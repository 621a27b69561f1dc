#!/usr/bin/env python
"""End-to-end (pinned host -> device -> pinned host) frames/s of Engine.run_batch_host for several chunk sizes:
python tools/e2e_sweep.py [frames] [reps]   (2x_Compact_Pretrain, 1080p)"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from upscale_video_b200 import engine as E  # noqa: E402
from upscale_video_b200 import ncnn_model  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
eng = E.Engine.from_files(ncnn_model.packaged_model_dir(), "2x_Compact_Pretrain", 0)
h_in = torch.randint(0, 256, (n, 1080, 1920, 3), dtype=torch.uint8).pin_memory()
h_out = torch.empty((n, 2160, 3840, 3), dtype=torch.uint8).pin_memory()
d_in = h_in.cuda()
d_out = torch.empty((n, 2160, 3840, 3), dtype=torch.uint8, device="cuda")
for _ in range(3):
    eng.run_batch_device(d_in, d_out, n, 1080, 1920, sync=True)
t0 = time.perf_counter()
for _ in range(reps):
    eng.run_batch_device(d_in, d_out, n, 1080, 1920, sync=True)
print("device-resident: %.1f frames/s" % (n * reps / (time.perf_counter() - t0)))
for b in (0, 2, 3, 4, 5, 7, 14):
    eng.set_option(E.OPT_MAX_BATCH, b)
    eng.run_batch_host(h_in, h_out, n, 1080, 1920)
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.run_batch_host(h_in, h_out, n, 1080, 1920)
    print("max_batch %2d: %.1f frames/s end to end" % (b, n * reps / (time.perf_counter() - t0)))
