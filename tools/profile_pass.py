#!/usr/bin/env python
"""One or more passes of N synthetic 1080p frames through the engine -- the command ncu wraps (see profiles/README.md).

    ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 20 -c 2 -o gpurun_out/prof \
        python tools/profile_pass.py --frames 4 --passes 2
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from upscale_video_b200 import engine as E  # noqa: E402
from upscale_video_b200 import ncnn_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=4)
ap.add_argument("--passes", type=int, default=2)
ap.add_argument("--model", default="2x_Compact_Pretrain")
ap.add_argument("--h", type=int, default=1080)
ap.add_argument("--w", type=int, default=1920)
ap.add_argument("--tile", type=int, default=960)
ap.add_argument("--impl", type=int, default=0)
ap.add_argument("--ring", type=int, default=0)
ap.add_argument("--debug", type=int, default=0)
a = ap.parse_args()
eng = E.Engine.from_files(ncnn_model.packaged_model_dir(), a.model, 0)
eng.set_option(E.OPT_IMPL, a.impl)
eng.set_option(E.OPT_RING_ROWS, a.ring)
eng.set_option(E.OPT_PIPE_DEBUG, a.debug)
s = eng.scale
d_in = torch.randint(0, 256, (a.frames, a.h, a.w, 3), dtype=torch.uint8, device="cuda")
d_out = torch.empty((a.frames, a.h * s, a.w * s, 3), dtype=torch.uint8, device="cuda")
for _ in range(a.passes):
    eng.run_batch_device(d_in, d_out, a.frames, a.h, a.w, a.tile, 10, sync=True)
print("done", int(eng.stat(E.STAT_LAUNCHES)), "launches")
