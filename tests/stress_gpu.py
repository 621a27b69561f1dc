#!/usr/bin/env python
"""Stress run for the pipelined schedule (by hand under gpurun; not collected by pytest).

Random frame sizes / batch sizes / ring sizes / tilings; every result of the pipelined persistent kernel must be
bit-identical to the layer-by-layer schedule (same MMAs per row, different synchronisation), which in turn is
checked against the oracle by tests/test_gpu_parity.py.  Catches rare ordering bugs in the counter protocol.

    python tests/stress_gpu.py [iterations] [seed]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from upscale_video_b200 import engine as E  # noqa: E402
from upscale_video_b200 import ncnn_model  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 60
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
mdir = ncnn_model.packaged_model_dir()
models = ["2x_Compact_Pretrain", "4x_Compact_Pretrain", "1x_HurrDeblur_SubCompact_nf24-nc8_244k_net_g"]
eng = {m: (E.Engine.from_files(mdir, m, 0), E.Engine.from_files(mdir, m, 0)) for m in models}
for m in models:
    eng[m][1].set_option(E.OPT_IMPL, E.IMPL_TCGEN05)
t0 = time.time()
bad = 0
for i in range(iters):
    m = models[int(rng.integers(0, 10)) % 3 if i % 3 else 0]
    pipe, layer = eng[m]
    s = pipe.scale
    big = rng.random() < 0.25
    h = int(rng.integers(1, 1300 if big else 200))
    w = int(rng.integers(1, 2100 if big else 700))
    n = int(rng.integers(1, 4))
    tile = int(rng.choice([960, 960, 0, 500]))
    if tile == 0 and (w + 127) // 128 * (18 if s > 1 else 10) > 148:
        w = 900
    ring = int(rng.choice([0, 4, 5, 8, 16, 33]))
    mb = int(rng.choice([0, 1, 2]))
    x = torch.from_numpy(rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)).cuda()
    ya = torch.empty((n, h * s, w * s, 3), dtype=torch.uint8, device="cuda")
    yb = torch.empty_like(ya)
    pipe.set_option(E.OPT_RING_ROWS, ring)
    pipe.set_option(E.OPT_MAX_BATCH, mb)
    layer.set_option(E.OPT_MAX_BATCH, mb)
    pipe.reset_stats()
    try:  # the persistent schedule explicitly (the narrow HurrDeblur network defaults to one launch per layer) ...
        pipe.set_option(E.OPT_IMPL, E.IMPL_PIPELINED)
        pipe.run_batch_device(x, ya, n, h, w, tile, 10, sync=True)
    except E.EngineError:  # ... unless layers x bands does not fit the device: whatever `auto` picks
        pipe.set_option(E.OPT_IMPL, E.IMPL_AUTO)
        pipe.run_batch_device(x, ya, n, h, w, tile, 10, sync=True)
    used_pipe = pipe.stat(E.STAT_PIPE_LAUNCHES) > 0
    layer.run_batch_device(x, yb, n, h, w, tile, 10, sync=True)
    same = bool(torch.equal(ya, yb))
    bad += not same
    print("%3d %-20s n=%d %4dx%-4d tile=%3d ring=%2d max_batch=%d pipe=%d %s" % (i, m[:20], n, h, w, tile, ring, mb, used_pipe,
                                                                                  "ok" if same else "MISMATCH"), flush=True)
print("stress: %d iterations, %d mismatches, %.1f s" % (iters, bad, time.time() - t0))
sys.exit(1 if bad else 0)
