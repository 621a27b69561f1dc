#!/usr/bin/env python
"""A few launches of the denoise kernel on device-resident 1080p frames -- the command profiled under ncu
(profiles/r01n_nlm_*): python tools/nlm_pass.py [frames] [level] [launches]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from upscale_video_b200 import engine as E  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
level = float(sys.argv[2]) if len(sys.argv) > 2 else 3
launches = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dn = E.Denoiser(0)
d_in = torch.randint(0, 256, (n, 1080, 1920, 3), dtype=torch.uint8, device="cuda")
d_out = torch.empty_like(d_in)
stream = torch.cuda.ExternalStream(dn.stream)
dn.run_batch_device(d_in, d_out, n, 1080, 1920, level, sync=True)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(launches + 1)]
for i in range(launches):
    with torch.cuda.stream(stream):
        ev[i].record()
    dn.run_batch_device(d_in, d_out, n, 1080, 1920, level, sync=False)
with torch.cuda.stream(stream):
    ev[launches].record()
dn.synchronize()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(launches)]
print("ok checksum %d; %d frames per launch, level %g: ms per launch %s -> %.0f frames/s" % (
    int(d_out[0, ::97, ::89].to(torch.int64).sum()), n, level, ["%.3f" % m for m in ms], n / (min(ms) * 1e-3)))

import time  # noqa: E402

h_in = d_in.cpu().pin_memory()
h_out = torch.empty_like(h_in).pin_memory()
dn.run_batch_host(h_in, h_out, n, 1080, 1920, level)
t0 = time.perf_counter()
for _ in range(launches):
    dn.run_batch_host(h_in, h_out, n, 1080, 1920, level)
dt = (time.perf_counter() - t0) / launches
print("end to end (pinned host -> device -> pinned host, chunk %s): %.3f ms per %d frames -> %.0f frames/s; equal to device result: %s" % (
    os.environ.get("B2SR_NLM_CHUNK", "default"), dt * 1e3, n, n / dt, bool(torch.equal(h_out, d_out.cpu()))))
