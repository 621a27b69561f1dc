#!/usr/bin/env python
"""Steady-state throughput of upscale_video_b200.raw_stream on synthetic 1080p frames from an in-memory file
(no ffmpeg in this image): python tools/raw_stream_bench.py [frames]"""
import io
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from upscale_video_b200 import engine as E, ncnn_model, raw_stream  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
frames = np.random.default_rng(0).integers(0, 256, (n, 1080, 1920, 3), dtype=np.uint8).tobytes()
mdir = ncnn_model.packaged_model_dir()
up = E.Engine.from_files(mdir, "2x_Compact_Pretrain", 0)
pre = E.Engine.from_files(mdir, "1" + raw_stream.HURR, 0)


class Sink:
    def __init__(self):
        self.n = 0

    def write(self, b):
        self.n += len(b)

    def flush(self):
        pass


for models, kw in (([], {"upscaler": up}), (["a"], {"upscaler": up, "prepass": pre})):
    for overlap in (False, True):  # sequential loop vs reader / engine / writer threads (--overlap)
        raw_stream.stream(io.BytesIO(frames[:8 * 1080 * 1920 * 3]), Sink(), 1920, 1080, 2, models, chunk=8, overlap=overlap, **kw)  # warm
        sink = Sink()
        t = time.time()
        k = raw_stream.stream(io.BytesIO(frames), sink, 1920, 1080, 2, models, chunk=8, overlap=overlap, **kw)
        dt = time.time() - t
        print("raw_stream 1080p->4K models=%s overlap=%s: %d frames, %.1f fps, %.2f GB/s out" % (models, overlap, k, k / dt, sink.n / dt / 1e9))
