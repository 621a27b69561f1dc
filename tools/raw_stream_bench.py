#!/usr/bin/env python
"""Steady-state throughput of upscale_video_b200.raw_stream on synthetic 1080p bgr24 frames -> 4K into a sink (no ffmpeg in this
image): the end-to-end product path without the PNG hop, on one or several GPUs.

    python tools/raw_stream_bench.py [--frames 192] [--gpus 0] [--input file|memory] [--chain]

--gpus 0       : raw_stream.stream on one GPU (reader / engine / writer threads)
--gpus 0,1,... : raw_stream.stream_multi, one worker thread per entry over one dynamic chunk queue, in-order writer
--input file   : frames in a file on /dev/shm (seekable: stream_multi's workers read their chunks themselves with preadv)
--input memory : an in-memory pipe-like object (one reader thread)
Output goes to /dev/null (the bytes are produced and handed to write(); nothing is encoded)."""
import argparse
import io
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from upscale_video_b200 import raw_stream  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=960)
ap.add_argument("--gpus", default="0")
ap.add_argument("--input", default="file", choices=["file", "memory"])
ap.add_argument("--chain", action="store_true", help="-m a: HurrDeblur pre-pass in front of the upscaler")
ap.add_argument("--chunk", type=int, default=4)
ap.add_argument("--multi", action="store_true", help="one GPU through stream_multi (streaming workers) instead of stream")
ap.add_argument("--check", type=int, default=0, help="also verify the first N output frames against the single-GPU stream")
a = ap.parse_args()
gpus = [int(g) for g in a.gpus.split(",")]
H, W = 1080, 1920
rng = np.random.default_rng(0)
base = rng.integers(0, 256, (16, H, W, 3), dtype=np.uint8)
path = "/dev/shm/b2sr_raw_stream_bench.raw"
if a.input == "file":
    with open(path, "wb") as f:
        for i in range(a.frames):
            f.write(base[i % 16].data)


class Pipe(io.RawIOBase):  # pipe-like: sequential, not seekable
    def __init__(self, n):
        self.n, self.i, self.off = n, 0, 0

    def readable(self):
        return True

    def readinto(self, b):
        if self.i >= self.n:
            return 0
        src = memoryview(base[self.i % 16]).cast("B")
        k = min(len(b), len(src) - self.off)
        b[:k] = src[self.off:self.off + k]
        self.off += k
        if self.off == len(src):
            self.off, self.i = 0, self.i + 1
        return k


from upscale_video_b200 import engine as E, ncnn_model  # noqa: E402

mdir = ncnn_model.packaged_model_dir()
# one set of engines per worker, built once (a long-running stream creates them once too) and handed to every run
prebuilt = [(None, E.Engine.from_files(mdir, "1" + raw_stream.HURR, g) if a.chain else None, E.Engine.from_files(mdir, "2x_Compact_Pretrain", g))
            for g in gpus]


def run(n_frames, sink):
    fin = open(path, "rb") if a.input == "file" else Pipe(n_frames)
    models = ["a"] if a.chain else []
    pool = list(prebuilt)
    t = time.time()
    if len(gpus) == 1 and not a.multi:
        k = raw_stream.stream(fin, sink, W, H, 2, models, gpus[0], 8, max_frames=n_frames, overlap=True, upscaler=prebuilt[0][2],
                              prepass=prebuilt[0][1])
    else:
        import threading
        lock = threading.Lock()

        def take(g):
            with lock:
                i = next(j for j, e in enumerate(pool) if e[2].device == g)
                return pool.pop(i)
        k = raw_stream.stream_multi(fin, sink, W, H, 2, models, gpus, a.chunk, max_frames=n_frames, make_engines=take)
    return k, time.time() - t


sink = open("/dev/null", "wb")
run(min(a.frames, 8 * len(gpus)), sink)  # warm-up
k, dt = run(a.frames, sink)
res = {"what": "raw_stream 1080p bgr24 -> 4K, 2x_Compact_Pretrain%s" % (" after 1x_HurrDeblur (-m a)" if a.chain else ""), "gpus": gpus,
       "input": a.input, "frames": k, "seconds": dt, "frames_per_s": k / dt,
       "note": "engines prebuilt; the time includes allocating the pinned staging chunks of the call, reading the input and handing every output byte to write()"}
if a.check:
    one, many = io.BytesIO(), io.BytesIO()
    raw_stream.stream(io.BytesIO(base[:a.check].tobytes()), one, W, H, 2, ["a"] if a.chain else [], gpus[0], 4)
    raw_stream.stream_multi(io.BytesIO(base[:a.check].tobytes()), many, W, H, 2, ["a"] if a.chain else [], gpus, 2)
    res["identical_to_single_gpu_stream"] = one.getvalue() == many.getvalue()
print(json.dumps(res))
if a.input == "file":
    os.remove(path)
