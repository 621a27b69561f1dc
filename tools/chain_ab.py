import sys, os, json
sys.path.insert(0, ".")
import torch
import bench
from upscale_video_b200 import engine as E, ncnn_model
for n, b in ((8, 1), (2, 2), (4, 2), (16, 2), (2, 1), (8, 2), (2, 2), (8, 1)):
    os.environ["B2SR_BENCH_CHAIN_N"] = str(n); os.environ["B2SR_BENCH_CHAIN_BUFS"] = str(b)
    out = bench.side_configs(E, ncnn_model, torch, 0, steps=10)
    print(n, b, [(round(o["frames_per_s"], 1), round(o["hurrdeblur_alone_frames_per_s"])) for o in out if "chained" in o["workload"]], flush=True)
