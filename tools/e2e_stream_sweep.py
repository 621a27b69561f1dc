"""GPU: end-to-end rate of the streaming host call (b2sr_submit_batch_host / b2sr_wait_batch, two steps in flight) against the
chunk size (B2SR_OPT_MAX_BATCH) and the synchronous call, 16 x 1080p frames per step, 2x_Compact.

    python tools/e2e_stream_sweep.py [--steps 12]"""
import argparse
import sys
import time

sys.path.insert(0, ".")
import torch  # noqa: E402

from upscale_video_b200 import engine as E, ncnn_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=12)
a = ap.parse_args()
B, H, W = 16, 1080, 1920
eng = E.Engine.from_files(ncnn_model.packaged_model_dir(), "2x_Compact_Pretrain", 0)
pairs = [(torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8).pin_memory(), torch.empty((B, 2 * H, 2 * W, 3), dtype=torch.uint8).pin_memory())
         for _ in range(2)]
d_in, d_out = pairs[0][0].cuda(), torch.empty((B, 2 * H, 2 * W, 3), dtype=torch.uint8, device="cuda")


def stream(steps):
    t = [None, None]
    for k in range(steps + 1):
        if k < steps:
            t[k & 1] = eng.submit_batch_host(pairs[k & 1][0], pairs[k & 1][1], B, H, W)
        if k >= 1:
            eng.wait_batch(t[(k - 1) & 1])


for _ in range(3):
    eng.run_batch_device(d_in, d_out, B, H, W, sync=True)
t0 = time.perf_counter()
for _ in range(a.steps):
    eng.run_batch_device(d_in, d_out, B, H, W, sync=False)
eng.synchronize()
print("device resident: %.1f fps" % (B * a.steps / (time.perf_counter() - t0)))
for mb in (4, 2, 8, 16, 4):
    eng.set_option(E.OPT_MAX_BATCH, mb)
    stream(2)
    t0 = time.perf_counter()
    stream(a.steps)
    fs = B * a.steps / (time.perf_counter() - t0)
    eng.run_batch_host(pairs[0][0], pairs[0][1], B, H, W)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        eng.run_batch_host(pairs[0][0], pairs[0][1], B, H, W)
    print("chunk %2d: streaming %.1f fps, synchronous calls %.1f fps" % (mb, fs, B * a.steps / (time.perf_counter() - t0)), flush=True)
