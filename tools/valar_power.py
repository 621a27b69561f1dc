"""Board power / SM clock while 4x_Valar_v1 runs back to back at 540p (NVML sampling): which limit is the schedule at?

    python tools/valar_power.py [--model 4x_Valar_v1 --h 540 --w 960 --scale 4 --batch 4 --secs 5]"""
import argparse
import sys
import threading
import time

sys.path.insert(0, ".")
import pynvml  # noqa: E402
import torch  # noqa: E402

from upscale_video_b200 import engine as E, ncnn_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="4x_Valar_v1")
ap.add_argument("--h", type=int, default=540)
ap.add_argument("--w", type=int, default=960)
ap.add_argument("--scale", type=int, default=4)
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--secs", type=float, default=5.0)
ap.add_argument("--ablate", type=int, default=0, help="B2SR_OPT_ABLATE bits, switched on AFTER two real passes (the buffers keep real activations)")
a = ap.parse_args()
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
eng = E.Engine.from_files(ncnn_model.packaged_model_dir(), a.model, 0)
d_in = torch.randint(0, 256, (a.batch, a.h, a.w, 3), dtype=torch.uint8, device="cuda")
d_out = torch.empty((a.batch, a.h * a.scale, a.w * a.scale, 3), dtype=torch.uint8, device="cuda")
time.sleep(1)
print("idle power %.0f W" % (pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
for _ in range(2):
    eng.run_batch_device(d_in, d_out, a.batch, a.h, a.w, sync=True)
if a.ablate:
    eng.set_option(E.OPT_ABLATE, a.ablate)
    eng.run_batch_device(d_in, d_out, a.batch, a.h, a.w, sync=True)
samples, stop = [], [False]


def samp():
    while not stop[0]:
        samples.append((pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
        time.sleep(0.05)


th = threading.Thread(target=samp)
th.start()
t0 = time.time()
n = 0
while time.time() - t0 < a.secs:
    eng.run_batch_device(d_in, d_out, a.batch, a.h, a.w, sync=True)
    n += a.batch
dt = time.time() - t0
stop[0] = True
th.join()
s = samples[len(samples) // 3:]
pw = sum(x[0] for x in s) / len(s)
ck = sum(x[1] for x in s) / len(s)
reasons = 0
for x in s:
    reasons |= x[2]
print("ablate %d | %s %dx%d batch %d: %.1f fps  power %.0f W  sm clock %.0f MHz  throttle reasons 0x%x  => %.2f J/frame" % (
    a.ablate, a.model, a.h, a.w, a.batch, n / dt, pw, ck, reasons, pw * dt / n))
