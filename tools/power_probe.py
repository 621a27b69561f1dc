"""Board power / SM clock / energy per frame of the pipelined kernel on 16 x 1080p frames (NVML sampling).
The ablation numbers quoted in DESIGN.md section 4 (MMAs off, epilogue off) came from temporary kernel switches that
are not in the tree; this script reproduces the full-kernel line."""
import sys, time, threading, subprocess
sys.path.insert(0, ".")
import torch
from upscale_video_b200 import engine as E, ncnn_model
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
eng = E.Engine.from_files(ncnn_model.packaged_model_dir(), "2x_Compact_Pretrain", 0)
d_in = torch.randint(0, 256, (16, 1080, 1920, 3), dtype=torch.uint8, device="cuda")
d_out = torch.empty((16, 2160, 3840, 3), dtype=torch.uint8, device="cuda")
def run(mode, secs=4.0):
    eng.set_option(E.OPT_PIPE_DEBUG, mode)
    for _ in range(3): eng.run_batch_device(d_in, d_out, 16, 1080, 1920, sync=True)
    samples = []; stop = [False]
    def samp():
        while not stop[0]:
            samples.append((pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
            time.sleep(0.05)
    th = threading.Thread(target=samp); th.start()
    t0 = time.time(); n = 0
    while time.time() - t0 < secs:
        eng.run_batch_device(d_in, d_out, 16, 1080, 1920, sync=False); n += 16
        if n % 64 == 0: eng.synchronize()
    eng.synchronize(); dt = time.time() - t0
    stop[0] = True; th.join()
    s = samples[len(samples)//3:]
    pw = sum(a for a, b in s) / len(s); ck = sum(b for a, b in s) / len(s)
    print("mode %d: %.1f fps  power %.0f W  sm clock %.0f MHz  => %.2f J/frame" % (mode, n / dt, pw, ck, pw * dt / n))
time.sleep(1)
print("idle power %.0f W" % (pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
run(0)
