"""GPU: per-launch device time (CUDA events around every launch, B2SR_OPT_PROFILE) vs elapsed time of a pass, layer-by-layer schedule."""
import sys
sys.path.insert(0, ".")
import torch
from upscale_video_b200 import engine as E, ncnn_model
HURR = "1x_HurrDeblur_SubCompact_nf24-nc8_244k_net_g"
name = sys.argv[1] if len(sys.argv) > 1 else HURR
eng = E.Engine.from_files(ncnn_model.packaged_model_dir(), name, 0)
eng.set_option(E.OPT_IMPL, E.IMPL_TCGEN05)
n, h, w = 8, 1080, 1920
tile = 0 if eng.scale == 1 else 960
d_in = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda")
d_out = torch.empty((n, h * eng.scale, w * eng.scale, 3), dtype=torch.uint8, device="cuda")
stream = torch.cuda.ExternalStream(eng.stream)
for prof in (0, 1):
    eng.set_option(E.OPT_PROFILE, prof)
    eng.run_batch_device(d_in, d_out, n, h, w, tile, 10, sync=True)
    eng.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
    for _ in range(5):
        eng.run_batch_device(d_in, d_out, n, h, w, tile, 10, sync=False)
    with torch.cuda.stream(stream):
        e1.record()
    e1.synchronize()
    print("profile %d: %.3f ms per pass of %d frames; launches per pass %.0f; summed per-launch device time per pass %.3f ms (mid convs %.3f ms over %.0f launches)" % (
        prof, e0.elapsed_time(e1) / 5, n, eng.stat(E.STAT_LAUNCHES) / 5, eng.stat(E.STAT_ALL_MS) / 5 if prof else float("nan"),
        eng.stat(E.STAT_TC_MID_MS) / 5 if prof else float("nan"), eng.stat(E.STAT_TC_MID_COUNT) / 5 if prof else float("nan")))
