"""A few passes of 1080p frames through 1x_HurrDeblur (layer-by-layer tcgen05 schedule): the ncu target for its kernels."""
import sys
sys.path.insert(0, ".")
import torch
from upscale_video_b200 import engine as E, ncnn_model
eng = E.Engine.from_files(ncnn_model.packaged_model_dir(), "1x_HurrDeblur_SubCompact_nf24-nc8_244k_net_g", 0)
n = 4
d_in = torch.randint(0, 256, (n, 1080, 1920, 3), dtype=torch.uint8, device="cuda")
d_out = torch.empty_like(d_in)
for _ in range(3):
    eng.run_batch_device(d_in, d_out, n, 1080, 1920, 0, 0, sync=True)
print("launches", int(eng.stat(E.STAT_LAUNCHES)))
