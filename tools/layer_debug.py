import sys, numpy as np
sys.path.insert(0, ".")
import torch
from upscale_video_b200 import engine as E, ncnn_model
HURR = "1x_HurrDeblur_SubCompact_nf24-nc8_244k_net_g"
eng = E.Engine.from_files(ncnn_model.packaged_model_dir(), sys.argv[1] if len(sys.argv) > 1 else HURR, 0)
eng.set_option(E.OPT_IMPL, E.IMPL_TCGEN05)
n, h, w = 8, 1080, 1920
d_in = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda")
d_out = torch.empty((n, h * eng.scale, w * eng.scale, 3), dtype=torch.uint8, device="cuda")
tile = 0 if eng.scale == 1 else 960
eng.run_batch_device(d_in, d_out, n, h, w, tile, 10, sync=True)
eng.set_option(E.OPT_PIPE_DEBUG, 1)
eng.run_batch_device(d_in, d_out, n, h, w, tile, 10, sync=True)
