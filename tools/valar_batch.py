"""N passes of 4 device-resident 540p frames through 4x_Valar_v1 (the ncu target for its whole-pass DRAM traffic).

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -s 354 -c 354 --csv \
        --log-file gpurun_out/r02_valar_traffic.csv python tools/valar_batch.py --passes 2
(one pass = 1 prep + 351 convolution + 2 resize launches with one launch per convolution; --seg 1 = persistent segments)"""
import argparse
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402

from upscale_video_b200 import engine as E, ncnn_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--passes", type=int, default=2)
ap.add_argument("--n", type=int, default=4)
ap.add_argument("--seg", type=int, default=0)
ap.add_argument("--h", type=int, default=540)
a = ap.parse_args()
eng = E.Engine.from_files(ncnn_model.packaged_model_dir(), "4x_Valar_v1", 0)
eng.set_option(E.OPT_SEG_PIPE, a.seg)
d_in = torch.randint(0, 256, (a.n, a.h, 960, 3), dtype=torch.uint8, device="cuda")
d_out = torch.empty((a.n, a.h * 4, 3840, 3), dtype=torch.uint8, device="cuda")
for _ in range(a.passes):
    eng.run_batch_device(d_in, d_out, a.n, a.h, 960, sync=True)
print("launches", int(eng.stat(E.STAT_LAUNCHES)))
