"""GPU: HurrDeblur (nf = 24) pre-pass: persistent schedule (two CTAs per SM at 1080p: 10 x 15 = 150 CTAs) vs layer by layer."""
import sys
import numpy as np
sys.path.insert(0, ".")
import torch  # noqa: E402
from upscale_video_b200 import engine as E, ncnn_model  # noqa: E402
HURR = "1x_HurrDeblur_SubCompact_nf24-nc8_244k_net_g"
for (h, w, n) in ((1080, 1920, 8), (540, 960, 8), (70, 300, 3)):
    outs = {}
    for impl in (E.IMPL_AUTO, E.IMPL_TCGEN05):
        eng = E.Engine.from_files(ncnn_model.packaged_model_dir(), HURR, 0)
        eng.set_option(E.OPT_IMPL, impl)
        d_in = torch.from_numpy(np.random.default_rng(1).integers(0, 256, (n, h, w, 3), dtype=np.uint8)).cuda()
        d_out = torch.empty_like(d_in)
        eng.run_batch_device(d_in, d_out, n, h, w, 0, 0, sync=True)
        stream = torch.cuda.ExternalStream(eng.stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record()
        for _ in range(10):
            eng.run_batch_device(d_in, d_out, n, h, w, 0, 0, sync=False)
        with torch.cuda.stream(stream):
            e1.record()
        e1.synchronize()
        outs[impl] = d_out.cpu().numpy()
        print("%dx%d x%d impl %d: %.3f ms per frame, pipe launches %d, fallbacks %d" % (h, w, n, impl, e0.elapsed_time(e1) / 10 / n,
              int(eng.stat(E.STAT_PIPE_LAUNCHES)), int(eng.stat(E.STAT_PIPE_FALLBACKS))), flush=True)
        eng.close()
    print("   identical:", np.array_equal(outs[E.IMPL_AUTO], outs[E.IMPL_TCGEN05]))
