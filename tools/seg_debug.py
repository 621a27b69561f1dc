"""GPU: buffers after the first RRDB (ops 0..15) with persistent segments vs per-convolution launches."""
import sys
import numpy as np
sys.path.insert(0, ".")
from upscale_video_b200 import engine as E, ncnn_model  # noqa: E402
h, w = int(sys.argv[1]), int(sys.argv[2])
upto = int(sys.argv[3]) if len(sys.argv) > 3 else 15
rr = int(sys.argv[4]) if len(sys.argv) > 4 else 0
eng = E.Engine.from_files(ncnn_model.packaged_model_dir(), "4x_Valar_v1", 0)
if rr:
    eng.set_option(E.OPT_RING_ROWS, rr)
img = np.random.default_rng(0).integers(0, 256, (h, w, 3), dtype=np.uint8)
o = eng.program.ops[upto]
for buf, name, sl in ((o["out16_buf"], "fp16 out", slice(o["out16_off"], o["out16_off"] + o["cout"])), (o["out32_buf"], "fp32 out", slice(0, 64))):
    if buf < 0:
        continue
    res = {}
    for mode in (1, 0):
        eng.set_option(E.OPT_SEG_PIPE, mode)
        res[mode] = eng.debug_fused(img, upto, buf)[:, :, sl]
    d = np.abs(res[1] - res[0])
    print("%s (buffer %d): max |diff| %.4g, scale %.3g" % (name, buf, d.max(), np.abs(res[0]).max()))
    bad = d.max(axis=2) > 0
    print("   bad pixels per row:", {int(y): int(bad[y].sum()) for y in range(h) if bad[y].any()})
    badc = d.max(axis=(0, 1)) > 0
    print("   bad channels:", np.nonzero(badc)[0].tolist())
    ys, xs = np.nonzero(bad)
    if len(ys):
        print("   cols of bad pixels in first bad row:", xs[ys == ys.min()][:40].tolist())
