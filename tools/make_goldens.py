#!/usr/bin/env python
"""Freeze known-answer vectors for the frame-upscale path into tests/golden/*.npz.

The reference has no golden outputs, and /root/reference (models, sample.png) does not
exist on the GPU box, so this script -- run once in the CPU container -- stores small *input* crops of the
reference's only fixture (sample.png) and seeded noise together with the f64-oracle outputs computed from the
reference's ORIGINAL ncnn model files.  tests/ then check (a) the oracle against these vectors (regression +
.b2sr conversion), (b) the CUDA path against them (<= 1 LSB).

Usage: python tools/make_goldens.py [--ref /root/reference] [--out tests/golden]
"""
import argparse
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import oracle  # noqa: E402

HURR = "1x_HurrDeblur_SubCompact_nf24-nc8_244k_net_g"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden"))
    ap.add_argument("--valar", action="store_true")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    mdir = os.path.join(a.ref, "models")
    img = cv2.imread(os.path.join(a.ref, "sample.png"))  # BGR u8, as the reference reads it (:263, :487)
    assert img.shape == (1278, 1920, 3)
    rng = np.random.default_rng(0)
    models = {n: oracle.read_ncnn(os.path.join(mdir, n + ".param"), os.path.join(mdir, n + ".bin"))
              for n in ["2x_Compact_Pretrain", "4x_Compact_Pretrain", HURR]}

    def save(name, **kw):
        path = os.path.join(a.out, name + ".npz")
        np.savez_compressed(path, **kw)
        print("%-28s %8d B  %s" % (name, os.path.getsize(path), {k: v.shape for k, v in kw.items() if hasattr(v, "shape")}))

    # 1. natural-image crop, single tile
    x = img[300:364, 800:896].copy()
    save("compact2x_crop", model="2x_Compact_Pretrain", scale=2, x=x,
         y=oracle.upscale_image_array(models["2x_Compact_Pretrain"], x, 2, "f64"),
         y_f32=oracle.upscale_image_array(models["2x_Compact_Pretrain"], x, 2, "f32"))
    # 2. crops that straddle the reference's 960-px tile seams (reference :489, :409-427)
    x = img[600:624, 0:1000].copy()
    save("compact2x_seam_x", model="2x_Compact_Pretrain", scale=2, x=x,
         y=oracle.upscale_image_array(models["2x_Compact_Pretrain"], x, 2, "f64"))
    x = img[0:1000, 1000:1024].copy()
    save("compact2x_seam_y", model="2x_Compact_Pretrain", scale=2, x=x,
         y=oracle.upscale_image_array(models["2x_Compact_Pretrain"], x, 2, "f64"))
    # 3. 4x pixel-shuffle model
    x = img[700:748, 400:464].copy()
    save("compact4x_crop", model="4x_Compact_Pretrain", scale=4, x=x,
         y=oracle.upscale_image_array(models["4x_Compact_Pretrain"], x, 4, "f64"))
    # 4. 1x pre-pass (apply_model, untiled) and the chained config: u8 hop between the nets (:288 -> :487)
    x = img[500:564, 1200:1296].copy()
    y1 = oracle.apply_model_array(models[HURR], x, "f64")
    save("hurr1x_crop", model=HURR, scale=1, x=x, y=y1)
    save("chain_hurr_compact2x", model=HURR + "+2x_Compact_Pretrain", scale=2, x=x,
         y=oracle.upscale_image_array(models["2x_Compact_Pretrain"], y1, 2, "f64"))
    # 5. seeded uniform noise (saturates outputs; exercises clamping) incl. ragged sizes
    for name, (h, w) in {"noise_a": (40, 56), "noise_ragged": (37, 131)}.items():
        x = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        save("compact2x_" + name, model="2x_Compact_Pretrain", scale=2, x=x,
             y=oracle.upscale_image_array(models["2x_Compact_Pretrain"], x, 2, "f64"))
    x = rng.integers(0, 256, (33, 47, 3), dtype=np.uint8)
    save("hurr1x_noise", model=HURR, scale=1, x=x, y=oracle.apply_model_array(models[HURR], x, "f64"))
    # 6. unrounded float canvas for process_tile parity (reference :462-477)
    x = img[900:940, 100:164].copy()
    save("compact2x_canvas_f64", model="2x_Compact_Pretrain", scale=2, x=x,
         y=oracle.upscale_canvas(models["2x_Compact_Pretrain"], x, 2, "f64").astype(np.float32))
    if a.valar:
        v = oracle.read_ncnn(os.path.join(mdir, "4x_Valar_v1.param"), os.path.join(mdir, "4x_Valar_v1.bin"))
        x = img[640:672, 960:992].copy()
        save("valar4x_crop", model="4x_Valar_v1", scale=4, x=x, y=oracle.upscale_image_array(v, x, 4, "f64"))


if __name__ == "__main__":
    main()
