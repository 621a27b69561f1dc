#!/usr/bin/env python
"""Freeze known-answer vectors for the denoise pass (`-m n=<level>`) into tests/golden/nlm_*.npz.

Unlike the network goldens these come from the REFERENCE'S OWN DEPENDENCY: the denoise worker of the reference
(upscale/upscale_processing.py:350-362) is a single call to cv2.fastNlMeansDenoisingColored, and cv2 (opencv 4.13.0,
CPU path: no OpenCL in this image, so the cv2.UMat argument changes nothing) is importable here.  Inputs are crops of
the reference's sample.png and seeded synthetic images; outputs are what cv2 returns for
``fastNlMeansDenoisingColored(cv2.UMat(x), None, level, level, 5, 9)`` exactly as apply_denoise calls it.

Usage: python tools/make_nlm_goldens.py [--ref /root/reference] [--out tests/golden]
"""
import argparse
import os

import cv2
import numpy as np


def reference_call(x, level):
    """reference upscale/upscale_processing.py:352-354."""
    return cv2.fastNlMeansDenoisingColored(cv2.UMat(x), None, level, level, 5, 9).get()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden"))
    a = ap.parse_args()
    img = cv2.imread(os.path.join(a.ref, "sample.png"))
    assert img.shape == (1278, 1920, 3)
    rng = np.random.default_rng(0)

    def save(name, x, level):
        path = os.path.join(a.out, name + ".npz")
        y = reference_call(x, level)
        np.savez_compressed(path, x=x, y=y, level=level, cv2_version=cv2.__version__)
        print("%-22s %7d B  %s level %d, %.1f%% of values changed" % (name, os.path.getsize(path), x.shape, level, 100.0 * (x != y).mean()))

    save("nlm_crop_l3", img[300:396, 800:930].copy(), 3)        # natural image, the README's typical level
    save("nlm_crop_l10", img[640:700, 1000:1075].copy(), 10)    # strong smoothing: long weight table
    noisy = np.clip(img[900:960, 200:290].astype(np.float64) + rng.normal(0, 6, (60, 90, 3)), 0, 255).astype(np.uint8)
    save("nlm_noisy_l5", noisy, 5)                              # what the filter is for
    save("nlm_noise_l30", rng.integers(0, 256, (33, 47, 3), dtype=np.uint8), 30)  # saturating noise, maximum level, ragged size
    save("nlm_tiny_l3", rng.integers(0, 256, (3, 5, 3), dtype=np.uint8), 3)       # smaller than the 6-px border: multi-bounce reflect-101
    save("nlm_1px_l1", rng.integers(0, 256, (1, 1, 3), dtype=np.uint8), 1)


if __name__ == "__main__":
    main()
