#!/usr/bin/env python
"""Convert the reference's ncnn model pairs into this repo's ``.b2sr`` containers.

Usage: python tools/convert_models.py [--src /root/reference/models] [--dst upscale_video_b200/models] [names...]

The ncnn files are third-party weights that the reference ships under ``models/`` (reference
``upscale_processing.py:70-71`` loads ``<scale><model_file>.param/.bin``).  The GPU test box has no
``/root/reference``, so the converted containers are what tests, ``smoke()`` and ``bench.py`` load there.
The conversion is lossless: array element order is unchanged and fp32 weights are narrowed to fp16 only when
every value survives the round trip (true for 4x_Compact_Pretrain).
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from upscale_video_b200 import ncnn_model  # noqa: E402

DEFAULT = ["2x_Compact_Pretrain", "4x_Compact_Pretrain", "1x_HurrDeblur_SubCompact_nf24-nc8_244k_net_g", "4x_Valar_v1"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference/models")
    ap.add_argument("--dst", default=ncnn_model.packaged_model_dir())
    ap.add_argument("names", nargs="*", default=DEFAULT)
    a = ap.parse_args()
    os.makedirs(a.dst, exist_ok=True)
    for name in a.names:
        g = ncnn_model.load_ncnn(os.path.join(a.src, name + ".param"), os.path.join(a.src, name + ".bin"))
        out = os.path.join(a.dst, name + ".b2sr")
        ncnn_model.save_b2sr(g, out)
        back = ncnn_model.load_b2sr(out)
        for l0, l1 in zip(g.layers, back.layers):
            assert (l0.type, l0.name, l0.bottoms, l0.tops, l0.params) == (l1.type, l1.name, l1.bottoms, l1.tops, l1.params)
            for k in l0.weights:
                assert (l0.weights[k].astype("f4") == l1.weights[k].astype("f4")).all(), (name, l0.name, k)
        d = ncnn_model.compact_desc(g)
        print("%-50s layers=%4d params=%9d B -> %s (%d B) family=%s" % (
            name, len(g.layers), g.nbytes(), out, os.path.getsize(out), d))


if __name__ == "__main__":
    main()
