"""CPU oracle for the frame-upscale path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import this module; product code under ``upscale_video_b200/`` must never do so.

PARITY UNPINNED for ncnn's own arithmetic (see oracle.c header): the reference delegates it to the un-vendored, unpinned
``ncnn_vulkan`` wheel and ships no golden outputs.  PINNED for the rest: the glue below reproduces, bit for bit, files written by the
reference's own ``upscale_processing.py`` run with an ncnn stand-in (``tools/make_ref_glue_goldens.py`` ->
``tests/golden/ref_glue.npz``), and the graph interpreter equals the published PyTorch architectures the models were exported from
(``tests/test_architecture_cross_check.py``).  This file restates

* the ncnn model format and layer semantics for the layers the reference's models use
  (own, independent reader below -- it deliberately does not share code with
  ``upscale_video_b200/ncnn_model.py``), and
* the reference's own Python glue around the network, line by line:
  ``apply_model``   reference upscale/upscale_processing.py:258-299
  ``process_tile``  reference upscale/upscale_processing.py:395-477
  ``upscale_image`` reference upscale/upscale_processing.py:480-542
  (pixel order fed unswapped as ``PIXEL_BGR`` :265-270/:437-442, ``x * float32(1/255.0)`` :271-273/:443-445,
  ``* 255`` :284/:462, float64 canvas :497, ``cv2.imwrite`` rounding :288/:519).

Heavy loops (convolution, pixel shuffle, nearest resize) run in ``liboracle.so`` (oracle.c, built by
``oracle/Makefile``); everything else is numpy.  ``precision="f64"`` is the value goldens are frozen from,
``precision="f32"`` mirrors ncnn's CPU arithmetic and is what the CPU baseline times.
"""
from __future__ import annotations

import ctypes
import json
import math
import os
import struct
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        assert _LIB.oracle_abi_version() == 1
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# ------------------------------------------------------------------------------------------------
# independent model reader (ncnn .param/.bin and this repo's .b2sr container)
# ------------------------------------------------------------------------------------------------
def read_ncnn(param_path, bin_path):
    """-> list of dict(type, name, bottoms, tops, params{int: value|list}, arrays{str: ndarray})."""
    toks = [ln.split() for ln in open(param_path).read().splitlines() if ln.strip()]
    assert toks[0] == ["7767517"], "bad ncnn magic"
    blob = open(bin_path, "rb").read()
    pos, layers = 0, []
    for t in toks[2:]:
        nin, nout = int(t[2]), int(t[3])
        L = {"type": t[0], "name": t[1], "bottoms": t[4:4 + nin], "tops": t[4 + nin:4 + nin + nout],
             "params": {}, "arrays": {}}
        for kv in t[4 + nin + nout:]:
            k, v = kv.split("=")
            k = int(k)
            if k <= -23300:
                vals = v.split(",")
                L["params"][-23300 - k] = [float(x) for x in vals[1:1 + int(vals[0])]]
            else:
                L["params"][k] = float(v) if ("." in v or "e" in v) else int(v)
        if L["type"] == "Convolution":
            n = L["params"][6]
            tag = struct.unpack_from("<I", blob, pos)[0]
            pos += 4
            if tag == 0x01306B47:
                w = np.frombuffer(blob, "<f2", n, pos).astype(np.float32)
                pos += (2 * n + 3) // 4 * 4
            else:
                assert tag == 0, hex(tag)
                w = np.frombuffer(blob, "<f4", n, pos).copy()
                pos += 4 * n
            co, k = L["params"][0], L["params"][1]
            L["arrays"]["weight"] = w.reshape(co, n // (co * k * k), k, k)
            if L["params"].get(5, 0):
                L["arrays"]["bias"] = np.frombuffer(blob, "<f4", co, pos).copy()
                pos += 4 * co
        elif L["type"] == "PReLU":
            n = L["params"][0]
            L["arrays"]["slope"] = np.frombuffer(blob, "<f4", n, pos).copy()
            pos += 4 * n
        layers.append(L)
    assert pos == len(blob), "ncnn .bin not fully consumed (%d of %d)" % (pos, len(blob))
    return layers


def read_b2sr(path):
    data = open(path, "rb").read()
    assert data[:8] == b"B2SRv1\0\0"
    hlen = struct.unpack_from("<I", data, 8)[0]
    head = json.loads(data[12:12 + hlen])
    layers = []
    for e in head["layers"]:
        L = {"type": e["type"], "name": e["name"], "bottoms": e["bottoms"], "tops": e["tops"],
             "params": {int(k): v for k, v in e["params"].items()}, "arrays": {}}
        for key, a in e["arrays"].items():
            arr = np.frombuffer(data, "<" + {"float16": "f2", "float32": "f4"}[a["dtype"]],
                                int(np.prod(a["shape"])), 12 + hlen + a["offset"])
            L["arrays"][key] = arr.astype(np.float32).reshape(a["shape"])
        layers.append(L)
    return layers


def read_model(model_dir, stem):
    p, b = os.path.join(model_dir, stem + ".param"), os.path.join(model_dir, stem + ".bin")
    if os.path.exists(p) and os.path.exists(b):
        return read_ncnn(p, b)
    return read_b2sr(os.path.join(model_dir, stem + ".b2sr"))


# ------------------------------------------------------------------------------------------------
# ncnn layer semantics (HWC arrays)
# ------------------------------------------------------------------------------------------------
def conv(x, weight, bias, pad, act=0, slope=None, precision="f64"):
    real = np.float64 if precision == "f64" else np.float32
    x = np.ascontiguousarray(x, real)
    H, W, Cin = x.shape
    Cout, Cin2, k, _ = weight.shape
    assert Cin == Cin2, (Cin, Cin2)
    out = np.empty((H, W, Cout), real)
    w = np.ascontiguousarray(weight, np.float32)
    b = None if bias is None else np.ascontiguousarray(bias, np.float32)
    s = None if slope is None else np.ascontiguousarray(slope, np.float32)
    fn = getattr(lib(), "oracle_conv_" + precision)
    rc = fn(_ptr(x), H, W, Cin, _ptr(w), None if b is None else _ptr(b), Cout, k, pad, act,
            None if s is None else _ptr(s), _ptr(out))
    assert rc == 0, rc
    return out


def pixelshuffle(x, r, precision="f64"):
    if r == 1:
        return x
    H, W, C = x.shape
    out = np.empty((H * r, W * r, C // (r * r)), x.dtype)
    getattr(lib(), "oracle_pixelshuffle_" + precision)(_ptr(np.ascontiguousarray(x)), H, W, C // (r * r), r, _ptr(out))
    return out


def nearest(x, sy, sx, precision="f64"):
    H, W, C = x.shape
    OH, OW = int(H * sy), int(W * sx)
    out = np.empty((OH, OW, C), x.dtype)
    getattr(lib(), "oracle_nearest_" + precision)(_ptr(np.ascontiguousarray(x)), H, W, C, ctypes.c_float(sy),
                                                   ctypes.c_float(sx), OH, OW, _ptr(out))
    return out


def run_graph(layers, x, precision="f64", input_name="input", output_name="output", taps=None):
    """Interpret the ncnn graph on an HWC input.  ``taps``: optional dict filled with blob name -> array for
    every Convolution/PReLU output (layer-wise bring-up checks)."""
    real = np.float64 if precision == "f64" else np.float32
    blobs = {}
    i, n = 0, len(layers)
    last_use = {}  # blob -> index of the last layer that reads it: freed right after (a 540p RRDB frame has 2127 blobs)
    for j, L in enumerate(layers):
        for b in L["bottoms"]:
            last_use[b] = j
    while i < n:
        L = layers[i]
        t, P = L["type"], L["params"]
        if t == "Input":
            assert L["tops"] == [input_name], (L["tops"], input_name)
            blobs[L["tops"][0]] = np.ascontiguousarray(x, real)
        elif t == "Split":
            for top in L["tops"]:
                blobs[top] = blobs[L["bottoms"][0]]
        elif t == "Convolution":
            assert P.get(2, 1) == 1 and P.get(3, 1) == 1, "dilation/stride != 1"
            act, slope = P.get(9, 0), None
            if act == 2:
                slope = np.asarray(P[10], np.float32)
            elif act != 0:
                raise NotImplementedError("conv activation %r" % act)
            blobs[L["tops"][0]] = conv(blobs[L["bottoms"][0]], L["arrays"]["weight"], L["arrays"].get("bias"),
                                       P.get(4, 0), act, slope, precision)
            if taps is not None:
                taps[L["tops"][0]] = blobs[L["tops"][0]]
        elif t == "PReLU":
            v = blobs[L["bottoms"][0]]
            s = L["arrays"]["slope"].astype(real)
            blobs[L["tops"][0]] = np.where(v < 0, v * s, v)
            if taps is not None:
                taps[L["tops"][0]] = blobs[L["tops"][0]]
        elif t == "PixelShuffle":
            assert P.get(1, 0) == 0
            blobs[L["tops"][0]] = pixelshuffle(blobs[L["bottoms"][0]], P.get(0, 1), precision)
        elif t == "Interp":
            assert P.get(0, 0) == 1, "only nearest"
            sy, sx = float(P.get(1, 1.0)), float(P.get(2, 1.0))
            v = blobs[L["bottoms"][0]]
            blobs[L["tops"][0]] = v if (sy == 1.0 and sx == 1.0) else nearest(v, sy, sx, precision)
        elif t == "BinaryOp":
            assert P.get(0, 0) == 0 and len(L["bottoms"]) == 2, "only add"
            blobs[L["tops"][0]] = blobs[L["bottoms"][0]] + blobs[L["bottoms"][1]]
        elif t == "Eltwise":
            assert P.get(0, 0) == 1, "only sum"
            co = P.get(1, [1.0] * len(L["bottoms"]))
            # ncnn Eltwise SUM with coeffs: out = b0*c0 + b1*c1, then out += bi*ci
            acc = blobs[L["bottoms"][0]] * real(np.float32(co[0])) + blobs[L["bottoms"][1]] * real(np.float32(co[1]))
            for j in range(2, len(L["bottoms"])):
                acc = acc + blobs[L["bottoms"][j]] * real(np.float32(co[j]))
            blobs[L["tops"][0]] = acc
        elif t == "Concat":
            assert P.get(0, 0) == 0, "channel concat only"
            blobs[L["tops"][0]] = np.concatenate([blobs[b] for b in L["bottoms"]], axis=2)
        else:
            raise NotImplementedError(t)
        for b in L["bottoms"]:
            if last_use.get(b) == i and b != output_name:
                blobs.pop(b, None)
        i += 1
    return blobs[output_name]


# ------------------------------------------------------------------------------------------------
# the reference's glue, restated
# ------------------------------------------------------------------------------------------------
def from_pixels_normalize(img_u8):
    """ncnn.Mat.from_pixels(PIXEL_BGR) + substract_mean_normalize([], [1/255.0]*3)
    (reference upscale_processing.py:265-273, :437-445): no channel swap, float32 multiply."""
    return img_u8.astype(np.float32) * np.float32(1 / 255.0)


def saturate_u8(a):
    """cv2.imwrite of a floating image: rint (ties to even) then clamp (reference :288, :519)."""
    a = np.ascontiguousarray(a, np.float64)
    out = np.empty(a.shape, np.uint8)
    lib().oracle_saturate_u8_f64(_ptr(a), ctypes.c_size_t(a.size), _ptr(out))
    return out


def net_x255(layers, img_u8, precision="f64"):
    """What ``output_tile`` / ``output`` hold after ``* 255`` (reference :284, :462): float, unrounded."""
    y = run_graph(layers, from_pixels_normalize(img_u8), precision)
    if precision == "f32":
        return (y.astype(np.float32) * np.float32(255)).astype(np.float32)  # float32 array * python int
    return y * 255.0


def apply_model_array(layers, img_u8, precision="f64"):
    """reference apply_model :258-299 on an in-memory frame: untiled run, u8 via imwrite rounding."""
    return saturate_u8(net_x255(layers, img_u8, precision))


def tile_rects(height, width, tile_size=960, halo=10):
    """The tile geometry of reference process_tile :398-434 for every (y, x) of upscale_image :499-503.
    Yields (y, x, in_y0, in_y1, in_x0, in_x1, core_y0, core_y1, core_x0, core_x1)."""
    for y in range(math.ceil(height / tile_size)):
        for x in range(math.ceil(width / tile_size)):
            sy, ey = y * tile_size, min(y * tile_size + tile_size, height)
            sx, ex = x * tile_size, min(x * tile_size + tile_size, width)
            by0 = -halo if sy >= halo else 0
            by1 = halo if ey <= height - halo else 0
            bx0 = -halo if sx >= halo else 0
            bx1 = halo if ex <= width - halo else 0
            yield (y, x, sy + by0, ey + by1, sx + bx0, ex + bx1, sy, ey, sx, ex)


def process_tile(layers, img, tile_size, scale, y, x, height, width, output, precision="f64", halo=10):
    """reference process_tile :395-477 (the ncnn extractor run replaced by run_graph)."""
    for (ty, tx, iy0, iy1, ix0, ix1, cy0, cy1, cx0, cx1) in tile_rects(height, width, tile_size, halo):
        if (ty, tx) != (y, x):
            continue
        tile = img[iy0:iy1, ix0:ix1, :].copy()
        out_tile = net_x255(layers, tile, precision)
        by, bx = (cy0 - iy0) * scale, (cx0 - ix0) * scale
        output[cy0 * scale:cy1 * scale, cx0 * scale:cx1 * scale, :] = out_tile[
            by:by + (cy1 - cy0) * scale, bx:bx + (cx1 - cx0) * scale, :]


def upscale_canvas(layers, img_u8, scale, precision="f64", tile_size=960, halo=10):
    """reference upscale_image :487-516: float64 canvas assembled from tiles (before imwrite)."""
    height, width, ch = img_u8.shape
    output = np.zeros((height * scale, width * scale, ch))
    for y in range(math.ceil(height / tile_size)):
        for x in range(math.ceil(width / tile_size)):
            process_tile(layers, img_u8, tile_size, scale, y, x, height, width, output, precision, halo)
    return output


def upscale_image_array(layers, img_u8, scale, precision="f64", tile_size=960, halo=10):
    """reference upscale_image :480-519 on an in-memory frame -> the u8 image cv2.imwrite would store."""
    return saturate_u8(upscale_canvas(layers, img_u8, scale, precision, tile_size, halo))
