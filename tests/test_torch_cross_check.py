"""CPU: a second, independent restatement of the reference's graphs (SURVEY.md section 8c recommends two oracles).

The oracle (oracle/oracle.py + oracle.c: own model reader, direct convolution loops in C) is checked here against
torch's ATen operators in float64 (conv2d, prelu, leaky_relu, pixel_shuffle, nearest interpolate) interpreting the
graph as parsed by the PRODUCT's reader (upscale_video_b200/ncnn_model.py) -- two readers, two convolution
implementations, one answer.  It does not pin the oracle against ncnn itself (not installable);
it removes "both restatements share a bug in the arithmetic or in the .param/.bin grammar" from the list of worries.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import HURR
from oracle import oracle
from upscale_video_b200 import ncnn_model


def torch_run_graph(graph, x_hwc):
    """ncnn layer semantics (SURVEY.md section 2 op table) on torch float64, NCHW."""
    blobs = {}
    t64 = lambda a: torch.from_numpy(np.asarray(a, np.float64))
    for L in graph.layers:
        t = L.type
        if t == "Input":
            blobs[L.tops[0]] = t64(x_hwc).permute(2, 0, 1)[None]
        elif t == "Split":
            for top in L.tops:
                blobs[top] = blobs[L.bottoms[0]]
        elif t == "Convolution":
            cout, k, pad = L.p(0), L.p(1), L.p(4)
            v = blobs[L.bottoms[0]]
            w = t64(L.weights["weight"]).reshape(cout, v.shape[1], k, k)
            y = F.conv2d(v, w, t64(L.weights["bias"]) if L.p(5) else None, padding=pad)
            if L.p(9) == 2:
                y = F.leaky_relu(y, float(np.float32(L.p(10)[0] if isinstance(L.p(10), (list, tuple, np.ndarray)) else L.p(10))))
            else:
                assert L.p(9) == 0
            blobs[L.tops[0]] = y
        elif t == "PReLU":
            blobs[L.tops[0]] = F.prelu(blobs[L.bottoms[0]], t64(L.weights["slope"]))
        elif t == "PixelShuffle":
            blobs[L.tops[0]] = F.pixel_shuffle(blobs[L.bottoms[0]], L.p(0, 1))
        elif t == "Interp":
            assert L.p(0) == 1
            sy, sx = float(L.p(1, 1.0)), float(L.p(2, 1.0))
            v = blobs[L.bottoms[0]]
            blobs[L.tops[0]] = v if sy == sx == 1.0 else F.interpolate(v, scale_factor=(sy, sx), mode="nearest")
        elif t == "BinaryOp":
            assert L.p(0) == 0
            blobs[L.tops[0]] = blobs[L.bottoms[0]] + blobs[L.bottoms[1]]
        elif t == "Eltwise":
            assert L.p(0) == 1
            co = L.p(1, [1.0] * len(L.bottoms))
            blobs[L.tops[0]] = sum(blobs[b] * float(np.float32(c)) for b, c in zip(L.bottoms, co))
        elif t == "Concat":
            assert L.p(0) == 0
            blobs[L.tops[0]] = torch.cat([blobs[b] for b in L.bottoms], dim=1)
        else:
            raise NotImplementedError(t)
    return blobs["output"][0].permute(1, 2, 0).numpy()


def natural(h, w, seed):
    rng = np.random.default_rng(seed)
    base = np.linspace(20, 230, w)[None, :, None] * np.ones((h, 1, 3)) * np.array([1.0, 0.8, 0.6])
    return np.clip(base + rng.normal(0, 10, (h, w, 3)), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("stem,shape", [("2x_Compact_Pretrain", (40, 56)), ("4x_Compact_Pretrain", (24, 40)),
                                         (HURR, (48, 64)), ("4x_Valar_v1", (12, 20))])
def test_oracle_equals_torch_f64(stem, shape, model_dir, oracle_models):
    torch.set_num_threads(8)
    graph = ncnn_model.load_model(model_dir, stem)   # the product's reader
    layers = oracle_models(stem)                      # the oracle's own reader
    img = natural(*shape, seed=len(stem))
    x = oracle.from_pixels_normalize(img)             # the reference's pre-processing (float32 multiply)
    mine = oracle.run_graph(layers, x, "f64") * 255.0
    other = torch_run_graph(graph, x) * 255.0
    assert mine.shape == other.shape
    assert np.abs(mine - other).max() < 1e-8, np.abs(mine - other).max()
    assert np.array_equal(oracle.saturate_u8(mine), oracle.saturate_u8(other))


def test_noise_input_saturating(model_dir, oracle_models):
    img = np.random.default_rng(5).integers(0, 256, (32, 48, 3), dtype=np.uint8)
    x = oracle.from_pixels_normalize(img)
    a = oracle.run_graph(oracle_models("2x_Compact_Pretrain"), x, "f64") * 255.0
    b = torch_run_graph(ncnn_model.load_model(model_dir, "2x_Compact_Pretrain"), x) * 255.0
    assert np.abs(a - b).max() < 1e-8 and np.array_equal(oracle.saturate_u8(a), oracle.saturate_u8(b))
