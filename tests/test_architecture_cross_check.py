"""CPU: a THIRD reading of the reference's networks -- from the published PyTorch architecture definitions, not from the graph.

The oracle (oracle/) and tests/test_torch_cross_check.py both interpret the ncnn `.param` graph layer by layer, so they
share one reading of ncnn's wiring semantics (Split / Concat order, BinaryOp / Eltwise operands and coefficients, PixelShuffle
mode, Interp index rule).  The model files were exported from PyTorch networks whose `forward` is public:

  * `SRVGGNetCompact` (Real-ESRGAN `realesrgan/archs/srvgg_arch.py`): body = conv(3, nf) + PReLU(nf), num_conv x
    (conv(nf, nf) + PReLU(nf)), conv(nf, 3 s^2); out = pixel_shuffle(body(x), s) + interpolate(x, scale_factor=s, 'nearest')
    -- reference models/2x_Compact_Pretrain.param:1-42 (nf 64, 16 convs), 4x_Compact_Pretrain.param (s 4),
    1x_HurrDeblur_SubCompact_nf24-nc8_244k_net_g.param:1-26 (nf 24, 8 convs, s 1);
  * the ESRGAN "plus" RRDBNet (old-arch `RRDB_Net`, `ResidualDenseBlock_5C` with `plus=True`): x1 = lrelu(c1(x)),
    x2 = lrelu(c2(cat(x, x1))) + conv1x1(x), x3 = lrelu(c3(cat(x, x1, x2))), x4 = lrelu(c4(cat(..., x3))) + x2,
    x5 = c5(cat(..., x4)), out = 0.2 x5 + x; RRDB: 0.2 rdb3(rdb2(rdb1(x))) + x; trunk conv; fea + trunk; two
    (nearest x2, conv, lrelu 0.2); HRconv + lrelu; conv_last -- reference models/4x_Valar_v1.param:1-1208.

Here those `forward`s are written out on torch float64 and given the weights of the ncnn files, assigned by tensor SHAPE and
order of appearance only (no use of the graph's blob names or wiring).  If the oracle's reading of any wiring rule were wrong, the
two would disagree.  What stays unpinned is ncnn's own arithmetic (fp16 Vulkan shaders), which nothing here can run.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import HURR
from oracle import oracle
from upscale_video_b200 import ncnn_model


def _convs_and_slopes(graph):
    """(weight OIHW, bias) of every Convolution and the slope vector of every PReLU, in file order."""
    convs, slopes = [], []
    t64 = lambda a: torch.from_numpy(np.asarray(a, np.float64))  # noqa: E731
    for L in graph.layers:
        if L.type == "Convolution":
            cout, k = L.p(0), L.p(1)
            w = t64(L.weights["weight"])
            cin = w.numel() // (cout * k * k)
            convs.append((w.reshape(cout, cin, k, k), t64(L.weights["bias"]) if L.p(5) else None))
        elif L.type == "PReLU":
            slopes.append(t64(L.weights["slope"]))
    return convs, slopes


def srvgg_compact_forward(graph, x_hwc, scale):
    """SRVGGNetCompact.forward with the file's convolutions / PReLUs taken in order."""
    convs, slopes = _convs_and_slopes(graph)
    assert len(slopes) == len(convs) - 1  # every convolution but the last is followed by a PReLU
    x = torch.from_numpy(np.asarray(x_hwc, np.float64)).permute(2, 0, 1)[None]
    out = x
    for i, (w, b) in enumerate(convs):
        assert w.shape[2] == 3
        out = F.conv2d(out, w, b, padding=1)
        if i < len(slopes):
            out = F.prelu(out, slopes[i])
    assert out.shape[1] == 3 * scale * scale
    out = F.pixel_shuffle(out, scale)
    base = x if scale == 1 else F.interpolate(x, scale_factor=scale, mode="nearest")
    return (out + base)[0].permute(1, 2, 0).numpy()


def rrdbnet_plus_forward(graph, x_hwc, nf=64, gc=32, nb=23):
    """RRDB_Net.forward (ESRGAN 'plus' variant, upscale 4) with the file's convolutions sorted into the modules by shape."""
    convs, slopes = _convs_and_slopes(graph)
    assert not slopes
    c02 = float(np.float32(0.2))  # the exported file holds the architecture's 0.2 constants as float32 text (2.000000e-01 read into a float)
    lrelu = lambda v: F.leaky_relu(v, c02)  # noqa: E731
    conv = lambda v, wb: F.conv2d(v, wb[0], wb[1], padding=wb[0].shape[2] // 2)  # noqa: E731
    it = iter(convs)
    conv_first = next(it)
    assert tuple(conv_first[0].shape) == (nf, 3, 3, 3)
    rdbs = []
    for _ in range(nb * 3):
        block = {}
        for _ in range(6):  # five 3x3 convolutions with growing input width + the bias-less 1x1 shortcut, in whatever order the file has them
            wb = next(it)
            cout, cin, k, _ = wb[0].shape
            key = "c1x1" if k == 1 else "c%d" % (1 + (cin - nf) // gc)
            assert key not in block and (k == 3 or (cin, cout, wb[1]) == (nf, gc, None)), (key, tuple(wb[0].shape))
            block[key] = wb
        assert sorted(block) == ["c1", "c1x1", "c2", "c3", "c4", "c5"] and block["c5"][0].shape[0] == nf
        rdbs.append(block)
    trunk_conv, upconv1, upconv2, hr_conv, conv_last = (next(it) for _ in range(5))
    assert next(it, None) is None and tuple(conv_last[0].shape) == (3, nf, 3, 3)

    def rdb(x, m):
        x1 = lrelu(conv(x, m["c1"]))
        x2 = lrelu(conv(torch.cat((x, x1), 1), m["c2"])) + conv(x, m["c1x1"])
        x3 = lrelu(conv(torch.cat((x, x1, x2), 1), m["c3"]))
        x4 = lrelu(conv(torch.cat((x, x1, x2, x3), 1), m["c4"])) + x2
        x5 = conv(torch.cat((x, x1, x2, x3, x4), 1), m["c5"])
        return x5 * c02 + x

    x = torch.from_numpy(np.asarray(x_hwc, np.float64)).permute(2, 0, 1)[None]
    fea = conv(x, conv_first)
    body = fea
    for i in range(nb):
        out = body
        for j in range(3):
            out = rdb(out, rdbs[3 * i + j])
        body = out * c02 + body
    fea = fea + conv(body, trunk_conv)
    fea = lrelu(conv(F.interpolate(fea, scale_factor=2, mode="nearest"), upconv1))
    fea = lrelu(conv(F.interpolate(fea, scale_factor=2, mode="nearest"), upconv2))
    out = conv(lrelu(conv(fea, hr_conv)), conv_last)
    return out[0].permute(1, 2, 0).numpy()


def natural(h, w, seed):
    rng = np.random.default_rng(seed)
    base = np.linspace(20, 230, w)[None, :, None] * np.ones((h, 1, 3)) * np.array([1.0, 0.8, 0.6])
    edges = 40.0 * ((np.arange(w)[None, :, None] // 7 + np.arange(h)[:, None, None] // 5) % 2)
    return np.clip(base + edges + rng.normal(0, 10, (h, w, 3)), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("stem,scale,shape", [("2x_Compact_Pretrain", 2, (33, 47)), ("4x_Compact_Pretrain", 4, (21, 30)), (HURR, 1, (40, 52))])
def test_oracle_equals_srvgg_compact_definition(stem, scale, shape, model_dir, oracle_models):
    torch.set_num_threads(8)
    img = natural(*shape, seed=3 + scale)
    x = oracle.from_pixels_normalize(img)  # the reference's pre-processing (float32 multiply), upscale_processing.py:437-441
    mine = oracle.run_graph(oracle_models(stem), x, "f64") * 255.0
    arch = srvgg_compact_forward(ncnn_model.load_model(model_dir, stem), x, scale) * 255.0
    assert mine.shape == arch.shape == (shape[0] * scale, shape[1] * scale, 3)
    assert np.abs(mine - arch).max() < 1e-8, np.abs(mine - arch).max()
    assert np.array_equal(oracle.saturate_u8(mine), oracle.saturate_u8(arch))


def test_oracle_equals_rrdbnet_plus_definition(model_dir, oracle_models):
    torch.set_num_threads(8)
    img = natural(11, 17, seed=9)
    x = oracle.from_pixels_normalize(img)
    mine = oracle.run_graph(oracle_models("4x_Valar_v1"), x, "f64") * 255.0
    arch = rrdbnet_plus_forward(ncnn_model.load_model(model_dir, "4x_Valar_v1"), x) * 255.0
    assert mine.shape == arch.shape == (44, 68, 3)
    assert np.abs(mine - arch).max() < 1e-8, np.abs(mine - arch).max()
    assert np.array_equal(oracle.saturate_u8(mine), oracle.saturate_u8(arch))
