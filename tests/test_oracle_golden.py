"""CPU: the oracle (oracle/oracle.py + oracle.c) against the frozen known-answer vectors in tests/golden/.

The goldens were produced by tools/make_goldens.py from the reference's ORIGINAL ncnn .param/.bin files; here the
oracle reads this repo's converted .b2sr containers, so the test also pins the conversion.  (The reference itself
ships no expected outputs -- "parity unpinned", see oracle/oracle.py.)
"""
import numpy as np
import pytest

from conftest import HURR, golden
from oracle import oracle


@pytest.mark.parametrize("name,model,scale", [
    ("compact2x_crop", "2x_Compact_Pretrain", 2),
    ("compact2x_noise_a", "2x_Compact_Pretrain", 2),
    ("compact2x_noise_ragged", "2x_Compact_Pretrain", 2),
    ("compact4x_crop", "4x_Compact_Pretrain", 4),
])
def test_upscale_goldens_f64_exact(name, model, scale, oracle_models):
    g = golden(name)
    y = oracle.upscale_image_array(oracle_models(model), g["x"], scale, "f64")
    assert y.dtype == np.uint8 and np.array_equal(y, g["y"])


def test_seam_x_golden(oracle_models):
    """1000-px-wide strip: two tiles in x with the reference's 10-px halo (process_tile :409-427)."""
    g = golden("compact2x_seam_x")
    y = oracle.upscale_image_array(oracle_models("2x_Compact_Pretrain"), g["x"], 2, "f64")
    assert np.array_equal(y, g["y"])


def test_hurr_and_chain_goldens(oracle_models):
    g = golden("hurr1x_crop")
    y1 = oracle.apply_model_array(oracle_models(HURR), g["x"], "f64")
    assert np.array_equal(y1, g["y"])
    c = golden("chain_hurr_compact2x")
    y2 = oracle.upscale_image_array(oracle_models("2x_Compact_Pretrain"), y1, 2, "f64")
    assert np.array_equal(y2, c["y"])
    n = golden("hurr1x_noise")
    assert np.array_equal(oracle.apply_model_array(oracle_models(HURR), n["x"], "f64"), n["y"])


def test_f32_oracle_within_one_lsb(oracle_models):
    g = golden("compact2x_crop")
    y = oracle.upscale_image_array(oracle_models("2x_Compact_Pretrain"), g["x"], 2, "f32")
    assert np.array_equal(y, g["y_f32"])
    assert np.abs(y.astype(int) - g["y"].astype(int)).max() <= 1


def test_canvas_golden(oracle_models):
    g = golden("compact2x_canvas_f64")
    y = oracle.upscale_canvas(oracle_models("2x_Compact_Pretrain"), g["x"], 2, "f64")
    assert np.allclose(y, g["y"], atol=2e-4)


def test_layer_semantics_small():
    """Known answers for the ncnn layer restatements, by hand."""
    x = np.arange(2 * 2 * 4, dtype=np.float64).reshape(2, 2, 4)  # HWC, C = 1*r*r
    ps = oracle.pixelshuffle(x, 2)
    assert ps.shape == (4, 4, 1)
    # out[y*2+dy][x*2+dx] = in[y][x][dy*2+dx]
    assert ps[0, 0, 0] == x[0, 0, 0] and ps[0, 1, 0] == x[0, 0, 1] and ps[1, 0, 0] == x[0, 0, 2] and ps[3, 3, 0] == x[1, 1, 3]
    up = oracle.nearest(np.arange(6, dtype=np.float64).reshape(2, 3, 1), 2.0, 2.0)
    assert up.shape == (4, 6, 1) and np.array_equal(up[:, :, 0], np.repeat(np.repeat(np.arange(6).reshape(2, 3), 2, 0), 2, 1))
    # 3x3 conv, zero padding: an all-ones kernel over an all-ones image counts the in-image taps
    w = np.ones((1, 1, 3, 3), np.float32)
    y = oracle.conv(np.ones((3, 4, 1)), w, np.zeros(1, np.float32), 1)
    assert np.array_equal(y[:, :, 0], np.array([[4, 6, 6, 4], [6, 9, 9, 6], [4, 6, 6, 4]], float))
    # cv2.imwrite rounding: half to even, saturate
    assert list(oracle.saturate_u8(np.array([0.5, 1.5, 2.5, 254.5, 255.5, -3.0, 300.0]))) == [0, 2, 2, 254, 255, 0, 255]


def test_saturate_matches_cv2(tmp_path):
    """The rounding the oracle applies is what cv2.imwrite does to a float image (reference :288, :519)."""
    import cv2
    vals = np.linspace(-2, 258, 3 * 64 * 5).reshape(5, 64, 3)
    vals[0, :7, 0] = [0.5, 1.5, 2.5, 254.5, 255.5, 127.5, 128.5]
    p = str(tmp_path / "r.png")
    cv2.imwrite(p, vals)
    assert np.array_equal(cv2.imread(p), oracle.saturate_u8(vals))


def test_tile_rects_1080p():
    rects = list(oracle.tile_rects(1080, 1920))
    shapes = [(r[3] - r[2], r[5] - r[4]) for r in rects]
    assert shapes == [(970, 970), (970, 970), (130, 970), (130, 970)]
    assert [(r[3] - r[2], r[5] - r[4]) for r in oracle.tile_rects(540, 960)] == [(540, 960)]


def test_chain_sensitivity_of_the_oracle_itself(oracle_models):
    """BASELINE configs[2] (HurrDeblur -> u8 -> 2x_Compact): why "<= 1 LSB end to end" cannot be the bar for ANY
    implementation of the chain.  The reference quantises stage 1 to u8 on disk (apply_model :288) and stage 2 re-reads it
    (:487); a stage-1 result within 1 LSB of the oracle's is a different stage-2 INPUT, and the upscaler's gain exceeds 1
    around edges: perturbing 2 % of the oracle's own stage-2 input values by +-1 moves the oracle's own stage-2 output by
    more than 1 LSB in places.  The GPU tests therefore assert <= 1 LSB per stage against the oracle applied to the bytes
    that stage actually received, and <= 3 LSB with < 1 % of values beyond 1 LSB end to end."""
    y = golden("hurr1x_crop")["y"]
    rng = np.random.default_rng(0)
    mask = rng.random(y.shape) < 0.02
    p = np.clip(y.astype(int) + np.where(mask, rng.choice([-1, 1], y.shape), 0), 0, 255).astype(np.uint8)
    m = oracle_models("2x_Compact_Pretrain")
    a, b = oracle.upscale_image_array(m, y, 2, "f64"), oracle.upscale_image_array(m, p, 2, "f64")
    d = np.abs(a.astype(int) - b.astype(int))
    assert d.max() >= 2 and (d > 1).mean() < 0.01


def test_rrdb_op_semantics_hand_vectors():
    """Per-op known answers, computed by hand, for the ncnn layers only 4x_Valar_v1 uses (reference
    models/4x_Valar_v1.param:6-22, :1203): Convolution with fused LeakyReLU (9=2, -23310=1,0.2), bias-less 1x1 Convolution,
    Concat along channels, Eltwise SUM with coefficients (-23301=2,c0,c1: out = c0 * in0 + c1 * in1, operand order = bottom
    order) and Interp nearest x2 (0=1: source index = floor(dst / 2)).  ncnn's documented parameter ids, from the public
    operator table: Convolution 0 = num_output, 1 = kernel, 4 = pad, 5 = bias_term, 9 = activation_type (2 = LeakyReLU),
    10 = activation_params; Eltwise 0 = op_type (1 = SUM), 1 = coeffs; Interp 0 = resize_type (1 = nearest), 1/2 = scales;
    Concat 0 = axis (0 = channels of a CHW blob)."""
    def layer(t, bottoms, tops, params=None, arrays=None):
        return {"type": t, "name": tops[0], "bottoms": bottoms, "tops": tops, "params": params or {}, "arrays": arrays or {}}

    x = np.array([[[1.0, -2.0], [3.0, -4.0]],
                  [[-5.0, 6.0], [7.0, -8.0]]])  # H = 2, W = 2, C = 2 (HWC)
    w1 = np.array([[1.0, 1.0], [1.0, -1.0], [0.5, 0.0]], np.float32).reshape(3, 2, 1, 1)  # 1x1 conv 2 -> 3, OIHW
    b1 = np.array([0.0, 1.0, -10.0], np.float32)
    layers = [
        layer("Input", [], ["input"]),
        layer("Split", ["input"], ["a", "b", "c"]),
        # y = lrelu_0.2(conv1x1(x) + b): per pixel (s, d, h) = (x0 + x1, x0 - x1 + 1, 0.5 x0 - 10)
        layer("Convolution", ["a"], ["y"], {0: 3, 1: 1, 4: 0, 5: 1, 6: 6, 9: 2, 10: [0.2]}, {"weight": w1, "bias": b1}),
        # z = bias-less 1x1 conv 2 -> 2 that swaps the channels
        layer("Convolution", ["b"], ["z"], {0: 2, 1: 1, 4: 0, 5: 0, 6: 4},
              {"weight": np.array([[0.0, 1.0], [1.0, 0.0]], np.float32).reshape(2, 2, 1, 1)}),
        layer("Eltwise", ["z", "c"], ["e"], {0: 1, 1: [0.2, 1.0]}),  # e = 0.2 z + 1.0 x
        layer("Concat", ["y", "e"], ["cat"], {0: 0}),                # channels: y0 y1 y2 e0 e1
        layer("Interp", ["cat"], ["output"], {0: 1, 1: 2.0, 2: 2.0}),
    ]
    out = oracle.run_graph(layers, x, "f64")
    assert out.shape == (4, 4, 5)
    lrelu = lambda v: v if v >= 0 else 0.2 * v  # noqa: E731
    for yy in range(2):
        for xx in range(2):
            x0, x1 = x[yy, xx]
            want = [lrelu(x0 + x1), lrelu(x0 - x1 + 1.0), lrelu(0.5 * x0 - 10.0), 0.2 * x1 + x0, 0.2 * x0 + x1]
            for dy in range(2):
                for dx in range(2):  # nearest x2: every source pixel fills a 2 x 2 block
                    # (0.2 is stored as a float32 coefficient / slope in the .param file)
                    assert np.allclose(out[2 * yy + dy, 2 * xx + dx], want, rtol=0, atol=1e-7), (yy, xx, dy, dx)
    # spot values: pixel (0, 0) = (1, -2): y = (lrelu(-1), 4, lrelu(-9.5)) = (-0.2, 4, -1.9); e = (0.2 * -2 + 1, 0.2 * 1 - 2) = (0.6, -1.8)
    assert np.allclose(out[1, 1], [-0.2, 4.0, -1.9, 0.6, -1.8], atol=1e-7)
    # Eltwise operand order matters: swapping the bottoms must change the result
    layers[4] = layer("Eltwise", ["c", "z"], ["e"], {0: 1, 1: [0.2, 1.0]})
    assert np.allclose(oracle.run_graph(layers, x, "f64")[0, 0, 3:], [0.2 * 1.0 - 2.0, 0.2 * -2.0 + 1.0], atol=1e-7)


REF_GLUE_CASES = [  # (name, model, scale, reference function) -- tools/make_ref_glue_goldens.py
    ("hurr_upscale_14x1000", HURR, 1, "upscale_image"), ("hurr_upscale_1000x12", HURR, 1, "upscale_image"),
    ("hurr_upscale_12x965", HURR, 1, "upscale_image"), ("hurr_upscale_12x969", HURR, 1, "upscale_image"),
    ("hurr_upscale_12x970", HURR, 1, "upscale_image"), ("hurr_upscale_12x971", HURR, 1, "upscale_image"),
    ("hurr_upscale_969x11", HURR, 1, "upscale_image"), ("hurr_upscale_970x11", HURR, 1, "upscale_image"),
    ("hurr_upscale_9x9", HURR, 1, "upscale_image"), ("hurr_apply_40x60", HURR, 1, "apply_model"),
    ("compact2x_10x980", "2x_Compact_Pretrain", 2, "upscale_image"), ("compact2x_972x8", "2x_Compact_Pretrain", 2, "upscale_image"),
    ("compact4x_8x975", "4x_Compact_Pretrain", 4, "upscale_image"), ("valar4x_5x964", "4x_Valar_v1", 4, "upscale_image"),
]


def ref_glue_big_input():
    """Same generator as tools/make_ref_glue_goldens.py::big_case_input (the four-tile case stores no input)."""
    yy, xx = np.mgrid[0:975, 0:972]
    rng = np.random.default_rng(975972)
    base = np.stack([120 + 90 * np.sin(xx / 37.0) * np.cos(yy / 53.0), 40 + 0.2 * xx + 30 * ((xx // 16 + yy // 12) % 2), 200 - 0.15 * yy], -1)
    return np.clip(base + rng.normal(0, 6, base.shape), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("name,model,scale,fn", REF_GLUE_CASES)
def test_oracle_glue_equals_reference_code(name, model, scale, fn, oracle_models):
    """PINS THE GLUE: tests/golden/ref_glue.npz holds what the reference's OWN upscale_processing.py (imported unmodified in the
    build container, init_worker -> upscale_image -> process_tile / apply_model, PNG in, PNG out) wrote when only ncnn_vulkan was
    replaced by a numpy stand-in whose network run is the oracle's float32 layer interpreter.  The oracle's restatement of that
    glue -- tile grid, the '>= 10 px remain' halo rule on every side, crop of the halo, unswapped BGR, 1/255 and *255 in float32,
    float64 canvas, imwrite rounding -- must reproduce those files bit for bit (reference upscale_processing.py:258-299, :395-519)."""
    g = golden("ref_glue")
    x, want = g[name + "__x"], g[name + "__y"]
    layers = oracle_models(model)
    got = oracle.upscale_image_array(layers, x, scale, "f32") if fn == "upscale_image" else oracle.apply_model_array(layers, x, "f32")
    assert got.shape == want.shape and np.array_equal(got, want), (name, int(np.abs(got.astype(int) - want.astype(int)).max()))


def test_oracle_glue_equals_reference_code_four_tiles(oracle_models):
    """The same for a 975 x 972 frame: four tiles, three of them slivers whose halo exists on one side only (digest of the whole
    output + the strips around both seams, as written by the reference's own code)."""
    import hashlib
    g = golden("ref_glue")
    got = oracle.upscale_image_array(oracle_models(HURR), ref_glue_big_input(), 1, "f32")
    assert np.array_equal(got[940:975], g["hurr_upscale_975x972__rows_940_975"])
    assert np.array_equal(got[:, 940:972], g["hurr_upscale_975x972__cols_940_972"])
    assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).digest() == g["hurr_upscale_975x972__sha256"].tobytes()
