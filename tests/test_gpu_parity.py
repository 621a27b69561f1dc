"""GPU (-m gpu): the CUDA hot path, called through the C ABI (include/b2sr.h via upscale_video_b200/engine.py),
against the CPU oracle and the frozen goldens.

Bar (north_star): <= 1 LSB per 8-bit channel against the fp64 restatement of the reference graph with the
reference's tiling and rounding.  The engine stores activations in fp16 and accumulates in fp32, so a small
fraction of values may land on the other side of a rounding boundary; MAX_MISMATCH bounds that fraction.
"""
import math
import os

import numpy as np
import pytest

from conftest import HURR, golden
from oracle import oracle

pytestmark = pytest.mark.gpu

MAX_LSB = 1            # tolerance stated by BASELINE.json north_star
MAX_MISMATCH = 0.045   # fraction of u8 values allowed to differ by exactly 1 LSB (measured, profiles/r02zz_parity_measured.jsonl:
                       # 0.2-0.5 % on natural content, 2.5-3.0 % on uniform noise, 4x_Valar_v1 1.1-2.1 %)


@pytest.fixture(scope="module")
def E():
    from upscale_video_b200 import engine
    assert engine.device_count() >= 1, "no CUDA device: the product path has no CPU fallback"
    return engine


@pytest.fixture(scope="module")
def engines(E, model_dir):
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = E.Engine.from_files(model_dir, name, 0)  # default schedule: pipelined tcgen05 where it fits
        return cache[name]

    yield get
    for e in cache.values():
        e.close()


def natural(h, w, seed=0):
    """Smooth gradients + edges + noise: unlike uniform noise this does not saturate the outputs."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([128 + 90 * np.sin(xx / 37.0 + c) * np.cos(yy / 23.0 - c) for c in range(3)], -1)
    img += 40 * ((xx // 64 + yy // 48) % 2)[..., None]
    img += rng.normal(0, 6, (h, w, 3))
    return np.clip(img, 0, 255).astype(np.uint8)


def _record(name, **kv):
    """Measured parity figures go to gpurun_out/ (when that scratch dir exists) so tolerances can be set from data."""
    import json
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_measured.jsonl"), "a") as f:
            f.write(json.dumps(dict(name=name, **kv)) + "\n")


def assert_parity(got, ref, what="", max_mismatch=MAX_MISMATCH):
    assert got.shape == ref.shape and got.dtype == np.uint8, what
    d = np.abs(got.astype(np.int32) - ref.astype(np.int32))
    _record(what, max_lsb=int(d.max()) if d.size else 0, mismatch_fraction=float((d > 0).mean()) if d.size else 0.0, bound=max_mismatch,
            values=int(d.size))
    assert d.max() <= MAX_LSB, "%s: max |diff| = %d LSB at %s" % (what, d.max(), np.unravel_index(d.argmax(), d.shape))
    # the fraction bound is only meaningful on images with enough values (a 1x1 frame has 12)
    assert (d > 0).sum() <= max(2, max_mismatch * d.size), "%s: %.2f%% of values differ" % (what, 100 * (d > 0).mean())


# ---------------------------------------------------------------- goldens ----------------------------
@pytest.mark.parametrize("name,model", [
    ("compact2x_crop", "2x_Compact_Pretrain"),
    ("compact2x_noise_a", "2x_Compact_Pretrain"),
    ("compact2x_noise_ragged", "2x_Compact_Pretrain"),
    ("compact2x_seam_x", "2x_Compact_Pretrain"),
    ("compact2x_seam_y", "2x_Compact_Pretrain"),
    ("compact4x_crop", "4x_Compact_Pretrain"),
])
def test_upscale_goldens(name, model, engines):
    g = golden(name)
    assert_parity(engines(model).run_u8(g["x"]), g["y"], name)


@pytest.mark.parametrize("name,model,scale,fn", [
    ("hurr_upscale_14x1000", HURR, 1, "upscale_image"), ("hurr_upscale_1000x12", HURR, 1, "upscale_image"),
    ("hurr_upscale_12x965", HURR, 1, "upscale_image"), ("hurr_upscale_12x969", HURR, 1, "upscale_image"),
    ("hurr_upscale_12x970", HURR, 1, "upscale_image"), ("hurr_upscale_12x971", HURR, 1, "upscale_image"),
    ("hurr_upscale_969x11", HURR, 1, "upscale_image"), ("hurr_upscale_970x11", HURR, 1, "upscale_image"),
    ("hurr_upscale_9x9", HURR, 1, "upscale_image"), ("hurr_apply_40x60", HURR, 1, "apply_model"),
    ("compact2x_10x980", "2x_Compact_Pretrain", 2, "upscale_image"), ("compact2x_972x8", "2x_Compact_Pretrain", 2, "upscale_image"),
    ("compact4x_8x975", "4x_Compact_Pretrain", 4, "upscale_image"), ("valar4x_5x964", "4x_Valar_v1", 4, "upscale_image"),
])
def test_reference_code_goldens(name, model, scale, fn, engines, model_dir):
    """The device against files written by the reference's OWN upscale_processing.py (tools/make_ref_glue_goldens.py: the reference
    module imported unmodified, only ncnn_vulkan replaced by a numpy stand-in running the oracle's float32 layers): the tile grid,
    halo rule, crops and rounding the device reproduces are the reference's, not a restatement's."""
    if model == "4x_Valar_v1" and not os.path.exists(os.path.join(model_dir, "4x_Valar_v1.b2sr")):
        pytest.skip("4x_Valar_v1 not packaged")
    g = golden("ref_glue")
    x, want = g[name + "__x"], g[name + "__y"]
    eng = engines(model)
    got = eng.run_u8(x) if fn == "upscale_image" else eng.run_u8(x, tile=0, halo=0)
    assert_parity(got, want, name, max_mismatch=0.045)


def test_hurr_goldens_and_chain(engines):
    """1x pre-pass (apply_model, untiled) and the chained config with its u8 hop between the two networks
    (reference apply_model :288 writes u8, upscale_image :487 re-reads it)."""
    for name in ("hurr1x_crop", "hurr1x_noise"):
        g = golden(name)
        assert_parity(engines(HURR).run_u8(g["x"], tile=0, halo=0), g["y"], name)
    g = golden("hurr1x_crop")
    y1 = engines(HURR).run_u8(g["x"], tile=0, halo=0)
    c = golden("chain_hurr_compact2x")
    # the golden chains from the oracle's own stage-1 output; feed that to isolate stage 2, then the full chain
    assert_parity(engines("2x_Compact_Pretrain").run_u8(g["y"]), c["y"], "chain stage 2")
    # End to end, the bar is NOT +-1: stage 1 is within 1 LSB of the oracle, and a 1-LSB difference in the u8 image handed to
    # stage 2 is a different INPUT to a network whose gain around edges exceeds 1 -- the oracle itself moves by up to 3 LSB
    # when its stage-2 input is perturbed by +-1 (tests/test_oracle_golden.py::test_chain_sensitivity_of_the_oracle_itself).  So the honest statement
    # for configs[2] is: every stage <= 1 LSB against the oracle applied to the bytes that stage actually received
    # (asserted above and in tests/test_zz_gpu_cli.py), end-to-end <= 3 LSB with the > 1 LSB fraction reported.
    d = np.abs(engines("2x_Compact_Pretrain").run_u8(y1).astype(int) - c["y"].astype(int))
    _record_chain = dict(max_lsb=int(d.max()), frac_gt1=float((d > 1).mean()), frac_gt0=float((d > 0).mean()))
    assert d.max() <= 3 and (d > 1).mean() < 0.01, _record_chain


def test_float_canvas(engines):
    """b2sr_run_f32 = what process_tile scatters into the canvas before imwrite (reference :462-477)."""
    g = golden("compact2x_canvas_f64")
    y = engines("2x_Compact_Pretrain").run_f32(g["x"])
    assert y.dtype == np.float32 and np.abs(y - g["y"]).max() < 0.6  # fp16 activations: ~0.1-0.3 of an LSB


# ---------------------------------------------------------------- oracle, seeded, edge shapes ----------
@pytest.mark.parametrize("h,w", [(1, 1), (2, 3), (3, 127), (5, 128), (4, 129), (7, 130), (9, 255), (6, 257), (33, 385), (70, 64)])
def test_edge_shapes_vs_oracle(h, w, engines, oracle_models):
    """Band boundaries every 128 columns, single-row/column frames, ragged tails."""
    img = natural(h, w, seed=h * 1000 + w)
    ref = oracle.upscale_image_array(oracle_models("2x_Compact_Pretrain"), img, 2, "f64")
    assert_parity(engines("2x_Compact_Pretrain").run_u8(img), ref, "%dx%d" % (h, w))


@pytest.mark.parametrize("model,scale", [("4x_Compact_Pretrain", 4), (HURR, 1)])
def test_other_models_vs_oracle(model, scale, engines, oracle_models):
    img = natural(45, 150, seed=5)
    if scale == 1:
        ref = oracle.apply_model_array(oracle_models(model), img, "f64")
        got = engines(model).run_u8(img, tile=0, halo=0)
    else:
        ref = oracle.upscale_image_array(oracle_models(model), img, scale, "f64")
        got = engines(model).run_u8(img)
    assert_parity(got, ref, model)


@pytest.mark.parametrize("h,w", [(20, 960), (20, 961), (20, 969), (20, 970), (20, 971), (20, 1920), (20, 1921), (961, 40), (970, 24)])
def test_tile_boundary_sizes(h, w, engines, oracle_models):
    """Frame sizes on and around the reference's tile arithmetic (960-px tiles, a 10-px halo only where >= 10 px of image
    remain on that side, reference :409-427): one tile exactly, one tile + 1..11 columns (the second tile is narrower than
    its halo), two tiles exactly, and the same in y."""
    img = natural(h, w, seed=h + w)
    ref = oracle.upscale_image_array(oracle_models("2x_Compact_Pretrain"), img, 2, "f32")
    assert_parity(engines("2x_Compact_Pretrain").run_u8(img), ref, "%dx%d" % (h, w))


def test_tile_seams_four_tiles(engines, oracle_models):
    """A frame with seams in both directions (2 x 2 reference tiles): 980 x 1000."""
    img = natural(980, 1000, seed=11)
    rects = list(oracle.tile_rects(980, 1000))
    assert len(rects) == 4
    ref = oracle.upscale_image_array(oracle_models("2x_Compact_Pretrain"), img, 2, "f32")
    assert_parity(engines("2x_Compact_Pretrain").run_u8(img), ref, "980x1000")


def test_saturation_and_noise(engines, oracle_models):
    """Uniform noise drives outputs into both clamps (cv2.imwrite saturation)."""
    img = np.random.default_rng(3).integers(0, 256, (40, 200, 3), dtype=np.uint8)
    img[:8] = 0
    img[8:16] = 255
    ref = oracle.upscale_image_array(oracle_models("2x_Compact_Pretrain"), img, 2, "f64")
    got = engines("2x_Compact_Pretrain").run_u8(img)
    assert (ref == 0).any() and (ref == 255).any()
    assert_parity(got, ref, "noise")


def test_three_device_schedules_agree(E, engines, model_dir):
    """Independent device implementations of the same arithmetic (fp16 storage, fp32 accumulate): CUDA-core kernels,
    tcgen05 layer by layer, tcgen05 pipelined (one persistent kernel, L2 row rings).  The two tcgen05 schedules issue
    the same MMAs per row and must agree bit for bit."""
    img = natural(64, 300, seed=2)
    simple = E.Engine.from_files(model_dir, "2x_Compact_Pretrain", 0)
    simple.set_option(E.OPT_IMPL, E.IMPL_SIMPLE)
    a = simple.run_u8(img)
    assert simple.stat(E.STAT_TC_LAUNCHES) == 0
    layer = E.Engine.from_files(model_dir, "2x_Compact_Pretrain", 0)
    layer.set_option(E.OPT_IMPL, E.IMPL_TCGEN05)
    b = layer.run_u8(img)
    assert layer.stat(E.STAT_TC_LAUNCHES) == 18 and layer.stat(E.STAT_PIPE_LAUNCHES) == 0  # one launch per convolution
    pipe = engines("2x_Compact_Pretrain")
    pipe.set_option(E.OPT_IMPL, E.IMPL_PIPELINED)
    pipe.reset_stats()
    c = pipe.run_u8(img)
    assert pipe.stat(E.STAT_PIPE_LAUNCHES) == 1 and pipe.stat(E.STAT_LAUNCHES) == 1  # the persistent kernel reads the u8 frames itself
    pipe.set_option(E.OPT_IMPL, E.IMPL_AUTO)
    d = np.abs(a.astype(int) - b.astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 0.01
    assert np.array_equal(b, c)
    l_a, l_b = simple.debug_layer(img, 8), layer.debug_layer(img, 8)
    assert np.abs(l_a - l_b).max() < 0.05 * np.abs(l_a).max()
    # a frame with seams in x and y, small rings (forces ring wrap-around and back-pressure), both schedules
    img = natural(1000, 1100, seed=8)
    pipe.set_option(E.OPT_RING_ROWS, 8)
    c = pipe.run_u8(img)
    pipe.set_option(E.OPT_RING_ROWS, 0)
    assert np.array_equal(layer.run_u8(img), c)
    # the narrow nf = 24 network: `auto` = one launch per layer (measured 2.7x faster than its persistent grid); the
    # persistent schedule stays available on request and must agree bit for bit
    hurr = E.Engine.from_files(model_dir, HURR, 0)
    img = natural(90, 500, seed=3)
    a = hurr.run_u8(img, tile=0, halo=0)
    assert hurr.stat(E.STAT_PIPE_LAUNCHES) == 0 and hurr.stat(E.STAT_TC_LAUNCHES) == 10 and hurr.stat(E.STAT_PIPE_FALLBACKS) == 0
    hurr.set_option(E.OPT_IMPL, E.IMPL_PIPELINED)
    assert np.array_equal(hurr.run_u8(img, tile=0, halo=0), a) and hurr.stat(E.STAT_PIPE_LAUNCHES) == 1
    hurr.close()
    simple.close()
    layer.close()


def test_pipelined_schedule_random_shapes():
    """40 random (model, batch, size, tiling, ring size) draws: the persistent pipelined kernel must reproduce the
    layer-by-layer schedule bit for bit (tests/stress_gpu.py; ring sizes down to 4 rows force constant back-pressure)."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "stress_gpu.py"), "40", "7"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "40 iterations, 0 mismatches" in r.stdout


def test_wide_frame_falls_back_to_layer_schedule(E, model_dir, oracle_models):
    """layers x bands > SM count (18 x 16 for a 2048-wide untiled frame): auto mode runs layer by layer, forcing the
    pipelined schedule is a clean error."""
    eng = E.Engine.from_files(model_dir, "2x_Compact_Pretrain", 0)
    img = natural(24, 2048, seed=6)
    out = eng.run_u8(img, tile=0, halo=0)
    assert eng.stat(E.STAT_PIPE_LAUNCHES) == 0 and eng.stat(E.STAT_TC_LAUNCHES) == 18
    ref = oracle.upscale_image_array(oracle_models("2x_Compact_Pretrain"), img, 2, "f64", tile_size=10**6)
    assert_parity(out, ref, "wide")
    eng.set_option(E.OPT_IMPL, E.IMPL_PIPELINED)
    with pytest.raises(E.EngineError, match="pipelined"):
        eng.run_u8(img, tile=0, halo=0)
    eng.close()


# ---------------------------------------------------------------- full size (BASELINE configs) --------
def test_full_1080p_vs_oracle(engines, oracle_models):
    """BASELINE configs[1] frame size, reference tiling (4 tiles), whole frame against the oracle."""
    img = natural(1080, 1920, seed=21)
    ref = oracle.upscale_image_array(oracle_models("2x_Compact_Pretrain"), img, 2, "f32")
    assert_parity(engines("2x_Compact_Pretrain").run_u8(img), ref, "1080p")


def test_config3_chain_1080p(E, engines, oracle_models):
    """BASELINE configs[2]: 1080p, 1x_HurrDeblur (untiled apply_model, u8 on "disk") -> 2x_Compact (tiled upscale_image).
    Each stage is checked against the oracle on the SAME input (the second stage on the engine's own stage-1 output),
    so a legitimate 1-LSB difference after stage 1 is not mistaken for a stage-2 error."""
    img = natural(1080, 1920, seed=33)
    hurr, comp = engines(HURR), engines("2x_Compact_Pretrain")
    hurr.reset_stats()
    y1 = hurr.run_u8(img, tile=0, halo=0)
    # 10 layers x 15 bands = 150 CTAs > 148 SMs: this shape runs layer by layer (10 launches + prep)
    assert hurr.stat(E.STAT_PIPE_LAUNCHES) == 0 and hurr.stat(E.STAT_TC_LAUNCHES) == 10
    assert_parity(y1, oracle.apply_model_array(oracle_models(HURR), img, "f32"), "chain stage 1 (HurrDeblur 1080p)")
    y2 = comp.run_u8(y1)
    assert_parity(y2, oracle.upscale_image_array(oracle_models("2x_Compact_Pretrain"), y1, 2, "f32"), "chain stage 2 (2x_Compact 1080p)")


def test_config4_540p_4x_compact(E, engines, oracle_models):
    """BASELINE configs[3] names 4x_Valar_v1 "(4x pixel-shuffle)"; Valar is an RRDB without pixel shuffle (SURVEY section 0.1,
    out of scope this round), the reference's 4x pixel-shuffle model is 4x_Compact_Pretrain: 540p -> 2160p, one tile."""
    img = natural(540, 960, seed=44)
    eng = engines("4x_Compact_Pretrain")
    eng.reset_stats()
    out = eng.run_u8(img)
    assert eng.stat(E.STAT_PIPE_LAUNCHES) == 1 and out.shape == (2160, 3840, 3)
    assert_parity(out, oracle.upscale_image_array(oracle_models("4x_Compact_Pretrain"), img, 4, "f32"), "4x 540p")


def test_full_size_properties(E, engines):
    """Size-independent properties at 1080p / 540p: batch == per-frame, device tiling == pasting separately run
    tiles, determinism, host pipeline == device path."""
    import torch
    eng = engines("2x_Compact_Pretrain")
    frames = np.stack([natural(1080, 1920, seed=s) for s in (1, 2, 3)])
    single = [eng.run_u8(f) for f in frames]
    d_in = torch.from_numpy(frames).cuda()
    d_out = torch.empty((3, 2160, 3840, 3), dtype=torch.uint8, device="cuda")
    eng.set_option(E.OPT_MAX_BATCH, 2)  # 2 + 1 frames: exercises the multi-pass path
    eng.run_batch_device(d_in, d_out, 3, 1080, 1920, sync=True)
    eng.set_option(E.OPT_MAX_BATCH, 0)
    batch = d_out.cpu().numpy()
    for i in range(3):
        assert np.array_equal(batch[i], single[i]), "frame %d: batch != single" % i
    assert np.array_equal(eng.run_u8(frames[0]), single[0])  # determinism
    # tiling on the device == running every reference tile as its own untiled image and pasting the core
    f = frames[1]
    canvas = np.zeros((2160, 3840, 3), np.uint8)
    for (_, _, iy0, iy1, ix0, ix1, cy0, cy1, cx0, cx1) in oracle.tile_rects(1080, 1920):
        t = eng.run_u8(np.ascontiguousarray(f[iy0:iy1, ix0:ix1]), tile=0, halo=0)
        canvas[cy0 * 2:cy1 * 2, cx0 * 2:cx1 * 2] = t[(cy0 - iy0) * 2:(cy1 - iy0) * 2, (cx0 - ix0) * 2:(cx1 - ix0) * 2]
    assert np.array_equal(canvas, single[1])
    # host pipeline (pinned, double-buffered) == device path
    h_in = torch.from_numpy(frames).pin_memory()
    h_out = torch.empty((3, 2160, 3840, 3), dtype=torch.uint8).pin_memory()
    eng.set_option(E.OPT_MAX_BATCH, 1)
    eng.run_batch_host(h_in, h_out, 3, 1080, 1920)
    eng.set_option(E.OPT_MAX_BATCH, 0)
    assert np.array_equal(h_out.numpy(), batch)


def test_host_pipeline_tapered_chunks(E, engines):
    """b2sr_run_batch_host cuts a synchronous call into chunks of 1, B, ..., B, 1 frames (one-frame ends keep the exposed
    H2D / D2H short): whatever the schedule, the frames equal the device-resident batch."""
    import torch
    eng = engines("2x_Compact_Pretrain")
    for n in (1, 2, 3, 7, 11):
        frames = np.stack([natural(270, 480, seed=40 + s) for s in range(n)])
        d_in = torch.from_numpy(frames).cuda()
        d_out = torch.empty((n, 540, 960, 3), dtype=torch.uint8, device="cuda")
        eng.run_batch_device(d_in, d_out, n, 270, 480, sync=True)
        h_in = torch.from_numpy(frames).pin_memory()
        h_out = torch.zeros((n, 540, 960, 3), dtype=torch.uint8).pin_memory()
        eng.run_batch_host(h_in, h_out, n, 270, 480)
        assert np.array_equal(h_out.numpy(), d_out.cpu().numpy()), n


def test_host_pipeline_submit_wait(E, engines):
    """b2sr_submit_batch_host / b2sr_wait_batch: several submissions in flight on their own pinned buffers (more than the
    ticket ring holds, mixed sizes so the staging buffers are replaced under pending work, waits out of order, a synchronous
    call and a single-frame call in between) produce the frames of the device-resident batch."""
    import torch
    eng = engines("2x_Compact_Pretrain")
    jobs = []
    for k in range(11):
        n, (h, w) = 1 + (k * 3) % 6, ((96, 200), (270, 480), (64, 1000))[k % 3]
        frames = np.stack([natural(h, w, seed=500 + 10 * k + s) for s in range(n)])
        h_in = torch.from_numpy(frames).pin_memory()
        h_out = torch.zeros((n, 2 * h, 2 * w, 3), dtype=torch.uint8).pin_memory()
        jobs.append((frames, h_in, h_out, n, h, w, eng.submit_batch_host(h_in, h_out, n, h, w)))
        if k == 4:  # a synchronous call between submissions waits for them and must not disturb them
            assert np.array_equal(eng.run_u8(frames[0]), eng.run_u8(frames[0]))
        if k == 7:
            eng.wait_batch(jobs[6][6])
    for j in (10, 2, 9, 0, 1, 3, 4, 5, 6, 7, 8):  # any order; an early ticket's event slot has been reused by then
        eng.wait_batch(jobs[j][6])
    for frames, h_in, h_out, n, h, w, _ in jobs:
        d_in = torch.from_numpy(frames).cuda()
        d_out = torch.empty((n, 2 * h, 2 * w, 3), dtype=torch.uint8, device="cuda")
        eng.run_batch_device(d_in, d_out, n, h, w, sync=True)
        assert np.array_equal(h_out.numpy(), d_out.cpu().numpy()), (n, h, w)
    with pytest.raises(E.EngineError):
        eng.wait_batch(10 ** 6)
    # pageable (numpy) buffers: the copies become synchronous, the result is the same
    frames, _, h_out, n, h, w, _ = jobs[3]
    np_out = np.zeros((n, 2 * h, 2 * w, 3), np.uint8)
    eng.wait_batch(eng.submit_batch_host(frames, np_out, n, h, w))
    assert np.array_equal(np_out, h_out.numpy())
    # a rejected submission (bad geometry) leaves the pipeline usable
    frames, h_in, h_out, n, h, w, _ = jobs[1]
    with pytest.raises(E.EngineError):
        eng.submit_batch_host(h_in, h_out, n, 0, w)
    before = h_out.numpy().copy()
    h_out.zero_()
    eng.wait_batch(eng.submit_batch_host(h_in, h_out, n, h, w))
    assert np.array_equal(h_out.numpy(), before)


def test_strides_and_device_memory(E, engines):
    import torch
    eng = engines("2x_Compact_Pretrain")
    img = natural(37, 131, seed=9)
    ref = eng.run_u8(img)
    lib = E.load_library()
    # padded host rows
    pin = np.zeros((37, 131 * 3 + 13), np.uint8)
    pin[:, :131 * 3] = img.reshape(37, -1)
    pout = np.full((74, 262 * 3 + 5), 0xAB, np.uint8)
    rc = lib.b2sr_run_u8(eng._h, pin.ctypes.data, 37, 131, pin.strides[0], pout.ctypes.data, pout.strides[0], 960, 10, E.MEM_HOST)
    assert rc == 0 and np.array_equal(pout[:, :262 * 3].reshape(74, 262, 3), ref) and (pout[:, 262 * 3:] == 0xAB).all()
    # device-resident single frame
    d_in = torch.from_numpy(img).cuda()
    d_out = torch.zeros((74, 262, 3), dtype=torch.uint8, device="cuda")
    rc = lib.b2sr_run_u8(eng._h, d_in.data_ptr(), 37, 131, 0, d_out.data_ptr(), 0, 960, 10, E.MEM_DEVICE)
    assert rc == 0 and np.array_equal(d_out.cpu().numpy(), ref)
    # device frames at every byte alignment, and ending exactly at the end of their allocation: the first layer copies the
    # frame rows as ALIGNED 4-byte words (cp.async) and fetches the words that straddle either end of the buffer byte by byte
    for shift in (1, 2, 3):
        for (h, w) in ((37, 131), (5, 1), (3, 130), (9, 257)):
            im = natural(h, w, seed=100 + shift)
            want = eng.run_u8(im)
            buf = torch.zeros(h * w * 3 + shift, dtype=torch.uint8, device="cuda")  # frame occupies the LAST h*w*3 bytes
            buf[shift:] = torch.from_numpy(im.reshape(-1)).cuda()
            out = torch.zeros((h * 2, w * 2, 3), dtype=torch.uint8, device="cuda")
            rc = lib.b2sr_run_batch_device(eng._h, buf.data_ptr() + shift, out.data_ptr(), 1, h, w, 960, 10, 1)
            assert rc == 0 and np.array_equal(out.cpu().numpy(), want), (shift, h, w)


def test_errors_are_codes_not_crashes(E, engines):
    eng = engines("2x_Compact_Pretrain")
    lib = E.load_library()
    img = np.zeros((4, 4, 3), np.uint8)
    out = np.zeros((8, 8, 3), np.uint8)
    assert lib.b2sr_run_u8(eng._h, img.ctypes.data, 0, 4, 0, out.ctypes.data, 0, 960, 10, 0) == -1
    assert lib.b2sr_run_u8(eng._h, img.ctypes.data, 4, 4, 3, out.ctypes.data, 0, 960, 10, 0) == -1  # stride < row
    assert lib.b2sr_run_u8(eng._h, img.ctypes.data, 4, 4, 0, out.ctypes.data, 0, 10, 10, 0) == -1   # halo >= tile
    assert b"tile" in lib.b2sr_last_error()
    with pytest.raises(E.EngineError):
        eng.debug_layer(img, 17)
    assert np.array_equal(eng.run_u8(img), eng.run_u8(img))  # still usable afterwards


# ---------------------------------------------------------------- the drop-in worker functions --------
def test_worker_functions_on_png_files(engines, oracle_models, model_dir, tmp_path, monkeypatch):
    """init_worker / upscale_image / process_tile / apply_model on real PNG files, like the reference's callers
    (test_gpus.py:15-33, test_images.py:131-144)."""
    import cv2
    from upscale_video_b200 import upscale_processing as up
    monkeypatch.chdir(tmp_path)
    img = natural(40, 1000, seed=4)
    cv2.imwrite("1.extract.png", img)
    up.init_worker([0], 0, model_dir, "x_Compact_Pretrain", 2, "input", "output")
    items = up.upscale_image("1.extract.png", "1.png", 2, None, 1, 1, remove=False)
    assert items[-1] == ["info", "Upscaled 1/1"] and [i[0] for i in items[:-1]] == ["debug", "debug"]
    got = cv2.imread("1.png")
    ref = oracle.upscale_image_array(oracle_models("2x_Compact_Pretrain"), img, 2, "f64")
    assert_parity(got, ref, "upscale_image")
    # the reference's own loop over process_tile gives the same picture
    canvas = np.zeros((80, 2000, 3))
    logs = []
    for x in range(2):
        assert up.process_tile(img, 960, 2, 0, x, 40, 1000, canvas, logs) is None
    cv2.imwrite("loop.png", canvas)
    assert np.array_equal(cv2.imread("loop.png"), got)
    # 1x model through apply_model
    up.init_worker([0], 0, model_dir, "x_HurrDeblur_SubCompact_nf24-nc8_244k_net_g", 1, "input", "output")
    items = up.apply_model("1.extract.png", "1.anime.png", True)
    assert items == [["info", "Processed Model: 1.anime.png"]] and not os.path.exists("1.extract.png")
    assert_parity(cv2.imread("1.anime.png"), oracle.apply_model_array(oracle_models(HURR), img, "f64"), "apply_model")
    up._release_engine()


def test_log_items_equal_the_reference_codes(model_dir, tmp_path, monkeypatch):
    """The result protocol (lists of [level, message], reference logging_callback :40-51): the items `upscale_image` returns for
    every form of `frame_batch` (:521-540), without an output file (test_gpus.py:22-31), and `apply_model` (:298) are the ones the
    reference's own functions returned for the same calls (tests/golden/ref_glue.npz, tools/make_ref_glue_goldens.py)."""
    import json
    import cv2
    from upscale_video_b200 import upscale_processing as up
    g = golden("ref_glue")
    want = json.loads(bytes(g["log_items_json"]).decode())
    img = g["log_items_input"]
    monkeypatch.chdir(tmp_path)
    up.init_worker([0], 0, model_dir, "x_Compact_Pretrain", 2, "input", "output")
    for key, frame_batch, out_name in (("batch_none", None, "7.png"), ("batch_int", 3, "7.png"), ("batch_list", [7, 9], "7.png"),
                                       ("no_output_file", None, None)):
        cv2.imwrite("7.extract.png", img)
        got = up.upscale_image("7.extract.png", out_name, 2, frame_batch, 7, 12, remove=key != "no_output_file")
        assert got == want[key], key
        assert os.path.exists("7.extract.png") == (key == "no_output_file")
    up.init_worker([0], 0, model_dir, "x_HurrDeblur_SubCompact_nf24-nc8_244k_net_g", 1, "input", "output")
    cv2.imwrite("7.extract.png", img)
    assert up.apply_model("7.extract.png", "7.anime.png", True) == want["apply_model"]
    up._release_engine()


def test_upscale_frames_pool_two_workers_one_gpu(oracle_models, model_dir, tmp_path, monkeypatch):
    """`-g 0,0`: two spawned workers sharing GPU 0, dynamic frame queue, inputs deleted when done, missing
    inputs skipped (reference upscale_frames :545-601)."""
    import cv2
    from upscale_video_b200 import upscale_processing as up
    monkeypatch.chdir(tmp_path)
    frames = {n: natural(24 + n, 100 + 7 * n, seed=n) for n in (1, 2, 4, 5)}
    for n, im in frames.items():
        cv2.imwrite("%d.extract.png" % n, im)
    import multiprocessing.process as mpp
    used = next(mpp._process_counter)  # children already created by this interpreter (reference's workers_used)
    up.upscale_frames(1, 1, 5, "extract", 2, [0, 0], used, model_dir, "x_Compact_Pretrain", "input", "output", remove=True)
    assert not os.path.exists("3.png")
    for n, im in frames.items():
        assert not os.path.exists("%d.extract.png" % n)
        ref = oracle.upscale_image_array(oracle_models("2x_Compact_Pretrain"), im, 2, "f64")
        assert_parity(cv2.imread("%d.png" % n), ref, "frame %d" % n)
    # the next batch is served by the SAME worker processes (the reference re-creates pool, ncnn.Net and model per
    # 10-minute batch, :565-577): engines, CUDA contexts and scratch persist; the caller's workers_used bookkeeping (which
    # assumes fresh processes per call) no longer decides the GPU slot, so even a stale value is harmless
    assert len(up._pools) == 1
    pool = next(iter(up._pools.values()))
    pids = sorted(p.pid for p in pool._pool)
    more = {n: natural(30 + n, 90 + 5 * n, seed=40 + n) for n in (6, 7, 8)}
    for n, im in more.items():
        cv2.imwrite("%d.extract.png" % n, im)
    up.upscale_frames(2, 6, 8, "extract", 2, [0, 0], 0, model_dir, "x_Compact_Pretrain", "input", "output", remove=True)
    assert len(up._pools) == 1 and sorted(p.pid for p in pool._pool) == pids
    for n, im in more.items():
        assert_parity(cv2.imread("%d.png" % n), oracle.upscale_image_array(oracle_models("2x_Compact_Pretrain"), im, 2, "f64"), "frame %d" % n)
    up.release_workers()
    assert not up._pools


def test_pool_functions_leave_what_the_reference_codes_leave(model_dir, tmp_path, monkeypatch, caplog):
    """`process_model` then `upscale_frames` over a directory of frames with a gap, two workers on GPU 0: the files that exist
    afterwards and the lines logged in the parent are the ones the reference's OWN two functions (spawn pools and all) left in the
    build container for the same calls (tests/golden/ref_glue.npz "pool", tools/make_ref_glue_goldens.py::pool_behaviour)."""
    import json
    import logging
    import cv2
    import multiprocessing.process as mpp
    from upscale_video_b200 import upscale_processing as up
    want = json.loads(bytes(golden("ref_glue")["log_items_json"]).decode())["pool"]
    monkeypatch.chdir(tmp_path)
    for n in (1, 2, 4, 5):
        cv2.imwrite("%d.extract.png" % n, natural(10, 24, seed=70 + n))
    used = next(mpp._process_counter)
    with caplog.at_level(logging.DEBUG):
        up.process_model(5, model_dir, HURR[1:] if HURR[0].isdigit() else HURR, 1, "input", "output", "extract", "anime", [0, 0], used)
    lines = sorted([r.levelname, r.getMessage()] for r in caplog.records if r.name == "root")
    assert sorted(os.listdir(".")) == want["after_process_model"] and lines == want["log_process_model"]
    caplog.clear()
    with caplog.at_level(logging.DEBUG):
        up.upscale_frames(2, 1, 5, "anime", 2, [0, 0], used + 2, model_dir, "x_Compact_Pretrain", "input", "output")
    lines = sorted([r.levelname, r.getMessage()] for r in caplog.records if r.name == "root")
    assert sorted(os.listdir(".")) == want["after_upscale_frames"] and lines == want["log_upscale_frames"]
    up.release_workers()


def test_raw_stream_matches_worker_functions(engines, model_dir, tmp_path):
    """the raw-frame stream (no PNG hop, chained pre-pass on the device) produces exactly the pixels
    the per-frame worker functions produce: apply_model (u8) -> upscale_image."""
    import io
    from upscale_video_b200 import raw_stream
    frames = np.stack([natural(70, 1000, seed=s) for s in (1, 2, 3)])
    hurr, comp = engines(HURR), engines("2x_Compact_Pretrain")
    expect = np.stack([comp.run_u8(hurr.run_u8(f, tile=0, halo=0)) for f in frames])
    out = io.BytesIO()
    n = raw_stream.stream(io.BytesIO(frames.tobytes()), out, 1000, 70, scale=2, models=["a"], chunk=2, model_path=model_dir)
    assert n == 3
    assert np.array_equal(np.frombuffer(out.getvalue(), np.uint8).reshape(3, 140, 2000, 3), expect)
    out = io.BytesIO()  # upscale only, overlapped host pipeline
    raw_stream.stream(io.BytesIO(frames.tobytes()), out, 1000, 70, scale=2, chunk=2, model_path=model_dir)
    assert np.array_equal(np.frombuffer(out.getvalue(), np.uint8).reshape(3, 140, 2000, 3), np.stack([comp.run_u8(f) for f in frames]))


def test_raw_stream_valar(model_dir, E):
    """`raw_stream -m r -s 4`: the RRDB upscaler behind the raw-frame pipe (host pipeline, several frames per pass) equals the
    per-frame worker path bit for bit."""
    import io
    from upscale_video_b200 import raw_stream
    if not os.path.exists(os.path.join(model_dir, "4x_Valar_v1.b2sr")):
        pytest.skip("4x_Valar_v1 not packaged")
    frames = np.stack([natural(30, 140, seed=s) for s in (51, 52, 53)])
    out = io.BytesIO()
    n = raw_stream.stream(io.BytesIO(frames.tobytes()), out, 140, 30, scale=4, models=["r"], chunk=2, model_path=model_dir)
    assert n == 3
    eng = E.Engine.from_files(model_dir, "4x_Valar_v1", 0)
    assert np.array_equal(np.frombuffer(out.getvalue(), np.uint8).reshape(3, 120, 560, 3), np.stack([eng.run_u8(f) for f in frames]))
    eng.close()


def test_valar_rrdb_fused_tcgen05(E, model_dir, oracle_models):
    """4x_Valar_v1 (RRDB, reference models/4x_Valar_v1.param) on the fused tcgen05 graph kernels (b2sr_create_fused):
    golden crop, a two-tile frame against the oracle (seam at x = 960), band boundaries at 128 columns with a ragged
    last band, the float canvas, determinism, and agreement with the generic engine's fp32 CUDA-core kernels."""
    from upscale_video_b200 import ncnn_model
    if not os.path.exists(os.path.join(model_dir, "4x_Valar_v1.b2sr")):
        pytest.skip("4x_Valar_v1 not packaged")
    eng = E.Engine.from_files(model_dir, "4x_Valar_v1", 0)
    assert eng.fused and not eng.generic and eng.scale == 4
    g = golden("valar4x_crop")
    # 420 convolutions deep and without an input residual: fp16 activation storage (fp32 trunk, fp32 accumulation) puts a
    # few % of the u8 values on the other side of a rounding boundary (CPU emulation of the same storage: 1.3 %); never > 1 LSB
    out = eng.run_u8(g["x"])
    # 420 convolutions (1 head, 23 x 3 x (5 + the 1x1 shortcut), 1 trunk, 4 tail): every 1x1 shortcut rides on the launch
    # of the 3x3 convolution it is added to (-69); every 192 -> 64 convolution is ONE launch of 2-CTA clusters (its two
    # 32-channel halves share the input rows through TMA multicast)
    assert eng.stat(E.STAT_TC_LAUNCHES) == 420 - 69 and eng.stat(E.STAT_PIPE_LAUNCHES) == 0 and eng.stat(E.STAT_HMMA_LAUNCHES) == 0
    # the opt-in schedule: the 23 RRDBs as 23 persistent segment launches (15 convolutions = 18 stages x bands CTAs each,
    # dense-block buffers in L2-resident rings): head + 23 + trunk convolution + 4 tail convolutions, same bytes out
    eng.set_option(E.OPT_SEG_PIPE, 1)
    eng.reset_stats()
    piped = eng.run_u8(g["x"])
    assert eng.stat(E.STAT_PIPE_LAUNCHES) == 23 and eng.stat(E.STAT_TC_LAUNCHES) == 1 + 23 + 1 + 4
    assert np.array_equal(out, piped), "persistent segments and per-convolution launches must agree bit for bit"
    eng.set_option(E.OPT_SEG_PIPE, 0)
    assert_parity(out, g["y"], "valar golden (tcgen05)", max_mismatch=0.03)
    assert np.array_equal(out, eng.run_u8(g["x"]))
    img = natural(20, 980, seed=13)  # seam at x = 960
    models = oracle_models("4x_Valar_v1")
    ref = oracle.upscale_image_array(models, img, 4, "f32")
    out = eng.run_u8(img)
    assert_parity(out, ref, "valar 20x980 (tcgen05)", max_mismatch=0.03)
    canvas = eng.run_f32(img)
    assert np.array_equal(oracle.saturate_u8(canvas), out)
    img = natural(37, 300, seed=14)  # three bands, the last one 44 columns wide; CTA ranges cut inside bands
    assert_parity(eng.run_u8(img), oracle.upscale_image_array(models, img, 4, "f32"), "valar 37x300 (tcgen05)", max_mismatch=0.03)
    # worst case found for the storage types: salt-and-pepper extremes drive the largest activations through all 23 blocks
    # (CPU emulation of the device arithmetic: 0.70 LSB max float error with the fp32 trunk -- an all-fp16 trunk reaches
    # 1.17 LSB here and is why the trunk is kept in fp32, ncnn_model.compile_fused fp32_chain)
    rng = np.random.default_rng(5)
    rng.integers(0, 256, (40, 64, 3))  # (keeps the generator state of the CPU study this case comes from)
    harsh = np.where(rng.random((40, 64, 1)) > 0.5, 250, 5).astype(np.uint8).repeat(3, 2)
    assert_parity(eng.run_u8(harsh), oracle.upscale_image_array(models, harsh, 4, "f64"), "valar salt-and-pepper (tcgen05)", max_mismatch=0.10)
    # batch == single frame, bit for bit (accumulator homes follow plane rows, not CTA ranges)
    import torch
    frames = np.stack([natural(64, 200, seed=s) for s in (21, 22, 23)])
    d_in = torch.from_numpy(frames).cuda()
    d_out = torch.empty((3, 256, 800, 3), dtype=torch.uint8, device="cuda")
    eng.run_batch_device(d_in, d_out, 3, 64, 200, sync=True)
    for i in range(3):
        assert np.array_equal(d_out[i].cpu().numpy(), eng.run_u8(frames[i])), "valar frame %d: batch != single" % i
    # the two schedules on a batch with seams, ragged bands and several planes per pass, and with short rings (8 rows:
    # every stage is back-pressured all the time)
    frames = np.stack([natural(150, 1000, seed=s) for s in (31, 32)])
    d_in = torch.from_numpy(frames).cuda()
    d_a = torch.empty((2, 600, 4000, 3), dtype=torch.uint8, device="cuda")
    d_b, d_c = torch.empty_like(d_a), torch.empty_like(d_a)
    eng.set_option(E.OPT_SEG_PIPE, 1)
    eng.run_batch_device(d_in, d_a, 2, 150, 1000, sync=True)
    eng.set_option(E.OPT_RING_ROWS, 8)
    eng.run_batch_device(d_in, d_c, 2, 150, 1000, sync=True)
    eng.set_option(E.OPT_RING_ROWS, 0)
    eng.set_option(E.OPT_SEG_PIPE, 0)
    eng.run_batch_device(d_in, d_b, 2, 150, 1000, sync=True)
    assert torch.equal(d_a, d_b) and torch.equal(d_c, d_b), "valar: persistent segments != per-convolution launches"
    gen = E.Engine(ncnn_model.load_model(model_dir, "4x_Valar_v1"), 0, generic=True)
    gen.set_option(E.OPT_IMPL, E.IMPL_SIMPLE)
    d = np.abs(eng.run_u8(img).astype(int) - gen.run_u8(img).astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 0.05
    gen.close()
    eng.close()


def test_valar_cta_pair_schedule(E, model_dir, oracle_models):
    """The opt-in CTA-pair form of the dense-block convolutions (B2SR_OPT_PAIR2: 2-CTA clusters over band pairs,
    tcgen05.mma.cta_group::2 with M = 256, phantom rows above and below every CTA range): an odd band count (the last
    pair's second CTA has no columns), seams, several planes per pass, ranges cut inside bands.  The 32-channel convolutions
    accumulate in the same order as the default schedule; the 64-channel ones use a 6-block instead of an 8-block accumulator
    ring, so rows whose sum is split over an extension block differ in the last fp32 bit: the schedules agree within 1 LSB
    and both meet the oracle."""
    import torch
    if not os.path.exists(os.path.join(model_dir, "4x_Valar_v1.b2sr")):
        pytest.skip("4x_Valar_v1 not packaged")
    eng = E.Engine.from_files(model_dir, "4x_Valar_v1", 0)
    models = oracle_models("4x_Valar_v1")
    img = natural(37, 300, seed=14)  # three bands: one full pair + a pair whose second band is empty
    base = eng.run_u8(img)
    eng.set_option(E.OPT_PAIR2, 1)
    eng.reset_stats()
    pair = eng.run_u8(img)
    assert eng.stat(E.STAT_TC_LAUNCHES) == 420 - 69
    d = np.abs(pair.astype(int) - base.astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 0.06, "pair form vs default: max %d, %.3f%%" % (d.max(), 100 * (d > 0).mean())
    assert_parity(pair, oracle.upscale_image_array(models, img, 4, "f32"), "valar 37x300 (cta pairs)", max_mismatch=0.03)
    assert np.array_equal(pair, eng.run_u8(img)), "pair form is not deterministic"
    frames = np.stack([natural(150, 1000, seed=s) for s in (31, 32)])  # two tiles per frame (seam at 960), 8 + 1 bands
    d_in = torch.from_numpy(frames).cuda()
    d_a = torch.empty((2, 600, 4000, 3), dtype=torch.uint8, device="cuda")
    d_b = torch.empty_like(d_a)
    eng.run_batch_device(d_in, d_a, 2, 150, 1000, sync=True)
    eng.set_option(E.OPT_PAIR2, 0)
    eng.run_batch_device(d_in, d_b, 2, 150, 1000, sync=True)
    dd = (d_a.to(torch.int16) - d_b.to(torch.int16)).abs()
    assert int(dd.max()) <= 1 and float((dd > 0).float().mean()) < 0.06
    eng.close()


@pytest.mark.parametrize("shape", [(1, 1), (2, 3), (3, 129), (9, 128), (17, 257)])
def test_valar_edge_shapes(E, model_dir, oracle_models, shape):
    """Degenerate and band-boundary shapes through the fused tcgen05 graph kernels: 1-pixel planes (every tap but the
    centre is padding), widths at 128 / 129 / 257 columns (a band of one column), fewer rows than the accumulator ring."""
    if not os.path.exists(os.path.join(model_dir, "4x_Valar_v1.b2sr")):
        pytest.skip("4x_Valar_v1 not packaged")
    eng = E.Engine.from_files(model_dir, "4x_Valar_v1", 0)
    img = natural(shape[0], shape[1], seed=31 + shape[1])
    ref = oracle.upscale_image_array(oracle_models("4x_Valar_v1"), img, 4, "f64")
    assert_parity(eng.run_u8(img), ref, "valar %dx%d" % shape, max_mismatch=0.03)
    eng.close()


def test_valar_rrdb_generic_graph_engine(E, model_dir, oracle_models):
    """4x_Valar_v1 through the generic op-by-op graph engine (b2sr_create_graph; the cross-check implementation since
    the fused tcgen05 engine exists): golden crop on warp-level MMA and on fp32 CUDA cores."""
    from upscale_video_b200 import ncnn_model
    if not os.path.exists(os.path.join(model_dir, "4x_Valar_v1.b2sr")):
        pytest.skip("4x_Valar_v1 not packaged")
    eng = E.Engine(ncnn_model.load_model(model_dir, "4x_Valar_v1"), 0, generic=True)
    assert eng.generic and eng.scale == 4
    g = golden("valar4x_crop")
    out = eng.run_u8(g["x"])
    assert eng.stat(E.STAT_TC_LAUNCHES) == 0 and eng.stat(E.STAT_HMMA_LAUNCHES) >= 400
    assert_parity(out, g["y"], "valar golden (hmma)", max_mismatch=0.03)
    # fp32 CUDA-core kernels everywhere: essentially exact
    eng.set_option(E.OPT_IMPL, E.IMPL_SIMPLE)
    eng.reset_stats()
    out32 = eng.run_u8(g["x"])
    assert eng.stat(E.STAT_HMMA_LAUNCHES) == 0
    assert_parity(out32, g["y"], "valar golden (fp32)", max_mismatch=0.005)
    eng.close()


def test_compact_models_through_generic_engine(E, engines, model_dir, oracle_models):
    """The generic graph engine (fp32, CUDA cores) as a third device implementation of the Compact graphs."""
    from upscale_video_b200 import ncnn_model
    img = natural(50, 300, seed=17)
    for name, scale in (("2x_Compact_Pretrain", 2), (HURR, 1)):
        gen = E.Engine(ncnn_model.load_model(model_dir, name), 0, generic=True)
        gen.set_option(E.OPT_IMPL, E.IMPL_SIMPLE)  # fp32 CUDA-core kernels (the default would use HMMA for the 64->64 convolutions)
        tile = 960 if scale > 1 else 0
        a = gen.run_u8(img, tile=tile, halo=10 if tile else 0)
        b = engines(name).run_u8(img, tile=tile, halo=10 if tile else 0)
        d = np.abs(a.astype(int) - b.astype(int))
        assert d.max() <= 1 and (d > 0).mean() < 0.05, name
        ref = (oracle.upscale_image_array(oracle_models(name), img, scale, "f64") if scale > 1
               else oracle.apply_model_array(oracle_models(name), img, "f64"))
        assert np.abs(a.astype(int) - ref.astype(int)).max() <= 1 and (a != ref).mean() < 0.002, name  # fp32 path: almost exact
        gen.close()


def test_valar_540p_full_frame_vs_oracle(E, model_dir, oracle_models):
    """BASELINE configs[3] at its stated size: one 960x540 frame of natural-looking content through 4x_Valar_v1
    (reference models/4x_Valar_v1.param:1-1208, 420 convolutions, no input residual) against the f32 CPU oracle run on
    this box's host threads (~2 min on 16 threads); bar <= 1 LSB on all 24.9 M output values."""
    if not os.path.exists(os.path.join(model_dir, "4x_Valar_v1.b2sr")):
        pytest.skip("4x_Valar_v1 not packaged")
    staged = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "sample.png")
    if os.path.exists(staged):  # the reference's own fixture, when it was staged: a real photograph-like frame
        import cv2
        img = np.ascontiguousarray(cv2.imread(staged)[300:840, 400:1360])
        what = "sample.png[300:840, 400:1360]"
    else:
        img = natural(540, 960, seed=31)
        what = "natural(540, 960)"
    assert img.shape == (540, 960, 3)
    eng = E.Engine.from_files(model_dir, "4x_Valar_v1", 0)
    out = eng.run_u8(img)
    eng.close()
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    ref = oracle.upscale_image_array(oracle_models("4x_Valar_v1"), img, 4, "f32")
    d = np.abs(out.astype(np.int32) - ref.astype(np.int32))
    _record("valar_540p_full_frame", content=what, max_lsb=int(d.max()), mismatch_fraction=float((d > 0).mean()))
    assert out.shape == (2160, 3840, 3) and d.max() <= 1, "valar 540p %s: max |diff| = %d LSB" % (what, d.max())
    assert (d > 0).mean() <= 0.025, "valar 540p %s: %.2f%% of values differ" % (what, 100 * (d > 0).mean())


def test_persistent_grid_guard_falls_back(E, model_dir, oracle_models):
    """The persistent kernel's CTAs wait on one another, so it is launched cooperatively and only when layers x bands fits
    the SMs the context may use.  With the SM budget cut below 18 x 8 (B2SR_OPT_SM_LIMIT, what an MPS share or a second
    persistent kernel on the device amounts to) the pass must run layer by layer -- same bytes out -- not spin into a trap."""
    eng = E.Engine.from_files(model_dir, "2x_Compact_Pretrain", 0)
    img = natural(70, 1000, seed=41)
    a = eng.run_u8(img)
    assert eng.stat(E.STAT_PIPE_LAUNCHES) == 1 and eng.stat(E.STAT_PIPE_FALLBACKS) == 0
    eng.set_option(E.OPT_SM_LIMIT, 100)
    eng.reset_stats()
    b = eng.run_u8(img)
    assert eng.stat(E.STAT_PIPE_LAUNCHES) == 0 and eng.stat(E.STAT_PIPE_FALLBACKS) == 1 and eng.stat(E.STAT_TC_LAUNCHES) == 18
    assert np.array_equal(a, b)
    eng.set_option(E.OPT_IMPL, E.IMPL_PIPELINED)  # explicitly requested but not allowed to fit: an error item, not a hang
    with pytest.raises(E.EngineError, match="co-resident"):
        eng.run_u8(img)
    eng.set_option(E.OPT_IMPL, E.IMPL_AUTO)
    eng.set_option(E.OPT_SM_LIMIT, 0)
    eng.reset_stats()
    assert np.array_equal(eng.run_u8(img), a) and eng.stat(E.STAT_PIPE_LAUNCHES) == 1
    # two contexts on one device, both persistent, launched back to back from two host threads: cooperative launches
    # serialise instead of interleaving half-resident grids
    import threading
    other = E.Engine.from_files(model_dir, "2x_Compact_Pretrain", 0)
    res = [None, None]

    def work(i, e):
        for _ in range(4):
            res[i] = e.run_u8(img)
    ts = [threading.Thread(target=work, args=(0, eng)), threading.Thread(target=work, args=(1, other))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert np.array_equal(res[0], a) and np.array_equal(res[1], a)
    other.close()
    eng.close()


def test_raw_stream_multi_worker_identical(engines, model_dir, tmp_path):
    """raw_stream.stream_multi (`-g 0,0`: two worker threads with their own engines on GPU 0, one dynamic chunk queue, in-order
    writer) writes exactly the bytes of the single-GPU stream -- from a pipe and from a seekable file -- including the
    chained `-m a` mode."""
    import io
    from upscale_video_b200 import raw_stream
    frames = np.stack([natural(70, 1000, seed=s) for s in range(60, 67)])
    comp, hurr = engines("2x_Compact_Pretrain"), engines(HURR)
    expect = np.stack([comp.run_u8(f) for f in frames]).tobytes()
    out = io.BytesIO()
    assert raw_stream.stream_multi(io.BytesIO(frames.tobytes()), out, 1000, 70, scale=2, gpus=[0, 0], chunk=2, model_path=model_dir) == 7
    assert out.getvalue() == expect
    p = tmp_path / "in.raw"
    p.write_bytes(frames.tobytes())
    out = io.BytesIO()
    with open(p, "rb") as f:
        assert raw_stream.stream_multi(f, out, 1000, 70, scale=2, gpus=[0, 0, 0], chunk=2, model_path=model_dir) == 7
    assert out.getvalue() == expect
    out = io.BytesIO()
    raw_stream.stream_multi(io.BytesIO(frames.tobytes()), out, 1000, 70, scale=2, models=["a"], gpus=[0, 0], chunk=3, model_path=model_dir)
    assert out.getvalue() == np.stack([comp.run_u8(hurr.run_u8(f, tile=0, halo=0)) for f in frames]).tobytes()
