"""GPU (-m gpu): the denoise pass through the C ABI (b2sr_nlm_*, include/b2sr.h) against the oracle, the vectors
frozen from cv2, and -- where cv2 is importable on the box -- against cv2.fastNlMeansDenoisingColored itself,
called exactly as reference apply_denoise does (upscale/upscale_processing.py:352-354).  Bar: bit-exact (integer path).
"""
import os

import numpy as np
import pytest

from conftest import golden
from oracle import nlmeans as N
from test_nlmeans_oracle import NLM_GOLDENS, natural

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E():
    from upscale_video_b200 import engine
    assert engine.device_count() >= 1, "no CUDA device: the product path has no CPU fallback"
    return engine


@pytest.fixture(scope="module")
def dn(E):
    d = E.Denoiser(0)
    yield d
    d.close()


@pytest.mark.parametrize("name", NLM_GOLDENS)
def test_goldens_bit_exact(name, dn):
    g = golden(name)
    got = dn.run_u8(g["x"], int(g["level"]))
    assert np.array_equal(got, g["y"]), "%s: %d values differ" % (name, (got != g["y"]).sum())
    assert dn.launches >= 1


@pytest.mark.parametrize("h,w", [(1, 1), (1, 40), (2, 2), (16, 28), (17, 29), (15, 27), (100, 3), (6, 57), (7, 7), (33, 85)])
def test_edge_shapes_vs_oracle(h, w, dn):
    """One tile exactly (16 x 28), one pixel over / under, images smaller than the 6-px border, ragged tiles."""
    level = 1 + (h * 7 + w) % 12
    img = natural(h, w, seed=h * 1000 + w) if (h + w) % 2 else np.random.default_rng(h + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    assert np.array_equal(dn.run_u8(img, level), N.fast_nl_means_denoising_colored(img, level, level)), (h, w, level)


def test_separate_luma_and_colour_levels(dn):
    img = natural(40, 60, seed=5)
    assert np.array_equal(dn.run_u8(img, 4, 9), N.fast_nl_means_denoising_colored(img, 4, 9))
    assert np.array_equal(dn.run_u8(img, 2.5, 0.75), N.fast_nl_means_denoising_colored(img, 2.5, 0.75))
    # going back to earlier levels rebuilds the tables
    assert np.array_equal(dn.run_u8(img, 4, 9), N.fast_nl_means_denoising_colored(img, 4, 9))


def test_both_kernel_variants_mid_size(dn):
    """Levels 1..6 run the packed-shuffle variant (weight tables <= 409 entries), higher levels the generic one; a
    high-contrast image makes the packed variant's clamp bite.  Checked against cv2 itself when importable."""
    rng = np.random.default_rng(21)
    img = natural(150, 260, seed=21)
    img[::7] = 0
    img[3::7] = 255
    img[40:90, 100:180] = rng.integers(0, 256, (50, 80, 3), dtype=np.uint8)
    try:
        import cv2
        ref = lambda lv: cv2.fastNlMeansDenoisingColored(cv2.UMat(img), None, lv, lv, 5, 9).get()
    except ImportError:
        ref = lambda lv: N.fast_nl_means_denoising_colored(img, lv, lv)
    for level in (1, 6, 7, 12, 30):
        got = dn.run_u8(img, level)
        want = ref(level)
        assert np.array_equal(got, want), "level %d: %d values differ" % (level, (got != want).sum())


def test_constant_image_is_the_lab_round_trip(dn):
    """All template distances are 0, so every weight is equal: the result is Lab2LBGR(LBGR2Lab(x))."""
    img = np.empty((37, 59, 3), np.uint8)
    img[:] = (13, 200, 97)
    assert np.array_equal(dn.run_u8(img, 3), N.lab_to_bgr(N.bgr_to_lab(img)))


def test_full_1080p_against_cv2_and_properties(E, dn):
    """BASELINE-size frame: bit-exact against cv2 itself (or the oracle on crops when cv2 is absent), plus the
    size-independent properties: determinism, batch == single, strided device buffers == packed host buffers,
    interior of a crop == crop of the full frame (the filter looks 6 px around a pixel)."""
    import torch
    rng = np.random.default_rng(11)
    img = natural(1080, 1920, seed=11)
    img[200:400, 300:700] = rng.integers(0, 256, (200, 400, 3), dtype=np.uint8)  # a saturating-noise patch
    got = dn.run_u8(img, 3)
    try:
        import cv2
        ref = cv2.fastNlMeansDenoisingColored(cv2.UMat(img), None, 3, 3, 5, 9).get()
        assert np.array_equal(got, ref), "%d values differ from cv2" % (got != ref).sum()
    except ImportError:
        for (y, x) in [(0, 0), (1080 - 70, 1920 - 90), (500, 900)]:
            ref = N.fast_nl_means_denoising_colored(img[y:y + 70, x:x + 90], 3, 3)
            assert np.array_equal(got[y + 6:y + 64, x + 6:x + 84], ref[6:-6, 6:-6])
    assert np.array_equal(got, dn.run_u8(img, 3))
    # crop interior
    crop = dn.run_u8(img[300:420, 500:640], 3)
    assert np.array_equal(crop[6:-6, 6:-6], got[306:414, 506:634])
    # device-resident batch of 3 frames: frame 1 is `img`, the others differ
    frames = np.stack([np.roll(img, 17, axis=1), img, img[::-1].copy()])
    d_in = torch.from_numpy(frames).cuda()
    d_out = torch.empty_like(d_in)
    dn.run_batch_device(d_in, d_out, 3, 1080, 1920, 3, sync=True)
    out = d_out.cpu().numpy()
    assert np.array_equal(out[1], got)
    assert np.array_equal(out[2], got[::-1])  # reflect-101 borders commute with a vertical flip
    assert np.array_equal(out[0], dn.run_u8(frames[0], 3))
    # host batch pipeline (more frames than one staging chunk holds)
    n = 19  # 8 + 8 + 3: the third chunk reuses staging slot 0
    h_in = torch.from_numpy(np.stack([frames[i % 3] for i in range(n)])).pin_memory()
    h_out = torch.empty_like(h_in).pin_memory()
    dn.run_batch_host(h_in, h_out, n, 1080, 1920, 3)
    for i in range(n):
        assert np.array_equal(h_out[i].numpy(), out[i % 3]), i


def test_strided_rows_and_device_pointers(E, dn):
    import ctypes
    import torch
    img = natural(45, 70, seed=9)
    ref = N.fast_nl_means_denoising_colored(img, 5, 5)
    lib = E.load_library()
    # padded rows on the host
    src = np.zeros((45, 70 * 3 + 13), np.uint8)
    src[:, :210] = img.reshape(45, 210)
    dst = np.full((45, 70 * 3 + 5), 7, np.uint8)
    rc = lib.b2sr_nlm_run_u8(dn._h, src.ctypes.data, 45, 70, src.shape[1], dst.ctypes.data, dst.shape[1], 5.0, 5.0, 5, 9, E.MEM_HOST)
    assert rc == 0, lib.b2sr_last_error()
    assert np.array_equal(dst[:, :210].reshape(45, 70, 3), ref) and (dst[:, 210:] == 7).all()
    # device pointers with padded rows
    d_src = torch.from_numpy(src).cuda()
    d_dst = torch.full((45, 70 * 3 + 5), 9, dtype=torch.uint8, device="cuda")
    rc = lib.b2sr_nlm_run_u8(dn._h, ctypes.c_void_p(d_src.data_ptr()), 45, 70, src.shape[1], ctypes.c_void_p(d_dst.data_ptr()),
                             dst.shape[1], 5.0, 5.0, 5, 9, E.MEM_DEVICE)
    assert rc == 0, lib.b2sr_last_error()
    out = d_dst.cpu().numpy()
    assert np.array_equal(out[:, :210].reshape(45, 70, 3), ref) and (out[:, 210:] == 9).all()


def test_errors_are_codes_not_crashes(E, dn):
    lib = E.load_library()
    img = natural(8, 8, seed=1)
    out = np.empty_like(img)
    args = (dn._h, img.ctypes.data, 8, 8, 24, out.ctypes.data, 24)
    assert lib.b2sr_nlm_run_u8(*args, 3.0, 3.0, 7, 21, E.MEM_HOST) == -5 and b"5 / 9" in lib.b2sr_last_error()  # cv2's defaults, not the reference's
    assert lib.b2sr_nlm_run_u8(*args, 0.0, 3.0, 5, 9, E.MEM_HOST) == -1
    assert lib.b2sr_nlm_run_u8(dn._h, img.ctypes.data, 8, 8, 20, out.ctypes.data, 24, 3.0, 3.0, 5, 9, E.MEM_HOST) == -1
    assert lib.b2sr_nlm_run_u8(dn._h, None, 8, 8, 24, out.ctypes.data, 24, 3.0, 3.0, 5, 9, E.MEM_HOST) == -1
    assert lib.b2sr_nlm_run_u8(dn._h, img.ctypes.data, 8, 8, 24, img.ctypes.data, 24, 3.0, 3.0, 5, 9, E.MEM_HOST) == -1  # in place
    assert b"overlap" in lib.b2sr_last_error()
    with pytest.raises(E.EngineError):
        E.Denoiser(99)
    assert np.array_equal(dn.run_u8(img, 3), N.fast_nl_means_denoising_colored(img, 3, 3))  # still usable


def test_worker_functions_on_png_files(tmp_path, monkeypatch):
    """apply_denoise / process_denoise on PNG files like the reference's callers (test_images.py:82-87,
    fix_frames.py:211-214): N.<tag>.png -> N.denoise.png, inputs deleted when `remove`, missing inputs skipped, the
    pool size returned."""
    import cv2
    from upscale_video_b200 import upscale_processing as up
    monkeypatch.chdir(tmp_path)
    frames = {n: natural(30 + n, 50 + 3 * n, seed=n) for n in (1, 2, 4)}
    for n, im in frames.items():
        cv2.imwrite("%d.extract.png" % n, im)
    items = up.apply_denoise("1.extract.png", "1.denoise.png", 3, False)
    assert items == [["info", "Processed Denoise: 1.denoise.png"]] and os.path.exists("1.extract.png")
    assert np.array_equal(cv2.imread("1.denoise.png"), N.fast_nl_means_denoising_colored(frames[1], 3, 3))
    os.remove("1.denoise.png")
    if up.denoiser is not None:
        up.denoiser.close()
        up.denoiser = None
    procs = up.process_denoise(4, "extract", 6, remove=True)
    assert procs == min(os.cpu_count() or 1, up.DENOISE_MAX_WORKERS)
    assert not os.path.exists("3.denoise.png")
    for n, im in frames.items():
        assert not os.path.exists("%d.extract.png" % n)
        assert np.array_equal(cv2.imread("%d.denoise.png" % n), N.fast_nl_means_denoising_colored(im, 6, 6)), n


def test_raw_stream_denoise_chain(dn, tmp_path):
    """`raw_stream -m a,n=4 -s 2`: denoise -> HurrDeblur -> 2x_Compact behind the raw-frame pipe equals the per-frame worker
    functions applied in the reference's order (test_images.py:82-144); `-s 1 -m n=4` is the denoise pass alone."""
    import io
    from conftest import HURR
    from upscale_video_b200 import engine as E
    from upscale_video_b200 import ncnn_model, raw_stream
    mdir = ncnn_model.packaged_model_dir()
    frames = np.stack([natural(40, 300, seed=s) for s in (1, 2, 3)])
    den = np.stack([N.fast_nl_means_denoising_colored(f, 4, 4) for f in frames])
    out = io.BytesIO()
    assert raw_stream.stream(io.BytesIO(frames.tobytes()), out, 300, 40, scale=1, models=["n=4"], chunk=2) == 3
    assert out.getvalue() == den.tobytes()
    hurr, comp = E.Engine.from_files(mdir, HURR), E.Engine.from_files(mdir, "2x_Compact_Pretrain")
    expect = np.stack([comp.run_u8(hurr.run_u8(f, tile=0, halo=0)) for f in den])
    out = io.BytesIO()
    assert raw_stream.stream(io.BytesIO(frames.tobytes()), out, 300, 40, scale=2, models=["a", "n=4"], chunk=2, model_path=mdir) == 3
    assert np.array_equal(np.frombuffer(out.getvalue(), np.uint8).reshape(3, 80, 600, 3), expect)
    hurr.close()
    comp.close()
