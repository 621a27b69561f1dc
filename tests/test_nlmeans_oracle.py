"""CPU: the denoise pass (`-m n=<level>`, reference upscale/upscale_processing.py:350-392).

* the oracle (oracle/nlmeans.py) is pinned against the reference's own dependency, cv2.fastNlMeansDenoisingColored
  -- exhaustively for the colour conversions, on seeded images for every level the reference accepts (1..30) --
  and against the vectors frozen from cv2 (tests/golden/nlm_*.npz, tools/make_nlm_goldens.py);
* the product's host-side tables (C ABI, no device needed) equal the oracle's;
* the kernel's warp algorithm, emulated lane by lane with the product's tables, equals the goldens;
* the host logic of apply_denoise / process_denoise (error protocol, file naming).
"""
import os

import numpy as np
import pytest

from conftest import golden
from oracle import nlmeans as N

NLM_GOLDENS = ["nlm_crop_l3", "nlm_crop_l10", "nlm_noisy_l5", "nlm_noise_l30", "nlm_tiny_l3", "nlm_1px_l1"]


def natural(h, w, seed):
    rng = np.random.default_rng(seed)
    base = np.linspace(20, 230, w)[None, :, None] * np.ones((h, 1, 3)) * np.array([1.0, 0.8, 0.6])
    base += 25 * np.sin(np.arange(h) / 3.0)[:, None, None]
    return np.clip(base + rng.normal(0, 6, (h, w, 3)), 0, 255).astype(np.uint8)


def test_lab_exhaustive():
    """Both colour conversions against cv2 over all 2**24 inputs."""
    cv2 = pytest.importorskip("cv2")
    v = np.arange(1 << 24, dtype=np.uint32)
    allc = np.stack([v & 255, (v >> 8) & 255, v >> 16], -1).astype(np.uint8).reshape(4096, 4096, 3)
    assert np.array_equal(N.bgr_to_lab(allc), cv2.cvtColor(allc, cv2.COLOR_LBGR2Lab))
    assert np.array_equal(N.lab_to_bgr(allc), cv2.cvtColor(allc, cv2.COLOR_Lab2LBGR))


@pytest.mark.parametrize("level", list(range(1, 31)))
def test_oracle_equals_cv2_every_level(level):
    """The call of reference apply_denoise (:354), every level the CLI lets through (test_images.py:45-52)."""
    cv2 = pytest.importorskip("cv2")
    img = natural(20 + level % 7, 24 + level % 5, seed=level) if level % 3 else \
        np.random.default_rng(level).integers(0, 256, (19, 23, 3), dtype=np.uint8)
    ref = cv2.fastNlMeansDenoisingColored(cv2.UMat(img), None, level, level, 5, 9).get()
    assert np.array_equal(N.fast_nl_means_denoising_colored(img, level, level), ref)


def test_oracle_planes_equal_cv2():
    """The two plane filters separately (1 and 2 channels), different h for luma and colour."""
    cv2 = pytest.importorskip("cv2")
    lab = N.bgr_to_lab(natural(40, 50, seed=2))
    for h in (1, 3, 7.5, 30):
        assert np.array_equal(N.fast_nl_means_denoising(lab[..., 0], h), cv2.fastNlMeansDenoising(lab[..., 0].copy(), None, h, 5, 9))
        ab = np.ascontiguousarray(lab[..., 1:3])
        assert np.array_equal(N.fast_nl_means_denoising(ab, h), cv2.fastNlMeansDenoising(ab, None, h, 5, 9))
    img = natural(30, 41, seed=3)
    assert np.array_equal(N.fast_nl_means_denoising_colored(img, 4, 9), cv2.fastNlMeansDenoisingColored(img, None, 4, 9, 5, 9))


def test_oracle_equals_cv2_random_shapes_and_levels():
    """Property test: any small image (including ones narrower than the 6-px border), any level pair."""
    cv2 = pytest.importorskip("cv2")
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")

    @hyp.settings(max_examples=60, deadline=None, derandomize=True)
    @hyp.given(st.integers(1, 20), st.integers(1, 20), st.integers(1, 30), st.integers(1, 30), st.integers(0, 2 ** 31 - 1),
               st.sampled_from(["noise", "smooth", "two-level"]))
    def check(h, w, level, level_color, seed, kind):
        rng = np.random.default_rng(seed)
        if kind == "noise":
            img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        elif kind == "smooth":
            img = np.clip(rng.integers(0, 256, 3)[None, None, :] + rng.normal(0, 4, (h, w, 3)), 0, 255).astype(np.uint8)
        else:
            img = (rng.integers(0, 2, (h, w, 1)) * 255).astype(np.uint8).repeat(3, axis=2)
        ref = cv2.fastNlMeansDenoisingColored(img, None, level, level_color, 5, 9)
        assert np.array_equal(N.fast_nl_means_denoising_colored(img, level, level_color), ref)

    check()


@pytest.mark.parametrize("name", NLM_GOLDENS)
def test_oracle_goldens(name):
    g = golden(name)
    assert np.array_equal(N.fast_nl_means_denoising_colored(g["x"], int(g["level"]), int(g["level"])), g["y"])


def test_product_tables_equal_oracle():
    """libb2sr builds its own tables in C++ (csrc/nlm_host.inl); they must be the oracle's, entry for entry."""
    from upscale_video_b200 import engine as E
    mine, ref = E.nlm_lab_tables(), N.lab_tables()
    for k in mine:
        assert np.array_equal(mine[k].reshape(-1), np.asarray(ref[k]).reshape(-1)), k
    for h in list(range(1, 31)) + [0.5, 2.5, 12.25]:
        for cn in (1, 2):
            assert np.array_equal(E.nlm_weight_table(h, cn), N.weight_table(h, cn)), (h, cn)
    with pytest.raises(E.EngineError):
        E.nlm_weight_table(0, 1)


@pytest.mark.parametrize("name", ["nlm_noise_l30", "nlm_tiny_l3", "nlm_1px_l1", "nlm_noisy_l5"])
def test_kernel_warp_algorithm_emulated(name):
    """nlm_kernel's tile / lane / sliding-window arithmetic (tests/nlm_warp_emulator.py) with the product's tables."""
    import nlm_warp_emulator as EM
    from upscale_video_b200 import engine as E
    g = golden(name)
    level = int(g["level"])
    tabs = []
    for cn in (1, 2):
        t = E.nlm_weight_table(level, cn)
        n = len(t)
        while n > 1 and t[n - 1] == 0:
            n -= 1
        tabs.append(t[:n].astype(np.int64))
    lab = EM.run(N.bgr_to_lab(g["x"]), tabs[0], tabs[1])
    assert np.array_equal(N.lab_to_bgr(lab), g["y"])


@pytest.mark.parametrize("th", [8, 12])
def test_kernel_other_tile_heights_emulated(th):
    """The experiment switch B2SR_NLM_TH (8 or 12 rows per warp tile instead of 16) changes only the tiling."""
    import nlm_warp_emulator as EM
    from upscale_video_b200 import engine as E
    g = golden("nlm_noise_l30")
    tabs = [E.nlm_weight_table(30, cn).astype(np.int64) for cn in (1, 2)]
    lab = EM.run(N.bgr_to_lab(g["x"]), tabs[0], tabs[1], packed=False, th=th)
    assert np.array_equal(N.lab_to_bgr(lab), g["y"])


def test_apply_denoise_reports_errors_as_items(tmp_path, monkeypatch):
    """Without a device the worker returns error items and leaves the input in place (never a silent CPU result)."""
    cv2 = pytest.importorskip("cv2")
    from upscale_video_b200 import engine as E
    from upscale_video_b200 import upscale_processing as up
    if E.device_count() > 0:
        pytest.skip("a CUDA device is visible: the error path is not reachable")
    monkeypatch.chdir(tmp_path)
    cv2.imwrite("1.extract.png", natural(8, 9, seed=1))
    items = up.apply_denoise("1.extract.png", "1.denoise.png", 3, True)
    assert [i[0] for i in items] == ["error", "error"] and items[0][1] == "Denoise failed"
    assert "no CUDA device" in str(items[1][1])
    assert os.path.exists("1.extract.png") and not os.path.exists("1.denoise.png")


def test_denoiser_has_no_cpu_path():
    from upscale_video_b200 import engine as E
    if E.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(E.EngineError, match="no CUDA device"):
        E.Denoiser(0)


def test_test_images_cli_accepts_denoise_option():
    """`-m n=K` is parsed like reference test_images.py:45-52 (clamped to 30, <= 0 disables)."""
    import inspect
    from upscale_video_b200 import test_images
    src = inspect.getsource(test_images.process_image)
    assert "process_denoise(input_frames, input_file_tag, denoise, remove=False)" in src and "min(int(denoise[0][1]), 30)" in src
