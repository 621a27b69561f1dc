"""GPU (-m gpu): the reference's OWN command-line tools, byte for byte, on this engine.

`__graft_entry__.build()` stages `/root/reference/test_gpus.py`, `test_images.py`, `sample.png` and the small original
model files (`.param` / `.bin`) unmodified under `baseline/_ref/` (git-ignored; it travels to the GPU box with the
snapshot).  With `upscale_video_b200/compat` in front of PYTHONPATH those scripts import
`upscale.upscale_processing` = the B200 drop-in and `ncnn_vulkan.ncnn` = the device-enumeration stub, run in a fresh
interpreter like a user would run them, and the PNGs they write are compared with the oracle.  Reference call sites
exercised: test_gpus.py:38-112 (device listing, spawn pool of init_worker, `runs` x upscale_image on sample.png),
test_images.py:18-159 (process_denoise -> process_model -> upscale_frames, renames), and through them the reference's
ORIGINAL ncnn model files read by the product loader.

Also here: the repo's own port of the calibrator, `python -m upscale_video_b200.test_gpus -g 0,0 -r 4`."""
import hashlib
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import HURR, ROOT
from oracle import nlmeans as N
from oracle import oracle
from test_nlmeans_oracle import natural

pytestmark = pytest.mark.gpu

REF = os.path.join(ROOT, "baseline", "_ref")
COMPAT = os.path.join(ROOT, "upscale_video_b200", "compat")


def _env():
    return dict(os.environ, PYTHONPATH=os.pathsep.join([COMPAT, ROOT, os.environ.get("PYTHONPATH", "")]))


def _staged(name):
    p = os.path.join(REF, name)
    if not os.path.exists(p):
        pytest.skip("baseline/_ref/%s is not staged (build() stages it where /root/reference exists)" % name)
    return p


def test_staged_scripts_are_the_reference_files_unmodified():
    """When the reference tree is present (the build container) the staged copies must be byte-identical to it."""
    for name in ("test_gpus.py", "test_images.py"):
        p = _staged(name)
        src = os.path.join("/root/reference", name)
        if os.path.exists(src):
            assert hashlib.sha256(open(p, "rb").read()).digest() == hashlib.sha256(open(src, "rb").read()).digest()
        text = open(p).read()
        assert "from upscale.upscale_processing import" in text and "upscale_video_b200" not in text


def test_reference_test_gpus_unmodified():
    """`python test_gpus.py -g 0,0 -s 2 -r 4` (reference README.md:39-49): two workers on GPU 0, four timed frames."""
    script = _staged("test_gpus.py")
    _staged("sample.png")
    r = subprocess.run([sys.executable, script, "-g", "0,0", "-s", "2", "-r", "4"], cwd=REF, env=_env(), capture_output=True, text=True,
                       timeout=600)
    log = r.stdout
    assert r.returncode == 0, log[-3000:] + r.stderr[-3000:]
    assert re.search(r"GPU count: [1-9]", log) and "Default GPU: 0" in log and "GPU 0: Discrete / NVIDIA" in log
    assert log.count("Testing GPU: 0") == 4 and log.count("seconds to upscale sample.png") == 4, log[-3000:]
    assert "seconds total to run tests." in log and "[ERROR]" not in log, log[-3000:]


def test_reference_test_images_unmodified(tmp_path, oracle_models):
    """`python test_images.py -i 5 -t tmp -o out -s 2` and `-i 6 -m n=3,a` (reference README.md:65-78), reading the
    reference's original 2x_Compact_Pretrain / HurrDeblur .param + .bin files from the directory beside the script."""
    import cv2
    script = _staged("test_images.py")
    _staged(os.path.join("models", "2x_Compact_Pretrain.bin"))
    tmp, out = tmp_path / "tmp", tmp_path / "out"
    (tmp / "upscale_video").mkdir(parents=True)
    out.mkdir()
    img = natural(200, 1000, seed=5)  # crosses the x = 960 tile seam
    small = natural(80, 112, seed=6)
    cv2.imwrite(str(tmp / "upscale_video" / "5.extract.png"), img)
    cv2.imwrite(str(tmp / "upscale_video" / "6.extract.png"), small)

    def run(frames, models=None):
        cmd = [sys.executable, script, "-i", frames, "-t", str(tmp), "-o", str(out), "-s", "2", "-g", "0"]
        if models:
            cmd += ["-m", models]
        r = subprocess.run(cmd, cwd=str(tmp_path), env=_env(), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "Completed" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
        assert "[ERROR]" not in r.stdout

    run("5")
    got = cv2.imread(str(out / "5.png"))
    ref = oracle.upscale_image_array(oracle_models("2x_Compact_Pretrain"), img, 2, "f64")
    d = np.abs(got.astype(np.int32) - ref.astype(np.int32))
    assert got.shape == (400, 2000, 3) and d.max() <= 1 and (d > 0).mean() < 0.06
    assert (out / "5.extract.png").exists()

    run("6", "n=3,a")
    den = N.fast_nl_means_denoising_colored(small, 3, 3)
    assert np.array_equal(cv2.imread(str(out / "6.denoise.png")), den)  # bit-exact vs cv2's algorithm
    anime = cv2.imread(str(out / "6.anime.png"))
    d = np.abs(anime.astype(np.int32) - oracle.apply_model_array(oracle_models(HURR), den, "f64").astype(np.int32))
    assert d.max() <= 1
    got = cv2.imread(str(out / "6.n=3.a.png"))
    ref = oracle.upscale_image_array(oracle_models("2x_Compact_Pretrain"), anime, 2, "f64")
    d = np.abs(got.astype(np.int32) - ref.astype(np.int32))
    assert got.shape == (160, 224, 3) and d.max() <= 1 and (d > 0).mean() < 0.06


def test_repo_test_gpus_module(tmp_path):
    """The repo's port of the calibrator (SURVEY 8a row `test_gpus.py` loop): `-g 0,0 -r 4` on a synthetic frame and on
    a PNG given with -i; same log lines as the reference's tool."""
    import cv2
    png = tmp_path / "frame.png"
    cv2.imwrite(str(png), natural(300, 1000, seed=9))
    for extra in ([], ["-i", str(png)]):
        r = subprocess.run([sys.executable, "-m", "upscale_video_b200.test_gpus", "-g", "0,0", "-s", "2", "-r", "4"] + extra, cwd=ROOT,
                           env=dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", "")), capture_output=True,
                           text=True, timeout=600)
        log = r.stdout
        assert r.returncode == 0, log[-3000:] + r.stderr[-3000:]
        assert re.search(r"GPU count: [1-9]", log) and "GPU 0: Discrete / NVIDIA" in log
        assert log.count("Testing GPU: 0") == 4 and len(re.findall(r"[0-9.]+ seconds to upscale ", log)) == 4, log[-3000:]
        assert "seconds total to run tests." in log and "[ERROR]" not in log
