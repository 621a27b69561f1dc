"""GPU (-m gpu): the reference's test_images.py command line, one fresh interpreter per invocation like the real CLI
(the `workers_used` bookkeeping of reference init_worker :59 assumes the parent has spawned nothing before)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import HURR, ROOT
from oracle import nlmeans as N
from oracle import oracle
from test_nlmeans_oracle import natural

pytestmark = pytest.mark.gpu


def run_cli(tmp, out, frames, scale, models=None):
    cmd = [sys.executable, "-m", "upscale_video_b200.test_images", "-i", frames, "-t", str(tmp), "-o", str(out), "-s", str(scale), "-g", "0"]
    if models:
        cmd += ["-m", models]
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "Completed" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_cli_test_images_540p_config0(tmp_path, oracle_models):
    """BASELINE configs[0]: one 540p PNG frame through `test_images.py -s 2` (reference test_images.py:18-159), then a
    small frame with `-m n=3,a` (denoise -> anime -> upscale, result renamed 2.n=3.a.png, intermediates kept)."""
    import cv2
    tmp, out = tmp_path / "tmp", tmp_path / "out"
    (tmp / "upscale_video").mkdir(parents=True)
    out.mkdir()
    img = natural(540, 960, seed=77)
    crop = img[:96, :128].copy()
    cv2.imwrite(str(tmp / "upscale_video" / "1.extract.png"), img)
    cv2.imwrite(str(tmp / "upscale_video" / "2.extract.png"), crop)
    run_cli(tmp, out, "1", 2)
    got = cv2.imread(str(out / "1.png"))
    ref = oracle.upscale_image_array(oracle_models("2x_Compact_Pretrain"), img, 2, "f64")
    d = np.abs(got.astype(np.int32) - ref.astype(np.int32))
    assert got.shape == (1080, 1920, 3) and d.max() <= 1 and (d > 0).mean() < 0.06
    assert (out / "1.extract.png").exists()  # remove=False on this path
    run_cli(tmp, out, "2", 2, "n=3,a")
    got = cv2.imread(str(out / "2.n=3.a.png"))
    den = N.fast_nl_means_denoising_colored(crop, 3, 3)
    assert np.array_equal(cv2.imread(str(out / "2.denoise.png")), den)
    # stage by stage against the oracle applied to the file the previous stage actually wrote (u8 hops in between)
    anime = cv2.imread(str(out / "2.anime.png"))
    d = np.abs(anime.astype(np.int32) - oracle.apply_model_array(oracle_models(HURR), den, "f64").astype(np.int32))
    assert d.max() <= 1
    ref = oracle.upscale_image_array(oracle_models("2x_Compact_Pretrain"), anime, 2, "f64")
    d = np.abs(got.astype(np.int32) - ref.astype(np.int32))
    assert got.shape == (192, 256, 3) and d.max() <= 1 and (d > 0).mean() < 0.06
