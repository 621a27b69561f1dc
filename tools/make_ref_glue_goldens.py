#!/usr/bin/env python
"""Golden vectors from the REFERENCE'S OWN CODE: /root/reference/upscale/upscale_processing.py is imported unmodified and its
`init_worker`, `upscale_image` (-> `process_tile`) and `apply_model` are run on PNG files, with only the two third-party modules it
cannot import here replaced:

  * `ncnn_vulkan.ncnn` -> a numpy stand-in for the five ncnn calls on the path (`Net.load_param/load_model`, `Mat.from_pixels`,
    `Mat.substract_mean_normalize`, `Extractor.input/extract`, `np.array(mat)`), whose network run is the oracle's layer
    interpreter in float32;
  * `wakepy` -> an empty module (imported at the top of the reference file, never used on this path).

So everything the reference itself contains on the path -- tile grid, the "halo only where >= 10 px remain" rule, the `.copy()` crop,
unswapped BGR, `1 / 255.0` normalisation, `* 255`, the float64 canvas, the crop of the halo out of every tile's output, cv2.imwrite's
rounding, model file naming `<scale><model_file>.{param,bin}` -- is executed from the reference's own lines, and the files it writes
are frozen as tests/golden/ref_glue.npz.  tests/test_oracle_golden.py::test_oracle_glue_equals_reference_code then requires the
oracle's restatement of that glue (oracle.upscale_image_array / apply_model_array, precision "f32") to reproduce them bit for bit.
What this does NOT pin is ncnn's layer arithmetic (the stand-in's network run is the oracle's): DESIGN.md section 1.

Run in the build container (needs /root/reference):   python tools/make_ref_glue_goldens.py
"""
import importlib.util
import os
import sys
import tempfile
import types

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

REF = "/root/reference"
HURR = "x_HurrDeblur_SubCompact_nf24-nc8_244k_net_g"


# ---- numpy stand-in for ncnn_vulkan.ncnn: only what upscale_processing.py:54-73, :265-281, :437-453 call ----
class _Mat:
    class PixelType:
        PIXEL_BGR = 2  # (from_pixels with PIXEL_BGR keeps the byte order it is given)

    def __init__(self, chw):
        self.chw = chw

    @staticmethod
    def from_pixels(pixels, pixel_type, w, h):
        assert pixel_type == _Mat.PixelType.PIXEL_BGR and pixels.dtype == np.uint8 and pixels.shape == (h, w, 3)
        return _Mat(np.ascontiguousarray(pixels.transpose(2, 0, 1)).astype(np.float32))  # planar float32, channel order unchanged

    def substract_mean_normalize(self, mean_vals, norm_vals):
        assert list(mean_vals) == [] and len(norm_vals) == 3
        for c in range(3):  # ncnn takes the factors as C floats
            self.chw[c] *= np.float32(norm_vals[c])

    def __array__(self, dtype=None, copy=None):
        return self.chw if dtype is None else self.chw.astype(dtype)


class _Extractor:
    def __init__(self, net):
        self.net, self.inputs = net, {}

    def input(self, name, mat):
        self.inputs[name] = mat

    def extract(self, name):
        (in_name, mat), = self.inputs.items()
        x = np.ascontiguousarray(mat.chw.transpose(1, 2, 0))
        y = oracle.run_graph(self.net.layers, x, "f32", input_name=in_name, output_name=name)
        return 0, _Mat(np.ascontiguousarray(y.astype(np.float32).transpose(2, 0, 1)))


class _Net:
    def __init__(self):
        self.opt = types.SimpleNamespace(use_vulkan_compute=False)
        self.param = self.layers = None
        self.device = None

    def set_vulkan_device(self, i):
        self.device = i

    def load_param(self, path):
        self.param = path

    def load_model(self, path):
        assert self.param is not None and os.path.splitext(self.param)[0] == os.path.splitext(path)[0]
        self.layers = oracle.read_ncnn(self.param, path)

    def create_extractor(self):
        return _Extractor(self)


def big_case_input():
    """The four-tile case's input is generated, not stored (the test calls this too); its output is stored as a digest plus the
    strips around both seams."""
    yy, xx = np.mgrid[0:975, 0:972]
    rng = np.random.default_rng(975972)
    base = np.stack([120 + 90 * np.sin(xx / 37.0) * np.cos(yy / 53.0), 40 + 0.2 * xx + 30 * ((xx // 16 + yy // 12) % 2), 200 - 0.15 * yy], -1)
    return np.clip(base + rng.normal(0, 6, base.shape), 0, 255).astype(np.uint8)


def import_reference():
    ncnn = types.SimpleNamespace(Net=_Net, Mat=_Mat, destroy_gpu_instance=lambda: None)
    pkg = types.ModuleType("ncnn_vulkan")
    pkg.ncnn = ncnn
    sys.modules["ncnn_vulkan"] = pkg
    wk = types.ModuleType("wakepy")
    wk.keep = types.SimpleNamespace()
    sys.modules["wakepy"] = wk
    spec = importlib.util.spec_from_file_location("ref_upscale_processing", os.path.join(REF, "upscale", "upscale_processing.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def pool_behaviour(sample):
    """The reference's `process_model` and `upscale_frames` themselves (spawn pools, one worker per -g entry, upscale_processing.py
    :302-347, :545-601) on a scratch directory of PNG frames with a gap: which files exist afterwards, what was logged.  The
    third-party modules are importable stubs here (tools/ref_stubs), so the spawned workers import the reference by its real name."""
    import importlib
    import logging
    stubs = os.path.join(ROOT, "tools", "ref_stubs")
    for pth in (REF, os.path.join(ROOT, "tools"), stubs):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    for m in ("ncnn_vulkan", "wakepy", "upscale", "upscale.upscale_processing"):
        sys.modules.pop(m, None)
    refpkg = importlib.import_module("upscale.upscale_processing")
    assert refpkg.__file__ == os.path.join(REF, "upscale", "upscale_processing.py")
    records = []

    class Collect(logging.Handler):
        def emit(self, record):
            records.append([record.levelname, record.getMessage()])

    root = logging.getLogger()
    handler, level = Collect(), root.level
    root.addHandler(handler)
    root.setLevel(logging.DEBUG)
    result = {}
    with tempfile.TemporaryDirectory() as tmp:
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            frames = {}
            for n in (1, 2, 4, 5):  # frame 3 is missing: finished by an earlier run (the resume contract)
                frames[n] = np.ascontiguousarray(sample[100 * n:100 * n + 10, 300:324])
                cv2.imwrite("%d.extract.png" % n, frames[n])
            workers = 0
            refpkg.process_model(5, os.path.join(REF, "models"), HURR, 1, "input", "output", "extract", "anime", [0, 0], workers)
            result["after_process_model"] = sorted(os.listdir("."))
            result["log_process_model"] = sorted(records)
            del records[:]
            workers += 2
            refpkg.upscale_frames(2, 1, 5, "anime", 2, [0, 0], workers, os.path.join(REF, "models"), "x_Compact_Pretrain", "input", "output")
            result["after_upscale_frames"] = sorted(os.listdir("."))
            result["log_upscale_frames"] = sorted(records)
            del records[:]
            for n in frames:
                assert cv2.imread("%d.png" % n).shape == (20, 48, 3)
        finally:
            os.chdir(cwd)
            root.removeHandler(handler)
            root.setLevel(level)
    print("pool behaviour:", result)
    return result


def main():
    ref = import_reference()
    import multiprocessing
    multiprocessing.current_process()._identity = (1,)  # init_worker :59 derives the GPU slot from the pool worker's identity
    sample = cv2.imread(os.path.join(REF, "sample.png"))
    rng = np.random.default_rng(11)
    noise = lambda h, w: rng.integers(0, 256, (h, w, 3), dtype=np.uint8)  # noqa: E731
    # (name, model_file, scale, input image, which reference function)
    cases = [
        # tile geometry on the cheap 24-channel network: seams at 960 in x and y, the ">= 10 px remain" rule on both sides of it
        ("hurr_upscale_14x1000", HURR, 1, sample[300:314, 700:1700], "upscale_image"),
        ("hurr_upscale_1000x12", HURR, 1, sample[100:1100, 900:912], "upscale_image"),
        ("hurr_upscale_12x965", HURR, 1, noise(12, 965), "upscale_image"),    # 960 <= 965 - 10 is false: no right halo on tile 0
        ("hurr_upscale_12x969", HURR, 1, noise(12, 969), "upscale_image"),
        ("hurr_upscale_12x970", HURR, 1, noise(12, 970), "upscale_image"),    # exactly 10 px remain: halo
        ("hurr_upscale_12x971", HURR, 1, noise(12, 971), "upscale_image"),
        ("hurr_upscale_969x11", HURR, 1, noise(969, 11), "upscale_image"),
        ("hurr_upscale_970x11", HURR, 1, noise(970, 11), "upscale_image"),
        ("hurr_upscale_9x9", HURR, 1, noise(9, 9), "upscale_image"),          # smaller than the halo
        ("hurr_upscale_975x972", HURR, 1, big_case_input(), "upscale_image"),  # four tiles, three of them slivers
        ("hurr_apply_40x60", HURR, 1, sample[500:540, 900:960], "apply_model"),  # the untiled pre-pass, u8 through imwrite
        # the upscalers: crop of every tile's halo at scale 2 and 4
        ("compact2x_10x980", "x_Compact_Pretrain", 2, sample[640:650, 500:1480], "upscale_image"),
        ("compact2x_972x8", "x_Compact_Pretrain", 2, sample[150:1122, 1000:1008], "upscale_image"),
        ("compact4x_8x975", "x_Compact_Pretrain", 4, sample[700:708, 300:1275], "upscale_image"),
        ("valar4x_5x964", "x_Valar_v1", 4, sample[400:405, 200:1164], "upscale_image"),  # RRDB: no halo on tile 0's right, sliver tile
    ]
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, model_file, scale, img, fn in cases:
            ref.init_worker([0], 0, os.path.join(REF, "models"), model_file, scale, "input", "output")
            src, dst = os.path.join(tmp, name + ".in.png"), os.path.join(tmp, name + ".out.png")
            cv2.imwrite(src, img)
            if fn == "upscale_image":
                items = ref.upscale_image(src, dst, scale, None, 1, 1, remove=True)
            else:
                items = ref.apply_model(src, dst, True)
            assert not any(level == "error" for level, _ in items), items
            assert not os.path.exists(src), "the reference removes its input"
            got = cv2.imread(dst)
            assert got.shape == (img.shape[0] * scale, img.shape[1] * scale, 3)
            if name == "hurr_upscale_975x972":
                import hashlib
                out[name + "__sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(got).tobytes()).digest(), np.uint8)
                out[name + "__rows_940_975"] = np.ascontiguousarray(got[940:975])
                out[name + "__cols_940_972"] = np.ascontiguousarray(got[:, 940:972])
            else:
                out[name + "__x"] = np.ascontiguousarray(img)
                out[name + "__y"] = got
            print("%-26s %s -> %s  (%s, scale %d, %s)" % (name, img.shape, got.shape, fn, scale, model_file), flush=True)
    # the log items the reference's functions return (the result protocol of logging_callback, upscale_processing.py:40-51):
    # every frame_batch form of upscale_image :521-540 and apply_model :298, with relative file names in a scratch directory
    import json
    logs = {}
    img = sample[600:612, 500:1480]  # two tiles
    with tempfile.TemporaryDirectory() as tmp:
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            ref.init_worker([0], 0, os.path.join(REF, "models"), "x_Compact_Pretrain", 2, "input", "output")
            for key, frame_batch, out_name in (("batch_none", None, "7.png"), ("batch_int", 3, "7.png"), ("batch_list", [7, 9], "7.png"),
                                               ("no_output_file", None, None)):
                cv2.imwrite("7.extract.png", img)
                logs[key] = ref.upscale_image("7.extract.png", out_name, 2, frame_batch, 7, 12, remove=key != "no_output_file")
            ref.init_worker([0], 0, os.path.join(REF, "models"), HURR, 1, "input", "output")
            cv2.imwrite("7.extract.png", img)
            logs["apply_model"] = ref.apply_model("7.extract.png", "7.anime.png", True)
        finally:
            os.chdir(cwd)
    logs["pool"] = pool_behaviour(sample)
    out["log_items_input"] = np.ascontiguousarray(img)
    out["log_items_json"] = np.frombuffer(json.dumps(logs).encode(), np.uint8)
    print("log items:", json.dumps(logs)[:300], "...")
    path = os.path.join(ROOT, "tests", "golden", "ref_glue.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
