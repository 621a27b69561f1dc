"""CPU: host-side logic of the drop-in worker module (upscale_video_b200/upscale_processing.py) and the
multi-process plumbing (gloo, world_size 2)."""
import logging
import os
import sys

import numpy as np
import pytest

from oracle import oracle
from upscale_video_b200 import parallel
from upscale_video_b200 import upscale_processing as up

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_get_frames():
    assert up.get_frames("1,4-6,9") == [1, 4, 5, 6, 9]
    assert up.get_frames("7") == [7]


@pytest.mark.parametrize("h,w", [(1080, 1920), (540, 960), (1278, 1920), (24, 1000), (969, 961), (1930, 975), (5, 5), (960, 960), (970, 1940)])
def test_tile_rect_matches_oracle(h, w):
    """tile_rect restates reference process_tile :398-427; the oracle restates it independently."""
    import math
    rects = {(r[0], r[1]): r[2:] for r in oracle.tile_rects(h, w)}
    for y in range(math.ceil(h / 960)):
        for x in range(math.ceil(w / 960)):
            (iy0, iy1, ix0, ix1), (cy0, cy1, cx0, cx1) = up.tile_rect(y, x, 960, h, w)
            assert (iy0, iy1, ix0, ix1, cy0, cy1, cx0, cx1) == rects[(y, x)]
    assert len(rects) == math.ceil(h / 960) * math.ceil(w / 960)


def test_logging_callback_exits_on_error(caplog):
    with caplog.at_level(logging.DEBUG):
        up.logging_callback([["debug", "d"], ["info", "i"]])
    assert "i" in caplog.text
    with pytest.raises(SystemExit):
        up.logging_callback([["info", "ok"], ["error", "boom"], ["info", "never"]])


def test_cli_fails_loudly_without_a_device(tmp_path):
    """No GPU: the worker initialiser must not throw the Pool into a respawn loop and the error item must end the parent
    (the reference's sys.exit inside the Pool callback would leave pool.join() hanging on current CPython)."""
    import subprocess
    import sys
    import cv2
    from upscale_video_b200 import engine
    if engine.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    (tmp_path / "t" / "upscale_video").mkdir(parents=True)
    (tmp_path / "o").mkdir()
    cv2.imwrite(str(tmp_path / "t" / "upscale_video" / "1.extract.png"), np.zeros((20, 30, 3), np.uint8))
    for models in (None, "n=3"):
        cmd = [sys.executable, "-m", "upscale_video_b200.test_images", "-i", "1", "-t", str(tmp_path / "t"), "-o", str(tmp_path / "o"),
               "-s", "2", "-g", "0"] + (["-m", models] if models else [])
        r = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=120)
        assert r.returncode != 0 and "Error - Exiting" in r.stderr and "no CUDA device" in (r.stdout + r.stderr), r.stdout + r.stderr
        assert not (tmp_path / "o" / "1.png").exists()


def test_logging_callback_on_pool_thread_defers_exit():
    import threading
    seen = []

    def handler():  # what the Pool's result-handler thread does
        try:
            up.logging_callback([["error", "boom"]])
            seen.append("returned")
        except SystemExit:
            seen.append("exit")

    t = threading.Thread(target=handler)
    t.start()
    t.join()
    assert seen == ["returned"] and up._failed
    with pytest.raises(SystemExit):
        up._exit_if_failed()
    assert not up._failed
    up._exit_if_failed()  # nothing pending: no exit


class _FakeNet:
    """Stands in for the engine in host-logic tests only (no arithmetic is checked with it)."""
    scale = 2

    def __init__(self, fail=False):
        self.fail, self.calls, self.closed = fail, [], False

    def run_u8(self, img, tile=960, halo=10):
        self.calls.append(("u8", img.shape, tile, halo))
        if self.fail:
            raise RuntimeError("device lost")
        return np.zeros((img.shape[0] * self.scale, img.shape[1] * self.scale, 3), np.uint8)

    def run_f32(self, img, tile=960, halo=10):
        self.calls.append(("f32", img.shape, tile, halo))
        if self.fail:
            raise RuntimeError("device lost")
        return np.full((img.shape[0] * self.scale, img.shape[1] * self.scale, 3), 7.0, np.float32)

    def close(self):
        self.closed = True


@pytest.fixture
def frame_png(tmp_path):
    import cv2
    p = str(tmp_path / "3.extract.png")
    cv2.imwrite(p, np.random.default_rng(0).integers(0, 256, (30, 1000, 3), dtype=np.uint8))
    return p


def test_upscale_image_contract(frame_png, tmp_path, monkeypatch):
    import cv2
    net = _FakeNet()
    monkeypatch.setattr(up, "net", net)
    out = str(tmp_path / "3.png")
    items = up.upscale_image(frame_png, out, 2, 1, 3, 10, remove=True)
    assert net.calls == [("u8", (30, 1000, 3), 960, 10)]
    assert items[:2] == [["debug", "Processing Tile: 1/2"], ["debug", "Processing Tile: 2/2"]]
    assert items[-1] == ["info", "Upscaling Batch: 1 : Upscaled 3/10"]
    assert not os.path.exists(frame_png) and cv2.imread(out).shape == (60, 2000, 3)


def test_upscale_image_variants(frame_png, monkeypatch):
    monkeypatch.setattr(up, "net", _FakeNet())
    items = up.upscale_image(frame_png, None, 2, None, 3, 10, remove=False)  # test_gpus.py usage
    assert items[-1] == ["info", "Upscaled 3/10"] and os.path.exists(frame_png)
    items = up.upscale_image(frame_png, "x.png", 2, [3, 4], 3, 10, remove=False)
    assert items[-1] == ["info", "Upscaled x.png"]
    os.remove("x.png")


def test_errors_become_log_items(frame_png, monkeypatch):
    net = _FakeNet(fail=True)
    monkeypatch.setattr(up, "net", net)
    items = up.upscale_image(frame_png, None, 2, None, 1, 1, remove=True)
    assert [i[0] for i in items[-2:]] == ["error", "error"] and items[-2][1] == "Upscale failed"
    assert os.path.exists(frame_png), "input must survive a failed frame (resume contract)"
    assert net.closed and up.net is None
    monkeypatch.setattr(up, "net", _FakeNet(fail=True))
    items = up.apply_model(frame_png, "o.png", True)
    assert items[0] == ["error", "Model processing failed"] and os.path.exists(frame_png)
    # wrong scale is an error item too, not an exception
    monkeypatch.setattr(up, "net", _FakeNet())
    assert up.upscale_image(frame_png, None, 4, None, 1, 1, remove=False)[-2][0] == "error"


def test_process_tile_scatter(monkeypatch):
    net = _FakeNet()
    monkeypatch.setattr(up, "net", net)
    img = np.zeros((30, 1000, 3), np.uint8)
    canvas = np.zeros((60, 2000, 3))
    items = []
    assert up.process_tile(img, 960, 2, 0, 1, 30, 1000, canvas, items) is None
    assert net.calls == [("f32", (30, 50, 3), 0, 0)]  # 40 core columns + 10 halo on the left
    assert (canvas[:, 1920:] == 7.0).all() and (canvas[:, :1920] == 0).all()
    monkeypatch.setattr(up, "net", _FakeNet(fail=True))
    assert up.process_tile(img, 960, 2, 0, 0, 30, 1000, canvas, items) == -1 and items[0][0] == "error"


def test_shard_frames():
    frames = list(range(1, 12))
    parts = [parallel.shard_frames(frames, r, 4) for r in range(4)]
    assert sorted(sum(parts, [])) == frames and parts[1] == [2, 6, 10]


def _gloo_worker(rank, world, port, model_dir, q):
    import torch.distributed as dist
    from upscale_video_b200 import ncnn_model
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    loads = []

    def load():
        loads.append(1)
        return ncnn_model.pack_compact_blob(ncnn_model.load_model(model_dir, "2x_Compact_Pretrain"))

    desc, blob = parallel.broadcast_packed_model(load, rank, world, device=None)

    def load_fused():  # an RRDB graph travels as its fused program (op list + buffer table + weights)
        loads.append(2)
        return ncnn_model.compile_fused(ncnn_model.load_model(model_dir, "4x_Valar_v1"))

    prog = parallel.broadcast_packed_model(load_fused, rank, world, device=None)
    fused_sig = (len(prog.ops), len(prog.bufs), prog.scale, prog.ops[2]["sc_cin"], prog.ops[2]["sc_w_off"], prog.ops[-1]["final"],
                 float(prog.weights.astype("float64").sum()), int(prog.weights.size))
    q.put((rank, loads, (desc.cin, desc.nf, desc.n_mid, desc.scale, desc.input_blob), float(blob.sum()), blob.size,
           parallel.shard_frames(range(10), rank, world), fused_sig))
    dist.destroy_process_group()


def test_weight_broadcast_gloo_world2(model_dir):
    """N>1 path on CPU: rank 0 alone reads the model, rank 1 receives identical bytes; frames shard disjointly."""
    import multiprocessing as mp
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, model_dir, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    (r0, l0, d0, s0, n0, f0, g0), (r1, l1, d1, s1, n1, f1, g1) = res
    assert (l0, l1) == ([1, 2], []) and d0 == d1 == (3, 64, 16, 2, "input") and s0 == s1 and n0 == n1 == 598464 + 1100 + 1088
    assert sorted(f0 + f1) == list(range(10)) and not set(f0) & set(f1)
    assert g0 == g1 and g0[:3] == (353, 12, 4) and g0[3] == 64 and g0[5] == 1  # the fused Valar program arrives intact


class _FakeEngine:
    """Host-logic stand-in for raw_stream tests: 'upscales' by pixel repetition, records calls."""

    def __init__(self, scale):
        self.scale, self.calls = scale, []

    def run_batch_host(self, h_in, h_out, n, h, w, tile=960, halo=10):
        a = h_in.numpy() if hasattr(h_in, "numpy") else h_in
        o = h_out.numpy() if hasattr(h_out, "numpy") else h_out
        self.calls.append((n, h, w, tile, halo))
        o[:n] = np.repeat(np.repeat(a[:n], self.scale, 1), self.scale, 2)


def test_raw_stream_framing(tmp_path):
    """raw_stream: chunking, short last chunk, pipe-like short reads, rgb24 channel swap, truncated input."""
    import io
    from upscale_video_b200 import raw_stream
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 256, (5, 6, 8, 3), dtype=np.uint8)

    class Dribble(io.BytesIO):  # returns at most 100 bytes per read, like a pipe
        def read(self, n=-1):
            return super().read(min(n, 100) if n >= 0 else 100)

    eng = _FakeEngine(2)
    out = io.BytesIO()
    n = raw_stream.stream(Dribble(frames.tobytes()), out, 8, 6, scale=2, chunk=2, upscaler=eng)
    got = np.frombuffer(out.getvalue(), np.uint8).reshape(5, 12, 16, 3)
    assert n == 5 and [c[0] for c in eng.calls] == [2, 2, 1] and eng.calls[0][3:] == (960, 10)
    assert np.array_equal(got, np.repeat(np.repeat(frames, 2, 1), 2, 2))
    # rgb24: the engine sees BGR, the stream stays RGB
    eng = _FakeEngine(2)
    out = io.BytesIO()
    raw_stream.stream(io.BytesIO(frames.tobytes()), out, 8, 6, scale=2, chunk=4, pix_fmt="rgb24", upscaler=eng)
    assert np.array_equal(np.frombuffer(out.getvalue(), np.uint8).reshape(5, 12, 16, 3), np.repeat(np.repeat(frames, 2, 1), 2, 2))
    # pre-pass only (scale 1, -m a): untiled calls
    pre = _FakeEngine(1)
    out = io.BytesIO()
    raw_stream.stream(io.BytesIO(frames.tobytes()), out, 8, 6, scale=1, models=["a"], chunk=3, prepass=pre)
    assert [c[3:] for c in pre.calls] == [(0, 0), (0, 0)] and out.getvalue() == frames.tobytes()
    # -m n=K: the denoiser sees every frame first, with the clamped level; then the upscaler sees its output
    class FakeDenoiser:
        def __init__(self):
            self.calls = []

        def run_batch_host(self, h_in, h_out, n, h, w, level, level_color=None):
            a = h_in.numpy() if hasattr(h_in, "numpy") else h_in
            o = h_out.numpy() if hasattr(h_out, "numpy") else h_out
            self.calls.append((n, h, w, level))
            o[:n] = 255 - a[:n]

    dn, eng = FakeDenoiser(), _FakeEngine(2)
    out = io.BytesIO()
    raw_stream.stream(io.BytesIO(frames.tobytes()), out, 8, 6, scale=2, models=["n=45"], chunk=2, upscaler=eng, denoiser=dn)
    assert dn.calls == [(2, 6, 8, 30), (2, 6, 8, 30), (1, 6, 8, 30)] and [c[0] for c in eng.calls] == [2, 2, 1]
    assert np.array_equal(np.frombuffer(out.getvalue(), np.uint8).reshape(5, 12, 16, 3), np.repeat(np.repeat(255 - frames, 2, 1), 2, 2))
    dn = FakeDenoiser()  # denoise only (scale 1)
    out = io.BytesIO()
    raw_stream.stream(io.BytesIO(frames.tobytes()), out, 8, 6, scale=1, models=["n=3"], chunk=4, denoiser=dn)
    assert out.getvalue() == (255 - frames).tobytes() and dn.calls[0][3] == 3
    assert raw_stream.denoise_level(["a", "n=0"]) is None and raw_stream.denoise_level(["r"]) is None
    # max_frames and truncation
    out = io.BytesIO()
    assert raw_stream.stream(io.BytesIO(frames.tobytes()), out, 8, 6, scale=2, chunk=2, max_frames=3, upscaler=_FakeEngine(2)) == 3
    with pytest.raises(ValueError, match="truncated"):
        raw_stream.stream(io.BytesIO(frames.tobytes()[:-7]), io.BytesIO(), 8, 6, scale=2, chunk=2, upscaler=_FakeEngine(2))


def test_raw_stream_overlapped_pump():
    """raw_stream --overlap: reader / engine / writer on three threads produce the same bytes, in order, as the sequential
    loop; errors on any stage surface on the caller's thread and nothing hangs."""
    import io
    import time
    from upscale_video_b200 import raw_stream
    rng = np.random.default_rng(1)
    frames = rng.integers(0, 256, (23, 6, 8, 3), dtype=np.uint8)
    want = np.repeat(np.repeat(frames, 2, 1), 2, 2).tobytes()

    class Dribble(io.BytesIO):
        def read(self, n=-1):
            return super().read(min(n, 173) if n >= 0 else 173)

    class SlowEngine(_FakeEngine):  # lets the reader run ahead and the writer lag behind
        def run_batch_host(self, *a, **k):
            time.sleep(0.002)
            super().run_batch_host(*a, **k)

    for chunk in (1, 2, 5, 23, 40):
        eng = SlowEngine(2)
        out = io.BytesIO()
        assert raw_stream.stream(Dribble(frames.tobytes()), out, 8, 6, scale=2, chunk=chunk, upscaler=eng, overlap=True) == 23
        assert out.getvalue() == want and sum(c[0] for c in eng.calls) == 23
    out = io.BytesIO()  # rgb24 + max_frames
    assert raw_stream.stream(io.BytesIO(frames.tobytes()), out, 8, 6, scale=2, chunk=4, pix_fmt="rgb24", max_frames=10,
                             upscaler=_FakeEngine(2), overlap=True) == 10
    assert out.getvalue() == want[:10 * 12 * 16 * 3]
    assert raw_stream.stream(io.BytesIO(b""), io.BytesIO(), 8, 6, scale=2, chunk=4, upscaler=_FakeEngine(2), overlap=True) == 0
    with pytest.raises(ValueError, match="truncated"):  # reader error
        raw_stream.stream(io.BytesIO(frames.tobytes()[:-7]), io.BytesIO(), 8, 6, scale=2, chunk=2, upscaler=_FakeEngine(2), overlap=True)

    class Broken(_FakeEngine):  # engine error on the third chunk
        def run_batch_host(self, *a, **k):
            if len(self.calls) == 2:
                raise RuntimeError("device lost")
            super().run_batch_host(*a, **k)

    with pytest.raises(RuntimeError, match="device lost"):
        raw_stream.stream(io.BytesIO(frames.tobytes()), io.BytesIO(), 8, 6, scale=2, chunk=2, upscaler=Broken(2), overlap=True)

    class ClosedPipe(io.BytesIO):  # writer error
        def write(self, b):
            if self.tell() > 3000:
                raise BrokenPipeError("reader went away")
            return super().write(b)

    with pytest.raises(BrokenPipeError):
        raw_stream.stream(io.BytesIO(frames.tobytes()), ClosedPipe(), 8, 6, scale=2, chunk=2, upscaler=SlowEngine(2), overlap=True)


def test_denoise_worker_device_choice(monkeypatch):
    monkeypatch.setenv("B2SR_DENOISE_GPUS", "2, 5")
    assert up._denoise_device() == 2  # not a pool worker: slot 0
    monkeypatch.setattr(up.multiprocessing.current_process(), "_identity", (4,), raising=False)
    assert up._denoise_device() == 5  # fourth child of the parent -> slot 3 -> listed[1]
    monkeypatch.delenv("B2SR_DENOISE_GPUS")
    monkeypatch.setattr(up._engine, "device_count", lambda: 3)
    assert up._denoise_device() == 0  # nothing selected: device 0, never a GPU the user did not name


def test_compat_shim_serves_the_reference_cli_imports():
    """upscale_video_b200/compat: every name the reference's CLIs import from `upscale.upscale_processing` /
    `ncnn_vulkan` resolves to the drop-in, with the reference's argument names (checked against the reference sources
    when they are around: /root/reference in the build container, baseline/_ref on the GPU box)."""
    import ast
    import importlib
    import inspect
    compat = os.path.join(ROOT, "upscale_video_b200", "compat")
    sys.path.insert(0, compat)
    try:
        for m in ("upscale", "upscale.upscale_processing", "ncnn_vulkan"):
            sys.modules.pop(m, None)
        shim = importlib.import_module("upscale.upscale_processing")
        ncnn = importlib.import_module("ncnn_vulkan").ncnn
        assert callable(ncnn.get_gpu_count) and callable(ncnn.get_default_gpu_index) and callable(ncnn.get_gpu_info)
        wanted = set()
        for base in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
            for name in ("test_gpus.py", "test_images.py", "upscale_video.py"):
                p = os.path.join(base, name)
                if not os.path.exists(p):
                    continue
                for node in ast.walk(ast.parse(open(p).read())):
                    if isinstance(node, ast.ImportFrom) and node.module == "upscale.upscale_processing":
                        wanted |= {a.name for a in node.names}
        wanted.discard("process_file")  # the ffmpeg pipeline around the hot path: out of scope
        for name in wanted | {"init_worker", "upscale_image", "upscale_frames", "process_model", "process_denoise", "get_frames"}:
            assert callable(getattr(shim, name)), name
        ref_src = "/root/reference/upscale/upscale_processing.py"
        if os.path.exists(ref_src):
            ref_args = {n.name: [a.arg for a in n.args.args] for n in ast.parse(open(ref_src).read()).body if isinstance(n, ast.FunctionDef)}
            for name in ("init_worker", "upscale_image", "upscale_frames", "process_model", "process_denoise", "apply_model",
                         "apply_denoise", "process_tile", "logging_callback", "get_frames"):
                assert list(inspect.signature(getattr(shim, name)).parameters) == ref_args[name], name
    finally:
        sys.path.remove(compat)
        for m in ("upscale", "upscale.upscale_processing", "ncnn_vulkan"):
            sys.modules.pop(m, None)


def test_raw_stream_multi_gpu_dynamic_queue(tmp_path):
    """raw_stream.stream_multi: one worker thread per -g entry over one dynamic queue of chunks (reference upscale_frames
    :565-598 with raw chunks instead of PNG names), in-order writer; bytes equal the single-GPU stream for a pipe input and
    for a seekable file (which the workers read themselves with preadv); errors surface, nothing hangs."""
    import io
    import random
    import threading
    import time
    from upscale_video_b200 import raw_stream
    rng = np.random.default_rng(2)
    frames = rng.integers(0, 256, (37, 6, 8, 3), dtype=np.uint8)
    want = np.repeat(np.repeat(frames, 2, 1), 2, 2).tobytes()
    seen = {}
    lock = threading.Lock()

    class Jitter(_FakeEngine):  # workers finish out of order
        def __init__(self, gpu):
            super().__init__(2)
            self.gpu = gpu

        def run_batch_host(self, h_in, h_out, n, h, w, *a, **k):
            time.sleep(random.random() * 0.004)
            with lock:
                seen[self.gpu] = seen.get(self.gpu, 0) + n
            super().run_batch_host(h_in, h_out, n, h, w, *a, **k)

    class Pipe(io.BytesIO):  # not seekable: one reader thread
        def seekable(self):
            return False

        def fileno(self):
            raise OSError("no fd")

    for chunk in (1, 3, 4, 40):
        seen.clear()
        out = io.BytesIO()
        n = raw_stream.stream_multi(Pipe(frames.tobytes()), out, 8, 6, scale=2, gpus=[0, 1, 1, 2], chunk=chunk,
                                    make_engines=lambda g: (None, None, Jitter(g)))
        assert n == 37 and out.getvalue() == want and sum(seen.values()) == 37
    assert len(seen) == 1  # one chunk of 40: one worker did everything
    path = tmp_path / "frames.raw"
    path.write_bytes(frames.tobytes())
    for chunk, mf in ((2, None), (5, 11), (4, 37)):
        seen.clear()
        out = io.BytesIO()
        with open(path, "rb") as f:
            n = raw_stream.stream_multi(f, out, 8, 6, scale=2, gpus=[0, 1, 2], chunk=chunk, max_frames=mf, make_engines=lambda g: (None, None, Jitter(g)))
        k = 37 if mf is None else mf
        assert n == k and out.getvalue() == want[:k * 12 * 16 * 3]
    assert len(seen) == 3  # every worker took chunks

    # engines with the streaming call (Engine.submit_batch_host / wait_batch): a worker keeps two chunks in flight
    class Deferred(Jitter):
        def __init__(self, gpu):
            super().__init__(gpu)
            self.jobs, self.tickets, self.in_flight, self.max_in_flight = {}, 0, 0, 0

        def submit_batch_host(self, h_in, h_out, n, h, w, *a, **k):
            self.tickets += 1
            self.jobs[self.tickets] = (h_in, h_out, n, h, w)
            self.in_flight += 1
            self.max_in_flight = max(self.max_in_flight, self.in_flight)
            return self.tickets

        def wait_batch(self, ticket):
            assert ticket == min(self.jobs), "submissions are waited for in order"
            self.run_batch_host(*self.jobs.pop(ticket))  # (the output appears only now: nothing may be written before its wait)
            self.in_flight -= 1

    class Deferred1(Deferred):
        def __init__(self, gpu):
            super().__init__(gpu)
            self.scale = 1

    made = []
    for chunk, gpus in ((1, [0]), (3, [0, 1, 1]), (40, [0, 1])):
        seen.clear()
        out = io.BytesIO()
        n = raw_stream.stream_multi(Pipe(frames.tobytes()), out, 8, 6, scale=2, gpus=gpus, chunk=chunk,
                                    make_engines=lambda g: (None, None, made.append(Deferred(g)) or made[-1]))
        assert n == 37 and out.getvalue() == want and sum(seen.values()) == 37
    assert all(e.in_flight == 0 and not e.jobs and e.max_in_flight <= 2 for e in made) and any(e.max_in_flight == 2 for e in made)

    class BrokenWait(Deferred):
        def wait_batch(self, ticket):
            raise RuntimeError("device lost in flight")

    with pytest.raises(RuntimeError, match="device lost in flight"):
        raw_stream.stream_multi(Pipe(frames.tobytes()), io.BytesIO(), 8, 6, scale=2, gpus=[0, 1], chunk=2,
                                make_engines=lambda g: (None, None, BrokenWait(g) if g == 1 else Deferred(g)))
    # rgb24 in and out, denoise -> pre-pass -> upscale chain per worker
    class Neg:
        def run_batch_host(self, h_in, h_out, n, h, w, level, level_color=None):
            a = h_in.numpy() if hasattr(h_in, "numpy") else h_in
            o = h_out.numpy() if hasattr(h_out, "numpy") else h_out
            o[:n] = 255 - a[:n]

    out = io.BytesIO()
    raw_stream.stream_multi(Pipe(frames.tobytes()), out, 8, 6, scale=2, models=["n=3", "a"], gpus=[0, 0], chunk=3, pix_fmt="rgb24",
                            make_engines=lambda g: (Neg(), _FakeEngine(1), _FakeEngine(2)))
    assert out.getvalue() == np.repeat(np.repeat(255 - frames, 2, 1), 2, 2).tobytes()
    # the same chain with a streaming last stage: the earlier stages of chunk k+1 run while chunk k's upscale is in flight, each
    # chunk on its own set of intermediate buffers (Deferred produces its output only at the wait, from the buffers it was given)
    for chunk in (1, 3):
        out = io.BytesIO()
        raw_stream.stream_multi(Pipe(frames.tobytes()), out, 8, 6, scale=2, models=["n=3", "a"], gpus=[0, 0], chunk=chunk, pix_fmt="rgb24",
                                make_engines=lambda g: (Neg(), _FakeEngine(1), Deferred(g)))
        assert out.getvalue() == np.repeat(np.repeat(255 - frames, 2, 1), 2, 2).tobytes()
    out = io.BytesIO()  # scale 1: the pre-pass is the last stage
    raw_stream.stream_multi(Pipe(frames.tobytes()), out, 8, 6, scale=1, models=["n=3", "a"], gpus=[0, 1], chunk=2,
                            make_engines=lambda g: (Neg(), Deferred1(g), None))
    assert out.getvalue() == (255 - frames).tobytes()
    # truncated inputs, a failing engine, a closed output pipe
    with pytest.raises(ValueError, match="truncated"):
        raw_stream.stream_multi(Pipe(frames.tobytes()[:-5]), io.BytesIO(), 8, 6, scale=2, gpus=[0, 1], chunk=2, make_engines=lambda g: (None, None, _FakeEngine(2)))
    path.write_bytes(frames.tobytes()[:-5])
    with open(path, "rb") as f, pytest.raises(ValueError, match="truncated"):
        raw_stream.stream_multi(f, io.BytesIO(), 8, 6, scale=2, gpus=[0, 1], chunk=2, make_engines=lambda g: (None, None, _FakeEngine(2)))

    class Broken(_FakeEngine):
        def run_batch_host(self, *a, **k):
            raise RuntimeError("device lost")

    with pytest.raises(RuntimeError, match="device lost"):
        raw_stream.stream_multi(Pipe(frames.tobytes()), io.BytesIO(), 8, 6, scale=2, gpus=[0, 1], chunk=2,
                                make_engines=lambda g: (None, None, Broken(2) if g == 1 else Jitter(g)))

    class ClosedPipe(io.BytesIO):
        def write(self, b):
            if self.tell() > 3000:
                raise BrokenPipeError("reader went away")
            return super().write(b)

    with pytest.raises(BrokenPipeError):
        raw_stream.stream_multi(Pipe(frames.tobytes()), ClosedPipe(), 8, 6, scale=2, gpus=[0, 1, 2], chunk=2, make_engines=lambda g: (None, None, Jitter(g)))


def test_host_helpers_behave_like_the_reference_code(caplog):
    """`get_frames` and `logging_callback` against the reference's own functions, imported unmodified with the two modules it
    cannot import here stubbed (tools/make_ref_glue_goldens.py::import_reference) -- only where /root/reference exists."""
    import logging
    if not os.path.exists("/root/reference/upscale/upscale_processing.py"):
        pytest.skip("the reference tree is only present in the build container")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import make_ref_glue_goldens as glue
    finally:
        sys.path.pop(0)
    keep = {m: sys.modules.get(m) for m in ("ncnn_vulkan", "wakepy")}
    try:
        ref = glue.import_reference()
    finally:
        for m, v in keep.items():
            if v is None:
                sys.modules.pop(m, None)
            else:
                sys.modules[m] = v
    from upscale_video_b200 import upscale_processing as up

    def outcome(fn, *a):
        try:
            return ("ok", fn(*a))
        except SystemExit as e:
            return ("exit", str(e.code))
        except Exception as e:  # noqa: BLE001 -- the exception TYPE is part of the behaviour compared
            return ("raise", type(e).__name__)

    for spec in ("1", "1,3-5", "7-7", "5-3", "10,2,2", "1-3,2-4", " 4 , 6-8", "", "a", "1-", "1-2-3", "3,,4"):
        assert outcome(up.get_frames, spec) == outcome(ref.get_frames, spec), spec
    for items in ([], [["info", "a"], ["debug", "b"]], [["info", "a"], ["error", "boom"], ["info", "never logged"]], [["error", ValueError("x")]],
                  [["warning", "ignored level"]]):
        caplog.clear()
        with caplog.at_level(logging.DEBUG):
            mine = outcome(up.logging_callback, items)
        mine_log = [(r.levelname, r.getMessage()) for r in caplog.records]
        caplog.clear()
        with caplog.at_level(logging.DEBUG):
            theirs = outcome(ref.logging_callback, items)
        assert mine == theirs and mine_log == [(r.levelname, r.getMessage()) for r in caplog.records], items
