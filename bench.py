#!/usr/bin/env python
"""bench.py -- 1080p frames/s of the frame-upscale hot path (2x_Compact_Pretrain) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A *step* is one pass of the hot path (reference upscale_frames -> upscale_image -> process_tile -> network,
upscale/upscale_processing.py:480-598) over one batch of B synthetic 1920x1080 BGR u8 frames per GPU, with the
reference's 960-px tiling + 10-px halo reproduced on the device (4 planes per frame).  Frames shard across
ranks with no data-path collective (weak scaling: B frames per rank); NCCL only broadcasts the packed weight
blob from rank 0 at start-up.

The one JSON line printed by rank 0 follows the contract in the task statement:
  value      frames/s with inputs and outputs resident in HBM (CUDA events on the engine's stream, max over ranks)
  e2e        frames/s through Engine.submit_batch_host / wait_batch (and, beside it, the synchronous run_batch_host): pinned host
             buffers, every step's H2D and D2H inside the timed region
  roofline   the dominant kernel (the persistent whole-network tcgen05 kernel): algorithmic FLOPs per launch / mean
             launch time (per-launch CUDA events recorded on the engine's stream inside the timed region)
  cpu_baseline  the CPU oracle (oracle/, a port of the reference graph + glue -- ncnn itself is not installable)
             timed on the host cores on a bounded sample of the same workload
`--impl reference` times that CPU port alone, with every host thread, on the same config (rank 0 only).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = "2x_Compact_Pretrain"
H, W, SCALE = 1080, 1920, 2
TILE, HALO = 960, 10
MAC_PER_PX_MID = 64 * 64 * 9                 # one nf->nf convolution
MAC_PER_PX_NET = 598464                      # whole 2x_Compact graph
METRIC = "1080p frames/sec (2x_Compact_Pretrain)"
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r02_traffic.json")


def load_traffic(key, frames_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed ncu capture of THIS
    round's kernel (profiles/r02_traffic.json, written by tools/ncu_traffic.py from the ncu CSV kept beside it).  Returns
    (bytes or None, note).  The capture's launch shape must be the bench's; otherwise the figure is scaled per frame and says so."""
    try:
        d = json.load(open(TRAFFIC_FILE))[key]
    except Exception:
        return None, "no ncu capture of the current kernel committed (profiles/r02_traffic.json)"
    per_frame = (d["dram_read_bytes"] + d["dram_write_bytes"]) / float(d["frames_per_launch"])
    note = "%s: %.3f GB read + %.3f GB written per launch of %d frames (%s)" % (
        d["kernel"], d["dram_read_bytes"] / 1e9, d["dram_write_bytes"] / 1e9, d["frames_per_launch"], d["source"])
    if abs(frames_per_launch - d["frames_per_launch"]) > 1e-6:
        note += "; scaled per frame to this run's %g frames per launch" % frames_per_launch
    return per_frame * frames_per_launch, note


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap", nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_port_fps(sample_hw=(H, W), reps=1, threads=None):
    """Times the CPU oracle (f32, what ncnn's CPU path computes in) on one whole synthetic 1080p frame -- all four
    reference tiles (970x970, 970x970, 130x970, 130x970), i.e. the bench's own config -- or, with a smaller `sample_hw`,
    on a crop scaled by area.  Returns (fps, threads, description, seconds per repetition)."""
    from oracle import oracle
    from upscale_video_b200 import ncnn_model
    threads = threads or os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(threads)
    layers = oracle.read_model(ncnn_model.packaged_model_dir(), MODEL)
    rng = np.random.default_rng(0)
    h, w = sample_hw
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    oracle.upscale_image_array(layers, img[:32, :64], SCALE, "f32")  # warm the library
    t0 = time.time()
    for _ in range(reps):
        oracle.upscale_image_array(layers, img, SCALE, "f32", TILE, HALO)
    dt = (time.time() - t0) / reps
    fps = (h * w) / float(H * W) / dt
    if (h, w) == (H, W):
        desc = "one whole 1080p frame per repetition (4 reference tiles, 960 + 10), f32 C oracle with OpenMP"
    else:
        desc = "%dx%d crop (1/%.1f of a 1080p frame), f32 C oracle with OpenMP, scaled by area" % (h, w, H * W / float(h * w))
    return fps, threads, desc, dt


def cpu_torch_fps(threads=None, reps=1):
    """the same graph, tiling and pre/post-processing on torch-CPU (oneDNN conv2d, fp32, channels
    first), `torch.set_num_threads(all host threads)` -- a second CPU restatement timed next to the C oracle.  One whole
    1080p frame (4 reference tiles) per repetition.  Returns (fps, threads, seconds per repetition)."""
    import torch
    import torch.nn.functional as F
    from upscale_video_b200 import ncnn_model
    from upscale_video_b200.upscale_processing import tile_rect
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    desc, blob = ncnn_model.pack_compact_blob(ncnn_model.load_model(ncnn_model.packaged_model_dir(), MODEL))
    blob = torch.from_numpy(np.ascontiguousarray(blob, np.float32))
    layers, off = [], 0
    shapes = [(desc.nf, desc.cin)] + [(desc.nf, desc.nf)] * desc.n_mid + [(desc.cin * desc.scale ** 2, desc.nf)]
    for i, (co, ci) in enumerate(shapes):
        w = blob[off:off + co * ci * 9].reshape(co, ci, 3, 3); off += co * ci * 9
        b = blob[off:off + co]; off += co
        sl = None
        if i < len(shapes) - 1:
            sl = blob[off:off + co]; off += co
        layers.append((w, b, sl))
    assert off == blob.numel()
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)

    def frame():
        out = np.zeros((H * SCALE, W * SCALE, 3), np.float64)
        for ty in range((H + TILE - 1) // TILE):
            for tx in range((W + TILE - 1) // TILE):
                (iy0, iy1, ix0, ix1), (cy0, cy1, cx0, cx1) = tile_rect(ty, tx, TILE, H, W, HALO)
                x0 = torch.from_numpy(img[iy0:iy1, ix0:ix1].astype(np.float32)).permute(2, 0, 1)[None] * np.float32(1 / 255.0)
                v = x0
                for w, b, sl in layers:
                    v = F.conv2d(v, w, b, padding=1)
                    if sl is not None:
                        v = F.prelu(v, sl)
                y = F.pixel_shuffle(v, SCALE) + F.interpolate(x0, scale_factor=SCALE, mode="nearest")
                y = (y[0].permute(1, 2, 0) * 255.0).numpy()
                oy, ox = (cy0 - iy0) * SCALE, (cx0 - ix0) * SCALE
                out[cy0 * SCALE:cy1 * SCALE, cx0 * SCALE:cx1 * SCALE] = y[oy:oy + (cy1 - cy0) * SCALE, ox:ox + (cx1 - cx0) * SCALE]
        return np.clip(np.rint(out), 0, 255).astype(np.uint8)

    with torch.no_grad():
        t0 = time.time()
        for _ in range(reps):
            frame()
        dt = (time.time() - t0) / reps
    return 1.0 / dt, threads, dt


REF_MAX_TIMED, REF_MAX_WARM = 3, 1  # one step = one whole frame = ~13 s on 16 threads: the arm times at most this many


def run_reference(args, rank):
    """--impl reference: the reference's arithmetic on the host CPU (port: ncnn_vulkan cannot be installed), on the bench's
    own config: every step is one whole synthetic 1080p frame through all four reference tiles."""
    if rank != 0:
        return
    warm, steps = min(args.warmup, REF_MAX_WARM), max(1, min(args.steps, REF_MAX_TIMED))
    fps_list = []
    info = None
    for i in range(warm + steps):
        fps, threads, desc, dt = cpu_port_fps(reps=1)
        if i >= warm:
            fps_list.append(fps)
        info = (threads, desc, dt)
    fps = float(np.mean(fps_list))
    sample = info[1] + "; %d step(s) timed after %d warm-up (requested %d / %d: capped so the arm ends within minutes)" % (
        steps, warm, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1000.0 / fps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "synthetic 1080p RGB batch, 2x_Compact_Pretrain, reference tiling 960+10 (BASELINE configs[1])",
                   "frame": [H, W, 3], "tile": TILE, "halo": HALO, "frames_per_step": 1, "steps_requested": args.steps,
                   "warmup_requested": args.warmup, "step": info[1]},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": info[0], "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_JSON_OUT or sys.stdout, flush=True)


def _event_timed(torch, stream_int, device, fn, steps, warm=2):
    """ms per call of `fn` (asynchronous launches on the engine's stream), CUDA events on that stream."""
    stream = torch.cuda.ExternalStream(stream_int, device=torch.device("cuda", device))
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
    for _ in range(steps):
        fn()
    with torch.cuda.stream(stream):
        e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / steps


def valar_540p(E, ncnn_model, torch, device, peaks, steps=10):
    """BASELINE configs[3]: synthetic 540p, 4x_Valar_v1 (RRDB) -- a side measurement reported next to the headline (not
    part of `value`): 4 device-resident frames per step, `steps` steps timed with CUDA events on the engine's stream,
    with its own roofline block (bound: tensor; algorithmic FLOPs = 2 x 18 068 160 MAC per input pixel)."""
    try:
        eng = E.Engine.from_files(ncnn_model.packaged_model_dir(), "4x_Valar_v1", device)
    except Exception as e:  # model not packaged
        return {"unavailable": str(e)[:200]}
    n, h, w = 4, 540, 960
    d_in = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda")
    d_out = torch.empty((n, h * 4, w * 4, 3), dtype=torch.uint8, device="cuda")
    eng.run_batch_device(d_in, d_out, n, h, w, TILE, HALO, sync=True)
    eng.reset_stats()
    ms = _event_timed(torch, eng.stream, device, lambda: eng.run_batch_device(d_in, d_out, n, h, w, TILE, HALO, sync=False), steps)
    eng.synchronize()
    launches = eng.stat(E.STAT_TC_LAUNCHES) / (steps + 2)
    pipes = eng.stat(E.STAT_PIPE_LAUNCHES) / (steps + 2)
    eng.close()
    fps = n / (ms * 1e-3)
    mac_px = 18068160
    tf = fps * 2.0 * mac_px * h * w / 1e12
    peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    traffic, tnote = load_traffic("valar_540p", n)
    return {"workload": "synthetic 540p, 4x_Valar_v1 RRDB (BASELINE configs[3]), tcgen05 graph kernels, 1xB200",
            "frames_per_s": fps, "ms_per_frame": 1e3 / fps, "tflops": tf,
            "frames_per_step": n, "steps": steps, "timer": "CUDA events on the engine's stream", "tcgen05_launches_per_step": launches,
            "persistent_rrdb_launches_per_step": pipes,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                         "frac_of_burst_peak": tf / peaks["bf16_tflops"], "flop_per_step": 2.0 * mac_px * h * w * n,
                         "traffic": traffic, "traffic_note": tnote,
                         "algorithmic_bytes_per_step": n * (h * w * 3 + 16 * h * w * 3)}}


def side_configs(E, ncnn_model, torch, device, steps=10):
    """The other BASELINE configs, each as a short device-resident measurement next to the headline (not part of `value`),
    timed with CUDA events over `steps` steps: configs[2] = 1x_HurrDeblur (whole frame, u8 out) -> 2x_Compact chained on the
    device at 1080p; configs[3]'s actual 4x pixel-shuffle reading (4x_Compact_Pretrain at 540p); the denoise pass."""
    out = []
    mdir = ncnn_model.packaged_model_dir()
    dev = torch.device("cuda", device)
    try:
        hurr = E.Engine.from_files(mdir, "1x_HurrDeblur_SubCompact_nf24-nc8_244k_net_g", device)
        comp = E.Engine.from_files(mdir, "2x_Compact_Pretrain", device)
        # Frames per hand-over, measured on one box with equal warm-up and duration (tools/chain_ab.py): 2 or 4 frames through two
        # hand-over buffers 316-317 fps, 8 frames 310, 16 frames 299.
        n = int(os.environ.get("B2SR_BENCH_CHAIN_N", "2"))
        nbuf = int(os.environ.get("B2SR_BENCH_CHAIN_BUFS", "2"))
        steps = max(1, steps * 16 // n)  # the same 160 frames per measurement (and 32 of warm-up) whatever the hand-over size:
        # the board's power governor settles over some 100 ms, so runs of different length or warmth do not compare
        d_in = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device="cuda")
        d_mids = [torch.empty_like(d_in), torch.empty_like(d_in)]  # two hand-over buffers: the pre-pass of step k+1 may run under step k's upscale
        d_mid = d_mids[0]
        d_out = torch.empty((n, 2 * H, 2 * W, 3), dtype=torch.uint8, device="cuda")
        sh, sc = torch.cuda.ExternalStream(hurr.stream, device=dev), torch.cuda.ExternalStream(comp.stream, device=dev)
        mid_ready = [torch.cuda.Event(), torch.cuda.Event()]
        mid_free = [torch.cuda.Event(), torch.cuda.Event()]
        turn = [0]

        def chain():  # the two engines own one stream each: events order each hand-over buffer's producer and consumer, nothing blocks the host
            k = turn[0] % nbuf
            turn[0] += 1
            sh.wait_event(mid_free[k])
            hurr.run_batch_device(d_in, d_mids[k], n, H, W, 0, 0, sync=False)       # apply_model: untiled, u8 out
            mid_ready[k].record(sh)
            sc.wait_event(mid_ready[k])
            comp.run_batch_device(d_mids[k], d_out, n, H, W, TILE, HALO, sync=False)  # upscale_image: 960 + 10 tiling
            mid_free[k].record(sc)
        mid_free[0].record(sc)
        mid_free[1].record(sc)
        for _ in range(max(2, 32 // n)):
            chain()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(sh)
        for _ in range(steps):
            chain()
        e1.record(sc)
        e1.synchronize()
        ms = e0.elapsed_time(e1) / steps
        hurr.reset_stats()
        ms_h = _event_timed(torch, hurr.stream, device, lambda: hurr.run_batch_device(d_in, d_mid, n, H, W, 0, 0, sync=False), steps)
        fps = n / (ms * 1e-3)
        out.append({"workload": "synthetic 1080p, 1x_HurrDeblur nf24 -> u8 -> 2x_Compact chained on the device (BASELINE configs[2])",
                    "frames_per_s": fps, "tflops": fps * 2.0 * (42768 + MAC_PER_PX_NET) * H * W / 1e12, "frames_per_step": n,
                    "steps": steps, "timer": "CUDA events (first engine's stream -> second engine's stream)",
                    "hurrdeblur_alone_frames_per_s": n / (ms_h * 1e-3), "hurrdeblur_alone_tflops": n / (ms_h * 1e-3) * 2.0 * 42768 * H * W / 1e12,
                    "hurrdeblur_schedule": "pipelined" if hurr.stat(E.STAT_PIPE_LAUNCHES) > 0 else "layer by layer"})
        hurr.close()
        comp.close()
        c4 = E.Engine.from_files(mdir, "4x_Compact_Pretrain", device)
        n, h, w = 16, 540, 960
        d_in = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda")
        d_out = torch.empty((n, 4 * h, 4 * w, 3), dtype=torch.uint8, device="cuda")
        ms = _event_timed(torch, c4.stream, device, lambda: c4.run_batch_device(d_in, d_out, n, h, w, TILE, HALO, sync=False), steps)
        fps = n / (ms * 1e-3)
        out.append({"workload": "synthetic 540p, 4x_Compact_Pretrain (the 4x pixel-shuffle model of BASELINE configs[3])",
                    "frames_per_s": fps, "tflops": fps * 2.0 * 619200 * h * w / 1e12, "frames_per_step": n, "steps": steps,
                    "timer": "CUDA events on the engine's stream"})
        c4.close()
    except Exception as e:
        out.append({"unavailable": str(e)[:200]})
    out.append(denoise_1080p(E, torch, device))
    return out


def denoise_1080p(E, torch, device, level=3):
    """the `-m n=<level>` denoise pass (reference apply_denoise, upscale_processing.py:350-362) at
    1080p -- device-resident frames/s, the same through pinned host buffers, and cv2.fastNlMeansDenoisingColored itself
    (the reference's implementation) on the host cores, on one whole frame."""
    try:
        dn = E.Denoiser(device)
        n = 16
        gen = torch.Generator(device="cuda").manual_seed(99)
        d_in = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device="cuda", generator=gen)
        d_out = torch.empty_like(d_in)
        stream = torch.cuda.ExternalStream(dn.stream, device=torch.device("cuda", device))
        for _ in range(2):
            dn.run_batch_device(d_in, d_out, n, H, W, level, sync=True)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 10
        with torch.cuda.stream(stream):
            ev0.record()
        for _ in range(steps):
            dn.run_batch_device(d_in, d_out, n, H, W, level, sync=False)
        with torch.cuda.stream(stream):
            ev1.record()
        dn.synchronize()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / steps
        h_in = torch.empty((n, H, W, 3), dtype=torch.uint8).pin_memory()
        h_in.copy_(d_in.cpu())
        h_out = torch.empty_like(h_in).pin_memory()
        dn.run_batch_host(h_in, h_out, n, H, W, level)
        t0 = time.perf_counter()
        for _ in range(3):
            dn.run_batch_host(h_in, h_out, n, H, W, level)
        e2e = 3 * n / (time.perf_counter() - t0)
        res = {"workload": "synthetic 1080p, fastNlMeansDenoisingColored(h=%d, hColor=%d, 5, 9) (-m n=%d), bit-exact vs cv2" % (level, level, level),
               "frames_per_s": n / (ms * 1e-3), "ms_per_launch": ms, "frames_per_launch": n, "e2e_frames_per_s": e2e,
               "kernel": "nlm_kernel (one launch per batch)", "bound": "integer ALU / shuffle issue",
               "algorithmic_GBps": n * H * W * 6 / (ms * 1e-3) / 1e9}
        try:
            import cv2
            img = h_in[0].numpy()
            cv2.fastNlMeansDenoisingColored(cv2.UMat(img[:64, :64].copy()), None, level, level, 5, 9).get()  # spin up cv2's thread pool
            t0 = time.perf_counter()
            ref = cv2.fastNlMeansDenoisingColored(cv2.UMat(img), None, level, level, 5, 9).get()
            dt = time.perf_counter() - t0
            res["cpu_baseline"] = {"value": 1.0 / dt, "unit": "frames/s", "cores": cv2.getNumThreads(), "kind": "reference",
                                   "sample": "one whole 1080p frame through cv2 %s (CPU path)" % cv2.__version__,
                                   "identical_to_gpu_result": bool(np.array_equal(ref, h_out[0].numpy()))}
        except ImportError:
            pass
        dn.close()
        return res
    except Exception as e:
        return {"workload": "denoise 1080p", "unavailable": str(e)[:200]}


_JSON_OUT = None  # the process's original stdout (see main)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16, help="frames per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the short side measurements of BASELINE configs[2] and [3]")
    ap.add_argument("--content", default="noise", choices=["noise", "natural"],
                    help="synthetic frame content: uniform random bytes (default; worst case for switching power) or smooth "
                         "gradients + edges + mild noise")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries ONE JSON line: everything libraries write to file descriptor 1 meanwhile (NCCL's version banner ignores
    # NCCL_DEBUG_FILE on this image) goes to stderr; the line itself is written to the saved descriptor
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    from upscale_video_b200 import engine as E
    from upscale_video_b200 import ncnn_model
    from upscale_video_b200 import parallel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path to time)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep NCCL's version banner off stdout: one JSON line there
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # rank 0 reads the model; everyone else receives the packed blob over NCCL (north_star: weight broadcast)
    packed = parallel.broadcast_packed_model(
        (lambda: ncnn_model.pack_compact_blob(ncnn_model.load_model(ncnn_model.packaged_model_dir(), MODEL))),
        rank, world, device=torch.device("cuda", local_rank))
    eng = E.Engine(device=local_rank, packed=packed)

    B = args.batch
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    def make_frames(content):
        if content == "noise":
            return torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device="cuda", generator=gen)
        yy = torch.arange(H, device="cuda", dtype=torch.float32)[None, :, None, None]
        xx = torch.arange(W, device="cuda", dtype=torch.float32)[None, None, :, None]
        ch = torch.arange(3, device="cuda", dtype=torch.float32)[None, None, None, :]
        fr = torch.arange(B, device="cuda", dtype=torch.float32)[:, None, None, None]
        img = 128 + 90 * torch.sin(xx / 37.0 + ch + 0.3 * fr) * torch.cos(yy / 23.0 - ch) + 40 * (((xx // 64) + (yy // 48)) % 2)
        img = img + 6 * torch.randn((B, H, W, 3), device="cuda", generator=gen)
        return img.clamp(0, 255).to(torch.uint8).contiguous()

    d_in = make_frames(args.content)
    d_out = torch.empty((B, H * SCALE, W * SCALE, 3), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        eng.run_batch_device(d_in, d_out, B, H, W, TILE, HALO, sync=False)

    for _ in range(args.warmup):
        step()
    eng.synchronize()
    eng.set_option(E.OPT_PROFILE, 1)
    eng.reset_stats()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record()
    for _ in range(args.steps):
        step()
    with torch.cuda.stream(stream):
        ev1.record()
    barrier()
    sampler.stop_flag = True
    ms = ev0.elapsed_time(ev1)
    launches = eng.stat(E.STAT_LAUNCHES)
    mid_ms, mid_n = eng.stat(E.STAT_TC_MID_MS), eng.stat(E.STAT_TC_MID_COUNT)
    pipe_ms, pipe_n = eng.stat(E.STAT_PIPE_MS), eng.stat(E.STAT_PIPE_LAUNCHES)
    all_ms = eng.stat(E.STAT_ALL_MS)
    eng.set_option(E.OPT_PROFILE, 0)
    eng.reset_stats()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * args.steps / (ms_max / 1000.0)

    # ---- end to end: pinned host frames in, pinned host frames out, copies inside the timed region ----
    e2e = None
    h_in = h_out = None
    if not args.no_e2e:
        h_in = torch.empty((B, H, W, 3), dtype=torch.uint8).pin_memory()
        h_in.copy_(d_in.cpu())
        h_out = torch.empty((B, H * SCALE, W * SCALE, 3), dtype=torch.uint8).pin_memory()
        e_steps = max(4, args.steps)

        def timed(fn):
            barrier()
            t0 = time.perf_counter()
            fn()
            barrier()
            te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if dist is not None:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            return world * B * e_steps / float(te.item())

        # (a) one synchronous call per step (b2sr_run_batch_host): returns when the step's last D2H has landed
        def sync_steps():
            for _ in range(e_steps):
                eng.run_batch_host(h_in, h_out, B, H, W, TILE, HALO)
        for _ in range(2):
            eng.run_batch_host(h_in, h_out, B, H, W, TILE, HALO)
        fps_sync = timed(sync_steps)
        checksum = int(h_out[0, ::97, ::89].to(torch.int64).sum())
        # (b) the streaming form of the same call (b2sr_submit_batch_host / b2sr_wait_batch): two steps in flight on two pairs of
        # pinned buffers, so step k+1's first H2D runs under step k's network and step k's last D2H under step k+1's; every step's
        # result is waited for and read on the host inside the timed region
        h_in2, h_out2 = h_in.clone().pin_memory(), torch.empty_like(h_out).pin_memory()
        pairs = [(h_in, h_out), (h_in2, h_out2)]
        sums = []

        def stream_steps():
            tickets = [None, None]
            for k in range(e_steps + 1):
                if k < e_steps:
                    tickets[k & 1] = eng.submit_batch_host(pairs[k & 1][0], pairs[k & 1][1], B, H, W, TILE, HALO)
                if k >= 1:
                    eng.wait_batch(tickets[(k - 1) & 1])
                    sums.append(int(pairs[(k - 1) & 1][1][0, ::97, ::89].to(torch.int64).sum()))
        e_keep = e_steps
        e_steps = 2
        stream_steps()  # warm-up (pins, plans)
        e_steps = e_keep
        sums.clear()
        fps_stream = timed(stream_steps)
        assert all(v == checksum for v in sums), "streamed steps disagree with the synchronous call"
        e2e = {"value": fps_stream, "unit": "frames/s", "h2d_bytes_per_step": int(h_in.numel()),
               "d2h_bytes_per_step": int(h_out.numel()), "steps": e_steps,
               "api": "b2sr_submit_batch_host / b2sr_wait_batch, two steps in flight on two pairs of pinned host buffers; every step's output "
                      "is waited for and read on the host inside the timed region",
               "timer": "host perf_counter from the first submit to the last wait", "result_checksum": checksum,
               "synchronous_call_value": fps_sync,
               "synchronous_call_note": "one b2sr_run_batch_host per step (returns when the step's last D2H has landed)"}

    # ---- the same workload on the other synthetic content (side measurement: switching power depends on the data) ----
    alt = None
    if not args.no_extra and world == 1:
        try:
            other = "natural" if args.content == "noise" else "noise"
            d_in.copy_(make_frames(other))
            for _ in range(3):
                step()
            eng.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                a0.record()
            for _ in range(args.steps):
                step()
            with torch.cuda.stream(stream):
                a1.record()
            eng.synchronize()
            torch.cuda.synchronize()
            fps_alt = B * args.steps / (a0.elapsed_time(a1) / 1000.0)
            alt = {"workload": "the headline workload (BASELINE configs[1]) with content = %s (smooth gradients + edges + mild noise)" % other
                   if other == "natural" else "the headline workload (BASELINE configs[1]) with content = noise",
                   "frames_per_s": fps_alt, "tflops": fps_alt * 2.0 * MAC_PER_PX_NET * H * W / 1e12, "frames_per_step": B, "steps": args.steps}
        except Exception as e:  # a side measurement must never take the headline line down
            alt = {"workload": "headline workload on the other content", "unavailable": str(e)[:200]}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks, peak_src = load_peaks()
    peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    if pipe_n:
        # dominant kernel = the persistent whole-network kernel (one launch per pass of frames_per_launch frames)
        frames_per_launch = B * args.steps / pipe_n
        flop_launch = 2.0 * MAC_PER_PX_NET * frames_per_launch * H * W  # exact frame, no halo/padding/junk columns
        k_ms, k_n = pipe_ms, pipe_n
        kernel = ("tc_pipe_kernel<64,16,2> (whole 2x_Compact network in one persistent launch: CTA = layer x band, "
                  "18 tcgen05 conv stages chained through L2-resident row rings)")
    else:
        # layer-by-layer schedule: dominant kernel = the 64->64 convolution (16 launches per pass)
        frames_per_launch = B * args.steps * 16.0 / max(mid_n, 1)
        flop_launch = 2.0 * MAC_PER_PX_MID * frames_per_launch * H * W
        k_ms, k_n = mid_ms, mid_n
        kernel = "tc_conv_kernel<64,64,0> (3x3 conv 64->64 + bias + PReLU, tcgen05 kind::f16)"
    ach = flop_launch / (k_ms / max(k_n, 1) * 1e-3) / 1e12 if k_n else None
    roofline = {
        "bound": "tensor", "kernel": kernel,
        "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": (ach / peak if ach else None),
        "peak_source": peak_src + ", bf16_tflops_sustained (kernel timed inside a long step)",
        "frac_of_burst_peak": (ach / peaks["bf16_tflops"] if ach else None),
        "flop_per_launch": flop_launch, "frames_per_launch": frames_per_launch, "launches_timed": k_n,
        "mean_launch_ms": k_ms / max(k_n, 1),
        "share_of_step": (k_ms / all_ms if all_ms else None),
        "whole_net_tflops": value / world * 2.0 * MAC_PER_PX_NET * H * W / 1e12,
    }
    roofline["traffic"], roofline["traffic_note"] = load_traffic("compact2x_1080p", frames_per_launch)
    roofline["algorithmic_bytes_per_launch"] = frames_per_launch * (H * W * 3 + SCALE * SCALE * H * W * 3)
    cpu = cpu_torch = None
    if not args.no_cpu_baseline and world == 1:  # reported on rank 0 at N = 1 only
        fps, threads, desc, dt = cpu_port_fps(reps=1)
        cpu = {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": desc + ", 1 repetition",
               "seconds": dt}
        try:  # the second CPU restatement, torch-CPU / oneDNN, same frame and tiling
            tfps, tthreads, tdt = cpu_torch_fps(threads)
            cpu_torch = {"value": tfps, "unit": "frames/s", "cores": tthreads, "kind": "port (torch %s CPU conv2d, oneDNN, fp32)" % torch.__version__,
                         "sample": "one whole 1080p frame (4 reference tiles), 1 repetition", "seconds": tdt,
                         "faster_than_c_oracle": bool(tfps > fps)}
        except Exception as e:
            cpu_torch = {"unavailable": str(e)[:200]}
    io_mb = (d_in.numel() + d_out.numel()) >> 20
    extra = None
    if not args.no_extra and world == 1:
        eng.close()
        del d_in, d_out, h_in, h_out
        torch.cuda.empty_cache()
        extra = [alt, valar_540p(E, ncnn_model, torch, local_rank, peaks)] + side_configs(E, ncnn_model, torch, local_rank)
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": {"workload": "synthetic 1080p RGB batch, 2x_Compact_Pretrain, 1xB200 per rank (BASELINE configs[1])",
                   "frames_per_gpu_per_step": B, "frame": [H, W, 3], "tile": TILE, "halo": HALO, "content": args.content,
                   "arithmetic": "fp16 weights and activations (tcgen05 kind::f16), fp32 accumulation and epilogue, u8 in/out",
                   "l2": "inputs+outputs per step are %d MB per GPU, larger than the 126 MB L2" % io_mb,
                   "parallelism": "frames sharded over %d rank(s), no data-path collective; weights NCCL-broadcast" % world},
        "roofline": roofline, "cpu_baseline": cpu, "cpu_baseline_torch": cpu_torch, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": sampler.summary(), "other_configs": extra,
    }
    print(json.dumps(line), file=_JSON_OUT or sys.stdout, flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
