"""Raw-frame streaming worker: the frame-upscale path without the PNG round trips.

The reference moves every frame through two PNG files -- ffmpeg writes ``N.extract.png`` (``extract_frames``, reference
upscale/upscale_processing.py:203-255), the worker ``cv2.imread``s it and ``cv2.imwrite``s ``N.png`` (:487, :519), ffmpeg
reads those back (``merge_frames`` :604-686) -- at ~68 ms decode + ~475 ms encode per 1080p->4K frame per core
(SURVEY.md section 3), more than 100x the GPU time of this engine.  This module is the drop-in for that leg:

    ffmpeg -i in.mkv -f rawvideo -pix_fmt bgr24 - \\
      | python -m upscale_video_b200.raw_stream -s 2 --width 1920 --height 1080 [-m a] [-g 0] \\
      | ffmpeg -f rawvideo -pix_fmt bgr24 -s 3840x2160 -r 24 -i - -c:v ... out.mkv

Frames are read in chunks, pushed through ``Engine.run_batch_host`` (pinned staging, H2D / network / D2H overlapped on
three streams) and written out in order.  Arithmetic is exactly ``upscale_image``'s: reference tiling (960 + 10),
``* 255``, ``cv2.imwrite`` rounding -- only the container around the pixels changes.  ``-m a`` chains the 1x
HurrDeblur pre-pass in front of the upscaler like ``process_file`` does with ``process_model`` (:888-909), keeping the
reference's u8 quantisation between the two networks but never leaving the GPU (no ``N.anime.png`` hop).  ``-m n=K``
puts the NL-means denoise pass (``apply_denoise`` :350-362, bit-identical to cv2) in front of both, in the reference's
order denoise -> anime -> upscale (test_images.py:82-110).

Pixel order: the reference's network sees OpenCV's BGR (``cv2.imread`` -> ``PIXEL_BGR`` unswapped, :263-270), so
``--pix_fmt bgr24`` (default) is byte-compatible; with ``rgb24`` the channels are swapped on the way in and out.
"""
from __future__ import annotations

import argparse
import sys

import numpy as np

from . import engine as _engine
from . import ncnn_model

HURR = "x_HurrDeblur_SubCompact_nf24-nc8_244k_net_g"


def _read_frame_into(f, dst):
    """Fill the contiguous u8 array ``dst`` (one frame) straight from ``f`` -- no intermediate bytes objects; pipes may
    return short reads.  Returns False at a clean EOF; raises ValueError on a truncated frame."""
    mv = memoryview(dst).cast("B")
    n, got = len(mv), 0
    while got < n:
        k = f.readinto(mv[got:]) if hasattr(f, "readinto") else None
        if k is None:  # file-like without readinto
            b = f.read(n - got)
            k = len(b)
            mv[got:got + k] = b
        if not k:
            break
        got += k
    if got == 0:
        return False
    if got != n:
        raise ValueError("truncated input: %d bytes of a %d-byte frame" % (got, n))
    return True


class _Stage:
    """Pinned (when torch + CUDA are present) or plain host buffers for one chunk of frames."""

    def __init__(self, shape):
        self.tensor = None
        try:
            import torch
            if torch.cuda.is_available():
                self.tensor = torch.empty(shape, dtype=torch.uint8).pin_memory()
                self.array = self.tensor.numpy()
                return
        except Exception:
            pass
        self.array = np.empty(shape, np.uint8)


_stage_cache = {}  # shape -> idle _Stage objects: pinned allocations are slow (~0.1 s per 100 MB), repeated calls reuse them


def _take_stage(shape):
    idle = _stage_cache.get(tuple(shape))
    return idle.pop() if idle else _Stage(shape)


def _give_stage(st):
    _stage_cache.setdefault(tuple(st.array.shape), []).append(st)


def denoise_level(models):
    """``n=K`` among the model options -> K clamped to 1..30, or None (reference test_images.py:45-52)."""
    for m in models:
        if m.startswith("n="):
            k = min(int(m.split("=")[1]), 30)
            return k if k > 0 else None
    return None


def stream(fin, fout, width, height, scale=2, models=(), gpu=0, chunk=8, pix_fmt="bgr24", model_path=None, max_frames=None,
           upscaler=None, prepass=None, denoiser=None, overlap=False):
    """Upscale raw ``height x width x 3`` u8 frames from ``fin`` to ``fout``.  Returns the number of frames written.
    ``upscaler`` / ``prepass`` / ``denoiser`` may be passed in as ready engines (tests); otherwise they are built from the
    model options.  ``overlap=True`` reads the next chunk and writes the previous one on helper threads while the engines
    work on the current one (``--overlap`` on the command line)."""
    model_path = model_path or ncnn_model.packaged_model_dir()
    level = denoise_level(models)
    if denoiser is None and level is not None:
        denoiser = _engine.Denoiser(gpu)
    if upscaler is None and scale > 1:
        # process_file :911-918: `-m r` selects <scale>x_Valar_v1 (RRDB, tcgen05 graph kernels), else <scale>x_Compact_Pretrain
        upscaler = _engine.Engine.from_files(model_path, str(scale) + ("x_Valar_v1" if "r" in models else "x_Compact_Pretrain"), gpu)
    if prepass is None and "a" in models:
        prepass = _engine.Engine.from_files(model_path, "1" + HURR, gpu)
    if upscaler is None and prepass is None and denoiser is None:
        raise ValueError("nothing to do: scale 1 and no pre-pass model")
    if denoiser is not None and level is None:
        raise ValueError("a denoiser needs its level: add 'n=K' to the model options")
    s = upscaler.scale if upscaler is not None else 1
    st_in = _Stage((chunk, height, width, 3))
    st_dn = _Stage((chunk, height, width, 3)) if (denoiser is not None and (upscaler is not None or prepass is not None)) else None
    st_mid = _Stage((chunk, height, width, 3)) if (prepass is not None and upscaler is not None) else None
    st_out = _Stage((chunk, height * s, width * s, 3))
    dev = None
    if st_mid is not None and st_in.tensor is not None:  # device-resident buffers for the chained mode
        import torch
        torch.cuda.set_device(gpu)
        dev = (torch, torch.empty((chunk, height, width, 3), dtype=torch.uint8, device="cuda"),
               torch.empty((chunk, height, width, 3), dtype=torch.uint8, device="cuda"),
               torch.empty((chunk, height * s, width * s, 3), dtype=torch.uint8, device="cuda"))
    swap = pix_fmt == "rgb24"

    def read_chunk(st_in, want):
        """Fill up to ``want`` frames of ``st_in`` from the input; returns the number read (short = end of input)."""
        n = 0
        while n < want:
            if not _read_frame_into(fin, st_in.array[n]):
                break
            if swap:
                st_in.array[n] = st_in.array[n][:, :, ::-1].copy()
            n += 1
        return n

    def process(st_in, st_out, n):
        """Denoise -> 1x pre-pass -> upscaler over the first ``n`` frames of ``st_in`` into ``st_out``."""
        src = st_in
        if denoiser is not None:  # fastNlMeansDenoisingColored(frame, None, K, K, 5, 9) on every frame first
            dst = st_dn if st_dn is not None else st_out
            denoiser.run_batch_host(src.tensor if src.tensor is not None else src.array,
                                    dst.tensor if dst.tensor is not None else dst.array, n, height, width, level)
            src = dst
        if prepass is not None and upscaler is not None and dev is not None:
            # chained on the device: frames go up once, the 1x model's u8 output (apply_model :263-288 rounding) feeds the
            # upscaler (upscale_image :489-519 tiling) from device memory, results come down once
            torch, d_in, d_mid, d_out = dev
            d_in[:n].copy_(src.tensor[:n])
            torch.cuda.synchronize()
            prepass.run_batch_device(d_in, d_mid, n, height, width, tile=0, halo=0, sync=True)
            upscaler.run_batch_device(d_mid, d_out, n, height, width, sync=True)
            st_out.tensor[:n].copy_(d_out[:n])
            torch.cuda.synchronize()
        else:
            if prepass is not None:  # 1x model over whole frames, untiled, u8 out (apply_model :263-288)
                dst = st_mid if upscaler is not None else st_out
                prepass.run_batch_host(src.tensor if src.tensor is not None else src.array,
                                       dst.tensor if dst.tensor is not None else dst.array, n, height, width, tile=0, halo=0)
                src = dst
            if upscaler is not None:  # reference tiling (upscale_image :489-519)
                upscaler.run_batch_host(src.tensor if src.tensor is not None else src.array,
                                        st_out.tensor if st_out.tensor is not None else st_out.array, n, height, width)

    def write_chunk(st_out, n):
        out = st_out.array[:n]
        fout.write(np.ascontiguousarray(out[:, :, :, ::-1]).data if swap else memoryview(out).cast("B"))

    if overlap:
        written = _pump(read_chunk, process, write_chunk, chunk, max_frames,
                        [st_in] + [_Stage((chunk, height, width, 3)) for _ in range(2)],
                        [st_out] + [_Stage((chunk, height * s, width * s, 3)) for _ in range(2)])
        fout.flush()
        return written
    written = 0
    while max_frames is None or written < max_frames:
        want = chunk if max_frames is None else min(chunk, max_frames - written)
        n = read_chunk(st_in, want)
        if n == 0:
            break
        process(st_in, st_out, n)
        write_chunk(st_out, n)
        written += n
        if n < want:
            break
    fout.flush()
    return written


class _Worker:
    """The engines of one ``-g`` entry plus its private intermediate buffers (one per worker thread of ``stream_multi``)."""

    def __init__(self, gpu, width, height, scale, models, chunk, model_path, level, make_engines=None):
        self.gpu, self.width, self.height, self.level = gpu, width, height, level
        if make_engines is not None:  # tests: (denoiser, prepass, upscaler) for this worker
            self.denoiser, self.prepass, self.upscaler = make_engines(gpu)
        else:
            self.denoiser = _engine.Denoiser(gpu) if level is not None else None
            self.upscaler = (_engine.Engine.from_files(model_path, str(scale) + ("x_Valar_v1" if "r" in models else "x_Compact_Pretrain"), gpu)
                             if scale > 1 else None)
            self.prepass = _engine.Engine.from_files(model_path, "1" + HURR, gpu) if "a" in models else None
        if self.upscaler is None and self.prepass is None and self.denoiser is None:
            raise ValueError("nothing to do: scale 1 and no pre-pass model")
        self.scale = self.upscaler.scale if self.upscaler is not None else 1
        stages = sum(e is not None for e in (self.denoiser, self.prepass, self.upscaler))
        # ping / pong between the stages, two sets: the set of a chunk whose last stage is still in flight is not reused by the next
        self.tmps = [[_Stage((chunk, height, width, 3)) for _ in range(min(2, stages - 1))] for _ in range(2)]
        self.tmp = self.tmps[0]
        self.last = self.upscaler if self.upscaler is not None else self.prepass  # the stage that may run without the host waiting
        self.can_stream = self.last is not None and hasattr(self.last, "submit_batch_host")
        self.n_submitted = 0

    def _steps(self, n, last_async):
        steps = []
        if self.denoiser is not None:
            steps.append(lambda a, b: self.denoiser.run_batch_host(a, b, n, self.height, self.width, self.level))
        if self.prepass is not None:
            if last_async and self.upscaler is None:
                steps.append(lambda a, b: self.prepass.submit_batch_host(a, b, n, self.height, self.width, tile=0, halo=0))
            else:
                steps.append(lambda a, b: self.prepass.run_batch_host(a, b, n, self.height, self.width, tile=0, halo=0))
        if self.upscaler is not None:
            if last_async:
                steps.append(lambda a, b: self.upscaler.submit_batch_host(a, b, n, self.height, self.width))
            else:
                steps.append(lambda a, b: self.upscaler.run_batch_host(a, b, n, self.height, self.width))
        return steps

    def _run(self, st_in, st_out, n, last_async, tmp):
        buf = lambda st: st.tensor if st.tensor is not None else st.array  # noqa: E731
        steps = self._steps(n, last_async)
        src, ret = st_in, None
        for i, step in enumerate(steps):
            dst = st_out if i == len(steps) - 1 else tmp[i % 2]
            ret = step(buf(src), buf(dst))
            src = dst
        return ret

    def process(self, st_in, st_out, n):
        """denoise -> 1x pre-pass -> upscaler (the reference's order, test_images.py:82-110) over ``n`` frames."""
        self._run(st_in, st_out, n, False, self.tmp)

    def submit(self, st_in, st_out, n):
        """``process`` with the LAST stage submitted instead of run (``can_stream`` workers): the earlier stages run
        synchronously -- on the same GPU as, and concurrently with, the previous chunk's last stage -- and the ticket returned is
        for ``last.wait_batch``.  At most two chunks may be in flight: each uses its own set of intermediate buffers."""
        self.n_submitted += 1
        return self._run(st_in, st_out, n, True, self.tmps[self.n_submitted & 1])


def stream_multi(fin, fout, width, height, scale=2, models=(), gpus=(0,), chunk=4, pix_fmt="bgr24", model_path=None, max_frames=None,
                 make_engines=None):
    """``stream`` over several GPUs: one worker thread per ``gpus`` entry (repeats allowed, like ``-g 0,0,1``), each with its
    own engines, taking chunks of frames from ONE dynamic queue -- the reference's frame sharding (``Pool.apply_async`` per
    frame over one worker per ``-g`` entry, upscale_processing.py:565-598) with chunks of raw frames instead of PNG file names
    -- and one writer that emits the chunks in input order.  The engine calls, pipe reads and writes all release the GIL.

    Input: a pipe is read by one reader thread; a seekable file (``fin.fileno()`` + ``seekable()``) is read by one reader
    thread PER WORKER with ``os.preadv`` at the chunks' offsets, so that the input side scales with the number of GPUs too
    (and stays overlapped with the engines).
    Every chunk gets its input AND output staging slots in sequence order before it is queued, and output slots are freed
    in sequence order by the writer, so the oldest chunk in flight can always finish (no slot deadlock).
    Returns the number of frames written; bytes out are identical to ``stream`` on one GPU."""
    import os
    import queue
    import threading
    model_path = model_path or ncnn_model.packaged_model_dir()
    models = list(models)
    level = denoise_level(models)
    gpus = list(gpus)
    nw = len(gpus)
    if nw < 1:
        raise ValueError("no GPU given")
    swap = pix_fmt == "rgb24"
    frame_bytes = height * width * 3
    direct = False
    try:
        direct = bool(fin.seekable()) and fin.fileno() >= 0 and hasattr(os, "preadv")
    except Exception:
        direct = False
    start_off = fin.tell() if direct else 0
    truncated = 0
    if direct:  # a seekable input's length is known: no speculative chunks past its end
        avail = max(0, os.fstat(fin.fileno()).st_size - start_off)
        truncated = avail % frame_bytes
        max_frames = avail // frame_bytes if max_frames is None else min(max_frames, avail // frame_bytes)
    workers, errors = [None] * nw, []
    stop = threading.Event()

    def build(i):
        try:
            workers[i] = _Worker(gpus[i], width, height, scale, models, chunk, model_path, level, make_engines)
        except BaseException as e:  # noqa: BLE001 -- re-raised on the caller's thread
            errors.append(e)

    ts = [threading.Thread(target=build, args=(i,)) for i in range(nw)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errors:
        raise errors[0]
    s = workers[0].scale
    n_slots = 3 * nw + 1  # per worker: two chunks in flight + one being read / waiting for the in-order writer
    free_in, free_out, work, toread = queue.Queue(), queue.Queue(), queue.Queue(), queue.Queue()
    all_slots = [_take_stage((chunk, height, width, 3)) for _ in range(n_slots)] + [_take_stage((chunk, height * s, width * s, 3)) for _ in range(n_slots)]
    for st in all_slots[:n_slots]:
        free_in.put(st)
    for st in all_slots[n_slots:]:
        free_out.put(st)
    done, done_cv = {}, threading.Condition()  # seq -> (out slot, n) | None (= end of stream marker)

    def get(q):
        while not stop.is_set():
            try:
                return q.get(timeout=0.1)
            except queue.Empty:
                continue
        return None

    def fail(e):
        errors.append(e)
        stop.set()
        with done_cv:
            done_cv.notify_all()

    def dispatcher():
        """Numbers the chunks, reserves their slots in order, and (pipe input) reads them."""
        seq, left = 0, max_frames
        try:
            while not stop.is_set() and (left is None or left > 0):
                st_in, st_out = get(free_in), get(free_out)
                if st_in is None or st_out is None:
                    break
                want = chunk if left is None else min(chunk, left)
                if direct:
                    n = want
                    toread.put((seq, st_in, st_out, want))  # a file reader thread fills the slot, then queues the chunk
                else:
                    n = 0
                    while n < want and _read_frame_into(fin, st_in.array[n]):
                        n += 1
                    if n == 0:
                        break
                    work.put((seq, st_in, st_out, n))
                seq += 1
                if left is not None:
                    left -= n
                if not direct and n < want:
                    break
        except BaseException as e:  # noqa: BLE001
            fail(e)
        finally:
            for _ in range(nw):
                (toread if direct else work).put(None)
            with done_cv:
                done.setdefault(("end", seq), None)
                done_cv.notify_all()

    def file_reader():
        """Seekable input: preadv the chunk at its offset into its slot (several of these run in parallel)."""
        try:
            while True:
                item = get(toread)
                if item is None:
                    break
                seq, st_in, st_out, want = item
                mv = memoryview(st_in.array).cast("B")[:want * frame_bytes]
                off, got = start_off + seq * chunk * frame_bytes, 0
                while got < len(mv):
                    k = os.preadv(fin.fileno(), [mv[got:]], off + got)
                    if k <= 0:
                        break
                    got += k
                if got != len(mv):
                    raise ValueError("short read: %d of %d bytes at offset %d" % (got, len(mv), off))
                work.put((seq, st_in, st_out, want))
        except BaseException as e:  # noqa: BLE001
            fail(e)
        finally:
            work.put(None)

    def worker(i):
        """One ``-g`` entry.  The worker keeps TWO chunks in flight on its last engine (``Engine.submit_batch_host`` /
        ``wait_batch``): the next chunk's first H2D copy (and its earlier stages) run under this chunk's network and this chunk's
        last D2H under the next one's -- but a chunk is never held back waiting for a successor that is not already queued."""
        w = workers[i]
        pending = None  # (ticket, seq, st_in, st_out, n) submitted, not yet waited for

        def publish(seq, st_in, st_out, n):
            free_in.put(st_in)
            with done_cv:
                done[seq] = (st_out, n)
                done_cv.notify_all()

        def finish(p):
            w.last.wait_batch(p[0])
            publish(*p[1:])

        try:
            while True:
                if pending is not None:
                    try:
                        item = work.get_nowait()
                    except queue.Empty:
                        p, pending = pending, None
                        finish(p)
                        continue
                else:
                    item = get(work)
                if item is None:
                    break
                seq, st_in, st_out, n = item
                if n > 0 and swap:
                    st_in.array[:n] = st_in.array[:n, :, :, ::-1].copy()
                if n > 0 and w.can_stream:
                    nxt_p = (w.submit(st_in, st_out, n), seq, st_in, st_out, n)
                    p, pending = pending, nxt_p
                    if p is not None:
                        finish(p)
                    continue
                if pending is not None:
                    p, pending = pending, None
                    finish(p)
                if n > 0:
                    w.process(st_in, st_out, n)
                publish(seq, st_in, st_out, n)
            if pending is not None:
                p, pending = pending, None
                finish(p)
        except BaseException as e:  # noqa: BLE001
            fail(e)
        finally:
            if pending is not None:  # an error elsewhere: the device may still be copying into this chunk's staging slots
                try:
                    w.last.wait_batch(pending[0])
                except BaseException:  # noqa: BLE001 -- already failing
                    pass

    threads = [threading.Thread(target=dispatcher, daemon=True)] + [threading.Thread(target=worker, args=(i,), daemon=True) for i in range(nw)]
    if direct:
        threads += [threading.Thread(target=file_reader, daemon=True) for _ in range(nw)]
    for t in threads:
        t.start()
    written, nxt = 0, 0
    try:
        while True:
            with done_cv:
                while nxt not in done and ("end", nxt) not in done and not stop.is_set():
                    done_cv.wait(timeout=0.1)
                if stop.is_set() and errors:
                    break
                if nxt not in done:
                    break  # ("end", nxt): every chunk before the end marker has been written
                st_out, n = done.pop(nxt)
            if n > 0:
                out = st_out.array[:n]
                fout.write(np.ascontiguousarray(out[:, :, :, ::-1]).data if swap else memoryview(out).cast("B"))
                written += n
            free_out.put(st_out)
            nxt += 1
    except BaseException as e:  # noqa: BLE001 -- e.g. a closed output pipe
        errors.append(e)
    finally:
        stop.set()
        for t in threads:
            t.join(timeout=5)
        if not any(t.is_alive() for t in threads):  # (a thread stuck in a pipe read keeps its slot: do not recycle then)
            for st in all_slots:
                _give_stage(st)
    if errors:
        raise errors[0]
    fout.flush()
    if truncated and (max_frames is None or written >= max_frames):
        raise ValueError("truncated input: %d bytes of a %d-byte frame" % (truncated, frame_bytes))
    return written


def _pump(read_chunk, process, write_chunk, chunk, max_frames, in_slots, out_slots):
    """Three-stage pipeline over rings of staging buffers: a reader thread fills input slots while the caller's thread runs
    the engines and a writer thread drains output slots (file I/O and the ctypes engine calls all release the GIL).  Frames
    leave in input order; an exception on any stage (e.g. a truncated frame) stops the others and is re-raised here."""
    import queue
    import threading
    free_in, filled, free_out, to_write = queue.Queue(), queue.Queue(), queue.Queue(), queue.Queue()
    for st in in_slots:
        free_in.put(st)
    for st in out_slots:
        free_out.put(st)
    errors = []
    stop = threading.Event()

    def reader():
        try:
            left = max_frames
            while not stop.is_set() and (left is None or left > 0):
                st = free_in.get()
                if st is None:
                    break
                want = chunk if left is None else min(chunk, left)
                n = read_chunk(st, want)
                if n:
                    filled.put((st, n))
                    if left is not None:
                        left -= n
                if n < want:
                    break
        except BaseException as e:  # noqa: BLE001 -- re-raised on the caller's thread
            errors.append(e)
        finally:
            filled.put(None)

    def writer():
        while True:
            item = to_write.get()
            if item is None:
                break
            st, n = item
            try:
                if not errors:
                    write_chunk(st, n)
            except BaseException as e:  # noqa: BLE001 -- e.g. a closed pipe; re-raised on the caller's thread
                errors.append(e)
                stop.set()
            free_out.put(st)  # always hand the slot back: the caller's thread must never starve on free_out

    tr, tw = threading.Thread(target=reader, daemon=True), threading.Thread(target=writer, daemon=True)
    tr.start()
    tw.start()
    written = 0
    try:
        while True:
            item = filled.get()
            if item is None:
                break
            st_in, n = item
            if errors:
                free_in.put(st_in)
                continue
            st_out = free_out.get()
            process(st_in, st_out, n)
            free_in.put(st_in)
            to_write.put((st_out, n))
            written += n
    except BaseException as e:  # noqa: BLE001
        errors.append(e)
    finally:
        stop.set()
        free_in.put(None)   # wake a reader waiting for a slot
        to_write.put(None)
        tw.join()
        tr.join(timeout=5)  # a reader blocked inside a pipe read is abandoned (daemon thread)
    if errors:
        raise errors[0]
    return written


def main(argv=None):
    ap = argparse.ArgumentParser(description="Upscale a raw bgr24/rgb24 frame stream (stdin -> stdout) on one or several GPUs")
    ap.add_argument("--width", type=int, required=True)
    ap.add_argument("--height", type=int, required=True)
    ap.add_argument("-s", "--scale", type=int, default=2, help="Scale 1, 2 or 4 (1 = pre-pass only). Default is 2.")
    ap.add_argument("-m", "--models", help="'a' adds the 1x anime touch-up model before upscaling, 'n={level}' NL-means noise reduction "
                                           "first, 'r' uses the real-life model 4x_Valar_v1 as the upscaler (like upscale_video.py -m a,n=3,r).")
    ap.add_argument("-g", "--gpus", default="0", help="GPU numbers, one worker per entry, repeats allowed (like upscale_video.py -g 0,0,1). "
                                                         "Several entries: one dynamic queue of frame chunks over all workers, output in order.")
    ap.add_argument("--chunk", type=int, default=8, help="frames per host<->device chunk")
    ap.add_argument("--pix_fmt", default="bgr24", choices=["bgr24", "rgb24"])
    ap.add_argument("--overlap", action="store_true", default=True, help="read, compute and write on three threads (default)")
    ap.add_argument("--no-overlap", dest="overlap", action="store_false", help="one thread: read, compute, write in turn")
    ap.add_argument("--model_path")
    ap.add_argument("-i", "--input", help="raw input file (default stdin)")
    ap.add_argument("-o", "--output", help="raw output file (default stdout)")
    a = ap.parse_args(argv)
    models = a.models.split(",") if a.models else []
    for m in models:
        if m not in ("a", "r") and not (m.startswith("n=") and m[2:].lstrip("-").isdigit()):
            sys.exit("unknown model option %r ('a' = HurrDeblur pre-pass, 'n=K' = denoise, 'r' = Valar upscaler)" % m)
    if "r" in models and a.scale != 4:
        sys.exit("-m r needs -s 4 (the reference ships 4x_Valar_v1 only)")
    fin = open(a.input, "rb") if a.input else sys.stdin.buffer
    fout = open(a.output, "wb") if a.output else sys.stdout.buffer
    try:
        gpus = [int(g) for g in a.gpus.split(",")]
    except ValueError:
        sys.exit("Invalid gpus")
    if len(gpus) > 1 or a.overlap:
        # (one GPU too: stream_multi's worker keeps two chunks in flight on the engine, measured 388 against 374 fps for `stream`)
        n = stream_multi(fin, fout, a.width, a.height, a.scale, models, gpus, a.chunk if len(gpus) == 1 else min(a.chunk, 4), a.pix_fmt,
                         a.model_path)
    else:
        n = stream(fin, fout, a.width, a.height, a.scale, models, gpus[0], a.chunk, a.pix_fmt, a.model_path, overlap=a.overlap)
    print("raw_stream: %d frames" % n, file=sys.stderr)


if __name__ == "__main__":
    main()
