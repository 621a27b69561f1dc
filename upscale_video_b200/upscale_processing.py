"""Drop-in for the frame-upscale worker loop of the reference's ``upscale/upscale_processing.py``.

Same function names, argument meaning, return values and error behaviour as the reference for the functions on
the hot path:

    get_frames           reference :27-37
    logging_callback     reference :40-51
    init_worker          reference :54-73     (process-global engine instead of ``ncnn.Net``)
    apply_model          reference :258-299   (whole-frame 1x model, e.g. HurrDeblur ``-m a``)
    process_model        reference :302-347
    apply_denoise        reference :350-362   (``-m n=<level>``: cv2.fastNlMeansDenoisingColored on the GPU, bit-exact)
    process_denoise      reference :365-392
    process_tile         reference :395-477
    upscale_image        reference :480-542
    upscale_frames       reference :545-601

What changed underneath: ``net`` is a :class:`upscale_video_b200.engine.Engine` (C ABI ``include/b2sr.h`` over
sm_100a CUDA).  ``upscale_image`` hands the whole frame to the engine in one call -- the 960-px tiling with its
10-px halo, the ``* 255``, the crop into the canvas and ``cv2.imwrite``'s rounding are reproduced on the device
(``b2sr_run_u8``) -- instead of looping over tiles in Python.  ``process_tile`` keeps the reference's per-tile
contract for callers that use it directly.  Errors are returned as log items, never raised (:289-293, :454-459).

The ffmpeg stages and the batch bookkeeping of ``process_file`` are outside the hot path and are not
reimplemented here.
"""
from __future__ import annotations

import logging
import math
import multiprocessing
import os
import sys
import threading

import cv2
import numpy as np

from . import engine as _engine

net = None
init_error = None  # why this worker has no engine (set by init_worker), reported by its tasks as error items
model_input_name = "input"
model_output_name = "output"

TILE_SIZE = 960  # reference :489
TILE_HALO = 10   # reference :409-427


def get_frames(x):
    """``"1,4-6"`` -> ``[1, 4, 5, 6]`` (reference :27-37)."""
    frames = []
    for part in x.split(","):
        if "-" in part:
            lo, hi = (int(v) for v in part.split("-"))
            frames.extend(range(lo, hi + 1))
        else:
            frames.append(int(part))
    return frames


_failed = False  # an error item arrived on a pool's result-handler thread; the pool functions exit after join()


def logging_callback(log_list):
    """Pool callback: log the worker's items; the first ``error`` item ends the parent (reference :40-51).

    The reference calls ``sys.exit`` right here.  Pool callbacks run on the pool's result-handler thread, where
    ``SystemExit`` only kills that thread and leaves ``pool.join()`` waiting forever (CPython 3.8+), so on that thread the
    failure is recorded instead and ``upscale_frames`` / ``process_model`` / ``process_denoise`` exit the parent as soon
    as their pool has drained; a direct call from the main thread exits immediately, like the reference."""
    global _failed
    failed = False
    for level, message in log_list:
        if level == "info":
            logging.info(message)
        elif level == "debug":
            logging.debug(message)
        elif level == "error":
            logging.error(message)
            failed = True
        if failed:
            if threading.current_thread() is threading.main_thread():
                sys.exit("Error - Exiting")
            _failed = True  # keep logging: the items after the first error carry the reason


def _exit_if_failed():
    global _failed
    if _failed:
        _failed = False
        sys.exit("Error - Exiting")


def _worker_slot(workers_used):
    """Index of this pool worker among its siblings, derived like the reference does from the parent's global
    child counter (reference :59); 0 when called outside a pool (tests, single-process use)."""
    ident = multiprocessing.current_process()._identity
    return (ident[0] - 1 - workers_used) if ident else 0


def init_worker(gpus, workers_used, model_path, model_file, scale, model_input, model_output):
    """Pool initializer: bind this worker to ``gpus[slot]`` and load ``<scale><model_file>`` once
    (reference :54-73).  ``model_input``/``model_output`` are kept for signature compatibility; the engine
    recognises the graph's input/output blobs structurally and checks the names match."""
    global net, model_input_name, model_output_name, init_error
    model_input_name = model_input
    model_output_name = model_output
    net, init_error = None, None
    gpu = _worker_slot(workers_used)
    if gpu > len(gpus) - 1:
        # The reference exits the worker here (:61-63), which makes the Pool respawn it forever.  Same message, but the
        # worker stays alive and every task it receives returns the error as log items, so the parent stops.
        logging.error("Unable to assign GPU to new worker.")
        init_error = "Unable to assign GPU to new worker."
        return
    try:
        net = _engine.Engine.from_files(model_path, str(scale) + model_file, device=int(gpus[gpu]))
    except Exception as e:  # no device, missing library, bad model file: reported by the first task, never a respawn loop
        logging.error(e)
        init_error = e
        return
    for given, found in ((model_input, net.desc.input_blob), (model_output, net.desc.output_blob)):
        if given and given != found:
            logging.warning("model blob name %r differs from the graph's %r; using the graph's", given, found)


def _release_engine():
    """What ``ncnn.destroy_gpu_instance()`` does for the reference after a failure (:292, :458)."""
    global net
    try:
        if net is not None:
            net.close()
    finally:
        net = None


def apply_model(input_file, output_file, remove):
    """Run the loaded 1x model over a whole frame, untiled (reference :258-299)."""
    logging_items = []
    img = cv2.imread(input_file)
    try:
        if net is None:
            raise _engine.EngineError(init_error or "no engine: init_worker has not run in this process")
        output = net.run_u8(img, tile=0, halo=0)
        if output_file:
            cv2.imwrite(output_file, output)
    except Exception as e:  # same breadth as the reference: any failure becomes log items
        logging_items.append(["error", "Model processing failed"])
        logging_items.append(["error", e])
        _release_engine()
        return logging_items
    if remove:
        os.remove(input_file)
    logging_items.append(["info", "Processed Model: " + str(output_file)])
    return logging_items


# ---- persistent per-GPU workers ---------------------------------------------------------------
# The reference builds a new pool -- new processes, new ncnn.Net, model files re-read, shaders rebuilt -- for every call of
# upscale_frames / process_model, i.e. once per 10-minute batch (:565-577 called from :923-947).  Here the pool of a given
# (gpus, model) is created once per parent process and reused by later calls: the workers keep their CUDA context, engine,
# TMA descriptors and scratch.  B2SR_PERSISTENT_WORKERS=0 restores a pool per call.
_pools = {}  # (gpus, model_path, model_file, scale) -> Pool


def _persistent():
    return os.environ.get("B2SR_PERSISTENT_WORKERS", "1") != "0"


def _init_pool_worker(gpus, ident_base, model_path, model_file, scale, model_input, model_output):
    """Initializer of the pools this module creates itself: like ``init_worker``, with the worker slot taken relative to
    the parent's child counter AT POOL CREATION (``ident_base``) instead of the caller's ``workers_used`` bookkeeping --
    which assumes every earlier call spawned fresh processes, no longer true once pools are reused."""
    init_worker(gpus, ident_base, model_path, model_file, scale, model_input, model_output)


def _pool(gpus, workers_used, model_path, model_file, scale, model_input, model_output):
    """(pool, reused).  `spawn`, like the reference (:321, :565): no CUDA context may be inherited from the parent."""
    key = (tuple(int(g) for g in gpus), os.path.abspath(model_path), model_file, int(scale))
    if _persistent() and key in _pools:
        return _pools[key], True
    # the parent's child counter: the next process it starts gets identity (base + 1,)
    base = next(multiprocessing.process._process_counter)
    pool = multiprocessing.get_context("spawn").Pool(
        processes=len(gpus), initializer=_init_pool_worker,
        initargs=(list(gpus), base, model_path, model_file, scale, model_input, model_output))
    if _persistent():
        _pools[key] = pool
    return pool, False


def _finish(pool, results):
    """What ``pool.close(); pool.join()`` does for the reference (:600-601): return when every task has run.  A reused
    pool stays open for the next call."""
    if _persistent():
        for r in results:
            r.wait()
    else:
        pool.close()
        pool.join()
    _exit_if_failed()


def release_workers():
    """Stop the persistent workers (also registered with ``atexit``)."""
    while _pools:
        _, pool = _pools.popitem()
        pool.close()
        pool.join()


import atexit  # noqa: E402

atexit.register(release_workers)


def process_model(frames_count, model_path, model_file, scale, model_input, model_output, input_file_tag,
                  output_file_tag, gpus, workers_used, remove=True):
    """One ``apply_model`` task per existing ``N.<input_file_tag>.png`` (reference :302-347)."""
    frames = range(1, frames_count + 1) if isinstance(frames_count, int) else frames_count
    pool, _ = _pool(gpus, workers_used, model_path, model_file, scale, model_input, model_output)
    results = []
    for frame in frames:
        input_file_name = "%s.%s.png" % (frame, input_file_tag)
        output_file_name = "%s.%s.png" % (frame, output_file_tag)
        if os.path.exists(input_file_name):
            results.append(pool.apply_async(apply_model, args=(input_file_name, output_file_name, remove), callback=logging_callback))
    _finish(pool, results)


denoiser = None  # process-global, created by the first apply_denoise task of a worker
DENOISE_MAX_WORKERS = 8  # the reference sizes this pool by CPU count because NL-means runs on the CPU there; here the
                         # workers only decode/encode PNGs around a GPU call, and each one holds a CUDA context


def _denoise_device():
    """GPU of this denoise worker.  ``process_denoise`` has no ``gpus`` argument in the reference (NL-means runs on the CPU
    there), so workers are spread round-robin over the devices listed in the environment variable ``B2SR_DENOISE_GPUS``
    (``"0,2"``; inherited by the spawned workers; ``test_images`` sets it from ``-g``) or, without it, all use device 0 --
    a worker never touches a GPU the user did not select."""
    ident = multiprocessing.current_process()._identity
    slot = (ident[0] - 1) if ident else 0
    listed = [int(g) for g in os.environ.get("B2SR_DENOISE_GPUS", "").split(",") if g.strip()]
    if listed:
        return listed[slot % len(listed)]
    return 0


def apply_denoise(input_file_name, output_file_name, denoise, remove):
    """One frame through ``fastNlMeansDenoisingColored(img, None, denoise, denoise, 5, 9)`` (reference :350-362).
    The filter runs on the GPU (``b2sr_nlm_run_u8``) in OpenCV's own fixed-point arithmetic: the PNG written is
    byte-identical to what OpenCV's CPU implementation produces (the reference passes a ``cv2.UMat``; on a machine with an
    OpenCL runtime OpenCV may take its OpenCL kernel instead, whose floating-point result is not bit-defined).  Unlike the reference, a failure is reported as error items (the convention
    of the other workers, :289-293) instead of vanishing inside ``apply_async``."""
    global denoiser
    try:
        img = cv2.imread(input_file_name)
        if denoiser is None:
            denoiser = _engine.Denoiser(device=_denoise_device())
        output = denoiser.run_u8(img, denoise, denoise, 5, 9)
        cv2.imwrite(output_file_name, output)
    except Exception as e:
        if denoiser is not None:
            denoiser.close()
        denoiser = None
        return [["error", "Denoise failed"], ["error", e]]
    if remove:
        os.remove(input_file_name)
    return [["info", "Processed Denoise: " + output_file_name]]


def process_denoise(frames_count, input_file_tag, denoise, remove=True):
    """One ``apply_denoise`` task per existing ``N.<input_file_tag>.png`` -> ``N.denoise.png``; returns the number
    of pool processes, which callers add to ``workers_used`` (reference :365-392)."""
    frames = range(1, frames_count + 1) if isinstance(frames_count, int) else frames_count
    pool = multiprocessing.get_context("spawn").Pool(processes=min(os.cpu_count() or 1, DENOISE_MAX_WORKERS))
    for frame in frames:
        input_file_name = str(frame) + "." + input_file_tag + ".png"
        output_file_name = str(frame) + ".denoise.png"
        if os.path.exists(input_file_name):
            pool.apply_async(apply_denoise, args=(input_file_name, output_file_name, denoise, remove),
                             callback=logging_callback)
    pool.close()
    pool.join()
    _exit_if_failed()
    return pool._processes


def tile_rect(y, x, tile_size, height, width, halo=TILE_HALO):
    """Input rectangle (with halo) and core rectangle of tile (y, x) -- reference :398-427.  A halo is added on
    a side only when at least ``halo`` pixels of image remain there."""
    y0, y1 = y * tile_size, min(y * tile_size + tile_size, height)
    x0, x1 = x * tile_size, min(x * tile_size + tile_size, width)
    top = halo if y0 >= halo else 0
    bottom = halo if y1 <= height - halo else 0
    left = halo if x0 >= halo else 0
    right = halo if x1 <= width - halo else 0
    return (y0 - top, y1 + bottom, x0 - left, x1 + right), (y0, y1, x0, x1)


def process_tile(img, tile_size, scale, y, x, height, width, output, logging_items):
    """Upscale tile (y, x) of ``img`` and write its core into the float canvas ``output`` (reference :395-477).
    Returns -1 (after appending error items) when the engine fails, else None."""
    (iy0, iy1, ix0, ix1), (cy0, cy1, cx0, cx1) = tile_rect(y, x, tile_size, height, width)
    input_tile = np.ascontiguousarray(img[iy0:iy1, ix0:ix1, :])
    try:
        if net is None:
            raise _engine.EngineError(init_error or "no engine: init_worker has not run in this process")
        output_tile = net.run_f32(input_tile, tile=0, halo=0)  # the tile is one zero-padded plane, `* 255` applied
    except Exception as e:
        logging_items.append(["error", "Upscale failed"])
        logging_items.append(["error", e])
        logging.error(e)
        _release_engine()
        return -1
    oy, ox = (cy0 - iy0) * scale, (cx0 - ix0) * scale
    output[cy0 * scale:cy1 * scale, cx0 * scale:cx1 * scale, :] = output_tile[
        oy:oy + (cy1 - cy0) * scale, ox:ox + (cx1 - cx0) * scale, :]
    return None


def upscale_image(input_file_name, output_file_name, scale, frame_batch, frame, end_frame, remove=True):
    """One frame: read PNG, upscale with the reference's tiling, write PNG, delete the input, return log items
    (reference :480-542).  ``output_file_name`` may be None (test_gpus.py), ``frame_batch`` None, int or list."""
    logging_items = []
    img = cv2.imread(input_file_name)
    height, width, _ = img.shape
    tiles_x = math.ceil(width / TILE_SIZE)
    tiles_y = math.ceil(height / TILE_SIZE)
    for tile_idx in range(1, tiles_x * tiles_y + 1):
        logging_items.append(["debug", f"Processing Tile: {tile_idx}/{tiles_x * tiles_y}"])
    try:
        if net is None:
            raise _engine.EngineError(init_error or "no engine: init_worker has not run in this process")
        if scale != net.scale:
            raise ValueError("engine was initialised for scale %d, upscale_image called with %d" % (net.scale, scale))
        output = net.run_u8(img, tile=TILE_SIZE, halo=TILE_HALO)  # all tiles of the frame in one device pass
    except Exception as e:
        logging_items.append(["error", "Upscale failed"])
        logging_items.append(["error", e])
        logging.error(e)
        _release_engine()
        return logging_items
    if output_file_name:
        cv2.imwrite(output_file_name, output)
    if remove:
        os.remove(input_file_name)
    if frame_batch:
        if isinstance(frame_batch, int):
            logging_items.append(["info", "Upscaling Batch: %s : Upscaled %s/%s" % (frame_batch, frame, end_frame)])
        else:
            logging_items.append(["info", "Upscaled " + str(output_file_name)])
    else:
        logging_items.append(["info", "Upscaled %s/%s" % (frame, end_frame)])
    return logging_items


def upscale_frames(frame_batch, start_frame, end_frame, input_file_tag, scale, gpus, workers_used, model_path,
                   model_file, model_input, model_output, remove=True):
    """Frame sharding: one pool worker per ``gpus`` entry, one ``upscale_image`` task per existing
    ``N.<input_file_tag>.png`` (reference :545-601).  Frames whose input is missing were finished by an earlier
    run and are skipped -- the reference's resume contract."""
    frames = frame_batch if (frame_batch and isinstance(frame_batch, list)) else range(start_frame, end_frame + 1)
    pool, _ = _pool(gpus, workers_used, model_path, model_file, scale, model_input, model_output)
    results = []
    for frame in frames:
        input_file_name = "%s.%s.png" % (frame, input_file_tag)
        output_file_name = "%s.png" % frame
        if os.path.exists(input_file_name):
            results.append(pool.apply_async(upscale_image,
                                            args=(input_file_name, output_file_name, scale, frame_batch, frame, end_frame, remove),
                                            callback=logging_callback))
    _finish(pool, results)
