"""``python -m upscale_video_b200.test_gpus [-g 0,0,1] [-s 2] [-r 10] [-i frame.png]``

GPU listing + worker-count calibration loop with the flags and log output of the reference's ``test_gpus.py``
(reference test_gpus.py:15-35 ``upscale_images``, :38-112 ``run_tests``, :115-127 CLI): enumerate devices, start a
spawn pool with one worker per ``-g`` entry (repeats = several workers on one GPU), push ``runs`` frames through
``upscale_image`` without writing output, log seconds per frame and in total.

The reference times its bundled ``sample.png``; that file belongs to the reference repo, so here ``-i`` names the
frame to use and, without it, a synthetic 1920x1278 frame (sample.png's size) is written to a temp file.

Adapted from ``test_gpus.py`` of davlee1972/upscale_video -- Copyright (c) 2022, David Lee (MIT licence) -- the command line,
flags, file naming and log lines are that tool's contract and are kept; the engine underneath is this repository's.
"""
import argparse
import logging
import multiprocessing
import os
import sys
import tempfile
import time

from . import engine
from . import ncnn_model
from .upscale_processing import init_worker, logging_callback, upscale_image


def upscale_images(input_file_name, output_file_name, scale, gpus):
    """One timed task (reference test_gpus.py:15-35)."""
    logging_items = []
    i = int(multiprocessing.current_process()._identity[0]) - 1
    start = time.time()
    logging_items.append(["info", "Testing GPU: " + str(gpus[i])])
    ret = upscale_image(input_file_name, output_file_name, scale, None, 1, 1, remove=False)
    logging_items += [item for item in ret if item[0] == "error"]  # the reference drops `ret`; errors must not vanish
    total = time.time() - start
    logging_items.append(["info", str(total) + " seconds to upscale " + os.path.basename(input_file_name)])
    return logging_items


def run_tests(gpus=None, scale=2, runs=10, input_file=None, model_path=None):
    logging.basicConfig(level=logging.INFO, format="[%(asctime)s] [%(levelname)s] %(message)s", datefmt="%Y-%m-%d %H:%M:%S",
                        stream=sys.stdout)
    gpu_count = engine.device_count()
    logging.info("Searching for CUDA (sm_100) GPUs")
    logging.info("====================================")
    logging.info("GPU count: " + str(gpu_count))
    logging.info("====================================")
    logging.info("Default GPU: " + str(engine.default_device()))
    logging.info("====================================")
    for i in range(gpu_count):
        logging.info("GPU " + str(i) + ": Discrete / " + engine.device_name(i))
    if gpus is None:
        return
    gpus = [int(g) for g in gpus.split(",")] if gpus else [0]
    model_path = model_path or ncnn_model.packaged_model_dir()
    tmp = None
    if input_file is None:
        import cv2
        import numpy as np
        tmp = tempfile.NamedTemporaryFile(suffix=".png", delete=False)
        tmp.close()
        yy, xx = np.mgrid[0:1278, 0:1920]
        img = np.stack([128 + 100 * np.sin(xx / 41.0 + c) * np.cos(yy / 29.0) for c in range(3)], -1).astype(np.uint8)
        cv2.imwrite(tmp.name, img)
        input_file = tmp.name
    pool = multiprocessing.get_context("spawn").Pool(
        processes=len(gpus), initializer=init_worker,
        initargs=(gpus, 0, model_path, "x_Compact_Pretrain", scale, "input", "output"))
    logging.info("")
    logging.info("Starting test runs")
    logging.info("====================================")
    start = time.time()
    for _ in range(runs):
        pool.apply_async(upscale_images, args=(input_file, None, scale, gpus), callback=logging_callback)
    pool.close()
    pool.join()
    total = time.time() - start
    logging.info("====================================")
    logging.info(str(total) + " seconds total to run tests.")
    if tmp is not None:
        os.remove(tmp.name)


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="Test GPU - List GPUs")
    parser.add_argument("-g", "--gpus", help="Optional gpus to test. Example 0,1,1,2. Default is 0.")
    parser.add_argument("-s", "--scale", type=int, default=2, help="Scale 2 or 4. Default is 2.")
    parser.add_argument("-r", "--runs", type=int, default=10, help="Number of tests")
    parser.add_argument("-i", "--input", help="PNG frame to time (default: synthetic 1920x1278)")
    parser.add_argument("--model_path", help="Directory with <scale>x_Compact_Pretrain.param/.bin or .b2sr (default: packaged models)")
    args = parser.parse_args()
    run_tests(args.gpus, args.scale, args.runs, args.input, args.model_path)
