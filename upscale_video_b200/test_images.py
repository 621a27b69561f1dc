"""``python -m upscale_video_b200.test_images -i 1,3,5-7 -t <temp> -o <out> [-s 2] [-m a] [-g 0,1]``

Upscale chosen extracted frames into an output directory, with the flags and file naming of the reference's
``test_images.py`` (reference test_images.py:18-159 ``process_image``, :162-207 CLI): copies
``<temp>/upscale_video/N.extract.png`` to the output directory, optionally runs the 1x anime model (``-m a``,
HurrDeblur through ``process_model``), then ``upscale_frames``; with ``-m`` the result is renamed
``N.<models>.png``.

``-m r`` selects the 4x_Valar_v1 RRDB model (fused tcgen05 graph kernels, b2sr_create_fused); ``-m n=K`` runs the
NL-means denoise pass first (``process_denoise``, level clamped to 1..30 like reference test_images.py:45-52), on the
GPU and bit-identical to OpenCV's result.

Adapted from ``test_images.py`` of davlee1972/upscale_video -- Copyright (c) 2022, David Lee (MIT licence) -- the command line,
flags, file naming and log lines are that tool's contract and are kept; the engine underneath is this repository's.
"""
import argparse
import logging
import os
import shutil
import sys
import tempfile

from . import ncnn_model
from .upscale_processing import get_frames, process_denoise, process_model, upscale_frames


def process_image(input_frames, temp_dir, output_dir, scale, models, gpus, model_path=None):
    logging.basicConfig(level=logging.INFO, format="[%(asctime)s] [%(levelname)s] %(message)s", datefmt="%Y-%m-%d %H:%M:%S",
                        stream=sys.stdout)
    if scale not in [1, 2, 4]:
        sys.exit("Scale must be 1, 2 or 4")
    models = models.split(",") if models else []
    denoise = [m.split("=") for m in models if m.startswith("n=")]
    if denoise:
        denoise = min(int(denoise[0][1]), 30)
        if denoise <= 0:
            denoise = None
    if "r" in models:
        scale = 4  # real-life imaging model is 4x only (reference test_images.py:40-41)
    if gpus:
        try:
            gpus = [int(g) for g in gpus.split(",")]
        except ValueError:
            logging.error("Invalid gpus")
            sys.exit("Error - Exiting")
    else:
        gpus = [0]
    model_path = model_path or ncnn_model.packaged_model_dir()
    input_frames = get_frames(input_frames)
    temp_dir = os.path.abspath(os.path.join(temp_dir or tempfile.gettempdir(), "upscale_video"))
    output_dir = os.path.abspath(output_dir)
    for frame in input_frames:
        shutil.copyfile(os.path.join(temp_dir, str(frame) + ".extract.png"), os.path.join(output_dir, str(frame) + ".extract.png"))
    os.chdir(output_dir)
    workers_used = 0
    input_file_tag = "extract"
    if denoise:
        logging.info("Starting denoise touchup...")
        # process_denoise has no gpus argument in the reference (NL-means runs on the CPU there): hand the -g selection to
        # the spawned denoise workers through the environment so that `-g 1` touches GPU 1 only
        os.environ["B2SR_DENOISE_GPUS"] = ",".join(str(g) for g in sorted(set(gpus)))
        workers_used += process_denoise(input_frames, input_file_tag, denoise, remove=False)
        input_file_tag = "denoise"
    if "a" in models:
        logging.info("Starting anime touchup...")
        process_model(input_frames, model_path, "x_HurrDeblur_SubCompact_nf24-nc8_244k_net_g", 1, "input", "output",
                      input_file_tag, "anime", gpus, workers_used, remove=False)
        workers_used += len(gpus)
        input_file_tag = "anime"
    for frame in input_frames:
        try:
            os.remove(str(frame) + ".png")
        except OSError:
            pass
    if scale > 1:
        logging.info("Starting upscale processing...")
        upscale_frames(input_frames, input_frames[-1], input_frames[-1], input_file_tag, scale, gpus, workers_used, model_path,
                       "x_Valar_v1" if "r" in models else "x_Compact_Pretrain", "input", "output", remove=False)
    if models:
        for frame in input_frames:
            src = str(frame) + (".png" if scale > 1 else "." + input_file_tag + ".png")
            shutil.move(os.path.join(output_dir, src), os.path.join(output_dir, str(frame) + "." + ".".join(models) + ".png"))
    logging.info("Completed")


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="Test Image Upscaler")
    parser.add_argument("-i", "--input_frames", required=True, help="List of input frames in format like 1,3,5-7,10-12,15")
    parser.add_argument("-t", "--temp_dir", help="Temp directory where extracted frames are saved. Default is tempfile.gettempdir().")
    parser.add_argument("-o", "--output_dir", required=True, help="Output directory where test images will be saved")
    parser.add_argument("-s", "--scale", type=int, default=2, help="Scale 1, 2 or 4. Default is 2.")
    parser.add_argument("-m", "--models", help="'a' adds the 1x anime touch-up model before upscaling, 'n={denoise level}' NL-means noise reduction, 'r' uses the 4x real-life model (Valar). Example: -m a,n=3,r")
    parser.add_argument("-g", "--gpus", help="Optional gpu #s to use. Example 0,1,3. Default is 0.")
    parser.add_argument("--model_path", help="Directory with the model files (default: packaged models)")
    args = parser.parse_args()
    process_image(args.input_frames, args.temp_dir, args.output_dir, args.scale, args.models, args.gpus, args.model_path)
