"""``upscale.upscale_processing`` as the reference's CLIs import it (test_gpus.py:10, test_images.py:10-15),
backed by the B200 engine: every name is the drop-in from ``upscale_video_b200.upscale_processing``."""
from upscale_video_b200.upscale_processing import (  # noqa: F401
    apply_denoise, apply_model, get_frames, init_worker, logging_callback, process_denoise, process_model,
    process_tile, upscale_frames, upscale_image,
)
