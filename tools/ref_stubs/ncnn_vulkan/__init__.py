"""Stand-in for the `ncnn_vulkan` wheel, importable by name (also in `spawn`ed pool workers): used ONLY by
tools/make_ref_glue_goldens.py to run the reference's own upscale_processing.py in the build container.  The five ncnn calls
on the path are implemented there (numpy + the oracle's float32 layer interpreter)."""
import types

import make_ref_glue_goldens as _g

ncnn = types.SimpleNamespace(Net=_g._Net, Mat=_g._Mat, destroy_gpu_instance=lambda: None)
