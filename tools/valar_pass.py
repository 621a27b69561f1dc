import sys, numpy as np
sys.path.insert(0, ".")
from upscale_video_b200 import engine as E, ncnn_model
eng = E.Engine.from_files(ncnn_model.packaged_model_dir(), "4x_Valar_v1", 0)
img = np.random.default_rng(0).integers(0,256,(540,960,3),dtype=np.uint8)
eng.run_u8(img)
