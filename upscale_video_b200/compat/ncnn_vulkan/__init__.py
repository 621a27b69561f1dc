"""The slice of ``ncnn_vulkan`` the reference's ``test_gpus.py`` touches directly (:8, :47-67): device enumeration.
Everything else the reference does with ncnn happens inside the worker functions, which ``upscale.upscale_processing``
(this directory) replaces wholesale."""
from upscale_video_b200 import engine as _engine


class _GpuInfo:
    def __init__(self, index):
        self._index = index

    def type(self):
        return 0  # index into the reference's ["Discrete", "Integrated", "Virtual", "CPU"] (test_gpus.py:56)

    def device_name(self):
        return _engine.device_name(self._index)


class _Ncnn:
    @staticmethod
    def get_gpu_count():
        return _engine.device_count()

    @staticmethod
    def get_default_gpu_index():
        return _engine.default_device()

    @staticmethod
    def get_gpu_info(index):
        return _GpuInfo(index)

    @staticmethod
    def destroy_gpu_instance():
        return None


ncnn = _Ncnn()
