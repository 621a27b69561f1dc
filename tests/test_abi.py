"""CPU: the C-ABI shared library builds, loads without a GPU driver, exports every symbol include/b2sr.h declares,
and fails loudly (no CPU fallback) when asked to compute without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from upscale_video_b200 import build, engine
    build.build_lib()
    return engine.load_library()


def _header_functions():
    text = open(os.path.join(ROOT, "include", "b2sr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2sr_[a-z0-9_]+)\s*\(", text)))


def test_exports_every_declared_symbol(lib):
    from upscale_video_b200 import engine
    declared = _header_functions()
    assert declared == sorted(engine.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.b2sr_abi_version() == 1


def test_no_libcuda_link_dependency():
    """The library must dlopen on a box without the NVIDIA driver (driver API is resolved at run time)."""
    import subprocess
    from upscale_video_b200 import engine
    out = subprocess.run(["ldd", engine.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libcudart" not in out and "not found" not in out


def test_net_desc_layout():
    from upscale_video_b200 import engine
    assert ctypes.sizeof(engine.NetDesc) == 16 * 4


def test_argument_validation_without_device(lib):
    from upscale_video_b200 import engine
    h = ctypes.c_void_p()
    nd = engine.NetDesc(family=99, cin=3, nf=64, n_mid=16, scale=2)
    blob = np.zeros(4, np.float32)
    assert lib.b2sr_create(ctypes.byref(h), 0, blob.ctypes.data, blob.nbytes, ctypes.byref(nd)) == -5
    assert b"family" in lib.b2sr_last_error()
    nd = engine.NetDesc(family=1, cin=3, nf=64, n_mid=16, scale=3)
    assert lib.b2sr_create(ctypes.byref(h), 0, blob.ctypes.data, blob.nbytes, ctypes.byref(nd)) == -5
    nd = engine.NetDesc(family=1, cin=3, nf=64, n_mid=16, scale=2)
    assert lib.b2sr_create(ctypes.byref(h), 0, blob.ctypes.data, blob.nbytes, ctypes.byref(nd)) == -1  # blob size
    assert h.value is None
    assert lib.b2sr_run_u8(None, None, 1, 1, 0, None, 0, 0, 0, 0) == -1


def test_no_cpu_fallback(model_dir):
    """Without a visible sm_100 device the engine must refuse to exist (never compute on the CPU)."""
    import torch
    from upscale_video_b200 import engine
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible; the refusal path is covered on the CPU box")
    assert engine.device_count() == 0 and engine.default_device() == -1
    with pytest.raises(engine.EngineError, match="no CUDA device"):
        engine.Engine.from_files(model_dir, "2x_Compact_Pretrain", 0)


def test_product_code_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under upscale_video_b200/ may reference it."""
    pkg = os.path.join(ROOT, "upscale_video_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src, f


def _fused_arrays(prog):
    from upscale_video_b200 import engine
    ops = (engine.FusedOp * len(prog.ops))()
    for a, o in zip(ops, prog.ops):
        for k in ("type", "res", "in_buf", "in_off", "cin", "k", "cout", "act", "slope", "nres", "w_off", "b_off",
                  "out16_buf", "out16_off", "out32_buf", "out32_off", "r", "final", "sc_cin", "sc_coef_v", "sc_coef_r", "sc_w_off"):
            setattr(a, k, o[k])
        for q in range(2):
            a.res_buf[q], a.res_off[q], a.coef_v[q], a.coef_r[q] = o["res_buf"][q], o["res_off"][q], o["coef_v"][q], o["coef_r"][q]
    bufs = (engine.FusedBuf * len(prog.bufs))()
    for a, b in zip(bufs, prog.bufs):
        a.channels, a.dtype, a.res = b["channels"], b["dtype"], b["res"]
    return ops, bufs


def test_fused_program_validation_without_device(lib, model_dir):
    """b2sr_create_fused checks the whole program before it touches the device: the real 4x_Valar_v1 program passes the
    checks (and then fails with NODEVICE here), broken programs are refused with INVALID / UNSUPPORTED and a message."""
    import copy
    import torch
    from upscale_video_b200 import engine, ncnn_model
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    assert ctypes.sizeof(engine.FusedOp) == 136 and ctypes.sizeof(engine.FusedBuf) == 16
    prog = ncnn_model.compile_fused(ncnn_model.load_model(model_dir, "4x_Valar_v1"))
    w = np.ascontiguousarray(prog.weights, np.float32)

    def create(p, scale=None):
        ops, bufs = _fused_arrays(p)
        h = ctypes.c_void_p()
        rc = lib.b2sr_create_fused(ctypes.byref(h), 0, ops, len(p.ops), bufs, len(p.bufs), scale or p.scale, w.ctypes.data, w.nbytes)
        assert h.value is None
        return rc, lib.b2sr_last_error().decode()

    rc, msg = create(prog)
    assert rc == -3 and "no CUDA device" in msg  # valid program: only the device is missing
    cases = []
    first = lambda pred: next(i for i, o in enumerate(prog.ops) if pred(o))  # noqa: E731
    i_res = first(lambda o: o["nres"] > 0)
    i_wide = first(lambda o: o["cin"] > 64)
    i_sc = first(lambda o: o["sc_cin"] > 0)
    i_near = first(lambda o: o["type"] == ncnn_model.FOP_NEAREST)
    bad = copy.deepcopy(prog); bad.ops[i_wide]["in_off"] = 4; cases.append((bad, "convolution input"))     # view not 16-byte aligned
    bad = copy.deepcopy(prog); bad.ops[i_wide]["cin"] = 208; cases.append((bad, "convolution input"))      # more than 3 groups of 64
    bad = copy.deepcopy(prog); bad.ops[i_res]["res_buf"][0] = 99; cases.append((bad, "bad residual"))
    bad = copy.deepcopy(prog); bad.ops[1]["out16_off"] = 176; cases.append((bad, "bad output view"))       # slice runs past the buffer
    bad = copy.deepcopy(prog); bad.ops[2]["w_off"] = len(w); cases.append((bad, "weights outside"))
    bad = copy.deepcopy(prog); bad.ops[-1]["final"] = 0; cases.append((bad, "convolution output"))
    bad = copy.deepcopy(prog); bad.ops[i_near]["r"] = 3; cases.append((bad, "nearest"))
    bad = copy.deepcopy(prog); bad.bufs[0]["dtype"] = 3; cases.append((bad, "buffer 0"))
    bad = copy.deepcopy(prog); bad.ops[i_sc]["sc_cin"] = 24; cases.append((bad, "shortcut"))               # not a multiple of 16
    bad = copy.deepcopy(prog); bad.ops[i_sc]["sc_w_off"] = len(w) - 5; cases.append((bad, "shortcut weights"))
    for p, frag in cases:
        rc, msg = create(p)
        assert rc in (-1, -5) and frag in msg, (rc, msg, frag)
    rc, msg = create(prog, scale=3)
    assert rc == -5 and "scale" in msg
