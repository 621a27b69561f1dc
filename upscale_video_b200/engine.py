"""ctypes binding of libb2sr.so (C ABI: include/b2sr.h) -- the object the worker functions hold instead of the
reference's process-global ``ncnn.Net`` (reference upscale/upscale_processing.py:22,57-73).

There is deliberately no CPU implementation behind this class: if the shared library is missing or no sm_100
device is visible, construction raises :class:`EngineError`.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import ncnn_model

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb2sr.so")

MEM_HOST, MEM_DEVICE = 0, 1
OPT_IMPL, OPT_PROFILE, OPT_MAX_BATCH, OPT_RING_ROWS, OPT_PIPE_DEBUG, OPT_SM_LIMIT, OPT_SEG_PIPE, OPT_PAIR2, OPT_ABLATE = 1, 2, 3, 4, 5, 6, 7, 8, 9
STAT_LAUNCHES, STAT_TC_LAUNCHES, STAT_TC_MID_MS, STAT_TC_MID_COUNT, STAT_ALL_MS, STAT_TC_MID_PIXELS = 1, 2, 3, 4, 5, 6
STAT_PIPE_LAUNCHES, STAT_PIPE_MS, STAT_HMMA_LAUNCHES, STAT_PIPE_FALLBACKS = 7, 8, 9, 10
IMPL_AUTO, IMPL_SIMPLE, IMPL_TCGEN05, IMPL_PIPELINED = 0, 1, 2, 3  # IMPL_TCGEN05 = tcgen05 kernels launched layer by layer

# every symbol include/b2sr.h declares (tests check the library exports exactly these)
SYMBOLS = [
    "b2sr_abi_version", "b2sr_device_count", "b2sr_default_device", "b2sr_device_name", "b2sr_create", "b2sr_create_graph",
    "b2sr_create_fused", "b2sr_fused_describe_segments", "b2sr_debug_fused", "b2sr_destroy",
    "b2sr_run_u8", "b2sr_run_f32", "b2sr_run_batch_device", "b2sr_run_batch_host", "b2sr_submit_batch_host", "b2sr_wait_batch",
    "b2sr_debug_layer",
    "b2sr_set_option", "b2sr_get_stat", "b2sr_reset_stats", "b2sr_synchronize", "b2sr_stream", "b2sr_last_error",
    "b2sr_bcast_weights", "b2sr_nccl_unique_id", "b2sr_nccl_comm_init", "b2sr_nccl_comm_destroy",
    "b2sr_nlm_create", "b2sr_nlm_destroy", "b2sr_nlm_run_u8", "b2sr_nlm_run_batch_device", "b2sr_nlm_run_batch_host",
    "b2sr_nlm_synchronize", "b2sr_nlm_stream", "b2sr_nlm_launches", "b2sr_nlm_weight_table", "b2sr_nlm_lab_tables",
]


class EngineError(RuntimeError):
    pass


class GraphOp(ctypes.Structure):  # b2sr_graph_op
    _fields_ = [("type", ctypes.c_int32), ("nin", ctypes.c_int32), ("inp", ctypes.c_int32 * 6), ("out", ctypes.c_int32),
                ("cin", ctypes.c_int32), ("cout", ctypes.c_int32), ("k", ctypes.c_int32), ("act", ctypes.c_int32),
                ("slope", ctypes.c_float), ("coef", ctypes.c_float * 2), ("plain", ctypes.c_int32), ("r", ctypes.c_int32),
                ("w_off", ctypes.c_int64), ("b_off", ctypes.c_int64),
                ("in_c", ctypes.c_int32 * 6), ("in_off", ctypes.c_int32 * 6), ("in_ld", ctypes.c_int32 * 6),
                ("out_c", ctypes.c_int32), ("out_off", ctypes.c_int32), ("out_ld", ctypes.c_int32),
                ("in_res", ctypes.c_int32), ("out_res", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class FusedBuf(ctypes.Structure):  # b2sr_fused_buf
    _fields_ = [("channels", ctypes.c_int32), ("dtype", ctypes.c_int32), ("res", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class FusedOp(ctypes.Structure):  # b2sr_fused_op
    _fields_ = [("type", ctypes.c_int32), ("res", ctypes.c_int32), ("in_buf", ctypes.c_int32), ("in_off", ctypes.c_int32),
                ("cin", ctypes.c_int32), ("k", ctypes.c_int32), ("cout", ctypes.c_int32), ("act", ctypes.c_int32),
                ("slope", ctypes.c_float), ("nres", ctypes.c_int32), ("w_off", ctypes.c_int64), ("b_off", ctypes.c_int64),
                ("res_buf", ctypes.c_int32 * 2), ("res_off", ctypes.c_int32 * 2), ("coef_v", ctypes.c_float * 2),
                ("coef_r", ctypes.c_float * 2), ("out16_buf", ctypes.c_int32), ("out16_off", ctypes.c_int32),
                ("out32_buf", ctypes.c_int32), ("out32_off", ctypes.c_int32), ("r", ctypes.c_int32), ("final", ctypes.c_int32),
                ("sc_cin", ctypes.c_int32), ("sc_coef_v", ctypes.c_float), ("sc_coef_r", ctypes.c_float), ("reserved", ctypes.c_int32),
                ("sc_w_off", ctypes.c_int64)]


class NetDesc(ctypes.Structure):
    _fields_ = [("family", ctypes.c_int32), ("cin", ctypes.c_int32), ("nf", ctypes.c_int32), ("n_mid", ctypes.c_int32),
                ("scale", ctypes.c_int32), ("reserved", ctypes.c_int32 * 11)]


_lib = None


def load_library(path: str = LIB_PATH):
    """dlopen libb2sr.so and declare prototypes.  Never falls back to anything else."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise EngineError("libb2sr.so not built (%s); run `python -m upscale_video_b200.build`" % path)
    lib = ctypes.CDLL(path)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    lib.b2sr_abi_version.restype = i32
    lib.b2sr_device_count.restype = i32
    lib.b2sr_default_device.restype = i32
    lib.b2sr_device_name.argtypes = [i32, ctypes.c_char_p, i32]
    lib.b2sr_create.argtypes = [ctypes.POINTER(vp), i32, vp, ctypes.c_size_t, ctypes.POINTER(NetDesc)]
    lib.b2sr_create_graph.argtypes = [ctypes.POINTER(vp), i32, ctypes.POINTER(GraphOp), i32, i32, i32, i32, i32, vp, ctypes.c_size_t]
    lib.b2sr_create_fused.argtypes = [ctypes.POINTER(vp), i32, ctypes.POINTER(FusedOp), i32, ctypes.POINTER(FusedBuf), i32, i32, vp, ctypes.c_size_t]
    lib.b2sr_fused_describe_segments.argtypes = [ctypes.POINTER(FusedOp), i32, ctypes.POINTER(FusedBuf), i32, i32, vp, i32]
    lib.b2sr_debug_fused.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    lib.b2sr_destroy.argtypes = [vp]
    lib.b2sr_destroy.restype = None
    lib.b2sr_run_u8.argtypes = [vp, vp, i32, i32, i32, vp, i32, i32, i32, i32]
    lib.b2sr_run_f32.argtypes = [vp, vp, i32, i32, i32, vp, i32, i32, i32, i32]
    lib.b2sr_run_batch_device.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32]
    lib.b2sr_run_batch_host.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32]
    lib.b2sr_submit_batch_host.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, ctypes.POINTER(ctypes.c_uint64)]
    lib.b2sr_wait_batch.argtypes = [vp, ctypes.c_uint64]
    lib.b2sr_debug_layer.argtypes = [vp, vp, i32, i32, i32, vp]
    lib.b2sr_set_option.argtypes = [vp, i32, i64]
    lib.b2sr_get_stat.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_double)]
    lib.b2sr_reset_stats.argtypes = [vp]
    lib.b2sr_synchronize.argtypes = [vp]
    lib.b2sr_stream.argtypes = [vp]
    lib.b2sr_stream.restype = vp
    lib.b2sr_last_error.restype = ctypes.c_char_p
    lib.b2sr_bcast_weights.argtypes = [vp, vp, i32]
    lib.b2sr_nccl_unique_id.argtypes = [vp]
    lib.b2sr_nccl_comm_init.argtypes = [ctypes.POINTER(vp), i32, i32, vp, i32]
    lib.b2sr_nccl_comm_destroy.argtypes = [vp]
    f32 = ctypes.c_float
    lib.b2sr_nlm_create.argtypes = [ctypes.POINTER(vp), i32]
    lib.b2sr_nlm_destroy.argtypes = [vp]
    lib.b2sr_nlm_destroy.restype = None
    lib.b2sr_nlm_run_u8.argtypes = [vp, vp, i32, i32, i32, vp, i32, f32, f32, i32, i32, i32]
    lib.b2sr_nlm_run_batch_device.argtypes = [vp, vp, vp, i32, i32, i32, f32, f32, i32]
    lib.b2sr_nlm_run_batch_host.argtypes = [vp, vp, vp, i32, i32, i32, f32, f32]
    lib.b2sr_nlm_synchronize.argtypes = [vp]
    lib.b2sr_nlm_stream.argtypes = [vp]
    lib.b2sr_nlm_stream.restype = vp
    lib.b2sr_nlm_launches.argtypes = [vp]
    lib.b2sr_nlm_launches.restype = ctypes.c_double
    lib.b2sr_nlm_weight_table.argtypes = [f32, i32, vp, i32]
    lib.b2sr_nlm_lab_tables.argtypes = [vp, vp, vp, vp, vp]
    if lib.b2sr_abi_version() != 1:
        raise EngineError("libb2sr.so ABI version %d, expected 1" % lib.b2sr_abi_version())
    _lib = lib
    return lib


def _check(rc, what):
    if rc != 0:
        raise EngineError("%s failed (%d): %s" % (what, rc, load_library().b2sr_last_error().decode(errors="replace")))


def device_count() -> int:
    """``ncnn.get_gpu_count()`` (reference test_gpus.py:47)."""
    return load_library().b2sr_device_count()


def default_device() -> int:
    """``ncnn.get_default_gpu_index()`` (reference test_gpus.py:53)."""
    return load_library().b2sr_default_device()


def device_name(dev: int) -> str:
    """``ncnn.get_gpu_info(i).device_name()`` (reference test_gpus.py:59-66)."""
    buf = ctypes.create_string_buffer(256)
    _check(load_library().b2sr_device_name(dev, buf, 256), "b2sr_device_name")
    return buf.value.decode()


def _ptr(a):
    """Host numpy array, torch tensor (host or cuda) or raw integer address -> (void*, is_device)."""
    if isinstance(a, np.ndarray):
        return ctypes.c_void_p(a.ctypes.data), False
    if isinstance(a, int):
        return ctypes.c_void_p(a), True
    if hasattr(a, "data_ptr"):
        return ctypes.c_void_p(a.data_ptr()), bool(getattr(a, "is_cuda", False))
    raise TypeError("unsupported buffer type %r" % type(a))


def _fused_arrays(prog):
    """ctypes arrays of b2sr_fused_op / b2sr_fused_buf for a ``ncnn_model.FusedProgram``."""
    ops = (FusedOp * len(prog.ops))()
    for a, o in zip(ops, prog.ops):
        for k in ("type", "res", "in_buf", "in_off", "cin", "k", "cout", "act", "slope", "nres", "w_off", "b_off",
                  "out16_buf", "out16_off", "out32_buf", "out32_off", "r", "final", "sc_cin", "sc_coef_v", "sc_coef_r", "sc_w_off"):
            setattr(a, k, o[k])
        for q in range(2):
            a.res_buf[q], a.res_off[q], a.coef_v[q], a.coef_r[q] = o["res_buf"][q], o["res_off"][q], o["coef_v"][q], o["coef_r"][q]
    bufs = (FusedBuf * len(prog.bufs))()
    for a, b in zip(bufs, prog.bufs):
        a.channels, a.dtype, a.res = b["channels"], b["dtype"], b["res"]
    return ops, bufs


def fused_segments(prog, sms: int = 148):
    """Host-only: the persistent segments ``b2sr_create_fused`` would form for ``prog`` on a device with ``sms`` SMs
    (list of dicts; see b2sr_fused_describe_segments in include/b2sr.h).  Needs no GPU."""
    lib = load_library()
    ops, bufs = _fused_arrays(prog)
    n = lib.b2sr_fused_describe_segments(ops, len(prog.ops), bufs, len(prog.bufs), sms, None, 0)
    if n < 0:
        _check(n, "b2sr_fused_describe_segments")
    out = np.zeros(n, np.int32)
    lib.b2sr_fused_describe_segments(ops, len(prog.ops), bufs, len(prog.bufs), sms, out.ctypes.data, n)
    v, pos, segs = out.tolist(), 1, []
    for _ in range(v[0]):
        ob, oe, ns, ni = v[pos:pos + 4]
        pos += 4
        stages = []
        for _ in range(ns):
            r = v[pos:pos + 13]
            pos += 13
            stages.append(dict(op=r[0], half=r[1], variant=r[2], in_inst=r[3], grp_ring=r[4:7], out16_inst=r[7], out32_inst=r[8],
                               res_inst=r[9:11], gate_op=r[11], bp_op=r[12]))
        inst = []
        for _ in range(ni):
            inst.append(dict(buf=v[pos], last_reader=v[pos + 1]))
            pos += 2
        segs.append(dict(op_begin=ob, op_end=oe, stages=stages, inst=inst))
    return segs


class Engine:
    """One network bound to one GPU (one per worker process, like the reference's ``net``)."""

    def __init__(self, graph: ncnn_model.Graph = None, device: int = 0, packed=None, generic: bool = False, program=None):
        """``graph``: a loaded model; or ``packed`` = (CompactDesc, fp32 blob) as produced by
        ``ncnn_model.pack_compact_blob`` (what a rank receives from the start-up weight broadcast).
        SRVGGNetCompact graphs run on the Compact tcgen05 kernels; graphs that lower to fused convolutions (the RRDB
        4x_Valar_v1) on the tcgen05 graph kernels (b2sr_create_fused); anything else, or any graph when
        ``generic=True`` (cross-checks), on the generic op-by-op graph engine (b2sr_create_graph).  ``program`` = a
        ``ncnn_model.FusedProgram`` (what a rank receives when the broadcast model is an RRDB graph)."""
        self._h = None
        self._lib = load_library()
        self.generic = False
        self.fused = False
        self.program = None
        if program is not None:  # a FusedProgram received from the start-up broadcast (parallel.broadcast_packed_model)
            self._init_fused(program, device)
            return
        if packed is None and not generic and ncnn_model.compact_desc(graph) is None:
            prog = ncnn_model.compile_fused(graph)
            if prog is not None:
                self._init_fused(prog, device)
                return
        if packed is None and (generic or ncnn_model.compact_desc(graph) is None):
            self._init_generic(graph, device)
            return
        desc, blob = packed if packed is not None else ncnn_model.pack_compact_blob(graph)
        blob = np.ascontiguousarray(blob, np.float32)
        self.desc = desc
        self.scale = desc.scale
        self.device = device
        nd = NetDesc(family=ncnn_model.FAMILY_COMPACT, cin=desc.cin, nf=desc.nf, n_mid=desc.n_mid, scale=desc.scale)
        h = ctypes.c_void_p()
        _check(self._lib.b2sr_create(ctypes.byref(h), device, blob.ctypes.data, blob.nbytes, ctypes.byref(nd)), "b2sr_create")
        self._h = h

    def _init_fused(self, prog, device):
        """RRDB-style graphs: every convolution (+ bias, LeakyReLU, residual adds) on the tcgen05 graph kernel."""
        ops, bufs = _fused_arrays(prog)
        w = np.ascontiguousarray(prog.weights, np.float32)
        h = ctypes.c_void_p()
        _check(self._lib.b2sr_create_fused(ctypes.byref(h), device, ops, len(prog.ops), bufs, len(prog.bufs), prog.scale,
                                           w.ctypes.data, w.nbytes), "b2sr_create_fused")
        self._h = h
        self.desc = ncnn_model.CompactDesc(3, 0, 0, prog.scale, 3 * prog.scale * prog.scale, prog.input_blob, prog.output_blob)
        self.scale = prog.scale
        self.device = device
        self.fused = True
        self.program = prog

    def debug_fused(self, img: np.ndarray, upto: int, buf: int) -> np.ndarray:
        """Bring-up: buffer ``buf`` after ops [0, upto] of the fused program on one untiled image."""
        img = np.ascontiguousarray(img, np.uint8)
        h, w, _ = img.shape
        b = self.program.bufs[buf]
        out = np.empty((h * b["res"], w * b["res"], b["channels"]), np.float32)
        _check(self._lib.b2sr_debug_fused(self._h, img.ctypes.data, h, w, upto, buf, out.ctypes.data), "b2sr_debug_fused")
        return out

    def _init_generic(self, graph, device):
        """Graphs that are not SRVGGNetCompact (4x_Valar_v1): the generic graph engine (b2sr_create_graph)."""
        prog = ncnn_model.compile_graph(graph)
        arr = (GraphOp * len(prog.ops))()
        for a, o in zip(arr, prog.ops):
            a.type, a.nin, a.out, a.cin, a.cout, a.k, a.act = o["type"], o["nin"], o["out"], o["cin"], o["cout"], o["k"], o["act"]
            for j, v in enumerate(o["in"]):
                a.inp[j], a.in_c[j], a.in_off[j], a.in_ld[j] = v, o["in_c"][j], o["in_off"][j], o["in_ld"][j]
            a.out_c, a.out_off, a.out_ld, a.in_res, a.out_res = o["out_c"], o["out_off"], o["out_ld"], o["in_res"], o["out_res"]
            a.slope, a.plain, a.r, a.w_off, a.b_off = o["slope"], o["plain"], o["r"], o["w_off"], o["b_off"]
            a.coef[0], a.coef[1] = o["coef"]
        w = np.ascontiguousarray(prog.weights, np.float32)
        h = ctypes.c_void_p()
        _check(self._lib.b2sr_create_graph(ctypes.byref(h), device, arr, len(prog.ops), prog.n_slots, prog.in_slot, prog.out_slot,
                                           prog.scale, w.ctypes.data, w.nbytes), "b2sr_create_graph")
        self._h = h
        self.desc = ncnn_model.CompactDesc(3, 0, 0, prog.scale, 3 * prog.scale * prog.scale, prog.input_blob, prog.output_blob)
        self.scale = prog.scale
        self.device = device
        self.generic = True

    @classmethod
    def from_files(cls, model_path: str, stem: str, device: int = 0) -> "Engine":
        return cls(ncnn_model.load_model(model_path, stem), device)

    def close(self):
        if self._h is not None:
            self._lib.b2sr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- single frame -------------------------------------------------------------------------
    def run_u8(self, img: np.ndarray, tile: int = 960, halo: int = 10) -> np.ndarray:
        """u8 HWC frame -> u8 HWC frame, the image the reference's ``cv2.imwrite(output)`` stores
        (upscale_image :487-519 with tile=960/halo=10; apply_model :263-288 with tile=0)."""
        img = np.ascontiguousarray(img, np.uint8)
        h, w, ch = img.shape
        assert ch == 3
        out = np.empty((h * self.scale, w * self.scale, 3), np.uint8)
        _check(self._lib.b2sr_run_u8(self._h, img.ctypes.data, h, w, 0, out.ctypes.data, 0, tile, halo, MEM_HOST), "b2sr_run_u8")
        return out

    def run_f32(self, img: np.ndarray, tile: int = 960, halo: int = 10) -> np.ndarray:
        """The float canvas before imwrite (``output_tile * 255`` scattered into ``output``, :462-477)."""
        img = np.ascontiguousarray(img, np.uint8)
        h, w, ch = img.shape
        assert ch == 3
        out = np.empty((h * self.scale, w * self.scale, 3), np.float32)
        _check(self._lib.b2sr_run_f32(self._h, img.ctypes.data, h, w, 0, out.ctypes.data, 0, tile, halo, MEM_HOST), "b2sr_run_f32")
        return out

    # ---- batches ------------------------------------------------------------------------------
    def run_batch_device(self, d_in, d_out, n: int, h: int, w: int, tile: int = 960, halo: int = 10, sync: bool = False):
        """Device-resident packed frames (torch cuda uint8 tensors or raw device addresses)."""
        pi, _ = _ptr(d_in)
        po, _ = _ptr(d_out)
        _check(self._lib.b2sr_run_batch_device(self._h, pi, po, n, h, w, tile, halo, int(sync)), "b2sr_run_batch_device")

    def run_batch_host(self, h_in, h_out, n: int, h: int, w: int, tile: int = 960, halo: int = 10):
        """Host frames (ideally pinned) through the double-buffered H2D -> network -> D2H pipeline."""
        pi, _ = _ptr(h_in)
        po, _ = _ptr(h_out)
        _check(self._lib.b2sr_run_batch_host(self._h, pi, po, n, h, w, tile, halo), "b2sr_run_batch_host")

    def submit_batch_host(self, h_in, h_out, n: int, h: int, w: int, tile: int = 960, halo: int = 10) -> int:
        """``run_batch_host`` without waiting: returns a ticket for ``wait_batch``.  Keep ``h_in`` / ``h_out`` (pinned) alive and
        untouched until then; two submissions in flight on two buffer pairs hide every copy behind the network."""
        pi, _ = _ptr(h_in)
        po, _ = _ptr(h_out)
        ticket = ctypes.c_uint64(0)
        _check(self._lib.b2sr_submit_batch_host(self._h, pi, po, n, h, w, tile, halo, ctypes.byref(ticket)), "b2sr_submit_batch_host")
        return int(ticket.value)

    def wait_batch(self, ticket: int):
        """Returns once the submission's last output frame is in its host buffer (submissions complete in order)."""
        _check(self._lib.b2sr_wait_batch(self._h, ticket), "b2sr_wait_batch")

    # ---- bring-up / measurement ---------------------------------------------------------------
    def debug_layer(self, img: np.ndarray, layer: int) -> np.ndarray:
        img = np.ascontiguousarray(img, np.uint8)
        h, w, _ = img.shape
        out = np.empty((h, w, self.desc.nf), np.float32)
        _check(self._lib.b2sr_debug_layer(self._h, img.ctypes.data, h, w, layer, out.ctypes.data), "b2sr_debug_layer")
        return out

    def set_option(self, key: int, value: int):
        _check(self._lib.b2sr_set_option(self._h, key, value), "b2sr_set_option")

    def stat(self, key: int) -> float:
        v = ctypes.c_double()
        _check(self._lib.b2sr_get_stat(self._h, key, ctypes.byref(v)), "b2sr_get_stat")
        return v.value

    def reset_stats(self):
        _check(self._lib.b2sr_reset_stats(self._h), "b2sr_reset_stats")

    def synchronize(self):
        _check(self._lib.b2sr_synchronize(self._h), "b2sr_synchronize")

    @property
    def stream(self) -> int:
        return int(self._lib.b2sr_stream(self._h) or 0)

    def bcast_weights(self, comm: "NcclComm", root: int = 0):
        """Make this engine's device-side parameters equal rank ``root``'s (``b2sr_bcast_weights``): the start-up weight
        broadcast for ranks that did not read the model files."""
        _check(self._lib.b2sr_bcast_weights(self._h, comm.handle, root), "b2sr_bcast_weights")


class NcclComm:
    """An ncclComm_t made through the C ABI's helpers (``b2sr_nccl_unique_id`` / ``b2sr_nccl_comm_init``) -- for callers that do
    not own a communicator.  ``exchange(id_bytes_or_None) -> id_bytes`` ships rank 0's 128-byte id to every rank (e.g.
    a torch.distributed broadcast over gloo, a file, an environment variable)."""

    def __init__(self, rank: int, world: int, device: int, exchange):
        lib = load_library()
        buf = ctypes.create_string_buffer(128)
        if rank == 0:
            _check(lib.b2sr_nccl_unique_id(buf), "b2sr_nccl_unique_id")
        ident = exchange(bytes(buf.raw) if rank == 0 else None)
        h = ctypes.c_void_p()
        _check(lib.b2sr_nccl_comm_init(ctypes.byref(h), world, rank, bytes(ident), device), "b2sr_nccl_comm_init")
        self.handle, self._lib = h, lib

    def close(self):
        if self.handle:
            self._lib.b2sr_nccl_comm_destroy(self.handle)
            self.handle = None


class Denoiser:
    """``cv2.fastNlMeansDenoisingColored(img, None, h, hColor, 5, 9)`` on one GPU -- the object the denoise worker
    holds (reference upscale/upscale_processing.py:350-362).  Bit-identical to cv2's CPU result; no CPU path."""

    def __init__(self, device: int = 0):
        self._h = None
        self._lib = load_library()
        h = ctypes.c_void_p()
        _check(self._lib.b2sr_nlm_create(ctypes.byref(h), device), "b2sr_nlm_create")
        self._h = h
        self.device = device

    def close(self):
        if self._h is not None:
            self._lib.b2sr_nlm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run_u8(self, img: np.ndarray, h: float, h_color: float = None, template_window: int = 5, search_window: int = 9) -> np.ndarray:
        """BGR u8 HWC frame -> denoised BGR u8 HWC frame (arguments as cv2's, window sizes 5 / 9 only)."""
        img = np.ascontiguousarray(img, np.uint8)
        hh, ww, ch = img.shape
        assert ch == 3
        out = np.empty_like(img)
        _check(self._lib.b2sr_nlm_run_u8(self._h, img.ctypes.data, hh, ww, ww * 3, out.ctypes.data, ww * 3, float(h),
                                         float(h if h_color is None else h_color), template_window, search_window, MEM_HOST),
               "b2sr_nlm_run_u8")
        return out

    def run_batch_device(self, d_in, d_out, n: int, h: int, w: int, level: float, level_color: float = None, sync: bool = False):
        """n packed device-resident frames (torch cuda uint8 tensors or raw device addresses), one launch."""
        pi, _ = _ptr(d_in)
        po, _ = _ptr(d_out)
        _check(self._lib.b2sr_nlm_run_batch_device(self._h, pi, po, n, h, w, float(level),
                                                   float(level if level_color is None else level_color), int(sync)),
               "b2sr_nlm_run_batch_device")

    def run_batch_host(self, h_in, h_out, n: int, h: int, w: int, level: float, level_color: float = None):
        """n packed host frames (ideally pinned) through the double-buffered H2D -> kernel -> D2H pipeline."""
        pi, _ = _ptr(h_in)
        po, _ = _ptr(h_out)
        _check(self._lib.b2sr_nlm_run_batch_host(self._h, pi, po, n, h, w, float(level),
                                                 float(level if level_color is None else level_color)),
               "b2sr_nlm_run_batch_host")

    def synchronize(self):
        _check(self._lib.b2sr_nlm_synchronize(self._h), "b2sr_nlm_synchronize")

    @property
    def stream(self) -> int:
        return int(self._lib.b2sr_nlm_stream(self._h) or 0)

    @property
    def launches(self) -> int:
        return int(self._lib.b2sr_nlm_launches(self._h))


def nlm_weight_table(h: float, channels: int) -> np.ndarray:
    """Host-only: the fixed-point weight table the kernel would use for level ``h`` (no device needed)."""
    lib = load_library()
    n = lib.b2sr_nlm_weight_table(float(h), channels, None, 0)
    if n < 0:
        _check(n, "b2sr_nlm_weight_table")
    out = np.empty(n, np.int32)
    lib.b2sr_nlm_weight_table(float(h), channels, out.ctypes.data, n)
    return out


def nlm_lab_tables() -> dict:
    """Host-only: the Lab conversion tables the kernel uses."""
    t = dict(fwd=np.empty(9, np.int32), inv=np.empty(9, np.int32), l2y=np.empty(256, np.int32), l2fy=np.empty(256, np.int32),
             cbrt=np.empty(3072, np.uint16))
    _check(load_library().b2sr_nlm_lab_tables(t["fwd"].ctypes.data, t["inv"].ctypes.data, t["l2y"].ctypes.data,
                                              t["l2fy"].ctypes.data, t["cbrt"].ctypes.data), "b2sr_nlm_lab_tables")
    return t
