"""Multi-GPU plumbing for the frame-upscale path: one process per GPU, frames sharded, no data-path collective.

The reference's only parallelism is frame-level: ``upscale_frames`` hands one frame per task to a pool with one
worker per ``-g`` entry (reference upscale/upscale_processing.py:565-598) and every worker loads the model files
itself in ``init_worker`` (:70-71).  Under ``torchrun`` the equivalent is: rank r owns GPU LOCAL_RANK and the
frames ``shard_frames`` assigns to it, and the packed weight blob is read once by rank 0 and broadcast
(``broadcast_packed_model``; NCCL over NVLink on GPUs, gloo in the CPU tests).  Nothing else is communicated.
"""
from __future__ import annotations

import json

import numpy as np

from . import ncnn_model


def shard_frames(frames, rank: int, world: int):
    """Static round-robin assignment frame i -> rank i mod world (the pool's dynamic queue, made deterministic)."""
    frames = list(frames)
    return frames[rank::world]


def _encode_packed(desc: ncnn_model.CompactDesc, blob: np.ndarray) -> np.ndarray:
    head = json.dumps({"cin": desc.cin, "nf": desc.nf, "n_mid": desc.n_mid, "scale": desc.scale, "cout_last": desc.cout_last,
                       "input_blob": desc.input_blob, "output_blob": desc.output_blob}).encode()
    head += b" " * ((-len(head)) % 4)
    return np.concatenate([np.array([len(head)], np.uint32).view(np.uint8), np.frombuffer(head, np.uint8),
                           np.ascontiguousarray(blob, np.float32).view(np.uint8)])


def _encode_fused(prog: ncnn_model.FusedProgram) -> np.ndarray:
    """A fused program (RRDB graphs: op list + buffer table + fp32 weight blob, 67 MB for 4x_Valar_v1) as one byte array."""
    head = json.dumps({"fused": 1, "ops": prog.ops, "bufs": prog.bufs, "scale": prog.scale, "input_blob": prog.input_blob,
                       "output_blob": prog.output_blob}).encode()
    head += b" " * ((-len(head)) % 4)
    return np.concatenate([np.array([len(head)], np.uint32).view(np.uint8), np.frombuffer(head, np.uint8),
                           np.ascontiguousarray(prog.weights, np.float32).view(np.uint8)])


def _decode_packed(buf: np.ndarray):
    hlen = int(buf[:4].view(np.uint32)[0])
    d = json.loads(bytes(buf[4:4 + hlen]).decode())
    blob = buf[4 + hlen:].view(np.float32).copy()
    if d.get("fused"):
        return ncnn_model.FusedProgram(d["ops"], d["bufs"], d["scale"], blob, d["input_blob"], d["output_blob"])
    return ncnn_model.CompactDesc(d["cin"], d["nf"], d["n_mid"], d["scale"], d["cout_last"], d["input_blob"], d["output_blob"]), blob


def broadcast_packed_model(load_fn, rank: int, world: int, device=None, src: int = 0):
    """Returns ``(CompactDesc, fp32 blob)`` -- or, when ``load_fn`` returns a ``FusedProgram`` (4x_Valar_v1 lowered by
    ``ncnn_model.compile_fused``), that program -- on every rank; only rank ``src`` calls ``load_fn`` (= touches the files).
    ``device``: torch device of the staging tensor (cuda for the nccl backend, None/cpu for gloo).  Feed the result to
    ``Engine(device=..., packed=...)`` / ``Engine(device=..., program=...)``."""
    if world <= 1:
        return load_fn()
    import torch
    import torch.distributed as dist
    if rank == src:
        loaded = load_fn()
        payload = _encode_fused(loaded) if isinstance(loaded, ncnn_model.FusedProgram) else _encode_packed(*loaded)
        n = torch.tensor([payload.size], dtype=torch.int64, device=device)
    else:
        payload = None
        n = torch.zeros(1, dtype=torch.int64, device=device)
    dist.broadcast(n, src)
    if rank == src:
        t = torch.from_numpy(payload.copy()).to(device) if device is not None else torch.from_numpy(payload.copy())
    else:
        t = torch.empty(int(n.item()), dtype=torch.uint8, device=device)
    dist.broadcast(t, src)
    return _decode_packed(t.cpu().numpy())
