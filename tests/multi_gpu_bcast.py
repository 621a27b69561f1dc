"""Launched by tests/test_zz_gpu_multi.py under torchrun (world size 2, one GPU per rank): the C ABI's start-up weight
broadcast.  Rank 0 reads the model; rank 1 creates its engine from the same description with an all-zero blob;
``b2sr_bcast_weights`` over an ncclComm_t made with the ABI's own helpers makes rank 1's engine produce rank 0's bytes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from upscale_video_b200 import engine as E, ncnn_model  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("gloo")  # only to ship the 128-byte NCCL id and compare results
mdir = ncnn_model.packaged_model_dir()
ok = True
for stem in ("2x_Compact_Pretrain", "4x_Valar_v1"):
    graph = ncnn_model.load_model(mdir, stem)
    if stem.startswith("2x"):
        desc, blob = ncnn_model.pack_compact_blob(graph)
        eng = E.Engine(device=local, packed=(desc, blob if rank == 0 else np.zeros_like(blob)))
    else:
        prog = ncnn_model.compile_fused(graph)
        if rank != 0:
            prog.weights = np.zeros_like(prog.weights)
        eng = E.Engine(device=local, program=prog)
    img = np.random.default_rng(3).integers(0, 256, (40, 300, 3), dtype=np.uint8)
    before = eng.run_u8(img)

    def exchange(ident):
        box = [ident]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    comm = E.NcclComm(rank, world, local, exchange)
    eng.bcast_weights(comm, 0)
    after = eng.run_u8(img)
    outs = [None] * world
    dist.all_gather_object(outs, (before.tobytes(), after.tobytes()))
    if rank == 0:
        same_after = all(o[1] == outs[0][1] for o in outs)
        differed_before = all(o[0] != outs[0][0] for o in outs[1:])
        print("%s: ranks agree after the broadcast: %s; rank > 0 differed before it: %s" % (stem, same_after, differed_before), flush=True)
        ok = ok and same_after and differed_before and outs[0][0] == outs[0][1]
    comm.close()
    eng.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
