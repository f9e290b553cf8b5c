"""Run under torchrun on N GPUs: the row-sharded result must equal the single-GPU result bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multigpu_check.py
Every rank also runs the unsharded light field on its own GPU as the comparison."""
import faulthandler
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from remotesensingproject_b200 import api
from remotesensingproject_b200.shard import shard_table
from remotesensingproject_b200.synth import make_light_field_np


def main():
    faulthandler.enable()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    failures = 0
    cases = [dict(S=7, V=96, U=80, C=3, D=24, mode="ftc", scale=1.0), dict(S=6, V=128, U=70, C=1, D=40, mode="ftc", scale=-1.0),
             dict(S=5, V=50, U=64, C=3, D=16, mode="depth2d", scale=1.0),
             dict(S=5, V=64, U=72, C=3, D=16, mode="ftc", scale=1.0, gather="full"),
             dict(S=6, V=80, U=90, C=3, D=20, mode="ftc", scale=1.0, halo="nccl"),
             # boundary 50 is a multiple of 2 only: levels 0-1 sharded, levels 2-3 replicated on every rank
             dict(S=5, V=100, U=90, C=3, D=16, mode="ftc", scale=-1.0),
             dict(S=4, V=100, U=90, C=1, D=24, mode="ftc", scale=1.0, u8=True),
             # lock-step row blocks (every rank evaluates its own rows) instead of the pass-balanced depth kernel
             dict(S=6, V=96, U=80, C=3, D=24, mode="ftc", scale=1.0, balance="0"),
             dict(S=6, V=96, U=80, C=3, D=24, mode="ftc", scale=1.0, balance="0", halo="nccl"),
             # unequal row blocks (cut by a made-up weight profile): shares of a pass span several owners
             dict(S=7, V=128, U=80, C=3, D=40, mode="ftc", scale=1.0, skew=True),
             # the geometry class of the bench: tensor-memory depth kernel (S >= 20, RGB), several chunks of hypotheses
             dict(S=28, V=64, U=160, C=3, D=72, mode="ftc", scale=1.0),
             dict(S=24, V=48, U=128, C=3, D=40, mode="depth2d", scale=1.0),
             # morphological opening of the edge mask (5x5 ellipse, core.hpp:759-769): every rank opens its rows plus a
             # margin taken from its own copy of the stack; coloured maps: quantile fit over the whole plane
             dict(S=6, V=96, U=80, C=3, D=24, mode="ftc", scale=1.0, opening=5, coloured=True),
             dict(S=5, V=64, U=72, C=1, D=16, mode="depth2d", scale=1.0, opening=3),
             # 16-bit stack, scaled by its maximum: all-reduce of the per-rank maxima, 16-bit raw rows gathered for the blur
             dict(S=5, V=96, U=80, C=3, D=16, mode="ftc", scale=-1.0, u16=True)]
    for i, c in enumerate(cases):
        os.environ.pop("RSLF_MEDIAN_GATHER", None)
        os.environ.pop("RSLF_HALO", None)
        os.environ.pop("RSLF_BALANCE", None)
        if c.get("balance"):
            os.environ["RSLF_BALANCE"] = c["balance"]           # "0": lock-step row blocks
        if c.get("halo"):
            os.environ["RSLF_HALO"] = c["halo"]                 # "nccl": small all-gather instead of peer-to-peer stores
        if c.get("gather"):
            os.environ["RSLF_MEDIAN_GATHER"] = c["gather"]      # whole-plane all-gather instead of the halo exchange
        epis, _ = make_light_field_np(c["S"], c["V"], c["U"], c["C"], dmin=-1.0, dmax=2.0, seed=300 + i, layers=5)
        if c["scale"] < 0:
            epis = (epis * 200.0 + 5.0).astype(np.float32)
        if c.get("u8"):
            epis = np.clip(np.rint(epis * 255.0), 0, 255).astype(np.uint8)
        if c.get("u16"):
            epis = np.clip(np.rint(epis * 200.0), 0, 65535).astype(np.uint16)
        p = api.default_params()
        if c.get("opening"):
            p.edge_confidence_opening_size = c["opening"]
        lut = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "colormap_jet.npy"))
        # single-GPU result (every rank computes it on its own device)
        ctx1 = api.Context(local)
        ctx1.upload_epis(epis, c["scale"])
        if c["mode"] == "ftc":
            ctx1.fine_to_coarse_run(-1.0, 2.0, c["D"], p)
            ref_map, ref_valid = ctx1.fine_to_coarse_get()
            ref_bgr = ctx1.fine_to_coarse_coloured(lut, p)[0] if c.get("coloured") else None
        else:
            ctx1.depth2d_run(-1.0, 2.0, c["D"], p)
            r = ctx1.depth2d_get()
            ref_map, ref_valid = r["best_depth"], r["edge_mask"]
        ref_samples = ctx1.timing()["samples"]
        ctx1.close()
        # sharded
        weights = None
        if c.get("skew"):
            weights = [1.0 + 9.0 * (v < c["V"] // 4) for v in range(c["V"])]      # the first quarter of the rows weighs 10x
        starts = shard_table(c["V"], c["U"], world, pyramid=(c["mode"] == "ftc"), weights=weights)
        v0, v1 = starts[rank], starts[rank + 1]
        ctx = api.Context(local)
        uid = [api.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)
        ctx.upload_epis(np.ascontiguousarray(epis[v0:v1]), c["scale"])
        ctx.set_row_shards(starts)
        if c["mode"] == "ftc":
            ctx.fine_to_coarse_run(-1.0, 2.0, c["D"], p)
            m, k = ctx.fine_to_coarse_get()
            bgr_ok = True
            if c.get("coloured"):
                bgr_ok = np.array_equal(ctx.fine_to_coarse_coloured(lut, p)[0], ref_bgr[:, v0:v1])
        else:
            bgr_ok = True
            ctx.depth2d_run(-1.0, 2.0, c["D"], p)
            r = ctx.depth2d_get()
            m, k = r["best_depth"], r["edge_mask"]
        samples = torch.tensor([ctx.timing()["samples"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(samples)
        ok = np.array_equal(m, ref_map[:, v0:v1]) and np.array_equal(k, ref_valid[:, v0:v1]) and float(samples[0]) == ref_samples and bgr_ok
        print("rank %d case %d rows [%d,%d): %s (samples %.0f vs %.0f)" % (rank, i, v0, v1, "OK" if ok else "MISMATCH",
                                                                          float(samples[0]), ref_samples), flush=True)
        failures += 0 if ok else 1
        ctx.close()
    t = torch.tensor([failures], device="cuda")
    dist.all_reduce(t)
    dist.destroy_process_group()
    if int(t[0]):
        sys.exit(1)
    if rank == 0:
        print("multi-GPU parity OK on %d ranks" % world)


if __name__ == "__main__":
    main()
