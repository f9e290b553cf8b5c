"""Host-side logic of the row-sharded (one process per GPU) path, on CPU with gloo, world_size 2.

What the GPU path exchanges per pass is emulated with the oracle: every rank computes the
per-row (row-local) part of a Depth1D pile pass on its block of rows, the ranks all-gather
the depth / mask rows of line s_hat, and the cross-row selective median on the gathered planes
must reproduce the single-process result on the rank's rows.  Also checks the shard tables."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from remotesensingproject_b200.shard import balanced_blocks, pyramid_levels, row_shards, shard_table  # noqa: E402


def test_shard_tables():
    # C3 on 8 ranks: boundaries multiples of 8 (levels 0-3 are sharded, 4-6 replicated), blocks balanced within 8 rows
    starts = shard_table(1080, 1920, 8)
    assert starts[0] == 0 and starts[-1] == 1080 and all(b % 8 == 0 for b in starts[1:-1])
    rows = [b - a for a, b in zip(starts[:-1], starts[1:])]
    assert max(rows) - min(rows) <= 8
    levels = len(pyramid_levels(1080, 1920))
    assert levels == 7
    for p in range(4):                      # every sharded level keeps >= 2 rows per rank
        b = [s >> p for s in starts[:-1]] + [pyramid_levels(1080, 1920)[p][0]]
        assert all(y - x >= 2 for x, y in zip(b[:-1], b[1:]))
    assert shard_table(1080, 1920, 2) == [0, 544, 1080]
    assert pyramid_levels(600, 1200) == [(600, 1200), (300, 600), (150, 300), (75, 150), (38, 75), (19, 38)]
    assert row_shards(10, 3) == [(0, 3), (3, 6), (6, 10)]
    assert shard_table(540, 960, 4, pyramid=False) == [0, 135, 270, 405, 540]
    with pytest.raises(ValueError):
        shard_table(12, 60, 16)           # fewer rows than ranks


def test_work_balanced_blocks():
    rng = np.random.default_rng(0)
    for _ in range(100):
        V = int(rng.integers(40, 1200))
        world = int(rng.choice([2, 3, 4, 8]))
        align = int(rng.choice([1, 2, 4, 8, 16]))
        if (V + align - 1) // align < world:
            continue
        w = rng.random(V) ** 3 * 10 + 0.01
        st = balanced_blocks(list(w), world, align)
        assert len(st) == world + 1 and st[0] == 0 and st[-1] == V
        assert all(b > a for a, b in zip(st[:-1], st[1:])) and all(x % align == 0 for x in st[1:-1])
        heavy = max(w[a:b].sum() for a, b in zip(st[:-1], st[1:]))
        even = max(w[a:b].sum() for a, b in row_shards(V, world, align))
        assert heavy <= even * 1.0001          # never worse than equal row counts
    # a skewed profile: the balanced cut gives the heavy rows to a smaller block
    w = [10.0] * 100 + [1.0] * 900
    st = shard_table(1000, 1500, 4, weights=w)
    assert st[1] < 250 and st[-1] == 1000


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from remotesensingproject_b200.synth import make_light_field_np
    oracle.set_num_threads(1)
    S, V, U, C, D = 7, 22, 48, 3, 16
    epis, _ = make_light_field_np(S, V, U, C, dmin=-1.0, dmax=2.0, seed=26, layers=5)
    epis = oracle.normalise(epis, 1.0)
    starts = shard_table(V, U, world, pyramid=False)
    tabs = [None] * world
    dist.all_gather_object(tabs, starts)
    assert all(t == starts for t in tabs)                        # every rank derives the same table
    v0, v1 = starts[rank], starts[rank + 1]
    full = oracle.depth1d_pile(epis, -1.0, 2.0, D)
    # row-local part on the rank's rows only
    mine = oracle.depth1d_pile(epis[v0:v1], -1.0, 2.0, D)
    assert np.array_equal(mine["raw_depth"], full["raw_depth"][v0:v1])
    assert np.array_equal(mine["edge_mask"], full["edge_mask"][v0:v1])
    # per-pass exchange: depth + mask rows of line s_hat from every rank (padded to equal size)
    maxrows = max(b - a for a, b in zip(starts[:-1], starts[1:]))
    send = torch.zeros((maxrows, U, 2), dtype=torch.float32)
    send[: v1 - v0, :, 0] = torch.from_numpy(mine["raw_depth"])
    send[: v1 - v0, :, 1] = torch.from_numpy(mine["edge_mask"].astype(np.float32))
    recv = [torch.zeros_like(send) for _ in range(world)]
    dist.all_gather(recv, send)
    g_depth = np.concatenate([recv[r][: starts[r + 1] - starts[r], :, 0].numpy() for r in range(world)])
    g_mask = np.concatenate([recv[r][: starts[r + 1] - starts[r], :, 1].numpy() for r in range(world)]).astype(np.uint8)
    filt = oracle.selective_median(g_depth, g_mask, epis, S // 2)
    ok = np.array_equal(filt[v0:v1], full["best_depth"][v0:v1])
    # a rank that filtered only its own rows (no exchange) must differ somewhere near the border: the exchange matters
    alone = oracle.selective_median(mine["raw_depth"], mine["edge_mask"], epis[v0:v1], S // 2)
    out.put((rank, bool(ok), bool(np.array_equal(alone, full["best_depth"][v0:v1]))))
    dist.destroy_process_group()


def test_median_exchange_with_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert not all(alone for _, _, alone in res), "the halo exchange made no difference: test field too plain"


def test_pass_balanced_shares_partition_every_pass():
    """Pass-balanced multi-GPU mode (k_balance.cuh): the ranks' shares of a pass tile the concatenated work lists exactly,
    differ by at most one pixel, and every owner's list is fully covered (host restatement of the device arithmetic)."""
    import ctypes as C
    from remotesensingproject_b200 import api
    lib = api.lib()
    rng = np.random.default_rng(5)
    for world in (1, 2, 3, 4, 8, 16):
        for trial in range(20):
            counts = rng.integers(0, 5000, size=world)
            if trial % 5 == 0:
                counts[rng.integers(0, world)] = 0
            if trial == 7:
                counts[:] = 0
            arr = (C.c_int * world)(*[int(x) for x in counts])
            T = int(counts.sum())
            covered = np.zeros(world, dtype=np.int64)
            nxt, sizes = 0, []
            for r in range(world):
                first, cnt = C.c_longlong(), C.c_int()
                frm = (C.c_int * world)()
                assert lib.rslf_balance_share(arr, world, r, C.byref(first), C.byref(cnt), frm) == 0
                assert first.value == nxt                       # contiguous, in rank order
                nxt += cnt.value
                sizes.append(cnt.value)
                assert sum(frm) == cnt.value
                covered += np.array(list(frm), dtype=np.int64)
            assert nxt == T and max(sizes) - min(sizes) <= 1
            np.testing.assert_array_equal(covered, counts)
    assert lib.rslf_balance_share(arr, 17, 0, C.byref(first), C.byref(cnt), None) != 0
