"""Host-side planning of the tensor-memory depth kernel (k_depth_tm.cuh: depth_tm_build, exposed as
rslf_plan_depth_tm): no device needed.  Invariants the kernel relies on:
  - every 4-view block of the padded stack is staged in exactly one round;
  - the views are [registers][tensor memory][shared memory] in ascending s, or reversed (the planner takes the
    order that needs fewer staging rounds, then less shared memory);
  - a round's staged blocks do not overlap and fit the area; rounds that are converted in place (registers, shared
    memory) have rows at least as large as the radiance rows, so block i's packed destination never overlaps the
    staging area of a later block;
  - twelve warps fit the SM's shared memory."""
import ctypes as C

import pytest

from remotesensingproject_b200 import api

SMEM = 227 * 1024
REG, TMEM, SMEMK = 0, 1, 2


MUST_FIT = {(100, 3, 256), (64, 3, 128), (200, 1, 512), (30, 1, 128), (24, 3, 40)}


def plan(S, Cc, D, s_hat, dmin=-1.0, dmax=4.0, slope=1.0, limit=SMEM):
    out = (C.c_int * 8)()
    rounds = (C.c_int * 24)()
    blocks = (C.c_int * 128)()
    rc = api.lib().rslf_plan_depth_tm(S, Cc, D, s_hat, C.c_float(dmin), C.c_float(dmax), C.c_float(slope), C.c_size_t(limit),
                                      out, rounds, blocks)
    assert rc == 0
    return list(out), list(rounds), list(blocks)


@pytest.mark.parametrize("S,Cc,D", [(100, 3, 256), (50, 3, 64), (64, 3, 128), (24, 3, 40), (200, 1, 512), (30, 1, 128), (97, 3, 33)])
def test_layout_invariants(S, Cc, D):
    Spad = (S + 3) // 4 * 4
    nblocks = Spad // 4
    for s_hat in sorted({0, 1, S // 4, S // 2, 3 * S // 4, S - 2, S - 1}):
        out, rounds, blocks = plan(S, Cc, D, s_hat)
        fits, RV, TV, SV, reg_last, nrounds, warp_bytes, warps = out
        if (S, Cc, D) in MUST_FIT:
            assert fits == 1, (S, Cc, D, s_hat)
        if not fits:            # wide chunks (few hypotheses over a wide range) at the rim: depth_kernel takes the pass
            continue
        assert RV + TV + SV == Spad and RV == 16 and TV % 4 == 0 and SV % 4 == 0
        assert TV * Cc <= 128                                   # columns of one warp's TMEM window
        assert warp_bytes * warps + 64 <= SMEM
        assert reg_last in (0, 1)
        # storage classes in stack order
        order = [(REG, RV), (TMEM, TV), (SMEMK, SV)] if not reg_last else [(SMEMK, SV), (TMEM, TV), (REG, RV)]
        kind_of_block, b = {}, 0
        for kind, n in order:
            for _ in range(n // 4):
                kind_of_block[b] = kind
                b += 1
        seen = set()
        area = (warp_bytes - 16) // 4                           # upper bound in floats (records included)
        for r in range(nrounds):
            b0, nb, kind = rounds[3 * r: 3 * r + 3]
            assert nb > 0
            used_end = 0
            for i in range(nb):
                blk = b0 + i
                assert blk not in seen and kind_of_block[blk] == kind
                seen.add(blk)
                off, pitch = blocks[2 * blk], blocks[2 * blk + 1]
                assert pitch % 4 == 0 and pitch >= 4
                assert off >= used_end                         # blocks of a round do not overlap
                used_end = off + 4 * pitch
                if kind != TMEM:
                    assert pitch >= Cc * 32                      # in-place conversion: row >= radiance row
                    assert off >= i * 4 * Cc * 32               # packed destination of block i ends before later blocks start
            assert used_end <= area
        assert seen == set(range(nblocks))
        assert rounds[3 * (nrounds - 1) + 2] == SMEMK or SV == 0   # the shared-memory round is the last one (its radiances stay)


def test_does_not_fit_is_reported():
    out, _, _ = plan(100, 3, 256, 50, limit=64 * 1024)
    assert out[0] == 0
    out, _, _ = plan(12, 3, 64, 6)                              # fewer views than the register tier + one block
    assert out[0] == 0
    out, _, _ = plan(400, 3, 64, 10)                            # more 4-view blocks than the layout table holds
    assert out[0] == 0
