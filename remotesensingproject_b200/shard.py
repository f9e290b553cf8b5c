"""Row sharding of a light field across ranks (the reference's OpenMP axis, core.hpp:799).

Contiguous blocks of image rows per rank; block boundaries are multiples of
2^(levels-1) rows so that every pyramid level splits at the same image position
and bound propagation (rows 2v, 2v+1 -> v, rslf_fine_to_coarse.hpp:201-294)
stays rank-local.
"""


def pyramid_levels(V, U, max_pyr_depth=-1):
    """Level sizes of FineToCoarse (ftc.hpp:130): while V > 10 and U > 10, halve with cvRound (half to even)."""
    dims = []
    while V > 10 and U > 10 and (max_pyr_depth < 1 or len(dims) < max_pyr_depth):
        dims.append((V, U))
        V, U = int(round(V * 0.5)), int(round(U * 0.5))     # Python's round is half-to-even, like cvRound
    return dims


def row_shards(V, world, align=1):
    """[(v0, v1)] per rank: contiguous, covering [0, V), boundaries multiples of `align`."""
    if world < 1:
        raise ValueError("world must be >= 1")
    blocks = (V + align - 1) // align
    out = []
    for r in range(world):
        b0 = blocks * r // world
        b1 = blocks * (r + 1) // world
        out.append((min(V, b0 * align), min(V, b1 * align)))
    return out


def shard_table(V, U, world, max_pyr_depth=-1, pyramid=True, min_rows=16):
    """Row boundaries [b0 .. b_world] for `world` ranks.

    With the fine-to-coarse pyramid the library shards level p by rows as long as every boundary is a multiple
    of 2^p and replicates the remaining (coarse, tiny) levels on every rank.  The boundaries are therefore
    aligned to 2^k with k as large as possible while a rank keeps about `min_rows` rows at level k: fine enough
    for balanced blocks, coarse enough that only a few thousand pixels are computed redundantly."""
    levels = len(pyramid_levels(V, U, max_pyr_depth)) if pyramid else 1
    k = 0
    while k + 1 < levels and (V // world) >> (k + 1) >= min_rows:
        k += 1
    align = 1 << k
    sh = row_shards(V, world, align)
    starts = [a for a, _ in sh] + [V]
    if any(b <= a for a, b in zip(starts[:-1], starts[1:])):
        raise ValueError("%d rows cannot be split over %d ranks" % (V, world))
    return starts
