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


def balanced_blocks(weights, world, align):
    """Contiguous partition of len(weights) rows into `world` blocks with boundaries at multiples of `align`
    that minimises the heaviest block (binary search on the cap + greedy sweep; every block gets >= 1 unit)."""
    V = len(weights)
    units = [(i, min(V, i + align)) for i in range(0, V, align)]
    w = [float(sum(weights[a:b])) for a, b in units]
    n = len(w)
    if n < world:
        raise ValueError("%d rows cannot be split over %d ranks at alignment %d" % (V, world, align))

    def cuts(cap):
        out, acc, used = [], 0.0, 1
        for i, x in enumerate(w):
            left_units, left_blocks = n - i, world - used
            if acc > 0 and (acc + x > cap or left_units == left_blocks):   # close the block; keep a unit for every later block
                if used == world:
                    return None
                out.append(i)
                acc, used = 0.0, used + 1
            acc += x
            if acc > cap and x <= cap:
                return None
        return out

    lo, hi = max(w), sum(w)
    best = None
    for _ in range(50):
        mid = 0.5 * (lo + hi)
        c = cuts(mid)
        if c is not None and len(c) <= world - 1:
            best, hi = c, mid
        else:
            lo = mid
    if best is None:
        best = cuts(sum(w)) or []
    # pad with extra cuts (splitting the last blocks) if the greedy used fewer than world blocks
    cut = sorted(set(best))
    i = n - 1
    while len(cut) < world - 1:
        if i not in cut and i > 0:
            cut.append(i)
        i -= 1
        cut = sorted(set(cut))
    return [0] + [units[c][0] for c in cut] + [V]


def shard_table(V, U, world, max_pyr_depth=-1, pyramid=True, min_rows=16, weights=None):
    """Row boundaries [b0 .. b_world] for `world` ranks.

    With the fine-to-coarse pyramid the library shards level p by rows as long as every boundary is a multiple
    of 2^p and replicates the remaining (coarse, tiny) levels on every rank.  The boundaries are therefore
    aligned to 2^k with k as large as possible while a rank keeps about `min_rows` rows at level k: fine enough
    for balanced blocks, coarse enough that only a few thousand pixels are computed redundantly.

    weights (optional, one per row): expected work per row, e.g. the number of edge-confident pixels; the blocks
    are then balanced by weight instead of by row count (the passes of all ranks advance in lock-step through the
    per-pass halo exchange, so the heaviest block sets the pace)."""
    levels = len(pyramid_levels(V, U, max_pyr_depth)) if pyramid else 1
    k = 0
    while k + 1 < levels and (V // world) >> (k + 1) >= min_rows:
        k += 1
    align = 1 << k
    if weights is not None and world > 1:
        starts = balanced_blocks(list(weights), world, align)
    else:
        sh = row_shards(V, world, align)
        starts = [a for a, _ in sh] + [V]
    if len(starts) != world + 1 or any(b <= a for a, b in zip(starts[:-1], starts[1:])):
        raise ValueError("%d rows cannot be split over %d ranks" % (V, world))
    return starts


def row_work_estimate(epis, scale_factor=-1.0, edge_threshold=0.02, shadow_level=0.0866, views=5):
    """Work estimate per image row for balanced sharding: number of edge-confident pixels of a few views
    (a torch restatement of the edge-confidence criterion, rslf_depth_computation_core.hpp:426-478; only a
    heuristic for the partition, results never depend on it).  epis: torch tensor [V][S][U][C], any device."""
    import torch
    V, S, U, C = epis.shape
    x = epis.float()
    if epis.dtype == torch.uint8:
        x = x / 255.0
    else:
        x = x / (x.max() if scale_factor < 0 else scale_factor)
    picks = sorted({min(S - 1, max(0, int(round(i * (S - 1) / max(1, views - 1))))) for i in range(views)})
    idx = torch.arange(U, device=x.device)
    w = torch.zeros(V, dtype=torch.float64, device=x.device)
    for s in picks:
        img = x[:, s]                                         # [V][U][C]
        ce = torch.zeros((V, U), dtype=torch.float32, device=x.device)
        for k in range(-4, 5):
            if k == 0:
                continue
            j = idx + k
            j = torch.where(j < 0, -j, j)
            j = torch.where(j > U - 1, 2 * (U - 1) - j, j).clamp_(0, U - 1)          # BORDER_REFLECT_101
            ce += ((img - img[:, j]) ** 2).sum(-1)
        norm = img.abs().squeeze(-1) * 1.7320508 if C == 1 else img.norm(dim=-1)
        w += ((ce > edge_threshold) & (norm >= shadow_level)).sum(dim=1).double()
    return (w + 1.0).cpu().tolist()
