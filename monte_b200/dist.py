"""One-process-per-GPU sharding of the two hot paths (torch.distributed is plumbing only).

MC  : histories are independent.  Every rank runs photons n in its slice of [0, per) for the same
      views (history ids are global, so the union is identical to a single-GPU run) and the integer
      tallies are summed with ONE reduce to rank 0 per batch of views.  No other data-path exchange.
FDK : the volume is cut into z-slabs, the filter into view ranges.  One exchange step: every rank's
      filtered views are gathered by all ranks (all_gather when views divide evenly, else one
      broadcast per rank); after that each rank backprojects its own slab and nothing is exchanged.

The compute callables are injected, so the same code is exercised on CPU with the gloo backend
(tests/test_dist_cpu.py) and on B200s with NCCL (bench.py, tests/test_dist_gpu.py).
"""
import torch
import torch.distributed as dist


def split_range(n, world, rank):
    """Contiguous balanced split of range(n): sizes differ by at most one, earlier ranks get the extra."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def balanced_split(cost, world, align=1):
    """Contiguous split of range(len(cost)) into `world` pieces of nearly equal total cost: piece r ends
    where the running sum first reaches (r+1)/world of the total.  Every piece gets at least one item
    when len(cost) >= world.  With align > 1 a cut is moved to the nearest multiple of `align` when that
    keeps the pieces non-empty (the backprojector works in z-blocks of 16 slices at absolute multiples
    of 16, so aligned slabs waste no partial block).  Returns the list of (lo, hi)."""
    n = len(cost)
    acc, total = [0.0], 0.0
    for c in cost:
        total += float(c)
        acc.append(total)
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        lo = cuts[-1] + 1 if n >= world else cuts[-1]          # leave at least one item to the previous piece
        hi = n - (world - r) if n >= world else n              # ... and to each of the following ones
        k = lo
        while k < hi and acc[k] < target:
            k += 1
        if k > lo and target - acc[k - 1] < acc[k] - target:   # the nearer of the two candidate cuts
            k -= 1
        k = min(max(k, lo), hi)
        if align > 1:
            ka = (k + align // 2) // align * align
            if lo <= ka <= hi:
                k = ka
        cuts.append(k)
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def fdk_slice_cost(g, overhead=0.1, stride=8):
    """Relative backprojection cost of every z-slice of the volume: the fraction of its (voxel, view)
    pairs that project onto the detector (the others are skipped, recon/bp3d20.cpp:116) plus a constant
    for the per-column work that is done regardless.  Sampled on a coarse (s, t, view) grid with the
    reference's projection formulas (bp3d20.cpp:99-113); used to cut z-slabs of equal work, because at
    wide cone angles the end slices see the detector in few views or none."""
    import numpy as np
    X = (g.x0 + g.vox * np.arange(0, g.nx, stride))[None, :, None]
    Y = (g.y0 - g.vox * np.arange(0, g.ny, stride))[:, None, None]
    beta = np.deg2rad(g.angle0_deg + g.angle_step_deg * np.arange(0, g.n_views, stride))[None, None, :]
    k = g.dsd / (X * np.cos(beta) + Y * np.sin(beta) + g.dso)
    u_ok = np.abs(k * (-X * np.sin(beta) + Y * np.cos(beta))) <= g.half_u
    ks = np.sort(k[u_ok])                                      # |k Z| <= half_v  <=>  k <= half_v / |Z|
    Z = np.abs(g.z0 - g.vox * np.arange(g.nz))
    frac = np.searchsorted(ks, g.half_v / np.maximum(Z, 1e-12), side="right") / float(k.size)
    return frac + overhead


def is_dist():
    return dist.is_available() and dist.is_initialized()


def world():
    return (dist.get_rank(), dist.get_world_size()) if is_dist() else (0, 1)


# ----------------------------------------------------------------------------------- MC
def mc_sharded_step(run_local, image0, image5, per, views, reduce=True):
    """run_local(image0, image5, per, views, n_range) adds this rank's photons into the (zeroed)
    image tensors [n_views][ny][nx] int32; then the view slices are summed onto rank 0."""
    rank, ws = world()
    n_range = split_range(per, ws, rank)
    run_local(image0, image5, per, views, n_range)
    if ws > 1 and reduce:
        vb, ve = views
        dist.reduce(image0[vb:ve], dst=0, op=dist.ReduceOp.SUM)
        dist.reduce(image5[vb:ve], dst=0, op=dist.ReduceOp.SUM)
    return n_range


# ----------------------------------------------------------------------------------- FDK
def fdk_gather_filtered(filt_rows, n_views, nv, ws):
    """filt_rows: [n_views*nv + 2][pitch] tensor in which this rank has filled the rows of its own
    view range (split_range(n_views, ws, rank)); afterwards every rank holds all views."""
    if ws == 1:
        return
    pitch = filt_rows.shape[1]
    body = filt_rows[: n_views * nv].view(-1)
    if n_views % ws == 0:
        chunk = (n_views // ws) * nv * pitch
        rank = dist.get_rank()
        dist.all_gather_into_tensor(body, body[rank * chunk:(rank + 1) * chunk])
    else:
        for r in range(ws):
            lo, hi = split_range(n_views, ws, r)
            if hi > lo:
                dist.broadcast(body[lo * nv * pitch: hi * nv * pitch], src=r)


def fdk_sharded(filter_views, pad, backproject_slab, filt_rows, n_views, nv, nz):
    """filter_views(view_lo, view_hi) fills this rank's rows of filt_rows; pad() fixes the duplicated
    column / trailing rows; backproject_slab(z_lo, z_hi) reconstructs this rank's slab.
    Returns (view range, z range) of this rank."""
    rank, ws = world()
    v_lo, v_hi = split_range(n_views, ws, rank)
    z_lo, z_hi = split_range(nz, ws, rank)
    filter_views(v_lo, v_hi)
    fdk_gather_filtered(filt_rows, n_views, nv, ws)
    pad()
    backproject_slab(z_lo, z_hi)
    return (v_lo, v_hi), (z_lo, z_hi)


def fdk_sharded_pipelined(filter_views, pad_views, backproject_views, filt_rows, n_views, nv, nz, z_ranges=None):
    """Same result as fdk_sharded, bit for bit, with the exchange hidden behind the backprojection:
    every rank filters its own views, then the view pieces are broadcast in ascending order (all
    broadcasts are queued at once on NCCL's stream) and piece r is backprojected — continuing the
    fp32 partial sums — as soon as pieces r and r+1 have landed (piece r's pad fix-up needs the first
    rows of piece r+1).  Views are consumed in ascending order, exactly as on one GPU.
    filter_views(lo, hi); pad_views(lo, hi); backproject_views(z_lo, z_hi, v_lo, v_hi, continue_sum).
    z_ranges: one (z_lo, z_hi) per rank (e.g. balanced_split(fdk_slice_cost(g), world)); default: equal
    thickness.  Any contiguous partition gives the same voxels bit for bit."""
    rank, ws = world()
    z_lo, z_hi = z_ranges[rank] if z_ranges is not None else split_range(nz, ws, rank)
    pieces = [split_range(n_views, ws, r) for r in range(ws)]
    v_lo, v_hi = pieces[rank]
    filter_views(v_lo, v_hi)
    works = [None] * ws
    if ws > 1:
        pitch = filt_rows.shape[1]
        body = filt_rows[: n_views * nv].view(-1)
        for r, (lo, hi) in enumerate(pieces):
            if hi > lo:
                works[r] = dist.broadcast(body[lo * nv * pitch: hi * nv * pitch], src=r, async_op=True)
    started = False
    for r, (lo, hi) in enumerate(pieces):
        if hi <= lo:
            continue
        for q in (r, r + 1):                       # this piece and the one its last view reaches into
            if q < ws and works[q] is not None:
                works[q].wait()
                works[q] = None
        pad_views(lo, hi)
        backproject_views(z_lo, z_hi, lo, hi, started)
        started = True
    return (v_lo, v_hi), (z_lo, z_hi)


def max_over_ranks(value, device):
    """max of a python float over ranks (timing rule: a multi-GPU time is the slowest rank's)"""
    if not is_dist() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
