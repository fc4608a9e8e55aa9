"""One-process-per-GPU sharding of the two hot paths (torch.distributed is plumbing only).

MC  : histories are independent.  Every rank runs photons n in its slice of [0, per) for the same
      views (history ids are global, so the union is identical to a single-GPU run) and the integer
      tallies are summed with ONE reduce to rank 0 per batch of views.  No other data-path exchange.
FDK : the volume is cut into z-slabs of equal work (fdk_slice_cost, balanced_split), the filter into view
      ranges.  One exchange step, three forms with identical results: fdk_sharded (every rank gathers all
      filtered views), fdk_sharded_pipelined (view pieces broadcast in order, overlapped with the
      backprojection), fdk_sharded_band (one all_to_all of just the detector rows each slab reads) and
      fdk_sharded_peers (no collective at all: the backprojector's pair conversion loads the band out of the
      peers' buffers, opened through CUDA IPC -- the one bench.py uses).  After that each rank backprojects its
      own slab; nothing else is exchanged.

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


def balanced_blocks(cost, world, block=16):
    """Partition range(len(cost)) into per-rank lists of ranges whose cuts all lie on multiples of `block`
    (the backprojector's z-block: an unaligned cut makes both neighbours compute the straddled block).
    With whole blocks as the unit a contiguous split cannot balance a volume whose end blocks are cheap
    (off the detector) and whose middle blocks are expensive: the expensive zone is therefore split
    contiguously and evenly, and the cheap blocks outside it are then handed, cheapest-loaded rank
    first, to the ranks as a second range.  Returns [[(lo, hi), ...] per rank]; every slice is covered
    exactly once."""
    n = len(cost)
    nb = (n + block - 1) // block
    bc = [float(sum(cost[b * block:min((b + 1) * block, n)])) for b in range(nb)]
    if nb <= world or max(bc) <= 0:
        return [[r] if r[1] > r[0] else [] for r in balanced_split(cost, world, block)]
    heavy = [b for b in range(nb) if bc[b] > 0.25 * max(bc)]
    h0, h1 = heavy[0], heavy[-1] + 1                            # the expensive zone (contiguous by construction of the cost)
    if h1 - h0 < world:
        return [[r] if r[1] > r[0] else [] for r in balanced_split(cost, world, block)]
    mid = balanced_split(bc[h0:h1], world)
    parts = [[(h0 + a, h0 + b)] for a, b in mid]
    load = [sum(bc[h0 + a:h0 + b]) for a, b in mid]
    for b in sorted(list(range(0, h0)) + list(range(h1, nb)), key=lambda q: -bc[q]):
        r = min(range(world), key=lambda q: load[q])
        parts[r].append((b, b + 1))
        load[r] += bc[b]
    out = []
    for pr in parts:
        pr.sort()
        merged = []
        for a, b in pr:
            if merged and merged[-1][1] == a:
                merged[-1] = (merged[-1][0], b)
            else:
                merged.append((a, b))
        out.append([(a * block, min(b * block, n)) for a, b in merged])
    return out


def fdk_z_partition(g, world_size, block=16, min_gain=0.03):
    """z-ranges per rank for the sharded backprojection: the contiguous equal-work split, unless handing
    the cheap end blocks out separately (balanced_blocks) lowers the slowest rank's load by > min_gain."""
    cost = fdk_slice_cost(g)
    contiguous = [[r] for r in balanced_split(cost, world_size, block)]
    blocks = balanced_blocks(cost, world_size, block)

    def worst(parts):
        return max(sum(float(cost[a:b].sum()) for a, b in pr) for pr in parts)
    return blocks if worst(blocks) < (1.0 - min_gain) * worst(contiguous) else contiguous


def fdk_slice_cost(g, overhead=0.32, partial_penalty=0.45, stride=8):
    """Relative backprojection cost of every z-slice of the volume, used to cut z-slabs of equal work: at
    wide cone angles the end slices see the detector in few views or none.  Model, calibrated on one
    B200 at C3 by timing slabs (scripts/fdk_slab_cost.py): cost = overhead + f * (1 + partial_penalty * [f
    below its maximum]) where f is the fraction of the slice's (voxel, view) pairs that project onto the
    detector (the others are skipped, recon/bp3d20.cpp:116), `overhead` the per-column work done
    regardless of f while any view sees the slice (0.042 ms per slice against 0.131 ms per fully visible
    slice; z-blocks that no view sees return at once), and the penalty the
    slower general path that columns near the detector edge take.  f is sampled on a coarse (s, t, view)
    grid with the reference's projection formulas (bp3d20.cpp:99-113)."""
    import numpy as np
    X = (g.x0 + g.vox * np.arange(0, g.nx, stride))[None, :, None]
    Y = (g.y0 - g.vox * np.arange(0, g.ny, stride))[:, None, None]
    beta = np.deg2rad(g.angle0_deg + g.angle_step_deg * np.arange(0, g.n_views, stride))[None, None, :]
    k = g.dsd / (X * np.cos(beta) + Y * np.sin(beta) + g.dso)
    u_ok = np.abs(k * (-X * np.sin(beta) + Y * np.cos(beta))) <= g.half_u
    ks = np.sort(k[u_ok])                                      # |k Z| <= half_v  <=>  k <= half_v / |Z|
    Z = np.abs(g.z0 - g.vox * np.arange(g.nz))
    frac = np.searchsorted(ks, g.half_v / np.maximum(Z, 1e-12), side="right") / float(k.size)
    partial = (frac > 0) & (frac < 0.97 * frac.max())
    # slices no view can see cost (almost) nothing: their z-blocks return at once
    return np.where(frac > 0, overhead, min(overhead, 0.02)) + frac * (1.0 + partial_penalty * partial)


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
    z_ranges: per rank one (z_lo, z_hi) or a list of them (balanced_split / balanced_blocks of
    fdk_slice_cost(g)); default: equal thickness.  Any partition gives the same voxels bit for bit."""
    rank, ws = world()
    mine = z_ranges[rank] if z_ranges is not None else split_range(nz, ws, rank)
    if len(mine) == 2 and not isinstance(mine[0], (tuple, list)):
        mine = [tuple(mine)]                                    # one (z_lo, z_hi)
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
        for (z_lo, z_hi) in mine:                  # several ranges per rank: see balanced_blocks
            if z_hi > z_lo:
                backproject_views(z_lo, z_hi, lo, hi, started)
        started = True
    return (v_lo, v_hi), (mine[0] if len(mine) == 1 else list(mine))


def _merge_rows(ranges, nv):
    """sorted, merged, clipped list of (row_lo, row_hi)"""
    out = []
    for a, b in sorted((max(0, a), min(nv, b)) for a, b in ranges):
        if b <= a:
            continue
        if out and a <= out[-1][1]:
            out[-1] = (out[-1][0], max(out[-1][1], b))
        else:
            out.append((a, b))
    return out


def fdk_sharded_band(filter_views, pad, backproject_slab, slab_rows, filt_rows, n_views, nv, z_ranges):
    """Same voxels as fdk_sharded / fdk_sharded_pipelined, bit for bit, with a much smaller exchange: a
    z-slab reads only a band of axial detector rows of every view (slab_rows(z_lo, z_hi) -> (row_lo,
    row_hi), monte_gpu_fdk_slab_rows), plus rows 0..3 that the previous view's last row reaches into.
    Every rank filters its own views, packs for each peer the rows that peer's slabs need, and ONE
    all_to_all_single moves them (C3 on 8 ranks: ~0.25 GB per rank instead of 2 GB); the rows no slab of
    this rank reads stay unwritten.  Then pad() and backproject_slab(z_lo, z_hi) for each of the rank's
    ranges run over all views in one go.  z_ranges: per rank a (z_lo, z_hi) or a list of them."""
    rank, ws = world()
    norm = [[tuple(zr)] if len(zr) == 2 and not isinstance(zr[0], (tuple, list)) else [tuple(q) for q in zr] for zr in z_ranges]
    pieces = [split_range(n_views, ws, r) for r in range(ws)]
    v_lo, v_hi = pieces[rank]
    filter_views(v_lo, v_hi)
    if ws > 1:
        pitch = filt_rows.shape[1]
        f3 = filt_rows[: n_views * nv].view(n_views, nv, pitch)
        need = []                                   # rows of every view that rank r reads
        for r in range(ws):
            rr = []
            for z_lo, z_hi in norm[r]:
                a, b = slab_rows(z_lo, z_hi)
                if b > a:
                    rr += [(a, b + 1), (0, 4)]     # + the partner row of the last pair, + the next view's first rows
            need.append(_merge_rows(rr, nv))
        n_rows = [sum(b - a for a, b in need[r]) for r in range(ws)]
        mine = v_hi - v_lo
        in_splits = [mine * n_rows[r] * pitch if r != rank else 0 for r in range(ws)]
        out_splits = [(pieces[q][1] - pieces[q][0]) * n_rows[rank] * pitch if q != rank else 0 for q in range(ws)]
        parts = [f3[v_lo:v_hi, a:b, :].reshape(-1) for r in range(ws) if r != rank for a, b in need[r]]
        send = torch.cat(parts) if parts else filt_rows.new_empty(0)
        recv = filt_rows.new_empty(sum(out_splits))
        dist.all_to_all_single(recv, send, out_splits, in_splits)
        o = 0
        for q in range(ws):
            if q == rank:
                continue
            nq = pieces[q][1] - pieces[q][0]
            for a, b in need[rank]:
                n = nq * (b - a) * pitch
                f3[pieces[q][0]:pieces[q][1], a:b, :] = recv[o:o + n].view(nq, b - a, pitch)
                o += n
    pad()
    for z_lo, z_hi in norm[rank]:
        if z_hi > z_lo:
            backproject_slab(z_lo, z_hi)
    return (v_lo, v_hi), (norm[rank][0] if len(norm[rank]) == 1 else norm[rank])


class PeerRows:
    """The padded-row buffers of all ranks, visible to every rank: each rank exports its own buffer (CUDA IPC handle of
    the allocation + offset, api.ipc_export) and opens the others' once (api.ipc_open).  ptrs[r] is the address, in THIS
    process, of rank r's buffer; v_end[r] the end of the view range rank r filters.  Set-up is a collective."""

    def __init__(self, api, filt_rows, n_views):
        rank, ws = world()
        self.api, self.rank, self.ws = api, rank, ws
        self.v_end = [split_range(n_views, ws, r)[1] for r in range(ws)]
        mine = api.ipc_export(filt_rows)
        handles = [None] * ws
        if ws > 1:
            dist.all_gather_object(handles, mine)
        else:
            handles[0] = mine
        self.opened = []
        self.ptrs = []
        for r in range(ws):
            if r == rank:
                self.ptrs.append(filt_rows.data_ptr())
            else:
                p = api.ipc_open(*handles[r])
                self.opened.append(p)
                self.ptrs.append(p)
        self.token = torch.zeros(1, device=filt_rows.device) if filt_rows.is_cuda else None

    def fence(self):
        """every rank's work issued so far on the current stream is done before any rank's later work starts: a
        one-element all-reduce ON THE STREAM (no host synchronisation)"""
        if self.ws > 1:
            if self.token is not None:
                dist.all_reduce(self.token)
            else:
                dist.barrier()

    def close(self):
        for p in self.opened:
            self.api.ipc_close(p)
        self.opened = []


def fdk_sharded_peers(filter_views, backproject_peers, peers, n_views, z_ranges):
    """Same voxels as the other fdk_sharded_* forms, bit for bit, with NO collective on the data path: every rank filters
    its own views into its own buffer; after a fence each rank backprojects its z ranges with
    backproject_peers(z_lo, z_hi, ptrs, v_end) (monte_gpu_fdk_backproject_peers_dev), which loads the detector-row band
    the slab reads straight out of the peers' buffers over NVLink while converting it to the backprojector's pair
    layout -- no packing, no all_to_all, no unpacking, and only the rows a slab reads travel.  A second fence keeps
    everybody's rows in place until all peers have read them."""
    rank, ws = world()
    norm = [[tuple(zr)] if len(zr) == 2 and not isinstance(zr[0], (tuple, list)) else [tuple(q) for q in zr] for zr in z_ranges]
    v_lo, v_hi = split_range(n_views, ws, rank)
    filter_views(v_lo, v_hi)
    peers.fence()
    for z_lo, z_hi in norm[rank]:
        if z_hi > z_lo:
            backproject_peers(z_lo, z_hi, peers.ptrs, peers.v_end)
    peers.fence()
    return (v_lo, v_hi), (norm[rank][0] if len(norm[rank]) == 1 else norm[rank])


def max_over_ranks(value, device):
    """max of a python float over ranks (timing rule: a multi-GPU time is the slowest rank's)"""
    if not is_dist() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
