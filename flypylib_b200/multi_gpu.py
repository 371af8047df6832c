"""Multi-GPU plumbing for the T-bar path: one process per GPU, ``torch.distributed`` (NCCL over
NVLink on the GPU box, gloo in the CPU tests).

Replaces the reference's in-graph tower replication (flypylib/multi_gpu.py:20-61 ``make_parallel``: one
tile per /gpu:i per predict step, outputs concatenated on /cpu:0) and the process fan-out of
``full_roi_inference`` (flypylib/fplobjdetect.py:841-986).  The path shards over independent units --
tile layers of the reference grid for the forward pass, substacks (+ buffer) for ``voxel2obj`` -- so the
only collective is a small all-gather of the per-rank detection lists.
"""
import numpy as np


def partition_layers(n_layers, world_size):
    """Contiguous split of n_layers tile layers over world_size ranks (first ranks take the remainder):
    25 layers on 8 ranks -> 4,3,3,3,3,3,3,3.  Returns [(begin, end)] * world_size."""
    base, rem = divmod(int(n_layers), int(world_size))
    out, b = [], 0
    for r in range(world_size):
        n = base + (1 if r < rem else 0)
        out.append((b, b + n))
        b += n
    return out


def tile_layers(size, rf_offset, out_sz):
    """Number of tile origins along one axis: len(np.mgrid[off : size-off : out_sz]) (fplnetwork.py:149-154)."""
    span = size - 2 * rf_offset
    return 0 if span <= 0 else -(-span // out_sz)


def slab_for_layers(begin, end, size, rf_offset, out_sz):
    """Input z-range [z0, z1) a rank needs to evaluate tile layers [begin, end), and the z-range of the
    prediction rows it produces (reference grid: layer k reads [k*out, k*out + out + 2*off))."""
    if end <= begin:
        return (0, 0), (0, 0)
    z0 = begin * out_sz
    z1 = min(size, end * out_sz + 2 * rf_offset)
    p0 = rf_offset + begin * out_sz
    p1 = min(size - rf_offset, rf_offset + end * out_sz)
    return (z0, z1), (p0, p1)


def shard_plan(Z, rf_offset, granularity, world_size):
    """z-slab sharding of ONE (Z,Y,X) volume over world_size ranks for ``FplNetwork.infer``: the
    prediction planes [off, Z-off) are cut into groups of ``granularity`` planes (``FplNetwork.slab_granularity``:
    rf_stride for the VGGs, the tile pitch for the U-Nets), the groups are dealt contiguously
    (``partition_layers``), and rank r reads the image planes its groups need -- its slab plus a 2*rf_offset
    receptive-field halo, nothing is recomputed beyond that.  Returns per rank
    ``((in0, in1), (own0, own1))``: image planes to read and prediction planes owned; the owned ranges partition
    [0, Z) (the first / last non-empty rank also owns the zero border planes), which is what
    ``voxel2obj_global`` expects.  Ranks without work get empty ranges."""
    Z, off, gran = int(Z), int(rf_offset), int(granularity)
    span = Z - 2 * off
    n_groups = 0 if span <= 0 else -(-span // gran)
    if n_groups == 0:                       # thinner than the receptive field: rank 0 returns the all-zero map
        return [((0, Z), (0, Z))] + [((Z, Z), (Z, Z))] * (world_size - 1)
    parts = partition_layers(n_groups, world_size)
    busy = [r for r, (a, b) in enumerate(parts) if b > a]
    plans = []
    for r, (g0, g1) in enumerate(parts):
        if g1 <= g0:
            edge = Z if r > busy[-1] else 0
            plans.append(((edge, edge), (edge, edge)))
            continue
        in0 = g0 * gran
        in1 = Z if r == busy[-1] else g1 * gran + 2 * off
        own0 = 0 if r == busy[0] else in0 + off
        own1 = Z if r == busy[-1] else in1 - off
        plans.append(((in0, in1), (own0, own1)))
    return plans


def allgather_detections(rows, group=None):
    """All-gather variable-length detection lists.  rows: (K,4) float64 tensor (x,y,z,conf) on the
    device the process group works on.  Returns the list of per-rank (K_r,4) tensors."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    m = max(int(c.item()) for c in counts)
    pad = torch.zeros((max(m, 1), 4), dtype=torch.float64, device=rows.device)
    pad[:rows.shape[0]] = rows
    bufs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return [b[:int(c.item())] for b, c in zip(bufs, counts)]


def merge_detections(parts):
    """Concatenate per-rank detections into one {'locs','conf'} dict in the reference's emission order
    (confidence descending; ties by (z,y,x) ascending, i.e. by flat index of a common volume)."""
    rows = np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1, 4) for p in parts], 0) \
        if parts else np.zeros((0, 4))
    if rows.shape[0]:
        order = np.lexsort((rows[:, 0], rows[:, 1], rows[:, 2], -rows[:, 3]))
        rows = rows[order]
    return {'locs': rows[:, :3].copy(), 'conf': rows[:, 3].copy()}


def slab_detections(pred_slab_fn, Z, rank, world, obj_min_dist, smoothing_sigma, buffer_sz, thd=0):
    """Detections rank `rank` of `world` owns under the substack semantics of full_roi_inference
    (flypylib/fplobjdetect.py:841-986, :1031-1034): slab [z0,z1) of the volume, extended by buffer_sz planes on
    each side, is an independent voxel2obj (own percentile, own greedy NMS); detections outside [z0,z1) are
    dropped (the reference drops the buffer zone, fplobjdetect.py:239-250).  pred_slab_fn(lo, hi) returns the
    CUDA float32 probability map of planes [lo,hi).  Returns (K,4) float64 rows (x,y,z,conf), z global."""
    from . import fplobjdetect
    z0, z1 = partition_layers(Z, world)[rank]
    if z1 <= z0:
        return np.zeros((0, 4))
    lo, hi = max(0, z0 - buffer_sz), min(Z, z1 + buffer_sz)
    out = fplobjdetect.voxel2obj_device(pred_slab_fn(lo, hi), obj_min_dist, smoothing_sigma, (0, 0, lo), 0, thd)
    rows = np.concatenate([out['locs'], out['conf'][:, None]], 1)
    return rows[(rows[:, 2] >= z0) & (rows[:, 2] < z1)]


def detect_substacks(network, image_dev, normalize, obj_min_dist, smoothing_sigma, buffer_sz, thd=0,
                     group=None):
    """z-slab sharded T-bar detection on the ranks of `group` (one process per GPU): every rank runs
    infer + voxel2obj on its slab (+ buffer) with no communication, then the per-slab detection lists are
    all-gathered (NCCL) -- the only collective of the path.  Returns the merged dict on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    Z = int(image_dev.shape[0])

    def pred_slab(lo, hi):
        return network.infer_device(image_dev[lo:hi].contiguous(), normalize=normalize)

    rows = slab_detections(pred_slab, Z, rank, world, obj_min_dist, smoothing_sigma, buffer_sz, thd)
    if world == 1:
        return merge_detections([rows])
    parts = allgather_detections(torch.from_numpy(np.ascontiguousarray(rows)).to(image_dev.device), group)
    return merge_detections([p.cpu().numpy() for p in parts])


# ---------------------------------------------------------------------------------------------------
# Sub-layer load balance for the U-Nets (their output depends on the tile phase, so work can only be cut on the
# reference tile grid): the (z, y) ROWS of tiles are dealt to the ranks instead of whole layers -- 25 layers on 8 ranks
# balance 25/32 = 0.78, 625 rows balance 0.99 -- and the rows a rank computes for a neighbour's layer travel to the
# plane owner over NCCL P2P (one partial layer per rank, ~1 GB at 2048^2 x 82 planes: ~2 ms on NVLink).
# ---------------------------------------------------------------------------------------------------
def row_plan(Z, Y, rf_offset, out_sz, world_size):
    """Deal the tile rows of the reference grid.  Returns (pieces, plans): ``pieces[r]`` = list of
    ``(kz, ya, yb, owner)`` -- rank r evaluates the tiles of layer kz, rows [ya, yb) (all x), and the prediction
    planes of layer kz belong to rank ``owner``; ``plans`` = ``shard_plan`` (plane ownership by whole layers, what
    ``voxel2obj_global`` works on)."""
    nz, ny = tile_layers(Z, rf_offset, out_sz), tile_layers(Y, rf_offset, out_sz)
    plans = shard_plan(Z, rf_offset, out_sz, world_size)
    own_layers = partition_layers(nz, world_size)

    def owner(kz):
        for r, (a, b) in enumerate(own_layers):
            if a <= kz < b:
                return r
        raise AssertionError(kz)

    pieces = []
    for a, b in partition_layers(nz * ny, world_size):
        mine = []
        if b > a:
            for kz in range(a // ny, (b - 1) // ny + 1):
                ya, yb = max(a, kz * ny) - kz * ny, min(b, (kz + 1) * ny) - kz * ny
                mine.append((kz, ya, yb, owner(kz)))
        pieces.append(mine)
    return pieces, plans


def piece_geometry(piece, Z, Y, rf_offset, out_sz):
    """((z0, z1), (y0, y1)) image block a piece reads and ((pz0, pz1), (py0, py1)) prediction block it produces
    (global coordinates; the far faces clip at the volume, like the reference's zero-padded edge tiles)."""
    kz, ya, yb, _ = piece
    off = rf_offset
    z0, y0 = kz * out_sz, ya * out_sz
    z1, y1 = min(Z, z0 + out_sz + 2 * off), min(Y, yb * out_sz + 2 * off)
    pz0, py0 = z0 + off, y0 + off
    pz1, py1 = min(Z - off, pz0 + out_sz), min(Y - off, y0 + off + (yb - ya) * out_sz)
    return ((z0, z1), (y0, y1)), ((pz0, pz1), (py0, py1))


def image_planes_for_pieces(my_pieces, Z, Y, rf_offset, out_sz):
    """z-range of the image the pieces of one rank read (empty: (Z, Z))."""
    if not my_pieces:
        return (Z, Z)
    g = [piece_geometry(p, Z, Y, rf_offset, out_sz)[0][0] for p in my_pieces]
    return (min(a for a, _ in g), max(b for _, b in g))


def infer_rows_sharded(compute_block, pieces, plans, rank, Z, Y, X, rf_offset, out_sz, halo, device, group=None,
                       dtype=None):
    """Forward pass of ONE volume with the tile rows dealt to the ranks.  ``compute_block((z0,z1),(y0,y1))`` returns
    the prediction of that image block evaluated as an independent volume (``FplNetwork.infer_device`` on the block:
    its origin lies on the reference tile grid, so every interior value equals the whole-volume one).  Returns
    ``(ext, e0)``: this rank's extended slab [e0, e0+len) with its OWNED planes complete (``plans[rank][1]``) and the
    ``halo`` margin planes still to be filled by the detection's halo exchange."""
    import torch
    import torch.distributed as dist
    dtype = dtype or torch.float32
    own0, own1 = plans[rank][1]
    e0, e1 = (max(0, own0 - halo), min(Z, own1 + halo)) if own1 > own0 else (own0, own0)
    ext = torch.zeros((e1 - e0, Y, X), dtype=dtype, device=device)
    sends, keep = [], []
    for piece in pieces[rank]:
        (zr, yr), ((pz0, pz1), (py0, py1)) = piece_geometry(piece, Z, Y, rf_offset, out_sz)
        if pz1 <= pz0 or py1 <= py0:
            continue
        sub = compute_block(zr, yr)                          # (z1-z0, y1-y0, X), border rf_offset wide = 0
        part = sub[rf_offset:rf_offset + (pz1 - pz0), rf_offset:rf_offset + (py1 - py0)]
        if piece[3] == rank:
            ext[pz0 - e0:pz1 - e0, py0:py1] = part
        else:
            buf = part.contiguous()
            keep.append(buf)
            sends.append((piece[3], buf))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        ops, recvs = [], []
        for dst, buf in sends:
            ops.append(dist.P2POp(dist.isend, buf, dst, group=group))
        for src, theirs in enumerate(pieces):               # every rank knows the whole plan: post the matching receives
            if src == rank:
                continue
            for piece in theirs:
                if piece[3] != rank:
                    continue
                _, ((pz0, pz1), (py0, py1)) = piece_geometry(piece, Z, Y, rf_offset, out_sz)
                if pz1 <= pz0 or py1 <= py0:
                    continue
                tmp = torch.empty((pz1 - pz0, py1 - py0, X), dtype=dtype, device=device)
                recvs.append((tmp, (pz0, pz1, py0, py1)))
                ops.append(dist.P2POp(dist.irecv, tmp, src, group=group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for tmp, (pz0, pz1, py0, py1) in recvs:
            ext[pz0 - e0:pz1 - e0, py0:py1] = tmp
    else:
        assert not sends, "pieces for other ranks need torch.distributed"
    return ext, e0


def detect_volume_rows_sharded(network, image_slab, slab_z0, Z, pieces, plans, normalize, obj_min_dist, smoothing_sigma,
                               volume_offset=(0, 0, 0), buffer_sz=0, thd=0, group=None, return_stats=False):
    """``detect_volume_sharded`` with tile-row load balance (U-Nets): ``image_slab`` holds the image planes
    [slab_z0, slab_z0+len) this rank's pieces read (``image_planes_for_pieces``)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    Y, X = int(image_slab.shape[1]), int(image_slab.shape[2])
    off, out_sz = int(network.rf_offset[0]), int(network.infer_sz[0]) - 2 * int(network.rf_offset[0])

    def compute_block(zr, yr):
        blk = image_slab[zr[0] - slab_z0:zr[1] - slab_z0, yr[0]:yr[1]].contiguous()
        return network.infer_device(blk, normalize=normalize)

    h = halo_planes(obj_min_dist, smoothing_sigma)
    ext, e0 = infer_rows_sharded(compute_block, pieces, plans, rank, Z, Y, X, off, out_sz, h, image_slab.device, group)
    own0, own1 = plans[rank][1]
    pred = ext[own0 - e0:own1 - e0]
    coll = _DistCollectives(group, all_ranges=[p[1] for p in plans])
    return voxel2obj_global([pred], [(own0, own1)], Z, obj_min_dist, smoothing_sigma, volume_offset, buffer_sz, thd,
                            coll=coll, return_stats=return_stats, ext_slabs=[ext if own1 > own0 else None])


def halo_planes(obj_min_dist, smoothing_sigma):
    """Planes of the probability map a rank needs beyond its own on each side for the exact-global voxel2obj:
    r (suppression ball) + lw (Gaussian half width)."""
    from . import fplobjdetect as P
    _, lw = P._gaussian_taps(smoothing_sigma)
    return int(obj_min_dist) + max(int(lw), 0)


def detect_volume_sharded(network, image_slab, Z, plans, normalize, obj_min_dist, smoothing_sigma,
                          volume_offset=(0, 0, 0), buffer_sz=0, thd=0, group=None, return_stats=False):
    """The whole T-bar path on ONE (Z,Y,X) volume sharded as z-slabs over the ranks of ``group`` (one process per
    GPU): replaces the reference's tower replication (flypylib/multi_gpu.py:20-61, fplnetwork.py:130-134) followed
    by a single ``voxel2obj``.  ``image_slab`` holds this rank's planes ``plans[rank][0]`` (``shard_plan``).
    Forward pass: ``FplNetwork.infer_slab_device`` -- receptive-field halo only, no communication.  Detection:
    ``voxel2obj_global`` (exact-global semantics): every rank returns the detection list of the single-GPU
    call, bit for bit."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    (in0, in1), (own0, own1) = plans[rank]
    Y, X = int(image_slab.shape[1]), int(image_slab.shape[2])
    ext = None
    if in1 > in0:
        # the forward pass writes this rank's planes straight into the extended slab the detection works on (its halo
        # margins are filled by the exchange): no staging copy of the probability map
        h = halo_planes(obj_min_dist, smoothing_sigma)
        e0, e1 = max(0, own0 - h), min(Z, own1 + h)
        ext = torch.empty((e1 - e0, Y, X), dtype=torch.float32, device=image_slab.device)
        _, first, last = network.infer_slab_device(image_slab, Z, in0, normalize=normalize, pred=ext, pred_z0=e0)
        assert (first, last) == (own0, own1), ((first, last), (own0, own1))
        pred = ext[own0 - e0:own1 - e0]
    else:
        pred = torch.zeros((0, Y, X), dtype=torch.float32, device=image_slab.device)
    coll = _DistCollectives(group, all_ranges=[p[1] for p in plans])
    return voxel2obj_global([pred], [(own0, own1)], Z, obj_min_dist, smoothing_sigma, volume_offset, buffer_sz, thd,
                            coll=coll, return_stats=return_stats, ext_slabs=[ext])


# ---------------------------------------------------------------------------------------------------
# Exact global voxel2obj on z-slabs (semantics S2): the result is bit-identical to ONE
# voxel2obj call on the whole volume (flypylib/fplobjdetect.py:132-257), whatever the number of ranks.
#   1. halo exchange of the probability map: r + lw planes per cut (lw = Gaussian half width), so that every
#      rank can smooth its owned planes plus an r-wide halo exactly;
#   2. the 97th percentile of the WHOLE padded volume: three radix levels, each a per-rank histogram of the
#      owned planes + a sum all-reduce of 2048 counters (the border zeros of the padded volume enter as a count);
#   3. greedy NMS as rounds: every rank decides only for the voxels it owns; the points selected in a round
#      are all-gathered and every rank suppresses all balls that reach into its extended slab -- its validity
#      map then equals the single-GPU one at the start of every round;
#   4. all-gather of the owned detections, global order (conf desc, flat index asc), buffer crop, offset.
# The collectives are tiny (KBs per round); `_DistCollectives` maps them to torch.distributed (NCCL on the
# GPU box), `_LocalCollectives` runs all ranks in one process on one GPU (tests).
# ---------------------------------------------------------------------------------------------------
class _DistCollectives(object):
    """One local rank; collectives over torch.distributed (device tensors -> NCCL)."""

    def __init__(self, group=None, all_ranges=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.local_ranks = [self.rank]
        self.all_ranges = [tuple(r) for r in all_ranges] if all_ranges is not None else None

    def halo(self, slabs, ranges, Z, h, ext=None):
        """-> [(extended slab, its first plane)].  ``ext`` (optional): the extended slab preallocated by the caller with
        the owned planes already in place (``slabs[0]`` is then a view of it) -- no staging copy."""
        import torch
        dist = self.dist
        slab, (z0, z1) = slabs[0], ranges[0]
        all_ranges = self.all_ranges
        if all_ranges is None:                                  # callers that know the plan pass it (no collective)
            all_ranges = [None] * self.world
            dist.all_gather_object(all_ranges, (z0, z1), group=self.group)
            self.all_ranges = [tuple(r) for r in all_ranges]
        assert tuple(all_ranges[self.rank]) == (z0, z1)
        if z1 <= z0:                                            # a rank without planes takes no part in the exchange
            return [(slab, z0)]
        e0, e1 = max(0, z0 - h), min(Z, z1 + h)
        if ext is not None and ext[0] is not None:
            ext = ext[0]
            assert int(ext.shape[0]) == e1 - e0
        else:
            ext = torch.empty((e1 - e0,) + tuple(slab.shape[1:]), dtype=slab.dtype, device=slab.device)
            ext[z0 - e0:z1 - e0] = slab
        ops = []
        for peer, (p0, p1) in enumerate(all_ranges):
            if peer == self.rank or p1 <= p0:
                continue
            pe0, pe1 = max(0, p0 - h), min(Z, p1 + h)
            lo, hi = max(z0, pe0), min(z1, pe1)                 # my planes the peer's extended slab needs
            if hi > lo:
                ops.append(dist.P2POp(dist.isend, slab[lo - z0:hi - z0].contiguous(), peer, group=self.group))
            lo, hi = max(p0, e0), min(p1, e1)                   # the peer's planes my extended slab needs
            if hi > lo:
                ops.append(dist.P2POp(dist.irecv, ext[lo - e0:hi - e0], peer, group=self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return [(ext, e0)]

    def known_ranges(self, ranges):
        return self.all_ranges

    def gather_blocks(self, blocks):
        """One all-gather of equal-sized blocks -> [(world, *block.shape)]."""
        import torch
        b = blocks[0]
        out = torch.empty(self.world * b.numel(), dtype=b.dtype, device=b.device)
        self.dist.all_gather_into_tensor(out, b.reshape(-1), group=self.group)
        return [out.view((self.world,) + tuple(b.shape))]

    def allreduce(self, tensors):
        self.dist.all_reduce(tensors[0], group=self.group)
        return tensors

    def allgather(self, tensors, scalars=None):
        """Concatenation of every rank's rows; with ``scalars`` also the sum of one integer per rank (the row
        counts and the scalars travel in the same small all-gather)."""
        import torch
        t = tensors[0]
        head = torch.tensor([t.shape[0], int(scalars[0]) if scalars is not None else 0], dtype=torch.int64,
                            device=t.device)
        heads = torch.empty(self.world * 2, dtype=torch.int64, device=t.device)
        self.dist.all_gather_into_tensor(heads, head, group=self.group)
        heads = heads.view(self.world, 2).cpu().numpy()
        counts, total = [int(c) for c in heads[:, 0]], int(heads[:, 1].sum())
        m = max(counts)
        if m == 0:
            out = [t[:0].clone()]
        else:
            pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            pad[:t.shape[0]] = t
            bufs = torch.empty(self.world * pad.numel(), dtype=t.dtype, device=t.device)
            self.dist.all_gather_into_tensor(bufs, pad.view(-1), group=self.group)
            bufs = bufs.view((self.world, m) + tuple(t.shape[1:]))
            out = [torch.cat([bufs[r, :c] for r, c in enumerate(counts)], 0)]
        return out if scalars is None else (out, total)


class _LocalCollectives(object):
    """All ranks in this process (same device): the collectives are plain tensor operations."""

    def __init__(self, world):
        self.world, self.rank, self.local_ranks = world, 0, list(range(world))

    def known_ranges(self, ranges):
        return list(ranges)

    def gather_blocks(self, blocks):
        import torch
        g = torch.stack(blocks, 0)
        return [g for _ in blocks]

    def halo(self, slabs, ranges, Z, h, ext=None):
        import torch
        full = torch.cat(slabs, 0)
        base = ranges[0][0]
        out = []
        for (z0, z1) in ranges:
            e0, e1 = max(0, z0 - h), min(Z, z1 + h)
            out.append((full[e0 - base:e1 - base].contiguous(), e0))
        return out

    def allreduce(self, tensors):
        total = sum(tensors[1:], tensors[0].clone())
        return [total.clone() for _ in tensors]

    def allgather(self, tensors, scalars=None):
        import torch
        cat = torch.cat(tensors, 0)
        out = [cat.clone() for _ in tensors]
        return out if scalars is None else (out, int(sum(int(v) for v in scalars)))


def _key2f(k):
    k = np.uint32(k)
    u = (k & np.uint32(0x7fffffff)) if (k & np.uint32(0x80000000)) else ~k
    return np.array([u], dtype=np.uint32).view(np.float32)[0]


def _scan_level(hist, rank, prefix, mask, shift, bins, extra_zeros):
    """One radix-select level on the host (the device version is select_scan_kernel in csrc/detect.cu):
    add the implicit border zeros to the class of +0.0, find the bin holding `rank`."""
    h = np.asarray(hist, dtype=np.uint64).copy()
    zkey = 0x80000000
    if (zkey & mask) == prefix:
        h[(zkey >> shift) & (bins - 1)] += np.uint64(extra_zeros)
    cum = np.cumsum(h[:bins], dtype=np.uint64)                    # counts stay far below 2^64
    sel = int(np.searchsorted(cum, np.uint64(rank), side='right'))
    sel = min(sel, bins - 1)
    before = int(cum[sel - 1]) if sel else 0
    return rank - before, prefix | (sel << shift), mask | ((bins - 1) << shift)


def voxel2obj_global(pred_slabs, ranges, Z, obj_min_dist, smoothing_sigma, volume_offset=(0, 0, 0), buffer_sz=0,
                     thd=0, coll=None, return_stats=False, ext_slabs=None):
    """voxel2obj of the whole (Z,Y,X) map whose planes [z0,z1) are held by different ranks; bit-identical to
    the single call.  ``pred_slabs`` / ``ranges``: CUDA float32 slabs and their (z0,z1) for the LOCAL ranks
    (one entry with torch.distributed, all ranks with ``_LocalCollectives``).  Every rank returns the full
    ``{'locs','conf'}`` dict."""
    import ctypes
    import torch
    from . import _lib, fplobjdetect as P
    if coll is None:
        coll = _DistCollectives()
    lib = _lib.lib()
    nl = len(pred_slabs)
    Y, X = int(pred_slabs[0].shape[1]), int(pred_slabs[0].shape[2])
    dev = pred_slabs[0].device
    devi = dev.index
    ctx = _lib.context(devi)
    st = _lib.current_stream_ptr(devi)
    p, _keep = P._make_params((Z, Y, X), obj_min_dist, smoothing_sigma, volume_offset, buffer_sz, thd)
    r, lw = int(p.obj_min_dist), max(int(p.lw), 0)
    empty = {'locs': np.zeros((0, 3)), 'conf': np.zeros(0)}
    stats = {'threshold': float('nan'), 'rounds': 0}
    stages = stats.setdefault('stage_ms', {}) if return_stats == 'stages' else None
    import time as _time
    _t = [_time.perf_counter()]

    def mark(name):
        """stage timing (only with return_stats='stages': it synchronises the device at every stage boundary)"""
        if stages is not None:
            torch.cuda.synchronize()
            now = _time.perf_counter()
            stages[name] = stages.get(name, 0.0) + (now - _t[0]) * 1e3
            _t[0] = now

    def done(out):
        return (out, stats) if return_stats else out

    with torch.cuda.device(devi):
        # 1. halo exchange + exact smoothing of [z0-r, z1+r)
        mark('setup')
        # (ext_slabs: extended slabs preallocated by the caller, owned planes already in place -- see halo_planes())
        ext = coll.halo([s.contiguous() for s in pred_slabs], ranges, Z, r + lw, ext=ext_slabs)
        mark('halo_exchange')
        smooth, s_lo = [], []
        for (e, e0), (z0, z1) in zip(ext, ranges):
            d_s = torch.empty_like(e)
            if e.shape[0] > 0:
                _lib.check(lib.fpl_v2o_smooth(ctx.handle, e.data_ptr(), int(e.shape[0]), Y, X, ctypes.byref(p),
                                              d_s.data_ptr(), st), "fpl_v2o_smooth")
            s0, s1 = max(0, z0 - r), min(Z, z1 + r)
            smooth.append(d_s[s0 - e0:s1 - e0]); s_lo.append(s0)
        del ext
        mark('smooth_exact')
        # 2. global percentile (np.percentile of the padded volume): three all-reduced radix levels
        n_pad = (Z + 2 * r) * (Y + 2 * r) * (X + 2 * r)
        extra = n_pad - Z * Y * X
        targets = [int(p.rank_lo)] + ([int(p.rank_hi)] if p.rank_hi != p.rank_lo else [])
        states = [[t, 0, 0] for t in targets]                       # [rank, prefix, mask]
        for li, (shift, bins) in enumerate(((21, 2048), (10, 2048), (0, 1024))):
            seen = {}                                               # the two order statistics usually share a class
            for stt in states:
                key = (stt[1], stt[2])
                if key not in seen:
                    hs = []
                    for sm, lo, (z0, z1) in zip(smooth, s_lo, ranges):
                        own = sm[z0 - lo:z1 - lo]
                        h = torch.zeros(2049, dtype=torch.int64, device=dev)
                        if own.numel():
                            _lib.check(lib.fpl_v2o_hist_level(ctx.handle, own.data_ptr(), int(own.numel()), stt[1], stt[2],
                                                              shift, bins, h.data_ptr(),
                                                              h.data_ptr() + 2048 * 8 if li == 0 else None, st),
                                       "fpl_v2o_hist_level")
                        hs.append(h)
                    seen[key] = coll.allreduce(hs)[0].cpu().numpy()
                tot = seen[key]
                if li == 0 and int(tot[2048]) > 0:                  # NaNs: percentile is NaN, nothing is selected
                    return done(empty)
                stt[0], stt[1], stt[2] = _scan_level(tot[:2048], stt[0], stt[1], stt[2], shift, bins, extra)
        a, b = _key2f(states[0][1]), _key2f(states[-1][1])
        g = np.float32(p.gamma)
        dlt = np.float32(b - a)
        res = np.float32(a + np.float32(dlt * g))
        if g >= np.float32(0.5):
            res = np.float32(b - np.float32(dlt * np.float32(np.float32(1.0) - g)))
        pr = float(res)
        thresh = float('nan') if (pr != pr or p.thd != p.thd) else max(pr, float(p.thd))
        stats['threshold'] = thresh
        mark('percentile_3_levels_allreduce')
        # 3. NMS rounds
        cand_bound = max(1, n_pad - targets[0])
        sessions, sel_bufs = [], []
        for sm, lo, (z0, z1) in zip(smooth, s_lo, ranges):
            ze = int(sm.shape[0])
            if ze == 0 or z1 <= z0:
                sessions.append(None); sel_bufs.append(None); continue
            cap = P._default_capacity((ze, Y, X), r)
            h = ctypes.c_void_p()
            ncand = ctypes.c_int64()
            _lib.check(lib.fpl_v2o_slab_begin(ctx.handle, sm.data_ptr(), ze, Y, X, ctypes.byref(p), thresh, z0 - lo, z1 - lo,
                                              int(min(ze * Y * X, cand_bound)), cap, ctypes.byref(h), ctypes.byref(ncand), st),
                       "fpl_v2o_slab_begin")
            sessions.append(h); sel_bufs.append(None)
        mark('dense_pass_local_maxima')
        # every round: decision half on every rank -> ONE all-gather of fixed-size blocks (header + selected points) ->
        # update half; the host reads the gathered headers once per round (termination, overflow)
        all_rng = coll.known_ranges(ranges)
        max_planes = max([b_ - a_ for a_, b_ in all_rng] + [1])
        sel_cap = P._default_capacity((max_planes + 2 * r, Y, X), r)
        blocks = [torch.zeros((1 + sel_cap, 3), dtype=torch.int64, device=dev) for _ in sessions]
        while True:
            for h, blk, lo in zip(sessions, blocks, s_lo):
                if h is not None:
                    _lib.check(lib.fpl_v2o_slab_round_pack(h, blk.data_ptr(), sel_cap, lo, st), "fpl_v2o_slab_round_pack")
            mark('round_decide')
            gathered = coll.gather_blocks(blocks)
            mark('round_allgather')
            for h, lo, gb in zip(sessions, s_lo, gathered):
                if h is not None:
                    _lib.check(lib.fpl_v2o_slab_apply_blocks(h, gb.data_ptr(), int(gb.shape[0]), sel_cap, lo, st),
                               "fpl_v2o_slab_apply_blocks")
            heads = gathered[0][:, 0, :].cpu().numpy()              # the one host synchronisation of the round
            mark('round_suppress')
            if heads[:, 2].any():
                raise RuntimeError("voxel2obj_global: list overflow in an NMS round (capacity %d)" % sel_cap)
            if int(heads[:, 1].sum()) == 0:
                break
            stats['rounds'] += 1
            if int(heads[:, 0].sum()) == 0:
                raise RuntimeError("voxel2obj_global: NMS round made no progress (internal error)")
        # 4. owned detections -> global list
        rows = []
        for h, lo, sm in zip(sessions, s_lo, smooth):
            if h is None:
                rows.append(torch.zeros((0, 4), dtype=torch.float64, device=dev)); continue
            cap = P._default_capacity((int(sm.shape[0]), Y, X), r)
            out = torch.empty((cap, 4), dtype=torch.float64, device=dev)
            cnt, rounds = ctypes.c_int64(), ctypes.c_int64()
            _lib.check(lib.fpl_v2o_slab_end(h, out.data_ptr(), cap, ctypes.byref(cnt), ctypes.byref(rounds), st),
                       "fpl_v2o_slab_end")
            o = out[:cnt.value].clone()
            o[:, 0] += lo
            rows.append(o)
        ar = coll.allgather(rows)[0]
        if ar.shape[0] > 1:
            # emission order of the greedy loop = (conf desc, flat index asc): two stable sorts on the device (a host
            # np.lexsort of the 7e4 detections of a 1024^3 volume costs 5 ms -- a sixth of the 8-GPU step)
            flat = (ar[:, 0] * Y + ar[:, 1]) * X + ar[:, 2]            # exact in float64 (< 2^53)
            ar = ar[torch.argsort(flat, stable=True)]
            ar = ar[torch.argsort(ar[:, 3], descending=True, stable=True)]
        allrows = ar.cpu().numpy()
        mark('final_allgather')
    if allrows.shape[0] == 0:
        return done(empty)
    z, y, x, c = allrows[:, 0], allrows[:, 1], allrows[:, 2], allrows[:, 3]
    bx, by, bz = (int(p.buffer_xyz[i]) for i in range(3))
    keep = (x >= bx) & (y >= by) & (z >= bz) & (x < X - bx) & (y < Y - by) & (z < Z - bz)
    locs = np.stack([x[keep] + p.offset_xyz[0], y[keep] + p.offset_xyz[1], z[keep] + p.offset_xyz[2]], 1)
    return done({'locs': locs, 'conf': c[keep].copy()})
