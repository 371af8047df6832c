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
