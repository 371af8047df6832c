"""CPU (gloo, world_size 2): host-side logic of the N>1 path -- layer partition, slab bounds,
variable-length detection all-gather, merge order."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flypylib_b200 import multi_gpu


def test_partition_and_slabs():
    assert multi_gpu.partition_layers(25, 8) == [(0, 4), (4, 7), (7, 10), (10, 13), (13, 16), (16, 19), (19, 22), (22, 25)]
    assert multi_gpu.partition_layers(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]
    # unet_like2 on 2048: 25 layers of 82
    assert multi_gpu.tile_layers(2048, 9, 82) == 25
    assert multi_gpu.tile_layers(1024, 10, 80) == 13
    assert multi_gpu.tile_layers(15, 9, 82) == 0
    (z0, z1), (p0, p1) = multi_gpu.slab_for_layers(4, 7, 2048, 9, 82)
    assert (z0, z1) == (328, 592) and (p0, p1) == (337, 583)          # halo = 2*rf_offset = 18 planes
    (z0, z1), (p0, p1) = multi_gpu.slab_for_layers(22, 25, 2048, 9, 82)
    assert z1 == 2048 and p1 == 2039
    # slabs of consecutive ranks tile the prediction rows exactly
    rows = []
    for b, e in multi_gpu.partition_layers(25, 8):
        rows.append(multi_gpu.slab_for_layers(b, e, 2048, 9, 82)[1])
    assert rows[0][0] == 9 and rows[-1][1] == 2039
    assert all(rows[i][1] == rows[i + 1][0] for i in range(7))


def test_merge_order():
    a = np.array([[1, 2, 3, 0.5], [4, 5, 6, 0.9]])
    b = np.array([[7, 8, 9, 0.9], [0, 0, 0, 0.1]])
    m = multi_gpu.merge_detections([a, b, np.zeros((0, 4))])
    assert np.array_equal(m['conf'], [0.9, 0.9, 0.5, 0.1])
    assert np.array_equal(m['locs'][0], [4, 5, 6]) and np.array_equal(m['locs'][1], [7, 8, 9])
    e = multi_gpu.merge_detections([])
    assert e['locs'].shape == (0, 3) and e['conf'].shape == (0,)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(rank)
        k = [3, 0][rank] if world == 2 else rank
        rows = torch.from_numpy(np.concatenate([rng.integers(0, 100, (k, 3)).astype(np.float64),
                                                rng.random((k, 1))], 1))
        parts = multi_gpu.allgather_detections(rows)
        assert len(parts) == world
        assert parts[rank].shape == (k, 4) and torch.equal(parts[rank], rows)
        merged = multi_gpu.merge_detections([p.numpy() for p in parts])
        q.put((rank, [tuple(p.shape) for p in parts], merged['conf'].tolist()))
    finally:
        dist.destroy_process_group()


def test_allgather_detections_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == res[1][1] == [(3, 4), (0, 4)]      # one rank contributes an empty list
    assert res[0][2] == res[1][2] and res[0][2] == sorted(res[0][2], reverse=True)


# ---- exact-global voxel2obj (S2): host logic and collectives -------------------------------------------
def test_radix_scan_levels_select_order_statistics():
    """multi_gpu._scan_level (host half of the all-reduced radix select) finds the same order statistic as a
    sort, including the implicit border zeros of the padded volume."""
    rng = np.random.default_rng(0)
    v = np.concatenate([np.abs(rng.standard_normal(150000)) * 0.3, -rng.random(5000), np.zeros(300)]).astype(np.float32)

    def f2key(a):
        u = a.view(np.uint32)
        return np.where(u & np.uint32(0x80000000), ~u, u | np.uint32(0x80000000)).astype(np.uint32)

    keys = f2key(v)
    extra = 40000
    allv = np.sort(np.concatenate([v, np.zeros(extra, np.float32)]))
    for target in [0, 4999, 5000, 30000, 45299, 45300, 120000, v.size + extra - 1]:
        rank, prefix, mask = target, 0, 0
        for shift, bins in ((21, 2048), (10, 2048), (0, 1024)):
            sel = (keys & np.uint32(mask)) == np.uint32(prefix)
            h = np.bincount(((keys[sel] >> np.uint32(shift)) & np.uint32(bins - 1)).astype(np.int64), minlength=2048)
            rank, prefix, mask = multi_gpu._scan_level(h, rank, prefix, mask, shift, bins, extra)
        assert mask == 0xffffffff
        got = multi_gpu._key2f(prefix)
        assert got == allv[target] or (got == 0 and allv[target] == 0), (target, got, allv[target])


def _coll_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        coll = multi_gpu._DistCollectives()
        Z, h = 23, 4
        full = torch.arange(Z * 3 * 2, dtype=torch.float32).view(Z, 3, 2)
        cuts = [(0, 3), (3, 5), (5, 23)][:world] if world == 3 else [(0, 9), (9, 23)]
        z0, z1 = cuts[rank]
        (ext, e0), = coll.halo([full[z0:z1].contiguous()], [(z0, z1)], Z, h)
        ok_halo = e0 == max(0, z0 - h) and torch.equal(ext, full[e0:min(Z, z1 + h)])
        tot = coll.allreduce([torch.tensor([rank + 1, 10 * (rank + 1)])])[0].tolist()
        rows = torch.full((rank + 1, 3), rank, dtype=torch.int64)
        gathered = coll.allgather([rows])[0]
        (g2,), total = coll.allgather([rows[:0]], scalars=[rank + 5])      # nobody has rows; scalars still summed
        assert g2.shape == (0, 3) and total == sum(r + 5 for r in range(world))
        (g3,), total3 = coll.allgather([rows], scalars=[1])
        assert torch.equal(g3, gathered) and total3 == world
        # a rank without planes (more ranks than slabs) takes no part in the halo exchange; plan known up front
        cuts2 = [(0, 12), (12, 23), (23, 23)][:world] if world == 3 else [(0, 23), (23, 23)]
        coll2 = multi_gpu._DistCollectives(all_ranges=cuts2)
        a0, a1 = cuts2[rank]
        (ext2, f0), = coll2.halo([full[a0:a1].contiguous()], [(a0, a1)], Z, h)
        ok_halo = ok_halo and (a1 == a0 or (f0 == max(0, a0 - h) and torch.equal(ext2, full[f0:min(Z, a1 + h)])))
        q.put((rank, bool(ok_halo), tot, gathered.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_s2_collectives_gloo(world):
    """Halo exchange (slabs thinner than the halo: several peers feed one extended slab), sum all-reduce and
    variable-length all-gather of multi_gpu._DistCollectives over gloo."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + os.getpid() % 300 + world
    procs = [ctx.Process(target=_coll_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    s = sum(range(1, world + 1))
    want_rows = [[r, r, r] for r in range(world) for _ in range(r + 1)]
    for rank, ok_halo, tot, gathered in res:
        assert ok_halo, rank
        assert tot == [s, 10 * s]
        assert gathered == want_rows


def test_shard_plan_partitions_the_volume():
    """multi_gpu.shard_plan: owned prediction planes partition [0,Z); every rank reads its planes plus a
    2*rf_offset halo; cuts lie on the network grid (SURVEY 8e 'Forward pass')."""
    for Z, off, gran, world in [(1024, 10, 4, 8), (1024, 10, 4, 2), (2048, 9, 82, 8), (256, 7, 4, 4), (333, 10, 4, 5),
                                (30, 10, 4, 4), (15, 10, 4, 2), (270, 9, 82, 3), (100, 9, 82, 2)]:
        plans = multi_gpu.shard_plan(Z, off, gran, world)
        assert len(plans) == world
        pos = 0
        for (in0, in1), (own0, own1) in plans:
            if own1 == own0:
                assert in1 == in0
                continue
            assert own0 == pos
            pos = own1
            assert in0 % gran == 0 and in0 <= own0 and in1 >= own1
            if in0 > 0:
                assert own0 == in0 + off
            if in1 < Z:
                assert own1 == in1 - off and (in1 - in0 - 2 * off) % gran == 0
        assert pos == Z
    # the north-star split: unet_like2 on 2048^3 over 8 ranks = whole tile layers 4,3,3,3,3,3,3,3; 18 halo planes
    plans = multi_gpu.shard_plan(2048, 9, 82, 8)
    assert [(p[0][1] - p[0][0] - 18 + 81) // 82 for p in plans] == [4, 3, 3, 3, 3, 3, 3, 3]
    # VGG on 1024^3: 251 groups of 4 planes -> 32/31 groups per rank (balance 0.98)
    plans = multi_gpu.shard_plan(1024, 10, 4, 8)
    assert sorted({p[1][1] - p[1][0] for p in plans[1:-1]}) == [124, 128]


def test_row_plan_covers_every_tile_row_once():
    """multi_gpu.row_plan: the (z, y) tile rows of the reference grid are dealt contiguously and evenly; every row
    is evaluated by exactly one rank; the prediction blocks of the pieces tile [off, Z-off) x [off, Y-off)."""
    for Z, Y, off, out, world in [(2048, 2048, 9, 82, 8), (270, 200, 9, 82, 3), (300, 120, 6, 90, 4), (100, 100, 9, 82, 2)]:
        pieces, plans = multi_gpu.row_plan(Z, Y, off, out, world)
        nz, ny = multi_gpu.tile_layers(Z, off, out), multi_gpu.tile_layers(Y, off, out)
        seen = np.zeros((nz, ny), int)
        cover = np.zeros((Z, Y), int)
        counts = []
        for r, mine in enumerate(pieces):
            counts.append(sum(yb - ya for _, ya, yb, _ in mine))
            for piece in mine:
                kz, ya, yb, owner = piece
                seen[kz, ya:yb] += 1
                o0, o1 = plans[owner][1]
                (zr, yr), ((pz0, pz1), (py0, py1)) = multi_gpu.piece_geometry(piece, Z, Y, off, out)
                assert o0 <= pz0 and pz1 <= o1                      # the owner of the layer owns these planes
                assert zr[0] == kz * out and yr[0] == ya * out and zr[1] <= Z and yr[1] <= Y
                cover[pz0:pz1, py0:py1] += 1
            lo, hi = multi_gpu.image_planes_for_pieces(mine, Z, Y, off, out)
            assert all(lo <= multi_gpu.piece_geometry(p, Z, Y, off, out)[0][0][0] for p in mine)
        assert (seen == 1).all()
        assert max(counts) - min(counts) <= 1
        assert (cover[off:Z - off, off:Y - off] == 1).all() and cover.sum() == (Z - 2 * off) * (Y - 2 * off)
    pieces, _ = multi_gpu.row_plan(2048, 2048, 9, 82, 8)
    assert [sum(yb - ya for _, ya, yb, _ in m) for m in pieces] == [79] + [78] * 7       # 625 rows: balance 0.99


def _rows_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Z, Y, X, off, out, halo = 60, 50, 7, 2, 8, 3
        full = torch.zeros((Z, Y, X))
        full[off:Z - off, off:Y - off, off:X - off] = torch.arange((Z - 2 * off) * (Y - 2 * off) * (X - 2 * off),
                                                                 dtype=torch.float32).view(Z - 2 * off, Y - 2 * off, X - 2 * off) + 1

        def compute_block(zr, yr):            # "prediction of the block evaluated on its own": right values inside, 0 border
            sub = full[zr[0]:zr[1], yr[0]:yr[1]].clone()
            sub[:off] = 0; sub[:, :off] = 0
            if zr[1] < Z: sub[-off:] = 0
            if yr[1] < Y: sub[:, -off:] = 0
            return sub
        pieces, plans = multi_gpu.row_plan(Z, Y, off, out, world)
        ext, e0 = multi_gpu.infer_rows_sharded(compute_block, pieces, plans, rank, Z, Y, X, off, out, halo, torch.device("cpu"))
        own0, own1 = plans[rank][1]
        ok = torch.equal(ext[own0 - e0:own1 - e0], full[own0:own1]) and e0 == max(0, own0 - halo)
        q.put((rank, bool(ok), [len(m) for m in pieces]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_row_sharded_forward_exchange_gloo(world):
    """infer_rows_sharded over gloo with a stand-in for the network: the rows a rank evaluates for a neighbour's layer
    arrive at the plane owner; every rank ends up with exactly its owned planes of the whole-volume result."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30300 + os.getpid() % 300 + world
    procs = [ctx.Process(target=_rows_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
