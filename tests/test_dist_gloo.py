"""Host-side logic of the multi-GPU path on CPU: band partition and the
accumulator all-reduce over a world_size-2 gloo group (no GPU needed)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vnlb_oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_partition_rows_covers_every_reference_row_once():
    from vnlb_b200.dist import partition_rows
    for h, ps, world in [(480, 7, 1), (480, 7, 2), (1080, 7, 8), (64, 7, 4), (33, 3, 3)]:
        bands = [partition_rows(h, ps, world, r) for r in range(world)]
        assert bands[0][0] == 0 and bands[-1][1] == h
        for a, b in zip(bands[:-1], bands[1:]):
            assert a[1] == b[0]
        sizes = [min(b[1], h - ps + 1) - b[0] for b in bands]
        assert max(sizes) - min(sizes) <= 1 and min(sizes) > 0


def test_weighted_partition_balances_and_covers():
    from vnlb_b200.dist import partition_rows_weighted
    rs = np.random.RandomState(0)
    for h, world in [(480, 2), (3840, 8), (64, 8)]:
        w = torch.from_numpy(rs.rand(h).astype(np.float32) * (1 + 4 * (np.arange(h) > h // 2)))
        bands = [partition_rows_weighted(w, 7, world, r) for r in range(world)]
        assert bands[0][0] == 0 and bands[-1][1] == h
        for a, b in zip(bands[:-1], bands[1:]):
            assert a[1] == b[0] and a[1] > a[0]
        if h > 400:
            sums = [float(w[a:min(b, h - 6)].sum()) for a, b in bands]
            assert max(sums) / (sum(sums) / world) < 1.05


def test_halo_and_exchange_plan():
    """halo = window radius + patch extent + flow drift; give/take are mirror images of each other across ranks and
    only neighbours exchange when the bands are wider than the halo."""
    from vnlb_b200 import get_params
    from vnlb_b200.dist import exchange_plan, halo_rows, make_layout
    params = get_params(20.)
    assert halo_rows(params, 0.0) == 13 + 6
    assert halo_rows(params, 2.0) == 13 + 6 + 15            # 6 frames x (2 + 0.5) px
    for h, world in [(1080, 8), (480, 2), (1080, 4)]:
        halo = halo_rows(params, 2.0)
        bands, tiles = make_layout(h, 7, world, halo)
        assert tiles[0][0] == 0 and tiles[-1][1] == h
        plans = [exchange_plan(bands, tiles, r) for r in range(world)]
        for r in range(world):
            give, take = plans[r]
            assert set(give) <= {r - 1, r + 1} and set(take) <= {r - 1, r + 1}
            for s, rows in give.items():
                assert plans[s][1][r] == rows                # what I give s is what s takes from me
                assert rows[0] >= bands[s][0] and rows[1] <= bands[s][1]
    # bands narrower than the halo: the plan reaches beyond the neighbours and still mirrors
    bands, tiles = make_layout(64, 7, 8, 19)
    plans = [exchange_plan(bands, tiles, r) for r in range(8)]
    assert any(abs(s - 3) > 1 for s in plans[3][0])
    for r in range(8):
        for s, rows in plans[r][0].items():
            assert plans[s][1][r] == rows


def test_rebalanced_cuts_equalise_predicted_time_within_the_margin():
    """Host arithmetic of the mid-step load balancing: equal speeds and equal remaining work leave the borders alone; a
    slow rank gives rows away; borders never move more than `margin` rows from the layout and bands never vanish."""
    from vnlb_b200.dist import partition_rows, rebalanced_cuts
    h, ps, world, margin = 1080, 7, 8, 14
    bands = [partition_rows(h, ps, world, r) for r in range(world)]
    rem = np.full(h, 1000.0)
    rem[h - ps + 1:] = 0
    same = rebalanced_cuts(rem, [50.0] * world, bands, bands, margin, ps)
    assert all(abs(a[0] - b[0]) <= 1 for a, b in zip(same, bands))
    speeds = [50.0] * world
    speeds[3] = 40.0                                            # rank 3 is 20 % slower: it must shrink
    new = rebalanced_cuts(rem, speeds, bands, bands, margin, ps)
    assert new[0][0] == 0 and new[-1][1] == h
    for a, b in zip(new[:-1], new[1:]):
        assert a[1] == b[0] and a[1] - a[0] >= ps
    assert (new[3][1] - new[3][0]) < (bands[3][1] - bands[3][0])
    assert all(abs(n[0] - o[0]) <= margin for n, o in zip(new, bands))
    t_old = [rem[a:b].sum() / speeds[r] for r, (a, b) in enumerate(bands)]
    t_new = [sum(rem[y] / speeds[[r for r, (a, b) in enumerate(bands) if a <= y < b][0]] for y in range(a2, b2))
             for (a2, b2) in new]
    assert max(t_new) < max(t_old)                             # the slowest rank finishes earlier
    # a second rebalance starts from the moved bands but stays inside the margin of the ORIGINAL layout
    again = rebalanced_cuts(rem, speeds, new, bands, margin, ps)
    assert all(abs(n[0] - o[0]) <= margin for n, o in zip(again, bands))
    # degenerate: nothing left anywhere
    assert rebalanced_cuts(np.zeros(h), speeds, bands, bands, margin, ps) == bands


def _migrate_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vnlb_b200.dist import BandRebalancer, make_layout
    T, H, W, halo, margin = 2, 60, 8, 5, 4
    bands, tiles = make_layout(H, 7, world, halo, None, margin)
    ya, yb = tiles[rank]
    glob = (np.random.RandomState(1).rand(T, H, W) < 0.5).astype(np.int8)      # the "current" global mask state
    mask = torch.zeros((T, yb - ya, W), dtype=torch.int8)
    y0, y1 = bands[rank]
    mask[:, y0 - ya:y1 - ya] = torch.from_numpy(glob[:, y0:y1])
    reb = BandRebalancer(bands, tiles, rank, None, margin, H, W, T, 7, 2, 3)
    new = [(0, bands[0][1] + 3), (bands[0][1] + 3, H)] if world == 2 else \
        [(0, bands[0][1] - 2), (bands[0][1] - 2, bands[1][1] + 4), (bands[1][1] + 4, H)]
    reb._migrate(mask, new)
    n0, n1 = new[rank]
    ok = np.array_equal(mask[:, n0 - ya:n1 - ya].numpy(), glob[:, n0:n1])       # my new band carries the current state
    outside = mask.clone()
    outside[:, n0 - ya:n1 - ya] = 0
    out["ok_%d" % rank] = bool(ok) and int(outside.abs().sum()) == 0              # and nothing is set outside it
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("world", [2, 3])
def test_mask_rows_follow_the_moving_borders(world):
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_migrate_worker, args=(world, port, out), nprocs=world, join=True)
        assert all(out["ok_%d" % r] for r in range(world)), dict(out)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vnlb_b200.dist import exchange_accumulators, exchange_halo, exchange_plan, gather_bands, make_layout
    # Every rank aggregates the groups of its own band of reference pixels into the accumulators of its row TILE
    # (band + halo, tile-local indices; oracle kernels on CPU), the border rows are exchanged, every rank normalises:
    # the gathered result must equal the single-process aggregation + normalisation of all groups.
    T, C, H, W, K, G, halo = 3, 3, 40, 28, 6, 60, 9
    rs = np.random.RandomState(0)
    ref_y = rs.randint(0, H - 6, (G,))
    ty = np.clip(ref_y[:, None] + rs.randint(-3, 4, (G, K)), 0, H - 7)      # neighbours within the halo (3 + 6 rows)
    ty[:, 0] = ref_y
    tx = rs.randint(0, W - 6, (G, K)); tt = rs.randint(0, T - 1, (G, K))
    patches = (rs.rand(G, K, 2, C, 7, 7) * 255).astype(np.float32)
    noisy = (rs.rand(T, C, H, W) * 255).astype(np.float32)
    bands, tiles = make_layout(H, 7, world, halo)
    (y0, y1), (ya, yb) = bands[rank], tiles[rank]
    give, take = exchange_plan(bands, tiles, rank)
    mine = (ref_y >= y0) & (ref_y < y1)
    hb = yb - ya
    inds_local = (tt * C * hb * W + (ty - ya) * W + tx).astype(np.int64)
    deno = np.zeros((T, C, hb, W), np.float32); weights = np.zeros((T, hb, W), np.float32)
    orc.agg_patches(deno, weights, patches[mine], inds_local[mine])
    deno, weights = torch.from_numpy(deno), torch.from_numpy(weights)
    nbytes = exchange_accumulators(deno, weights, ya, give, take)
    dn = deno.numpy().copy()
    orc.normalize(dn, weights.numpy(), noisy[:, :, ya:yb])
    dn = torch.from_numpy(dn)
    exchange_halo(dn, ya, give, take)
    full = gather_bands(dn, ya, bands, rank, H)
    # single-process reference
    fd = np.zeros((T, C, H, W), np.float32); fw = np.zeros((T, H, W), np.float32)
    orc.agg_patches(fd, fw, patches, (tt * C * H * W + ty * W + tx).astype(np.int64))
    orc.normalize(fd, fw, noisy)
    ok_w = np.array_equal(weights.numpy()[:, y0 - ya:y1 - ya], fw[:, y0:y1])
    out["w_equal_%d" % rank] = bool(ok_w)
    out["full_close_%d" % rank] = bool(np.allclose(full.numpy(), fd, rtol=1e-5, atol=1e-3))
    out["tile_close_%d" % rank] = bool(np.allclose(dn.numpy(), fd[:, :, ya:yb], rtol=1e-5, atol=1e-3))   # halo rows too
    out["covered_%d" % rank] = int(mine.sum())
    out["bytes_%d" % rank] = int(nbytes)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("world", [2, 3])
def test_band_tiles_plus_border_exchange_equal_single_process(world):
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        for r in range(world):
            assert out["w_equal_%d" % r] and out["full_close_%d" % r] and out["tile_close_%d" % r], (r, dict(out))
            assert 0 < out["covered_%d" % r] < 60
        # only border strips travel: far less than a frame per rank (T x (C+1) x H x W x 4 = 53760 B)
        assert 0 < out["bytes_0"] < 3 * 4 * 40 * 28 * 4 // 2
