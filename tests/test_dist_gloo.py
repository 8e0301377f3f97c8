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


def test_snake_partition_is_a_partition():
    from vnlb_b200.dist import partition_rows_snake
    for h, world in [(480, 2), (3840, 8), (960, 4)]:
        rows = []
        for r in range(world):
            for (a, b) in partition_rows_snake(h, 7, world, r):
                rows += list(range(a, min(b, h - 6)))
        assert sorted(rows) == list(range(h - 6))


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


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vnlb_b200.dist import allreduce_accumulators, partition_rows
    from vnlb_b200.utils import AttrDict
    # every rank aggregates the groups of its own band of reference pixels (oracle kernels on CPU),
    # then the accumulators are summed: the result must equal the single-process aggregation.
    T, C, H, W, K = 3, 3, 24, 28, 6
    rs = np.random.RandomState(0)
    ty = rs.randint(0, H - 6, (40, K)); tx = rs.randint(0, W - 6, (40, K)); tt = rs.randint(0, T - 1, (40, K))
    inds = (tt * C * H * W + ty * W + tx).astype(np.int64)
    patches = (rs.rand(40, K, 2, C, 7, 7) * 255).astype(np.float32)
    ref_y = ty[:, 0]                                           # band membership by the reference pixel's row
    y0, y1 = partition_rows(H, 7, world, rank)
    mine = (ref_y >= y0) & (ref_y < y1)
    deno = np.zeros((T, C, H, W), np.float32); weights = np.zeros((T, H, W), np.float32)
    orc.agg_patches(deno, weights, patches[mine], inds[mine])
    images = AttrDict(deno=torch.from_numpy(deno), weights=torch.from_numpy(weights))
    allreduce_accumulators(images)
    if rank == 0:
        fd = np.zeros((T, C, H, W), np.float32); fw = np.zeros((T, H, W), np.float32)
        orc.agg_patches(fd, fw, patches, inds)
        out["w_equal"] = bool(np.array_equal(images.weights.numpy(), fw))
        out["d_close"] = bool(np.allclose(images.deno.numpy(), fd, rtol=1e-5, atol=1e-3))
        out["covered"] = int(mine.sum())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_band_aggregation_plus_allreduce_equals_single_process():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert out["w_equal"] and out["d_close"] and 0 < out["covered"] < 40
