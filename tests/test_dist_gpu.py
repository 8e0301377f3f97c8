"""Multi-GPU path on real GPUs (skipped with fewer than 2): vnlb_b200.dist.denoise_distributed over NCCL on a band + halo
partition must match the single-GPU call within the PSNR tolerance of the north star (0.02 dB); max-abs parity is not
promised across band borders (per-rank greedy masks)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, use_flow, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import vnlb_b200
    from vnlb_b200 import dist as vdist, synth
    T, H, W, sigma = 8, 240, 320, 20.
    clean, flows = synth.synth_video(T, H, W, 123, return_flows=True)
    noisy = synth.add_noise(clean, sigma, 123)
    flows = flows if use_flow else None
    stats = {}
    deno, basic, _ = vdist.denoise_distributed(torch.from_numpy(noisy).pin_memory(), sigma, flows=flows, stats=stats,
                                               device=dev, max_flow=2.0 if use_flow else None)
    if rank == 0:
        d1, b1, _ = vnlb_b200.denoise(noisy, sigma, gpuid=0, verbose=False, flows=flows)
        ps = lambda x: float(vnlb_b200.compute_psnrs(x.cpu().numpy(), clean).mean())
        out["psnr_n"] = (ps(deno), ps(basic))
        out["psnr_1"] = (ps(d1), ps(b1))
        out["maxabs"] = float((deno - d1).abs().max())
        # away from the band border the two runs see the same neighbourhoods but not the same greedy draws
        out["layout"] = stats["layout"]
        out["exchange_bytes"] = stats["exchange_bytes"]
        out["groups"] = stats["ngroups"]
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("use_flow", [False, True])
def test_two_gpus_match_one_gpu_within_psnr_tolerance(use_flow):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, use_flow, out), nprocs=world, join=True)
        out = dict(out)
    print(out)
    assert abs(out["psnr_n"][0] - out["psnr_1"][0]) <= 0.02, out       # final estimate
    assert abs(out["psnr_n"][1] - out["psnr_1"][1]) <= 0.02, out       # basic estimate
    lay = out["layout"]
    assert lay["rows_copied"] < lay["rows_total"]                        # band + halo only
    frame_bytes = 8 * 4 * 240 * 320 * 4
    assert 0 < out["exchange_bytes"]["accumulators"] < frame_bytes // 2  # border strips, not frames
