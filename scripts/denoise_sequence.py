#!/usr/bin/env python
"""Denoise a frame sequence with vnlb_b200 (equivalent of the reference's
scripts/process_video_sequence.py / example.py, which import packages that do not exist there).

  python scripts/denoise_sequence.py --frames DIR --sigma 20 --out OUT [--clean DIR] [--add-noise]
                                     [--fflow ff.npy --bflow bf.npy] [--nframes N] [--schedule fast|parity]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", required=True)
    ap.add_argument("--sigma", type=float, required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--clean", default=None, help="directory of clean frames for PSNR")
    ap.add_argument("--add-noise", action="store_true", help="treat --frames as clean and add N(0, sigma^2)")
    ap.add_argument("--fflow", default=None)
    ap.add_argument("--bflow", default=None)
    ap.add_argument("--nframes", type=int, default=None)
    ap.add_argument("--schedule", default="fast", choices=["fast", "parity"])
    ap.add_argument("--gpuid", type=int, default=0)
    args = ap.parse_args()
    import vnlb_b200
    from vnlb_b200 import video_io
    vid = video_io.read_video_sequence(args.frames, args.nframes)
    clean = video_io.read_video_sequence(args.clean, args.nframes) if args.clean else None
    if args.add_noise:
        clean = vid
        vid = (vid + np.random.RandomState(123).randn(*vid.shape).astype(np.float32) * args.sigma).astype(np.float32)
    flows = None
    if args.fflow and args.bflow:
        flows = {"fflow": np.load(args.fflow), "bflow": np.load(args.bflow)}
    deno, basic, dt = vnlb_b200.denoise(vid, args.sigma, gpuid=args.gpuid, flows=flows, schedule=args.schedule,
                                        verbose=False)
    t, c, h, w = vid.shape
    print("denoised %dx%dx%d in %.3f s (%.2f Mpx/s)" % (w, h, t, dt, t * h * w / 1e6 / dt))
    if clean is not None:
        for name, x in (("noisy", vid), ("basic", basic), ("deno", deno)):
            print("PSNR %-5s %.3f dB" % (name, float(vnlb_b200.compute_psnrs(x, clean).mean())))
    video_io.save_video_sequence(deno.cpu().numpy(), os.path.join(args.out, "deno"))
    video_io.save_video_sequence(basic.cpu().numpy(), os.path.join(args.out, "basic"))


if __name__ == "__main__":
    main()
