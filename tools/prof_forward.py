"""Tiny driver for ncu captures: N forwards of the pro x4 model on synthetic input (GPU box only).

    python tools/prof_forward.py [--batch 8] [--hw 256 256] [--iters 2]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hitsir_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--hw", type=int, nargs=2, default=[256, 256])
    ap.add_argument("--iters", type=int, default=2)
    args = ap.parse_args()
    torch.manual_seed(0)
    model = hitsir_b200.HiT_SIR(True, True, True, **hitsir_b200.PRO_KWARGS).eval().to("cuda:0")
    x = torch.rand(args.batch, 3, args.hw[0], args.hw[1], device="cuda:0")
    with torch.no_grad():
        for _ in range(args.iters):
            y = model(x)
    torch.cuda.synchronize()
    print("ok", tuple(y.shape), float(y.float().mean()))


if __name__ == "__main__":
    main()
