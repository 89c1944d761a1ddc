"""Stage-by-stage check of the tcgen05 window self-correlation kernel (csrc/scc_umma.cu).  GPU box only.

    python tools/scc_check.py [--hw 56 72] [--batch 1] [--mode stress] [--blocks 0 1 2 3 4 5]

For every block j of layer 0 it reads the kernel's window-0 intermediates through the `block0.j.sccdbg` tap
(G | TPT | corr | KP | Mblk, fp32) and compares them with the same quantities computed here in fp32 from the
kernel's own bf16 input tokens (`block0.j.qkv` tap), then compares the kernel output (`block0.j.scc`) with the
CPU oracle.  The oracle import makes this test infrastructure, not product code.
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.weights import synthetic_image  # noqa: E402
from tests.helpers import rel_l2  # noqa: E402


def pos(c):
    half, r = divmod(c, 90)
    return half * 96 + (r // 15) * 16 + r % 15


def tile_geom(w):
    if w == 4:
        return 4, 4, 1
    if w == 8:
        return 8, 8, 1
    return 16, 8, w // 16


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hw", type=int, nargs=2, default=[56, 72])
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--mode", default="stress")
    ap.add_argument("--wins", type=int, nargs="*", default=[4, 8, 16, 32, 48, 64])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--one", type=int, default=0, help="(internal) run a single window size in this process")
    args = ap.parse_args()
    if not args.one:
        # one process per window size: a trapped kernel poisons the CUDA context of its process only
        import subprocess
        for w in args.wins:
            cmd = [sys.executable, os.path.abspath(__file__), "--hw", str(args.hw[0]), str(args.hw[1]), "--batch", str(args.batch),
                   "--mode", args.mode, "--seed", str(args.seed), "--one", str(w)]
            try:
                r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
                print(r.stdout[-6000:], flush=True)
                print(f"[w={w}] exit code {r.returncode}", flush=True)
            except subprocess.TimeoutExpired as e:
                print(f"[w={w}] TIMEOUT\n{(e.stdout or b'')[-3000:]}", flush=True)
        return
    import hitsir_b200
    from oracle.hitsir_oracle import HiTSIROracle, OracleConfig
    from oracle.weights import fill_state_dict
    # a one-block model whose only window is args.one
    kw = dict(hitsir_b200.PRO_KWARGS)
    kw.update(depths=[1], num_heads=[6], hier_win_ratios=[args.one / 8.0])
    model = hitsir_b200.HiT_SIR(True, True, True, **kw).eval()
    sd0 = fill_state_dict(model.state_dict(), args.seed, args.mode)
    model.load_state_dict(sd0, strict=True)
    oracle = HiTSIROracle(sd0, OracleConfig(True, True, True, depths=(1,), num_heads=(6,), hier_win_ratios=(args.one / 8.0,)))
    sd = oracle.sd
    x = synthetic_image(args.batch, args.hw[0], args.hw[1], seed=77)
    taps = {}
    with torch.no_grad():
        oracle.forward(x, taps)
    dev = torch.device("cuda:0")
    model = model.to(dev)
    xd = x.to(dev)
    B, H, W = args.batch, args.hw[0], args.hw[1]
    wins = [args.one]
    perm = torch.tensor([pos(c) for c in range(180)])

    def tap(name, numel):
        dst = torch.full((numel,), float("nan"), device=dev)
        model.set_tap(dev, name, dst, stop=True)
        with torch.no_grad():
            model(xd)
        torch.cuda.synchronize()
        return dst.cpu()

    for j in [0]:
        w = wins[j]
        base = min(w, 8)
        r = w // base
        L, Lb = w * w, base * base
        Hp, Wp = -(-H // w) * w, -(-W // w) * w
        p = f"layers.0.residual_group.blocks.{j}.correlation."
        qkv = tap(f"block0.{j}.qkv", B * Hp * Wp * 180).view(B, Hp, Wp, 180)
        print(f"block0.{j} w={w} Hp={Hp} Wp={Wp}: qkv rel_l2 vs oracle = {rel_l2(qkv, taps[f'block0.{j}.qkv']):.3e}", flush=True)
        dbg = tap(f"block0.{j}.sccdbg", 63488 + 32)
        tl = dbg[63488:63488 + 16]
        print("    timeline (cycles since MMA-thread window start): " + " ".join(f"{int((v - tl[0]) % (1 << 24))}" for v in tl.tolist()), flush=True)
        G = dbg[:24576].view(128, 192)
        TPT = dbg[24576:36864].view(192, 64)
        corr = dbg[36864:49152].view(128, 96)
        KP = dbg[49152:61440].view(128, 96)
        Mb = dbg[61440:63488].view(128, 16)
        # expected, from the kernel's own input tokens of window 0
        tw = qkv[0, :w, :w, :].reshape(L, 180)
        T = torch.zeros(L, 192)
        T[:, perm] = tw
        for h in range(6):
            T[:, 16 * h + 15] = 1.0
        G_ref = T[:, :128].t() @ T
        wsl = sd[p + "spatial_linear.weight"].view(r, r)
        bsl = sd[p + "spatial_linear.bias"].item()
        P = torch.zeros(64, L)
        for l in range(L):
            ly, lx = divmod(l, w)
            P[(ly // r) * base + lx // r, l] = wsl[ly % r, lx % r]
        TPT_raw = T.t() @ P.t()                       # [192, 64]
        padrow = torch.tensor([(c % 16) == 15 for c in range(96)])
        TPT_ref = TPT_raw.clone()
        vpart = TPT_ref[96:, :]
        vpart[:, :Lb] += bsl
        vpart[padrow, :] = 0
        vpart[:, Lb:] = 0
        Wk = torch.zeros(96, 192)
        w1, b1 = sd[p + "k_generate1.weight"], sd[p + "k_generate1.bias"]
        w2, b2 = sd[p + "k_generate2.weight"], sd[p + "k_generate2.bias"]
        for h in range(6):
            Wk[16 * h:16 * h + 15, 16 * h:16 * h + 15] = 0.5 * w1
            Wk[16 * h:16 * h + 15, 96 + 16 * h:96 + 16 * h + 15] = 0.5 * w2
            Wk[16 * h:16 * h + 15, 16 * h + 15] += 0.5 * (b1 + b2)
        corr_ref = (G_ref[:96] @ Wk.t()) / L
        KP_ref = TPT_raw.t() @ Wk.t()                 # [64, 96]
        KP_ref[:Lb, ~padrow] += bsl
        KP_ref[Lb:, :] = 0
        KP_ref[:, padrow] = 0
        VP = TPT_ref[96:, :].t()                      # [64, 96]
        Mfull = VP.t() @ KP_ref / 15.0                # [(h,j)][(h',i)]
        Mb_ref = torch.stack([Mfull[rw, 16 * (rw // 16):16 * (rw // 16) + 16] for rw in range(96)])

        def rep(name, got, ref):
            nan = int(torch.isnan(got).sum())
            print(f"    {name:6s} rel_l2={rel_l2(torch.nan_to_num(got), ref):.3e} max_abs={(torch.nan_to_num(got) - ref).abs().max().item():.3e} "
                  f"ref_absmax={ref.abs().max().item():.3e} nan={nan}", flush=True)
        rep("G", G[:96], G_ref[:96])
        rep("TPT", TPT, TPT_ref)
        rep("corr", corr[:96], corr_ref)
        rep("KP", KP[:64], KP_ref)
        rep("Mblk", Mb[:96], Mb_ref)
        ref = taps[f"block0.{j}.scc"].contiguous()
        got = tap(f"block0.{j}.scc", ref.numel()).view(ref.shape)
        rep("out_s", got[..., :90], ref[..., :90])
        rep("out_c", got[..., 90:], ref[..., 90:])
    model.set_tap(dev, None)


if __name__ == "__main__":
    main()
