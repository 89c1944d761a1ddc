"""Benchmark of the HiT-SIR-pro forward pass (BASELINE.json metric: output megapixels/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg4|cfg5|cfg1]

One step = one forward of one batch of synthetic LR images through the public nn.Module (-> C ABI -> CUDA).
N=1 workload: BASELINE.json configs[1] (32 x 3 x 256 x 256 LR, x4 'nearest+conv', pro config).  For N>1 the
launcher is torchrun (one rank per GPU); every rank runs its own batch (weak scaling) and the SR outputs are
all-gathered over NCCL once per step.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (flags, upsampler, upscale, batch, H, W, description)
    "cfg1": ((1, 1, 1), "nearest+conv", 4, 1, 64, 64, "HiT-SIR-pro x4, 1x 64x64 LR"),
    "cfg2": ((1, 1, 1), "nearest+conv", 4, 32, 256, 256, "HiT-SIR-pro x4 batched inference, 32x 256x256 LR"),
    "cfg3": ((1, 1, 1), "pixelshuffle", 2, 8, 576, 576, "HiT-SIR-pro x2, the 8 halo tiles (576x576) of a 1920x1080 frame as one batch (tile t -> rank t % N)"),
    "cfg4": ((1, 1, 1), "nearest+conv", 4, 16, 512, 512, "hitsir_pro_gan generator x4, 16x 512x512 LR"),
    "cfg5": ((0, 0, 0), "nearest+conv", 4, 32, 256, 256, "ablation HiT-SIR-pro x4 casa=False mulsizeconvextract=False, 32x 256x256 LR"),
}
# algorithmic GFLOP per image of the reference formulation (conv+GEMM+bmm, FMA=2; torch.utils.flop_counter on the
# reference module)
GFLOP_PER_IMAGE = {"cfg1": 101.81, "cfg2": 1575.8, "cfg3": 7110.2, "cfg4": 6260.6, "cfg5": 1417.1}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(bf16=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), hbm=p["hbm_gbs"], src="measured")
    return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# analytic algorithmic FLOPs (reference formulation, FMA=2) of one launch of each contraction category
def category_flops(cat, B, H, W):
    N = B * H * W
    table = {
        "gemm_fc1_gelu": 2 * N * 180 * 360, "gemm_fc2_ln": 2 * N * 360 * 180, "gemm_proj_ln": 2 * N * 180 * 180,
        "conv_layer": 2 * N * 9 * 180 * 180, "conv_after_body": 2 * N * 9 * 180 * 180, "conv_ua": 2 * N * 9 * 180 * 180,
        "conv_before_upsample": 2 * N * 9 * 180 * 64, "conv_up1": 2 * 4 * N * 9 * 64 * 64, "conv_up2": 2 * 16 * N * 9 * 64 * 64,
        "conv_hr": 2 * 16 * N * 9 * 64 * 64, "conv_last": 2 * 16 * N * 9 * 64 * 3,
        "gemm_first_msgate": 2 * N * 3 * 180 * (9 + 25 + 49 + 81 + 1), "gemm_first_last": 2 * N * 720 * 180,
        "gemm_first": 2 * N * 27 * 180,
        "ffn_tail": 2 * N * 360 * 180 + 2 * N * 360 * 25,          # fc2 + the depthwise 5x5 (both counted by the reference flop counter)
    }
    if cat in table:
        return table[cat]
    if cat.startswith("scc_w"):
        w = int(cat[5:])
        Hp, Wp = -(-H // w) * w, -(-W // w) * w
        Np = B * Hp * Wp
        Lb = min(w, 8) ** 2
        r2 = (w * w) // Lb
        # k-gen (2 x 15x15 per head), pooling Linear(r^2,1) on k and v, S-SC (q k^T, corr v), C-SC (q^T k, corr v^T)
        return Np * (2 * 2 * 6 * 15 * 15 + 2 * 2 * 90 + 2 * 2 * Lb * 90 + 2 * 2 * 90 * 90)
    return 0


# algorithmic (compulsory, unpadded) HBM bytes of one launch of each bandwidth-bound category: DESIGN.md section 3
def category_bytes(cat, B, H, W):
    N = B * H * W
    table = {
        "dwconv5": N * (360 * 2 + 360 * 2),                       # bf16 hidden in, bf16 hidden out
        "ffn_tail": N * (360 * 2 + 180 * 4 + 180 * 4),            # bf16 hidden in, fp32 residual in, fp32 stream out
        "cast_shadow": N * (180 * 4 + 180 * 2),
        "qkv_build": N * (180 * 4 + 180 * 2),                     # fp32 stream in, bf16 window tokens out
        "sca_stats": N * 180 * 4,                                 # fp32 stream in (statistics out are negligible)
        "gemm_proj_ln": N * (180 * 2 + 180 * 4 + 180 * 4 + 180 * 2),   # bf16 A, fp32 residual in, fp32 stream + bf16 shadow out
        "gemm_fc2_ln": N * (360 * 2 + 180 * 4 + 180 * 4 + 180 * 2),
        "gemm_fc1_gelu": N * (180 * 2 + 360 * 2),
        "ln_rows": N * (180 * 4 + 180 * 4),
        "fusion_combine": N * (5 * 180 * 4 + 180 * 2),
        "upsample2": 0,
    }
    if cat in table:
        return table[cat]
    if cat.startswith("scc_w"):
        return N * (180 * 2 + 180 * 2)                            # window tokens in (once), self-correlation out
    return 0


def measured_traffic(cat):
    """DRAM bytes per launch of `cat` from the committed ncu capture (profiles/traffic.json), or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        return json.load(open(path)).get(cat)
    return None


def run_ours(args):
    import torch.distributed as dist
    import hitsir_b200
    from hitsir_b200.sharding import ShardedSR

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    flags, up, scale, B, H, W, desc = WORKLOADS[args.workload]
    torch.manual_seed(0)
    kw = dict(hitsir_b200.PRO_KWARGS)
    kw.update(upsampler=up, upscale=scale)
    model = hitsir_b200.HiT_SIR(*[bool(f) for f in flags], **kw).eval().to(dev)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    x = torch.rand(B, 3, H, W, device=dev, generator=g)
    sharded = ShardedSR(model, scale)
    out_mp = B * 3 // 3 * (H * scale) * (W * scale) / 1e6          # output megapixels per rank per step
    gathered = torch.empty((world * B, 3, H * scale, W * scale), device=dev) if world > 1 else None
    # Output gather: peer-to-peer copies over NVLink on a side stream (sharding.PeerGather), double-buffered so that the gather of step
    # i overlaps the forward of step i + 1; every gather is complete before the timed region ends.  HITSIR_GATHER=nccl (or a platform
    # without symmetric memory) uses one NCCL all-gather per step on the compute stream instead.
    peer = None
    if world > 1 and os.environ.get("HITSIR_GATHER", "p2p") == "p2p":
        try:
            from hitsir_b200.sharding import PeerGather
            peer = PeerGather((B, 3, H * scale, W * scale), torch.float32, dev)
        except Exception as e:                                  # plumbing fallback only: the compute path is unchanged
            if rank == 0:
                print(f"[bench] PeerGather unavailable ({type(e).__name__}: {e}); using NCCL all_gather", file=sys.stderr)
            peer = None
        ok = torch.tensor([1 if peer is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)               # all ranks take the same path
        if ok.item() == 0:
            peer = None
    state = {"i": 0}

    def step():
        y = model(x)
        if world > 1:
            if peer is not None:
                peer.start(y, state["i"] & 1)                   # waits (on its own stream) for this forward only
                state["i"] += 1
            else:
                dist.all_gather_into_tensor(gathered, y)
        return y

    def drain():
        if peer is not None:
            peer.wait(0)

    with torch.no_grad():
        for _ in range(args.warmup):
            step()
        drain()
        if peer is not None:                                    # untimed check: the peer-to-peer gather delivers what NCCL's all_gather delivers
            yv = model(x)
            peer.start(yv, 0)
            got = peer.wait(0)
            dist.all_gather_into_tensor(gathered, yv)
            if not torch.equal(got.view_as(gathered), gathered):
                raise RuntimeError("PeerGather result differs from NCCL all_gather")
        torch.cuda.synchronize()
        launches_per_step = model.last_launch_count
        model.profile_enable(dev, True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        drain()                                                 # the last gathers have landed before the clock stops
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        clocks = sampler.stop()
        ms_total = e0.elapsed_time(e1)
        prof = model.profile_read(dev)
        model.profile_enable(dev, False)
        # ---- end-to-end through the public host-buffer API: every step copies its inputs from pinned host memory and its result
        # back (HostPipeline: double-buffered, so the PCIe copies of step i overlap the kernels of its neighbours)
        xh = [x.cpu().pin_memory() for _ in range(2)]
        yh = [torch.empty((B, 3, H * scale, W * scale), dtype=torch.float32).pin_memory() for _ in range(2)]
        model.forward_host(xh[0], yh[0])
        pipe = hitsir_b200.HostPipeline(model, dev)
        pipe.submit(xh[0], yh[0])
        pipe.wait()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        n_e2e = max(2, min(args.steps, 6))
        for i in range(n_e2e):
            pipe.submit(xh[i & 1], yh[i & 1])
        pipe.wait()
        e2e_s = (time.perf_counter() - t0) / n_e2e
    t = torch.tensor([ms_total, e2e_s * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t[0].item() / args.steps
    e2e_ms = t[1].item()
    value = world * out_mp / (ms_step / 1e3)
    pk = peaks()
    # dominant kernel category of the step
    cat, (cat_ms, cat_n) = max(prof.items(), key=lambda kv: kv[1][0])
    fl = category_flops(cat, B, H, W)
    by = category_bytes(cat, B, H, W)
    N_tok = B * H * W
    t_launch = cat_ms / cat_n / 1e3
    tflops = fl / t_launch / 1e12 if fl else None
    gbs = by / t_launch / 1e9 if by else None
    # the roofline that binds the dominant kernel: whichever of its two fractions is larger
    f_t = tflops / pk["bf16_sustained"] if tflops else 0.0
    f_h = gbs / pk["hbm"] if gbs else 0.0
    if f_h >= f_t and gbs:
        roof = {"bound": "hbm", "kernel": cat, "achieved": round(gbs, 1), "peak": pk["hbm"], "unit": "GB/s", "frac": round(f_h, 4),
                "algorithmic_bytes_per_launch": by}
    else:
        roof = {"bound": "tensor", "kernel": cat, "achieved": round(tflops, 2) if tflops else None, "peak": pk["bf16_sustained"],
                "unit": "TFLOP/s", "frac": round(f_t, 4) if tflops else None, "algorithmic_flops_per_launch": fl}
    if cat == "ffn_tail":
        # the depthwise 5x5 of this kernel runs as warp-level m16n8k16 MMAs with block-diagonal taps (13 per 16 px x 8 ch output row):
        # 8x16-pixel tiles x 6 slices x 16 warps x 52.  Peak = one MMA per 7.34 cycles per SM sub-partition, measured with
        # tools/ubench/mma_rate.cu on this pool's B200 (4 sub-partitions x 148 SMs at the sampled SM clock).
        tiles = B * -(-H // 8) * -(-W // 16)
        mma = tiles * 6 * 16 * 52 / t_launch
        mma_peak = 148 * 4 * (clocks.get("sm_mhz") or 1965.0) * 1e6 / 7.34
        roof["warp_mma_m16n8k16"] = {"achieved_gmma_s": round(mma / 1e9, 2), "peak_gmma_s": round(mma_peak / 1e9, 2), "frac": round(mma / mma_peak, 4),
                                     "useful_flop_fraction": round(25 / (13 * 16), 3)}
    elif cat == "dwconv5":
        # stand-alone SIMT depthwise path (HITSIR_FFN=unfused): 25 FP32 FMA per hidden element, 148 SMs x 128 FMA/clk at the sampled SM clock
        fma = N_tok * 360 * 25 / t_launch
        fma_peak = 148 * 128 * (clocks.get("sm_mhz") or 1965.0) * 1e6
        roof["simt_fp32_fma"] = {"achieved_tfma_s": round(fma / 1e12, 2), "peak_tfma_s": round(fma_peak / 1e12, 2), "frac": round(fma / fma_peak, 4)}
    tr = measured_traffic(cat)
    roof["traffic"] = tr.get("bytes_per_launch") if isinstance(tr, dict) else tr
    if isinstance(tr, dict):
        roof["traffic_note"] = tr.get("note")
    whole = GFLOP_PER_IMAGE.get(args.workload, 0) * B / (ms_step / 1e3) / 1e3
    roof.update({"peak_source": pk["src"] + (" (sustained, kernel timed inside a long step)" if roof["bound"] == "tensor" else " (copy bandwidth)"),
                 "kernel_share_of_step": round(cat_ms / ms_total, 3), "launch_ms": round(t_launch * 1e3, 4),
                 "whole_forward_tflops": round(whole, 1), "whole_forward_frac": round(whole / pk["bf16_sustained"], 4)})
    breakdown = {k: {"ms_per_step": round(v[0] / args.steps, 3), "launches_per_step": v[1] // args.steps,
                     "tflops": round(category_flops(k, B, H, W) / (v[0] / v[1] / 1e3) / 1e12, 1) if category_flops(k, B, H, W) else None,
                     "gbs": round(category_bytes(k, B, H, W) / (v[0] / v[1] / 1e3) / 1e9, 0) if category_bytes(k, B, H, W) else None}
                 for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}
    line = {
        "metric": "HiT-SIR-pro x4 output megapixels/s", "value": round(value, 3), "unit": "MP/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}, random-init weights, per-GPU batch {B}",
                   "l2": "activations (>= 0.8 GB per tensor) exceed the 126 MB L2; no explicit flush",
                   "collective": ("none" if world == 1 else "all-gather of the fp32 SR outputs per step: NVLink peer-to-peer copies on a side stream, overlapped with the next forward"
                                  if peer is not None else "NCCL all_gather of the fp32 SR outputs per step")},
        "e2e": {"value": round(world * out_mp / (e2e_ms / 1e3), 3), "unit": "MP/s", "h2d_bytes_per_step": B * 3 * H * W * 4,
                "d2h_bytes_per_step": B * 3 * H * W * scale * scale * 4, "ms_per_step": round(e2e_ms, 3)},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks,
        "roofline": roof,
        "breakdown": breakdown,
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(budget_s=20.0)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(budget_s=20.0, hw=(64, 64), steps=None):
    """The oracle (CPU restatement of the reference, fp32, all host threads) timed on cfg1-sized patches.
    `rebuild_bias=True` reproduces the reference's per-forward rebuild of the relative-position bias
    (hit_sir_pro.py:477-503) so the number reflects the reference's own CPU cost."""
    from oracle.hitsir_oracle import HiTSIROracle, OracleConfig
    from oracle.weights import fill_state_dict
    import hitsir_b200
    torch.manual_seed(0)
    m = hitsir_b200.HiT_SIR(True, True, True, **hitsir_b200.PRO_KWARGS)
    o = HiTSIROracle(fill_state_dict(m.state_dict(), 0, "init"), OracleConfig())
    x = torch.rand(1, 3, hw[0], hw[1])
    times = []
    with torch.no_grad():
        o(x)                                   # warm-up
        t_start = time.perf_counter()
        while True:
            o._bias.clear()                    # reference semantics: bias tables rebuilt every forward
            t0 = time.perf_counter()
            o(x)
            times.append(time.perf_counter() - t0)
            if (steps is not None and len(times) >= steps) or (steps is None and time.perf_counter() - t_start > budget_s and len(times) >= 2):
                break
    times.sort()
    med = times[len(times) // 2]
    mp = hw[0] * 4 * hw[1] * 4 / 1e6
    return {"value": round(mp / med, 5), "unit": "MP/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{len(times)} forwards of 1x3x{hw[0]}x{hw[1]} LR (BASELINE cfg1) through oracle/hitsir_oracle.py, fp32, "
                      f"median {med:.2f} s, relative-position bias rebuilt per forward like the reference",
            "seconds_per_forward": round(med, 3)}


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path.  The reference is a Python module that
    cannot travel to the GPU box, so this times the oracle port (validated against the reference's golden
    vectors) on the box's host cores, on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # all the host threads this process may use: torchrun exports OMP_NUM_THREADS=1, which would make the reference arm single-threaded
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except (AttributeError, OSError):
        torch.set_num_threads(max(1, os.cpu_count() or 1))
    flags, up, scale, B, H, W, desc = WORKLOADS[args.workload]
    total = args.steps + args.warmup
    hw = (H, W) if total <= 3 else ((128, 128) if total <= 12 else (64, 64))
    hw = (min(hw[0], H), min(hw[1], W))
    from oracle.hitsir_oracle import HiTSIROracle, OracleConfig
    from oracle.weights import fill_state_dict
    import hitsir_b200
    kw = dict(hitsir_b200.PRO_KWARGS)
    kw.update(upsampler=up, upscale=scale)
    m = hitsir_b200.HiT_SIR(*[bool(f) for f in flags], **kw)
    o = HiTSIROracle(fill_state_dict(m.state_dict(), 0, "init"), OracleConfig(*[bool(f) for f in flags], upscale=scale, upsampler=up))
    x = torch.rand(1, 3, hw[0], hw[1])
    with torch.no_grad():
        for _ in range(args.warmup):
            o._bias.clear()
            o(x)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            o._bias.clear()
            o(x)
        dt = (time.perf_counter() - t0) / args.steps
    mp = hw[0] * scale * hw[1] * scale / 1e6
    v = round(mp / dt, 5)
    sample = (f"each step = 1 image crop {hw[0]}x{hw[1]} LR of the {B}x{H}x{W} batch, oracle port of the reference forward "
              f"(fp32, CPU, relative-position bias rebuilt per forward like the reference)")
    print(json.dumps({
        "impl": "reference", "metric": "HiT-SIR-pro x4 output megapixels/s", "value": v, "unit": "MP/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "sample": sample},
        "cpu_baseline": {"value": v, "unit": "MP/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
