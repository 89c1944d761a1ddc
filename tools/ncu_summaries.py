"""Turn ncu captures into the compact summaries committed under profiles/.

    python tools/ncu_summaries.py launches gpurun_out/launches.csv profiles/rXX_ncu_launch_list_summary.csv "<command line note>"
    python tools/ncu_summaries.py full gpurun_out/prof.ncu-rep profiles/rXX_ncu_full_summary.csv "<note>" [--traffic profiles/traffic.json]

`launches`: CSV written by `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ...`; per-kernel launch count,
total time and share of the forward (weight-packing kernels are dropped).
`full`: raw page of an `ncu --set full` report reduced to the columns the roofline discussion uses; with --traffic the DRAM bytes per
launch of the known kernel categories are written to the JSON that bench.py reads for `roofline.traffic`.
"""
import csv
import io
import json
import re
import subprocess
import sys

PACKING = ("pack_", "pos_table", "pooled_bias", "_image_kernel", "casa_bfrag", "fill_", "tap_kernel", "at::", "distribution_elementwise")
COLS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg",
        # tcgen05 side (only in --set full reports): cycles the tensor-core pipe is occupied by tcgen05.mma (operand fetch included; small-N
        # instructions keep it busy without filling the math units) and its shared-memory operand reads
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum"]
CATEGORY = {"ffn_tail_kernel": "ffn_tail", "qkv_casa_mma_kernel": "qkv_build", "qkv_casa_kernel": "qkv_build", "scc_dense_kernel<16>": "scc_w4",
            "scc_dense_kernel<64>": "scc_w8"}


def short(name):
    m = re.search(r"([A-Za-z_0-9]+(?:<[^>]*>)?)\(", name)
    return m.group(1) if m else name


def launches(src, dst, note):
    rows = [r for r in csv.reader(x for x in open(src) if not x.startswith("==")) if r]
    hdr = rows[0]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = {}
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        k = short(r[ik])
        if any(p in k for p in PACKING):
            continue
        unit = r[hdr.index("Metric Unit")]
        v = float(r[iv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        n, t = agg.get(k, (0, 0.0))
        agg[k] = (n + 1, t + v)
    total = sum(t for _, t in agg.values())
    with open(dst, "w") as f:
        f.write("# " + note + "\n")
        f.write("kernel,launches,total_ms,share_pct\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('"%s",%d,%.3f,%.2f\n' % (k, n, t, 100.0 * t / total))
        f.write("TOTAL,%d,%.3f,100.00\n" % (sum(n for n, _ in agg.values()), total))
    print("wrote", dst, "total %.3f ms" % total)


def full(src, dst, note, traffic=None):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(c) for c in COLS if c in hdr]
    ik = hdr.index("Kernel Name")
    with open(dst, "w") as f:
        f.write("# " + note + "\n")
        w = csv.writer(f)
        w.writerow(["Kernel Name"] + [hdr[i] for i in idx])
        w.writerow([""] + [units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[ik]] + [r[i] for i in idx])
    print("wrote", dst, len(rows) - 2, "launches")
    if traffic:
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        try:
            tr = json.load(open(traffic))
        except OSError:
            tr = {}
        for r in rows[2:]:
            cat = CATEGORY.get(short(r[ik]))
            if cat is None:
                continue
            b = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
            tr[cat] = {"bytes_per_launch": int(round(b, -3)),
                       "note": "dram__bytes_read.sum + dram__bytes_write.sum, one launch at cfg2 shapes (32x256x256), " + dst}
        json.dump(tr, open(traffic, "w"), indent=1)
        print("wrote", traffic)


if __name__ == "__main__":
    mode, src, dst, note = sys.argv[1:5]
    if mode == "launches":
        launches(src, dst, note)
    else:
        full(src, dst, note, sys.argv[6] if len(sys.argv) > 6 and sys.argv[5] == "--traffic" else None)
