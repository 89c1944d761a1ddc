mkdir -p gpurun_out
PKG=single-image-super-resolution-application_b200
echo "bitcmp rev vs fwd: $(timeout 200 python tools/bitcmp.py $PKG/libhitsir_b200.so $PKG/libhitsir_sccfwd.so 2>&1 | tail -1)"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_banded.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2z_pytest.log
for rep in 1 2; do for n in b200 sccfwd; do
  HITSIR_B200_LIB=$PWD/$PKG/libhitsir_$n.so timeout 120 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ab_$n.log 2> gpurun_out/ab_$n.err || { echo "$n FAILED"; continue; }
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads(open(f"gpurun_out/ab_{n}.log").read().strip().split("\n")[-1]); b = d["breakdown"]
print(f"{n:8s} {d['ms_per_step']:.2f} ms clk {d['clocks']['sm_mhz']} | " + ", ".join(f"{k} {b[k]['ms_per_step']:.3f}" for k in ("scc_w16","scc_w32","scc_w48","scc_w64","ffn_tail")))
PY
done; done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct
for n in b200 sccfwd; do
  HITSIR_B200_LIB=$PWD/$PKG/libhitsir_$n.so timeout 300 ncu --metrics $M --clock-control none -k "regex:scc_umma_kernel" --launch-skip 4 -c 4 --csv --log-file gpurun_out/r2z_scc_$n.csv python tools/prof_forward.py --batch 32 --iters 1 > /dev/null 2>&1; echo "ncu $n rc=$?"
done
