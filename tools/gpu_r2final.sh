set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -6 gpurun_out/r2h_pytest.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r2h_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2h_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2h_bench.log 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"
python tools/bench_summary.py gpurun_out/r2h_bench.log | head -12
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2h_bench_ref.log 2> gpurun_out/r2h_bench_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/r2h_bench_ref.log
for w in cfg1 cfg5; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2h_bench_$w.log 2>&1; python tools/bench_summary.py gpurun_out/r2h_bench_$w.log | head -3; done
