set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -25 gpurun_out/r2a_pytest.log
timeout 600 python tools/measure_taps.py > gpurun_out/r2a_taps.json 2> gpurun_out/r2a_taps.err; echo "taps rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.log 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2a_bench.log
tail -5 gpurun_out/r2a_bench.err
