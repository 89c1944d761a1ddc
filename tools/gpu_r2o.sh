set -x
mkdir -p gpurun_out
T=${1:-r2o}
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=" gpurun_out/${T}_pytest.log | head -20
timeout 400 python tools/measure_taps.py > gpurun_out/${T}_taps.json 2> gpurun_out/${T}_taps.err; echo "taps rc=$?"
timeout 300 python tools/time_finalize.py > gpurun_out/${T}_finalize.json 2> gpurun_out/${T}_finalize.err; cat gpurun_out/${T}_finalize.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python tools/bench_summary.py gpurun_out/${T}_bench.log | head -12
