set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
tail -4 gpurun_out/r2m_pytest.log
timeout 400 python tools/measure_taps.py > gpurun_out/r2m_taps.json 2> gpurun_out/r2m_taps.err; echo "taps rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2m_bench.log 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"
python tools/bench_summary.py gpurun_out/r2m_bench.log | head -8
for w in cfg1 cfg5; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2m_bench_$w.log 2>&1; python tools/bench_summary.py gpurun_out/r2m_bench_$w.log | head -2; done
M=launch__grid_size,launch__block_size,launch__registers_per_thread,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,smsp__cycles_active.avg,lts__t_bytes.sum
timeout 600 ncu --metrics $M --clock-control none -k "regex:qkv_casa_mma_kernel|fusion_combine_kernel" --launch-skip 1 -c 4 -o gpurun_out/r2m_qkv -f python tools/prof_forward.py --batch 32 --iters 1 > gpurun_out/r2m_ncu_qkv.log 2>&1; echo "ncu rc=$?"
