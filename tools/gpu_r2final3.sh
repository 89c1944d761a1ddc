# final round-2 capture (one GPU): tests, bench lines, then the ncu evidence of the same build (each ncu pass after the plain command exited 0)
set -x
mkdir -p gpurun_out
T=r2t
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=" gpurun_out/${T}_pytest.log | head -20
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python tools/bench_summary.py gpurun_out/${T}_bench.log | head -8
for w in cfg1 cfg5; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${T}_bench_$w.log 2>&1; python tools/bench_summary.py gpurun_out/${T}_bench_$w.log | head -2; done
M=launch__grid_size,launch__block_size,launch__registers_per_thread,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,smsp__cycles_active.avg,lts__t_bytes.sum,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed
timeout 300 python tools/prof_forward.py --batch 32 --iters 2 > gpurun_out/${T}_prof_plain.log 2>&1 || { tail -5 gpurun_out/${T}_prof_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python tools/prof_forward.py --batch 32 --iters 1 > gpurun_out/${T}_ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:ffn_tail_kernel" --launch-skip 2 -c 1 -o gpurun_out/${T}_ffn_tail_full -f python tools/prof_forward.py --batch 32 --iters 1 > gpurun_out/${T}_ncu_ffn_full.log 2>&1; echo "ffn full rc=$?"
timeout 1500 ncu --metrics $M --clock-control none -k "regex:ffn_tail_kernel|umma_gemm_tma_kernel|qkv_casa_mma_kernel|scc_umma_kernel|scc_dense_kernel|conv3_c64|conv_last_fold|sca_stats_kernel|fusion_combine_kernel" --launch-skip 2 -c 48 -o gpurun_out/${T}_hot -f python tools/prof_forward.py --batch 32 --iters 1 > gpurun_out/${T}_ncu_hot.log 2>&1; echo "hot rc=$?"
ls -la gpurun_out/${T}*.ncu-rep
