# round-2 profile capture (one GPU): launch list of one cfg2 forward, --set full of one ffn_tail launch (the dominant kernel), and the
# roofline metrics of every hot-kernel launch of one forward.  Each ncu pass runs only after the plain command exited 0.  Reports stay
# well below gpurun's 64 MiB return limit.
set -x
mkdir -p gpurun_out
M=launch__grid_size,launch__block_size,launch__registers_per_thread,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,smsp__cycles_active.avg,lts__t_bytes.sum
timeout 300 python tools/prof_forward.py --batch 32 --iters 2 > gpurun_out/r2_prof_plain.log 2>&1 || { tail -5 gpurun_out/r2_prof_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python tools/prof_forward.py --batch 32 --iters 1 > gpurun_out/r2_ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:ffn_tail_kernel" --launch-skip 2 -c 1 -o gpurun_out/r2_ffn_tail_full -f python tools/prof_forward.py --batch 32 --iters 1 > gpurun_out/r2_ncu_ffn_full.log 2>&1; echo "ffn full rc=$?"
timeout 1500 ncu --metrics $M --clock-control none -k "regex:ffn_tail_kernel|umma_gemm_tma_kernel|qkv_casa_mma_kernel|scc_umma_kernel|scc_dense_kernel|conv3_c64_kernel|sca_stats_kernel|fusion_combine_kernel" --launch-skip 2 -c 40 -o gpurun_out/r2_hot -f python tools/prof_forward.py --batch 32 --iters 1 > gpurun_out/r2_ncu_hot.log 2>&1; echo "hot rc=$?"
ls -la gpurun_out/*.ncu-rep
