set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
grep -n "passed\|failed\|FAILED\|Error" gpurun_out/r2b_pytest.log | head -40
timeout 600 python tools/stage_check.py --case ps2 --hw 192 192 --mode init --seed 11 > gpurun_out/r2b_stage_ps2_init.log 2>&1
timeout 600 python tools/stage_check.py --case pro --hw 192 192 --mode init --seed 13 > gpurun_out/r2b_stage_pro_init.log 2>&1
timeout 600 python tools/stage_check.py --case ps2 --hw 192 192 --mode stress --seed 12 > gpurun_out/r2b_stage_ps2_stress.log 2>&1
tail -12 gpurun_out/r2b_stage_ps2_init.log; tail -8 gpurun_out/r2b_stage_pro_init.log; tail -8 gpurun_out/r2b_stage_ps2_stress.log
