mkdir -p gpurun_out
PKG=single-image-super-resolution-application_b200
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2w_pytest.log
echo "bitcmp pdl vs nopdl: $(timeout 200 python tools/bitcmp.py $PKG/libhitsir_b200.so $PKG/libhitsir_nopdl.so 2>&1 | tail -1)"
for rep in 1 2; do
for n in b200 nopdl; do
  for w in cfg1 cfg2; do
    HITSIR_B200_LIB=$PWD/$PKG/libhitsir_$n.so timeout 120 python bench.py --workload $w --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ab_${n}_$w.log 2> gpurun_out/ab_${n}_$w.err || { echo "$n $w FAILED"; tail -3 gpurun_out/ab_${n}_$w.err; continue; }
    echo "$n $w: $(python tools/bench_summary.py gpurun_out/ab_${n}_$w.log | head -1)"
  done
done
done
