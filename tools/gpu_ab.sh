# usage: bash tools/gpu_ab.sh base v1 v2 ...  -- bitwise comparison of every variant against the first, then the A/B bench
mkdir -p gpurun_out
PKG=single-image-super-resolution-application_b200
b=$1; shift
for n in "$@"; do
  echo "bitcmp $b vs $n: $(timeout 200 python tools/bitcmp.py $PKG/libhitsir_$b.so $PKG/libhitsir_$n.so 2>&1 | tail -1)"
done
bash tools/ab_run.sh $b "$@"
