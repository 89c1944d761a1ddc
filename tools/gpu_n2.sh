mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "rc=$?"
tail -c 1800 gpurun_out/bench_n$N.log; tail -5 gpurun_out/bench_n$N.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_n{n}.log").read().strip().split("\n")[-1])
    print("N", d["n_gpus"], "value", d["value"], "ms", d["ms_per_step"], "collective:", d["config"]["collective"][:60])
    print("cfg4", d.get("cfg4")); print("cfg3", d.get("cfg3")); print("cfg3_exact", d.get("cfg3_exact"))
except Exception as e:
    print("parse failed", e)
PY
