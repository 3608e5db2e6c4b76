cd $GRAFT_REPO_ROOT
for di in 8 100000 1; do
  GBRS_DEEP_ITEMS=$di timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_di$di.json 2> gpurun_out/bench_di$di.err; echo "bench di=$di rc=$?"
  python - <<PY
import json
j=json.load(open("gpurun_out/bench_di$di.json")); print("di$di", j["ms_per_step"], j["roofline"]["per_kernel_ms"])
PY
done
