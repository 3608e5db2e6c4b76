cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_e2e2.json 2> gpurun_out/bench_e2e2.err; echo "bench rc=$?"
python - <<PY
import json
j=json.load(open("gpurun_out/bench_e2e2.json")); print(j["ms_per_step"], j["roofline"]["per_kernel_ms"]); print(j["e2e"])
PY
