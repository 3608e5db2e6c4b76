cd $GRAFT_REPO_ROOT
for v in m0_o0_b4 m0_o1_b4 m2_o0_b4 m2_o1_b4 m1_o0_b4; do
  GBRS_LIB_PATH=$GRAFT_REPO_ROOT/gbrs_b200/_C/libv_$v.so timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; echo "bench $v rc=$?"
  python - <<PY
import json
j=json.load(open("gpurun_out/bench_$v.json")); print("$v", j["ms_per_step"], j["roofline"]["per_kernel_ms"])
PY
done
