cd $GRAFT_REPO_ROOT
for v in p2p nvls; do
GBRS_XCHG=$v timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_8gpu_$v.json 2> gpurun_out/bench_8gpu_$v.err; echo "bench8 $v rc=$?"
python - <<PY
import json
j=json.load(open("gpurun_out/bench_8gpu_$v.json")); print("8gpu $v", j["ms_per_step"], j["value"], j["roofline"]["per_kernel_ms"], j["config"]["exchange"][:60])
PY
done
