cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/final
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/final/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/final/smoke.log
timeout 400 python bench.py > gpurun_out/final/bench_c2.json 2> gpurun_out/final/bench_c2.err; echo "bench c2 rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final/bench_ref.json 2> gpurun_out/final/bench_ref.err; echo "bench ref rc=$?"
timeout 300 python bench.py --workload c1 --no-cpu > gpurun_out/final/bench_c1.json 2> gpurun_out/final/bench_c1.err; echo "bench c1 rc=$?"
for m in 1 2 3; do timeout 300 python bench.py --model $m --steps 20 --warmup 3 --no-cpu > gpurun_out/final/bench_c2_m$m.json 2> gpurun_out/final/bench_c2_m$m.err; echo "bench m$m rc=$?"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/final/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_weights_m4|k_column_reduce|k_locus_acc|k_converge" -s 16 -c 4 -o gpurun_out/final/prof_r1g -f python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/final/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/final/prof_r1g.ncu-rep --page raw --csv > gpurun_out/final/raw_r1g.csv 2>/dev/null
for k in k_weights_m4 k_column_reduce k_locus_acc; do ncu -i gpurun_out/final/prof_r1g.ncu-rep --page source --csv --kernel-name regex:$k > gpurun_out/final/src_$k.csv 2>/dev/null; done
python - <<PY
import json
for f in ("bench_c2","bench_c1","bench_c2_m1","bench_c2_m2","bench_c2_m3","bench_ref"):
    try:
        j=json.load(open(f"gpurun_out/final/{f}.json")); print(f, j.get("ms_per_step"), j.get("value"), (j.get("roofline") or {}).get("per_kernel_ms"), (j.get("e2e") or {}).get("value"))
    except Exception as e: print(f, "ERR", e)
PY
