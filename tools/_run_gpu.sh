cd $GRAFT_REPO_ROOT
timeout 600 python tools/bench_compress.py > gpurun_out/bench_compress.json 2> gpurun_out/bench_compress.err; echo "compress bench rc=$?"
cat gpurun_out/bench_compress.json; tail -3 gpurun_out/bench_compress.err
