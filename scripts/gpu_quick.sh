cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 -x > gpurun_out/t_quick.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/t_quick.log
timeout 1200 python bench.py --steps 2 --warmup 3 --workload c4 --sites 100000000 --max-splits 64 > gpurun_out/bench_c4_full.json 2> gpurun_out/bench_c4_full.err; echo "bench c4 rc=$?"
cat gpurun_out/bench_c4_full.json; tail -5 gpurun_out/bench_c4_full.err
