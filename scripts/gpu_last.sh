cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"; cut -c1-160 gpurun_out/bench_c2.json
timeout 600 python bench.py --workload c4 --sites 100000000 --max-splits 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c4_full.json 2> gpurun_out/bench_c4_full.err; echo "bench c4 full rc=$?"; cut -c1-200 gpurun_out/bench_c4_full.json; tail -2 gpurun_out/bench_c4_full.err
