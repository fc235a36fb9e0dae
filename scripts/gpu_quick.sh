cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 -x -k "wide or thin" > gpurun_out/t_quick.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/t_quick.log
timeout 900 python bench.py --steps 2 --warmup 3 --workload c4 --no-cpu-baseline > gpurun_out/bench_c4_10m.json 2> gpurun_out/bench_c4_10m.err; echo "bench c4 rc=$?"
cat gpurun_out/bench_c4_10m.json | cut -c1-300; tail -5 gpurun_out/bench_c4_10m.err
