cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 --durations=8 > gpurun_out/t_quick.log 2>&1; echo "pytest rc=$?"
tail -22 gpurun_out/t_quick.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
cut -c1-200 gpurun_out/bench_quick.json; tail -5 gpurun_out/bench_quick.err
