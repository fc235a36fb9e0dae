cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 -x > gpurun_out/t_quick.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/t_quick.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
cat gpurun_out/bench_quick.json | cut -c1-200; tail -5 gpurun_out/bench_quick.err
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload c3 > gpurun_out/bench_c3_quick.json 2> gpurun_out/bench_c3_quick.err; echo "bench c3 rc=$?"
cat gpurun_out/bench_c3_quick.json | cut -c1-200
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload c5 > gpurun_out/bench_c5_quick.json 2> gpurun_out/bench_c5_quick.err; echo "bench c5 rc=$?"
cat gpurun_out/bench_c5_quick.json | cut -c1-200
