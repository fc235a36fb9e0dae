# Full check on one B200: all GPU tests, smoke(), the default bench line.   gpurun --timeout 2400 -- "bash scripts/gpu_check.sh"
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 -x > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/t_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
cut -c1-200 gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err
