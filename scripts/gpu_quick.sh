cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 -x > gpurun_out/t_quick.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/t_quick.log
for w in c2 c3 c5; do timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload $w > gpurun_out/bench_${w}_quick.json 2> gpurun_out/bench_${w}_quick.err; echo "bench $w rc=$?"; cut -c1-200 gpurun_out/bench_${w}_quick.json; done
timeout 300 python scripts/bench_c1.py | cut -c1-200
