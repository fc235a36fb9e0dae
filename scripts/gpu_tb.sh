# Quick loop on one B200: a subset of the GPU tests (pytest -k "$1") and one bench line (workload "$2", default c2).
#   gpurun --timeout 900 -- 'bash scripts/gpu_tb.sh "int32_gram or config2" c2'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 -k "${1:-int32_gram}" > gpurun_out/t_quick.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/t_quick.log
W=${2:-c2}
timeout 900 python bench.py --steps 5 --warmup 3 --workload $W --no-cpu-baseline > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; echo "bench rc=$?"
cut -c1-200 gpurun_out/bench_$W.json; tail -3 gpurun_out/bench_$W.err
