cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
W=${1:-c2}
timeout 900 python bench.py --steps 5 --warmup 3 --workload $W ${2:-} > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; echo "bench rc=$?"
cut -c1-700 gpurun_out/bench_$W.json; tail -3 gpurun_out/bench_$W.err
