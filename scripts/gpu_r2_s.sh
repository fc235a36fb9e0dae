# Round-2 run S (one B200): BASELINE config 4 at its full size (64 taxa, 10^8 sites), 64 of the thin splits.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python bench.py --workload c4 --sites 100000000 --max-splits 64 --steps 1 --warmup 3 --extras none --no-cpu-baseline > gpurun_out/r2s_bench_c4_full.json 2> gpurun_out/r2s_bench_c4_full.err; echo "bench c4 full rc=$?"; python scripts/bench_brief.py gpurun_out/r2s_bench_c4_full.json 2>/dev/null; tail -3 gpurun_out/r2s_bench_c4_full.err
