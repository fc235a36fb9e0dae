# One-shot validation of the warp-per-split subflattening scorer: A/B against the block-wide kernel, then the parity tests
# that touch subflattening scores, smoke() and one c3 bench line, all with SPB_SUBFLATTEN_WARP=1.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python scripts/ab_subflatten_warp.py > gpurun_out/ab_warp.log 2>&1; echo "ab rc=$?"; tail -3 gpurun_out/ab_warp.log
export SPB_SUBFLATTEN_WARP=1
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 -k "subflatten or config3 or erickson_svd or identity or four_pattern or random_tables" > gpurun_out/t_warp.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t_warp.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke_warp.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_warp.log
timeout 300 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_warp.json 2> gpurun_out/bench_c3_warp.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_c3_warp.json
