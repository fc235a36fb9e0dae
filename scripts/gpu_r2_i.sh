# Round-2 run I (one B200): GPU tests after the DMMA / table-join / 64-split scatter changes, strip A/B, full bench, launch list.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 -x > gpurun_out/r2i_tests.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2i_tests.log
SPB_STRIP_KERNEL=scan timeout 600 python bench.py --steps 3 --warmup 3 --extras none --no-cpu-baseline > gpurun_out/r2i_bench_c2_scan.json 2> gpurun_out/r2i_bench_c2_scan.err; echo "bench(strip=scan) rc=$?"; python scripts/bench_brief.py gpurun_out/r2i_bench_c2_scan.json 2>/dev/null | head -2
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"; python scripts/bench_brief.py gpurun_out/r2i_bench.json 2>/dev/null; tail -3 gpurun_out/r2i_bench.err
SPB_SMALL_EIG=jacobi timeout 600 python bench.py --steps 3 --warmup 3 --extras none --no-cpu-baseline > gpurun_out/r2i_bench_c2_jacobi.json 2> gpurun_out/r2i_bench_c2_jacobi.err; echo "bench(small eig=jacobi) rc=$?"; python scripts/bench_brief.py gpurun_out/r2i_bench_c2_jacobi.json 2>/dev/null | head -2
SPB_THIN_FILTER_BITS_PER_SLOT=2 timeout 600 python bench.py --workload c4 --max-splits 256 --steps 2 --warmup 3 --extras none --no-cpu-baseline > gpurun_out/r2i_bench_c4_filter2.json 2> gpurun_out/r2i_bench_c4_filter2.err; echo "bench c4 (filter 2 bits/slot) rc=$?"; python scripts/bench_brief.py gpurun_out/r2i_bench_c4_filter2.json 2>/dev/null | head -2
python scripts/ncu_step.py > gpurun_out/r2i_ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2i_launches_c2.csv python scripts/ncu_step.py > gpurun_out/r2i_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python scripts/launch_summary.py gpurun_out/r2i_launches_c2.csv "ncu launch list, one c2 step (2,035 splits), round 2 final (CTA-pair Gram, DMMA products, table-join strip)" > gpurun_out/r2i_launches_c2.md; head -30 gpurun_out/r2i_launches_c2.md
