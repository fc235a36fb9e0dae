# Round-2 run K (one B200): GPU tests after the DMMA inner products / thin-kernel changes, full bench, library-op timings.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 -x > gpurun_out/r2k_tests.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2k_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"; python scripts/bench_brief.py gpurun_out/r2k_bench.json 2>/dev/null; tail -3 gpurun_out/r2k_bench.err
SPB_DOT_KERNEL=simt timeout 600 python bench.py --steps 3 --warmup 3 --extras none --no-cpu-baseline > gpurun_out/r2k_bench_c2_simtdot.json 2> gpurun_out/r2k_bench_c2_simtdot.err; echo "bench(dot=simt) rc=$?"; python scripts/bench_brief.py gpurun_out/r2k_bench_c2_simtdot.json 2>/dev/null | head -2
timeout 600 python scripts/time_library_ops.py > gpurun_out/r2k_library_ops.json 2> gpurun_out/r2k_library_ops.err; echo "library ops rc=$?"; cat gpurun_out/r2k_library_ops.json; tail -3 gpurun_out/r2k_library_ops.err
