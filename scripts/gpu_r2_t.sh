# Round-2 run T (one B200): the whole GPU suite, smoke and the default bench on the final tree.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2t_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2t_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 > gpurun_out/r2t_tests.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2t_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"; python scripts/bench_brief.py gpurun_out/r2t_bench.json 2>/dev/null; tail -3 gpurun_out/r2t_bench.err
