# Round-2 run V (one B200): GPU tests of the solver's straggler compaction, then the c2 bench line.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 -x -k "straggler or score or reducible or config2 or int32 or erickson or flat" > gpurun_out/r2v_tests.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2v_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --extras none --no-cpu-baseline > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"; python scripts/bench_brief.py gpurun_out/r2v_bench.json 2>/dev/null | head -3; tail -3 gpurun_out/r2v_bench.err
