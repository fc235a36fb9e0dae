# Round-2 run H (one B200): G0 Q kernel variants (micro-benchmark), then the c2 step with the column kernel and with the DMMA kernels.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python scripts/symv_bench.py 4096 48 > gpurun_out/r2h_symv_4096.log 2>&1; echo "symv 4096 rc=$?"; cat gpurun_out/r2h_symv_4096.log | grep variant
timeout 300 python scripts/symv_bench.py 1024 512 > gpurun_out/r2h_symv_1024.log 2>&1; echo "symv 1024 rc=$?"; cat gpurun_out/r2h_symv_1024.log | grep variant
for v in 0 1 4 6; do
SPB_SYMV_KERNEL=$v timeout 600 python bench.py --steps 3 --warmup 3 --extras none --no-cpu-baseline > gpurun_out/r2h_bench_c2_v$v.json 2> gpurun_out/r2h_bench_c2_v$v.err; echo "bench symv=$v rc=$?"; python scripts/bench_brief.py gpurun_out/r2h_bench_c2_v$v.json | head -3
done
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 -x -k "score or config2 or int32 or gram" > gpurun_out/r2h_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2h_tests.log
