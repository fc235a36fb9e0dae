# Round-2 evidence run A (one B200): GPU tests, microbenchmarks (MMA rate, int8 peak), count-kernel A/B, bench + reference arm.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 -s > gpurun_out/r2a_tests.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2a_tests.log
timeout 120 scripts/bin/mma_rate > gpurun_out/r2a_mma_rate.jsonl 2> gpurun_out/r2a_mma_rate.err; echo "mma_rate rc=$?"; cat gpurun_out/r2a_mma_rate.jsonl
timeout 120 python scripts/int8_peak.py > gpurun_out/r2a_int8_peak.json 2> gpurun_out/r2a_int8_peak.err; echo "int8 rc=$?"; cat gpurun_out/r2a_int8_peak.json
timeout 600 python scripts/kernel_roofline.py > gpurun_out/r2a_roofline_stream.log 2>&1; echo "roofline(stream) rc=$?"; cat gpurun_out/r2a_roofline_stream.log
SPB_COUNT_KERNEL=cache timeout 600 python scripts/kernel_roofline.py > gpurun_out/r2a_roofline_cache.log 2>&1; echo "roofline(cache) rc=$?"; cat gpurun_out/r2a_roofline_cache.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2a_smoke.log
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; cat gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/r2a_bench_ref.json
