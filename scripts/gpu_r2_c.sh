# Round-2 evidence run C (one B200): GPU tests, MI diagnostic, c2 bench A/B (CTA-pair vs single-CTA Gram), count-kernel A/B.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python scripts/diag_mi.py > gpurun_out/r2c_diag_mi.log 2>&1; echo "diag_mi rc=$?"; tail -12 gpurun_out/r2c_diag_mi.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 -s > gpurun_out/r2c_tests.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2c_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --extras none --no-cpu-baseline > gpurun_out/r2c_bench_c2.json 2> gpurun_out/r2c_bench_c2.err; echo "bench rc=$?"; python scripts/bench_brief.py gpurun_out/r2c_bench_c2.json; tail -3 gpurun_out/r2c_bench_c2.err
SPB_GRAM_KERNEL=1cta timeout 600 python bench.py --steps 3 --warmup 3 --extras none --no-cpu-baseline > gpurun_out/r2c_bench_c2_1cta.json 2> gpurun_out/r2c_bench_c2_1cta.err; echo "bench(1cta) rc=$?"; python scripts/bench_brief.py gpurun_out/r2c_bench_c2_1cta.json
SPB_STREAM_LARGE=1 timeout 600 python bench.py --steps 3 --warmup 3 --extras none --no-cpu-baseline > gpurun_out/r2c_bench_c2_stream.json 2> gpurun_out/r2c_bench_c2_stream.err; echo "bench(stream) rc=$?"; python scripts/bench_brief.py gpurun_out/r2c_bench_c2_stream.json
for kern in class cache stream; do SPB_COUNT_KERNEL=$kern timeout 600 python scripts/kernel_roofline.py > gpurun_out/r2c_roofline_$kern.log 2>&1; echo "roofline($kern) rc=$?"; cat gpurun_out/r2c_roofline_$kern.log; done
python scripts/ncu_step.py --per-size 8 > gpurun_out/r2c_ncu_plain_small.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gram_u8_umma2|symv_cols" -c 12 -f -o gpurun_out/r2c_prof_c2 python scripts/ncu_step.py --per-size 8 > gpurun_out/r2c_ncu_full.log 2>&1; echo "ncu full rc=$?"
