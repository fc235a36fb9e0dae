# Round-2 evidence run B (one B200): MMA-rate microbenchmark, MI diagnostic, all GPU tests (CTA-pair Gram kernel, streamed
# scoring route), c2 bench A/B (pair vs single-CTA Gram, streamed vs batched eigen stage), ncu launch list + full capture.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 90 scripts/bin/mma_rate > gpurun_out/r2b_mma_rate.jsonl 2> gpurun_out/r2b_mma_rate.err; echo "mma_rate rc=$?"; cat gpurun_out/r2b_mma_rate.jsonl
timeout 300 python scripts/diag_mi.py > gpurun_out/r2b_diag_mi.log 2>&1; echo "diag_mi rc=$?"; tail -15 gpurun_out/r2b_diag_mi.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 -s > gpurun_out/r2b_tests.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2b_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --extras none > gpurun_out/r2b_bench_c2.json 2> gpurun_out/r2b_bench_c2.err; echo "bench rc=$?"; python scripts/bench_brief.py gpurun_out/r2b_bench_c2.json; tail -3 gpurun_out/r2b_bench_c2.err
SPB_STREAM_LARGE=0 timeout 600 python bench.py --steps 3 --warmup 3 --extras none --no-cpu-baseline > gpurun_out/r2b_bench_c2_batched.json 2> gpurun_out/r2b_bench_c2_batched.err; echo "bench(batched eigen) rc=$?"; python scripts/bench_brief.py gpurun_out/r2b_bench_c2_batched.json
SPB_STREAM_LARGE=0 SPB_GRAM_KERNEL=1cta timeout 600 python bench.py --steps 3 --warmup 3 --extras none --no-cpu-baseline > gpurun_out/r2b_bench_c2_1cta.json 2> gpurun_out/r2b_bench_c2_1cta.err; echo "bench(1cta, batched) rc=$?"; python scripts/bench_brief.py gpurun_out/r2b_bench_c2_1cta.json
python scripts/ncu_step.py > gpurun_out/r2b_ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2b_launches_c2.csv python scripts/ncu_step.py > gpurun_out/r2b_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python scripts/launch_summary.py gpurun_out/r2b_launches_c2.csv "ncu launch list, one c2 step (2,035 splits), round 2 (CTA-pair Gram + streamed eigen stage)" > gpurun_out/r2b_launches_c2.md; head -24 gpurun_out/r2b_launches_c2.md
python scripts/ncu_step.py --per-size 8 > gpurun_out/r2b_ncu_plain_small.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gram_u8_umma|symv_cols" -c 40 -f -o gpurun_out/r2b_prof_c2 python scripts/ncu_step.py --per-size 8 > gpurun_out/r2b_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | grep r2b
