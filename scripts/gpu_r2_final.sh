# Round-2 final evidence run (one B200): build check, smoke, GPU tests, reference arm, full bench, launch list, ncu --set full.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2z_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 > gpurun_out/r2z_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2z_tests.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err; echo "reference rc=$?"; cut -c1-260 gpurun_out/r2z_bench_reference.json
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"; python scripts/bench_brief.py gpurun_out/r2z_bench.json 2>/dev/null; tail -3 gpurun_out/r2z_bench.err
python scripts/ncu_step.py > gpurun_out/r2z_ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2z_launches_c2.csv python scripts/ncu_step.py > gpurun_out/r2z_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python scripts/launch_summary.py gpurun_out/r2z_launches_c2.csv "ncu launch list, one c2 step (2,035 splits), round 2 final (CTA-pair Gram, DMMA products and inner products, table-join strip, 512-matrix solver batches)" > gpurun_out/r2z_launches_c2.md; head -24 gpurun_out/r2z_launches_c2.md
python scripts/ncu_step.py --per-size 64 > gpurun_out/r2z_ncu_plain_small.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gram_u8_umma2|symv_dmma|hi_join|dot_dmma|strip_rows_dmma|gram_score_warp" -c 14 -f -o gpurun_out/r2z_prof_c2 python scripts/ncu_step.py --per-size 64 > gpurun_out/r2z_ncu_full.log 2>&1; echo "ncu c2 rc=$?"; ls -la gpurun_out/r2z_prof_c2.ncu-rep
