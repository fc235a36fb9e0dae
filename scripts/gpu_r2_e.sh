# Round-2 evidence run E (one B200): diagnostics, GPU tests, full bench, ncu of the c2 step (symv v3) and of the c3 scorer.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python scripts/diag_mi.py > gpurun_out/r2e_diag_mi.log 2>&1; echo "diag_mi rc=$?"; grep "matches golden" gpurun_out/r2e_diag_mi.log | sort | uniq -c
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 -s > gpurun_out/r2e_tests.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2e_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; python scripts/bench_brief.py gpurun_out/r2e_bench.json; tail -3 gpurun_out/r2e_bench.err
python scripts/ncu_step.py > gpurun_out/r2e_ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2e_launches_c2.csv python scripts/ncu_step.py > gpurun_out/r2e_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python scripts/launch_summary.py gpurun_out/r2e_launches_c2.csv "ncu launch list, one c2 step (2,035 splits), round 2 (CTA-pair Gram, symv v3)" > gpurun_out/r2e_launches_c2.md; head -22 gpurun_out/r2e_launches_c2.md
python scripts/ncu_step.py --per-size 8 > gpurun_out/r2e_ncu_plain_small.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gram_u8_umma2|symv_cols" -c 6 -f -o gpurun_out/r2e_prof_c2 python scripts/ncu_step.py --per-size 8 > gpurun_out/r2e_ncu_full.log 2>&1; echo "ncu c2 rc=$?"
python scripts/ncu_step.py --workload c3 --sites 1000000 > gpurun_out/r2e_ncu_plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"subflatten_score" -c 1 -f -o gpurun_out/r2e_prof_c3 python scripts/ncu_step.py --workload c3 --sites 1000000 > gpurun_out/r2e_ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
for kern in class cache; do SPB_COUNT_KERNEL=$kern python scripts/ncu_count.py > gpurun_out/r2e_count_plain_$kern.log 2>&1 && \
SPB_COUNT_KERNEL=$kern ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"count_" -c 1 -f -o gpurun_out/r2e_prof_count_$kern python scripts/ncu_count.py > gpurun_out/r2e_ncu_count_$kern.log 2>&1; echo "ncu count($kern) rc=$?"; cat gpurun_out/r2e_count_plain_$kern.log; done
