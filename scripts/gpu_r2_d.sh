# Round-2 evidence run D (one B200): MI / erickson diagnostics, GPU tests (triple-table subflattening scorer), full bench
# (c2 + c3 + c5 + c4 extras), A/B of the subflattening scorers, launch list + ncu of the c2 step and of the c3 scorer.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 60 scripts/bin/mma_rate > gpurun_out/r2d_mma_rate.jsonl 2> gpurun_out/r2d_mma_rate.err; echo "mma_rate rc=$?"; tail -3 gpurun_out/r2d_mma_rate.jsonl
timeout 300 python scripts/diag_mi.py > gpurun_out/r2d_diag_mi.log 2>&1; echo "diag_mi rc=$?"; grep -v "^      ('" gpurun_out/r2d_diag_mi.log | head -60; grep "WRONG" gpurun_out/r2d_diag_mi.log | head
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 -s > gpurun_out/r2d_tests.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2d_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"; python scripts/bench_brief.py gpurun_out/r2d_bench.json; tail -3 gpurun_out/r2d_bench.err
SPB_SUBFLATTEN_KERNEL=staged timeout 600 python bench.py --steps 3 --warmup 3 --workload c3 --no-cpu-baseline > gpurun_out/r2d_bench_c3_staged.json 2> gpurun_out/r2d_bench_c3_staged.err; echo "bench(c3 staged) rc=$?"; python scripts/bench_brief.py gpurun_out/r2d_bench_c3_staged.json
python scripts/ncu_step.py > gpurun_out/r2d_ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2d_launches_c2.csv python scripts/ncu_step.py > gpurun_out/r2d_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python scripts/launch_summary.py gpurun_out/r2d_launches_c2.csv "ncu launch list, one c2 step (2,035 splits), round 2 (CTA-pair Gram, pipelined symv)" > gpurun_out/r2d_launches_c2.md; head -30 gpurun_out/r2d_launches_c2.md
python scripts/ncu_step.py --per-size 8 > gpurun_out/r2d_ncu_plain_small.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gram_u8_umma2|symv_cols" -c 6 -f -o gpurun_out/r2d_prof_c2 python scripts/ncu_step.py --per-size 8 > gpurun_out/r2d_ncu_full.log 2>&1; echo "ncu c2 rc=$?"
python scripts/ncu_step.py --workload c3 --sites 1000000 > gpurun_out/r2d_ncu_plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"subflatten_score" -c 1 -f -o gpurun_out/r2d_prof_c3 python scripts/ncu_step.py --workload c3 --sites 1000000 > gpurun_out/r2d_ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
