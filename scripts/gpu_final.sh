# Full evidence run on one B200: tests, smoke, bench (c2, c3, c5, reference arm), ncu launch list + full capture of the Gram kernel.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"; cat gpurun_out/bench_c2.json
timeout 900 python bench.py --steps 3 --warmup 3 --workload c3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"; cat gpurun_out/bench_c3.json
timeout 900 python bench.py --steps 3 --warmup 3 --workload c5 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench c5 rc=$?"; cat gpurun_out/bench_c5.json
timeout 900 python bench.py --steps 2 --warmup 3 --workload c4 --max-splits 256 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"; cat gpurun_out/bench_c4.json
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_c2.json 2> gpurun_out/bench_ref_c2.err; echo "ref rc=$?"; cat gpurun_out/bench_ref_c2.json
python scripts/ncu_step.py > gpurun_out/ncu_plain_full.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_c2.csv python scripts/ncu_step.py > gpurun_out/ncu_full.log 2>&1; echo "ncu launches rc=$?"
python scripts/launch_summary.py gpurun_out/launches_c2.csv "ncu launch list, one c2 step (2,035 splits)" > gpurun_out/launches_c2.md; head -30 gpurun_out/launches_c2.md
python scripts/ncu_step.py --per-size 16 > gpurun_out/ncu_plain_small.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gram_u8_umma|symv_|u8_kernel|strip_|hi_strip" -c 80 -f -o gpurun_out/prof_c2 python scripts/ncu_step.py --per-size 16 > gpurun_out/ncu_small.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -5
timeout 300 python scripts/bench_c1.py > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo "bench c1 rc=$?"; cat gpurun_out/bench_c1.json
