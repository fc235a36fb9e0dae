cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T="timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300"
$T > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"
timeout 600 python scripts/phase_profile.py > gpurun_out/phase.log 2>&1; echo "phase rc=$?"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"
tail -5 gpurun_out/t_all.log; cat gpurun_out/phase.log; cat gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
python scripts/ncu_step.py > gpurun_out/ncu_plain_full.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_c2.csv python scripts/ncu_step.py > gpurun_out/ncu_full.log 2>&1; echo "ncu launches rc=$?"
python scripts/ncu_step.py --per-size 3 > gpurun_out/ncu_plain_small.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gram_u8_umma -o gpurun_out/prof_gram python scripts/ncu_step.py --per-size 3 > gpurun_out/ncu_small.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail; tail -3 gpurun_out/ncu_full.log gpurun_out/ncu_small.log
