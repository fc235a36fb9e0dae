cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 -k "score or scorer or config2 or flat or gram" > gpurun_out/t_quick.log 2>&1; echo "pytest rc=$?"
timeout 600 python scripts/phase_profile.py > gpurun_out/phase.log 2>&1; echo "phase rc=$?"
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
tail -5 gpurun_out/t_quick.log; cat gpurun_out/phase.log; cat gpurun_out/bench_quick.json; tail -5 gpurun_out/bench_quick.err
python scripts/ncu_step.py --per-size 64 > gpurun_out/ncu_plain_64.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_c2_ps64.csv python scripts/ncu_step.py --per-size 64 > gpurun_out/ncu_ps64.log 2>&1; echo "ncu launches rc=$?"
python scripts/launch_summary.py gpurun_out/launches_c2_ps64.csv "c2, 64 splits per size" | head -40
