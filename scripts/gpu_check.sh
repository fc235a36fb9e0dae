cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T="timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300"
$T > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python scripts/phase_profile.py > gpurun_out/phase.log 2>&1; echo "phase rc=$?"
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"
tail -30 gpurun_out/t_all.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/phase.log; cat gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
