# ncu launch list (gpu__time_duration.sum per launch) of one c2 step -> gpurun_out/launches_c2.md
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python scripts/ncu_step.py > gpurun_out/ncu_plain_full.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_c2.csv python scripts/ncu_step.py > gpurun_out/ncu_full.log 2>&1; echo "ncu launches rc=$?"
python scripts/launch_summary.py gpurun_out/launches_c2.csv "ncu launch list, one c2 step (2,035 splits)" > gpurun_out/launches_c2.md; head -40 gpurun_out/launches_c2.md
