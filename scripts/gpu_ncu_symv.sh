# ncu --set full of the int32 product kernel (32 matrices per launch) -> gpurun_out/prof_symv*
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python scripts/ncu_step.py --per-size 32 > gpurun_out/ncu_plain_small.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"symv_cols_i32" -c 2 -f -o gpurun_out/prof_symv python scripts/ncu_step.py --per-size 32 > gpurun_out/ncu_symv.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/prof_symv.ncu-rep --page raw --csv > gpurun_out/prof_symv_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_symv.ncu-rep --page details --csv > gpurun_out/prof_symv_details.csv 2>/dev/null
ls -la gpurun_out/prof_symv*
