cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 -x -k "count or fasta or config or smoke or pair" > gpurun_out/t_quick.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/t_quick.log
timeout 900 python scripts/kernel_roofline.py > gpurun_out/kernel_roofline.log 2>&1; echo "roofline rc=$?"; cat gpurun_out/kernel_roofline.log
