cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 -x > gpurun_out/t_quick.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/t_quick.log
