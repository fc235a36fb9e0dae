cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 -k "${1:-rank1 or erickson}" > gpurun_out/t_quick.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/t_quick.log
