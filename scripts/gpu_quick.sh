cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 -k "large_side or generate_alignment or fasta or quirks" > gpurun_out/t_quick.log 2>&1; echo "pytest rc=$?"
tail -22 gpurun_out/t_quick.log
