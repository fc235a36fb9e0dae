cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_plain.log 2>&1 && \
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python __graft_entry__.py smoke > gpurun_out/sanitizer_memcheck_smoke.log 2>&1; echo "memcheck rc=$?"
tail -12 gpurun_out/sanitizer_memcheck_smoke.log
