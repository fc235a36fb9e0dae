# Round-2 run X (one B200): smoke and the whole GPU suite on the final tree.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2x_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2x_smoke.log
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 > gpurun_out/r2x_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2x_tests.log
