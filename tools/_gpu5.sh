set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_gputest5.log 2>&1
tail -15 gpurun_out/r2_gputest5.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
