set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/quick_bg.py > gpurun_out/r2_bg_quick1.log 2>&1; cat gpurun_out/r2_bg_quick1.log
timeout 1200 python -m pytest tests -m gpu -q -k "beergame or runtime_instantiated or config5 or eight_million" > gpurun_out/r2_gputest6.log 2>&1; tail -5 gpurun_out/r2_gputest6.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
