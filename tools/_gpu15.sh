set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_split.py tests/test_gpu_parity.py -q -x > gpurun_out/r2_gputest15.log 2>&1; tail -6 gpurun_out/r2_gputest15.log
timeout 300 python tools/single_env_latency.py > gpurun_out/r2_single_env_latency_v2.log 2>&1; cat gpurun_out/r2_single_env_latency_v2.log
