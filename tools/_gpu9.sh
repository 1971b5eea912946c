set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sc_split -s 2 -c 1 -f -o gpurun_out/r2_split_f32_v2 python tools/prof_split.py "[5,4,7,10]" fast_f32 > gpurun_out/ncu_e.log 2>&1
tail -3 gpurun_out/ncu_e.log
for f in gym_supplychain_b200/_jit/*.name; do if grep -q "sc_split_rolloutIffLi5ELi4ELi7ELi10ELi2ELi2ELi4ELb0" $f; then cp ${f%.name}.cubin gpurun_out/r2_split_f32_v2.cubin; fi; done
ls -la gpurun_out/*.cubin
timeout 900 python -m pytest tests/test_gpu_split.py tests/test_gpu_edges.py tests/test_gpu_fullsize.py -q > gpurun_out/r2_gputest9.log 2>&1; tail -4 gpurun_out/r2_gputest9.log
timeout 300 python tools/single_env_latency.py > gpurun_out/r2_single_env_latency.log 2>&1; cat gpurun_out/r2_single_env_latency.log
