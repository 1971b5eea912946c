set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py > gpurun_out/r2_gputest2.log 2>&1
tail -5 gpurun_out/r2_gputest2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke2.log 2>&1; tail -3 gpurun_out/r2_smoke2.log
timeout 900 python bench.py > gpurun_out/r2_bench_n1_v1.json 2> gpurun_out/r2_bench_n1_v1.err; tail -c 600 gpurun_out/r2_bench_n1_v1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_v1.json 2> gpurun_out/r2_bench_ref_v1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sc_layered_rollout -s 2 -c 1 -f -o gpurun_out/r2_ntom_mixed_v11 python tools/prof_rollout.py ntom mixed > gpurun_out/ncu_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bg_kernel -s 2 -c 1 -f -o gpurun_out/r2_bg_v2 python tools/prof_bg.py > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out | tail -12
head -c 3000 gpurun_out/r2_bench_n1_v1.json
