set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_gputest11.log 2>&1; tail -6 gpurun_out/r2_gputest11.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_ref_v2.json 2> gpurun_out/r2_bench_ref_v2.err
timeout 900 python bench.py > gpurun_out/r2_bench_n1_v2.json 2> gpurun_out/r2_bench_n1_v2.err; tail -c 400 gpurun_out/r2_bench_n1_v2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_v12.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
head -c 600 gpurun_out/r2_bench_n1_v2.json
