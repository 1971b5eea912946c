set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
M="l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum,gpu__time_duration.sum"
timeout 300 ncu --metrics $M --clock-control none -k regex:sc_layered_rollout -s 2 -c 1 --csv --log-file gpurun_out/r2_bank_bulk.csv python tools/prof_rollout.py ntom mixed > /dev/null 2>&1
SCG_LIB=$PWD/gym_supplychain_b200/libscg_nobulk.so timeout 300 ncu --metrics $M --clock-control none -k regex:sc_layered_rollout -s 2 -c 1 --csv --log-file gpurun_out/r2_bank_nobulk.csv python tools/prof_rollout.py ntom mixed > /dev/null 2>&1
tail -5 gpurun_out/r2_bank_bulk.csv gpurun_out/r2_bank_nobulk.csv
timeout 300 python bench.py --config c2 --no-cpu-baseline --no-other > gpurun_out/r2_bench_c2_v2.json 2> gpurun_out/r2_bench_c2_v2.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c2_v2.json'))['variants']
for k in ('step_api_one_launch_per_period','step_api_360_periods_in_one_cuda_graph','fused_360_period_rollout'): print(k, d[k])"
