#!/bin/bash
# Full GPU-side validation in one gpurun call: GPU tests, smoke, both bench arms, then (optional, `full`) the secondary
# measurements and captures that profiles/ quotes.  Outputs under gpurun_out/.
#   gpurun --timeout 3600 -- 'bash tools/gpu_validate.sh [tag] [full]'
set -x
cd ${GRAFT_REPO_ROOT:-.}
tag=${1:-final}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/gputest_$tag.log 2>&1; tail -4 gpurun_out/gputest_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err
timeout 900 python bench.py > gpurun_out/bench_n1_$tag.json 2> gpurun_out/bench_n1_$tag.err; tail -c 300 gpurun_out/bench_n1_$tag.err
head -c 400 gpurun_out/bench_n1_$tag.json
[ "$2" = "full" ] || exit 0
CASES="ntom:mixed,ntom:fast_f32,2perstage:mixed,2perstage:fast_f32,linear:mixed,linear:fast_f32,v0:mixed,v0:fast_f32,oneonen:mixed,oneonen:fast_f32,multiproduct:mixed,multiproduct:fast_f32"
timeout 600 python tools/quick_rollout.py --cases $CASES > gpurun_out/quick_rollout_$tag.log 2>&1
timeout 600 python tools/midsize_chains.py > gpurun_out/midsize_$tag.log 2>&1
timeout 900 python tools/split_check.py --modes fast_f32,mixed --configs d:0 2>&1 | grep -v compiling > gpurun_out/split_$tag.log
timeout 600 python tools/quick_bg.py > gpurun_out/quick_bg_$tag.log 2>&1
# captures (after the un-profiled runs above exited): launch list of the bench command, full capture of the headline kernel
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-other --no-cpu-baseline > gpurun_out/ncu_bench_$tag.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:sc_layered_rollout -s 2 -c 1 -f -o gpurun_out/r2_ntom_mixed_$tag \
    python tools/prof_rollout.py ntom mixed > gpurun_out/ncu_ntom_$tag.log 2>&1; tail -2 gpurun_out/ncu_ntom_$tag.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:sc_split_rollout -s 2 -c 1 -f -o gpurun_out/r2_split_mixed_$tag \
    python tools/prof_split.py "[5,4,7,10]" mixed > gpurun_out/ncu_split_$tag.log 2>&1; tail -2 gpurun_out/ncu_split_$tag.log
tail -n 30 gpurun_out/quick_rollout_$tag.log gpurun_out/midsize_$tag.log gpurun_out/split_$tag.log gpurun_out/quick_bg_$tag.log
