#!/bin/bash
# Full GPU-side validation in one gpurun call: GPU tests, smoke, both bench arms.  Outputs under gpurun_out/.
#   gpurun --timeout 3600 -- 'bash tools/gpu_validate.sh [tag]'
set -x
cd ${GRAFT_REPO_ROOT:-.}
tag=${1:-final}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/gputest_$tag.log 2>&1; tail -4 gpurun_out/gputest_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err
timeout 900 python bench.py > gpurun_out/bench_n1_$tag.json 2> gpurun_out/bench_n1_$tag.err; tail -c 300 gpurun_out/bench_n1_$tag.err
head -c 400 gpurun_out/bench_n1_$tag.json
