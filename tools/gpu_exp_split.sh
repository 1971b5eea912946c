#!/bin/bash
# Node-parallel kernel, fp64 state: occupancy / tile variants on one B200 (parity against the C oracle on every line).
#   gpurun --timeout 1500 -- 'bash tools/gpu_exp_split.sh [tag]'
cd ${GRAFT_REPO_ROOT:-.}
tag=${1:-v1}
out=gpurun_out/split_exp_$tag.log
mkdir -p gpurun_out
{
echo "== tests/test_gpu_split.py"
timeout 900 python -m pytest tests/test_gpu_split.py -m gpu -q -x 2>&1 | tail -3
echo "== planned shapes"
timeout 600 python tools/split_check.py --modes mixed,fast_f32 --configs d:0 2>&1 | grep -v compiling
echo "== 5-4-7-10 mixed: 12 warps per SM x 168 registers, one action tile"
SCG_SPLIT_TILES=1 timeout 600 python tools/split_check.py --shapes "[5,4,7,10]" --modes mixed --configs 4:12 2>&1 | grep -v compiling
echo "== 5-4-7-10 mixed: 10 warps per SM, two tiles / one tile"
SCG_SPLIT_TILES=2 timeout 600 python tools/split_check.py --shapes "[5,4,7,10]" --modes mixed --configs 4:10 2>&1 | grep -v compiling
SCG_SPLIT_TILES=1 timeout 600 python tools/split_check.py --shapes "[5,4,7,10]" --modes mixed --configs 4:10 2>&1 | grep -v compiling
echo "== 10 per stage mixed: 16 warps per SM x 128 registers"
SCG_SPLIT_TILES=1 timeout 600 python tools/split_check.py --shapes "10" --modes mixed --configs 8:16 2>&1 | grep -v compiling
} > $out 2>&1
cat $out
