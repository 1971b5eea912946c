#!/bin/bash
# A/B on one B200: one-warp blocks (bulk-copy addresses in uniform registers) against multi-warp blocks.
#   gpurun --timeout 1500 -- 'bash tools/gpu_exp_onewarp.sh [tag]'
# libscg.so = the shipped per-kernel choice; libscg_w0.so = no kernel one-warp (build.py --tag=w0 -DSCG_ONE_WARP=-1);
# libscg_w1all.so = every kernel one-warp (--tag=w1all -DSCG_ONE_WARP=1).
cd ${GRAFT_REPO_ROOT:-.}
tag=${1:-v1}
out=gpurun_out/onewarp_$tag.log
mkdir -p gpurun_out
L=gym_supplychain_b200
CASES="ntom:mixed,ntom:fast_f32,2perstage:mixed,2perstage:fast_f32,linear:mixed,linear:fast_f32,v0:mixed,v0:fast_f32,oneonen:mixed,oneonen:fast_f32,multiproduct:mixed,multiproduct:fast_f32"
{
echo "== shipped launch shapes (libscg.so)"
timeout 600 python tools/quick_rollout.py --cases $CASES
echo "== multi-warp blocks everywhere (libscg_w0.so)"
SCG_LIB=$L/libscg_w0.so timeout 600 python tools/quick_rollout.py --cases $CASES
echo "== one-warp blocks everywhere, same warps per SM (libscg_w1all.so)"
SCG_LIB=$L/libscg_w1all.so timeout 600 python tools/quick_rollout.py --cases $CASES
echo "== run-time instantiated mid-size chains: planned block size"
timeout 600 python tools/midsize_chains.py
echo "== run-time instantiated mid-size chains: one-warp blocks (SCG_JIT_WARPS=1)"
SCG_JIT_WARPS=1 timeout 600 python tools/midsize_chains.py
echo "== node-parallel kernel: planned block size (fp64 state: 1 warp, fp32: 4)"
timeout 600 python tools/split_check.py --modes fast_f32,mixed --configs d:0 2>&1 | grep -v compiling
} > $out 2>&1
cat $out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:sc_layered_rollout -s 2 -c 1 -f -o gpurun_out/r2_ntom_mixed_v12 \
    python tools/prof_rollout.py ntom mixed > gpurun_out/ncu_ntom_v12.log 2>&1; tail -2 gpurun_out/ncu_ntom_v12.log
