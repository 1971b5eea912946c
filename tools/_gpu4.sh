set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python tools/split_check.py --configs 4:3,8:3,4:2 > gpurun_out/r2_split2.log 2>&1
grep -v "^\[scg\]" gpurun_out/r2_split2.log
export SCG_SPLIT_G=4 SCG_SPLIT_MINB=3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sc_split -s 2 -c 1 -f -o gpurun_out/r2_split_f32_v1 python tools/prof_split.py "[5,4,7,10]" fast_f32 > gpurun_out/ncu_c.log 2>&1
tail -2 gpurun_out/ncu_c.log
unset SCG_SPLIT_G SCG_SPLIT_MINB
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sc_layered_rollout -s 2 -c 1 -f -o gpurun_out/r2_scv0_mixed_v1 python tools/prof_rollout.py v0 mixed 1048576 64 > gpurun_out/ncu_d.log 2>&1
tail -2 gpurun_out/ncu_d.log
cp gym_supplychain_b200/_jit/*.name gpurun_out/ 2>/dev/null
ls gym_supplychain_b200/_jit | head -40
