set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_sync1.log; : > $L
for tag in "" _sync64 _sync128 _sync256 _t128 _base; do echo "== lib$tag" >> $L; SCG_LIB=$PWD/gym_supplychain_b200/libscg$tag.so timeout 300 python tools/quick_rollout.py --cases ntom:mixed,2perstage:mixed --iters 6 >> $L 2>&1; done
echo "== split default" >> $L; timeout 600 python tools/split_check.py --configs 4:0 --shapes "[5,4,7,10]" >> $L 2>&1
echo "== split sync" >> $L; SCG_JIT_DEFINES="-DSCG_SYNC_PERIOD=1" timeout 600 python tools/split_check.py --configs 4:0 --shapes "[5,4,7,10]" >> $L 2>&1
echo "== split sync 8 warps" >> $L; SCG_SPLIT_WARPS=8 SCG_JIT_DEFINES="-DSCG_SYNC_PERIOD=1" timeout 600 python tools/split_check.py --configs 4:0 --shapes "[5,4,7,10]" >> $L 2>&1
echo "== split 8 warps nosync" >> $L; SCG_SPLIT_WARPS=8 timeout 600 python tools/split_check.py --configs 4:0 --shapes "[5,4,7,10]" >> $L 2>&1
grep -v "^\[scg\]" $L
