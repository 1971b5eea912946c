set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=gpurun_out/r2_baked1.log; : > $L
echo "== prebuilt" >> $L; timeout 600 python tools/quick_rollout.py >> $L 2>&1
echo "== jit unbaked" >> $L; SCG_JIT_FORCE=1 timeout 600 python tools/quick_rollout.py >> $L 2>&1
echo "== jit baked" >> $L; SCG_JIT_FORCE=1 SCG_JIT_BAKE=1 timeout 600 python tools/quick_rollout.py >> $L 2>&1
grep -v "^\[scg\]" $L
