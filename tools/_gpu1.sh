set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_ab1.log 2>&1
for tag in "" _base _r200t64 _r216t32 _r184t32 _r168t32; do echo "== lib$tag" >> gpurun_out/r2_ab1.log; SCG_LIB=$PWD/gym_supplychain_b200/libscg$tag.so timeout 300 python tools/quick_rollout.py --cases ntom:mixed --iters 6 >> gpurun_out/r2_ab1.log 2>&1; done
echo "== default all cases" >> gpurun_out/r2_ab1.log
timeout 600 python tools/quick_rollout.py >> gpurun_out/r2_ab1.log 2>&1
echo "== base all cases" >> gpurun_out/r2_ab1.log
SCG_LIB=$PWD/gym_supplychain_b200/libscg_base.so timeout 600 python tools/quick_rollout.py --cases 2perstage:mixed,linear:mixed >> gpurun_out/r2_ab1.log 2>&1
timeout 900 python tools/error_survey.py > gpurun_out/r2_err1.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_gputest1.log 2>&1
tail -5 gpurun_out/r2_gputest1.log
cat gpurun_out/r2_ab1.log gpurun_out/r2_err1.log
