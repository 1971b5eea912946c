set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2_gputest14.log 2>&1; tail -4 gpurun_out/r2_gputest14.log; timeout 600 python tools/split_check.py --configs 4:0,8:0 > gpurun_out/r2_split3.log 2>&1; grep -v "^\[scg\]" gpurun_out/r2_split3.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== fixed ring ON (new launch plan)" > gpurun_out/r2_mid3.log; SCG_SPLIT_G=0 timeout 600 python /dev/stdin >> gpurun_out/r2_mid3.log 2>&1 <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import gym_supplychain_b200 as gsc
for kw in (dict(nodes_per_echelon=[2,3,2,4], num_products=3, avg_leadtime=3, max_leadtime=3), dict(nodes_per_echelon=[3,3,3,3], num_products=3), dict(nodes_per_echelon=[4,4,4,4], num_products=2), dict(nodes_per_echelon=[2,2,2,2], num_products=2), dict(nodes_per_echelon=[2,3,2,4], num_products=1, avg_leadtime=3, max_leadtime=3)):
    for mode in ("mixed", "fast_f32"):
        n, T = 1 << 19, 64
        env = gsc.make("supplychain-ntom-v0", num_envs=n, mode=mode, seed=0, auto_reset=True, **kw)
        env.reset()
        A, O = env.n_actions, env.n_obs
        acts = torch.empty((T, n, A), device="cuda").uniform_(-1, 1)
        obs = torch.empty((T, n, O), device="cuda"); rew = torch.empty((T, n), device="cuda"); done = torch.empty((T, n), dtype=torch.uint8, device="cuda")
        for _ in range(2): env.rollout(T, acts, out=(obs, rew, done))
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(4): env.rollout(T, acts, out=(obs, rew, done))
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 4
        v = n * T / (ms * 1e-3); B = 4 * A + 4 * O + 5
        print(kw, mode, "variant", env.scg_info.kernel_variant, "%.3f ms %.3e steps/s frac %.3f" % (ms, v, v * B / 1e9 / 6549.1), flush=True)
        del env, acts, obs, rew, done; torch.cuda.empty_cache()
PY
grep -v "^\[scg\]" gpurun_out/r2_mid3.log
