set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cat > /tmp/scv0.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import gym_supplychain_b200 as gsc
for env_id, kw in (("supplychain-v0", {}), ("supplychain-ntom-v0", dict(num_products=3, stochastic_leadtimes=True, avg_leadtime=2, max_leadtime=4))):
    for mode in ("mixed", "fast_f32"):
        n, T = 1 << 20, 72
        if "ntom" in env_id: n, T = 1 << 19, 64
        env = gsc.make(env_id, num_envs=n, mode=mode, seed=0, auto_reset=True, **kw)
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
        print(env_id, kw, mode, "variant", env.scg_info.kernel_variant, "%.3f ms %.3e steps/s frac %.3f rew_sum %.9e" % (ms, v, v * B / 1e9 / 6549.1, rew.double().sum().item()), flush=True)
        del env, acts, obs, rew, done; torch.cuda.empty_cache()
PY
timeout 600 python /tmp/scv0.py > gpurun_out/r2_scv0_quick3.log 2>&1; grep -v "^\[scg\]" gpurun_out/r2_scv0_quick3.log
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/gputest_v15.log 2>&1; tail -4 gpurun_out/gputest_v15.log
