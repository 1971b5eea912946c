"""Drop-in single-env use (num_envs=None: numpy in/out, old-gym API, exact_f64): wall time per env.step() on the GPU box."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_supplychain_b200 as gsc

cases = [("supplychain-ntom-v0", {}), ("supplychain-ntom-v0", {"mode": "fast_f64"}), ("supplychain-ntom-v0", {"mode": "mixed"}),
         ("supplychain-2perstage-v0", {}), ("supplychain-2perstage-v0", {"mode": "fast_f64"}), ("beergame-v2", {}), ("beergame-v0", {})]
for env_id, kw in cases:
    env = gsc.make(env_id, **kw)
    env.seed(0)
    obs = env.reset()
    acts = ([env.action_space.sample() for _ in range(64)] if env.action_space is not None
            else [np.random.RandomState(i).randint(0, 5, 4) for i in range(64)])
    for a in acts[:10]:
        env.step(a)
    env.reset()
    t0 = time.perf_counter(); n = 0
    for ep in range(3):
        done = False
        while not done:
            obs, r, done, info = env.step(acts[n % 64]); n += 1
        env.reset()
    dt = time.perf_counter() - t0
    print("%-26s %-20s single env: %6.1f us per step() (%d steps incl. %d resets)" % (env_id, kw.get("mode", "default (exact)"), 1e6 * dt / n, n, 3), flush=True)
