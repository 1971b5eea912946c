"""gym.vector / Stable-Baselines-style adapter over the batched envs (ecosystem plumbing).

The reference ships single envs and leaves vectorisation to the agent library (SubprocVecEnv etc., one Python
process per env).  This adapter exposes the attributes those libraries read -- `num_envs`, `single_action_space`,
`single_observation_space`, `action_space`, `observation_space`, `reset()`, `step()`, `step_async()/step_wait()`,
`close()` -- on top of ONE handle that simulates all envs on the GPU.  Auto-reset follows the gym-vector convention
(the observation returned with done=True is the first one of the next episode; the episode return of the finished
episode is reported in `infos`).

    venv = GpuVecEnv("supplychain-ntom-v0", num_envs=4096, numpy_io=True)     # numpy in/out like SubprocVecEnv
    obs = venv.reset(); obs, rewards, dones, infos = venv.step(actions)
"""
import numpy as np

from . import registration


class GpuVecEnv(object):
    def __init__(self, env_id, num_envs, numpy_io=False, **kwargs):
        kwargs.setdefault("auto_reset", True)
        self.env = registration.make(env_id, num_envs=num_envs, **kwargs)
        self.num_envs = int(num_envs)
        self.numpy_io = bool(numpy_io)
        self.single_action_space = self.action_space = self.env.action_space
        self.single_observation_space = self.observation_space = self.env.observation_space
        self._pending = None
        self._ep_ret = None

    def seed(self, seed=None):
        self.env.seed(seed)
        return [seed] * self.num_envs

    def _out(self, x):
        return x.cpu().numpy() if self.numpy_io else x

    def reset(self):
        import torch
        obs = self.env.reset()
        self._ep_ret = torch.zeros(self.num_envs, dtype=torch.float64, device=obs.device)
        return self._out(obs)

    def step_async(self, actions):
        self._pending = actions

    def step_wait(self):
        obs, rew, done, info = self.env.step(self._pending)
        self._pending = None
        self._ep_ret += rew.double()
        finished = done.nonzero().flatten()
        infos = dict(info) if isinstance(info, dict) else {}
        if finished.numel():
            # SB3 / gym.wrappers.RecordEpisodeStatistics style: returns of the episodes that just ended
            infos["episode"] = {"env_index": self._out(finished), "r": self._out(self._ep_ret[finished].clone())}
            self._ep_ret[finished] = 0.0
        return self._out(obs), self._out(rew), self._out(done), infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        self.env.close()

    def __len__(self):
        return self.num_envs
