"""Oracle (test infrastructure): batched classic-control envs + the reference's parallel-env
wrappers, restated in NumPy fp32.

Dynamics: the reference gets CartPole/Pendulum from the un-vendored
ClassicControlEnvironments.jl@main (call sites README.md:50,76, benchmark/bench_utils.jl:14,20).
Restated from the Gymnasium CartPole-v1 / Pendulum-v1 equations; PARITY UNPINNED.
Every fp32 operation is written out one rounding at a time (no FMA), sin/cos are the
correctly rounded fp32 values (computed in fp64, rounded once), so the CUDA kernels can
reproduce trajectories bit-for-bit.

Wrappers follow the reference line by line:
  ParallelEnv.act            environment_wrappers/multithreadedParallelEnv.jl:47-74
  MonitorWrapper             environment_wrappers/monitorWrapperEnv.jl:36-70
  RunningMeanStd             environment_wrappers/normalizeWrapperEnv.jl:8-50
  NormalizeWrapper           environment_wrappers/normalizeWrapperEnv.jl:111-197
"""
from collections import deque

import numpy as np

from . import philox

f32 = np.float32


def _sincos32(theta):
    t64 = theta.astype(np.float64)
    return np.sin(t64).astype(f32), np.cos(t64).astype(f32)


class CartPoleBatch:
    """Gymnasium CartPole-v1, Euler integrator, fp32."""
    kind = "cartpole"
    obs_dim = 4
    act_kind = "discrete"
    n_actions = 2

    GRAVITY = f32(9.8)
    MASSPOLE = f32(0.1)
    TOTAL_MASS = f32(0.1) + f32(1.0)
    LENGTH = f32(0.5)
    POLEMASS_LENGTH = f32(0.1) * f32(0.5)
    FORCE_MAG = f32(10.0)
    TAU = f32(0.02)
    THETA_THR = f32(12 * 2 * np.pi / 360)
    X_THR = f32(2.4)
    FOUR_THIRDS = f32(4.0 / 3.0)

    def __init__(self, n_envs, seed=0, max_steps=500, gid_offset=0, act_start=1):
        self.n = n_envs
        self.seed = seed
        self.max_steps = max_steps
        self.gid = np.arange(n_envs, dtype=np.int64) + gid_offset
        self.act_start = act_start
        self.state = np.zeros((4, n_envs), dtype=f32)
        self.steps = np.zeros(n_envs, dtype=np.int32)
        self.episode = np.zeros(n_envs, dtype=np.int64)
        self.terminated = np.zeros(n_envs, dtype=bool)
        self.truncated = np.zeros(n_envs, dtype=bool)
        self.reset_all()

    def reset_idx(self, idx):
        xs = philox.philox4x32(self.gid[idx], self.episode[idx], 0, philox.TAG_RESET, self.seed)
        for k in range(4):
            u = philox.u01_f32(xs[k])
            self.state[k, idx] = f32(-0.05) + f32(0.1) * u
        self.episode[idx] += 1
        self.steps[idx] = 0
        self.terminated[idx] = False
        self.truncated[idx] = False

    def reset_all(self):
        self.reset_idx(np.arange(self.n))

    def obs(self):
        return self.state.T.copy()  # (n, 4)

    def step(self, actions):
        """actions: env-space ints (act_start based). Returns rewards (n,) f32."""
        a = np.asarray(actions).astype(np.int64) - self.act_start
        x, x_dot, th, th_dot = (self.state[k] for k in range(4))
        force = np.where(a == 1, self.FORCE_MAG, -self.FORCE_MAG).astype(f32)
        sinth, costh = _sincos32(th)
        temp = (force + (self.POLEMASS_LENGTH * (th_dot * th_dot)) * sinth) / self.TOTAL_MASS
        thetaacc = ((self.GRAVITY * sinth) - (costh * temp)) / (
            self.LENGTH * (self.FOUR_THIRDS - ((self.MASSPOLE * (costh * costh)) / self.TOTAL_MASS)))
        xacc = temp - (((self.POLEMASS_LENGTH * thetaacc) * costh) / self.TOTAL_MASS)
        x_new = x + self.TAU * x_dot
        x_dot_new = x_dot + self.TAU * xacc
        th_new = th + self.TAU * th_dot
        th_dot_new = th_dot + self.TAU * thetaacc
        self.state = np.stack([x_new, x_dot_new, th_new, th_dot_new]).astype(f32)
        self.steps += 1
        self.terminated = ((x_new < -self.X_THR) | (x_new > self.X_THR)
                           | (th_new < -self.THETA_THR) | (th_new > self.THETA_THR))
        self.truncated = self.steps >= self.max_steps
        return np.ones(self.n, dtype=f32)


class PendulumBatch:
    """Gymnasium Pendulum-v1 (g=10), fp32, 200-step TimeLimit."""
    kind = "pendulum"
    obs_dim = 3
    act_kind = "continuous"
    act_dim = 1
    act_low = np.array([-2.0], dtype=f32)
    act_high = np.array([2.0], dtype=f32)

    MAX_SPEED = f32(8.0)
    MAX_TORQUE = f32(2.0)
    DT = f32(0.05)
    PI = f32(np.pi)
    TWO_PI = f32(2 * np.pi)

    def __init__(self, n_envs, seed=0, max_steps=200, gid_offset=0):
        self.n = n_envs
        self.seed = seed
        self.max_steps = max_steps
        self.gid = np.arange(n_envs, dtype=np.int64) + gid_offset
        self.state = np.zeros((2, n_envs), dtype=f32)
        self.steps = np.zeros(n_envs, dtype=np.int32)
        self.episode = np.zeros(n_envs, dtype=np.int64)
        self.terminated = np.zeros(n_envs, dtype=bool)
        self.truncated = np.zeros(n_envs, dtype=bool)
        self.reset_all()

    def reset_idx(self, idx):
        xs = philox.philox4x32(self.gid[idx], self.episode[idx], 0, philox.TAG_RESET, self.seed)
        self.state[0, idx] = -self.PI + self.TWO_PI * philox.u01_f32(xs[0])
        self.state[1, idx] = f32(-1.0) + f32(2.0) * philox.u01_f32(xs[1])
        self.episode[idx] += 1
        self.steps[idx] = 0
        self.truncated[idx] = False

    def reset_all(self):
        self.reset_idx(np.arange(self.n))

    def obs(self):
        s, c = _sincos32(self.state[0])
        return np.stack([c, s, self.state[1]], axis=1).astype(f32)

    def step(self, actions):
        """actions: (n, 1) or (n,) fp32 torques (env-space; clipped again like Gymnasium)."""
        u = np.clip(np.asarray(actions, dtype=f32).reshape(self.n), -self.MAX_TORQUE, self.MAX_TORQUE)
        th, thdot = self.state[0], self.state[1]
        xp = th + self.PI
        an = (xp - self.TWO_PI * np.floor(xp / self.TWO_PI)) - self.PI
        cost = ((an * an) + (f32(0.1) * (thdot * thdot))) + (f32(0.001) * (u * u))
        sinth, _ = _sincos32(th)
        newthdot = thdot + (((f32(15.0) * sinth) + (f32(3.0) * u)) * self.DT)
        newthdot = np.clip(newthdot, -self.MAX_SPEED, self.MAX_SPEED).astype(f32)
        newth = th + newthdot * self.DT
        self.state = np.stack([newth, newthdot]).astype(f32)
        self.steps += 1
        self.truncated = self.steps >= self.max_steps
        return (-cost).astype(f32)


class SyntheticBatch:
    """Synthetic env for the rollout-only sweep: obs = Philox U(-1,1)^D,
    reward = U(0,1), terminates w.p. 1/200, truncates at 500. Action is ignored."""
    kind = "synthetic"
    act_kind = "discrete"
    n_actions = 2

    def __init__(self, n_envs, obs_dim, seed=0, max_steps=500, gid_offset=0, act_start=1):
        self.n = n_envs
        self.obs_dim = obs_dim
        self.seed = seed
        self.max_steps = max_steps
        self.act_start = act_start
        self.gid = np.arange(n_envs, dtype=np.int64) + gid_offset
        self.steps = np.zeros(n_envs, dtype=np.int32)
        self.life = np.zeros(n_envs, dtype=np.int64)
        self.episode = np.zeros(n_envs, dtype=np.int64)
        self.terminated = np.zeros(n_envs, dtype=bool)
        self.truncated = np.zeros(n_envs, dtype=bool)

    def reset_idx(self, idx):
        self.episode[idx] += 1
        self.steps[idx] = 0
        self.terminated[idx] = False
        self.truncated[idx] = False

    def reset_all(self):
        self.reset_idx(np.arange(self.n))

    def obs(self):
        out = np.zeros((self.n, self.obs_dim), dtype=f32)
        for b in range((self.obs_dim + 3) // 4):
            xs = philox.philox4x32(self.gid, self.life, b, philox.TAG_SYN_OBS, self.seed)
            for j in range(4):
                if 4 * b + j < self.obs_dim:
                    out[:, 4 * b + j] = f32(-1.0) + f32(2.0) * philox.u01_f32(xs[j])
        return out

    def step(self, actions):
        xs = philox.philox4x32(self.gid, self.life, 0, philox.TAG_SYN_DYN, self.seed)
        r = philox.u01_f32(xs[0])
        self.terminated = philox.u01_f32(xs[1]) < f32(1.0 / 200.0)
        self.life += 1
        self.steps += 1
        self.truncated = self.steps >= self.max_steps
        return r


class ScalingBatch:
    """ScalingWrapperEnv (environment_wrappers/scalingWrapperEnv.jl:14-49,72-120) around every env of a batch: observations
    of a Box space are mapped to [-1, 1] with the pre-computed factors `scale = 2 / (high - low)`, `offset = low` as
    `(x - offset) * scale - 1` (:72-75), actions come in [-1, 1] and are mapped back with `(a + 1) / scale + offset` (:77-80)
    before the wrapped env's act!.  Everything else is forwarded (:122-132).  fp32 like the spaces' eltype."""

    def __init__(self, batch, obs_low, obs_high, act_low=None, act_high=None):
        self.b = batch
        self.orig_obs_low, self.orig_obs_high = np.asarray(obs_low, dtype=f32), np.asarray(obs_high, dtype=f32)
        self.orig_act_low = np.asarray(batch.act_low if act_low is None else act_low, dtype=f32)
        self.orig_act_high = np.asarray(batch.act_high if act_high is None else act_high, dtype=f32)
        with np.errstate(divide="ignore"):
            self.obs_scale = (f32(2.0) / (self.orig_obs_high - self.orig_obs_low)).astype(f32)       # :37-38
            self.act_scale = (f32(2.0) / (self.orig_act_high - self.orig_act_low)).astype(f32)       # :42-43
        self.obs_offset, self.act_offset = self.orig_obs_low, self.orig_act_low
        self.act_low = -np.ones_like(self.orig_act_low)                                               # :30-34
        self.act_high = np.ones_like(self.orig_act_high)

    def __getattr__(self, name):            # n, kind, obs_dim, act_kind, act_dim, steps, terminated, truncated, ...
        return getattr(self.b, name)

    def scale_obs(self, x):
        return (((np.asarray(x, dtype=f32) - self.obs_offset).astype(f32) * self.obs_scale).astype(f32) - f32(1.0)).astype(f32)

    def unscale_action(self, a):
        a = np.asarray(a, dtype=f32).reshape(self.b.n, -1)
        return (((a + f32(1.0)).astype(f32) / self.act_scale).astype(f32) + self.act_offset).astype(f32)

    def obs(self):
        return self.scale_obs(self.b.obs())

    def step(self, actions):
        return self.b.step(self.unscale_action(actions))

    def reset_idx(self, idx):
        self.b.reset_idx(idx)

    def reset_all(self):
        self.b.reset_all()


PENDULUM_OBS_LOW = np.array([-1.0, -1.0, -8.0], dtype=f32)       # Pendulum-v1 observation space (cos, sin, angular velocity)
PENDULUM_OBS_HIGH = np.array([1.0, 1.0, 8.0], dtype=f32)


class ParallelEnv:
    """Vector env with auto-reset. environment_wrappers/multithreadedParallelEnv.jl:47-74:
    reward; flags read BEFORE reset; terminal_observation stored iff truncated; reset iff
    terminated or truncated; observe after the step returns the post-reset obs."""

    def __init__(self, batch):
        self.b = batch
        self.n = batch.n

    def reset(self):
        self.b.reset_all()

    def observe(self):
        return self.b.obs()

    def act(self, actions):
        rewards = self.b.step(actions)
        term = self.b.terminated.copy()
        trunc = self.b.truncated.copy()
        info = {"terminal_observation": None, "episode_r": None, "episode_l": None}
        if trunc.any():
            info["terminal_observation"] = self.b.obs()  # rows valid where trunc
        done = term | trunc
        if done.any():
            self.b.reset_idx(np.nonzero(done)[0])
        return rewards, term, trunc, info


class MonitorWrapper:
    """environment_wrappers/monitorWrapperEnv.jl:36-70."""

    def __init__(self, env, stats_window=100):
        self.env = env
        self.n = env.n
        self.ep_ret = np.zeros(env.n, dtype=f32)
        self.ep_len = np.zeros(env.n, dtype=np.int64)
        self.returns = deque(maxlen=stats_window)
        self.lengths = deque(maxlen=stats_window)
        self.total_episodes = 0

    def reset(self):
        self.env.reset()
        self.ep_ret[:] = 0
        self.ep_len[:] = 0

    def observe(self):
        return self.env.observe()

    def act(self, actions):
        rewards, term, trunc, info = self.env.act(actions)
        self.ep_ret += rewards
        self.ep_len += 1
        done = term | trunc
        info["episode_r"] = np.where(done, self.ep_ret, f32(0)).astype(f32)
        info["episode_l"] = np.where(done, self.ep_len, 0)
        for i in np.nonzero(done)[0]:
            self.returns.append(self.ep_ret[i])
            self.lengths.append(int(self.ep_len[i]))
            self.total_episodes += 1
            self.ep_ret[i] = 0
            self.ep_len[i] = 0
        return rewards, term, trunc, info

    def log_stats(self):
        if len(self.returns) == 0:
            return None
        return float(np.mean(np.array(self.returns, dtype=f32))), float(np.mean(self.lengths))


class RunningMeanStd:
    """environment_wrappers/normalizeWrapperEnv.jl:8-50. mean 0 / var 1 / count 0 at start;
    the first update overwrites; later updates are the parallel (Chan) merge; batch variance
    is the population variance."""

    def __init__(self, shape):
        self.mean = np.zeros(shape, dtype=f32)
        self.var = np.ones(shape, dtype=f32)
        self.count = 0

    def update(self, batch):
        """batch: (..., n) — statistics over the last axis (normalizeWrapperEnv.jl:21-26)."""
        batch = np.asarray(batch, dtype=f32)
        bm = batch.mean(axis=-1, dtype=f32)
        bv = batch.var(axis=-1, dtype=f32)
        self.update_from_moments(bm, bv, batch.shape[-1])

    def update_from_moments(self, bm, bv, bc):
        bm = np.asarray(bm, dtype=f32)
        bv = np.asarray(bv, dtype=f32)
        if self.count == 0:
            self.mean = bm.reshape(self.mean.shape).copy()
            self.var = bv.reshape(self.var.shape).copy()
            self.count = bc
        else:
            delta = bm.reshape(self.mean.shape) - self.mean
            total = self.count + bc
            new_mean = self.mean + delta * f32(bc) / f32(total)
            m_a = self.var * f32(self.count)
            m_b = bv.reshape(self.var.shape) * f32(bc)
            m2 = m_a + m_b + delta ** 2 * f32(self.count) * f32(bc) / f32(total)
            self.mean = new_mean.astype(f32)
            self.var = (m2 / f32(total)).astype(f32)
            self.count = total


class NormalizeWrapper:
    """environment_wrappers/normalizeWrapperEnv.jl:52-197."""

    def __init__(self, env, obs_dim, training=True, norm_obs=True, norm_reward=True,
                 clip_obs=10.0, clip_reward=10.0, gamma=0.99, epsilon=1e-8):
        self.env = env
        self.n = env.n
        self.obs_rms = RunningMeanStd((obs_dim,))
        self.ret_rms = RunningMeanStd(())
        self.returns = np.zeros(env.n, dtype=f32)
        self.training = training
        self.norm_obs = norm_obs
        self.norm_reward = norm_reward
        self.clip_obs = f32(clip_obs)
        self.clip_reward = f32(clip_reward)
        self.gamma = f32(gamma)
        self.epsilon = f32(epsilon)
        self.old_obs = None
        self.old_rewards = None

    def reset(self):  # :111-121 — stats are NOT updated on reset
        self.env.reset()
        self.old_obs = self.env.observe()
        self.returns[:] = 0

    def normalize_obs(self, obs):  # :174-186
        if not self.norm_obs:
            return obs
        o = (obs - self.obs_rms.mean) / np.sqrt(self.obs_rms.var + self.epsilon)
        return np.clip(o, -self.clip_obs, self.clip_obs).astype(f32)

    def observe(self):  # :123-137 — stats updated on EVERY observe while training
        obs = self.env.observe()
        self.old_obs = obs.copy()
        if self.training and self.norm_obs:
            self.obs_rms.update(obs.T)  # (obs_dim, n) batch, stats over envs
        return self.normalize_obs(obs)

    def act(self, actions):  # :139-165
        rewards, term, trunc, info = self.env.act(actions)
        self.old_rewards = rewards.copy()
        if self.training and self.norm_reward:
            self.returns = (self.returns * self.gamma + rewards).astype(f32)  # :167-171
            self.ret_rms.update(self.returns.reshape(1, -1))
        if self.norm_reward:  # :188-197 (no mean subtraction)
            rewards = np.clip(rewards / np.sqrt(self.ret_rms.var + self.epsilon),
                              -self.clip_reward, self.clip_reward).astype(f32)
        done = term | trunc
        self.returns[done] = 0
        if trunc.any() and info.get("terminal_observation") is not None:
            info["terminal_observation"] = self.normalize_obs(info["terminal_observation"])
        return rewards, term, trunc, info

    # monitor pass-through
    def log_stats(self):
        return self.env.log_stats() if hasattr(self.env, "log_stats") else None
