"""Time the fused rollout + the update for a workload under the current DRIL_* tuning env vars."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dril_b200 as D  # noqa: E402
from dril_b200 import _lib as L  # noqa: E402

kind, n, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
hidden = [int(x) for x in sys.argv[4].split(",")]
norm = len(sys.argv) > 5 and sys.argv[5] == "norm"
env = D.CudaBatchedEnv(kind, n, seed=0, monitor_window=100, normalize=D.NormalizeConfig() if norm else None)
layer = D.ActorCriticLayer(env.observation_space(), env.action_space(), hidden_dims=hidden)
alg = D.PPO(n_steps=T, batch_size=T * n // 4, epochs=4)
agent = D.Agent(layer, alg, rng=np.random.default_rng(0))
buf = D.RolloutBuffer(env.observation_space(), env.action_space(), alg.gae_lambda, alg.gamma, T, n)
ctx = agent.ctx
hyper = alg.hyper()
for it in range(3):
    L.check(ctx.lib.dril_ppo_iteration_async(env.h, agent.device.h, buf.h, C.byref(hyper), alg.epochs, alg.batch_size, 1, it * 4))
ctx.synchronize()
ctx.set_profiling(True)
ctx.reset_profile()
K = 5
for it in range(K):
    L.check(ctx.lib.dril_ppo_iteration_async(env.h, agent.device.h, buf.h, C.byref(hyper), alg.epochs, alg.batch_size, 1, 12 + it * 4))
ctx.synchronize()
prof = ctx.profile()
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("DRIL_"))
print(f"[{kind} n={n} T={T} hidden={hidden} norm={norm}] {tag}: " +
      " ".join(f"{k}={v[0] / K:.3f}ms" for k, v in prof.items() if v[1]))
