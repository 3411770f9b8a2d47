"""Synthetic rollout-only sweep (BASELINE.json configs[4]): fused rollout + GAE throughput vs the
HBM roofline for n_envs x obs_dim, hidden [64,64] and a thin [8] policy. Prints one CSV row per point."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dril_b200 as D  # noqa: E402
from dril_b200 import _lib as L  # noqa: E402

peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
T = 128
print("env,n_envs,obs_dim,hidden,rollout_ms,gae_ms,env_steps_per_s,alg_bytes_per_step,achieved_GBps,frac_of_hbm_peak,fwd_GFLOPs")
points = [("synthetic", n, d, h) for h in ([64, 64], [8]) for d in (4, 16, 64) for n in (1 << 10, 1 << 14, 1 << 17, 1 << 20)
          if n * d <= (1 << 24)]
points += [("cartpole", n, 4, [64, 64]) for n in (1 << 10, 1 << 12, 1 << 14, 1 << 16, 1 << 18)]
ctx = D.Context.default()
for kind, n, d, hidden in points:
    env = D.CudaBatchedEnv(kind, n, obs_dim=d, seed=0, monitor_window=100)
    layer = D.ActorCriticLayer(env.observation_space(), env.action_space(), hidden_dims=hidden)
    alg = D.PPO(n_steps=T)
    agent = D.Agent(layer, alg, rng=np.random.default_rng(0))
    buf = D.RolloutBuffer(env.observation_space(), env.action_space(), alg.gae_lambda, alg.gamma, T, n)
    for _ in range(2):
        D.collect_rollout(buf, agent, alg, env)
    ctx.set_profiling(True); ctx.reset_profile()
    K = 3
    for _ in range(K):
        D.collect_rollout(buf, agent, alg, env)
    prof = ctx.profile(); ctx.set_profiling(False)
    ro, gae = prof["rollout"][0] / K, prof["gae"][0] / K
    steps = n * T
    bytes_per = 4 * d + 34
    flops = 2 * sum(i * o for net in (0, 1) for (i, o) in layer.layer_dims(net))
    gbps = steps * bytes_per / ((ro + gae) * 1e-3) / 1e9
    print(f"{kind},{n},{d},{'x'.join(map(str, hidden))},{ro:.3f},{gae:.3f},{steps / ((ro + gae) * 1e-3):.4g},{bytes_per},{gbps:.1f},"
          f"{gbps / peak:.4f},{steps * flops / (ro * 1e-3) / 1e9:.0f}")
    buf.close(); env.close(); agent.device.close()
