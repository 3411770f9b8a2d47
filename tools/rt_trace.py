"""Per-step phase timeline of the tensor-core rollout kernel (CTA 0). Needs a trace build:
   DRIL_NVCC_EXTRA=-DTC_TRACE python dril.jl_b200/build.py --force"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dril_b200 as D
from dril_b200 import _lib as L
n, T = 4096, 128
env = D.CudaBatchedEnv("cartpole", n, seed=0, monitor_window=100)
layer = D.ActorCriticLayer(env.observation_space(), env.action_space(), hidden_dims=[64, 64])
alg = D.PPO(n_steps=T, batch_size=T * n // 4, epochs=4)
agent = D.Agent(layer, alg, rng=np.random.default_rng(0))
buf = D.RolloutBuffer(env.observation_space(), env.action_space(), alg.gae_lambda, alg.gamma, T, n)
ctx = agent.ctx; hyper = alg.hyper()
for k in range(2):
    L.check(ctx.lib.dril_ppo_iteration_async(env.h, agent.device.h, buf.h, C.byref(hyper), alg.epochs, alg.batch_size, 1, k))
ctx.synchronize()
out = (C.c_longlong * (4 * 8 * 8))()
ctx.lib.dril_debug_rt_trace.argtypes = [C.c_void_p]
L.check(ctx.lib.dril_debug_rt_trace(out))
tr = np.array(out).reshape(4, 8, 8)
names = ["L0", "sync1+issue", "window", "wait+copy", "sync2+head", "sync3", "decide", "select+reset"]
for fq in range(4):
    for st in range(2, 5):
        row = tr[fq, st, :8]
        nxt = tr[fq, st + 1, 0]
        seg = [row[i + 1] - row[i] for i in range(7)] + [nxt - row[7]]
        print(f"warp{[0, 5, 6, 7][fq]} step{8 + st}: " + " ".join(f"{names[i]}={seg[i]}" for i in range(8)) + f" | step total {nxt - row[0]}")
