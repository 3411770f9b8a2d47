"""Host-side cost of enqueueing one PPO iteration (empty launch queue) vs its device time."""
import ctypes as C, os, sys, time
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
def it(k):
    L.check(ctx.lib.dril_ppo_iteration_async(env.h, agent.device.h, buf.h, C.byref(hyper), alg.epochs, alg.batch_size, 1, k))
st = L.IterStats()
for k in range(3): it(k)
ctx.synchronize()
for rep in range(5):
    ctx.synchronize()
    t0 = time.perf_counter(); it(10 + rep); t1 = time.perf_counter()
    L.check(ctx.lib.dril_iteration_result(agent.device.h, C.byref(st))); t2 = time.perf_counter()
    print(f"enqueue {1e3 * (t1 - t0):.3f} ms, until result {1e3 * (t2 - t0):.3f} ms, device rollout+update {st.rollout_ms + st.update_ms:.3f} ms")
