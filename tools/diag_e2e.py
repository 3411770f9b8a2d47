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
for k in range(3): it(k)
ctx.synchronize()
K = 20
for flush in (0, 1, 0, 1):
    ctx.synchronize(); t0 = time.perf_counter(); ctx.event_record(0)
    for k in range(K):
        if flush: ctx.flush_l2()
        it(10 + k)
    ctx.event_record(1); t_enq = time.perf_counter() - t0
    ms = ctx.event_elapsed_ms(0, 1)
    print(f"flush={flush}: {ms / K:.3f} ms/iter (device), enqueue took {t_enq * 1e3 / K:.3f} ms/iter on the host")
ctx.event_record(2)
for k in range(10): ctx.flush_l2()
ctx.event_record(3)
print("flush alone: %.3f ms" % (ctx.event_elapsed_ms(2, 3) / 10))
for rep in range(3):
    ctx.synchronize(); t0 = time.perf_counter()
    out = D.train(agent, env, alg, n * T * K)
    ctx.synchronize(); dt = time.perf_counter() - t0
    print(f"train() {K} iterations: {dt * 1e3 / K:.3f} ms/iter wall; timers {out[1]}")
