"""Phase timeline of the tensor-core loss/grad kernel (CTA 0). Needs a trace build:
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
for k in range(3):
    L.check(ctx.lib.dril_ppo_iteration_async(env.h, agent.device.h, buf.h, C.byref(hyper), alg.epochs, alg.batch_size, 1, k))
ctx.synchronize()
out = (C.c_longlong * (2 * 32 * 16))()
ctx.lib.dril_debug_tc_trace.argtypes = [C.c_void_p]
L.check(ctx.lib.dril_debug_tc_trace(out))
tr = np.array(out).reshape(2, 32, 16)
t0 = tr[tr > 0].min()
names = ["start", "L0 done", "G1 issued", "G1 done", "head done", "acquired", "staged", "reduced", "images", "G2G3 issued", "G2G3 done",
         "dz0 staged", "reduced/release"]
for ps, nm in ((0, "actor"), (16, "critic")):
    for it in range(4):
        for g in range(2):
            row = tr[g, ps + it, :13]
            if row[0] == 0:
                continue
            rel = row - t0
            print(f"{nm} g{g} tile{it}: start {rel[0]:7d} | " + " ".join(f"{names[i + 1]}+{row[i + 1] - row[i]}" for i in range(12)) +
                  f" | total {row[12] - row[0]}")

tl = tr[1, 31, :11]
print("tail (CTA 0): wait for all CTAs + barrier 1 = %d, slice reduction = %d, block sum + barrier 2 = %d, norm/stats/Adam = %d cycles" %
      (tl[1] - tl[0], tl[2] - tl[1], tl[3] - tl[2], tl[4] - tl[3]))
print("  reduction: index loads %d, partial loads + sums %d, smem exchange %d, rest %d | after barrier 2: sq sum %d, stop flag %d, accumulators %d, Adam %d" %
      (tl[5] - tl[1], tl[6] - tl[5], tl[7] - tl[6], tl[2] - tl[7], tl[8] - tl[3], tl[9] - tl[8], tl[10] - tl[9], tl[4] - tl[10]))
