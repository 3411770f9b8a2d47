"""Phase timeline of the features-on-lanes loss/grad kernel (CTA 0, thread 0 of each group). Needs a trace build:
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
names = ["record wait", "A (H0, images)", "st wait+sync+G1 issue", "G1 wait", "B1 (H1)", "sync", "out partials", "sync", "head", "sync",
         "B2 (dZ1, images)", "sync+G2G3 issue", "G2 wait", "C (dZ0 sums)", "G3 wait", "flush+sync"]
t0 = tr[tr > 0].min()
tot = np.zeros(16)
cnt = 0
for it in range(8):
    for g in range(2):
        row = tr[g, it]
        nxt = tr[g, it + 1][0]
        if row[0] == 0 or row[15] == 0:
            continue
        d = [row[i + 1] - row[i] for i in range(15)] + [nxt - row[15] if nxt > row[15] else 0]
        print(f"g{g} tile{it}: start {row[0] - t0:7d} | " + " ".join(f"{d[i]}" for i in range(16)) + f" | total {sum(d)}")
        if nxt > row[15]:
            tot += np.array(d); cnt += 1
print("mean over %d tiles:" % cnt)
for i in range(16):
    print(f"  {names[i]:24s} {tot[i] / max(cnt, 1):8.0f}")
print(f"  {'total':24s} {tot.sum() / max(cnt, 1):8.0f}")
tl = tr[1, 31, :11]
print("tail (CTA 0): wait for all CTAs + barrier 1 = %d, slice reduction = %d, block sum + barrier 2 = %d, norm/stats/Adam = %d cycles" %
      (tl[1] - tl[0], tl[2] - tl[1], tl[3] - tl[2], tl[4] - tl[3]))
k = tr[0, 30]
print("kernel (CTA 0, cycles): prologue %d | loop until g0 done %d, g1 done %d | sync %d | end-of-pass reductions %d | tail %d | total %d" %
      (k[1] - k[0], k[4] - k[1], k[5] - k[1], k[2] - max(k[4], k[5]), k[3] - k[2], k[6] - k[3], k[6] - k[0]))
print("end of pass: smem writes + warp sums %d | sync %d | global writes %d | sync %d | dealloc %d" % (k[7] - k[2], k[8] - k[7], k[9] - k[8], k[10] - k[9], k[3] - k[10]))
tl = tr[1, 31, :11]
print("  tail reduction: index loads %d, partial loads + sums %d, smem exchange %d, rest %d | after barrier 2: sq sum %d, stop flag %d, accumulators %d, Adam %d" %
      (tl[5] - tl[1], tl[6] - tl[5], tl[7] - tl[6], tl[2] - tl[7], tl[8] - tl[3], tl[9] - tl[8], tl[10] - tl[9], tl[4] - tl[10]))
st, en = tr[0, 28], tr[0, 29]
order = np.argsort(st)
prev = None
for i in order:
    if st[i] == 0:
        continue
    print("launch slot %2d: duration %6.1f us%s" % (i, (en[i] - st[i]) / 1e3, "" if prev is None else ", gap since previous end %6.1f us" % ((st[i] - prev) / 1e3)))
    prev = en[i]
