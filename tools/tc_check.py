"""Tensor-core loss/grad kernel vs the CUDA-core kernel vs the oracle, per parameter block."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dril_b200 as D  # noqa: E402
from oracle import policy as OP, ppo as OO  # noqa: E402

f32 = np.float32
spec = OP.PolicySpec(4, [64, 64], "discrete", 2, act_start=1)
rng = np.random.default_rng(0)
flat = (OP.init_params(spec, seed=3) + rng.normal(size=spec.n_params()).astype(f32) * 0.05).astype(f32)
p = D.DevicePolicy(D.Context.default(), 4, [64, 64], D.Discrete(2, 1))
p.set_params(flat)
blocks, off = [], 0
for net in ("actor", "critic"):
    for li, (i, o) in enumerate(spec.layer_dims(0 if net == "actor" else 1)):
        blocks.append((f"{net}.W{li}", off, off + i * o)); off += i * o
        blocks.append((f"{net}.b{li}", off, off + o)); off += o
for B in [int(x) for x in (sys.argv[1:] or ["128", "1000", "4097"])]:
    obs = rng.normal(size=(B, 4)).astype(f32)
    actions = rng.integers(1, 3, (B, 1))
    v0, lp0, _ = OP.evaluate_actions(spec, flat, obs, actions)
    old_lp = (lp0 + rng.normal(size=B).astype(f32) * 0.2).astype(f32)
    old_v = (v0 + rng.normal(size=B).astype(f32) * 0.3).astype(f32)
    adv, ret = rng.normal(size=B).astype(f32), rng.normal(size=B).astype(f32)
    alg = D.PPO(ent_coef=0.01)
    cfg = OO.PPOConfig(ent_coef=0.01)
    eloss, estats, eg = OO.ppo_loss_and_grads(spec, flat, obs, actions, adv, ret, old_lp, old_v, cfg)
    res = {}
    for tc in (0, 1):
        D.set_option("tc", tc)
        res[tc] = p.loss_grad(obs, actions, adv, ret, old_lp, old_v, alg.hyper())
    D.set_option("tc", 0)
    rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-30))
    print(f"B={B}: loss oracle {eloss:.6f} ffma {res[0][0]:.6f} tc {res[1][0]:.6f}")
    print("   stats tc   ", {k: round(v, 6) for k, v in res[1][1].items()})
    print("   stats ffma ", {k: round(v, 6) for k, v in res[0][1].items()})
    print(f"   grad relerr vs oracle: ffma {rel(res[0][2], eg):.2e}  tc {rel(res[1][2], eg):.2e}")
    for name, a, b in blocks:
        print(f"      {name:10s} ffma {rel(res[0][2][a:b], eg[a:b]):.2e}  tc {rel(res[1][2][a:b], eg[a:b]):.2e}")
    a, b = blocks[0][1], blocks[0][2]
    gt, ge = res[1][2][a:b].reshape(4, 64), eg[a:b].reshape(4, 64)
    for d in range(4):
        print(f"      actor.W0 row d={d}: relerr {rel(gt[d], ge[d]):.2e}  tc[:4] {gt[d][:4]}  oracle[:4] {ge[d][:4]}")
