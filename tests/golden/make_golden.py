"""Generates the committed golden fixtures `tests/golden/*.npz`.

The reference is Julia (no `julia` in this image or on the GPU box) and its own tests hold closed-form values, not
golden files, so there are two kinds of fixture here:

* `ref_*`  — the literal inputs and expected values of the reference's own tests for this path
  (test/test_gae.jl, test/test_buffers.jl, test/test_normalize_wrapper.jl, test/test_distributions.jl), written out as
  arrays.  The expected values are evaluated from the closed forms stated in those tests in float64, NOT by running the
  oracle: they pin the oracle (tests/test_golden_cpu.py) and the CUDA path (tests/test_gpu_golden.py) alike.
* `orc_*`  — seeded input/output vectors produced by the oracle (`oracle/`), frozen so that (a) a later edit of the oracle
  that changes results is caught on CPU and (b) the GPU parity tests have expected values that do not depend on
  executing the oracle on the GPU box.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import envs as OE, philox as OPH, policy as OP, ppo as OO  # noqa: E402

f32 = np.float32


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}.npz  {os.path.getsize(path) / 1024:.1f} KB  {sorted(arrays)}")


# --------------------------------------------------------------------------------------------------
# ref_*: the reference's own test vectors
# --------------------------------------------------------------------------------------------------
def _gae_sum64(r, v, gamma, lam, boot):
    """float64 GAE by its definition A_t = sum_k (gamma*lambda)^k delta_{t+k} (the comment block of
    test/test_gae.jl:38-46), with delta_T = r_T + gamma*boot - V_T (boot = 0 when terminated) — deliberately not the
    backward recursion the oracle and the kernel use."""
    r = np.asarray(r, np.float64); v = np.asarray(v, np.float64); T = len(r)
    vn = np.append(v[1:], boot)
    delta = r + gamma * vn - v
    return np.array([sum((gamma * lam) ** k * delta[t + k] for k in range(T - t)) for t in range(T)])


def ref_gae():
    """The scenarios of the reference's GAE tests, with their own parameters:
    A  test/test_gae.jl:1-71     8 steps, reward 1 at the end, V = 0.5, gamma .99, lambda .95, terminated
    B  test/test_gae.jl:73-115   4 steps, V = 0.3, five (gamma, lambda) pairs incl. Monte-Carlo and TD(0)
    C  test/test_gae.jl:176-220  32 steps = 4 episodes of 8, gamma = lambda = 1, V = 0: advantage = return = 1 everywhere
    D  test/test_buffers.jl:60-115  6 steps, gamma .9, lambda .8, V = 0.7: terminated vs truncated with bootstrap 0.2"""
    out = {}
    f = lambda x: np.float64(np.float32(x))           # the tests pass Float32 literals
    rA = np.zeros(8); rA[-1] = 1
    out["A_rewards"] = rA.astype(f32); out["A_values"] = np.full(8, 0.5, f32)
    out["A_gamma_lambda"] = np.array([f(0.99), f(0.95)])
    out["A_adv"] = _gae_sum64(rA, np.full(8, 0.5), f(0.99), f(0.95), 0.0)
    rB = np.zeros(4); rB[-1] = 1
    pairs = [(0.95, 0.9), (0.99, 0.95), (1.0, 1.0), (0.9, 0.0), (0.8, 0.5)]
    out["B_rewards"] = rB.astype(f32); out["B_values"] = np.full(4, 0.3, f32)
    out["B_gamma_lambda"] = np.array([[f(g), f(l)] for g, l in pairs])
    out["B_adv"] = np.array([_gae_sum64(rB, np.full(4, f(0.3)), f(g), f(l), 0.0) for g, l in pairs])
    rC = np.zeros(32); rC[7::8] = 1
    out["C_rewards"] = rC.astype(f32); out["C_values"] = np.zeros(32, f32); out["C_term"] = rC.astype(bool)
    out["C_adv"] = np.ones(32)
    rD = np.zeros(6); rD[-1] = 1
    out["D_rewards"] = rD.astype(f32); out["D_values"] = np.full(6, 0.7, f32)
    out["D_gamma_lambda"] = np.array([f(0.9), f(0.8)]); out["D_boot"] = np.array(f(0.2))
    out["D_adv_terminated"] = _gae_sum64(rD, np.full(6, f(0.7)), f(0.9), f(0.8), 0.0)
    out["D_adv_truncated"] = _gae_sum64(rD, np.full(6, f(0.7)), f(0.9), f(0.8), f(0.2))
    save("ref_gae", **out)


def ref_running_mean_std():
    """test/test_normalize_wrapper.jl:3-70: feeding batches one after another must equal the moments (population
    variance) of the concatenation."""
    rng = np.random.default_rng(11)
    batches = [rng.normal(loc=m, scale=s, size=(n, 3)) .astype(f32) for m, s, n in ((0, 1, 7), (3, 2, 5), (-1, .5, 11), (10, 4, 1))]
    allx = np.concatenate(batches).astype(np.float64)
    save("ref_running_mean_std", **{f"batch{i}": b for i, b in enumerate(batches)}, mean=allx.mean(0), var=allx.var(0),
         count=np.array(allx.shape[0]))


def ref_distributions():
    """test/test_distributions.jl:1-39 (DiagGaussian logpdf/entropy vs the textbook formulas),
    :94-118 (Categorical logpdf = log p[a], entropy = -sum p log p)."""
    rng = np.random.default_rng(5)
    mean = rng.normal(size=(6, 3)); log_std = rng.normal(size=3) * 0.3; x = rng.normal(size=(6, 3))
    var = np.exp(2 * log_std)
    logpdf = (-0.5 * ((x - mean) ** 2 / var + 2 * log_std + np.log(2 * np.pi))).sum(1)
    entropy = (0.5 * (1 + np.log(2 * np.pi)) + log_std).sum()
    logits = rng.normal(size=(6, 4))
    p = np.exp(logits - logits.max(1, keepdims=True)); p /= p.sum(1, keepdims=True)
    actions = rng.integers(0, 4, 6)
    save("ref_distributions", g_mean=mean.astype(f32), g_log_std=log_std.astype(f32), g_x=x.astype(f32),
         g_logpdf=logpdf, g_entropy=np.array(entropy), c_logits=logits.astype(f32), c_actions=actions,
         c_logpdf=np.log(p[np.arange(6), actions]), c_entropy=-(p * np.log(p)).sum(1))


# --------------------------------------------------------------------------------------------------
# orc_*: frozen oracle vectors
# --------------------------------------------------------------------------------------------------
def orc_philox():
    c = np.arange(8, dtype=np.uint32)
    out = OPH.philox4x32(c, c * 7 + 1, np.uint32(3), np.uint32(0x5EED), 0xDEADBEEF12345678)
    gid = np.arange(5)
    save("orc_philox", c0=c, out=np.stack(out), u64=OPH.sample_uniform64(gid, 9, 42), normals=OPH.normals(gid, 9, 3, 42),
         perm=OPH.feistel_permute(np.arange(1000), 1000, OPH.feistel_keys(3, 1, 77)))


def _replay(batch, actions):
    env = OE.ParallelEnv(batch)
    obs = [env.observe()]
    R, TE, TR, TO = [], [], [], []
    for a in actions:
        r, te, tr, info = env.act(a)
        R.append(r); TE.append(te); TR.append(tr)
        to = np.zeros_like(obs[0])
        for i in np.nonzero(tr)[0]:
            to[i] = info["terminal_observation"][i]
        TO.append(to); obs.append(env.observe())
    return dict(obs=np.stack(obs), rewards=np.stack(R), term=np.stack(TE), trunc=np.stack(TR), terminal_obs=np.stack(TO))


def orc_env_replay():
    rng = np.random.default_rng(0)
    n, steps = 16, 60
    a = rng.integers(1, 3, (steps, n))
    save("orc_cartpole_replay", actions=a, seed=np.array(5), max_steps=np.array(25),
         **_replay(OE.CartPoleBatch(n, seed=5, max_steps=25), a))
    a = rng.uniform(-2.5, 2.5, (steps, n, 1)).astype(f32)
    save("orc_pendulum_replay", actions=a, seed=np.array(5), max_steps=np.array(20),
         **_replay(OE.PendulumBatch(n, seed=5, max_steps=20), a))


def orc_gae():
    rng = np.random.default_rng(17)
    T, N = 24, 10
    r = rng.normal(size=(T, N)).astype(f32); v = rng.normal(size=(T, N)).astype(f32)
    term = rng.random((T, N)) < 0.08; trunc = (rng.random((T, N)) < 0.08) & ~term
    boot = np.where(trunc, rng.normal(size=(T, N)), 0).astype(f32); last = rng.normal(size=N).astype(f32)
    adv, ret = OO.gae_timemajor(r, v, term, trunc, boot, last, 0.99, 0.95)
    save("orc_gae", rewards=r, values=v, term=term, trunc=trunc, boot=boot, last_values=last, gamma=np.array(0.99),
         gae_lambda=np.array(0.95), advantages=adv, returns=ret)


SPECS = {
    "cartpole": lambda: OP.PolicySpec(4, [64, 64], "discrete", 2, act_start=1),
    "pendulum": lambda: OP.PolicySpec(3, [32, 16], "continuous", 1, act_low=[-2], act_high=[2]),
}


def _params(spec, seed):
    rng = np.random.default_rng(seed)
    return (OP.init_params(spec, seed=seed) + rng.normal(size=spec.n_params()).astype(f32) * 0.05).astype(f32)


def orc_rollout():
    """Fused-rollout fixtures with replayed actions: CartPole + Monitor ([64,64], the tensor-core path) and
    Pendulum + Monitor + Normalize (training statistics) with a small [32,16] net."""
    for kind, n, T, ms, norm in (("cartpole", 32, 16, 12, False), ("pendulum", 24, 12, 8, True)):
        spec = SPECS[kind](); flat = _params(spec, 2)
        rng = np.random.default_rng(4)
        forced = rng.integers(1, 3, (T, n)) if kind == "cartpole" else (rng.normal(size=(T, n, 1)) * 1.5).astype(f32)
        b = OE.CartPoleBatch(n, seed=9, max_steps=ms) if kind == "cartpole" else OE.PendulumBatch(n, seed=9, max_steps=ms)
        env = OE.MonitorWrapper(OE.ParallelEnv(b))
        if norm:
            env = OE.NormalizeWrapper(env, spec.obs_dim)
        ob = OO.collect_rollout_timemajor(env, spec, flat, T, forced_actions=forced)
        adv, ret = OO.gae_timemajor(ob["rewards"], ob["values"], ob["term"], ob["trunc"], ob["boot"], ob["last_values"], 0.97, 0.9)
        extra = {}
        if norm:
            extra = dict(obs_mean=env.obs_rms.mean, obs_var=env.obs_rms.var, obs_count=np.array(env.obs_rms.count),
                         ret_mean=np.array(env.ret_rms.mean), ret_var=np.array(env.ret_rms.var), ret_count=np.array(env.ret_rms.count))
        save(f"orc_rollout_{kind}", params=flat, forced=forced, n_envs=np.array(n), n_steps=np.array(T), max_steps=np.array(ms),
             seed=np.array(9), gamma=np.array(0.97), gae_lambda=np.array(0.9), hidden=np.array(spec.hidden),
             obs=ob["obs"], rewards=ob["rewards"], values=ob["values"], logprobs=ob["logprobs"], term=ob["term"], trunc=ob["trunc"],
             boot=ob["boot"], last_values=ob["last_values"], episode_r=ob["episode_r"], episode_l=ob["episode_l"],
             advantages=adv, returns=ret, **extra)


def _minibatch(spec, flat, B, rng):
    obs = rng.normal(size=(B, spec.obs_dim)).astype(f32)
    if spec.act_kind == "discrete":
        actions = rng.integers(spec.act_start, spec.act_start + spec.act_n, (B, 1))
    else:
        actions = rng.normal(size=(B, spec.act_n)).astype(f32)
    v0, lp0, _ = OP.evaluate_actions(spec, flat, obs, actions)
    old_lp = (lp0 + rng.normal(size=B).astype(f32) * 0.2).astype(f32)
    old_v = (v0 + rng.normal(size=B).astype(f32) * 0.3).astype(f32)
    return obs, actions, rng.normal(size=B).astype(f32), rng.normal(size=B).astype(f32), old_lp, old_v


def orc_lossgrad():
    """PPO loss, statistics and the flat gradient on one minibatch (ppo.jl:365-407), two hyper-parameter sets;
    then three clip + Adam steps (ppo.jl:209-239) from the same parameters."""
    for kind, B in (("cartpole", 200), ("pendulum", 150)):
        spec = SPECS[kind](); flat = _params(spec, 3)
        rng = np.random.default_rng(B)
        mb = _minibatch(spec, flat, B, rng)
        out = dict(params=flat, hidden=np.array(spec.hidden), obs=mb[0], actions=mb[1], advantages=mb[2], returns=mb[3],
                   old_logprobs=mb[4], old_values=mb[5])
        hypers = [dict(ent_coef=0.01, clip_range_vf=None, normalize_advantage=True, vf_coef=0.5),
                  dict(ent_coef=0.02, clip_range_vf=0.2, normalize_advantage=False, vf_coef=0.7)]
        for i, h in enumerate(hypers):
            loss, stats, g = OO.ppo_loss_and_grads(spec, flat, *mb, OO.PPOConfig(**h))
            out[f"h{i}_loss"] = np.array(loss, np.float64)
            out[f"h{i}_grads"] = g
            out[f"h{i}_stat_names"] = np.array(sorted(stats))
            out[f"h{i}_stats"] = np.array([stats[k] for k in sorted(stats)], np.float64)
        save(f"orc_lossgrad_{kind}", **out)
    # Adam: fixed gradients (one below the clip threshold), lr 3e-4, eps 1e-5, max_grad_norm 0.5
    spec = SPECS["pendulum"](); flat = _params(spec, 0)
    rng = np.random.default_rng(0)
    opt = OO.Adam(flat.size, lr=3e-4)
    cur, gs, norms, ps = flat, [], [], []
    for it in range(3):
        g = (rng.normal(size=flat.size) * (0.001 if it == 1 else 0.05)).astype(f32)
        gc, norm = OO.clip_grads(g, 0.5)
        cur = opt.step(cur, gc)
        gs.append(g); norms.append(norm); ps.append(cur)
    save("orc_adam", params=flat, hidden=np.array(spec.hidden), grads=np.stack(gs), norms=np.array(norms, np.float64),
         params_after=np.stack(ps), m=opt.m, v=opt.v)


if __name__ == "__main__":
    ref_gae(); ref_running_mean_std(); ref_distributions()
    orc_philox(); orc_env_replay(); orc_gae(); orc_rollout(); orc_lossgrad()
