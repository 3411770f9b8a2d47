"""GPU parity against the committed golden fixtures (tests/golden/*.npz): the CUDA path through the C ABI is compared
with frozen expected values only — nothing under oracle/ is executed here.  Tolerances are north_star's:
(a) flags bit-exact, observations 1e-6 relative; (b) GAE 1e-5; (c) loss and gradients 1e-4 relative."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
f32 = np.float32


def gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


@pytest.fixture(scope="module")
def D():
    import __graft_entry__
    __graft_entry__.build()
    import dril_b200
    return dril_b200


def _gae(D, r, v, term, trunc, boot, last, gamma, lam):
    z = lambda a, dt: np.ascontiguousarray(np.asarray(a, dt).reshape(len(r), -1))
    return D.gae_raw(z(r, f32), z(v, f32), z(term, bool), z(trunc, bool), z(boot, f32), np.asarray(last, f32).reshape(-1),
                     float(gamma), float(lam))


def test_ref_gae_scenarios(D):
    """test/test_gae.jl:1-71,73-115,176-220 and test/test_buffers.jl:60-115 through dril_gae_raw."""
    g = gold("ref_gae")
    T = 8
    last = np.zeros(T, bool); last[-1] = True
    adv, ret = _gae(D, g["A_rewards"], g["A_values"], last, np.zeros(T), np.zeros(T), [0], *g["A_gamma_lambda"])
    np.testing.assert_allclose(adv[:, 0], g["A_adv"], atol=1e-5)
    np.testing.assert_allclose(ret[:, 0], g["A_adv"] + 0.5, atol=1e-5)
    last = np.zeros(4, bool); last[-1] = True
    for (gm, lm), exp in zip(g["B_gamma_lambda"], g["B_adv"]):
        adv, _ = _gae(D, g["B_rewards"], g["B_values"], last, np.zeros(4), np.zeros(4), [0], gm, lm)
        np.testing.assert_allclose(adv[:, 0], exp, atol=1e-5)
    adv, ret = _gae(D, g["C_rewards"], g["C_values"], g["C_term"], np.zeros(32), np.zeros(32), [0], 1.0, 1.0)
    np.testing.assert_allclose(adv[:, 0], g["C_adv"], atol=1e-5)
    np.testing.assert_allclose(ret[:, 0], g["C_adv"], atol=1e-5)
    last = np.zeros(6, bool); last[-1] = True
    gd, ld = g["D_gamma_lambda"]
    a_te, _ = _gae(D, g["D_rewards"], g["D_values"], last, np.zeros(6), np.zeros(6), [0], gd, ld)
    bootv = np.zeros(6, f32); bootv[-1] = g["D_boot"]
    a_tr, _ = _gae(D, g["D_rewards"], g["D_values"], np.zeros(6), last, bootv, [0], gd, ld)
    np.testing.assert_allclose(a_te[:, 0], g["D_adv_terminated"], atol=1e-5)
    np.testing.assert_allclose(a_tr[:, 0], g["D_adv_truncated"], atol=1e-5)
    # rollout cut in the middle of an episode: the bootstrap comes from last_values (trajectory.jl:65-70)
    a_cut, _ = _gae(D, g["D_rewards"], g["D_values"], np.zeros(6), np.zeros(6), np.zeros(6), [g["D_boot"]], gd, ld)
    np.testing.assert_allclose(a_cut[:, 0], g["D_adv_truncated"], atol=1e-5)


def test_orc_gae(D):
    g = gold("orc_gae")
    adv, ret = D.gae_raw(g["rewards"], g["values"], g["term"], g["trunc"], g["boot"], g["last_values"], float(g["gamma"]),
                         float(g["gae_lambda"]))
    np.testing.assert_allclose(adv, g["advantages"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(ret, g["returns"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("kind", ["cartpole", "pendulum"])
def test_orc_env_replay(D, kind):
    """north_star (a): replayed action sequences -> bit-exact flags, observations within 1e-6 relative."""
    g = gold(f"orc_{kind}_replay")
    n = g["actions"].shape[1]
    env = D.CudaBatchedEnv(kind, n, max_steps=int(g["max_steps"]), seed=int(g["seed"]))
    np.testing.assert_allclose(env.observe(), g["obs"][0], rtol=1e-6, atol=1e-7)
    for t, a in enumerate(g["actions"]):
        r, te, tr, infos = env.act(a)
        np.testing.assert_array_equal(te, g["term"][t])
        np.testing.assert_array_equal(tr, g["trunc"][t])
        np.testing.assert_allclose(r, g["rewards"][t], rtol=1e-6, atol=0)
        np.testing.assert_allclose(env.observe(), g["obs"][t + 1], rtol=1e-6, atol=1e-7)
        for i in np.nonzero(tr)[0]:
            np.testing.assert_allclose(infos[i]["terminal_observation"], g["terminal_obs"][t, i], rtol=1e-6, atol=1e-7)
    env.close()


@pytest.mark.parametrize("kind", ["cartpole", "pendulum"])
def test_orc_fused_rollout(D, kind):
    """Fused rollout + GAE with replayed actions: CartPole [64,64] (tensor-core rollout path) with Monitor, and
    Pendulum with Monitor + training NormalizeWrapperEnv (general path, running statistics)."""
    g = gold(f"orc_rollout_{kind}")
    n, T = int(g["n_envs"]), int(g["n_steps"])
    norm = kind == "pendulum"
    env = D.CudaBatchedEnv(kind, n, max_steps=int(g["max_steps"]), seed=int(g["seed"]), monitor_window=100,
                           normalize=D.NormalizeConfig() if norm else None)
    layer = D.ActorCriticLayer(env.observation_space(), env.action_space(), hidden_dims=[int(h) for h in g["hidden"]])
    alg = D.PPO(n_steps=T, gamma=float(g["gamma"]), gae_lambda=float(g["gae_lambda"]))
    agent = D.Agent(layer, alg, rng=np.random.default_rng(0))
    agent.set_parameters(g["params"])
    buf = D.RolloutBuffer(env.observation_space(), env.action_space(), alg.gae_lambda, alg.gamma, T, n)
    _, ok = D.collect_rollout(buf, agent, alg, env, forced_actions=g["forced"])
    assert ok
    flags = buf.download("flags")
    te, tr = (flags & 1).astype(bool), ((flags >> 1) & 1).astype(bool)
    np.testing.assert_array_equal(te, g["term"])
    np.testing.assert_array_equal(tr, g["trunc"])
    tol = dict(rtol=3e-5, atol=3e-5) if norm else dict(rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(buf.download("obs"), g["obs"], rtol=3e-5 if norm else 1e-6, atol=3e-5 if norm else 1e-7)
    np.testing.assert_allclose(buf.download("rewards"), g["rewards"], **(tol if norm else dict(rtol=1e-6, atol=0)))
    np.testing.assert_allclose(buf.download("values"), g["values"], **tol)
    np.testing.assert_allclose(buf.download("logprobs"), g["logprobs"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(buf.download("last_values"), g["last_values"], **tol)
    np.testing.assert_allclose(np.where(tr, buf.download("boot"), 0), g["boot"], **tol)
    np.testing.assert_allclose(buf.download("advantages"), g["advantages"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(buf.download("returns"), g["returns"], rtol=1e-4, atol=1e-4)
    done = te | tr
    assert done.any()
    np.testing.assert_allclose(buf.download("episode_r")[done], g["episode_r"][done], rtol=1e-5)
    np.testing.assert_array_equal(buf.download("episode_l")[done], g["episode_l"][done])
    if norm:
        s = env.norm_stats()
        assert s["obs_count"] == int(g["obs_count"]) and s["ret_count"] == int(g["ret_count"])
        np.testing.assert_allclose(s["obs_mean"], g["obs_mean"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(s["obs_var"], g["obs_var"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose([s["ret_mean"], s["ret_var"]], [g["ret_mean"], g["ret_var"]], rtol=1e-5, atol=1e-6)
    buf.close(); env.close()


def _policy(D, kind, hidden, flat):
    space = D.Discrete(2, 1) if kind == "cartpole" else D.Box([-2], [2])
    p = D.DevicePolicy(D.Context.default(), 4 if kind == "cartpole" else 3, [int(h) for h in hidden], space)
    p.set_params(np.ascontiguousarray(flat, f32))
    return p


@pytest.mark.parametrize("kind", ["cartpole", "pendulum"])
def test_orc_loss_and_gradients(D, kind):
    """north_star (c): loss and gradients on the same minibatch and parameters within 1e-4 relative."""
    g = gold(f"orc_lossgrad_{kind}")
    p = _policy(D, kind, g["hidden"], g["params"])
    mb = tuple(g[k] for k in ("obs", "actions", "advantages", "returns", "old_logprobs", "old_values"))
    algs = [D.PPO(ent_coef=0.01), D.PPO(ent_coef=0.02, clip_range_vf=0.2, normalize_advantage=False, vf_coef=0.7)]
    for i, alg in enumerate(algs):
        loss, stats, grads = p.loss_grad(*mb, alg.hyper())
        eloss = float(g[f"h{i}_loss"])
        assert abs(loss - eloss) <= 1e-4 * max(1.0, abs(eloss))
        for k, v in zip(g[f"h{i}_stat_names"], g[f"h{i}_stats"]):
            assert abs(stats[str(k)] - v) <= 1e-4 * max(1.0, abs(v)), (k, stats[str(k)], v)
        eg = g[f"h{i}_grads"].astype(np.float64)
        assert np.linalg.norm(grads - eg) <= 1e-4 * np.linalg.norm(eg)
    p.close()


def test_orc_adam(D):
    g = gold("orc_adam")
    p = _policy(D, "pendulum", g["hidden"], g["params"])
    for it in range(3):
        norm = p.optimizer_step(g["grads"][it], D.PPO().hyper())
        assert abs(norm - g["norms"][it]) <= 1e-5 * g["norms"][it]
        np.testing.assert_allclose(p.get_params(), g["params_after"][it], rtol=1e-6, atol=1e-7)
    m, v, step = p.get_opt_state()
    assert step == 3
    np.testing.assert_allclose(m, g["m"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(v, g["v"], rtol=1e-5, atol=1e-10)
    p.close()
