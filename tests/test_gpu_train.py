"""(d) end-to-end: PPO on the batched CartPole reaches a return of 500 (north_star), and the
reference's learning bar in spirit (test/test_ppo_integration.jl:1-40)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def D():
    import __graft_entry__
    __graft_entry__.build()
    import dril_b200
    return dril_b200


def test_cartpole_reaches_500(D):
    n, T = 4096, 128
    env = D.CudaBatchedEnv("cartpole", n, seed=0, monitor_window=100)
    layer = D.ActorCriticLayer(env.observation_space(), env.action_space(), hidden_dims=[64, 64])
    alg = D.PPO(n_steps=T, batch_size=T * n // 4, epochs=4, learning_rate=1e-3, ent_coef=0.0)
    logger = D.DictLogger()
    agent = D.Agent(layer, alg, rng=np.random.default_rng(0), logger=logger)
    best = 0.0
    for it in range(150):
        out = D.train(agent, env, alg, n * T)
        assert out is not None and np.isfinite(out[0]["losses"]).all()
        best = max(best, env.monitor_stats()["ep_len_mean"])
        if best >= 499.5:
            break
    # the last-100-episodes window (monitorWrapperEnv.jl:64-70) is all 500-step episodes
    assert best >= 499.5, best
    # deterministic evaluation through the compat path (evaluation.jl:54-143) also holds the pole
    eval_env = D.CudaBatchedEnv("cartpole", 16, max_steps=500, seed=123, monitor_window=100)
    res = D.evaluate_agent(agent, eval_env, n_eval_episodes=16, deterministic=True)
    assert res["mean_reward"] >= 475, res
    pol = D.extract_policy(agent)
    a = pol(np.zeros(4, np.float32))
    assert a in (1, 2)
    assert "train/loss" in logger.scalars and "env/ep_rew_mean" in logger.scalars


def test_pendulum_normalized_improves(D):
    n, T = 1024, 128
    env = D.CudaBatchedEnv("pendulum", n, seed=0, monitor_window=100, normalize=D.NormalizeConfig())
    layer = D.ActorCriticLayer(env.observation_space(), env.action_space(), hidden_dims=[64, 64])
    alg = D.PPO(n_steps=T, batch_size=T * n // 8, epochs=6, learning_rate=1e-3, gamma=0.95)
    agent = D.Agent(layer, alg, rng=np.random.default_rng(0))
    D.train(agent, env, alg, n * T * 2)
    first = env.monitor_stats()["ep_rew_mean"]
    D.train(agent, env, alg, n * T * 40)
    last = env.monitor_stats()["ep_rew_mean"]
    assert last > first + 200, (first, last)


def test_callbacks_and_save_load(D, tmp_path):
    """test/test_callbacks.jl:24-38,70-99 and test/test_ppo_integration.jl:42-83."""
    seen = {}

    class CB(D.AbstractCallback):
        def on_training_start(self, loc):
            seen["keys"] = set(loc)
            return True

        def on_rollout_end(self, loc):
            seen["i"] = loc["i"]
            return loc["i"] < 2

    env = D.CudaBatchedEnv("cartpole", 8, seed=0, monitor_window=100, normalize=D.NormalizeConfig())
    layer = D.ActorCriticLayer(env.observation_space(), env.action_space(), hidden_dims=[16, 16])
    alg = D.PPO(n_steps=64, batch_size=64, epochs=2)
    agent = D.Agent(layer, alg, rng=np.random.default_rng(0))
    p_init = agent.train_state.parameters.copy()
    out = D.train(agent, env, alg, 8 * 64 * 5, callbacks=[CB()])
    assert out is None and seen["i"] == 2 and D.steps_taken(agent) == 2 * 8 * 64
    # on_rollout_end runs between the rollout and the update (ppo.jl:179-186): the stop at i = 2 skips the second update,
    # and the aborted train! keeps the parameters of the first one (host copy included)
    assert agent.stats.gradient_updates == 2 * (8 * 64 // 64)
    assert not np.array_equal(agent.train_state.parameters, p_init)
    np.testing.assert_array_equal(agent.train_state.parameters, agent.device.get_params())
    need = {"agent", "env", "alg", "iterations", "total_steps", "max_steps", "n_steps", "n_envs", "roll_buffer",
            "total_fps", "callbacks"}
    assert need <= seen["keys"], need - seen["keys"]
    path = D.save_policy_params_and_state(agent, str(tmp_path / "agent"))
    p0 = agent.sync_from_device().copy()
    agent2 = D.Agent(layer, alg, rng=np.random.default_rng(5))
    D.load_policy_params_and_state(agent2, path)
    np.testing.assert_array_equal(agent2.train_state.parameters, p0)
    obs = np.random.default_rng(0).normal(size=(5, 4)).astype(np.float32)
    np.testing.assert_array_equal(D.predict_actions(agent, obs, deterministic=True),
                                  D.predict_actions(agent2, obs, deterministic=True))


def test_pipelined_train_matches_sequential(D):
    """train! without callbacks enqueues iteration i+1 before reading iteration i; with a (no-op) callback it runs one
    iteration at a time.  Both must produce the same parameters, statistics and monitor window bit for bit."""
    class Noop(D.AbstractCallback):
        pass

    def run(callbacks):
        n, T = 256, 32
        env = D.CudaBatchedEnv("cartpole", n, seed=5, monitor_window=100)
        layer = D.ActorCriticLayer(env.observation_space(), env.action_space(), hidden_dims=[64, 64])
        alg = D.PPO(n_steps=T, batch_size=T * n // 2, epochs=2)
        logger = D.DictLogger()
        agent = D.Agent(layer, alg, rng=np.random.default_rng(3), logger=logger)
        out = D.train(agent, env, alg, n * T * 7, callbacks=callbacks)
        assert out is not None
        return agent.train_state.parameters.copy(), out[0], env.monitor_stats(), logger.scalars

    p_a, s_a, m_a, l_a = run(None)
    p_b, s_b, m_b, l_b = run([Noop()])
    np.testing.assert_array_equal(p_a, p_b)
    for k in s_a:
        if k != "fps":
            np.testing.assert_array_equal(s_a[k], s_b[k], err_msg=k)
    assert len(s_a["losses"]) == 7
    assert m_a == m_b
    assert [v for _, v in l_a["env/ep_rew_mean"]] == [v for _, v in l_b["env/ep_rew_mean"]]


def test_iteration_explained_variance_matches_buffer(D):
    """The explained variance reported by a train! iteration (moments fused into the GAE pass) equals the value
    computed from the buffer's stored values / returns (algorithms/ppo.jl:256), also from the NumPy formula."""
    n, T = 300, 40
    env = D.CudaBatchedEnv("cartpole", n, seed=11, monitor_window=100)
    layer = D.ActorCriticLayer(env.observation_space(), env.action_space(), hidden_dims=[64, 64])
    alg = D.PPO(n_steps=T, batch_size=T * n // 2, epochs=1)
    agent = D.Agent(layer, alg, rng=np.random.default_rng(2))
    out = D.train(agent, env, alg, n * T)
    buf = next(iter(agent._roll_buffers.values()))
    v, r = buf.download("values").astype(np.float64).ravel(), buf.download("returns").astype(np.float64).ravel()
    ev = 1.0 - np.var(v - r, ddof=1) / np.var(r, ddof=1)
    assert abs(out[0]["explained_variances"][0] - ev) < 1e-5
    assert abs(buf.explained_variance() - ev) < 1e-5


def test_callbacks_early_stopping(D):
    """test/test_callbacks.jl:56-100: a hook returning false stops training; steps_taken is 0 after an on_training_start /
    on_rollout_start stop and exactly one rollout (512) when on_step stops at a threshold of 500."""
    def setup():
        env = D.CudaBatchedEnv("cartpole", 8, seed=0, monitor_window=100, normalize=D.NormalizeConfig(gamma=0.99))
        layer = D.ActorCriticLayer(env.observation_space(), env.action_space())
        alg = D.PPO(ent_coef=0.1, n_steps=64, batch_size=64, epochs=10)
        return D.Agent(layer, alg, rng=np.random.default_rng(0)), env, alg

    class StopAtStart(D.AbstractCallback):
        def on_training_start(self, loc): return False

    class StopAtRollout(D.AbstractCallback):
        def on_rollout_start(self, loc): return False

    class StopOnStep(D.AbstractCallback):
        def __init__(self, threshold): self.threshold, self.calls, self.last_i = threshold, 0, 0

        def on_step(self, loc):
            self.calls += 1
            self.last_i = loc["i"]
            return D.steps_taken(loc["agent"]) < self.threshold

    for cb in (StopAtStart(), StopAtRollout()):
        agent, env, alg = setup()
        assert D.train(agent, env, alg, 3000, callbacks=[cb]) is None
        assert D.steps_taken(agent) == 0
    agent, env, alg = setup()
    cb = StopOnStep(500)
    assert D.train(agent, env, alg, 3000, callbacks=[cb]) is None
    assert D.steps_taken(agent) == 512
    assert cb.calls == 64 + 1 and cb.last_i == 1          # 64 hooks of rollout 1, the first hook of rollout 2 stops
    # collect_rollout!(...; callbacks) reports the failure like rollout_buffer.jl:53-57
    buf = D.RolloutBuffer(env.observation_space(), env.action_space(), alg.gae_lambda, alg.gamma, 64, 8)
    fps, ok = D.collect_rollout(buf, agent, alg, env, callbacks=[cb])
    assert not ok
    buf.close()


def test_on_step_sees_advancing_env_and_chunked_equals_fused(D):
    """trajectory.jl:34-39: on_step(i) runs before env step i, so the hook observes an env that has advanced i - 1 steps, and
    a `false` stops the collection there.  The chunked collection (one launch per step) fills the buffer exactly like the
    fused rollout."""
    n, T = 24, 12

    def setup():
        env = D.CudaBatchedEnv("pendulum", n, seed=3, monitor_window=100, normalize=D.NormalizeConfig())
        layer = D.ActorCriticLayer(env.observation_space(), env.action_space(), hidden_dims=[16, 16])
        alg = D.PPO(n_steps=T, batch_size=64, epochs=1)
        agent = D.Agent(layer, alg, rng=np.random.default_rng(1))
        buf = D.RolloutBuffer(env.observation_space(), env.action_space(), alg.gae_lambda, alg.gamma, T, n)
        return env, alg, agent, buf

    class Watch(D.AbstractCallback):
        def __init__(self, stop_at=None): self.steps, self.counts, self.stop_at = [], [], stop_at

        def on_step(self, loc):
            self.steps.append(int(loc["env"].get_state()[1][0]))
            self.counts.append(loc["env"].norm_stats()["obs_count"])
            return self.stop_at is None or loc["i"] < self.stop_at

    env, alg, agent, buf = setup()
    w = Watch()
    fps, ok = D.collect_rollout(buf, agent, alg, env, callbacks=[w])
    assert ok and w.steps == list(range(T))                    # env step counter seen by hook i is i - 1 (no episode ends in 12 steps)
    assert w.counts == [n * (i + 1) for i in range(T)]         # observe() before the loop precedes hook 1 (trajectory.jl:32), one per step after it
    env2, alg2, agent2, buf2 = setup()
    D.collect_rollout(buf2, agent2, alg2, env2)
    for k in ("obs", "actions", "rewards", "values", "logprobs", "advantages", "returns", "boot", "last_values", "flags"):
        np.testing.assert_allclose(buf.download(k), buf2.download(k), rtol=1e-6, atol=1e-6, err_msg=k)
    assert env.norm_stats()["obs_count"] == env2.norm_stats()["obs_count"] == n * (T + 1)
    # abort in the middle of a rollout: the env has advanced stop_at - 1 steps, collect_rollout! reports failure
    env3, alg3, agent3, buf3 = setup()
    w3 = Watch(stop_at=5)
    fps, ok = D.collect_rollout(buf3, agent3, alg3, env3, callbacks=[w3])
    assert not ok and len(w3.steps) == 5 and int(env3.get_state()[1][0]) == 4
    for b in (buf, buf2, buf3):
        b.close()


def test_evaluate_agent_vs_oracle_loop(D):
    """evaluate_agent (src/evaluation.jl:54-143) with the episode loop on the device against the same loop over the oracle
    env with deterministic (mode) actions: episode lengths / returns in the reference's (step, env) order."""
    from oracle import envs as OE, policy as OP
    n = 37
    spec = OP.PolicySpec(4, [64, 64], "discrete", 2, act_start=1)
    flat = OP.init_params(spec, seed=5)
    env = D.CudaBatchedEnv("cartpole", n, max_steps=60, seed=9, monitor_window=100)
    layer = D.ActorCriticLayer(env.observation_space(), env.action_space(), hidden_dims=[64, 64])
    agent = D.Agent(layer, D.PPO(), rng=np.random.default_rng(0))
    agent.set_parameters(flat)
    n_eval = 50
    er, el = D.evaluate_agent(agent, env, n_eval_episodes=n_eval, deterministic=True, return_stats=False, chunk_steps=7)
    o = OE.MonitorWrapper(OE.ParallelEnv(OE.CartPoleBatch(n, seed=9, max_steps=60)))
    o.reset()
    obs = o.observe()
    exp_r, exp_l = [], []
    while len(exp_r) < n_eval:
        a, _, _ = OP.forward(spec, flat, obs, deterministic=True)
        r, term, trunc, infos = o.act(a)
        obs = o.observe()
        for i in range(n):
            if (term[i] or trunc[i]) and len(exp_r) < n_eval:
                exp_r.append(infos["episode_r"][i]); exp_l.append(int(infos["episode_l"][i]))
    np.testing.assert_array_equal(el, np.asarray(exp_l))
    np.testing.assert_allclose(er, np.asarray(exp_r, np.float32), rtol=1e-6)
    # statistics tuple and the host-loop path over the AbstractParallelEnv interface agree
    fresh = lambda: D.CudaBatchedEnv("cartpole", n, max_steps=60, seed=9, monitor_window=100)
    st = D.evaluate_agent(agent, fresh(), n_eval_episodes=n_eval, deterministic=True)
    assert abs(st["mean_reward"] - float(np.mean(exp_r))) < 1e-4 and abs(st["mean_length"] - float(np.mean(exp_l))) < 1e-9
    assert abs(st["std_reward"] - float(np.std(np.asarray(exp_r, np.float64), ddof=1))) < 1e-4
    er_h, el_h = D.evaluate_agent(agent, fresh(), n_eval_episodes=n_eval, deterministic=True, return_stats=False, on_device=False)
    np.testing.assert_array_equal(el_h, el)
    # stochastic evaluation on an unmonitored, normalised env: returns are sums of the normalised step rewards
    env_u = D.CudaBatchedEnv("cartpole", 16, max_steps=30, seed=2, normalize=D.NormalizeConfig(training=False))
    er_u, el_u = D.evaluate_agent(agent, env_u, n_eval_episodes=20, deterministic=False, return_stats=False, warn=False)
    assert (el_u >= 1).all() and (el_u <= 30).all() and np.isfinite(er_u).all()
    np.testing.assert_allclose(er_u, el_u / np.sqrt(1.0 + 1e-8), rtol=1e-5)   # ret_var = 1 (untrained stats): r / sqrt(1 + eps)


def test_normalization_stats_save_load_sync(D, tmp_path):
    """save_normalization_stats / load_normalization_stats! / sync_normalization_stats! (normalizeWrapperEnv.jl:261-309)."""
    env = D.CudaBatchedEnv("pendulum", 32, seed=1, normalize=D.NormalizeConfig(clip_obs=5.0, gamma=0.9))
    layer = D.ActorCriticLayer(env.observation_space(), env.action_space(), hidden_dims=[16])
    alg = D.PPO(n_steps=16, batch_size=128, epochs=1)
    agent = D.Agent(layer, alg, rng=np.random.default_rng(0))
    D.train(agent, env, alg, 32 * 16 * 2)
    s = env.norm_stats()
    assert s["obs_count"] == 32 * (16 + 1) * 2 and s["ret_count"] == 32 * 16 * 2
    path = D.save_normalization_stats(env, str(tmp_path / "norm"))
    d = np.load(path)
    assert set(d.files) == {"obs_mean", "obs_var", "obs_count", "ret_mean", "ret_var", "ret_count", "clip_obs", "clip_reward", "gamma", "epsilon"}
    assert float(d["clip_obs"]) == 5.0 and abs(float(d["gamma"]) - 0.9) < 1e-7
    env2 = D.CudaBatchedEnv("pendulum", 8, seed=2, normalize=D.NormalizeConfig(training=False))
    D.load_normalization_stats(env2, path)
    s2 = env2.norm_stats()
    for k in ("obs_mean", "obs_var"):
        np.testing.assert_array_equal(s2[k], s[k])
    assert (s2["obs_count"], s2["ret_count"], s2["ret_mean"], s2["ret_var"]) == (s["obs_count"], s["ret_count"], s["ret_mean"], s["ret_var"])
    env3 = D.CudaBatchedEnv("pendulum", 5, seed=3, normalize=D.NormalizeConfig(training=False))
    D.sync_normalization_stats(env3, env)
    np.testing.assert_array_equal(env3.norm_stats()["obs_mean"], s["obs_mean"])
    # the synced eval env normalises like the training env: same raw state -> same observation
    policy = D.extract_policy(agent, env)
    raw, _ = env3.get_original() if False else (None, None)
    o3 = env3.observe()
    raw3, _ = env3.get_original()
    exp = np.clip((raw3 - s["obs_mean"]) / np.sqrt(s["obs_var"] + 1e-8), -10, 10)
    np.testing.assert_allclose(o3, exp, rtol=1e-5, atol=1e-6)
    a = policy(raw3, deterministic=True)
    assert a.shape == (5, 1) and (np.abs(a) <= 2).all()


def test_matrix_observations_flatten(D):
    """Multi-dimensional Box observations (test/test_buffers.jl:280-314): the layer's feature extractor is a Flatten
    (layers/layer_helpers.jl:13-25); buffers and observe() keep the observation shape."""
    n, T = 20, 6
    env = D.CudaBatchedEnv("synthetic", n, obs_shape=(2, 3), max_steps=4, seed=1, monitor_window=10)
    assert env.observation_space().size() == (2, 3) and env.observe().shape == (n, 2, 3)
    layer = D.ActorCriticLayer(env.observation_space(), env.action_space(), hidden_dims=[8, 8])
    assert layer.obs_dim == 6 and layer.parameterlength() == 2 * (6 * 8 + 8 + 8 * 8 + 8) + (8 * 2 + 2) + (8 + 1)
    alg = D.PPO(n_steps=T, batch_size=32, epochs=1)
    agent = D.Agent(layer, alg, rng=np.random.default_rng(0))
    buf = D.RolloutBuffer(env.observation_space(), env.action_space(), alg.gae_lambda, alg.gamma, T, n)
    D.collect_rollout(buf, agent, alg, env)
    obs = buf.download("obs")
    assert obs.shape == (T, n, 2, 3)
    v, lp, _ = agent.device.evaluate(obs.reshape(T * n, 6), buf.download("actions").reshape(T * n, 1))
    np.testing.assert_allclose(lp, buf.download("logprobs").reshape(-1), atol=1e-5)
    np.testing.assert_allclose(v, buf.download("values").reshape(-1), atol=1e-6)
    r, term, trunc, infos = env.act(np.ones(n, np.int64))
    assert any("terminal_observation" in i and i["terminal_observation"].shape == (2, 3) for i in infos) or not trunc.any()
    assert D.train(agent, env, alg, n * T * 2) is not None
    buf.close()
