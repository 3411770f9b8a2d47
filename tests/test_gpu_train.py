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
    out = D.train(agent, env, alg, 8 * 64 * 5, callbacks=[CB()])
    assert out is None and seen["i"] == 2 and D.steps_taken(agent) == 2 * 8 * 64
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
