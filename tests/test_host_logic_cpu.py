"""Host-side logic of the reference mirror that needs no device: the on_step hook schedule of collect_trajectories
(buffers/trajectory.jl:34-39) as evaluated in front of the fused rollout, and get_hparams (logging/logging_utils.jl:11-35)."""
import types

import dril_b200 as D
from dril_b200 import api


class _Env:
    def observation_space(self): return D.Box([-1, -1], [1, 1])
    def action_space(self): return D.Discrete(2)


def _agent(steps=0):
    return types.SimpleNamespace(stats=types.SimpleNamespace(steps_taken=steps))


def test_on_step_hooks_schedule_and_abort():
    calls = []

    class Count(D.AbstractCallback):
        def on_step(self, loc):
            calls.append((loc["i"], set(loc) >= {"agent", "env", "alg", "n_steps", "n_envs", "callbacks", "i"}))
            return True

    class StopAt(D.AbstractCallback):
        def __init__(self, k): self.k = k
        def on_step(self, loc): return loc["i"] < self.k

    class NoStep(D.AbstractCallback):        # does not override on_step: costs nothing, never consulted
        pass

    env, alg = _Env(), D.PPO(n_steps=5)
    assert api._on_step_hooks(None, _agent(), env, alg, 5, 8)
    assert api._on_step_hooks([NoStep()], _agent(), env, alg, 5, 8) and not calls
    assert api._on_step_hooks([Count(), NoStep()], _agent(), env, alg, 5, 8)
    assert [c[0] for c in calls] == [1, 2, 3, 4, 5] and all(c[1] for c in calls)      # one call per env step, 1-based like Julia
    calls.clear()
    assert not api._on_step_hooks([Count(), StopAt(3)], _agent(), env, alg, 5, 8)
    assert [c[0] for c in calls] == [1, 2, 3]                                          # `all` short-circuits at the first false


def test_threshold_callback_matches_reference_count():
    """test/test_callbacks.jl:92-99: steps_taken only moves once per rollout, so a threshold of 500 with 8 envs x 64 steps
    lets exactly one rollout through."""
    class Threshold(D.AbstractCallback):
        def on_step(self, loc): return loc["agent"].stats.steps_taken < 500

    env, alg, agent = _Env(), D.PPO(n_steps=64), _agent(0)
    rollouts = 0
    while api._on_step_hooks([Threshold()], agent, env, alg, 64, 8):
        agent.stats.steps_taken += 64 * 8          # add_step! after the rollout (ppo.jl:173)
        rollouts += 1
    assert rollouts == 1 and agent.stats.steps_taken == 512


def test_get_hparams_keys():
    hp = D.get_hparams(D.PPO(n_steps=128, batch_size=64, epochs=4))
    for k in ("gamma", "gae_lambda", "clip_range", "ent_coef", "vf_coef", "max_grad_norm", "n_steps", "batch_size", "epochs",
              "learning_rate"):
        assert k in hp, k
