"""Host-side logic of the reference mirror that needs no device: which callbacks switch collect_rollout! to chunked collection
(buffers/trajectory.jl:34-39), and get_hparams (logging/logging_utils.jl:11-35)."""
import types

import dril_b200 as D
from dril_b200 import api


class _Env:
    def observation_space(self): return D.Box([-1, -1], [1, 1])
    def action_space(self): return D.Discrete(2)


def _agent(steps=0):
    return types.SimpleNamespace(stats=types.SimpleNamespace(steps_taken=steps))


def test_on_step_callbacks_filter():
    """Only callbacks that override on_step switch collect_rollout! to chunked collection (one env step per launch, so the hook
    of step i sees the env after i - 1 steps, trajectory.jl:34-39); the others keep the fused rollout."""
    class Count(D.AbstractCallback):
        def on_step(self, loc): return True

    class NoStep(D.AbstractCallback):
        def on_rollout_end(self, loc): return True

    a, b = Count(), NoStep()
    assert api._on_step_callbacks(None) == [] and api._on_step_callbacks([]) == []
    assert api._on_step_callbacks([b]) == []
    assert api._on_step_callbacks([b, a]) == [a]


def test_get_hparams_keys():
    hp = D.get_hparams(D.PPO(n_steps=128, batch_size=64, epochs=4))
    for k in ("gamma", "gae_lambda", "clip_range", "ent_coef", "vf_coef", "max_grad_norm", "n_steps", "batch_size", "epochs",
              "learning_rate"):
        assert k in hp, k
