"""CPU oracle for the DRiL.jl PPO rollout-and-update hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / reported
baseline.  The product path (``dril.jl_b200``) never imports this package and
fails loudly when ``libdril_b200.so`` is missing.

It is a NumPy fp32 restatement of the reference algorithm, every function
citing the reference file:line it follows (paths relative to /root/reference).

Parity status (see DESIGN.md "Oracle"):
  * GAE, bootstrap rules, RunningMeanStd, Normalize/Monitor wrappers,
    Categorical/DiagGaussian formulas: PINNED by re-running the reference's own
    closed-form tests (tests/test_oracle_pins.py cites each test file:line).
  * CartPole / Pendulum dynamics: the reference takes them from the un-vendored
    ClassicControlEnvironments.jl@main (test/Project.toml:23); restated here from
    the Gymnasium CartPole-v1 / Pendulum-v1 equations.  PARITY UNPINNED.
  * PPO loss gradients (Zygote) and Adam (Optimisers.jl 0.4): third-party, no
    reference test pins values.  The analytic backward here is cross-checked
    against torch fp64 autograd of the same restated loss.  PARITY UNPINNED
    against DRiL.jl itself (no Julia toolchain in this image).
"""
from . import philox, envs, policy, ppo  # noqa: F401
