"""Box / Discrete boundary types against the scenarios of the reference's test/test_spaces.jl (creation, validation,
sampling, containment, equality).  No device needed."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("dril_spaces", os.path.join(ROOT, "dril.jl_b200", "spaces.py"))
S = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(S)
f32 = np.float32


def test_box_creation_and_validation():            # test_spaces.jl:2-45
    b = S.Box(f32([-2, -1]), f32([1, 3]))
    assert b.shape == (2,) and b.low.dtype == f32 and (b.low == [-2, -1]).all() and (b.high == [1, 3]).all()
    assert S.Box(f32([-5]), f32([10])).shape == (1,)
    assert S.Box(f32([-1, 0, -10]), f32([1, 5, 0])).shape == (3,)
    with pytest.raises(AssertionError):
        S.Box(f32([-1]), f32([1, 2]))               # mismatched shapes
    with pytest.raises(AssertionError):
        S.Box(f32([1, -1]), f32([0, 1]))            # low > high
    S.Box(f32([1, 2]), f32([1, 2]))                 # low == high is valid


def test_box_sampling_and_containment():            # test_spaces.jl:47-175
    space = S.Box(f32([-3, 0]), f32([2, 10]))
    rng = np.random.default_rng(42)
    x = space.sample(rng)
    assert x.shape == (2,) and x.dtype == f32 and x in space
    assert space.sample() in space
    xs = space.sample(rng, 5)
    assert len(xs) == 5 and all(v.dtype == f32 and v in space for v in xs)
    space = S.Box(f32([-2, 1]), f32([0, 5]))
    for v in ([-1, 3], [-2, 1], [0, 5], [-1.5, 2.5]):
        assert f32(v) in space
    for v in ([0.5, 3], [-1, 0.5], [-3, 6], [-1, 3, 0]):
        assert f32(v) not in space
    tiny = S.Box(f32([0]), f32([1e-6]))
    assert f32([0]) in tiny and f32([1e-6]) in tiny and f32([1e-5]) not in tiny
    point = S.Box(f32([1, 2]), f32([1, 2]))
    assert f32([1, 2]) in point and f32([1, 2.1]) not in point
    assert (point.sample(rng) == [1, 2]).all()


def test_discrete_properties_sampling_containment():    # test_spaces.jl:177-270
    d = S.Discrete(5)
    assert d.n == 5 and d.start == 1 and d.size() == (1,)
    assert S.Discrete(4, -2).start == -2
    assert S.Discrete(5, 0) == S.Discrete(5, 0) and S.Discrete(5, 0) != S.Discrete(5, 1) and S.Discrete(5, 0) != S.Discrete(4, 0)
    rng = np.random.default_rng(42)
    s0 = S.Discrete(5, 0)
    v = s0.sample(rng)
    assert isinstance(v, int) and 0 <= v <= 4 and v in s0
    vs = s0.sample(rng, 200)
    assert len(vs) == 200 and set(vs) == {0, 1, 2, 3, 4}
    assert set(S.Discrete(3, 1).sample(rng, 100)) == {1, 2, 3}
    assert set(S.Discrete(4, -1).sample(rng, 100)) == {-1, 0, 1, 2}
    assert all(k in s0 for k in range(5)) and all(k not in s0 for k in (-1, 5, 6, 10))
    for bad in (1.0, 1.5, "1", [1], True):
        assert bad not in s0
    with pytest.raises(AssertionError):
        S.Discrete(0)
