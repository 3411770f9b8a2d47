"""Box / Discrete spaces — mirror of src/spaces.jl:28-44,157-164 (boundary types)."""
import numpy as np


class Box:
    def __init__(self, low, high, shape=None):
        if shape is not None and np.isscalar(low):
            low = np.full(shape, low, dtype=np.float32)
            high = np.full(shape, high, dtype=np.float32)
        self.low = np.asarray(low, dtype=np.float32)
        self.high = np.asarray(high, dtype=np.float32)
        assert self.low.shape == self.high.shape, "Low and high arrays must have the same shape"
        assert (self.low <= self.high).all(), "All low values must be <= corresponding high values"
        self.shape = self.low.shape

    def size(self):
        return self.shape

    def __eq__(self, o):
        return isinstance(o, Box) and self.shape == o.shape and (self.low == o.low).all() and (self.high == o.high).all()

    def __contains__(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and (x >= self.low).all() and (x <= self.high).all()

    def sample(self, rng=None, n=None):
        """rand([rng,] space[, n]) — src/spaces.jl:60-106: uniform in [low, high] per dimension, Float32; with n a list of n
        samples (Julia returns a Vector of Vectors)."""
        rng = rng if rng is not None else np.random.default_rng()
        one = lambda: (self.low + rng.random(self.shape, dtype=np.float32) * (self.high - self.low)).astype(np.float32)
        return one() if n is None else [one() for _ in range(n)]

    def __repr__(self):
        return f"Box(shape={self.shape})"


class Discrete:
    """Discrete(n, start=1): values start .. start+n-1 (Julia convention; Gym is start=0)."""

    def __init__(self, n, start=1):
        assert n > 0, "n must be positive"
        self.n, self.start = int(n), int(start)

    def size(self):
        return (1,)

    def __eq__(self, o):
        return isinstance(o, Discrete) and self.n == o.n and self.start == o.start

    def __contains__(self, x):
        return isinstance(x, (int, np.integer)) and not isinstance(x, (bool, np.bool_)) and self.start <= x <= self.start + self.n - 1

    def sample(self, rng=None, n=None):
        """rand([rng,] space[, n]) — src/spaces.jl:190-227: integers start .. start+n-1."""
        rng = rng if rng is not None else np.random.default_rng()
        if n is None:
            return int(rng.integers(self.start, self.start + self.n))
        return [int(v) for v in rng.integers(self.start, self.start + self.n, size=n)]

    def __repr__(self):
        return f"Discrete({self.n}, {self.start})"
