"""Oracle (test infrastructure): actor-critic layer, distributions and the analytic backward.

Follows:
  MLP shape / init          layers/layer_helpers.jl:27-57, layers/layer_constructors.jl:16-20,61-65
  parameter order           layers/layer_lux.jl:4-52 (actor_head layers, critic_head layers, log_std);
                            Lux Dense weight is (out,in) column-major == row-major [in][out]
  forward + sample          layers/layer_forward.jl:3-13 (continuous), :30-39 (discrete)
  evaluate_actions          layers/layer_methods.jl:28-37, :46-55
  predict_values/actions    layers/layer_methods.jl:3-26, :57-61
  Categorical               DRiLDistributions/categorical.jl:20-52
  DiagGaussian              DRiLDistributions/diagGaussian.jl:4-47
  adapters                  adapters/default_adapters.jl:4-11 (clamp), :34-40 (identity)
"""
import numpy as np

from . import philox

f32 = np.float32
LOG2PI = f32(np.log(2 * np.pi))


class PolicySpec:
    def __init__(self, obs_dim, hidden, act_kind, act_n, act_start=1, act_low=None, act_high=None):
        assert act_kind in ("discrete", "continuous")
        self.obs_dim = int(obs_dim)
        self.hidden = [int(h) for h in hidden]
        self.act_kind = act_kind
        self.act_n = int(act_n)          # n actions (discrete) or act_dim (continuous)
        self.act_start = int(act_start)
        self.act_low = None if act_low is None else np.asarray(act_low, dtype=f32).reshape(-1)
        self.act_high = None if act_high is None else np.asarray(act_high, dtype=f32).reshape(-1)

    def layer_dims(self, net):
        """[(in,out)...] for net 0 (actor) / 1 (critic). layer_helpers.jl:27-57: empty hidden ->
        single Dense(in -> 1) (hard-coded 1, :33)."""
        out = self.act_n if net == 0 else 1
        if not self.hidden:
            return [(self.obs_dim, 1)]
        dims = [(self.obs_dim, self.hidden[0])]
        for i in range(1, len(self.hidden)):
            dims.append((self.hidden[i - 1], self.hidden[i]))
        dims.append((self.hidden[-1], out))
        return dims

    def n_params(self):
        n = 0
        for net in (0, 1):
            for (i, o) in self.layer_dims(net):
                n += i * o + o
        if self.act_kind == "continuous":
            n += self.act_n
        return n


def unflatten(spec, flat):
    """flat fp32 vector (ComponentVector order) -> {'actor': [(W[in,out], b)...], 'critic': [...], 'log_std'}."""
    flat = np.asarray(flat, dtype=f32)
    p = 0
    out = {}
    for net, name in ((0, "actor"), (1, "critic")):
        layers = []
        for (i, o) in spec.layer_dims(net):
            W = flat[p:p + i * o].reshape(i, o)
            p += i * o
            b = flat[p:p + o]
            p += o
            layers.append((W, b))
        out[name] = layers
    if spec.act_kind == "continuous":
        out["log_std"] = flat[p:p + spec.act_n]
        p += spec.act_n
    assert p == flat.size, (p, flat.size)
    return out


def flatten(spec, params):
    parts = []
    for name in ("actor", "critic"):
        for (W, b) in params[name]:
            parts += [np.asarray(W, dtype=f32).reshape(-1), np.asarray(b, dtype=f32).reshape(-1)]
    if spec.act_kind == "continuous":
        parts.append(np.asarray(params["log_std"], dtype=f32).reshape(-1))
    return np.concatenate(parts).astype(f32)


def _orthogonal(rng, out_dims, in_dims, gain):
    """WeightInitializers.orthogonal analogue (QR of a Gaussian matrix); returns (out,in)."""
    rows, cols = out_dims, in_dims
    a = rng.standard_normal((max(rows, cols), min(rows, cols)))
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.diag(r))
    if rows < cols:
        q = q.T
    return (gain * q[:rows, :cols]).astype(f32)


def init_params(spec, seed=0, log_std_init=0.0):
    """Orthogonal init, gains sqrt(2) / 0.01 / 1.0, zero bias (layer_constructors.jl:16-20,61-65).
    The exact Julia RNG stream is unpinned and unnecessary: weights are passed in."""
    rng = np.random.default_rng(seed)
    params = {}
    for net, name, out_gain in ((0, "actor", 0.01), (1, "critic", 1.0)):
        dims = spec.layer_dims(net)
        layers = []
        for li, (i, o) in enumerate(dims):
            gain = out_gain if li == len(dims) - 1 else np.sqrt(2.0)
            W_oi = _orthogonal(rng, o, i, gain)       # Lux (out,in)
            layers.append((np.ascontiguousarray(W_oi.T), np.zeros(o, dtype=f32)))
        params[name] = layers
    if spec.act_kind == "continuous":
        params["log_std"] = np.full(spec.act_n, log_std_init, dtype=f32)
    return flatten(spec, params)


def mlp_forward(layers, x, keep=False):
    """x: (B, in). tanh on all but the last Dense (layer_helpers.jl:33-56)."""
    acts = [x]
    h = x
    for li, (W, b) in enumerate(layers):
        z = (h @ W + b).astype(f32)
        h = np.tanh(z).astype(f32) if li < len(layers) - 1 else z
        acts.append(h)
    return (h, acts) if keep else h


def softmax(logits):
    m = logits.max(axis=1, keepdims=True)
    e = np.exp(logits - m).astype(f32)
    return (e / e.sum(axis=1, keepdims=True, dtype=f32)).astype(f32)


def categorical_sample(probs, u64, start):
    """categorical.jl:44-49: first index with cumsum(p) >= u (fp32 cumsum vs Float64 u)."""
    cum = np.cumsum(probs, axis=1, dtype=f32)
    ge = cum.astype(np.float64) >= u64[:, None]
    idx = np.where(ge.any(axis=1), ge.argmax(axis=1), probs.shape[1] - 1)  # findfirst->nothing guard
    return idx.astype(np.int64) + start


def categorical_logpdf(probs, actions, start):  # categorical.jl:29-36
    idx = np.asarray(actions).astype(np.int64).reshape(-1) - start
    return np.log(probs[np.arange(probs.shape[0]), idx]).astype(f32)


def categorical_entropy(probs):  # categorical.jl:38-40
    return (-(probs * np.log(probs)).sum(axis=1, dtype=f32)).astype(f32)


def gaussian_logpdf(mean, log_std, x):  # diagGaussian.jl:26-38
    k = mean.shape[1]
    ls_sum = log_std.sum(dtype=f32)
    diff = x - mean
    var_inv = np.exp(f32(-2) * log_std).astype(f32)
    dss = (diff * diff * var_inv).sum(axis=1, dtype=f32)
    return (f32(-0.5) * (f32(2) * ls_sum + dss + f32(k) * LOG2PI)).astype(f32)


def gaussian_entropy(log_std, B):  # diagGaussian.jl:40-45
    k = log_std.size
    return np.full(B, f32(0.5) * f32(k) * (f32(1) + LOG2PI) + log_std.sum(dtype=f32), dtype=f32)


def to_env(spec, actions):
    """adapters/default_adapters.jl:4-11 (ClampAdapter) / :34-40 (DiscreteAdapter)."""
    if spec.act_kind == "continuous":
        return np.clip(actions, spec.act_low, spec.act_high).astype(f32)
    return actions


def forward(spec, flat, obs, env_gid=None, step_idx=0, seed=0, forced_actions=None,
            deterministic=False):
    """The layer call (layer_forward.jl:3-13 / :30-39): returns (actions, values, logprobs).
    Sampling uses the Philox SAMPLE stream keyed by (env_gid, step_idx)."""
    p = unflatten(spec, flat)
    obs = np.asarray(obs, dtype=f32)
    B = obs.shape[0]
    out = mlp_forward(p["actor"], obs)
    values = mlp_forward(p["critic"], obs).reshape(B)
    if env_gid is None:
        env_gid = np.arange(B)
    if spec.act_kind == "discrete":
        probs = softmax(out)
        if forced_actions is not None:
            actions = np.asarray(forced_actions).astype(np.int64).reshape(B)
        elif deterministic:
            actions = probs.argmax(axis=1).astype(np.int64) + spec.act_start  # categorical.jl:42
        else:
            u = philox.sample_uniform64(env_gid, step_idx, seed)
            actions = categorical_sample(probs, u, spec.act_start)
        logp = categorical_logpdf(probs, actions, spec.act_start)
    else:
        log_std = p["log_std"]
        if forced_actions is not None:
            actions = np.asarray(forced_actions, dtype=f32).reshape(B, spec.act_n)
        elif deterministic:
            actions = out.copy()
        else:
            eps = philox.normals(env_gid, step_idx, spec.act_n, seed)
            actions = (out + np.exp(log_std).astype(f32) * eps).astype(f32)  # diagGaussian.jl:13-17
        logp = gaussian_logpdf(out, log_std, actions)
    return actions, values.astype(f32), logp


def predict_values(spec, flat, obs):  # layer_methods.jl:57-61
    p = unflatten(spec, flat)
    obs = np.asarray(obs, dtype=f32)
    return mlp_forward(p["critic"], obs).reshape(obs.shape[0]).astype(f32)


def evaluate_actions(spec, flat, obs, actions):
    """layer_methods.jl:28-37 / :46-55 -> (values, logprobs, entropy)."""
    p = unflatten(spec, flat)
    obs = np.asarray(obs, dtype=f32)
    B = obs.shape[0]
    out = mlp_forward(p["actor"], obs)
    values = mlp_forward(p["critic"], obs).reshape(B)
    if spec.act_kind == "discrete":
        probs = softmax(out)
        return values, categorical_logpdf(probs, actions, spec.act_start), categorical_entropy(probs)
    a = np.asarray(actions, dtype=f32).reshape(B, spec.act_n)
    return values, gaussian_logpdf(out, p["log_std"], a), gaussian_entropy(p["log_std"], B)
