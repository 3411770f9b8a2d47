"""Counter-based RNG spec shared by the oracle and the CUDA library (test infrastructure).

The reference draws from Julia's task-local Xoshiro (layers/layer_forward.jl:10,36) and
from ``agent.rng`` for shuffling (algorithms/ppo.jl:194); those streams cannot be
reproduced on a GPU, so north_star prescribes a counter-based Philox.  This file is the
normative statement of that stream so that sampled actions, reset states and minibatch
permutations can be compared element-for-element between the oracle and the kernels.

Philox4x32-10 (Salmon et al., SC'11), key = (seed_lo, seed_hi),
counter = (c0, c1, c2, tag):

  tag 1 RESET    c0 = global env id, c1 = episode index, c2 = 0      -> reset uniforms
  tag 2 SAMPLE   c0 = global env id, c1 = policy step index, c2 = block -> action noise
  tag 3 SYN_OBS  c0 = global env id, c1 = env lifetime step, c2 = block -> synthetic obs
  tag 4 SHUFFLE  c0 = epoch counter, c1 = rank, c2 = 0               -> Feistel round keys
  tag 5 SYN_DYN  c0 = global env id, c1 = env lifetime step, c2 = 0  -> synthetic reward/term
"""
import numpy as np

TAG_RESET, TAG_SAMPLE, TAG_SYN_OBS, TAG_SHUFFLE, TAG_SYN_DYN = 1, 2, 3, 4, 5

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32(c0, c1, c2, c3, seed):
    """Vectorised Philox4x32-10. c* broadcastable uint32-valued arrays; seed: python int (64 bit).

    Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(
        *(np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3)))
    c0, c1, c2, c3 = (c.copy() for c in (c0, c1, c2, c3))
    k0 = int(seed) & 0xFFFFFFFF
    k1 = (int(seed) >> 32) & 0xFFFFFFFF
    for r in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def u01_f32(x):
    """uint32 -> float32 uniform in [0, 1): top 24 bits * 2^-24 (exact in fp32)."""
    return ((np.asarray(x, dtype=np.uint32) >> np.uint32(8)).astype(np.float32)
            * np.float32(2.0 ** -24))


def u01_f32_open(x):
    """uint32 -> float32 uniform in (0, 1]: (top 24 bits + 1) * 2^-24."""
    return (((np.asarray(x, dtype=np.uint32) >> np.uint32(8)).astype(np.float32)
             + np.float32(1.0)) * np.float32(2.0 ** -24))


def u01_f64(x0, x1):
    """two uint32 -> float64 uniform in [0,1) with 53 random bits (Julia rand(Float64) analogue,
    DRiLDistributions/categorical.jl:46)."""
    bits = (np.asarray(x0, dtype=np.uint64) << np.uint64(21)) | (
        np.asarray(x1, dtype=np.uint64) >> np.uint64(11))
    return bits.astype(np.float64) * (2.0 ** -53)


def box_muller(xa, xb):
    """Two uint32 words -> two fp32 standard normals (z_cos, z_sin). fp32 arithmetic."""
    u1 = u01_f32_open(xa)
    u2 = u01_f32(xb)
    r = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
    th = (np.float32(6.2831853071795864) * u2).astype(np.float32)
    return (r * np.cos(th)).astype(np.float32), (r * np.sin(th)).astype(np.float32)


def normals(env_gid, step_idx, n, seed):
    """n fp32 N(0,1) draws per env for the SAMPLE stream: block b gives draws 4b..4b+3 as
    (cos(x0,x1), sin(x0,x1), cos(x2,x3), sin(x2,x3))."""
    env_gid = np.asarray(env_gid)
    out = np.zeros(env_gid.shape + (n,), dtype=np.float32)
    for b in range((n + 3) // 4):
        x0, x1, x2, x3 = philox4x32(env_gid, step_idx, b, TAG_SAMPLE, seed)
        za, zb = box_muller(x0, x1)
        zc, zd = box_muller(x2, x3)
        for j, z in enumerate((za, zb, zc, zd)):
            if 4 * b + j < n:
                out[..., 4 * b + j] = z
    return out


def sample_uniform64(env_gid, step_idx, seed):
    """The float64 uniform used for categorical inverse-CDF sampling."""
    x0, x1, _, _ = philox4x32(env_gid, step_idx, 0, TAG_SAMPLE, seed)
    return u01_f64(x0, x1)


# ----------------------------------------------------------------------------------------
# Minibatch shuffle: a keyed bijection on [0, n) (4-round Feistel + cycle walking).
# Replaces MLUtils.DataLoader(shuffle=true, rng=agent.rng) (algorithms/ppo.jl:188-195), whose
# permutation order is unpinned by the reference; any uniform-looking permutation per epoch
# preserves the semantics (every sample exactly once per epoch, batches of batch_size,
# last batch partial).
# ----------------------------------------------------------------------------------------
def feistel_keys(epoch_counter, rank, seed):
    return [int(k) for k in philox4x32(epoch_counter, rank, 0, TAG_SHUFFLE, seed)]


def _feistel_round_fn(r, key, half_mask):
    h = (r.astype(np.uint64) + np.uint64(key)) & _MASK
    h = (h * np.uint64(0x9E3779B1)) & _MASK
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(0x85EBCA77)) & _MASK
    h ^= h >> np.uint64(13)
    return h & np.uint64(half_mask)


def feistel_half_bits(n):
    bits = max(2, int(n - 1).bit_length())
    return (bits + 1) // 2


def feistel_permute(idx, n, keys):
    """perm[i] for i in idx: bijection on [0,n)."""
    hb = feistel_half_bits(n)
    half_mask = (1 << hb) - 1
    x = np.asarray(idx, dtype=np.uint64).copy()
    todo = np.ones(x.shape, dtype=bool)
    while True:
        l = (x >> np.uint64(hb)) & np.uint64(half_mask)
        r = x & np.uint64(half_mask)
        for k in keys:
            l, r = r, l ^ _feistel_round_fn(r, k, half_mask)
        y = (l << np.uint64(hb)) | r
        x = np.where(todo, y, x)
        todo = todo & (x >= np.uint64(n))
        if not todo.any():
            break
    return x.astype(np.int64)
