"""N>1 host-side logic on CPU (gloo, world_size 2): data-parallel sharding over envs is exact.
  * global env ids make the env RNG streams independent of the rank split;
  * allreduce(sum) of per-rank gradient sums scaled by 1/B_global, with the minibatch advantage
    moments allreduced first, equals the single-process gradient on the union minibatch
    (per-minibatch normalisation over the GLOBAL minibatch, ppo.jl:350-356, 375)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import envs as OE, policy as OP, ppo as OO


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec = OP.PolicySpec(4, [16, 16], "discrete", 2, act_start=1)
    flat = OP.init_params(spec, seed=0)
    n_local, T = 6, 12
    env = OE.ParallelEnv(OE.CartPoleBatch(n_local, seed=3, gid_offset=rank * n_local))
    gids = np.arange(n_local) + rank * n_local
    buf = OO.collect_rollout_timemajor(env, spec, flat, T, policy_seed=5, env_gid=gids)
    adv, ret = OO.gae_timemajor(buf["rewards"], buf["values"], buf["term"], buf["trunc"], buf["boot"], buf["last_values"], 0.99, 0.95)
    # global advantage moments -> every rank normalises with the same mean / Bessel std
    a = adv.reshape(-1).astype(np.float64)
    mom = torch.tensor([a.sum(), (a * a).sum(), a.size], dtype=torch.float64)
    dist.all_reduce(mom)
    n = mom[2].item()
    mean = mom[0].item() / n
    std = np.sqrt((mom[1].item() - n * mean * mean) / (n - 1))
    adv_n = ((adv.reshape(-1) - np.float32(mean)) / (np.float32(std) + np.float32(1e-8))).astype(np.float32)
    cfg = OO.PPOConfig(normalize_advantage=False, ent_coef=0.01)
    B = T * n_local
    _, stats, g = OO.ppo_loss_and_grads(spec, flat, buf["obs"].reshape(B, -1), buf["actions"].reshape(B, -1), adv_n,
                                        ret.reshape(-1), buf["logprobs"].reshape(-1), buf["values"].reshape(-1), cfg)
    gt = torch.tensor(g.astype(np.float64) * B)          # local SUM (oracle returns the local mean)
    dist.all_reduce(gt)
    g_global = (gt / n).numpy()
    if rank == 0:
        np.save(out, dict(g=g_global, obs=buf["obs"], adv=adv, mean=mean, std=std), allow_pickle=True)
    # keep the shards for the single-process comparison
    np.save(out + f".rank{rank}.npy", dict(obs=buf["obs"], actions=buf["actions"], adv=adv, ret=ret, logp=buf["logprobs"],
                                           val=buf["values"]), allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_gradient_equals_single_process(tmp_path):
    out = str(tmp_path / "dp.npy")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    dp = np.load(out, allow_pickle=True).item()
    shards = [np.load(out + f".rank{r}.npy", allow_pickle=True).item() for r in range(2)]
    spec = OP.PolicySpec(4, [16, 16], "discrete", 2, act_start=1)
    flat = OP.init_params(spec, seed=0)
    # single process over all 12 envs: same env streams thanks to global env ids
    env = OE.ParallelEnv(OE.CartPoleBatch(12, seed=3))
    buf = OO.collect_rollout_timemajor(env, spec, flat, 12, policy_seed=5)
    np.testing.assert_array_equal(buf["obs"][:, :6], shards[0]["obs"])
    np.testing.assert_array_equal(buf["obs"][:, 6:], shards[1]["obs"])
    cat = lambda k: np.concatenate([s[k].reshape(s[k].shape[0] * s[k].shape[1], -1) for s in shards])
    cfg = OO.PPOConfig(normalize_advantage=True, ent_coef=0.01)
    _, _, g = OO.ppo_loss_and_grads(spec, flat, cat("obs"), cat("actions"), cat("adv").reshape(-1), cat("ret").reshape(-1),
                                    cat("logp").reshape(-1), cat("val").reshape(-1), cfg)
    np.testing.assert_allclose(dp["g"], g, rtol=2e-4, atol=1e-6)
