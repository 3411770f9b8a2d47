"""Multi-GPU consistency check (run under torch.distributed.run): the peer-memory allreduce path and the NCCL
path give the same parameters, and all ranks hold bit-identical parameters after training."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dril_b200 as D  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ.setdefault("NCCL_DEBUG", "WARN")
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def run(use_p2p):
    ctx = D.Context(device=local, seed=0)
    uid = [D.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(rank, world, uid[0])
    n, T = 1024, 32
    env = D.CudaBatchedEnv("cartpole", n, seed=0, ctx=ctx, monitor_window=100, gid_offset=rank * n)
    layer = D.ActorCriticLayer(env.observation_space(), env.action_space(), hidden_dims=[64, 64])
    alg = D.PPO(n_steps=T, batch_size=n * T // 4, epochs=3, learning_rate=1e-3)
    agent = D.Agent(layer, alg, rng=np.random.default_rng(0), ctx=ctx)
    if use_p2p:
        def all_gather(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out
        ctx.comm_p2p_setup(all_gather, agent.device.n_params + 8)
    out = D.train(agent, env, alg, n * T * 4)
    assert out is not None and np.isfinite(out[0]["losses"]).all()
    params = agent.train_state.parameters.copy()
    gathered = [None] * world
    dist.all_gather_object(gathered, params)
    for r in range(world):
        assert np.array_equal(gathered[r], gathered[0]), f"rank {r} parameters differ from rank 0 (p2p={use_p2p})"
    return params, out[0]


p_nccl, s_nccl = run(False)
p_p2p, s_p2p = run(True)
diff = np.abs(p_nccl - p_p2p).max()
if rank == 0:
    print(f"world={world}: ranks bit-identical on both paths; max |param(nccl) - param(p2p)| = {diff:.3e}; "
          f"losses nccl {s_nccl['losses']} p2p {s_p2p['losses']}", flush=True)
assert diff <= 1e-5, diff
dist.barrier()
dist.destroy_process_group()
