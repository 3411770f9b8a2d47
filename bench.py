#!/usr/bin/env python
"""bench.py — end-to-end PPO env-steps/s (BASELINE.json metric) on N B200s of one node.

A "step" is one PPO iteration of the hot path: fused rollout (n_steps x n_envs env steps with the
actor-critic in the loop) -> GAE -> `epochs` x `n_minibatches` fused loss/grad + clip + Adam
updates, over synthetic env batches of the named shape with random-init weights.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4]

N > 1 is launched by the driver with torch.distributed.run (one rank per GPU): envs shard across
ranks (weak scaling: per-GPU env count fixed), gradients are allreduced over NVLink peer memory per minibatch
(NCCL fallback with DRIL_NO_P2P=1).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on (fits one GPU)
    "c2": dict(name="PPO CartPole, 4096 batched envs/GPU, n_steps=128, hidden [64,64], discrete categorical policy",
               kind="cartpole", n_envs=4096, n_steps=128, hidden=[64, 64], normalize=False),
    "c3": dict(name="PPO Pendulum diag-Gaussian, 16384 envs/GPU, n_steps=128, hidden [128,128,64], NormalizeWrapperEnv obs+reward",
               kind="pendulum", n_envs=16384, n_steps=128, hidden=[128, 128, 64], normalize=True),
    "c4": dict(name="PPO CartPole, 65536 envs/GPU data-parallel, n_steps=128, hidden [64,64]",
               kind="cartpole", n_envs=65536, n_steps=128, hidden=[64, 64], normalize=False),
}
EPOCHS, N_MINIBATCHES = 4, 4          # BASELINE.json does not fix these; stated with every number
METRIC, UNIT = "end-to-end PPO env-steps/sec", "env-steps/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"],
                    sm_max_mhz=d.get("sm_max_mhz", 1965.0), source="measured")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, sm_max_mhz=1965.0, source="fallback")


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region through NVML (a polling thread in this
    process; an external `nvidia-smi -lms` loop measurably slowed the timed kernels)."""

    def __init__(self, gpu_index):
        self.idx, self.samples, self.stop_flag, self.th, self.err = gpu_index, [], False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.idx]) if vis and vis.split(",")[self.idx].isdigit() else self.idx
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def _poll(self):
        # every NVML query stalls the GPU front end for ~1.5 ms (measured: 4 samples cost 13 % of a 50 ms region), so the
        # region is sampled sparsely: once 10 ms after the start, then every 100 ms; the max clock is read before it
        nv = self.nv
        time.sleep(0.01)
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((sm, self.mx, rs, None))
            except Exception as e:  # pragma: no cover
                self.err = repr(e)
                return
            for _ in range(10):
                if self.stop_flag:
                    break
                time.sleep(0.01)

    def stop(self):
        if self.th is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % self.err]}
        self.stop_flag = True
        self.th.join(timeout=1.0)
        nv = self.nv
        names = {getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
                 getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                 getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                 getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap"}
        reasons = set()
        for (_, _, rs, _) in self.samples:
            for bit, name in names.items():
                if rs & bit:
                    reasons.add(name)
        sm = [x[0] for x in self.samples]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(x[1] for x in self.samples) if sm else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "how": "NVML (clock + event reasons) inside the timed region: 10 ms after its start, then every 100 ms"}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep NCCL's version banner off stdout (one JSON line only)
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        return rank, world, local, dist
    return rank, world, local, None


def run_ours(args):
    import dril_b200 as D
    import __graft_entry__
    rank, world, local, dist = dist_setup(args.gpus)
    if rank == 0:
        __graft_entry__.build()
    if dist:
        dist.barrier()
    w = WORKLOADS[args.workload]
    n_envs, n_steps = w["n_envs"], w["n_steps"]
    ctx = D.Context(device=local, seed=0)
    if world > 1:
        import torch
        uid = [D.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0])
    env = D.CudaBatchedEnv(w["kind"], n_envs, seed=0, ctx=ctx, monitor_window=100, gid_offset=rank * n_envs,
                           normalize=D.NormalizeConfig() if w["normalize"] else None)
    layer = D.ActorCriticLayer(env.observation_space(), env.action_space(), hidden_dims=w["hidden"])
    batch = n_steps * n_envs // N_MINIBATCHES                      # per-rank minibatch; global = batch * world
    alg = D.PPO(n_steps=n_steps, batch_size=batch, epochs=EPOCHS)
    agent = D.Agent(layer, alg, rng=np.random.default_rng(0), ctx=ctx)   # same init on every rank
    comm = "single GPU"
    if world > 1:
        comm = "NCCL allreduce of the flat gradient per minibatch"
        if os.environ.get("DRIL_NO_P2P", "0") != "1":
            def all_gather(b):
                out = [None] * world
                dist.all_gather_object(out, b)
                return out
            ctx.comm_p2p_setup(all_gather, agent.device.n_params + 8)
            comm = ("one-shot NVLink peer-memory allreduce (push) of the flat gradient inside the loss/grad kernel's fused tail; "
                    "advantage / explained-variance moments through a second small peer-memory kernel; no NCCL on the data path")
    buf = D.RolloutBuffer(env.observation_space(), env.action_space(), alg.gae_lambda, alg.gamma, n_steps, n_envs, ctx=ctx)
    import ctypes as C
    from dril_b200 import _lib as L
    lib, hyper = ctx.lib, alg.hyper()
    steps_per_iter = n_steps * n_envs

    pending = [0]
    last_stats = [None]

    def read_result():
        s_ = L.IterStats()
        L.check(lib.dril_iteration_result(agent.device.h, C.byref(s_)))
        pending[0] -= 1
        last_stats[0] = s_
        return s_

    def iteration_async():
        # results are read in FIFO order and at most 4 iterations may be in flight: reading an OLD result waits on that
        # iteration's event only, so the host stays ahead of the device and the stream never drains inside a timed loop
        if pending[0] >= int(os.environ.get('BENCH_PENDING', '3')):
            read_result()
        L.check(lib.dril_ppo_iteration_async(env.h, agent.device.h, buf.h, C.byref(hyper), alg.epochs, alg.batch_size,
                                             agent.shuffle_seed, agent.epoch_counter))
        agent.epoch_counter += alg.epochs
        pending[0] += 1

    def drain():
        while pending[0] > 0:
            read_result()
        return last_stats[0]

    def barrier():
        ctx.synchronize()
        if dist:
            dist.barrier()
        ctx.synchronize()

    # ---- warm-up --------------------------------------------------------------------------
    # at least W (>= 3) iterations, and at least 0.3 s of work so that a box that was idle reaches its sustained clocks
    # and the pipelined enqueue path is in steady state before anything is timed
    t_w = time.perf_counter()
    n_w = 0
    while n_w < max(args.warmup, 3) or time.perf_counter() - t_w < 0.3:
        iteration_async()
        n_w += 1
        if n_w % 3 == 0:
            drain()
    st = drain()
    # ---- value: K iterations resident on the device, CUDA events on the launching stream -----
    sampler = ClockSampler(local)
    barrier()
    if rank == 0 and not os.environ.get('BENCH_NO_SAMPLER'):
        sampler.start()
    launches0 = ctx.launch_count()
    ctx.event_record(0)
    for _ in range(args.steps):
        ctx.flush_l2()                  # inside the timed region: nothing stays L2-hot from the previous step
        iteration_async()
    ctx.event_record(1)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ctx.event_elapsed_ms(0, 1)
    launches = ctx.launch_count() - launches0
    st = drain()
    if dist:
        import torch
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = steps_per_iter * world * args.steps / (ms * 1e-3)

    # ---- e2e: the public API call (train!) with host parameters in and statistics out ---------
    D.train(agent, env, alg, steps_per_iter * 3)       # warm-up of the public path (its rollout buffer is allocated once per agent)
    barrier()
    t0 = time.perf_counter()
    out = D.train(agent, env, alg, steps_per_iter * args.steps)
    ctx.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist:
        import torch
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    assert out is not None and np.isfinite(out[0]["losses"]).all()
    e2e_value = steps_per_iter * world * args.steps / e2e_s
    n_params = agent.device.n_params
    h2d = int(n_params * 4 / args.steps + C.sizeof(L.PPOHyper))       # parameters once per train! + hyper per iteration
    d2h = int(n_params * 4 / args.steps + C.sizeof(L.IterStats) + 64)

    # ---- per-kernel CUDA-event profile of the same steps (roofline) -------------------------------
    ctx.set_profiling(True)
    ctx.reset_profile()
    for _ in range(args.steps):
        ctx.flush_l2()
        iteration_async()
    ctx.synchronize()
    prof = ctx.profile()
    ctx.set_profiling(False)
    st = drain()

    if rank != 0:
        if dist:
            dist.barrier()
            dist.destroy_process_group()
        return
    pk = peaks()
    D_obs = env.obs_dim
    act_bytes = 4
    spec_fwd_flops = 2 * sum(i * o for net in (0, 1) for (i, o) in layer.layer_dims(net))
    total_prof_ms = sum(v[0] for v in prof.values()) or 1.0
    kern = {}
    for name, (tms, n) in prof.items():
        if n:
            kern[name] = {"ms_per_step": tms / args.steps, "launches_per_step": n / args.steps, "share": tms / total_prof_ms}
    # algorithmic bytes / flops per launch (DESIGN.md "Kernels"; SURVEY §8d per-unit figures)
    ro = kern.get("rollout", {"ms_per_step": float("nan")})
    ro_bytes = (4 * D_obs + 17) * steps_per_iter
    ro_flops = spec_fwd_flops * steps_per_iter
    lg = kern.get("loss_grad", {"ms_per_step": float("nan"), "launches_per_step": 1})
    lg_bytes = (4 * D_obs + act_bytes + 12) * steps_per_iter * EPOCHS
    lg_flops = 3 * spec_fwd_flops * steps_per_iter * EPOCHS
    gae = kern.get("gae", {"ms_per_step": float("nan")})
    fp32_peak = 148 * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
    lg_path = agent.device.update_path()            # "tensor": tcgen05 3xTF32 kernel, "fp32": CUDA-core kernel
    hid_flops = 2 * sum(i * o for net in (0, 1) for (i, o) in layer.layer_dims(net)[1:-1])     # the square hidden GEMMs
    lg_tensor_flops = 3 * (3 * hid_flops) * steps_per_iter * EPOCHS if lg_path in ("tensor", "mma") else 0.0   # 3 GEMMs x 3 TF32 passes
    mma_tf32_peak = 148 * 512 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12    # mma.sync TF32 issue rate measured by tools/mma_probe.cu
    rooflines = {
        "rollout": {"bound": "hbm", "achieved": ro_bytes / (ro["ms_per_step"] * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                    "fp32_tflops": ro_flops / (ro["ms_per_step"] * 1e-3) / 1e12, "share": ro.get("share")},
        "gae": {"bound": "hbm", "achieved": 17 * steps_per_iter / (gae["ms_per_step"] * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                "share": gae.get("share")},
        "loss_grad": {"bound": "tensor", "achieved": lg_flops / (lg["ms_per_step"] * 1e-3) / 1e12, "peak": pk["bf16_sustained"],
                      "unit": "TFLOP/s", "hbm_gbs": lg_bytes / (lg["ms_per_step"] * 1e-3) / 1e9,
                      "frac_of_fp32_fma_peak": lg_flops / (lg["ms_per_step"] * 1e-3) / 1e12 / fp32_peak, "share": lg.get("share"),
                      "path": lg_path,
                      "tf32_issued_tflops": lg_tensor_flops / (lg["ms_per_step"] * 1e-3) / 1e12,
                      "frac_of_tf32_peak_issued": lg_tensor_flops / (lg["ms_per_step"] * 1e-3) / 1e12 / (pk["bf16_sustained"] / 2)},
    }
    for r in rooflines.values():
        r["frac"] = r["achieved"] / r["peak"]
    dominant = max(kern, key=lambda k: kern[k]["share"])
    roof = dict(rooflines.get(dominant, rooflines["loss_grad"]))
    if lg_path == "tensor":
        note = ("loss_grad runs its 64x64 GEMMs on tcgen05.mma kind::tf32 with the 3xTF32 split (fp32-level accuracy, parity 1e-4 "
                "holds at ~2e-6); achieved = fp32-equivalent algorithmic FLOPs (fwd + 2x bwd) / time, peak = measured dense bf16 by "
                "contract; tf32 peak is half of it and 3 tensor passes are issued per algorithmic FLOP (frac_of_tf32_peak_issued); "
                "the kernel is bound by its CUDA-core phases (tanh, loss head, thin-layer gradient reductions), see DESIGN.md")
    elif lg_path == "mma":
        rooflines["loss_grad"]["frac_of_mma_sync_tf32_peak_issued"] = rooflines["loss_grad"]["tf32_issued_tflops"] / mma_tf32_peak
        roof = dict(rooflines.get(dominant, rooflines["loss_grad"]))
        note = ("general-shape loss_grad kernel: hidden GEMMs on warp-level tensor-core tiles (mma.sync m16n8k8 TF32, 3xTF32 split, "
                "fp32-level accuracy), thin layers and loss head on CUDA cores; achieved = fp32-equivalent algorithmic FLOPs / time, "
                "peak = measured dense bf16 (tcgen05 path) by contract; frac_of_mma_sync_tf32_peak_issued counts the 3 issued passes "
                "against the 512 MAC/clk/SM the legacy tensor path sustains on sm_100a (profiles/r01_mma_sync_rates.txt)")
    else:
        note = ("fp32 CUDA-core kernels; frac is against the tensor peak by contract, frac_of_fp32_fma_peak is the pipe it actually "
                "runs on")
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures of workload C2
    # (profiles/r01_lossgrad_tc_final_summary.txt, profiles/r01_rollout_tc_final_summary.txt); null for other workloads
    traffic_c2 = {"loss_grad": 14.16e6 + 0.06e6, "rollout": 192.3e3} if (args.workload == "c2" and lg_path == "tensor") else {}
    if args.workload == "c3" and lg_path == "mma":
        # profiles/r01_lossgrad_mma_summary.txt: 137.7 MB read + 32.0 MB written per launch.  The reads are the 32-byte
        # sectors of the Feistel-order gather (4-byte fields of 524 288 scattered rows, two passes), 11x the algorithmic
        # bytes but 43 GB/s, far from a limit; the writes are the red.add traffic of the dW partials
        traffic_c2 = {"loss_grad": 137.66e6 + 32.02e6}
    roof.update({"kernel": dominant, "traffic": traffic_c2.get(dominant), "peak_source": pk["source"] + (" sustained bf16" if roof["bound"] == "tensor" else " copy"),
                 "note": note})

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (on-device env batches of the named shape, orthogonal random-init weights, seed 0)",
        "config": {"workload": w["name"], "epochs": EPOCHS, "minibatches_per_epoch": N_MINIBATCHES, "batch_size_per_gpu": batch,
                   "env_steps_per_step_per_gpu": steps_per_iter, "adam_steps_per_step": EPOCHS * N_MINIBATCHES,
                   "l2": "flushed: 256 MB memset on the stream before every timed step (inside the timed region)",
                   "warmup_iterations_run": n_w,        # W requested, extended to >= 0.3 s of untimed work
                   "parallelism": f"dp{world} over envs, {comm}" if world > 1 else "single GPU"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "dril_b200.train (train!): host parameters in, per-iteration learn_stats + final parameters out, wall clock"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "rooflines": rooflines,
        "kernels": kern,
        "last_iteration": {k: (v if np.isfinite(v) else None) for k, v in st.as_dict().items()},
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.workload)
    print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the restated CPU reference (oracle/, NumPy + threaded BLAS).
# DRiL.jl itself cannot run here: there is no Julia toolchain in the image.
# ---------------------------------------------------------------------------------------------
def _oracle_iteration_rate(workload, sample_envs, iterations):
    from oracle import envs as OE, policy as OP, ppo as OO
    w = WORKLOADS[workload]
    n_steps = w["n_steps"]
    if w["kind"] == "cartpole":
        spec = OP.PolicySpec(4, w["hidden"], "discrete", 2, act_start=1)
        env = OE.MonitorWrapper(OE.ParallelEnv(OE.CartPoleBatch(sample_envs, seed=0)))
    else:
        spec = OP.PolicySpec(3, w["hidden"], "continuous", 1, act_low=[-2], act_high=[2])
        env = OE.MonitorWrapper(OE.ParallelEnv(OE.PendulumBatch(sample_envs, seed=0)))
        if w["normalize"]:
            env = OE.NormalizeWrapper(env, 3)
    flat = OP.init_params(spec, seed=0)
    cfg = OO.PPOConfig(n_steps=n_steps, batch_size=n_steps * sample_envs // N_MINIBATCHES, epochs=EPOCHS)
    OO.train(env, spec, flat, cfg, n_steps * sample_envs)              # warm-up iteration
    t0 = time.perf_counter()
    OO.train(env, spec, flat, cfg, n_steps * sample_envs * iterations)
    dt = time.perf_counter() - t0
    return n_steps * sample_envs * iterations / dt, dt


def cpu_baseline(workload):
    sample_envs = max(64, WORKLOADS[workload]["n_envs"] // 8)
    rate, dt = _oracle_iteration_rate(workload, sample_envs, 2)
    return {"value": rate, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
            "sample": f"restated reference (Python/NumPy, threaded BLAS), not DRiL.jl: 2 PPO iterations of {sample_envs} envs x "
                      f"{WORKLOADS[workload]['n_steps']} steps, epochs {EPOCHS}, {N_MINIBATCHES} minibatches ({dt:.1f} s)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    sample_envs = max(64, w["n_envs"] // 8)
    from oracle import envs as OE, policy as OP, ppo as OO  # noqa: F401
    rates = []
    _oracle_iteration_rate(args.workload, sample_envs, 1)             # warm-up
    t_all = time.perf_counter()
    for _ in range(args.steps):
        r, _ = _oracle_iteration_rate(args.workload, sample_envs, 1)
        rates.append(r)
    total = time.perf_counter() - t_all
    value = float(np.mean(rates))
    sample = (f"restated reference (Python/NumPy oracle port), not DRiL.jl (no Julia toolchain): each step = 1 PPO iteration of "
              f"{sample_envs} envs x {w['n_steps']} steps (1/8 of the per-GPU workload), epochs {EPOCHS}, {N_MINIBATCHES} minibatches")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total / max(args.steps, 1) * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "epochs": EPOCHS, "minibatches_per_epoch": N_MINIBATCHES},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.workload is None:
        args.workload = "c2"
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
