#!/usr/bin/env python
"""Benchmark of the hot path named by BASELINE.json: differentiable env-steps/s (forward + adjoint).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--env AntEnv] [--num-envs 4096]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one SHAC short-horizon rollout of the workload: `horizon` env-steps forward for `num_envs`
environments per GPU, then the adjoint of all of them (BASELINE.json configs[1]: AntEnv, 4096 envs,
horizon 32).  Prints ONE JSON line (rank 0):

* ``value``    env-steps/s over all GPUs, kernel path: inputs resident in HBM, 2*horizon launches of OUR
               kernels per step through the C ABI, device-timed with CUDA events, max over ranks.
* ``e2e``      same metric through the reference-facing API (``envs.AntEnv.step`` ->
               ``dflex.sim.SemiImplicitIntegrator.forward`` -> autograd), every rollout's actions copied
               from pinned HOST memory inside the timed region and the loss + action gradients read back.
* ``roofline`` dominant kernel (the adjoint): algorithmic bytes per launch / measured launch time vs the
               measured HBM copy bandwidth (MEASURED_PEAKS.json).  This path is FP32-issue bound, not
               HBM bound, so ``frac`` is tiny by construction; ``traffic`` comes from
               the committed ncu capture.
* ``cpu_baseline`` the reference's own compiled CPU kernels (oracle/_ref, kind "reference") on a bounded
               sample, timed on this box's host (1 core: the reference CPU path is a serial loop).
* ``--impl reference`` times only that CPU arm and prints the same line shape.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MM_FREQ = {"AntEnv": 16, "HumanoidEnv": 48, "SNUHumanoidEnv": 8, "CartPoleSwingUpEnv": 4, "HopperEnv": 16, "CheetahEnv": 16}
SUBSTEPS = {"AntEnv": 16, "HumanoidEnv": 48, "SNUHumanoidEnv": 48, "CartPoleSwingUpEnv": 4, "HopperEnv": 16, "CheetahEnv": 16}
HORIZON = {"SNUHumanoidEnv": 128}
# reference CPU sample: (num_envs, env-steps) sized for ~10-30 s of single-core work
CPU_SAMPLE = {"AntEnv": (64, 24), "HumanoidEnv": (16, 8), "SNUHumanoidEnv": (8, 6), "CartPoleSwingUpEnv": (64, 240),
              "HopperEnv": (64, 32), "CheetahEnv": (64, 32)}


def algorithmic_bytes(Q, D, A, substeps, row=None, nseg=1):
    """Bytes one env-step MUST move per environment, fp32.  With row=None: SURVEY.md section 8d's figure for a
    state-only tape (row = Q + D).  With the kernels' actual tape row (q, qd + the forward intermediates the
    adjoint reads back instead of recomputing) and the D*D H^-1 block per mass-matrix
    update: the design's own algorithmic traffic.  Returns (forward, backward)."""
    row = (Q + D) if row is None else row
    hinv = 0 if row == Q + D else nseg * D * D
    fwd = 4 * (Q + D + A) + 4 * (Q + D) + 4 * (substeps * row + hinv)
    bwd = 4 * (substeps * row + hinv) + 4 * (Q + D) + 4 * A + 4 * (Q + D + A)
    return fwd, bwd


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for k, name in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_arm(env_name, budget_steps=None):
    """Time the reference's compiled CPU kernels on a bounded sample.  Returns the cpu_baseline dict."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import torch
    import ref_driver
    if not ref_driver.available():
        return {"value": None, "unit": "env-steps/s", "cores": 0, "kind": "reference",
                "sample": "oracle/_ref/kernels.so missing (build it in the container: python oracle/make_golden.py)"}
    torch.set_num_threads(1)
    n, steps = CPU_SAMPLE[env_name]
    if budget_steps:
        steps = budget_steps
    arrays = dict(np.load(os.path.join(ROOT, "diffrl_b200", "assets", env_name + ".npz")))
    tf, tb = ref_driver.time_env_steps(arrays, n, SUBSTEPS[env_name], MM_FREQ[env_name], 1.0 / 60.0, steps,
                                       ground=bool(arrays["ground"]))
    return {"value": n * steps / (tf + tb), "unit": "env-steps/s", "cores": 1, "kind": "reference",
            "sample": "%s: %d envs x %d env-steps forward+adjoint on the reference's own generated CPU kernels "
                      "(oracle/_ref/kernels.so, serial loop) in %.1f s; host has %d cores" % (env_name, n, steps, tf + tb, os.cpu_count()),
            "seconds": tf + tb}


def _cpu_worker(env_name, reps, out_q):
    """One host process of the reference arm: `reps` timed samples on one core."""
    try:
        secs = [cpu_reference_arm(env_name)["seconds"] for _ in range(reps)]
        out_q.put(secs)
    except Exception as exc:  # pragma: no cover
        out_q.put(exc)


def run_reference(args):
    """The reference's CPU implementation of the path on this box's host cores.  Its kernels are a serial
    loop (adjoint.py:1271-1279), so "all the host threads it can use" = P independent single-threaded
    processes, each running the same bounded sample concurrently (environments are independent)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    env = args.env
    n, steps = CPU_SAMPLE[env]
    probe = cpu_reference_arm(env) if args.warmup > 0 else None
    if probe is not None and probe["value"] is None:
        print(json.dumps({"impl": "reference", "unavailable": probe["sample"]}))
        return
    procs = max(1, min(args.cpu_procs or (os.cpu_count() or 1), 64))
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    workers = [ctx.Process(target=_cpu_worker, args=(env, args.steps, out_q)) for _ in range(procs)]
    for w in workers:
        w.start()
    results = [out_q.get() for _ in workers]
    for w in workers:
        w.join()
    for r in results:
        if isinstance(r, Exception):
            print(json.dumps({"impl": "reference", "unavailable": "worker failed: %r" % (r,)}))
            return
    per_step = [max(r[i] for r in results) for i in range(args.steps)]   # all workers run step i concurrently
    times = per_step
    value = procs * n * steps * len(times) / sum(times)
    base = {"value": value, "unit": "env-steps/s", "cores": procs, "kind": "reference",
            "sample": "%s: %d single-threaded processes x (%d envs x %d env-steps forward+adjoint) on the reference's own "
                      "generated CPU kernels (oracle/_ref/kernels.so); host has %d cores" % (env, procs, n, steps, os.cpu_count())}
    line = {"impl": "reference", "metric": "differentiable env-steps/s (fwd+bwd)", "value": value, "unit": "env-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s reference dflex CPU path, bounded sample: %d processes x %d envs x %d env-steps per step" % (env, procs, n, steps)},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION) off it
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    from diffrl_b200 import _capi
    import diffrl_b200.envs as envs
    from diffrl_b200.dflex_api.sim import _engine_for

    env_name, N, T = args.env, args.num_envs, args.horizon or HORIZON.get(args.env, 32)
    S, mm = SUBSTEPS[env_name], MM_FREQ[env_name]
    torch.manual_seed(1234 + rank)
    env = getattr(envs, env_name)(num_envs=N, device=str(dev), render=False, seed=rank, stochastic_init=False,
                                  no_grad=False, MM_caching_frequency=mm)
    eng = _engine_for(env.model)
    Q, D, M = eng.Q, eng.D, eng.M
    A = D + M
    lib = _capi.lib()
    dt = env.sim_dt

    # ------------------------------------------------------------ kernel path (inputs resident in HBM)
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    env.clear_grad(); env.reset()
    q0, qd0 = env.state.joint_q.detach().clone(), env.state.joint_qd.detach().clone()
    acts = [torch.zeros(N * D, device=dev) for _ in range(T)]
    muscs = [None] * T
    for t in range(T):
        a = torch.rand((N, env.num_actions), generator=g, device=dev) * 2 - 1
        if env_name == "SNUHumanoidEnv":
            muscs[t] = ((a * 0.5 + 0.5).view(-1) * env.muscle_strengths).contiguous()
        else:
            env.state.joint_act.zero_()
            env._apply_actions(a)
            acts[t] = env.state.joint_act.detach().clone()
    gq_seed, gqd_seed = torch.randn(N * Q, device=dev), torch.randn(N * D, device=dev)

    phase_events = []      # (start, forward done, backward done) of every timed rollout, on the launching stream

    def kernel_rollout(record=False):
        ev3 = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if record else None
        if record:
            ev3[0].record()
        q, qd, tapes = q0, qd0, []
        for t in range(T):
            q, qd, tape, _ = eng.forward(q, qd, acts[t], muscs[t], S, mm, dt)
            tapes.append(tape)
        if record:
            ev3[1].record()
        gq, gqd = gq_seed, gqd_seed
        for t in reversed(range(T)):
            gq, gqd, gact, gm = eng.backward(acts[t], muscs[t], tapes[t], gq, gqd, S, mm, dt)
        if record:
            ev3[2].record()
            phase_events.append(ev3)
        return gq

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        kernel_rollout()
    barrier()
    if args.ncu_range:
        torch.cuda.profiler.start()
        kernel_rollout()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    launches0 = lib.dfx_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with ClockSampler(local) as clocks:
        ev[0].record()
        for _ in range(args.steps):
            kernel_rollout(record=True)
        ev[1].record()
        barrier()
    kernel_ms = ev[0].elapsed_time(ev[1])
    launches = lib.dfx_launch_count() - launches0

    # average launch duration of the forward and of the adjoint kernel INSIDE the timed rollouts (every launch reads /
    # writes its own tape block, gigabytes per rollout, far beyond L2): the roofline denominators
    fwd_ms = sum(e[0].elapsed_time(e[1]) for e in phase_events) / (len(phase_events) * T)
    bwd_ms = sum(e[1].elapsed_time(e[2]) for e in phase_events) / (len(phase_events) * T)

    # ------------------------------------------------------------ end to end through env.step + autograd
    host_actions = torch.empty((T, N, env.num_actions), dtype=torch.float32).pin_memory()
    host_actions.copy_(torch.rand(host_actions.shape) * 2 - 1)
    host_grad = torch.empty((T, N, env.num_actions), dtype=torch.float32).pin_memory()
    host_loss = torch.empty((), dtype=torch.float32).pin_memory()
    comm = torch.zeros(16384, device=dev)   # size of the Ant actor's flattened gradient (cfg/shac/ant.yaml)

    def e2e_rollout():
        env.clear_grad()
        env.reset()
        env.initialize_trajectory()
        a_dev = host_actions.to(dev, non_blocking=True).requires_grad_()
        loss = torch.zeros((), device=dev)
        for t in range(T):
            obs, rew, done, _ = env.step(a_dev[t])
            loss = loss + rew.sum()
        loss.backward()
        if world > 1:
            # the per-rollout policy-gradient all-reduce of the data-parallel SHAC actor update; the
            # synthetic benchmark has no policy, so a buffer of the Ant actor's size carries the
            # action-gradient summary
            comm[: T * env.num_actions] = a_dev.grad.mean(dim=1).reshape(-1)
            dist.all_reduce(comm)
        host_grad.copy_(a_dev.grad, non_blocking=True)
        host_loss.copy_(loss.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(host_loss)

    for _ in range(max(1, args.warmup // 2)):
        e2e_rollout()
    barrier()
    e3 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e2e_steps = max(1, args.steps // 2)
    e3[0].record()
    for _ in range(e2e_steps):
        e2e_rollout()
    e3[1].record()
    barrier()
    eager_ms = e3[0].elapsed_time(e3[1])

    # the same rollout through the package's graphed-rollout API (one CUDA graph per rollout: H2D actions,
    # horizon x env.step, loss.backward, D2H loss + action gradients)
    e2e_ms, e2e_api = eager_ms, "envs.%s.step -> dflex.sim.SemiImplicitIntegrator.forward -> autograd (eager)" % env_name
    if args.e2e != "eager" and hasattr(env, "_start_state"):
        from diffrl_b200.rollout import GraphedRollout
        env.clear_grad(); env.reset(); env.initialize_trajectory()
        roll = GraphedRollout(env, T)
        roll.host_actions.copy_(host_actions)      # the pinned staging buffer the graph's H2D copy reads every rollout

        def graphed_rollout():
            env.clear_grad()
            env.reset()
            loss_h, grad_h = roll(None, sync=False)
            if world > 1:
                comm[: T * env.num_actions] = roll.actions.grad.mean(dim=1).reshape(-1)
                dist.all_reduce(comm)
            torch.cuda.current_stream().synchronize()
            return float(loss_h)

        for _ in range(max(1, args.warmup // 2)):
            graphed_rollout()
        barrier()
        if args.ncu_range_e2e:          # launch list of ONE end-to-end rollout (profiles/): ncu --profile-from-start off
            torch.cuda.profiler.start()
            graphed_rollout()
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        e3[0].record()
        for _ in range(e2e_steps):
            graphed_rollout()
        e3[1].record()
        barrier()
        e2e_ms = e3[0].elapsed_time(e3[1])
        e2e_api = "diffrl_b200.rollout.GraphedRollout(envs.%s): one CUDA graph = H2D actions + %d x env.step + backward + D2H" % (env_name, T)

    # ------------------------------------------------------------ reduce over ranks (max time)
    times = torch.tensor([kernel_ms, e2e_ms, eager_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    kernel_ms, e2e_ms, eager_ms = times.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_env_steps = world * N * T * args.steps
    value = total_env_steps / (kernel_ms * 1e-3)
    e2e_value = world * N * T * e2e_steps / (e2e_ms * 1e-3)
    peaks, peak_kind = measured_peaks()
    row = lib.dfx_pack_query(eng.pack, 8)   # DFX_QUERY_TAPE_ROW_FLOATS
    b_fwd, b_bwd = algorithmic_bytes(Q, D, A, S, row=row, nseg=(S + mm - 1) // mm)
    s_fwd, s_bwd = algorithmic_bytes(Q, D, A, S)
    achieved = N * b_bwd / (bwd_ms * 1e-3) / 1e9
    traffic, fp32 = None, None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            prof = json.load(f).get(env_name, {})
        traffic = prof.get("bwd_dram_bytes_per_launch")
        if "bwd_fp32_flop_per_launch" in prof and N == 4096:
            # SURVEY.md 8d's second yardstick: fp32 FLOP actually executed (counted by ncu for this launch shape) per
            # second against the CUDA-core peak 148 SMs x 128 lanes x 2 FLOP x the SM clock sampled during the run
            flop = prof["fwd_fp32_flop_per_launch"] + prof["bwd_fp32_flop_per_launch"]
            mhz = clocks.summary().get("sm_mhz") or 1965.0
            peak = 148 * 128 * 2 * mhz * 1e6
            fp32 = {"achieved_tflops": flop / ((fwd_ms + bwd_ms) * 1e-3) / 1e12, "peak_tflops": peak / 1e12,
                    "frac": flop / ((fwd_ms + bwd_ms) * 1e-3) / peak, "flop_per_env_step": flop / N,
                    "source": "profiles/r01_traffic.json (ncu instruction counts)"}
    cpu = cpu_reference_arm(env_name) if world == 1 and not args.no_cpu_baseline else None
    line = {
        "metric": "differentiable env-steps/s (fwd+bwd)", "value": value, "unit": "env-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": kernel_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s num_envs=%d/GPU SHAC short-horizon=%d (rollout forward + adjoint), substeps=%d, "
                               "mass_matrix_freq=%d" % (env_name, N, T, S, mm),
                   "parallelism": "env-sharded x%d, no data-path collective%s" % (world, "; 64 KB policy-gradient all-reduce per rollout (e2e)" if world > 1 else ""),
                   "cache": "per-rollout tape %.0f MB > 126 MB L2 (inputs larger than L2)" % (T * eng.tape_floats(S, mm) * 4 / 1e6),
                   "kernel_family": "tile (32 envs per CTA, lane = env)" if lib.dfx_pack_query(eng.pack, 9) else "lane group"},
        "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": int(host_actions.numel() * 4),
                "d2h_bytes_per_step": int(host_grad.numel() * 4 + 4), "api": e2e_api,
                "ms_per_step": e2e_ms / e2e_steps,
                "eager_env_step_loop": {"value": world * N * T * e2e_steps / (eager_ms * 1e-3), "ms_per_step": eager_ms / e2e_steps}},
        "gpu_launches": int(launches),
        "kernel_ms": {"forward_env_step": fwd_ms, "backward_env_step": bwd_ms},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_kind,
                     "kernel": ("dfx_tile_kernel<NW,BWD=1> (adjoint of one env-step, 32-environment tiles)"
                                if lib.dfx_pack_query(eng.pack, 9) else "dfx_step_kernel<G,BWD=1> (adjoint of one env-step)"),
                     "algorithmic_bytes_per_launch": N * b_bwd,
                     "survey_8d_state_only_bytes_per_launch": N * s_bwd,
                     "forward_kernel": {"achieved": N * b_fwd / (fwd_ms * 1e-3) / 1e9, "algorithmic_bytes_per_launch": N * b_fwd},
                     "fp32": fp32,
                     "note": "algorithmic bytes = this design's tape rows (q, qd + forward intermediates, 4*row B per env-substep) "
                             "+ H^-1 blocks + state I/O; the fused path is FP32-issue/latency bound, not HBM bound "
                             "(profiles/r01_ant_full.md, DESIGN.md section 3); launch durations are averages over the timed rollouts"},
        "clocks": clocks.summary(),
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--env", default="AntEnv", choices=sorted(SUBSTEPS))
    ap.add_argument("--num-envs", type=int, default=4096)
    ap.add_argument("--horizon", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e", default="graph", choices=["graph", "eager"])
    ap.add_argument("--cpu-procs", type=int, default=0, help="reference arm: host processes (default: all cores, max 64)")
    ap.add_argument("--ncu-range-e2e", action="store_true",
                    help="wrap one graphed end-to-end rollout in cudaProfilerStart/Stop (for ncu --profile-from-start off)")
    ap.add_argument("--ncu-range", action="store_true",
                    help="wrap ONE kernel-path step and ONE e2e step in cudaProfilerStart/Stop (use with ncu --profile-from-start off)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
