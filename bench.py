#!/usr/bin/env python3
"""bench.py -- headline benchmark of the SSE hot path: trajectory-control-steps/sec on N B200s.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA path through the C-ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU algorithm (oracle port) on the host cores

A "step" = one control step of every trajectory of the batch: n_sub SSE substeps + moment extraction + flags, ONE kernel
launch per rank.  Workload at N=1 = BASELINE.json configs[1]: quartic oscillator cooling, 1024 trajectories (N=171 grid
points, 80 substeps per control step); N>1 = the same per-GPU batch on every rank (weak scaling), trajectories sharded with no
data-path collective plus the exchange of the moment/reward block (stored by the SSE kernel into every rank's peer memory).
Inside the same driver-timed run the line also carries, under "extra", the other BASELINE configurations measured the same way
(CUDA events, L2 flushed between steps, max over ranks): config3 (inverted harmonic, 8192 trajectories/GPU), config4 (inverted
quartic, 8192 trajectories/GPU = 65,536 on 8 GPUs, with its own roofline) and five points of the grid-size sweep (N = 257 ... 8193).
Prints ONE JSON line (rank 0).
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

# algorithmic FP64 flops per trajectory-control-step (n_sub * F_sub * N + moments)
F_SUB = {"harmonic": 316, "inverted_harmonic": 364, "quartic": 568, "inverted_quartic": 568}
MOM_FLOPS = {"harmonic": 60, "inverted_harmonic": 60, "quartic": 250, "inverted_quartic": 250}


def flops_per_unit(task, n, n_sub):
    return float(n_sub * F_SUB[task] * n + MOM_FLOPS[task] * n)


def state_len(params):
    if "n_max" in params:
        return params["n_max"] + 1
    return 2 * int(params["x_max"] / params["grid_size"] + 0.5) + 1


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_port_rate(task, params, n_traj_per_thread, threads, fast=True, constant_force=False):
    """The oracle (CPU restatement of the reference algorithm) timed the way the reference runs: one trajectory per
    single-threaded worker, `threads` workers.  Returns (traj-control-steps/s, seconds)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import oracle_for, level_force
    from deepreinforcementlearningcontrolofquantumcartpoles_b200.states import initial_states
    psi0 = initial_states(params, threads * n_traj_per_thread, seed=3)
    rng = np.random.default_rng(7)
    noise = rng.standard_normal((threads * n_traj_per_thread, params["n_sub"], 2))
    acts = rng.integers(0, params["n_levels"], threads * n_traj_per_thread)
    if constant_force:                       # every worker keeps one force: reset_ab (LU + C = dt^3/12 H0^2 ...) runs once, outside the measurement
        acts[:] = 13
    oracles = [oracle_for(params, fast=fast) for _ in range(threads)]
    for t, o in enumerate(oracles):      # warm the 21-force cache is NOT done: the reference refactorises on every force change too
        st = psi0[t * n_traj_per_thread].copy()
        o.run(st, params["dt"], level_force(params, 13) if constant_force else 0.0, params["gamma"], noise[0][:2])

    def work(t):
        o = oracles[t]
        for j in range(n_traj_per_thread):
            b = t * n_traj_per_thread + j
            st = psi0[b].copy()
            o.run(st, params["dt"], level_force(params, int(acts[b])), params["gamma"], noise[b])     # ctypes releases the GIL

    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    return threads * n_traj_per_thread / dt, dt


def reference_build_call_pattern(task, params, n_ctrl=3):
    """The reference's own simulation*.cpp (built against the MKL-API shim, oracle/_ref) driven the way the reference drives it:
    one Python->C call per substep, force change (-> its reset_ab) once per control step.  Single process; the reference runs one such
    process per core.  Informational: the shim is slower than real MKL, so this is NOT the baseline the speed-up is quoted against."""
    try:
        import numpy as np
        from oracle.ref_module import RefModule, available
        if not available(task):
            return None
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from common import level_force
        from deepreinforcementlearningcontrolofquantumcartpoles_b200.states import initial_states
        ref = RefModule(task)
        psi = initial_states(params, 1, seed=5)[0]
        rng = np.random.default_rng(1)
        ref.step(psi, params["dt"], 0.0, params["gamma"], rng.standard_normal(2))
        t0 = time.perf_counter()
        for c in range(n_ctrl):
            F = level_force(params, int(rng.integers(0, params["n_levels"])))
            for s in range(params["n_sub"]):
                ref.step(psi, params["dt"], F, params["gamma"], rng.standard_normal(2))
        dt = time.perf_counter() - t0
        return {"value_per_core": n_ctrl / dt, "unit": "traj-control-steps/s", "kind": "reference sources + MKL-API shim, Python call per substep (as shipped)",
                "sample": "%d control steps, 1 process" % n_ctrl}
    except Exception as e:      # never let the informational leg break the bench
        return {"error": repr(e)}


def run_reference(args, task, params):
    """--impl reference: the reference's own CPU algorithm for the path on all host cores.  The MKL original cannot be built
    here (no MKL), so this is the oracle port (kind "port"); each step = a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_thread = args.ref_traj_per_thread
    if per_thread <= 0:
        r0, s0 = cpu_port_rate(task, params, 1, cores)          # calibration (also warms caches / builds the oracle)
        per_thread = max(1, int(2.0 / max(s0, 1e-3)))
    rates, secs = [], []
    for i in range(args.warmup + args.steps):
        r, s = cpu_port_rate(task, params, per_thread, cores)
        if i >= args.warmup:
            rates.append(r); secs.append(s)
    total_units = args.steps * cores * per_thread
    value = total_units / sum(secs)
    n = state_len(params)
    line = {"impl": "reference", "metric": "trajectory-control-steps/sec", "value": value, "unit": "traj-control-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(task, args.batch, n, params["n_sub"]), "task": task, "trajectories_per_gpu": args.batch, "global_trajectories": args.gpus * args.batch,
                       "state_len": n, "n_sub": params["n_sub"]},
            "measurement": {"sample": "%d trajectories x 1 control step per timed step (bounded sample of the workload above; the rate is per unit of work)" % (cores * per_thread)},
            "cpu_baseline": {"value": value, "unit": "traj-control-steps/s", "cores": cores, "kind": "port",
                             "sample": "%d threads x %d trajectories x 1 control step (%d substeps, N=%d), oracle/sse_oracle.c -Ofast, per-force refactorisation included" % (cores, per_thread, params["n_sub"], n)},
            "e2e": {"value": value, "unit": "traj-control-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def closed_loop_rate(sim, params, B, steps, warmup, out):
    """SURVEY 8f rows 2-3 measured beside the SSE step: observation -> direct_DQN (noisy nets, in-kernel Philox noise) -> epsilon-greedy ->
    SSE control step -> experience row, everything resident on the device (random-init weights of the reference's architecture)."""
    import numpy as np
    import torch
    from deepreinforcementlearningcontrolofquantumcartpoles_b200 import rollout as R, _lib as L
    rng = np.random.default_rng(0)
    pol = R.DirectDQNPolicy(sim.K, params["n_levels"])
    for idx, name in enumerate(L.POLICY_PARAMS):
        size = int(pol.lib.qc_policy_param_size(pol.h, idx))
        lo, hi = (0.01, 0.03) if name.endswith(("SW", "SB")) else (-0.05, 0.05)
        pol.set_param(name, rng.uniform(lo, hi, size).astype(np.float32))
    ring = R.ReplayRing(sim.K, max(4 * B, 1 << 16))
    obs = R.observation(out["moments"], 1.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = pol.launch_count()
    for i in range(warmup + steps):
        if i == warmup:
            torch.cuda.synchronize()
            e0.record()
        greedy = pol.forward(obs, noise="philox", seed=0, traj_offset=sim.traj_offset, counter=i, want_q=False)["greedy"]
        action, _ = pol.epsilon_greedy(greedy, 0.05, seed=0, traj_offset=sim.traj_offset, counter=i)
        sim.step(action, out=out)
        obs2 = R.observation(out["moments"], 1.0)
        ring.push(obs, obs2, action, out["aux"], reward_scale=-1.0, reward_stride=L.QC_AUX_COUNT)
        obs = obs2
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    p0.record()
    for i in range(steps):
        greedy = pol.forward(obs, noise="philox", seed=0, traj_offset=sim.traj_offset, counter=i, want_q=False)["greedy"]
        pol.epsilon_greedy(greedy, 0.05, seed=0, traj_offset=sim.traj_offset, counter=i)
    p1.record()
    torch.cuda.synchronize()
    return {"value": B / (ms * 1e-3), "unit": "traj-control-steps/s", "ms_per_step": ms, "policy_ms_per_step": p0.elapsed_time(p1) / steps,
            "policy": "direct_DQN %d-512-512-256-%d, factorised noisy layers; 512-wide layers as 3xTF32 on tcgen05 tensor cores, rest fp32 FMA" % (sim.K, params["n_levels"]),
            "policy_launches_per_step": (pol.launch_count() - l0) / (warmup + 2 * steps) + 1, "experience_rows": ring.total(),
            "note": "no L2 flush; device-resident loop, no host synchronisation inside the timed region"}


def workload_name(task, B, n, n_sub):
    return "%s SSE control step: %d trajectories/GPU, N=%d complex128, %d substeps/control step, 21 force levels" % (task, B, n, n_sub)


def median5(per_step_ms):
    """Median of the mean step time of 5 equal chunks of the timed region (SURVEY.md 8d asks for a median of 5)."""
    k = len(per_step_ms)
    if k < 5:
        return None
    c = k // 5
    means = sorted(sum(per_step_ms[i * c:(i + 1) * c]) / c for i in range(5))
    return means[2]


def measure_extra(task, params, B, steps, warmup, rank, world, local_rank, flush_buf, fp64_peak, tag):
    """One more workload measured like the headline: device-resident inputs, one launch per control step, per-step CUDA events with the L2
    flushed in between, max over ranks.  No exchange (the step alone); every rank runs its own shard of B trajectories."""
    import numpy as np
    import torch
    from deepreinforcementlearningcontrolofquantumcartpoles_b200 import BatchedSim, _lib as L
    from deepreinforcementlearningcontrolofquantumcartpoles_b200.states import initial_states
    dev = "cuda:%d" % local_rank
    n = state_len(params)
    sim = BatchedSim(params, batch=B, device=local_rank, seed=1, traj_offset=rank * B)
    psi0 = initial_states(params, min(B, 64), seed=100 + rank)
    sim.set_state(np.tile(psi0, ((B + psi0.shape[0] - 1) // psi0.shape[0], 1))[:B])
    g = torch.Generator(device=dev); g.manual_seed(1000 + rank)
    actions = torch.randint(0, params["n_levels"], (warmup + steps, B), device=dev, dtype=torch.int32, generator=g)
    out = sim.alloc_outputs()
    for i in range(warmup):
        sim.step(actions[i], out=out)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    l0 = sim.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i in range(steps):
        if flush_buf is not None:
            flush_buf.zero_()
        evs[i][0].record()
        sim.step(actions[warmup + i], out=out)
        evs[i][1].record()
    torch.cuda.synchronize()
    per = [a.elapsed_time(b) for a, b in evs]
    ms = sum(per) / steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    fpu = flops_per_unit(task, n, params["n_sub"])
    ach = B / (ms * 1e-3) * fpu / 1e12
    traffic = None
    try:
        static = json.load(open(os.path.join(ROOT, "profiles", "roofline_static.json")))
        if tag.startswith("BASELINE configs[4]"):              # sweep points: captured per (N, batch)
            traffic = static.get("sweep_N%d_B%d" % (n, B), {}).get("dram_bytes_per_launch")
        elif B == 8192:
            traffic = static.get(task, {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    rec = {"workload": workload_name(task, B, n, params["n_sub"]) + (" [%s]" % tag if tag else ""), "value": world * B / (ms * 1e-3), "unit": "traj-control-steps/s",
           "n_gpus": world, "global_trajectories": world * B, "steps": steps, "warmup": warmup, "ms_per_step": ms, "gpu_launches": int(sim.launch_count() - l0),
           "max_norm_deviation": float((out["aux"][:, L.QC_AUX_NORM] - 1).abs().max().item()),
           "roofline": {"bound": "fp64", "achieved": ach, "peak": fp64_peak / 1e12, "unit": "TFLOP/s", "frac": ach / (fp64_peak / 1e12), "traffic": traffic, "flops_per_unit": fpu,
                        "kernel": sim.kernel_info(), "hbm_algorithmic_bytes_per_launch": B * (32.0 * n + 8.0 * sim.K + 37), "hbm_algorithmic_GBps": B * (32.0 * n + 8.0 * sim.K + 37) / (ms * 1e-3) / 1e9}}
    del sim
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--task", default="quartic", help="quartic (BASELINE configs[1], default) | inverted_quartic | harmonic | inverted_harmonic")
    ap.add_argument("--batch", type=int, default=1024, help="trajectories per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-traj-per-thread", type=int, default=0, help="0 = sized for ~2 s of CPU work per timed step")
    ap.add_argument("--no-l2-flush", action="store_true")
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"],
                    help="N>1: how the per-step result block reaches every rank: stored by the SSE kernel into peer memory (fused) or pack + NCCL all-gather")
    ap.add_argument("--no-closed-loop", action="store_true", help="skip the policy + experience-row closed-loop measurement (N=1 only)")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra BASELINE configurations (config3, config4, sweep)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    from deepreinforcementlearningcontrolofquantumcartpoles_b200 import configs
    task = args.task
    params = configs.PRESETS[task]()
    if args.impl == "reference":
        run_reference(args, task, params)
        return

    import numpy as np
    import torch
    from deepreinforcementlearningcontrolofquantumcartpoles_b200 import BatchedSim, measure_peaks, dist as qdist, _lib as L
    from deepreinforcementlearningcontrolofquantumcartpoles_b200.states import initial_states

    # stdout carries exactly ONE line, the JSON record: libraries that write to file descriptor 1 (NCCL prints its version there) are
    # diverted to stderr for the whole run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    rank, local_rank, world = qdist.init_process_group()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the SSE hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = "cuda:%d" % local_rank
    B = args.batch
    n = state_len(params)
    K_steps, W_steps = args.steps, args.warmup

    sim = BatchedSim(params, batch=B, device=local_rank, seed=0, traj_offset=rank * B)
    psi0 = initial_states(params, min(B, 256), seed=rank)
    sim.set_state(np.tile(psi0, ((B + psi0.shape[0] - 1) // psi0.shape[0], 1))[:B])
    g = torch.Generator(device=dev); g.manual_seed(rank)
    actions = torch.randint(0, params["n_levels"], (W_steps + K_steps, B), device=dev, dtype=torch.int32, generator=g)   # redrawn each control step
    out = sim.alloc_outputs()
    gathered = torch.empty((world * B, sim.K + 5), dtype=torch.float64, device=dev) if world > 1 else None
    fused, gather_mode = None, ("nccl" if world > 1 else None)
    if world > 1 and args.gather == "fused":
        try:                                            # CUDA IPC can be unavailable in some containers: then every rank uses the NCCL path
            fused = qdist.FusedGather(sim, rank, world)
            ok = torch.ones(1, device=dev)
        except Exception as e:                          # noqa: BLE001
            sys.stderr.write("rank %d: fused result exchange unavailable (%s); using NCCL all-gather\n" % (rank, e))
            ok = torch.zeros(1, device=dev)
        torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN)
        if ok.item() < 1:
            if fused is not None:
                fused.close(collective=False)       # never used: no barriers (the failing ranks have nothing to close)
            fused = None
        gather_mode = "fused" if fused is not None else "nccl (fused unavailable)"
    consumed = torch.zeros(1, dtype=torch.float64, device=dev)
    main_stream = torch.cuda.current_stream()
    side = torch.cuda.Stream(device=dev) if world > 1 else None          # the learner's side: waits for and reads the gathered blocks

    def exchange():
        """Every rank obtains the [world*B, K+5] result block of a control step and reads it (the learner's consumption, here a checksum).
        Fused: the rows of step k are stored by the SSE kernel itself into its own rank's area.  The consumer (wait for every rank's flag of
        step k-1, pull the peers' rows over NVLink, read the block) runs on a second stream behind the launch of step k; step k+1 waits for
        it (four buffers make that safe, include/qcart.h).  The simulation never idles until the slowest rank has finished the current step:
        the exchange has to keep up with the stepping, it does not add latency to it."""
        if fused is not None:
            seq = fused.seq()
            if seq > 1:
                with torch.cuda.stream(side):
                    fused.wait(seq - 1)
                    consumed.add_(fused.block(seq - 1)[:, 0].sum())
                    ev = torch.cuda.Event(); ev.record(side)
                main_stream.wait_event(ev)              # step seq+1 runs behind the consumer of step seq-1 (buffer-reuse rule of include/qcart.h)
        elif world > 1:
            qdist.all_gather_block(qdist.pack_block(out["moments"], out["aux"], out["flags"]), world, gathered)
            consumed.add_(gathered[:, 0].sum())
    flush_buf = None if args.no_l2_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    # roofline denominators measured in-process (MEASURED_PEAKS.json has no FP64 / shared-memory entry)
    fp64_peak, smem_peak = measure_peaks(local_rank)

    for i in range(W_steps):
        sim.step(actions[i], out=out)
        exchange()
    torch.cuda.synchronize()
    launches0 = sim.launch_count()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K_steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K_steps)]
    t_wall0 = time.perf_counter()
    for i in range(K_steps):
        if flush_buf is not None:
            flush_buf.zero_()                           # L2 flush between timed iterations (untimed)
        e0, e1 = evs[i]
        e0.record()
        kev[i][0].record()
        sim.step(actions[W_steps + i], out=out)
        kev[i][1].record()
        exchange()
        e1.record()
    if fused is not None:                               # the last step's rows (outside the per-step events: nothing left to overlap them with)
        with torch.cuda.stream(side):
            fused.wait()
            consumed.add_(fused.block()[:, 0].sum())
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    launches = sim.launch_count() - launches0
    per_step = [a.elapsed_time(b) for a, b in evs]
    ms_total = sum(per_step)
    ms_kernel = sum(a.elapsed_time(b) for a, b in kev) / K_steps
    ms_med5 = median5(per_step)
    if world > 1:
        t = torch.tensor([ms_total, ms_med5 or 0.0], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms_total, ms_med5 = float(t[0].item()), (float(t[1].item()) or None)
    value = world * B * K_steps / (ms_total * 1e-3)
    norm_dev = float((out["aux"][:, L.QC_AUX_NORM] - 1).abs().max().item())
    gather_ok, gather_failed = None, None
    if fused is not None:                               # the fused block of the last step must equal a plain NCCL all-gather of the same outputs
        qdist.all_gather_block(qdist.pack_block(out["moments"], out["aux"], out["flags"]), world, gathered)
        gather_ok = bool(torch.equal(fused.block(), gathered))
        gather_failed = fused.failed_ranks()

    # ---- end-to-end through the host-buffer C-ABI call (pinned host memory, H2D + kernel + D2H per step; N>1: plus the exchange) ----
    act_host = actions.cpu().pin_memory()
    mom_h = torch.empty((B, sim.K), dtype=torch.float64).pin_memory()
    aux_h = torch.empty((B, L.QC_AUX_COUNT), dtype=torch.float64).pin_memory()
    flg_h = torch.empty((B,), dtype=torch.uint8).pin_memory()

    def e2e_step(i):
        sim.step_host(act_host[i], moments=mom_h, aux=aux_h, flags=flg_h)       # synchronous: results are in host memory on return
        if fused is not None:                           # every rank also needs all other ranks' rows of this step
            fused.wait()
            consumed.add_(fused.block()[:, 0].sum())
            torch.cuda.current_stream().synchronize()
        elif world > 1:
            exchange()
            torch.cuda.current_stream().synchronize()
    for i in range(min(3, W_steps)):
        e2e_step(i)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    t0 = time.perf_counter()
    for i in range(K_steps):
        e2e_step(W_steps + i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * B * K_steps / e2e_s
    if fused is not None:
        fused.close()
        sim.step(actions[0], out=out)

    closed_loop = None
    if world == 1 and not args.no_closed_loop:
        closed_loop = closed_loop_rate(sim, params, B, K_steps, min(W_steps, 10), out)
    main_kernel_info = sim.kernel_info()
    K_mom = sim.K
    del sim

    # ---- the other BASELINE configurations, inside the same driver-timed run ----
    extra = {}
    if not args.no_extra:
        xs, xw = max(5, min(20, K_steps // 10)), 3
        extra["config3"] = measure_extra("inverted_harmonic", configs.inverted_harmonic(), 8192, xs, xw, rank, world, local_rank, flush_buf, fp64_peak,
                                         "BASELINE configs[2]; herm_mode 0 = literal HERMITIAN/UPPER application of C (I:23,551); mode 1 differs by < 1e-7 after 5 substeps at F_max, which of the two MKL computes is unpinned")
        extra["config4"] = measure_extra("inverted_quartic", configs.inverted_quartic(), 8192, xs, xw, rank, world, local_rank, flush_buf, fp64_peak,
                                         "BASELINE configs[3]: 65,536 trajectories on 8 GPUs = 8192 per GPU")
        # grid-size sweep (BASELINE configs[4]): every rank steps its own batch of each size (weak scaling like the headline), so the 1/2/4/8-GPU
        # runs of the driver carry it too; ~0.3 s of device time per rank for all five sizes
        extra["sweep"] = [measure_extra("inverted_quartic", configs.quartic_sweep(npts), Bs, 3, 2, rank, world, local_rank, flush_buf, fp64_peak,
                                        "BASELINE configs[4], x_max 13, dt ~ h^2") for npts, Bs in ((257, 4096), (1025, 2048), (2049, 1024), (4097, 296), (8193, 148))]

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    fpu = flops_per_unit(task, n, params["n_sub"])
    achieved = B / (ms_kernel * 1e-3) * fpu / 1e12            # per GPU, dominant (only) kernel
    prof = {}
    pj = os.path.join(ROOT, "profiles", "roofline_static.json")
    if os.path.exists(pj):
        try:
            prof = json.load(open(pj)).get(task, {})
        except Exception:
            prof = {}
    hbm_bytes = B * (32.0 * n + 8.0 * K_mom + 8.0 * L.QC_AUX_COUNT + 5)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    roofline = {"bound": "fp64", "achieved": achieved, "peak": fp64_peak / 1e12, "unit": "TFLOP/s", "frac": achieved / (fp64_peak / 1e12),
                "traffic": prof.get("dram_bytes_per_launch"),
                "kernel": main_kernel_info, "kernel_ms": ms_kernel, "flops_per_unit": fpu,
                "peak_source": "measured in-process: dependency-free DFMA loop on all SMs (qc_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 entry",
                "smem": {"peak_TBps": smem_peak / 1e12, "bytes_per_launch": prof.get("smem_bytes_per_launch"),
                         "achieved_TBps": (prof.get("smem_bytes_per_launch") / (ms_kernel * 1e-3) / 1e12) if prof.get("smem_bytes_per_launch") else None},
                "hbm": {"algorithmic_bytes_per_launch": hbm_bytes, "achieved_GBps": hbm_bytes / (ms_kernel * 1e-3) / 1e9,
                        "peak_GBps": peaks.get("hbm_gbs", 6650.0), "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cpu_port_rate(task, params, 1, cores)                  # builds / loads the oracle, warms the caches
        _, s0 = cpu_port_rate(task, params, 16, cores)
        per_thread = max(2, int(16 * 3.0 / max(s0, 1e-3)))     # five (+ three) samples of ~3 s each
        runs = sorted(cpu_port_rate(task, params, per_thread, cores) for _ in range(5))
        v, secs = runs[2]
        runs_cf = sorted(cpu_port_rate(task, params, per_thread, cores, constant_force=True) for _ in range(3))
        v_cf = runs_cf[1][0]
        cpu_baseline = {"value": v, "unit": "traj-control-steps/s", "cores": cores, "kind": "port",
                        "value_per_core": v / cores,
                        "value_per_core_without_reset_ab": v_cf / cores,
                        "reset_ab_share": max(0.0, 1.0 - v / v_cf),
                        "sample": "median of 5 runs of %d threads x %d trajectories x 1 control step (%d substeps, N=%d), %.1f s each; oracle/sse_oracle.c built -Ofast (the reference's flag); "
                                  "force redrawn per control step, so the per-force refactorisation (reset_ab: band LU + H0^2..H0^5) is included as in the reference; "
                                  "value_per_core_without_reset_ab = same with one force per worker (median of 3)" % (cores, per_thread, params["n_sub"], n, secs),
                        "reference_build_as_shipped": reference_build_call_pattern(task, params)}

    line = {"metric": "trajectory-control-steps/sec", "value": value, "unit": "traj-control-steps/s", "n_gpus": world, "steps": K_steps,
            "warmup": W_steps, "ms_per_step": ms_total / K_steps, "ms_per_step_median_of_5": ms_med5, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(task, B, n, params["n_sub"]), "task": task, "trajectories_per_gpu": B, "global_trajectories": world * B,
                       "state_len": n, "n_sub": params["n_sub"]},
            "measurement": {"noise": "in-kernel Philox4x32-10 + Box-Muller",
                       "actions": "uniform over 21 levels, redrawn every control step (torch.Generator seed 0)",
                       "l2": "inputs (2.8 MB/GPU) fit L2; L2 flushed with a 256 MiB write between timed steps, per-step CUDA events summed" if flush_buf is not None else "no flush",
                       "parallelism": ("%d rank(s), trajectories sharded, [B,%d] f64 result block per step %s" % (world, K_mom + 5,
                                        "stored by the SSE kernel into its rank's gather area + flags to all ranks; each rank pulls and reads the block of step k-1 over NVLink on a second stream behind the launch of step k (fused, 4 buffers)"
                                        if gather_mode == "fused" else "by pack + NCCL all-gather [%s]" % gather_mode)) if world > 1 else "1 rank"},
            "clocks": clocks, "e2e": {"value": e2e_value, "unit": "traj-control-steps/s", "h2d_bytes_per_step": B * 4,
                                      "d2h_bytes_per_step": B * (K_mom * 8 + L.QC_AUX_COUNT * 8 + 1),
                                      "includes_exchange": world > 1},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline, "closed_loop": closed_loop, "extra": extra,
            "check": {"max_norm_deviation": norm_dev, "wall_s_timed_region": t_wall, "fused_gather_equals_nccl": gather_ok, "gather_failed_ranks": gather_failed,
                      "consumed_checksum": float(consumed.item())}}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
