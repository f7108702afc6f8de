#!/usr/bin/env python3
"""bench.py -- headline benchmark of the SSE hot path: trajectory-control-steps/sec on N B200s.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA path through the C-ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU algorithm (oracle port) on the host cores

A "step" = one control step of every trajectory of the batch: n_sub SSE substeps + moment extraction + flags, ONE kernel
launch per rank.  Workload at N=1 = BASELINE.json configs[1]: quartic oscillator cooling, 1024 trajectories (N=171 grid
points, 80 substeps per control step); N>1 = the same per-GPU batch on every rank (weak scaling), trajectories sharded with no
data-path collective plus the all-gather of the moment/reward block.  Prints ONE JSON line (rank 0).
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


def cpu_port_rate(task, params, n_traj_per_thread, threads, fast=True):
    """The oracle (CPU restatement of the reference algorithm) timed the way the reference runs: one trajectory per
    single-threaded worker, `threads` workers.  Returns (traj-control-steps/s, seconds)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import oracle_for, initial_states, level_force
    psi0 = initial_states(params, threads * n_traj_per_thread, seed=3)
    rng = np.random.default_rng(7)
    noise = rng.standard_normal((threads * n_traj_per_thread, params["n_sub"], 2))
    acts = rng.integers(0, params["n_levels"], threads * n_traj_per_thread)
    oracles = [oracle_for(params, fast=fast) for _ in range(threads)]
    for t, o in enumerate(oracles):      # warm the 21-force cache is NOT done: the reference refactorises on every force change too
        st = psi0[t * n_traj_per_thread].copy()
        o.run(st, params["dt"], 0.0, params["gamma"], noise[0][:2])

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
        from common import initial_states, level_force
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
            "config": {"workload": workload_name(task, args.batch, n, params["n_sub"]), "sample": "%d trajectories x 1 control step per timed step" % (cores * per_thread)},
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
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import initial_states

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

    def exchange():
        """Every rank obtains the [world*B, K+5] result block of this control step."""
        if fused is not None:
            fused.wait()                                # rows were stored by the kernel itself; this only waits for the other ranks' flags
        elif world > 1:
            qdist.all_gather_block(qdist.pack_block(out["moments"], out["aux"], out["flags"]), world, gathered)
    flush_buf = None if args.no_l2_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    # roofline denominators measured in-process (MEASURED_PEAKS.json has no FP64 / shared-memory entry)
    fp64_peak, smem_peak = measure_peaks(local_rank)

    def one_step(i):
        sim.step(actions[i], out=out)
        exchange()

    for i in range(W_steps):
        one_step(i)
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
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    launches = sim.launch_count() - launches0
    ms_total = sum(a.elapsed_time(b) for a, b in evs)
    ms_kernel = sum(a.elapsed_time(b) for a, b in kev) / K_steps
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms_total = float(t.item())
    value = world * B * K_steps / (ms_total * 1e-3)
    norm_dev = float((out["aux"][:, L.QC_AUX_NORM] - 1).abs().max().item())
    gather_ok = None
    if fused is not None:                               # the fused block of the last step must equal a plain NCCL all-gather of the same outputs
        qdist.all_gather_block(qdist.pack_block(out["moments"], out["aux"], out["flags"]), world, gathered)
        gather_ok = bool(torch.equal(fused.block(), gathered))
        fused.close()
        sim.step(actions[0], out=out)                   # (the end-to-end leg below runs without the exchange)

    # ---- end-to-end through the host-buffer C-ABI call (pinned host memory, H2D + kernel + D2H per step) ----
    act_host = actions.cpu().pin_memory()
    mom_h = torch.empty((B, sim.K), dtype=torch.float64).pin_memory()
    aux_h = torch.empty((B, L.QC_AUX_COUNT), dtype=torch.float64).pin_memory()
    flg_h = torch.empty((B,), dtype=torch.uint8).pin_memory()
    for i in range(min(3, W_steps)):
        sim.step_host(act_host[i], moments=mom_h, aux=aux_h, flags=flg_h)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    t0 = time.perf_counter()
    for i in range(K_steps):
        sim.step_host(act_host[W_steps + i], moments=mom_h, aux=aux_h, flags=flg_h)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * B * K_steps / e2e_s

    closed_loop = None
    if world == 1 and not args.no_closed_loop:
        closed_loop = closed_loop_rate(sim, params, B, K_steps, min(W_steps, 10), out)

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
    hbm_bytes = B * (32.0 * n + 8.0 * sim.K + 8.0 * L.QC_AUX_COUNT + 5)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    roofline = {"bound": "fp64", "achieved": achieved, "peak": fp64_peak / 1e12, "unit": "TFLOP/s", "frac": achieved / (fp64_peak / 1e12),
                "traffic": prof.get("dram_bytes_per_launch"),
                "kernel": sim.kernel_info(), "kernel_ms": ms_kernel, "flops_per_unit": fpu,
                "peak_source": "measured in-process: dependency-free DFMA loop on all SMs (qc_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 entry",
                "smem": {"peak_TBps": smem_peak / 1e12, "bytes_per_launch": prof.get("smem_bytes_per_launch"),
                         "achieved_TBps": (prof.get("smem_bytes_per_launch") / (ms_kernel * 1e-3) / 1e12) if prof.get("smem_bytes_per_launch") else None},
                "hbm": {"algorithmic_bytes_per_launch": hbm_bytes, "achieved_GBps": hbm_bytes / (ms_kernel * 1e-3) / 1e9,
                        "peak_GBps": peaks.get("hbm_gbs", 6650.0), "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        _, s0 = cpu_port_rate(task, params, 1, cores)
        per_thread = max(2, int(12.0 / max(s0, 1e-3)))
        v, secs = cpu_port_rate(task, params, per_thread, cores)
        cpu_baseline = {"value": v, "unit": "traj-control-steps/s", "cores": cores, "kind": "port",
                        "sample": "%d threads x %d trajectories x 1 control step (%d substeps, N=%d) in %.1f s; oracle/sse_oracle.c built -Ofast (the reference's flag), "
                                  "per-force refactorisation included as in the reference" % (cores, per_thread, params["n_sub"], n, secs),
                        "reference_build_as_shipped": reference_build_call_pattern(task, params)}

    line = {"metric": "trajectory-control-steps/sec", "value": value, "unit": "traj-control-steps/s", "n_gpus": world, "steps": K_steps,
            "warmup": W_steps, "ms_per_step": ms_total / K_steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(task, B, n, params["n_sub"]), "task": task, "trajectories_per_gpu": B, "global_trajectories": world * B,
                       "state_len": n, "n_sub": params["n_sub"], "noise": "in-kernel Philox4x32-10 + Box-Muller",
                       "actions": "uniform over 21 levels, redrawn every control step (torch.Generator seed 0)",
                       "l2": "inputs (2.8 MB/GPU) fit L2; L2 flushed with a 256 MiB write between timed steps, per-step CUDA events summed" if flush_buf is not None else "no flush",
                       "parallelism": ("%d rank(s), trajectories sharded, [B,%d] f64 result block per step %s" % (world, sim.K + 5,
                                        "stored by the SSE kernel into every rank's peer memory + flag wait (fused)" if gather_mode == "fused" else "by pack + NCCL all-gather [%s]" % gather_mode)) if world > 1 else "1 rank"},
            "clocks": clocks, "e2e": {"value": e2e_value, "unit": "traj-control-steps/s", "h2d_bytes_per_step": B * 4,
                                      "d2h_bytes_per_step": B * (sim.K * 8 + L.QC_AUX_COUNT * 8 + 1)},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline, "closed_loop": closed_loop,
            "check": {"max_norm_deviation": norm_dev, "wall_s_timed_region": t_wall, "fused_gather_equals_nccl": gather_ok}}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
