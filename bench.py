#!/usr/bin/env python
"""bench.py -- discretized intervals/sec of the batched SCvx linearize-and-discretize hot path.

Workload (BASELINE.json configs[2], the largest single-GPU configuration): 4096 satellites x K=200
nodes (815,104 intervals), tangential thrust 0.5, tf=2.  One STEP = one SCP linearization pass over the
batch: propagate every satellite (reference trajectory + extract_uk) and discretize every interval
(integrator_steps=101: 100 fixed fourth-order Runge-Kutta(-Nystrom) steps + 101-node trapezoid per interval).
The propagation runs beside the discretization (windows of k gated on its progress, mpc_propagate_discretize;
bit-identical to the two kernels back to back, which --no-overlap times instead).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                          # the reference algorithm on the host cores

Under torchrun (N>1) every rank processes its own 4096-satellite shard (weak scaling) and the discretized
matrices are all-gathered over NCCL/NVLink inside the timed step, as north_star asks.

Prints ONE JSON line (rank 0).  Timing: CUDA events on the launching stream, >= 3 warm-up steps, L2 flushed
between timed steps (256 MiB write), max over ranks, clocks sampled during the timed region.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HUBBLE = np.array([5371.4806e3, -4133.1393e3, 1399.9594e3, 4.6921e3, 4.9848e3, -3.2752e3, 12200.0])
FP64_NOMINAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12      # 148 SM x 64 DFMA/clk x 2 flop x 1.965 GHz = 37.2


def flops_per_interval(n_sub, include_j2=False):
    """ALGORITHMIC FP64 work of one interval: the arithmetic of the algorithm DESIGN.md section 4 states (42 live Phi
    entries, symmetric G, Nystrom's 3-stage fourth-order Runge-Kutta method in step-normalised variables with steps that
    span two quadrature nodes and a cubic-Hermite midpoint, symplectic inverse, 56 accumulators), FMA = 2 flop,
    mul/add = 1, MUFU seeds not counted.  One thread does one interval with no recomputation, so this equals the
    executed DFMA/DMUL/DADD count: per integrator step (= two quadrature nodes) 754 FMA + 242 mul + 99 add (J2: 798 /
    298 / 116), plus the last node and the Phi_end * [integrals] epilogue (1249 flop; J2 1310).  An odd n_sub runs one
    step per node (532 / 182 / 55; J2 571 / 230 / 69).  Counted from the SASS of the shipped kernel
    (scripts/sass_reuse.py) and cross-checked against ncu smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on."""
    tail = 1310 if include_j2 else 1249
    if n_sub % 2 == 0:
        return (n_sub // 2) * ((2 * 798 + 298 + 116) if include_j2 else (2 * 754 + 242 + 99)) + tail
    return n_sub * ((2 * 571 + 230 + 69) if include_j2 else (2 * 532 + 182 + 55)) + tail


def fp64_instr_per_interval(n_sub, include_j2=False):
    """FP64-pipe instructions (DFMA+DMUL+DADD) per interval: the pipe-occupancy view of the same work."""
    tail = 785 if include_j2 else 742
    if n_sub % 2 == 0:
        return (n_sub // 2) * ((798 + 298 + 116) if include_j2 else (754 + 242 + 99)) + tail
    return n_sub * ((571 + 230 + 69) if include_j2 else (532 + 182 + 55)) + tail


def bytes_per_interval():
    return 13 * 8 + 105 * 8       # read x_k, u_k, u_k+1; write 105 doubles


# ------------------------------------------------------------------------------------------ inputs

def make_constellation(n_sats, seed=20240531):
    """Hubble state rotated about z by 2*pi*i/N, speed scaled by 1+0.1*U[0,1), one scale from sat 0."""
    import mpconstellation_b200 as M
    scale = M.SatelliteScale(x=HUBBLE)
    const = scale.get_normalized_constants()
    y = scale.normalize_state(HUBBLE)
    rng = np.random.default_rng(seed)
    ang = 2 * np.pi * np.arange(n_sats) / max(n_sats, 1)
    ca, sa = np.cos(ang), np.sin(ang)
    f = 1 + 0.1 * rng.random(n_sats)
    Y = np.tile(y, (n_sats, 1))
    Y[:, 0], Y[:, 1] = ca * y[0] - sa * y[1], sa * y[0] + ca * y[1]
    Y[:, 3], Y[:, 4] = (ca * y[3] - sa * y[4]) * f, (sa * y[3] + ca * y[4]) * f
    Y[:, 5] = y[5] * f
    return Y, const


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index=0, period=0.004):
        super().__init__(daemon=True)
        self.period, self.samples, self.reasons, self.max_mhz = period, [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------ CPU arms

def cpu_port_baseline(n_sats, K, tf, n_sub, budget_s=12.0):
    """oracle/mpc_oracle.c (plain-C port of the same RK4/trapezoid algorithm, OpenMP over intervals) on a
    bounded sample of the workload: the first S satellites."""
    from oracle import c_oracle as C
    from oracle.mpc_oracle import OracleConstants
    Y, const_m = make_constellation(n_sats)
    const = OracleConstants(const_m.MU, const_m.R_E, const_m.J2, const_m.G0, const_m.ISP, const_m.S, const_m.R0, const_m.RHO)
    cores = C.max_threads()
    n_prop = max(1, int(np.ceil(1000 / (K - 1))))

    def run(S):
        t0 = time.perf_counter()
        x, u, st = C.propagate_batch(Y[:S], tf, const, C.CTRL_TANGENTIAL, (0.5, 0, 0), include_drag=False,
                                     include_J2=False, T=K, n_sub=n_prop)
        out = C.discretize_batch(x, u, tf, const, n_sub=n_sub)
        assert out[5].max() == 0
        return time.perf_counter() - t0
    S = min(n_sats, 2 * cores)
    t = run(S)                                   # calibration (also warms the OpenMP pool)
    S = int(min(n_sats, max(S, S * budget_s / max(t, 1e-3))))
    t = run(S)
    return {"value": S * (K - 1) / t, "unit": "intervals/s", "cores": cores, "kind": "port",
            "sample": f"first {S} of {n_sats} satellites x {K-1} intervals (propagate + discretize, n_sub={n_sub}), "
                      f"{t:.1f} s of oracle/mpc_oracle.c with OpenMP"}


def reference_arm(args):
    """--impl reference: the reference ALGORITHM (scipy RK45 + per-node numpy, one process pool over intervals
    per discretize call, satellites looped serially -- linearize_discretize.py:334-390, optimizer.py:243-249,
    simulator.py:41-45) restated in oracle/mpc_oracle.py, on all host cores, on a bounded sample per step.
    The reference itself is pure Python and cannot travel to the GPU box; the restatement makes the same
    library calls and is pinned to it bit-for-bit by tests/test_oracle_golden.py."""
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import mpc_oracle as O
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sats, K, tf = args.sats, args.nodes, args.tf
    sf = O.scale_factors(HUBBLE)
    const = O.normalized_constants(sf)
    Y, _ = make_constellation(n_sats)
    cores = os.cpu_count() or 1
    S = args.ref_sats
    ctrl = O.ctrl_tangential(0.5)

    def step():
        for s in range(S):
            x, t = O.propagate(Y[s], tf, ctrl, const, False, False, K)
            u = O.extract_uk(x, t, ctrl)
            O.discretize(x, u, tf, const, use_uniform_steps=True, integrator_steps=args.n_sub + 1, processes=cores)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = S * (K - 1) / dt
    # the reference's shipped default (use_uniform_steps=False: quadrature on the 4-8 accepted RK45 steps), once, for the record
    t1 = time.perf_counter()
    for s in range(S):
        x, t = O.propagate(Y[s], tf, ctrl, const, False, False, K)
        O.discretize(x, O.extract_uk(x, t, ctrl), tf, const, use_uniform_steps=False, processes=cores)
    default_mode = S * (K - 1) / (time.perf_counter() - t1)
    sample = (f"{S} of {n_sats} satellites x {K-1} intervals per step (propagate + discretize, use_uniform_steps=True, "
              f"integrator_steps={args.n_sub + 1}), mp.Pool({cores}) per discretize call as the reference does")
    print(json.dumps({
        "impl": "reference", "metric": "discretized intervals/sec", "value": val, "unit": "intervals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": val, "unit": "intervals/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "intervals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "default_mode_intervals_per_s": default_mode,
        "gpu_launches": 0,
    }))


def workload_config(args):
    return {"workload": f"{args.sats} satellites x K={args.nodes} nodes ({args.sats * (args.nodes - 1)} intervals) per GPU: "
                        "propagate (tangential thrust 0.5, no drag/J2) + discretize, BASELINE configs[2]",
            "sats_per_gpu": args.sats, "K": args.nodes, "tf": args.tf, "integrator_steps": args.n_sub + 1,
            "integrator": "fixed-step fourth-order Runge-Kutta-Nystrom (3 stages), one step per two quadrature nodes with a cubic-Hermite midpoint, trapezoid on all integrator_steps nodes", "l2": "flushed between timed steps (256 MiB write)",
            "parallelism": f"satellites sharded over {args.gpus} GPU(s)" + (f", all-gather of the SoA matrices inside the step ({args.gather})" if args.gpus > 1 else "")}


# ------------------------------------------------------------------------------------------ GPU arm

def gpu_arm(args):
    import torch
    import mpconstellation_b200 as M
    from mpconstellation_b200 import _lib
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    _lib.require_gpu()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    try:
        local_phys = int(vis.split(",")[local]) if vis else local      # NVML index of this rank's GPU
    except (ValueError, IndexError):
        local_phys = local
    dist = None
    if world > 1:
        import torch.distributed as dist
        # stdout carries exactly one JSON line: keep NCCL's version banner off it (NCCL_DEBUG=INFO etc. is respected)
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    N, K, tf, n_sub = args.sats, args.nodes, args.tf, args.n_sub
    n_int = N * (K - 1)
    Y, const = make_constellation(N * world)
    Y = Y[rank * N:(rank + 1) * N]
    ctrl = M.ConstantTangentialThrustController(tangential_thrust=0.5)

    # ---- device-resident step --------------------------------------------------------------------
    y0 = torch.from_numpy(Y).to(dev)
    tfd = torch.full((N,), tf, dtype=torch.float64, device=dev)
    x = torch.empty((N, 7, K), dtype=torch.float64, device=dev)
    u = torch.empty((N, 3, K), dtype=torch.float64, device=dev)
    stp = torch.empty(N, dtype=torch.int32, device=dev)
    std = torch.empty(n_int, dtype=torch.int32, device=dev)
    out = torch.empty((105, n_int), dtype=torch.float64, device=dev) if world == 1 else None
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    n_prop = M.batch.default_n_sub(K) if args.prop == "rk4" else 0   # 0: the reference's RK45, replayed (default)
    fused = None
    gather_mode = "none"
    if world > 1:
        from mpconstellation_b200 import distributed as D
        gather_mode = args.gather
        if gather_mode == "fused":
            try:
                fused = D.FusedGather(N * world, K, device=dev, mode=args.fused_mode, chunk_waves=args.chunk_waves)
            except Exception as exc:            # no peer mapping on this box: fall back to the NCCL baseline
                if rank == 0:
                    print(f"[bench] symmetric memory unavailable ({exc}); using the NCCL all-gather", file=sys.stderr)
                gather_mode = "nccl"
        if gather_mode == "nccl":
            local_out = torch.empty((105, n_int), dtype=torch.float64, device=dev)
            comm_stream = torch.cuda.Stream(dev)
            col_chunk = ((N + args.chunks - 1) // args.chunks) * (K - 1)

    # One linearization pass.  Default: the propagation is hidden behind the discretization (the intervals are
    # discretized window by window along k as the propagation publishes its progress, mpc_propagate_discretize);
    # --no-overlap runs the two kernels back to back (what the ncu launch list under profiles/ serialises anyway).
    overlap = not args.no_overlap and (world == 1 or (fused is not None and fused.overlap_ok and
                                                      args.fused_mode in ("unicast", "multicast")))

    def step():
        if overlap and world == 1:
            M.propagate_discretize_device(y0, tfd, ctrl, const, K, n_sub_prop=n_prop, n_sub_disc=n_sub, y=x, u_out=u,
                                          out=out, status_prop=stp, status_disc=std, n_windows=args.windows)
            return
        if overlap:
            # ... and every result is stored into all ranks' gathered buffers (peer stores over NVLink)
            fused.propagate_discretize(y0, tfd, ctrl, const, n_sub_prop=n_prop, n_sub=n_sub, y=x, u_out=u, status_prop=stp,
                                       barrier=True, n_windows=args.windows)
            return
        M.propagate_batch_device(y0, tfd, ctrl, const, include_drag=False, include_J2=False, T=K, n_sub=n_prop,
                                 y=x, u_out=u, status=stp)
        if world == 1:
            M.discretize_batch_device(x, u, tfd, const, n_sub=n_sub, out=out, status=std)
        elif fused is not None:
            # the kernel stores every result into all ranks' gathered buffers (peer stores over NVLink)
            fused.discretize(x, u, tfd, const, n_sub=n_sub, barrier=True)
        else:
            def produce(c0, c1):
                s0, s1 = c0 // (K - 1), c1 // (K - 1)
                M.discretize_batch_device(x[s0:s1], u[s0:s1], tfd[s0:s1], const, n_sub=n_sub, out=local_out,
                                          out_offset=c0, status=std[c0:c1])
            step.gathered = D.nccl_gather_chunks(local_out, args.chunks, side_stream=comm_stream, produce=produce,
                                                 chunk_cols=col_chunk)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    assert int(stp.max()) == 0 and int((fused.status if fused is not None else std).max()) == 0, "device status flags set"
    peak_tflops, _ = M.fp64_peak_tflops(local, repeats=5)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = M.launch_count()
    step_ms, disc_ms, prop_ms = [], [], []
    for _ in range(args.steps):
        flush.fill_(1.0)
        barrier()
        e0, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
        e0.record()
        step()
        e3.record()
        torch.cuda.synchronize(dev)
        step_ms.append(e0.elapsed_time(e3))
    launches = M.launch_count() - launches0
    # The two kernels on their own (same inputs, same L2 flush, same clock sampling window): the discretization
    # kernel's launch duration is what `roofline` is computed from.  Inside the overlapped step the same kernel runs
    # as windows beside the propagation, so it cannot be bracketed by events there.
    scratch = out if world == 1 else torch.empty((105, n_int), dtype=torch.float64, device=dev)
    for _ in range(args.steps):
        flush.fill_(1.0)
        torch.cuda.synchronize(dev)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        M.propagate_batch_device(y0, tfd, ctrl, const, include_drag=False, include_J2=False, T=K, n_sub=n_prop,
                                 y=x, u_out=u, status=stp)
        e1.record()
        M.discretize_batch_device(x, u, tfd, const, n_sub=n_sub, out=scratch, status=std)
        e2.record()
        torch.cuda.synchronize(dev)
        prop_ms.append(e0.elapsed_time(e1))
        disc_ms.append(e1.elapsed_time(e2))
    clocks = sampler.stop()
    if world > 1:
        del scratch
    # the reference's shipped default quadrature mode (adaptive RK45 nodes, replayed on the device): reported
    # next to the headline, which is the fixed-step RK4 / 101-node mode north_star names
    adaptive_ms = None
    if world == 1:
        nn = torch.empty(n_int, dtype=torch.int32, device=dev)
        ts = []
        for _ in range(4):
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
            M.discretize_batch_device(x, u, tfd, const, out=out, status=std, adaptive=dict(rtol=1e-3, atol=1e-6, max_step=1e-2),
                                      n_nodes=nn)
            eb.record()
            torch.cuda.synchronize(dev)
            ts.append(ea.elapsed_time(eb))
        adaptive_ms = float(min(ts[1:]))
        adaptive_nodes = (int(nn.min()), int(nn.max()))
        M.discretize_batch_device(x, u, tfd, const, n_sub=n_sub, out=out, status=std)   # restore the headline result
        torch.cuda.synchronize(dev)
    total_ms = float(np.sum(step_ms))
    if dist is not None:
        t = torch.tensor([total_ms, float(np.sum(disc_ms)), float(np.sum(prop_ms))], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, disc_total, prop_total = float(t[0]), float(t[1]), float(t[2])
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt[0])
    else:
        disc_total, prop_total = float(np.sum(disc_ms)), float(np.sum(prop_ms))
    prop_ms_avg = prop_total / args.steps
    ms_per_step = total_ms / args.steps
    value = world * n_int / (ms_per_step * 1e-3)
    disc_ms_avg = disc_total / args.steps

    # ---- end to end through the public host API (pinned host buffers, H2D + D2H inside the timed region) --
    # every rank keeps its host thread and its pinned buffers on the NUMA node of its own GPU (first touch)
    cpus0 = os.sched_getaffinity(0)
    cpus = M.bind_host_to_gpu(local_phys)
    y0_h = M.pinned_empty((N, 7))
    y0_h[:] = Y
    out_h = M.pinned_empty((105, n_int))
    y_h = M.pinned_empty((N, 7, K))
    u_h = M.pinned_empty((N, 3, K))
    st_h = M.pinned_empty(n_int, np.int32)
    e2e_t = []
    for i in range(2 + max(3, args.steps // 2)):
        barrier()
        t0 = time.perf_counter()
        res, _, _ = M.propagate_discretize(y0_h, tf, ctrl, const, T=K, n_sub_prop=n_prop, n_sub_disc=n_sub, out=out_h,
                                           y_out=y_h, u_out=u_h, status=st_h, device=local)
        dt = time.perf_counter() - t0
        if i >= 2:
            e2e_t.append(dt)
    e2e_s = float(np.mean(e2e_t))
    os.sched_setaffinity(0, cpus0)          # the CPU baseline below uses every host core again
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    # parity spot check inside the bench: device-resident result == host-API result
    if world == 1:
        dev_cols = out[:, :K - 1]
    elif fused is not None:
        dev_cols = fused.buf[:, rank * n_int:rank * n_int + K - 1]
    else:
        dev_cols = local_out[:, :K - 1]
    same = bool(np.array_equal(out_h[:, :K - 1], dev_cols.cpu().numpy()))
    gathered_ok = None
    if fused is not None:
        # every rank must hold every other rank's block: compare a column of each peer block with its owner
        probe = torch.stack([fused.buf[:, r * n_int + 7] for r in range(world)])
        mine = fused.buf[:, rank * n_int + 7].clone()
        allm = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allm, mine)
        gathered_ok = bool(all(torch.equal(probe[r], allm[r]) for r in range(world)))

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    fl = flops_per_interval(n_sub)
    achieved_tflops = fl * n_int / (disc_ms_avg * 1e-3) / 1e12
    # the CPU baseline is a rank-0, N=1 measurement (under torchrun OMP_NUM_THREADS=1 would make it a one-core number)
    cpu = cpu_port_baseline(N, K, tf, n_sub) if (world == 1 and not args.no_cpu_baseline) else None
    line = {
        "metric": "discretized intervals/sec", "value": value, "unit": "intervals/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args),
        "e2e": {"value": world * n_int / e2e_s, "unit": "intervals/s", "ms_per_step": e2e_s * 1e3,
                "h2d_bytes_per_step": int(y0_h.nbytes + N * 8),
                "d2h_bytes_per_step": int(out_h.nbytes + y_h.nbytes + u_h.nbytes + n_int * 4),
                "api": "mpconstellation_b200.propagate_discretize (C-ABI mpc_propagate_discretize_host), pinned host buffers",
                "host_cpus_rank0": len(cpus),
                "matches_device_path": same},
        "gather": {"mode": gather_mode + (":" + args.fused_mode if fused is not None else ""), "verified": gathered_ok,
                   "bytes_received_per_rank_per_step": int((world - 1) * n_int * 8 * (98 if fused is not None and fused.skip_const else 105)),
                   "note": "rows 42..48 of the SoA result (last row of A_k, constants) are written once by the owner and never sent"
                           if fused is not None and fused.skip_const else None} if world > 1 else None,
        "gpu_launches": int(launches),
        "kernel": {"discretize_ms": disc_ms_avg, "propagate_ms": prop_ms_avg,
                   "timing": "each kernel launched on its own right after the timed steps (CUDA events, L2 flushed); "
                             "back_to_back_ms is their sum, ms_per_step the step as shipped",
                   "back_to_back_ms": disc_ms_avg + prop_ms_avg,
                   "overlap": ("propagation hidden behind the discretization: windows along k gated by stream memory operations "
                               "(mpc_propagate_discretize), bit-identical results") if overlap else "none (kernels back to back)",
                   # propagation is sequential in tau (latency-bound), reported as satellite-steps/s
                   "propagator": "scipy RK45 replayed (max_step 0.001: 1000 steps x 6 stages per satellite), dense-output samples"
                                 if n_prop == 0 else f"fixed-step RK4, {(K - 1) * n_prop} steps x 4 stages per satellite",
                   "propagate_sat_steps_per_s": N * (1000 if n_prop == 0 else (K - 1) * n_prop) / (prop_ms_avg * 1e-3),
                   "discretize_intervals_per_s_per_gpu": n_int / (disc_ms_avg * 1e-3),
                   "default_mode_adaptive_rk45": None if adaptive_ms is None else {
                       "discretize_ms": adaptive_ms, "intervals_per_s": n_int / (adaptive_ms * 1e-3),
                       "nodes_per_interval": list(adaptive_nodes),
                       "note": "use_uniform_steps=False (reference default): quadrature on scipy-RK45 accepted steps"}},
        "roofline": {"bound": "fp64", "achieved": achieved_tflops, "peak": peak_tflops, "unit": "TFLOP/s",
                     "frac": achieved_tflops / peak_tflops,
                     # the conservative reading: the same flops over the WHOLE step (propagation, window gating and
                     # -- at N > 1 -- the exchange included)
                     "in_step": {"achieved": fl * n_int / (ms_per_step * 1e-3) / 1e12,
                                 "frac": fl * n_int / (ms_per_step * 1e-3) / 1e12 / peak_tflops},
                     # dram__bytes_read.sum + dram__bytes_write.sum of one full-batch launch, ncu --set full capture of
                     # this command (profiles/r01_n_discretize_pair_windows.txt, third kernel: 65.80 + 628.39 MB);
                     # algorithmic = 944 B x 815,104 = 769.5 MB (the 13 input doubles are shared by neighbours in L2)
                     "traffic": 694.19e6 if (N, K, n_sub) == (4096, 200, 100) else None,
                     "peak_source": "DFMA-chain microbenchmark (mpc_fp64_peak_probe) in this run; MEASURED_PEAKS.json has no FP64 entry",
                     "peak_nominal": FP64_NOMINAL_TFLOPS, "flop_per_interval": fl,
                     # SURVEY.md 8(d) counts the reference formulation (plain RK4 on 42+7 unknowns, dense Phi^-1
                     # products): 2084 n_sub + 1386 flop per interval.  The kernel executes fewer (Nystrom form,
                     # symplectic inverse, Euler identity), so `achieved`/`frac` above use the EXECUTED count (the
                     # conservative reading); this is the same throughput priced at the survey's count.
                     "survey_alg": {"flop_per_interval": 2084 * n_sub + 1386,
                                    "achieved": (2084 * n_sub + 1386) * n_int / (disc_ms_avg * 1e-3) / 1e12,
                                    "frac": (2084 * n_sub + 1386) * n_int / (disc_ms_avg * 1e-3) / 1e12 / peak_tflops},
                     "fp64_pipe_frac": fp64_instr_per_interval(n_sub) * n_int / (disc_ms_avg * 1e-3) / (peak_tflops * 1e12 / 2),
                     "fp64_pipe_note": "FP64 instructions issued / (measured DFMA issue rate): DMUL/DADD occupy a DFMA slot but count 1 flop",
                     "kernel": "mpc::discretize_pair_kernel",
                     "hbm": {"achieved": bytes_per_interval() * n_int / (disc_ms_avg * 1e-3) / 1e9, "peak": hbm_peak,
                             "unit": "GB/s", "peak_source": hbm_src, "bytes_per_interval": bytes_per_interval()}},
        "clocks": clocks,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sats", type=int, default=4096, help="satellites per GPU")
    ap.add_argument("--nodes", type=int, default=200, help="K temporal nodes")
    ap.add_argument("--tf", type=float, default=2.0)
    ap.add_argument("--n-sub", dest="n_sub", type=int, default=100, help="RK4 steps per interval (integrator_steps-1)")
    ap.add_argument("--chunks", type=int, default=8, help="compute/all-gather overlap chunks (N>1, --gather nccl)")
    ap.add_argument("--fused-mode", dest="fused_mode", default="unicast", choices=["unicast", "multicast", "push", "pushk"])
    ap.add_argument("--chunk-waves", dest="chunk_waves", type=int, default=1, help="--fused-mode push: kernel waves per pushed chunk")
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"],
                    help="N>1: all-gather by peer stores from inside the kernel (fused) or chunked NCCL all-gather")
    ap.add_argument("--ref-sats", dest="ref_sats", type=int, default=2, help="satellites per step of the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-overlap", dest="no_overlap", action="store_true",
                    help="run propagate and discretize back to back instead of overlapped")
    ap.add_argument("--prop", default="rk45", choices=["rk45", "rk4"],
                    help="propagator: the reference's RK45 replayed step for step (default) or fixed-step RK4")
    ap.add_argument("--windows", type=int, default=0, help="windows along k of the overlapped pass (0 = library default)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
