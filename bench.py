#!/usr/bin/env python
"""Benchmark of the batched interior-point hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload mpc_quadcopter|mpc_dcmotor|l1l2|sls] [--batch B_PER_GPU]

One "step" = one pass of the hot path over one batch of synthetic instances:
``set parameters + primal initial values  ->  ipmPD_CSsolver for every instance  ->  getOutputs``.
The default workload is the configuration the metric is quoted on: ``mpcmhe/mpc_quadcopter``
(T=100, ipmPD_CS, nonlinear), 65536 instances PER GPU (the whole named batch fits one GPU;
"scaling": "weak" -- every rank owns its own 65536 instances, no collective on the data path).

Printed by rank 0, ONE JSON line:
  value      converged solves/s, whole job, inputs/outputs resident in HBM (device pointers
             through the C ABI), CUDA events on the solver's stream, max over ranks
  e2e        the same metric through the host-buffer C ABI calls (tc_set / tc_solve /
             tc_get_outputs on pinned host memory): H2D of every input and D2H of every output
             and of status/iter inside the timed region
  roofline   dominant kernel (tc_solve_kernel): algorithmic group-boundary bytes actually
             executed (device counters x static per-group table) / CUDA-event kernel time
  cpu_baseline  the reference's own C solver (oracle/_ref, built from /root/reference) on all
             host cores, one process per core, bounded sample (N=1, rank 0 only)

``--impl reference`` times only that CPU arm (see run_reference).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "converged IPM solves/sec (batched, FP64)"
UNIT = "solves/s"
DEFAULT_BATCH = dict(mpc_quadcopter=65536, mpc_dcmotor=262144, l1l2=65536, sls=262144)
WORKLOAD_DESC = dict(
    mpc_quadcopter="mpcmhe/mpc_quadcopter nonlinear MPC T=100 (nU=700,nG=406,nF=300, KKT 1406), ipmPD_CS, mu0=1e-3, maxIter=1000",
    mpc_dcmotor="mpcmhe/mpc_dcmotor linear MPC T=30 (KKT 323), mu0=1e-3, maxIter=100",
    l1l2="examples/l1l2estimationCS N=200 (KKT 1792), mu0=1, maxIter=100",
    sls="examples/sls bounded least squares N=100 n=8 (KKT 24), mu0=1, maxIter=30")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mpc_quadcopter", choices=sorted(DEFAULT_BATCH))
    ap.add_argument("--batch", type=int, default=0, help="instances per GPU (default: the BASELINE batch)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="instances per core in the CPU baseline")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--pipeline", type=int, default=3,
                    help="solver handles kept in flight (own stream + scratchbook each): the straggler tail of "
                         "step k overlaps the bulk of steps k+1.. ; 1 = strictly one batch at a time")
    ap.add_argument("--driver", default="", choices=["", "warp", "lanes"],
                    help="device driver of the compiled back end (default: the emitter's default)")
    ap.add_argument("--backend", default="compiled", choices=["compiled", "tape"],
                    help="compiled: per-problem sm_100a module (default); tape: generic interpreter")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own C solver on the host cores (checker library, oracle/_ref)
# ------------------------------------------------------------------------------------------------
def _oracle_library(gen):
    """prebuilt reference library for this exact instruction list (kind 'reference'), else the
    port (kind 'port') -- bench.py's cpu_baseline / --impl reference are the only places
    outside tests/ and smoke() that touch oracle/"""
    from oracle.oracle import Oracle
    for kind, label in (("ref", "reference"), ("port", "port")):
        o = Oracle(gen, kind)
        if os.path.exists(o.so):
            return o.so, label
    o = Oracle(gen, "ref")
    return o.build(), "reference"      # only possible where /root/reference exists


def _cpu_inputs(name, gen, n, seed):
    from tenscalc_b200 import problems
    from oracle.oracle import pack_inputs
    P, V = problems.synthetic_inputs(name, gen, n, seed=seed)
    Pp, Vp = pack_inputs(gen, P, V, n)
    nY = sum(int(np.prod(s)) if s else 1 for _, s in gen.outputs)
    return Pp, Vp, nY


def cpu_baseline(name, gen, sargs, per_core, repeats=1):
    from oracle.cpu_runner import CpuPool, host_cores
    so, kind = _oracle_library(gen)
    cores = host_cores()
    n = cores * per_core
    Pp, Vp, nY = _cpu_inputs(name, gen, n, seed=12345)
    pool = CpuPool(so, cores)
    pool.solve(Pp[:cores], Vp[:cores], nY, sargs["mu0"], sargs["maxIter"])      # spawn + dlopen warm-up
    best = None
    for _ in range(repeats):
        st, it, _, sec, wall = pool.solve(Pp, Vp, nY, sargs["mu0"], sargs["maxIter"])
        if best is None or wall < best[4]:
            best = (st, it, None, sec, wall)
    pool.close()
    st, it, _, sec, wall = best
    conv = int((st == 0).sum())
    return dict(value=conv / wall, unit=UNIT, cores=cores, kind=kind,
                sample=f"{n} instances of the same synthetic workload ({per_core} per core, one process per core, "
                       f"gcc -O1), wall {wall:.2f} s, {conv} converged",
                us_per_newton_step=1e6 * wall / max(int(it.sum()), 1),
                us_per_newton_step_per_core=1e6 * float(sec.sum()) / max(int(it.sum()), 1),
                solver_ms_per_instance_one_core=1e3 * float(sec.mean()))


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path (its unmodified
    ipmPD_CSsolver.c + code printed by its own writeCinstructionsC, compiled with gcc -O1)
    on all host cores; each step = a bounded sample of the workload."""
    if rank != 0:
        return
    from tenscalc_b200 import problems
    from oracle.cpu_runner import CpuPool, host_cores
    name = args.workload
    gen, sargs = problems.BUILDERS[name]()
    so, kind = _oracle_library(gen)
    cores = host_cores()
    per_core = args.cpu_sample or dict(mpc_quadcopter=24, l1l2=96, mpc_dcmotor=2048, sls=8192)[name]
    n = cores * per_core
    Pp, Vp, nY = _cpu_inputs(name, gen, n, seed=12345)
    pool = CpuPool(so, cores)
    pool.solve(Pp[:cores], Vp[:cores], nY, sargs["mu0"], sargs["maxIter"])
    for _ in range(args.warmup):
        pool.solve(Pp, Vp, nY, sargs["mu0"], sargs["maxIter"])
    walls, conv, iters = [], 0, 0
    for _ in range(args.steps):
        st, it, _, _, wall = pool.solve(Pp, Vp, nY, sargs["mu0"], sargs["maxIter"])
        walls.append(wall)
        conv += int((st == 0).sum())
        iters += int(it.sum())
    pool.close()
    total = float(np.sum(walls))
    value = conv / total
    sample = f"{n} instances per step ({per_core} per core x {cores} processes), gcc -O1"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[name], "sample_per_step": n},
        "us_per_newton_step": 1e6 * total / max(iters, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f:
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=float(max(pw)))
        return out


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from tenscalc_b200 import metrics, problems
    from tenscalc_b200.runtime import pack_inputs

    name = args.workload
    B = args.batch or DEFAULT_BATCH[name]
    K, W = args.steps, max(args.warmup, 0)
    gen, sargs = problems.BUILDERS[name]()
    tables = metrics.group_tables(gen.problem)
    io = metrics.io_bytes_per_instance(gen)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:      # before CUDA is touched by this process
        per_core = args.cpu_sample or dict(mpc_quadcopter=48, l1l2=256, mpc_dcmotor=8192, sls=32768)[name]
        cpu = cpu_baseline(name, gen, sargs, per_core)

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic batch of this rank, in pinned host memory (C ABI layout double[B][n]) ----
    P, V = problems.synthetic_inputs(name, gen, B, seed=1000 + rank)
    Ps, Vs = pack_inputs(gen, P, V, B)
    h_in = {}
    for pre, blocks in (("P_", Ps), ("V_", Vs)):
        for nm, blk in blocks.items():
            t = torch.empty(blk.shape, dtype=torch.float64).pin_memory()
            t.numpy()[...] = blk
            h_in[pre + nm] = t
    del P, V, Ps, Vs
    out_sizes = [int(np.prod(s)) if s else 1 for _, s in gen.outputs]
    h_out = [torch.empty((B, n), dtype=torch.float64).pin_memory() for n in out_sizes]
    h_status = torch.empty(B, dtype=torch.int32).pin_memory()
    h_iter = torch.empty(B, dtype=torch.int32).pin_memory()
    d_in = {k: t.cuda(non_blocking=True) for k, t in h_in.items()}
    d_out = [torch.empty((B, n), dtype=torch.float64, device="cuda") for n in out_sizes]
    d_status = torch.zeros(B, dtype=torch.int32, device="cuda")
    d_iter = torch.zeros(B, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()

    # `depth` handles in flight: each owns a stream and a scratchbook, so consecutive steps are independent
    # launches and the few instances that need hundreds of Newton iterations (the tail of one
    # batch) run next to the bulk of the following batches instead of holding the GPU alone
    depth = max(1, min(args.pipeline, K))
    mopts = dict(driver=args.driver) if args.driver and args.backend == "compiled" else {}
    objs = [gen(B, device=local_rank, backend=args.backend, **mopts) for _ in range(depth)]
    obj = objs[0]
    assert obj.has_module() == (args.backend == "compiled")
    exts = [torch.cuda.ExternalStream(o.stream_ptr(), device=torch.device("cuda", local_rank)) for o in objs]
    mu0, maxIter = sargs["mu0"], sargs["maxIter"]
    d_outs = [[torch.empty((B, n), dtype=torch.float64, device="cuda") for n in out_sizes] for _ in objs]
    d_stats = [(torch.zeros(B, dtype=torch.int32, device="cuda"), torch.zeros(B, dtype=torch.int32, device="cuda"))
               for _ in objs]
    h_outs = [[torch.empty((B, n), dtype=torch.float64).pin_memory() for n in out_sizes] for _ in objs]
    h_stats = [(torch.empty(B, dtype=torch.int32).pin_memory(), torch.empty(B, dtype=torch.int32).pin_memory())
               for _ in objs]
    del d_out, d_status, d_iter, h_out, h_status, h_iter

    def step_device(i):
        o = objs[i % depth]
        for k, t in d_in.items():
            o.set_device_ptr(k, t.data_ptr(), t.shape[1])
        o.solve_device_ptr(d_stats[i % depth][0].data_ptr(), d_stats[i % depth][1].data_ptr(), mu0, maxIter)
        o.get_outputs_ptrs([t.data_ptr() for t in d_outs[i % depth]], device=True)

    def e2e_submit(i):
        o = objs[i % depth]
        for k, t in h_in.items():
            o.set_packed(k, t.numpy())
        o.solve_async(h_stats[i % depth][0].numpy(), h_stats[i % depth][1].numpy(), mu0, maxIter)

    def e2e_collect(i):        # blocks until step i is complete and its results are in host memory
        objs[i % depth].get_outputs_ptrs([t.data_ptr() for t in h_outs[i % depth]], device=False)

    def all_sync():
        for o in objs:
            o.synchronize()

    # ---- device-resident arm ----
    for i in range(W):
        step_device(i)
    all_sync()
    for o in objs:
        o.group_counts(reset=True)
    launches0 = sum(o.kernel_launches() for o in objs)
    clocks = ClockSampler(local_rank) if rank == 0 else None
    e0 = torch.cuda.Event(enable_timing=True)
    e1s = [torch.cuda.Event(enable_timing=True) for _ in objs]
    barrier()
    t_wall = time.perf_counter()
    e0.record(exts[0])
    for j in range(1, depth):          # every stream starts after the common start event
        exts[j].wait_event(e0)
    solve_ms = []
    for i in range(K):
        if i >= depth:                 # the handle is being reused: its previous solve time is final
            solve_ms.append(objs[i % depth].last_solve_ms())
        step_device(i)
    for j in range(depth):
        e1s[j].record(exts[j])
    all_sync()
    wall_dev_ms = 1e3 * (time.perf_counter() - t_wall)
    barrier()
    for j in range(min(depth, K)):
        solve_ms.append(objs[j].last_solve_ms())
    total_ms = max(e0.elapsed_time(e) for e in e1s)
    clk = clocks.stop() if clocks else None
    launches = sum(o.kernel_launches() for o in objs) - launches0
    counts = sum(o.group_counts(reset=True) for o in objs)
    last = (K - 1) % depth
    st = d_stats[last][0].cpu().numpy()
    it = d_stats[last][1].cpu().numpy()
    conv = int((st == 0).sum())
    newton = int(it.sum())

    # ---- end-to-end arm (host buffers through the C ABI) ----
    e2e = None
    if not args.no_e2e:
        for i in range(max(1, min(W, 3))):
            e2e_submit(i)
            e2e_collect(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(K):
            if i >= depth:
                e2e_collect(i - depth)
            e2e_submit(i)
        for i in range(max(0, K - depth), K):
            e2e_collect(i)
        all_sync()
        barrier()
        e2e_ms = 1e3 * (time.perf_counter() - t0)     # host clock: the region starts and ends on the host
        conv_e2e = int((h_stats[last][0].numpy() == 0).sum())
    else:
        e2e_ms, conv_e2e = float("nan"), 0

    # ---- reduce over ranks: time = max, work = sum ----
    red = torch.tensor([total_ms, e2e_ms, float(np.mean(solve_ms))], dtype=torch.float64, device="cuda")
    tot = torch.tensor([conv, newton, B, conv_e2e, launches], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    total_ms, e2e_ms, solve_ms_mean = [float(x) for x in red.cpu()]
    conv_all, newton_all, B_all, conv_e2e_all, launches_all = [float(x) for x in tot.cpu()]

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # rank-0 kernel: algorithmic bytes it executed per launch / its mean duration
        alg_bytes = float((counts * tables["boundary_bytes"]).sum()) / K
        if args.backend == "compiled":     # bytes the emitted code loads/stores (after per-chunk register reuse)
            from tenscalc_b200.emit_cuda import emit_module
            per_group = emit_module(gen)[1]["per_group"]
            issued = np.array([8.0 * (g["loads"] + g["stores"]) for g in per_group])
        else:
            issued = tables["vm_bytes"]
        vm_bytes = float((counts * issued).sum()) / K
        flops = float((counts * tables["flops"]).sum()) / K
        kms = float(np.mean(solve_ms))
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            key = f"{name}:{B}"
            if key in tj:
                traffic = tj[key]["dram_bytes_per_launch"]
        except Exception:
            pass
        # with several launches in flight the denominator is the time the GPU spent per launch
        # (timed region / launches), not the (overlapping) start-to-end duration of one launch
        per_launch_ms = kms if depth == 1 else total_ms / K
        achieved = alg_bytes / (per_launch_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": conv_all * K / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[name], "batch_per_gpu": B, "global_batch": int(B_all),
                       "parallelism": f"batch range-partitioned over {world} GPU(s), no collective",
                       "pipeline": f"{depth} solver handle(s) in flight per GPU (own stream and scratchbook each)",
                       "l2": f"per-GPU scratchbook {obj.scratch_bytes() / 1e9:.1f} GB per handle >> 126 MB L2 (inputs larger than L2)"},
            "solves_per_s_all_statuses": B_all * K / (total_ms * 1e-3),
            "converged_fraction": conv_all / B_all,
            "us_per_newton_step": 1e3 * total_ms / K / max(newton_all, 1.0),
            "newton_steps_per_solve": newton_all / B_all,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                         "kernel": "tcm_solve_kernel" if args.backend == "compiled" else "tc_solve_kernel", "kernel_ms": kms, "kernel_ms_per_launch_of_timed_region": per_launch_ms,
                         "kernel_share_of_step": min(1.0, kms * K / total_ms / depth),
                         "algorithmic_bytes_per_launch": alg_bytes, "issued_bytes_per_launch": vm_bytes, "backend": args.backend,
                         "algorithmic_gflops": flops / (per_launch_ms * 1e-3) / 1e9},
            "cpu_baseline": cpu,
            "e2e": None if args.no_e2e else {
                "value": conv_e2e_all * K / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / K,
                "h2d_bytes_per_step": int(io["h2d"] * B), "d2h_bytes_per_step": int(io["d2h"] * B)},
            "gpu_launches": int(launches_all),
            "clocks": clk,
        }
        print(json.dumps(line))
    for o in objs:
        o.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
