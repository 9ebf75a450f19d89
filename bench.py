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
    ap.add_argument("--workload", default="mpc_quadcopter", choices=sorted(DEFAULT_BATCH) + ["dense_ls"])
    ap.add_argument("--batch", type=int, default=0, help="instances per GPU (default: the BASELINE batch)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="instances per core in the CPU baseline")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--pipeline", type=int, default=7,
                    help="solver handles kept in flight (own stream + scratchbook each): the straggler tail of "
                         "step k overlaps the bulk of steps k+1.. ; 1 = strictly one batch at a time")
    ap.add_argument("--driver", default="", choices=["", "warp", "lanes"],
                    help="device driver of the compiled back end (default: the emitter's default)")
    ap.add_argument("--mopts", default="", help="emitter options of the compiled module, e.g. lockstep=4,max_loads=64")
    ap.add_argument("--parity", type=int, default=4096,
                    help="benchmarked instances re-solved by the CPU oracle after the timed region (0: skip)")
    ap.add_argument("--no-strong", action="store_true",
                    help="N>1: skip the strong-scaling leg (the BASELINE batch split over the N GPUs)")
    ap.add_argument("--phases", default="", help="iteration budget of the phases of a phased solve, e.g. 1x40,2x12,4x16,8x16,16x16,32 "
                                                  "(last repeats; default: the library's schedule)")
    ap.add_argument("--ordering", default="", help="elimination ordering of the generator (default: symamd)")
    ap.add_argument("--backend", default="compiled", choices=["compiled", "tape"],
                    help="compiled: per-problem sm_100a module (default); tape: generic interpreter")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own C solver on the host cores (checker library, oracle/_ref)
# ------------------------------------------------------------------------------------------------
def _oracle_library(gen, opt="-O1", build=True):
    """prebuilt reference library for this exact instruction list (kind 'reference'), else the
    port (kind 'port') -- bench.py's cpu_baseline / parity / --impl reference legs are the only
    places outside tests/ and smoke() that touch oracle/"""
    from oracle.oracle import Oracle
    for kind, label in (("ref", "reference"), ("port", "port")):
        o = Oracle(gen, kind, opt)
        if os.path.exists(o.so):
            return o.so, label
    if not build:
        return None, None
    if os.path.isdir("/root/reference"):
        return Oracle(gen, "ref", opt).build(), "reference"
    return Oracle(gen, "port", opt).build(), "port"      # gcc only: possible on the GPU box too


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

    def timed(lib):
        pool = CpuPool(lib, cores)
        pool.solve(Pp[:cores], Vp[:cores], nY, sargs["mu0"], sargs["maxIter"])      # spawn + dlopen warm-up
        best = None
        for _ in range(repeats):
            r = pool.solve(Pp, Vp, nY, sargs["mu0"], sargs["maxIter"])
            if best is None or r[4] < best[4]:
                best = r
        pool.close()
        return best

    st, it, _, sec, wall = timed(so)
    conv = int((st == 0).sum())
    out = dict(value=conv / wall, unit=UNIT, cores=cores, kind=kind,
               sample=f"{n} instances of the same synthetic workload ({per_core} per core, one process per core, "
                      f"gcc -O1), wall {wall:.2f} s, {conv} converged",
               us_per_newton_step=1e6 * wall / max(int(it.sum()), 1),
               us_per_newton_step_per_core=1e6 * float(sec.sum()) / max(int(it.sum()), 1),
               solver_ms_per_instance_one_core=1e3 * float(sec.mean()))
    # the same sample with the toolbox's default optimisation level (-O0, parameters4all.m:480-482) and with
    # -O2, where those libraries were prebuilt (oracle/build_all.py); -O1 above is the example scripts'
    # "fastest run-time" setting (l1l2estimationCS.m:22-25) and stays the headline baseline.  No
    # -march=native: the libraries are built in the build container and must run on the GPU box's
    # host, and -mfma would let gcc contract a*b+c, i.e. change the reference's rounding sequence.
    other = {}
    for opt in ("-O0", "-O2"):
        lib, k2 = _oracle_library(gen, opt, build=False)
        if lib:
            st2, it2, _, _, wall2 = timed(lib)
            other[opt] = dict(value=int((st2 == 0).sum()) / wall2, kind=k2,
                              status_iter_equal_to_O1=bool(np.array_equal(st2, st) and np.array_equal(it2, it)))
    if other:
        out["other_opt_levels"] = other
    return out


def parity_at_scale(name, gen, sargs, Pp, Vp, st_gpu, it_gpu, outs_gpu, n_check):
    """Re-solve `n_check` of the BENCHMARKED instances with the CPU oracle (the reference's own driver and
    emitter, oracle/_ref) and compare: exit status identical, |iter difference| <= 1, outputs of the
    converged instances within 1e-8 relative (BASELINE.json north_star).  The sample holds up to 512
    instances of every non-zero exit status, the 64 largest iteration counts, and a random fill.
    Pp/Vp: packed inputs [B, nP] / [B, nV] of this rank; outs_gpu: list of [B, n_k] host arrays."""
    from oracle.cpu_runner import CpuPool, host_cores
    so, kind = _oracle_library(gen)
    B = len(st_gpu)
    pick = []
    for code in np.unique(st_gpu):
        if code != 0:
            pick += list(np.nonzero(st_gpu == code)[0][:512])
    n_nonzero = len(pick)
    pick += list(np.argsort(-it_gpu.astype(np.int64), kind="stable")[:64])
    pick += list(np.random.default_rng(7).permutation(B))
    sel = np.array(list(dict.fromkeys(int(i) for i in pick))[:max(n_check, n_nonzero + 64)], dtype=np.int64)
    sel = sel[:min(len(sel), B)]
    nY = sum(int(np.prod(s)) if s else 1 for _, s in gen.outputs)
    pool = CpuPool(so, host_cores())
    t0 = time.perf_counter()
    st_o, it_o, Y, _, _ = pool.solve(np.ascontiguousarray(Pp[sel]), np.ascontiguousarray(Vp[sel]), nY,
                                     sargs["mu0"], sargs["maxIter"])
    res = dict(checked=int(len(sel)), oracle=kind, nonzero_status_in_sample=int((st_gpu[sel] != 0).sum()),
               top_iter_in_sample=64, max_iter_in_sample=int(it_gpu[sel].max()))
    st_g, it_g = st_gpu[sel], it_gpu[sel]
    dit = np.abs(it_g.astype(np.int64) - it_o.astype(np.int64))
    res["status_mismatches"] = int((st_g != st_o).sum())
    res["iter_equal"] = int((dit == 0).sum())
    res["iter_within_1"] = int((dit <= 1).sum())
    res["iter_beyond_1"] = int((dit > 1).sum())
    res["max_abs_iter_diff"] = int(dit.max())
    ok = (st_o == 0) & (st_g == 0) & (dit <= 1)
    Yg = np.concatenate([o[sel] for o in outs_gpu], axis=1)
    worst, off = 0.0, 0
    for _, shp in gen.outputs:
        n = int(np.prod(shp)) if shp else 1
        a, b = Yg[ok, off:off + n], Y[ok, off:off + n]
        if a.size:
            scale = np.maximum(np.abs(b).max(axis=1, keepdims=True), 1e-300)
            worst = max(worst, float(np.max(np.abs(a - b) / scale)))
        off += n
    res["max_rel_output_err"] = worst
    res["bit_identical_outputs"] = bool(np.array_equal(Yg[ok], Y[ok]))
    res["tolerance"] = "status ==, |d iter| <= 1, outputs <= 1e-8 relative (converged instances)"
    res["ok"] = bool(res["status_mismatches"] == 0 and res["iter_beyond_1"] == 0 and worst <= 1e-8)
    # libm noise floor: the oracle against ITSELF with the correctly rounded x^-1/2, x^-3/2 in place of glibc's
    # pow (they differ on ~8e-4 of arguments, one ULP): how many instances move when a libm result does
    try:
        from oracle.oracle import Oracle
        probe = Oracle(gen, "ref" if kind == "reference" else "port", variant="crpow")
        uses_pow = any(gen.problem.tab.types[i] in (33, 34) for g in gen.problem.groups for i in g)   # I_Dsqrt, I_DDsqrt
        if uses_pow and os.path.exists(probe.so):
            pool2 = CpuPool(probe.so, host_cores())
            st_p, it_p, _, _, _ = pool2.solve(np.ascontiguousarray(Pp[sel]), np.ascontiguousarray(Vp[sel]), nY,
                                              sargs["mu0"], sargs["maxIter"])
            pool2.close()
            d2 = np.abs(it_p.astype(np.int64) - it_o.astype(np.int64))
            res["libm_noise_floor"] = dict(
                what="oracle vs oracle built with correctly rounded pow(x,-.5), pow(x,-1.5) (1 ULP on ~8e-4 of calls)",
                status_mismatches=int((st_p != st_o).sum()), iter_beyond_1=int((d2 > 1).sum()),
                iter_equal=int((d2 == 0).sum()), max_abs_iter_diff=int(d2.max()))
            # the device against that oracle too: both use the correctly rounded powers
            d3 = np.abs(it_g.astype(np.int64) - it_p.astype(np.int64))
            res["vs_crpow_oracle"] = dict(status_mismatches=int((st_g != st_p).sum()), iter_equal=int((d3 == 0).sum()),
                                          iter_beyond_1=int((d3 > 1).sum()), max_abs_iter_diff=int(d3.max()))
    except Exception as e:      # the probe is optional evidence, never a gate
        res["libm_noise_floor"] = f"unavailable: {e}"
    pool.close()
    res["oracle_seconds"] = time.perf_counter() - t0
    return res


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path (its unmodified
    ipmPD_CSsolver.c + code printed by its own writeCinstructionsC, compiled with gcc -O1)
    on all host cores; each step = a bounded sample of the workload."""
    if rank != 0:
        return
    from tenscalc_b200 import problems
    from oracle.cpu_runner import CpuPool, host_cores
    name = args.workload
    gen, sargs = problems.BUILDERS[name](**(dict(ordering=args.ordering) if args.ordering else {}))
    so, kind = _oracle_library(gen)
    cores = host_cores()
    # about half a second of work per core and step: long enough that the step's wall time (= the slowest core's) is
    # not dominated by one unlucky slice of stragglers, short enough for --steps 20 --warmup 5 to end within a minute
    per_core = args.cpu_sample or dict(mpc_quadcopter=192, l1l2=320, mpc_dcmotor=4096, sls=8192)[name]
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
def _mopts(args):
    """emitter options of the compiled module from --driver / --mopts (integers unless quoted names)"""
    if args.backend != "compiled":
        return {}
    o = dict(driver=args.driver) if args.driver else {}
    for kv in filter(None, args.mopts.split(",")):
        k, v = kv.split("=")
        o[k] = int(v) if v.lstrip("-").isdigit() else v
    return o


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


def _measure(args, name, gen, sargs, B, K, W, rank, local_rank, world, barrier, with_e2e, parity_n, clocks_on):
    """One timed device-resident region (+ optional end-to-end region and parity check) on a batch of B
    instances per GPU.  Returns this rank's raw numbers."""
    import torch
    from tenscalc_b200 import problems
    from tenscalc_b200.runtime import pack_inputs

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
    d_in = {k: t.cuda(non_blocking=True) for k, t in h_in.items()}
    torch.cuda.synchronize()

    # `depth` handles in flight: each owns a stream and a scratchbook, so consecutive steps are independent
    # launches and the few instances that need hundreds of Newton iterations (the tail of one
    # batch) run next to the bulk of the following batches instead of holding the GPU alone
    depth = max(1, min(args.pipeline, K))
    mopts = _mopts(args)
    # as many handles as asked for and as fit: a handle costs its scratchbook (18 GB for mpc_quadcopter at 65 536
    # instances) plus the parking pool of the phased solve; 10 % of the device stays free for the I/O arrays
    free0, total_mem = torch.cuda.mem_get_info()
    objs = [gen(B, device=local_rank, backend=args.backend, **mopts)]
    objs[0].synchronize()
    per_handle = max(free0 - torch.cuda.mem_get_info()[0], objs[0].scratch_bytes()) * 1.08
    depth = max(1, min(depth, 1 + int((torch.cuda.mem_get_info()[0] - 0.10 * total_mem) // per_handle)))
    objs += [gen(B, device=local_rank, backend=args.backend, **mopts) for _ in range(depth - 1)]
    obj = objs[0]
    assert obj.has_module() == (args.backend == "compiled")
    if args.phases:
        ph = []
        for tok in args.phases.split(","):
            a, _, b = tok.partition("x")
            ph += [int(a)] * (int(b) if b else 1)
        for o in objs:
            o.set_phased(True, ph)
    exts = [torch.cuda.ExternalStream(o.stream_ptr(), device=torch.device("cuda", local_rank)) for o in objs]
    mu0, maxIter = sargs["mu0"], sargs["maxIter"]
    d_outs = [[torch.empty((B, n), dtype=torch.float64, device="cuda") for n in out_sizes] for _ in objs]
    d_stats = [(torch.zeros(B, dtype=torch.int32, device="cuda"), torch.zeros(B, dtype=torch.int32, device="cuda"))
               for _ in objs]
    # end-to-end arm: `nbuf` sets of pinned host result buffers per handle, so that the next batch of a handle is
    # queued (host -> device copy, solve, device -> host copy, all stream-ordered) while its previous one is still
    # running -- the host waits for a step only when it needs that step's buffers again.  Two sets when the host
    # has the memory for them (1.3 GB per set for mpc_quadcopter), else one (the host then blocks per step).
    set_bytes = 8 * B * sum(out_sizes) + 8 * B
    try:
        avail = next(int(ln.split()[1]) * 1024 for ln in open("/proc/meminfo") if ln.startswith("MemAvailable"))
    except (OSError, StopIteration):
        avail = 0
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
    max_sets = max(1, int(0.35 * avail // (set_bytes * local_world))) if avail else depth
    nbuf = 2 if max_sets >= 2 * depth else 1
    e2e_depth = min(depth, max_sets)                 # short of host memory: fewer handles take part in this arm
    nsets = e2e_depth * nbuf if with_e2e else 1
    h_outs = [[torch.empty((B, n), dtype=torch.float64).pin_memory() for n in out_sizes] for _ in range(nsets)]
    h_stats = [(torch.empty(B, dtype=torch.int32).pin_memory(), torch.empty(B, dtype=torch.int32).pin_memory())
               for _ in range(nsets)]

    def step_device(i):
        o = objs[i % depth]
        for k, t in d_in.items():
            o.set_device_ptr(k, t.data_ptr(), t.shape[1])
        o.solve_device_ptr(d_stats[i % depth][0].data_ptr(), d_stats[i % depth][1].data_ptr(), mu0, maxIter)
        o.get_outputs_ptrs([t.data_ptr() for t in d_outs[i % depth]], device=True)

    e2e_done = {}

    def e2e_submit(i):         # everything of step i is enqueued on its handle's stream; nothing blocks
        o = objs[i % e2e_depth]
        for k, t in h_in.items():
            o.set_packed(k, t.numpy())
        o.solve_async(h_stats[i % nsets][0].numpy(), h_stats[i % nsets][1].numpy(), mu0, maxIter)
        o.get_outputs_ptrs([t.data_ptr() for t in h_outs[i % nsets]], device=False, sync=False)
        e2e_done[i] = torch.cuda.Event()
        e2e_done[i].record(exts[i % e2e_depth])

    def e2e_collect(i):        # blocks until step i is complete and its results are in host memory
        e2e_done.pop(i).synchronize()

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
    clocks = ClockSampler(local_rank) if clocks_on else None
    e0 = torch.cuda.Event(enable_timing=True)
    e1s = [torch.cuda.Event(enable_timing=True) for _ in objs]
    barrier()
    e0.record(exts[0])
    for j in range(1, depth):          # every stream starts after the common start event
        exts[j].wait_event(e0)
    for i in range(K):
        step_device(i)                 # submission never blocks: K steps are queued on `depth` streams
    for j in range(depth):
        e1s[j].record(exts[j])
    all_sync()
    barrier()
    solve_ms = [objs[j].last_solve_ms() for j in range(min(depth, K))]
    total_ms = max(e0.elapsed_time(e) for e in e1s)
    clk = clocks.stop() if clocks else None
    launches = sum(o.kernel_launches() for o in objs) - launches0
    counts = sum(o.group_counts(reset=True) for o in objs)
    last = (K - 1) % depth
    st = d_stats[last][0].cpu().numpy()
    it = d_stats[last][1].cpu().numpy()

    # ---- parity at benchmark scale (outside every timed region; rank 0's shard) ----
    parity = None
    if parity_n and rank == 0:
        Pp = np.concatenate([h_in["P_" + n].numpy() for n, _ in gen.parameters], axis=1) if gen.parameters \
            else np.zeros((B, 0))
        Vp = np.concatenate([h_in["V_" + n].numpy() for n, _ in gen.variables], axis=1)
        outs = [t.cpu().numpy() for t in d_outs[last]]
        parity = parity_at_scale(name, gen, sargs, Pp, Vp, st, it, outs, parity_n)
        del outs

    # ---- end-to-end arm (host buffers through the C ABI) ----
    e2e_ms, conv_e2e = float("nan"), 0
    if with_e2e:
        for i in range(max(1, min(W, 3))):
            e2e_submit(i)
            e2e_collect(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(K):
            if i >= nsets:         # the host needs the buffers of step i - nsets again: that step must be complete
                e2e_collect(i - nsets)
            e2e_submit(i)
        for i in range(max(0, K - nsets), K):
            e2e_collect(i)
        all_sync()
        barrier()
        e2e_ms = 1e3 * (time.perf_counter() - t0)     # host clock: the region starts and ends on the host
        hl = (K - 1) % nsets
        conv_e2e = int((h_stats[hl][0].numpy() == 0).sum())
        if parity is not None:     # the host-buffer path must give the device-pointer path's results
            parity["e2e_equals_device_arm"] = bool(
                np.array_equal(h_stats[hl][0].numpy(), st) and np.array_equal(h_stats[hl][1].numpy(), it)
                and all(np.array_equal(a.numpy(), b.cpu().numpy()) for a, b in zip(h_outs[hl], d_outs[last])))
    res = dict(B=B, depth=depth, nbuf=nbuf, total_ms=total_ms, e2e_ms=e2e_ms, solve_ms=solve_ms, launches=launches,
               counts=counts, st=st, it=it, conv_e2e=conv_e2e, clk=clk, parity=parity,
               scratch_bytes=obj.scratch_bytes())
    for o in objs:
        o.close()
    del d_in, d_outs, d_stats, h_outs, h_stats, h_in
    torch.cuda.empty_cache()
    return res


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from tenscalc_b200 import metrics, problems

    name = args.workload
    B = args.batch or DEFAULT_BATCH[name]
    K, W = args.steps, max(args.warmup, 0)
    kw = dict(ordering=args.ordering) if args.ordering else {}
    gen, sargs = problems.BUILDERS[name](**kw)
    tables = metrics.group_tables(gen.problem)
    io = metrics.io_bytes_per_instance(gen)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:      # before CUDA is touched by this process
        per_core = args.cpu_sample or dict(mpc_quadcopter=384, l1l2=640, mpc_dcmotor=8192, sls=32768)[name]   # ~1 s per core
        cpu = cpu_baseline(name, gen, sargs, per_core)

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather_ranks(x):       # [world, len(x)] on every rank
        t = torch.tensor(x, dtype=torch.float64, device="cuda")
        if world == 1:
            return t.cpu().numpy()[None, :]
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return torch.stack(out).cpu().numpy()

    def histogram(st):         # exit-status histogram of the whole job
        t = torch.from_numpy(st.astype(np.int32)).cuda()
        if world > 1:
            out = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(out, t)
            t = torch.cat(out)
        codes, cnt = np.unique(t.cpu().numpy(), return_counts=True)
        return {str(int(c)): int(n) for c, n in zip(codes, cnt)}

    r = _measure(args, name, gen, sargs, B, K, W, rank, local_rank, world, barrier,
                 with_e2e=not args.no_e2e, parity_n=args.parity, clocks_on=(rank == 0))
    hist = histogram(r["st"])
    per_rank = gather_ranks([r["total_ms"], r["e2e_ms"], float(np.mean(r["solve_ms"])), float((r["st"] == 0).sum()),
                             float(r["it"].sum()), float(r["it"].max()), float(r["conv_e2e"]), float(r["launches"])])
    total_ms = float(per_rank[:, 0].max())        # time = max over ranks, work = sum over ranks
    e2e_ms = float(per_rank[:, 1].max())
    conv_all, newton_all = float(per_rank[:, 3].sum()), float(per_rank[:, 4].sum())
    conv_e2e_all, launches_all = float(per_rank[:, 6].sum()), float(per_rank[:, 7].sum())
    B_all = B * world
    depth = r["depth"]

    # ---- N>1: strong-scaling leg -- the BASELINE batch ("65536 instances sharded over 8 GPUs") split over ranks
    strong = None
    if world > 1 and not args.no_strong and (args.batch or DEFAULT_BATCH[name]) % world == 0:
        Bs = (args.batch or DEFAULT_BATCH[name]) // world
        rs = _measure(args, name, gen, sargs, Bs, K, W, rank, local_rank, world, barrier,
                      with_e2e=False, parity_n=0, clocks_on=False)
        pr2 = gather_ranks([rs["total_ms"], float((rs["st"] == 0).sum()), float(rs["it"].sum())])
        t2 = float(pr2[:, 0].max())
        strong = {"scaling": "strong", "global_batch": Bs * world, "batch_per_gpu": Bs,
                  "value": float(pr2[:, 1].sum()) * K / (t2 * 1e-3), "unit": UNIT, "ms_per_step": t2 / K,
                  "us_per_newton_step": 1e3 * t2 / K / max(float(pr2[:, 2].sum()), 1.0),
                  "per_rank_total_ms": {"min": float(pr2[:, 0].min()), "median": float(np.median(pr2[:, 0])), "max": t2}}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        counts = r["counts"]
        # rank-0 kernel: algorithmic bytes it executed per launch / its mean duration
        alg_bytes = float((counts * tables["boundary_bytes"]).sum()) / K
        if args.backend == "compiled":     # bytes the emitted code loads/stores (after per-chunk register reuse)
            from tenscalc_b200.build import default_module_opts
            from tenscalc_b200.emit_cuda import emit_module
            per_group = emit_module(gen, **{**default_module_opts(gen, **_mopts(args)), **_mopts(args)})[1]["per_group"]
            issued = np.array([8.0 * (g["loads"] + g["stores"]) for g in per_group])
        else:
            issued = tables["vm_bytes"]
        vm_bytes = float((counts * issued).sum()) / K
        flops = float((counts * tables["flops"]).sum()) / K
        kms = float(np.mean(r["solve_ms"]))
        traffic = traffic_note = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            key = f"{name}:{B}"
            if key in tj:
                traffic = tj[key]["dram_bytes_per_launch"]
                traffic_note = tj[key].get("note")
        except Exception:
            pass
        fp64 = None
        try:
            fp64 = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak.json")))
        except Exception:
            pass
        # with several launches in flight the denominator is the time the GPU spent per launch
        # (timed region / launches), not the (overlapping) start-to-end duration of one launch
        per_launch_ms = kms if depth == 1 else total_ms / K
        achieved = alg_bytes / (per_launch_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                "kernel": "tcm_solve_kernel" if args.backend == "compiled" else "tc_solve_kernel", "kernel_ms": kms,
                "kernel_ms_per_launch_of_timed_region": per_launch_ms,
                "kernel_share_of_step": min(1.0, kms * K / total_ms / depth),
                "algorithmic_bytes_per_launch": alg_bytes, "issued_bytes_per_launch": vm_bytes, "backend": args.backend,
                "algorithmic_flops_per_launch": flops,
                "algorithmic_gflops": flops / (per_launch_ms * 1e-3) / 1e9}
        if traffic_note:
            roof["traffic_note"] = traffic_note
        # both bounds of SURVEY 8(d): time the launch would take at the HBM peak on its algorithmic bytes and
        # at the measured FP64 peak (separately rounded DMUL+DADD, -fmad=false) on its algorithmic flops
        roof["t_hbm_ms"] = alg_bytes / (peak * 1e9) * 1e3
        if fp64 and fp64.get("dmul_dadd_tflops"):
            roof["t_fp64_ms"] = flops / (fp64["dmul_dadd_tflops"] * 1e12) * 1e3
            roof["fp64_peak_tflops"] = {"dfma": fp64.get("dfma_tflops"), "dmul_dadd": fp64.get("dmul_dadd_tflops"),
                                        "source": "profiles/fp64_peak.json (tools/bench/fp64peak.cu on B200)"}
            roof["binding_bound"] = "hbm" if roof["t_hbm_ms"] >= roof["t_fp64_ms"] else "fp64"
        line = {
            "metric": METRIC, "value": conv_all * K / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[name], "batch_per_gpu": B, "global_batch": int(B_all),
                       "parallelism": f"batch range-partitioned over {world} GPU(s), no collective",
                       "pipeline": f"{depth} solver handle(s) in flight per GPU (own stream and scratchbook each)",
                       "ordering": gen.info.get("ordering", "symamd"),
                       "l2": f"per-GPU scratchbook {r['scratch_bytes'] / 1e9:.1f} GB per handle >> 126 MB L2 (inputs larger than L2)"},
            "solves_per_s_all_statuses": B_all * K / (total_ms * 1e-3),
            "converged_fraction": conv_all / B_all,
            "status_histogram": hist,
            "us_per_newton_step": 1e3 * total_ms / K / max(newton_all, 1.0),
            "newton_steps_per_solve": newton_all / B_all,
            "parity": r["parity"],
            "per_rank": {"total_ms": {"min": float(per_rank[:, 0].min()), "median": float(np.median(per_rank[:, 0])),
                                      "max": total_ms},
                         "kernel_ms": {"min": float(per_rank[:, 2].min()), "max": float(per_rank[:, 2].max())},
                         "max_iter": {"min": float(per_rank[:, 5].min()), "max": float(per_rank[:, 5].max())},
                         "newton_steps": {"min": float(per_rank[:, 4].min()), "max": float(per_rank[:, 4].max())}},
            "roofline": roof,
            "cpu_baseline": cpu,
            "e2e": None if args.no_e2e else {
                "value": conv_e2e_all * K / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / K,
                "h2d_bytes_per_step": int(io["h2d"] * B), "d2h_bytes_per_step": int(io["d2h"] * B),
                "api": "tc_set (pinned host arrays) -> tc_solve_async -> tc_get_outputs_async into pinned host arrays; "
                       f"{r['nbuf']} set(s) of host result buffers per handle: the host waits for a step when it needs "
                       "that step's buffers again"},
            "gpu_launches": int(launches_all),
            "clocks": r["clk"],
        }
        if strong:
            line["strong_scaling"] = strong
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    # rank 0 prints ONE JSON line on stdout: keep NCCL's version banner (NCCL_DEBUG=VERSION, the default of some
    # images) out of it; an explicit INFO/TRACE setting of the caller is respected
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    args = parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.workload == "dense_ls":        # BASELINE config 5 runs on the dense path (bench_dense.py)
        import bench_dense
        if args.impl == "reference":
            bench_dense.run_reference(args, rank, METRIC, UNIT)
        else:
            bench_dense.run_ours(args, rank, local_rank, world, METRIC, UNIT, ClockSampler)
    elif args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
