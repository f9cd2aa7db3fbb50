#!/usr/bin/env python
"""bench.py -- stiff-ODE ensemble systems solved/sec (Robertson, RADAU_IIA_53) on N B200.

Workload (BASELINE.json configs[1]): 2^20 swept-rate Robertson members PER GPU, RADAU_IIA_53,
rtol = atol = 1e-6, newton tol 1e-7, dx_delta 1e-4, maxit 5000, refresh_jac 10, dt0 = 1e-6,
y0 = {1,0,0}, t in [0,40] (--span long: [0,1e11]).  One "step" = one solve of the whole ensemble.

  value  = members/s with inputs resident in HBM (device-resident C-ABI entry, CUDA-event timed)
  e2e    = members/s through rehuel_irk_ensemble() with HOST buffers (pinned), H2D + D2H inside
  roofline = FP64: flops from the per-member attempt / newton_iters counters the kernel returns
             (SURVEY.md 8d dense model, plus the executed-algorithm model) / kernel time,
             against the FP64 FMA peak measured on this device in this run
  cpu_baseline = the unmodified reference (oracle/_ref) on the host cores, one process per core

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--span short|long]
N > 1: launched by torchrun, one rank per GPU; members are sharded by index (rank r owns members
[r*M, (r+1)*M) of the sweep: weak scaling), the only collective is an all-reduce of statistics.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "stiff-ODE ensemble systems solved/sec (Robertson, RADAU_IIA_53)"
UNIT = "systems/s"
M_PER_GPU = 1 << 20
RTOL = ATOL = 1e-6
DT0 = 1e-6
SPANS = {"short": (0.0, 40.0), "long": (0.0, 1e11)}
FP64_NOMINAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12  # 37.2: SMs x FP64 lanes x 2 x max SM clock


# how the CPU arm was built (oracle/Makefile): the reference's own Makefile:3 adds -march=native -mtune=native, which is
# left out because the library is built in one container and run on another host
CPU_BUILD_NOTE = ("unmodified reference sources over oracle/arma_shim, g++ -std=c++11 -O2 -ffp-contract=off, no -DNDEBUG "
                  "(the LU sits in an assert), no -march=native (built on a different host)")


def workload_name(span: str, M: int) -> str:
    t0, t1 = SPANS[span]
    return (f"robertson-sweep M={M}/GPU (k1,k2,k3 = nominal*10^(u-1/2), splitmix64 seed 0x5EED), RADAU_IIA_53, "
            f"rtol=atol=1e-6, t=[{t0:g},{t1:g}], dt0=1e-6, y0={{1,0,0}}")


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation on the host cores, one process per core
# ---------------------------------------------------------------------------------------------
_WORKER = r"""
import sys, json, time
sys.path.insert(0, {root!r})
import numpy as np
import importlib.util, os
from oracle.refsolver import RefSolver, make_options
# the parameter generator is loaded by PATH: this process must not import (and thereby map) the product library
_spec = importlib.util.spec_from_file_location("_ensembles", os.path.join({root!r}, "rehuel_b200", "ensembles.py"))
ensembles = importlib.util.module_from_spec(_spec); _spec.loader.exec_module(ensembles)
kind, start, stride, total, t1 = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5])
s = RefSolver(kind)
P = ensembles.robertson_params(total)
o = make_options(rel_tol=1e-6, abs_tol=1e-6)
tic = time.perf_counter()
r = s.solve("robertson", 211, 0.0, t1, 1e-6, ensembles.ROBERTSON_Y0, P, o, start=start, stride=stride)
el = time.perf_counter() - tic
n = len(range(start, total, stride))
ok = int((r.status[start::stride] == 0).sum())
print(json.dumps(dict(n=n, ok=ok, seconds=el, attempts=int(r.counters["attempt"][start::stride].sum()))))
"""


def cpu_reference_run(span: str, sample: int, cores: int) -> dict:
    """Members 0..sample-1 of the sweep striped over `cores` processes; wall = slowest process."""
    kind = "reference" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "librehuel_ref.so")) else "port"
    if kind == "port" and not os.path.exists(os.path.join(ROOT, "oracle", "libirk_port.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"], stdout=subprocess.DEVNULL)
    t1 = SPANS[span][1]
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as fh:
        fh.write(_WORKER.format(root=ROOT))
        script = fh.name
    env = dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    tic = time.perf_counter()
    procs = [subprocess.Popen([sys.executable, script, kind, str(c), str(cores), str(sample), repr(t1)],
                              stdout=subprocess.PIPE, env=env, text=True) for c in range(cores)]
    outs = [json.loads(p.communicate()[0].strip().splitlines()[-1]) for p in procs]
    wall = time.perf_counter() - tic
    os.unlink(script)
    slowest = max(o["seconds"] for o in outs)
    n = sum(o["n"] for o in outs)
    assert n == sample and all(o["ok"] == o["n"] for o in outs)
    return dict(value=n / slowest, unit=UNIT, cores=cores, kind=kind,
                sample=f"members 0..{sample - 1} of the same sweep, striped over {cores} processes "
                       f"(one per core); time = slowest process ({slowest:.2f} s; wall incl. start-up {wall:.2f} s); "
                       + CPU_BUILD_NOTE,
                per_core=n / slowest / cores, seconds=slowest, attempts=sum(o["attempts"] for o in outs))


def host_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


_JSON_OUT = None


def claim_stdout() -> None:
    """stdout carries the ONE JSON line and nothing else: file descriptor 1 is pointed at stderr for the rest of
    the process (NCCL writes its version banner to fd 1 from C, below Python's sys.stdout) and the line goes to a
    duplicate of the original descriptor."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    per_step = (2048 if args.span == "short" else 32) * cores
    for _ in range(args.warmup):
        cpu_reference_run(args.span, max(cores, per_step // 8), cores)
    vals, secs = [], 0.0
    for _ in range(args.steps):
        r = cpu_reference_run(args.span, per_step, cores)
        vals.append(r["value"])
        secs += r["seconds"]
    value = per_step * args.steps / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.span, M_PER_GPU), "sample_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": r["kind"],
                         "sample": f"each step = members 0..{per_step - 1} of the sweep, one process per core, time = slowest process; " + CPU_BUILD_NOTE},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(suffix=".csv")
        self.fh = open(self.path, "w")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.fh,
                                         stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.proc.wait()
        self.fh.close()
        sm, mx, power, reasons = [], [], [], set()
        for ln in open(self.path):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [c for c, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "samples_under_load": len(busy),
                "power_w_max": max(power)}


def bind_to_gpu_numa_node(local_rank: int) -> str:
    """Pin this rank (and the pinned buffers it allocates afterwards: first touch) to the NUMA node
    its GPU hangs off; with several ranks per box the host<->device copies of the e2e path otherwise
    cross sockets.  Best effort: returns a note for the JSON line."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(local_rank), "pci_device_id", 0)
        bdf = f"{dom:04x}:{bus:02x}:{dev:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return "numa: single node"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return f"numa: gpu {bdf} -> node {node}, {len(allowed)} cpus"
        return f"numa: node {node} has no allowed cpus"
    except Exception as e:  # noqa: BLE001
        return f"numa: not bound ({type(e).__name__})"


# dram__bytes_read.sum + dram__bytes_write.sum of ensemble_irk_thread<RobertsonF,...> at 2^20 members, t in [0,40]: a
# CONSTANT from the ncu --set full capture of this command, not measured in the run (profiles/README.md names the file)
NCU_DRAM_BYTES_PER_LAUNCH = 526.37e6
NCU_DRAM_SOURCE = "profiles/r2_thread_v11_short_ncu_summary.txt (240.98 MB read + 285.39 MB written)"


def device_ms(fn, torch, reps=1, warm=1):
    """Device time of fn() (CUDA events on the current stream), best of `reps` after `warm` untimed calls."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def counter_sums(ens, torch):
    """Sums of the per-member counters of a DeviceEnsemble after a solve (python ints)."""
    c = ens.counters.to(torch.int64).sum(dim=1).tolist()
    from rehuel_b200 import capi
    d = dict(zip(capi.COUNTER_ROWS, c))
    d["members"] = ens.M
    d["n_failed"] = int((ens.status != 0).sum())
    d["max_attempt"] = int(ens.counter("attempt").max())
    return d


def model_flops(model, cs, booked=False):
    """Executed flops from counter sums.  With failed Newton solves in the run only the iterations of the successful ones
    are counted (flopmodel.AttemptFlops.total: a failed solve books more evaluations than a lane with a non-finite iterate
    executes); booked=True prices every booked iteration instead (an upper bound, reported next to it)."""
    it = None if (booked or cs["reject_newton"] == 0) else cs["newton_iters"]
    return model.total(cs["members"], cs["attempt"], cs["newton_success"], cs["reject_err"], cs["fun_evals"], newton_iters=it)


def run_gpu_arm(args) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist

    from rehuel_b200 import capi, ensembles, flopmodel, sharding
    from rehuel_b200.device import DeviceEnsemble

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    numa_note = bind_to_gpu_numa_node(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}"

    M = args.members
    t0, t1 = SPANS[args.span]
    handout = {"index": capi.ORDER_INDEX, "locality": capi.ORDER_LOCALITY, "locality_desc": capi.ORDER_LOCALITY_DESC}[args.handout]
    opts = capi.make_options(RTOL, ATOL, handout_order=handout)
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    # rank r owns members [r*M, (r+1)*M) of the sweep (weak scaling; any rank can generate its slice)
    P_host = torch.from_numpy(ensembles.robertson_params(M, start=rank * M)).pin_memory()
    y0_host = torch.from_numpy(ensembles.ROBERTSON_Y0.copy()).pin_memory()
    P_dev = P_host.to(dev)
    y0_dev = y0_host.to(dev)
    ens = DeviceEnsemble(capi.ROBERTSON, capi.RADAU_IIA_53, M, y0_dev, P_dev, handout_order=handout)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    peak_tflops = capi.fp64_peak_tflops()
    rob = flopmodel.thread_kernel(3, capi.RADAU_IIA_53, 11, 8, zero_col0_rows=(2,))

    # ---------------- value: inputs resident in HBM, device-timed
    for _ in range(args.warmup):
        ens.solve(t0, t1, DT0, opts)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = capi.kernel_launches()
    ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(args.steps)]
    barrier()
    wall0 = time.perf_counter()
    for a, m, b in ev:
        flush.fill_(1)                 # L2 flush between timed iterations (outside the event pair)
        a.record()
        ens.reorder()                  # hand-out order from the parameters (2 kernels + radix sort); part of the step
        m.record()
        ens.launch(t0, t1, DT0, opts)  # ONE launch of the integration kernel
        b.record()
    barrier()
    wall = time.perf_counter() - wall0
    launches = capi.kernel_launches() - launches0
    step_ms = [a.elapsed_time(b) for a, m, b in ev]
    kernel_ms = [m.elapsed_time(b) for a, m, b in ev]   # the integration kernel alone (roofline)
    total_ms = float(sum(step_ms))
    tt = torch.tensor([total_ms, float(sum(kernel_ms))], dtype=torch.float64, device=dev)
    ll = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ll, op=dist.ReduceOp.SUM)
    total_ms_max, kernel_ms_max = tt[0].item(), tt[1].item()

    # ---------------- statistics from the kernel's own counters (the one collective of the path)
    cs = counter_sums(ens, torch)
    uses_820 = cs["members"] + cs["reject_err"]
    fixed, extra820, per_newton = ensembles.flops_per_attempt_dense_model(3, 3, 11, 8)
    flops_dense = fixed * cs["attempt"] + per_newton * cs["newton_iters"] + extra820 * uses_820
    flops_exec = model_flops(rob, cs)
    flops_r1_local = 259.0 * cs["attempt"] + 369.0 * cs["newton_iters"] + 30.0 * uses_820 - 132.0 * 0.962 * cs["newton_success"]
    local = sharding.pack_stats(members=M, attempt=cs["attempt"], steps=cs["steps"], newton_iters=cs["newton_iters"],
                                fun_evals=cs["fun_evals"], jac_evals=cs["jac_evals"], reject_err=cs["reject_err"],
                                reject_newton=cs["reject_newton"], n_failed=cs["n_failed"], flops_dense_model=flops_dense,
                                flops_executed_model=flops_exec, max_attempt=cs["max_attempt"])
    tot = sharding.allreduce_stats(local, device=dev)
    mass_err = float((ens.y.sum(0) - 1.0).abs().max())

    # ---------------- e2e: the public C-ABI call with host buffers; the caller asks for the states and the status
    # (options.result_mask = 0): the ensemble sums of the counters come back with every call
    e2e_opts = capi.make_options(RTOL, ATOL, handout_order=handout, result_mask=0)

    def e2e_call():
        res = capi.Result()
        capi.check(capi.lib.rehuel_irk_ensemble(
            capi.ROBERTSON, capi.RADAU_IIA_53, t0, t1, DT0, M, 3,
            capi.C.cast(y0_host.data_ptr(), capi._PD), 1, capi.C.cast(P_host.data_ptr(), capi._PD),
            capi.C.byref(e2e_opts), 0, 1, capi.C.byref(res)))
        checksum = (res.sum.attempt, res.sum.n_failed, res.y[M - 1])
        # y rows and the statistics words of the (at most 8) chunks; the status row is only fetched for chunks with failed
        # members (it is zero otherwise, and the statistics say so: capi.cu finish_chunk) -- counted in full if any failed
        d2h = 3 * 8 * M + 13 * 8 * 8 + (4 * M if res.sum.n_failed else 0)
        capi.lib.rehuel_result_free(capi.C.byref(res))
        return d2h, checksum

    for _ in range(max(1, args.warmup)):
        e2e_call()
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        d2h_bytes, chk = e2e_call()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - e0
    e2e_s_max = max_over_ranks(e2e_s)
    e2e_value = world * M * args.steps / e2e_s_max
    assert chk[0] == cs["attempt"] and chk[1] == 0, "the C-ABI call and the device-resident solve disagree"

    # ---------------- the long span (t in [0,1e11], the reference's own Robertson set-up: test_adaptive_step/main.cpp:119-134)
    # as a second timed block: ~0.1 s per launch, so the clock record has samples under load
    long_line = None
    if args.span == "short" and not args.quick:
        lt0, lt1 = SPANS["long"]
        ens.solve(lt0, lt1, DT0, opts)
        barrier()
        lev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(2)) for _ in range(args.long_steps)]
        for a, b in lev:
            ens.reorder()
            a.record()
            ens.launch(lt0, lt1, DT0, opts)
            b.record()
        barrier()
        lms = max_over_ranks(sum(a.elapsed_time(b) for a, b in lev) / len(lev))
        lcs = counter_sums(ens, torch)
        lflops = model_flops(rob, lcs)
        long_line = {"bound": "fp64", "achieved": lflops / (lms * 1e-3) / 1e12, "peak": peak_tflops, "unit": "TFLOP/s",
                     "frac": lflops / (lms * 1e-3) / 1e12 / peak_tflops, "kernel_ms_per_launch": lms, "launches": len(lev),
                     "systems_per_s_per_gpu": M / (lms * 1e-3), "attempts_per_member": lcs["attempt"] / M,
                     "max_attempts": lcs["max_attempt"], "rejected_err_share": lcs["reject_err"] / lcs["attempt"],
                     "workload": workload_name("long", M), "model": "executed flops, as in roofline (rank 0's members)"}
    clocks = sampler.stop() if sampler else None  # sampled over the device-timed, the e2e and the long-span region

    # ---------------- the other BASELINE configs, one device-timed launch each (N = 1 only; outside the headline timing)
    configs = None
    if world == 1 and not args.quick:
        configs = other_configs(capi, ensembles, flopmodel, DeviceEnsemble, torch, dev, peak_tflops)

    # ---------------- strong scaling: fixed totals, whatever N is
    strong = None if args.quick else strong_block(capi, ensembles, DeviceEnsemble, torch, dist, dev, rank, world, barrier, max_over_ranks,
                                                  flopmodel, peak_tflops)

    # ---------------- CPU baseline beside it (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        cpu = cpu_reference_run(args.span, (4096 if args.span == "short" else 64) * cores, cores)

    if rank == 0:
        ms_per_step = total_ms_max / args.steps
        value = world * M * args.steps / (total_ms_max * 1e-3)
        kern_s = kernel_ms_max / args.steps * 1e-3
        ach_dense = tot["flops_dense_model"] / world / kern_s / 1e12   # per GPU, per launch
        ach_exec = tot["flops_executed_model"] / world / kern_s / 1e12
        flops_r1 = flops_r1_local * world  # rank 0's members stand for all (the sweeps are statistically alike)
        info = capi.kernel_info(capi.ROBERTSON, capi.RADAU_IIA_53)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.span, M), "members_total": world * M, "span": args.span,
                       "parallelism": f"member-sharded x{world} (no data-path collective)",
                       "l2": "flushed between timed steps (256 MiB fill, outside the per-step event pair)",
                       "timing": "sum of per-step CUDA-event durations on the launching stream, max over ranks",
                       "handout_order": args.handout + (" (Morton order of the rate constants, recomputed on the device inside every timed step "
                                                       "by two of our kernels and one cub::DeviceRadixSort -- library code on the path, ~0.1 ms"
                                                       + ("; walked from the large-k1 end, whose members need up to 39 attempts against 23, so the "
                                                          "queue drains on cheap members" if args.handout == "locality_desc" else "") + ")" if handout else ""),
                       "wall_s_timed_region_rank0": wall, "host_binding_rank0": numa_note},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(P_host.numel() * 8 + 24),
                    "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": 1e3 * e2e_s_max / args.steps,
                    "bytes_per_member": {"h2d": 24, "d2h": 24 if tot["n_failed"] == 0 else 28},
                    "api": "rehuel_irk_ensemble (C ABI, pinned host buffers; options.result_mask = 0: the states y[3][M] and "
                           "status[M] of every member plus the ensemble sums of all counters come back; the status row is "
                           "filled on the host for chunks whose statistics report no failed member and copied only otherwise; "
                           "t_final, err_last and the nine per-member counter arrays are not asked for)"},
            "gpu_launches": int(ll.item()),
            "clocks": clocks,
            "roofline": {
                "bound": "fp64", "achieved": ach_exec, "peak": peak_tflops, "unit": "TFLOP/s",
                "frac": ach_exec / peak_tflops, "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
                "traffic_note": "CONSTANT, not measured in this run: DRAM bytes read + written per launch of this kernel from the "
                                "ncu --set full capture of this command, " + NCU_DRAM_SOURCE + "; algorithmic I/O is 109 MB",
                "kernel_ms_per_launch": kernel_ms_max / args.steps,
                "model": "flops of the eigen-decoupled algorithm actually executed, from the kernel's own counter sums "
                         f"(rehuel_b200/flopmodel.py): {rob.fixed:g}*attempts + {rob.per_iteration:g}*newton_iterations + "
                         f"{rob.extra_820:g}*uses_of_8.20 - {rob.residual:g}*newton_solves (every solve is taken to skip its dead "
                         "final residual: a lower bound); FMA = 2, other FP64 ops = 1",
                "peak_source": "measured in this run: register-resident DFMA chains (rehuel_fp64_peak); MEASURED_PEAKS.json has no FP64 entry",
                "peak_nominal": FP64_NOMINAL_TFLOPS, "frac_of_nominal": ach_exec / FP64_NOMINAL_TFLOPS,
                "kernel": "ensemble_irk_thread<RobertsonF,3,1,1,EST_REUSE>", "regs": info["regs"],
                "spill_bytes": info["local_bytes"], "ctas_per_sm": info["ctas_per_sm"], "block": info["block"],
            },
            "roofline_round1_model": {
                "achieved": flops_r1 / world / kern_s / 1e12, "frac": flops_r1 / world / kern_s / 1e12 / peak_tflops, "unit": "TFLOP/s",
                "model": "the executed-flop count of the ROUND-1 kernel on the same counters (259*attempts + 369*newton_iterations + "
                         "30*uses_of_8.20 - 132*0.962*newton_solves): what this work cost before the multiplications by structural "
                         "zeros and ones, the dy product, the first iteration's full transform and two of the three square roots "
                         "were removed -- a work-equivalent rate for comparison with BENCH_r01, not a hardware fraction"},
            "roofline_dense_model": {
                "achieved": ach_dense, "frac": ach_dense / peak_tflops, "unit": "TFLOP/s",
                "model": "SURVEY.md 8(d) flops of the REFERENCE's dense (s n) x (s n) algorithm, which the kernel does not execute: "
                         "873*attempts + 321*newton_iters + 41*uses_of_8.20 -- a work-equivalent rate, not a hardware fraction"},
            "roofline_long": long_line,
            "configs": configs,
            "strong": strong,
            "cpu_baseline": ({k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")} if cpu else None),
            "stats": {"attempts_per_member": tot["attempt"] / tot["members"], "accepted_per_member": tot["steps"] / tot["members"],
                      "rejected_err": tot["reject_err"], "rejected_newton": tot["reject_newton"],
                      "newton_iters_per_attempt": tot["newton_iters"] / max(tot["attempt"], 1),
                      "fun_evals": tot["fun_evals"], "jac_evals": tot["jac_evals"], "failed_members": tot["n_failed"],
                      "max_attempts": tot["max_attempt"], "mass_conservation_max_abs_err_rank0": mass_err,
                      "step_ms": step_ms, "kernel_ms": kernel_ms},
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def other_configs(capi, ensembles, flopmodel, DeviceEnsemble, torch, dev, peak_tflops) -> dict:
    """BASELINE configs 3, 4, 5 and the dense CTA-per-system kernel: one warm-up and one device-timed launch each, with the
    executed-flop model of the kernel that runs them (rehuel_b200/flopmodel.py) against the FP64 peak of this run."""
    import numpy as np
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    o = capi.make_options(RTOL, ATOL)
    out = {}

    def one(name, ens, t1, model, note):
        ms = device_ms(lambda: ens.solve(0.0, t1, DT0, o), torch)
        cs = counter_sums(ens, torch)
        fl = model_flops(model, cs)
        out[name] = {"ms": ms, "members": ens.M, "systems_per_s": ens.M / (ms * 1e-3), "attempts_per_member": cs["attempt"] / ens.M,
                     "max_attempts": cs["max_attempt"], "newton_rejections": cs["reject_newton"], "failed_members": cs["n_failed"],
                     "roofline": {"bound": "fp64", "achieved": fl / (ms * 1e-3) / 1e12, "peak": peak_tflops, "unit": "TFLOP/s",
                                  "frac": fl / (ms * 1e-3) / 1e12 / peak_tflops,
                                  "model": f"executed flops (flopmodel.py): fixed {model.fixed:g}/attempt, {model.per_iteration:g}/Newton iteration"},
                     "kernel": note}
        if cs["reject_newton"]:  # failed solves: their iterations are booked as in the reference but only partly executed
            out[name]["roofline"]["model"] += ("; the iterations of FAILED Newton solves are left out (a lane with a non-finite iterate "
                                               "stops early and books the rest): a lower bound")
            out[name]["roofline"]["frac_if_every_booked_iteration_ran"] = model_flops(model, cs, booked=True) / (ms * 1e-3) / 1e12 / peak_tflops
        return ens

    # config 3: Van der Pol sweep, LOBATTO_IIIC_43, thread-per-system; stiff end handed out first
    Mv = 1 << 20
    ens = DeviceEnsemble(capi.VAN_DER_POL, capi.LOBATTO_IIIC_43, Mv, cu(ensembles.VDPOL_Y0.copy()), cu(ensembles.vdpol_params(Mv)),
                         handout_order=capi.ORDER_LOCALITY)
    one("config3_vdpol_2p20_lobatto43", ens, 2.0, flopmodel.thread_kernel(2, capi.LOBATTO_IIIC_43, 6, 7), "ensemble_irk_thread<VanDerPolF,3,1,1,EST_IDENTITY>")
    del ens
    # config 4: 1-D Brusselator MOL n = 64, LOBATTO_IIIC_85, warp-per-system band kernel (a sixteenth of the 65 536 members)
    Mb = 4096
    ens = DeviceEnsemble(capi.BRUSSELATOR_1D, capi.LOBATTO_IIIC_85, Mb, cu(ensembles.bruss1d_y0(32)), cu(ensembles.bruss1d_params(Mb)), n_hint=64)
    one("config4_bruss1d_n64_lobatto85_4096", ens, 10.0, flopmodel.band_kernel(64, 2, capi.LOBATTO_IIIC_85, 10, 8), "ensemble_irk_warp<Brusselator1dF<32>,5,1,2,EST_OWN_LU> (band kernel)")
    del ens
    # the literal "CTA-per-system shared-memory block LU": the same system declared dense, short span
    Mc = 296
    ens = DeviceEnsemble(capi.BRUSSELATOR_1D_DENSE, capi.LOBATTO_IIIC_85, Mc, cu(ensembles.bruss1d_y0(32)), cu(ensembles.bruss1d_params(Mc)), n_hint=64)
    one("dense_cta_bruss1d_n64_lobatto85_t0.5", ens, 0.5, flopmodel.dense_kernel(64, capi.LOBATTO_IIIC_85, 10, 64 * 8), "ensemble_irk_cta<Brusselator1dDenseF<32>,5,1,2,EST_OWN_LU,256>")
    del ens
    return out


def strong_block(capi, ensembles, DeviceEnsemble, torch, dist, dev, rank, world, barrier, max_over_ranks, flopmodel, peak_tflops) -> dict:
    """Strong scaling: fixed totals over the N ranks.
    config 2 at 2^20 members in all: rank r integrates the contiguous members [r M/N, (r+1) M/N).
    config 3 at 2^20 Van der Pol members in all, dealt i mod N.
    config 5 at 10 485 760 members in all (even index Robertson / RADAU_IIA_53 / t in [0,40], odd index Van der Pol /
    LOBATTO_IIIC_43 / t in [0,2], rtol = atol = 10^(-3-6v), newton tol = 0.1 rtol): cut into 256 chunks that the ranks CLAIM
    from a cursor in shared memory (rehuel_cursor) as they finish, eight chunks in flight per GPU, the stiff end of the
    sweep first (longest processing time first)."""
    import numpy as np
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    o = capi.make_options(RTOL, ATOL)
    out = {}
    # ---- config 2, 2^20 members in all
    Mt = 1 << 20
    lo, hi = Mt * rank // world, Mt * (rank + 1) // world
    ens = DeviceEnsemble(capi.ROBERTSON, capi.RADAU_IIA_53, hi - lo, cu(ensembles.ROBERTSON_Y0.copy()),
                         cu(ensembles.robertson_params(hi - lo, start=lo)), handout_order=capi.ORDER_LOCALITY_DESC)
    ens.solve(0.0, 40.0, DT0, o)
    barrier()
    ms = max_over_ranks(device_ms(lambda: ens.solve(0.0, 40.0, DT0, o), torch, reps=3, warm=0))
    out["config2_robertson_2p20_total"] = {"ms": ms, "systems_per_s": Mt / (ms * 1e-3), "members_total": Mt}
    del ens
    # ---- config 3, 2^20 Van der Pol members in all (mu_ref = 10^(-6 i/(M-1)), LOBATTO_IIIC_43), member i on rank i mod N
    # (cost grows along the index), stiff end handed out first
    Mv = 1 << 20
    Pv = ensembles.vdpol_params(Mv)[:, rank::world]
    ens = DeviceEnsemble(capi.VAN_DER_POL, capi.LOBATTO_IIIC_43, Pv.shape[1], cu(ensembles.VDPOL_Y0.copy()), cu(Pv), handout_order=capi.ORDER_LOCALITY)
    ens.solve(0.0, 2.0, DT0, o)
    barrier()
    ms = max_over_ranks(device_ms(lambda: ens.solve(0.0, 2.0, DT0, o), torch, reps=2, warm=0))
    out["config3_vdpol_2p20_total"] = {"ms": ms, "systems_per_s": Mv / (ms * 1e-3), "members_total": Mv,
                                       "max_attempts_rank0": int(ens.counter("attempt").max()),
                                       "newton_rejections_rank0": int(ens.counter("reject_newton").sum())}
    del ens
    # ---- config 5, 10 485 760 members in all, chunks claimed from the shared cursor
    Mt, n_chunks, depth = 10485760, 256, 8
    half = Mt // 2
    per = half // n_chunks                      # members of each class per chunk
    name = f"rehuel_bench_{os.environ.get('MASTER_PORT', '0')}_{os.getppid() if world > 1 else os.getpid()}"
    cur = capi.Cursor(name, True) if rank == 0 else None
    barrier()
    if rank != 0:
        cur = capi.Cursor(name, False)
    tol_all = ensembles.mixed_tolerances(Mt)
    streams = [(torch.cuda.Stream(), torch.cuda.Stream()) for _ in range(depth)]

    def chunk_ens(c):
        j0 = c * per
        tr, tv = tol_all[0::2][j0:j0 + per], tol_all[1::2][j0:j0 + per]
        Pr, Pv = ensembles.robertson_params(per, start=j0), ensembles.vdpol_params(per, start=j0, total=half)
        pair = []
        for prob, method, t1, y0h, P, tl in ((capi.ROBERTSON, capi.RADAU_IIA_53, 40.0, ensembles.ROBERTSON_Y0, Pr, tr),
                                             (capi.VAN_DER_POL, capi.LOBATTO_IIIC_43, 2.0, ensembles.VDPOL_Y0, Pv, tv)):
            tl_d = cu(tl)
            keys = cu(np.vstack([tl[None, :], P[:2]]))          # the tolerance is a cost dimension like any parameter
            pair.append((DeviceEnsemble(prob, method, per, cu(y0h.copy()), cu(P), rtol=tl_d, atol=tl_d.clone(),
                                        newton_tol=(0.1 * tl_d).contiguous(), handout_order=capi.ORDER_LOCALITY, handout_keys=keys), t1))
        return pair

    # every rank may end up with any chunk: all of them are staged on every GPU beforehand (synthetic inputs, ~0.4 GB)
    chunks = [chunk_ens(c) for c in range(n_chunks)]

    def run_all():
        mine = []
        main = torch.cuda.current_stream()
        free, busy = list(range(depth)), {}
        while True:
            for sl in [sl for sl, e in busy.items() if e.query()]:   # a slot is free again when its chunk is done
                del busy[sl]
                free.append(sl)
            if not free:
                time.sleep(0)
                continue
            c = cur.next(1)                     # only a rank with a free slot claims more work
            if c >= n_chunks:
                break
            c = n_chunks - 1 - c                # the Van der Pol members get stiffer along the index: costly chunks first
            slot = free.pop()
            for (de, t1), st in zip(chunks[c], streams[slot]):
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    de.solve(0.0, t1, DT0, o)
            # the slot is done when BOTH its streams are: join them on the second stream
            streams[slot][1].wait_stream(streams[slot][0])
            e = torch.cuda.Event()
            e.record(streams[slot][1])
            busy[slot] = e
            mine.append(c)
        for pair in streams:
            for st in pair:
                main.wait_stream(st)
        return mine

    def timed():
        barrier()
        if rank == 0:
            cur.reset()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        mine = run_all()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), mine

    timed()                                      # warm-up (also the first touch of every chunk's buffers)
    ms, mine = timed()
    ms = max_over_ranks(ms)
    failed = sum(int((de.status != 0).sum()) for c in mine for de, _ in chunks[c])
    att = sum(int(de.counter("attempt").sum()) for c in mine for de, _ in chunks[c])
    # executed flops of the chunks this rank ran: the thread kernel's models of the two member classes (flopmodel.py)
    models = (flopmodel.thread_kernel(3, capi.RADAU_IIA_53, 11, 8, zero_col0_rows=(2,)), flopmodel.thread_kernel(2, capi.LOBATTO_IIIC_43, 6, 7))
    flops = sum(model_flops(models[k], counter_sums(de, torch)) for c in mine for k, (de, _) in enumerate(chunks[c]))
    t = torch.tensor([float(failed), float(att), float(len(mine)), float(flops)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t)
        lo_hi = torch.tensor([float(len(mine)), float(-len(mine))], dtype=torch.float64, device=dev)
        dist.all_reduce(lo_hi, op=dist.ReduceOp.MAX)
        most, least = int(lo_hi[0].item()), int(-lo_hi[1].item())
    else:
        most = least = len(mine)
    out["config5_mixed_10485760_total"] = {"ms": ms, "systems_per_s": Mt / (ms * 1e-3), "members_total": Mt, "chunks": n_chunks,
                                           "chunks_in_flight_per_gpu": depth, "chunks_claimed_min_max": [least, most],
                                           "failed_members": int(t[0].item()), "attempts_per_member": t[1].item() / Mt,
                                           "roofline": {"bound": "fp64", "achieved": t[3].item() / (ms * 1e-3) / 1e12 / world, "peak": peak_tflops,
                                                        "unit": "TFLOP/s per GPU", "frac": t[3].item() / (ms * 1e-3) / 1e12 / world / peak_tflops,
                                                        "model": "executed flops of ensemble_irk_thread for the Robertson (RADAU_IIA_53) and Van der Pol "
                                                                 "(LOBATTO_IIIC_43) halves, from the counters of every chunk (iterations of failed "
                                                                 "Newton solves left out: a lower bound); wall time of the whole pipeline (ordering, "
                                                                 "512 launches, chunk claiming) in the denominator"},
                                           "balance": "chunks claimed from a shared-memory cursor (rehuel_cursor) as GPUs finish"}
    barrier()
    cur.close()
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--span", choices=tuple(SPANS), default="short")
    ap.add_argument("--members", type=int, default=M_PER_GPU, help="members per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline + e2e only (no long span, other configs, strong block)")
    ap.add_argument("--long-steps", type=int, default=4, help="timed launches of the long-span block")
    ap.add_argument("--handout", choices=("index", "locality", "locality_desc"), default="locality_desc",
                    help="order in which members are handed to the GPU lanes (results do not depend on it)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
