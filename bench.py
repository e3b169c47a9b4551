#!/usr/bin/env python
"""Benchmark of the matrix_fill hot path (BASELINE.json: Jacobian+residual elements/s).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle/_ref)

One "step" = one matrix_fill_full over the whole local mesh (residual + Jacobian): zero the MSR
values and residual, assemble every element, scatter.  At N=1 the workload is BASELINE.json
configs[1]: 3-D lid-driven cavity, Q2/P1 hex27, 100^3 = 1M elements (SURVEY.md §8d "C2").  At
N>1 every rank assembles its own 1M-element slab of an N-times longer box (weak scaling; the
element loop has no data-path collective -- SURVEY.md §8e).

`value`  : device-resident step (state already in HBM), elements/s over all ranks.
`e2e`    : the same step through goma_gpu_fill() with HOST buffers: H2D of x, assembly, D2H of
           the MSR values and the residual, all inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# SURVEY.md §8(d) / BASELINE.md §3 per-element algorithmic work, fixed per config
ALG_FLOPS_PER_ELEM = {"c2_hex27_ns": 1.2e6, "c3_hex27_ns_energy": 1.85e6, "c5_hex8_pspg_T_2Y": 0.25e6,
                      "c4_hex27_ale_ns": 3.6e6}  # C4: "3-4x the fixed-mesh count" (SURVEY.md §8d), lower end
FP64_PEAK_NOMINAL_TFLOPS = 40.0  # BASELINE.json north_star; replaced by the DFMA micro-benchmark when it runs


def cavity_problem(n, energy=False, x_len=1.0):
    from goma_b200.mesh import box_mesh
    from goma_b200.problem import Dirichlet, Problem

    m = box_mesh("HEX27", (n, n, n), lo=(0, 0, 0), hi=(x_len, 1.0, 1.0))
    bcs = [Dirichlet(v, s, 0.0) for s in (1, 2, 3, 4, 5) for v in "UVW"]
    bcs += [Dirichlet("U", 6, 1.0), Dirichlet("V", 6, 0.0), Dirichlet("W", 6, 0.0), Dirichlet("P", 7, 0.0)]
    kw = {}
    if energy:
        bcs += [Dirichlet("T", 1, 1.0), Dirichlet("T", 2, 0.0)]
        kw = dict(energy=True, k=0.0141, Cp=1.0, beta=1.0, Tref=0.0, gravity=(0.0, 0.0, -1.0), ns_source="BOUSSINESQ")
    return Problem(m, rho=1.0, mu=0.01, bcs=bcs, **kw)


def c5_problem_on(mesh):
    """BASELINE.json configs[4] / SURVEY.md §8d C5: hex8 Q1/Q1 PSPG (local) + energy + 2 species."""
    from goma_b200.problem import Dirichlet, Problem

    bcs = [Dirichlet(v, s, 0.0) for s in (1, 2, 3, 4, 5) for v in "UVW"]
    bcs += [Dirichlet("U", 6, 1.0), Dirichlet("V", 6, 0.0), Dirichlet("W", 6, 0.0), Dirichlet("P", 7, 0.0),
            Dirichlet("T", 1, 1.0), Dirichlet("T", 2, 0.0), Dirichlet("Y", 1, 0.7, species=0), Dirichlet("Y", 2, 0.2, species=1)]
    return Problem(mesh, interp="Q1Q1", pspg="local", ps_scaling=0.1, energy=True, n_species=2, rho=1.0, mu=0.01,
                   k=0.0141, Cp=1.0, beta=1.0, Tref=0.0, gravity=(0.0, 0.0, -1.0), ns_source="BOUSSINESQ",
                   diffusivity=(0.01, 0.02, 1.0, 1.0), bcs=bcs)


def c4_problem_on(mesh):
    """SURVEY.md §8d C4, 3-D variant: hex27 Q2/P1 Navier-Stokes on an ALE pseudo-solid mesh (ARBITRARY / NONLINEAR,
    lambda = mu = 1), volumetric assembly with Dirichlet conditions on the displacements."""
    from goma_b200.problem import Dirichlet, Problem

    bcs = [Dirichlet(v, s, 0.0) for s in (1, 2, 3, 4, 5) for v in ("U", "V", "W", "DX", "DY", "DZ")]
    bcs += [Dirichlet("U", 6, 1.0), Dirichlet("V", 6, 0.0), Dirichlet("W", 6, 0.0), Dirichlet("DZ", 6, 0.0),
            Dirichlet("P", 7, 0.0)]
    return Problem(mesh, ale=True, rho=1.0, mu=0.01, lame_mu=1.0, lame_lambda=1.0, bcs=bcs)


def cavity_problem_on(mesh, energy=False):
    """Same cards as cavity_problem on a given (sub-domain) mesh: empty node sets simply match nothing."""
    from goma_b200.problem import Dirichlet, Problem

    bcs = [Dirichlet(v, s, 0.0) for s in (1, 2, 3, 4, 5) for v in "UVW"]
    bcs += [Dirichlet("U", 6, 1.0), Dirichlet("V", 6, 0.0), Dirichlet("W", 6, 0.0), Dirichlet("P", 7, 0.0)]
    kw = {}
    if energy:
        bcs += [Dirichlet("T", 1, 1.0), Dirichlet("T", 2, 0.0)]
        kw = dict(energy=True, k=0.0141, Cp=1.0, beta=1.0, Tref=0.0, gravity=(0.0, 0.0, -1.0), ns_source="BOUSSINESQ")
    return Problem(mesh, rho=1.0, mu=0.01, bcs=bcs, **kw)


def synthetic_state(problem, seed):
    from tests.cases import make_state

    return make_state(problem, seed=seed)["x"]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        """One resident nvidia-smi in loop mode, started before the warm-up so that its start-up
        (which holds driver locks) stays outside the timed region; mark() opens the window."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 10:
                time.sleep(0.05)
        except OSError:
            self.proc = None
        self.first = 0

    def mark(self):
        self.first = len(self.rows)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.rows = self.rows[max(self.first - 1, 0):]  # samples taken during the timed window
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i] == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def reference_sample(cores, sample_n, nrep, energy=False):
    """The reference's own matrix_fill_full (oracle/_ref) on `cores` independent single-rank processes,
    one `sample_n`^3 hex27 sub-domain each.  Returns (elements/s, mean seconds/step)."""
    from oracle import ref_driver

    p = cavity_problem(sample_n, energy)
    x = synthetic_state(p, 1)
    n = len(x)
    tmp = tempfile.mkdtemp(prefix="goma_ref_bench_")
    dirs = []
    for c in range(cores):
        wd = os.path.join(tmp, f"p{c}")
        ref_driver.write_workdir(p, wd)
        with open(os.path.join(wd, "state.bin"), "wb") as f:
            f.write(np.array([n, 1, 1, 0], np.int32).tobytes())
            f.write(np.array([0.0, 0.0, 0.0, -1.0, -1.0], np.float64).tobytes())
            f.write(x.tobytes())
            f.write(np.zeros(4 * n).tobytes())
        dirs.append(wd)
    procs = [subprocess.Popen([ref_driver.REF_EXE, wd, "fill", str(nrep)], stdout=subprocess.PIPE,
                              stderr=subprocess.DEVNULL, text=True) for wd in dirs]
    means = []
    for pr in procs:
        out = pr.communicate()[0]
        if pr.returncode != 0:
            raise RuntimeError("goma_ref_fill failed in the CPU baseline")
        line = [l for l in out.splitlines() if l.startswith("fill[0]")][-1]
        means.append(float(line.split("mean_s=")[1].split()[0]))
    subprocess.call(["rm", "-rf", tmp])
    step_s = max(means)
    return cores * p.mesh.num_elems / step_s, step_s, p.mesh.num_elems


def port_sample(sample_n, nrep, energy=False):
    from goma_b200 import capi
    from oracle import port

    p = cavity_problem(sample_n, energy)
    st = {"x": synthetic_state(p, 1)}
    ija = capi.pattern_msr(p)
    t = []
    for _ in range(nrep):
        t0 = time.perf_counter()
        port.port_fill(p, ija, st)
        t.append(time.perf_counter() - t0)
    return p.mesh.num_elems / min(t), min(t), p.mesh.num_elems


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_baseline(budget_cores=None, nrep=8, energy=False):
    from oracle import ref_driver

    cores = budget_cores or host_cores()
    if ref_driver.ref_available():
        v, step_s, ne = reference_sample(cores, 8, nrep, energy)
        return {"value": v, "unit": "elements/s", "cores": cores, "kind": "reference",
                "sample": f"{cores} single-rank processes of the reference's matrix_fill_full, one 8^3 hex27 "
                          f"sub-domain ({ne} elements) each, {nrep} fills, mean {step_s:.3f} s/fill"}
    v, step_s, ne = port_sample(6, nrep, energy)
    return {"value": v, "unit": "elements/s", "cores": 1, "kind": "port",
            "sample": f"oracle/fill_port.c, one 6^3 hex27 mesh ({ne} elements), best of {nrep}"}


def run_reference_arm(args, rank):
    """--impl reference: the reference CPU path on all host cores, same config/metric/unit."""
    if rank != 0:
        return
    from oracle import ref_driver

    cores = host_cores()
    nrep = args.warmup + args.steps
    if ref_driver.ref_available():
        v, step_s, ne = reference_sample(cores, 8, nrep, args.energy)
        kind = "reference"
        sample = (f"{cores} single-rank reference processes x one 8^3 hex27 sub-domain ({ne} elements) each; "
                  f"{nrep} fills per process, mean s/fill of the slowest process")
    else:
        v, step_s, ne = port_sample(6, nrep, args.energy)
        cores, kind = 1, "port"
        sample = f"oracle/fill_port.c on one 6^3 hex27 mesh ({ne} elements)"
    line = {"impl": "reference", "metric": "jacobian_residual_elements_per_s", "value": v, "unit": "elements/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "elements_per_step": cores * ne},
            "cpu_baseline": {"value": v, "unit": "elements/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_name(args):
    if getattr(args, "config", "c2") == "c4":
        return (f"3D Navier-Stokes on an ALE pseudo-solid mesh, Q2/P1 hex27, {args.n}^3 elements per GPU (SURVEY C4, 3-D "
                "variant; not the headline config)")
    if getattr(args, "config", "c2") == "c5":
        return (f"3D NS + energy + 2 species, PSPG Q1/Q1 hex8, {args.n}^3 elements per GPU (BASELINE.json configs[4]; "
                "not the headline config)")
    phys = "NS+energy (Boussinesq)" if args.energy else "Navier-Stokes"
    return f"3D lid-driven cavity {phys}, Q2/P1 hex27, {args.n}^3 elements per GPU (BASELINE.json configs[1])"


def fp64_peak_tflops(device):
    """DFMA micro-benchmark through torch (fp64 matmul runs on the FP64 pipe): measured peak for the roofline."""
    import torch

    n = 4096
    a = torch.randn(n, n, dtype=torch.float64, device=device)
    b = torch.randn(n, n, dtype=torch.float64, device=device)
    torch.matmul(a, b)
    torch.cuda.synchronize(device)
    best = 0.0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize(device)
        best = max(best, 2 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    del a, b
    torch.cuda.empty_cache()
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--edge", dest="n", type=int, default=100, help="elements per direction per GPU (100 -> 1M hex27 elements)")
    ap.add_argument("--energy", action="store_true", help="config C3 physics (NS + energy) instead of C2")
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "c5"],
                    help="c2 = headline (default); c3 = --energy; c5 = hex8 PSPG + T + 2 species (kernel number only)")
    ap.add_argument("--scatter", type=int, default=2,
                    help="0 fp64 atomics, 1 coloured load+add+store, 2 coloured first-touch stores (default)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.config == "c3":
        args.energy = True
    if args.config in ("c4", "c5") and args.impl == "reference":
        raise SystemExit("the reference arm is defined on the headline config (c2) and c3")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge
    from goma_b200.matrix_fill import MatrixFill

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the GPU fill has no CPU fallback")
    ge.build(quiet=True)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    t_setup = time.perf_counter()
    sub = None
    make_on = (c5_problem_on if args.config == "c5" else c4_problem_on if args.config == "c4"
               else (lambda mesh: cavity_problem_on(mesh, args.energy)))
    if world == 1:
        if args.config in ("c4", "c5"):
            from goma_b200.mesh import box_mesh

            problem = make_on(box_mesh("HEX8" if args.config == "c5" else "HEX27", (args.n,) * 3))
        else:
            problem = cavity_problem(args.n, args.energy)
        num_owned_nodes = None
        ne_owned = problem.mesh.num_elems
    else:
        # weak scaling: rank r owns the r-th n^3 slab of a (world*n) x n x n cavity, plus its ghost column
        from goma_b200.dp_comm import exchange_dof, setup_peer_exchange, slab_subdomain

        sub = slab_subdomain(make_on, args.n, rank, world, elem_type="HEX8" if args.config == "c5" else "HEX27")
        problem = sub.problem
        num_owned_nodes = sub.num_owned_nodes
        ne_owned = int(sub.elem_owned.sum())
    x = synthetic_state(problem, 20261017 + rank)
    mf = MatrixFill(problem, device=local_rank, num_owned_nodes=num_owned_nodes)
    mf.set_option("scatter", args.scatter)
    ne = problem.mesh.num_elems  # assembled per step on this rank (owned + ghost elements)
    n_unk, nnz = mf.num_unknowns, mf.nnz_plus
    t_setup = time.perf_counter() - t_setup

    from goma_b200.matrix_fill import device_view

    bufs = mf.device_buffers()
    d_x = device_view(bufs.d_x, n_unk, dev)

    trace = [] if os.environ.get("GOMA_BENCH_TRACE") else None
    halo = None
    if sub is not None:
        # ghost refresh = goma_gpu_exchange_dof: one kernel pulling the ghost values out of the neighbours'
        # HBM over NVLink.  Checked once against the torch.distributed (NCCL send/recv) restatement.
        setup_peer_exchange(mf, sub)
        mf.matrix_fill_full(x, assemble_jacobian=False)  # state into HBM
        ref = d_x.clone()
        exchange_dof(ref, sub)
        torch.cuda.synchronize(dev)
        mf.exchange_dof(0)
        torch.cuda.synchronize(dev)
        same = bool(torch.equal(ref, d_x))
        if not same:
            raise SystemExit(f"rank {rank}: peer-memory exchange_dof differs from the NCCL send/recv result")
        halo = "goma_gpu_exchange_dof (one pull kernel over NVLink peer memory) before every fill; checked == NCCL send/recv"

    hU = [0.0, 0.0]

    def step():
        t0 = time.perf_counter()
        if sub is not None:
            mf.exchange_dof(0)  # ghost refresh before the fill (mm_sol_nonlinear.c:1273), same stream as the fill
        if problem.pspg:  # global_h_elem_siz / global_velocity_norm on the device (mm_sol_nonlinear.c:1184-1192)
            sums = mf.global_h_U(None if sub is None else sub.elem_owned)
            if world > 1:
                t = torch.tensor(sums, dtype=torch.float64, device=dev)
                dist.all_reduce(t)
                sums = t.cpu().numpy()
            hU[0], hU[1] = sums[0] / sums[1], sums[2] / sums[3]
        t1 = time.perf_counter()
        mf.fill_device(h_elem_avg=hU[0], U_norm=hU[1])
        if trace is not None:
            trace.append((t1 - t0, time.perf_counter() - t1))

    # state into HBM once (one residual-only host call), then the device-resident steps
    sampler = ClockSampler(local_rank)
    sampler.start()
    mf.matrix_fill_full(x, assemble_jacobian=False)
    for _ in range(args.warmup):
        step()
    sampler.mark()
    barrier()
    t0 = time.perf_counter()
    kernel_ms, launches = 0.0, 0
    for _ in range(args.steps):
        step()
        ms, nl = mf.last_stats()
        kernel_ms += ms
        launches += nl
    barrier()
    step_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    clocks = sampler.stop()
    if trace is not None:
        print(f"[bench trace rank {rank}] (exchange_s, fill_s) per step:", [(round(a * 1e3, 3), round(b * 1e3, 3)) for a, b in trace[-args.steps:]],
              file=sys.stderr)
    dev_ms = max_over_ranks(kernel_ms / args.steps)  # CUDA events on the library's stream: memsets + kernel(s)
    total_elems = sum_over_ranks(float(ne_owned))  # ghost elements are assembled twice but counted once
    value = total_elems / step_s

    # ---- the pass that follows the fill in the Newton loop, HBM-bound: row-sum scaling of the
    #      device-resident matrix (reads and writes every value once).  Not part of `value`.
    post = None
    if world == 1:
        mf.fill_device()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st_lib = torch.cuda.ExternalStream(bufs.stream, device=dev)
        ts = []
        for _ in range(3):
            ev0.record(st_lib)
            mf.row_sum_scale(want_scale=False)
            ev1.record(st_lib)
            torch.cuda.synchronize(dev)
            ts.append(ev0.elapsed_time(ev1))
        ms = min(ts[1:])
        post = {"row_sum_scale_ms": ms, "GB/s": 16.0 * nnz / ms / 1e6, "bytes": 16 * nnz,
                "note": "row_sum_scaling_scale on device: one read + one write of the MSR values"}
        # CSR hand-off to a GPU solver (§8f-2): structure once, values re-gathered after every fill
        free_b, _ = torch.cuda.mem_get_info(dev)
        if 12.0 * (nnz + n_unk) + 4.0 * nnz / 8 > 0.9 * free_b:  # colind + values (+ transient node-node lists)
            post["csr"] = "skipped: %.0f GB free on the device, the CSR copy needs %.0f GB" % (free_b / 1e9, 12e-9 * nnz)
            return_csr = False
        else:
            return_csr = True
        t0 = time.perf_counter()
        rowptr, colind, values = mf.csr(refresh_values=False) if return_csr else (None, None, None)
        torch.cuda.synchronize(dev)
        if return_csr:
            post["csr_structure_s"] = time.perf_counter() - t0
            ts = []
            for _ in range(3):
                ev0.record(st_lib)
                mf.csr(refresh_values=True)
                ev1.record(st_lib)
                torch.cuda.synchronize(dev)
                ts.append(ev0.elapsed_time(ev1))
            post["csr_values_ms"] = min(ts[1:])
            post["csr_values_GB/s"] = 16.0 * values.numel() / post["csr_values_ms"] / 1e6
        del rowptr, colind, values

    # ---- end to end through the C ABI with host buffers (pinned), copies inside the timed region
    e2e = None
    serial_e2e = False
    if not args.no_e2e and world > 1:
        # every rank pins nnz*8 bytes of host memory for the matrix: if the box cannot hold all of them at once the
        # ranks take turns (their PCIe links are independent; what is lost is only the overlap between ranks)
        import psutil

        need = 8.0 * (nnz + 1 + 2 * n_unk)
        avail = float(psutil.virtual_memory().available)
        serial_e2e = sum_over_ranks(need) > 0.6 * avail
    if not args.no_e2e and serial_e2e:
        t_tot = 0.0
        checksum = 0.0
        for turn in range(world):
            barrier()
            if turn == rank:
                hx = torch.empty(n_unk, dtype=torch.float64).pin_memory()
                ha = torch.empty(nnz + 1, dtype=torch.float64).pin_memory()
                hr = torch.empty(n_unk, dtype=torch.float64).pin_memory()
                hx.numpy()[:] = x
                ptrs = (hx.data_ptr(), 0, 0, 0, 0, ha.data_ptr(), hr.data_ptr())
                mf.fill_raw(ptrs)
                t0 = time.perf_counter()
                for _ in range(args.e2e_steps):
                    mf.fill_raw(ptrs)
                t_tot = (time.perf_counter() - t0) / args.e2e_steps
                checksum = float(hr.numpy().sum())
                del hx, ha, hr
        barrier()
        e2e_s = sum_over_ranks(t_tot)
        e2e = {"value": total_elems / e2e_s, "unit": "elements/s", "h2d_bytes_per_step": 8 * n_unk,
               "d2h_bytes_per_step": 8 * (nnz + 1) + 8 * n_unk, "ms_per_step": e2e_s * 1e3, "host_buffers": "pinned",
               "steps": args.e2e_steps, "resid_checksum": checksum,
               "note": "ranks measured one after the other (host RAM cannot pin every rank's matrix buffer at once); "
                       "value = all elements / sum of the per-rank times"}
    elif not args.no_e2e:
        try:
            hx = torch.empty(n_unk, dtype=torch.float64).pin_memory()
            ha = torch.empty(nnz + 1, dtype=torch.float64).pin_memory()
            hr = torch.empty(n_unk, dtype=torch.float64).pin_memory()
            pinned = True
        except RuntimeError:
            hx, ha, hr = (torch.empty(k, dtype=torch.float64) for k in (n_unk, nnz + 1, n_unk))
            pinned = False
        hx.numpy()[:] = x
        ptrs = (hx.data_ptr(), 0, 0, 0, 0, ha.data_ptr(), hr.data_ptr())
        mf.fill_raw(ptrs)  # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            mf.fill_raw(ptrs)
        barrier()
        e2e_s = max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
        e2e = {"value": total_elems / e2e_s, "unit": "elements/s", "h2d_bytes_per_step": 8 * n_unk,
               "d2h_bytes_per_step": 8 * (nnz + 1) + 8 * n_unk, "ms_per_step": e2e_s * 1e3,
               "host_buffers": "pinned" if pinned else "pageable", "steps": args.e2e_steps,
               "resid_checksum": float(hr.numpy().sum())}
        del hx, ha, hr

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        key = "c5_hex8_pspg_T_2Y" if args.config == "c5" else "c4_hex27_ale_ns" if args.config == "c4" else ("c3_hex27_ns_energy" if args.energy else "c2_hex27_ns")
        flops = ALG_FLOPS_PER_ELEM[key]
        bytes_per_elem = 8.0 * (nnz + n_unk) / ne + problem.mesh.npe * 4 + 8.0 * (3 * problem.mesh.num_nodes + n_unk) / ne
        kern_s = dev_ms * 1e-3
        traffic = None  # measured DRAM bytes per launch (ncu capture summarised under profiles/)
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")))
            if key in tj and launches:
                traffic = tj[key]["bytes_per_element"] * ne * args.steps / launches
        except OSError:
            pass
        fp64_meas = fp64_peak_tflops(dev)
        fp64_peak = max(fp64_meas, 1e-9)
        ach_tf = flops * ne / kern_s / 1e12
        ach_gbs = bytes_per_elem * ne / kern_s / 1e9
        line = {
            "metric": "jacobian_residual_elements_per_s", "value": value, "unit": "elements/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "elements_per_gpu": ne_owned,
                       "elements_assembled_per_gpu": ne, "unknowns_per_gpu": n_unk,
                       "halo": halo,
                       "nnz_per_gpu": nnz, "scatter": ["fp64 atomics", "coloured load+add+store", "coloured first-touch stores"][args.scatter],
                       "l2": "inputs larger than L2 (MSR values %.1f GB per GPU rewritten every step)" % (8e-9 * nnz),
                       "setup_s": round(t_setup, 1)},
            "clocks": clocks, "gpu_launches": launches, "device_ms_per_step": dev_ms,
            "roofline": {"bound": "fp64", "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": ach_tf / fp64_peak, "traffic": traffic,
                         "peak_source": "fp64 matmul micro-benchmark in this run (nominal %.0f)" % FP64_PEAK_NOMINAL_TFLOPS,
                         "flops_per_element": flops,
                         "note": "binding roof of the hex27 fill is the FP64 pipe (AI ~27 flop/B, SURVEY.md §8d)"},
            "roofline_hbm": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s",
                             "frac": ach_gbs / hbm_peak, "traffic": traffic, "peak_source": hbm_src,
                             "bytes_per_element": bytes_per_elem},
        }
        if post:
            post["hbm_frac"] = post["GB/s"] / hbm_peak
            line["post_fill"] = post
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(energy=args.energy)
            except Exception as ex:  # the baseline must never take the GPU number down with it
                line["cpu_baseline"] = {"value": None, "unit": "elements/s", "cores": 0, "kind": "reference",
                                        "sample": f"failed: {ex}"}
        print(json.dumps(line))
    mf.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
