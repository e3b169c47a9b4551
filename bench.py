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

# SURVEY.md §8(d) / BASELINE.md §3 per-element algorithmic work, fixed per config: the contract's
# numerator of `roofline.achieved`.  C4: "3-4x the fixed-mesh count" (SURVEY.md §8d), lower end.
ALG_FLOPS_PER_ELEM = {"c2_hex27_ns": 1.2e6, "c3_hex27_ns_energy": 1.85e6, "c5_hex8_pspg_T_2Y": 0.25e6,
                      "c4_hex27_ale_ns": 3.6e6}
# Counted, not estimated (BASELINE.md §3 "instrumented CPU restatement"):
#  restatement = floating-point operations oracle/fill_port.c executes per element, counted by running it with an
#                operation-counting number type (oracle/flop_count.cpp; `python bench.py --count-flops` re-counts).  The
#                restatement evaluates every (i,a; j,b) entry of App. A as written -- no factoring across (a,b) --
#                so it is an UPPER bound of the algorithmic work, 3x the survey's figure on the hex27 configs;
#  executed    = what the CUDA kernel issues per element.  hex27 tensor-core kernels: DMMA count x 512
#                flop (2016 node-pair + 385 set-up instructions on C2, 2576 + 420 on C3; the 27 -> 32 and 3 -> 8 tile
#                padding included) + the scalar FP64 instructions; C5 / C4: FP64-pipe active cycles of the ncu capture
#                x 64 lanes x 2 (an upper bound: every instruction counted as an FMA), profiles/r2j_fill_kernel_c*.txt.
COUNTED_FLOPS_PER_ELEM = {
    "c2_hex27_ns": {"restatement": 3.70e6, "executed": 1.33e6},
    "c3_hex27_ns_energy": {"restatement": 4.88e6, "executed": 1.65e6},
    "c5_hex8_pspg_T_2Y": {"restatement": 0.290e6, "executed": 0.21e6},
    "c4_hex27_ale_ns": {"restatement": 29.2e6, "executed": 14.1e6},
}
FP64_PEAK_NOMINAL_TFLOPS = 40.0  # BASELINE.json north_star; replaced by the DFMA micro-benchmark when it runs
CONFIGS = {  # name -> (roofline key, element type, default edge = the size BASELINE.json / SURVEY.md §8d name)
    "c2": ("c2_hex27_ns", "HEX27", 100),
    "c3": ("c3_hex27_ns_energy", "HEX27", 126),
    "c4": ("c4_hex27_ale_ns", "HEX27", 32),
    "c5": ("c5_hex8_pspg_T_2Y", "HEX8", 100),
}


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
    from goma_b200.state import make_state  # the seeded state of SURVEY.md §8d, shared with the parity tests

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
            "config": {"workload": workload_name(args.config, args.n), "elements_per_step": cores * ne},
            "cpu_baseline": {"value": v, "unit": "elements/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_name(cfg, n, scaling="weak", world=1, partition="brick"):
    per = (f"{n}^3 elements per GPU" if scaling == "weak" else
           f"ONE {n}^3 mesh split over {world} GPUs ({'bricks' if partition == 'brick' and world in (2, 4, 8) else 'x-slabs'})")
    if cfg == "c4":
        return f"3D Navier-Stokes on an ALE pseudo-solid mesh, Q2/P1 hex27, {per} (SURVEY C4, 3-D variant)"
    if cfg == "c5":
        return f"3D NS + energy + 2 species, PSPG Q1/Q1 hex8, {per} (BASELINE.json configs[4])"
    if cfg == "c3":
        return f"3D Boussinesq natural convection NS+energy, Q2/P1 hex27, {per} (BASELINE.json configs[2])"
    return f"3D lid-driven cavity Navier-Stokes, Q2/P1 hex27, {per} (BASELINE.json configs[1])"


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


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank (and with it the pinned host buffers it allocates: first touch) to the CPUs of the NUMA node its
    GPU hangs off, so that 8 ranks copying 45 GB each to the host do not all cross one socket's memory controller."""
    try:
        import torch

        prop = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (getattr(prop, "pci_domain_id", 0), prop.pci_bus_id, prop.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return {"numa_node": node, "note": "the box exposes no NUMA affinity for the GPU: ranks left unbound"}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"numa_node": node, "cpus": len(cpus)}
    except Exception:
        pass
    return None


def problem_maker(cfg):
    if cfg == "c5":
        return c5_problem_on
    if cfg == "c4":
        return c4_problem_on
    return lambda mesh: cavity_problem_on(mesh, cfg == "c3")


class Workload:
    """One config on this rank: problem, GPU context, sub-domain (N > 1)."""

    def __init__(self, cfg, n, rank, world, local_rank, scaling, scatter, layout="msr", partition="brick"):
        from goma_b200.matrix_fill import MatrixFill
        from goma_b200.mesh import box_mesh

        self.cfg, self.n, self.rank, self.world = cfg, n, rank, world
        self.key, elem, _ = CONFIGS[cfg]
        make_on = problem_maker(cfg)
        t0 = time.perf_counter()
        self.sub = None
        if world == 1:
            self.problem = make_on(box_mesh(elem, (n,) * 3))
            owned_nodes = None
            self.ne_owned = self.problem.mesh.num_elems
        else:
            from goma_b200.dp_comm import slab_subdomain

            if scaling == "strong" and partition == "brick" and world in (2, 4, 8):
                # ONE n^3 cavity cut into bricks (2x2x2 at 8 GPUs: every rank has 7 neighbours, ghost layers on its
                # three high sides: (n/2 + 1)^3 / (n/2)^3 = 6 % redundant elements against 12 % for x-slabs)
                from goma_b200.dp_comm import brick_subdomain

                self.parts = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
                self.sub = brick_subdomain(make_on, n, rank, self.parts, elem_type=elem)
            elif scaling == "strong":  # ... or into x-slabs of (almost) equal width
                cols = [round(r * n / world) for r in range(world + 1)]
                self.sub = slab_subdomain(make_on, n, rank, world, elem_type=elem, cols=cols, x_len=1.0)
            else:  # weak: rank r owns the r-th n^3 slab of a (world*n) x n x n box, plus its ghost column
                self.sub = slab_subdomain(make_on, n, rank, world, elem_type=elem)
            self.problem = self.sub.problem
            owned_nodes = self.sub.num_owned_nodes
            self.ne_owned = int(self.sub.elem_owned.sum())
        self.x = synthetic_state(self.problem, 20261017 + rank)
        self.mesh_gen_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        self.mf = MatrixFill(self.problem, device=local_rank, num_owned_nodes=owned_nodes, layout=layout)
        self.init_s = time.perf_counter() - t0
        self.mf.set_option("scatter", scatter)
        self.ne = self.problem.mesh.num_elems  # assembled per step on this rank (owned + ghost elements)
        self.n_unk, self.nnz = self.mf.num_unknowns, self.mf.nnz_plus
        self.hU = [0.0, 0.0]

    def step(self, dist=None, dev=None):
        mf = self.mf
        if self.sub is not None:
            mf.exchange_dof(0)  # ghost refresh (mm_sol_nonlinear.c:1273): own stream, overlaps the interior classes
        if self.problem.pspg:  # global_h_elem_siz / global_velocity_norm on the device (mm_sol_nonlinear.c:1184-1192)
            sums = mf.global_h_U(None if self.sub is None else self.sub.elem_owned)
            if self.world > 1:
                import torch

                t = torch.tensor(sums, dtype=torch.float64, device=dev)
                dist.all_reduce(t)
                sums = t.cpu().numpy()
            self.hU[0], self.hU[1] = sums[0] / sums[1], sums[2] / sums[3]
        mf.fill_device(h_elem_avg=self.hU[0], U_norm=self.hU[1])


def roofline_of(w, dev_ms, fp64_peak, hbm_peak, hbm_src, traffic=None):
    flops = ALG_FLOPS_PER_ELEM[w.key]
    bpe = 8.0 * (w.nnz + w.n_unk) / w.ne + w.problem.mesh.npe * 4 + 8.0 * (3 * w.problem.mesh.num_nodes + w.n_unk) / w.ne
    kern_s = dev_ms * 1e-3
    ach_tf = flops * w.ne / kern_s / 1e12
    ach_gbs = bpe * w.ne / kern_s / 1e9
    counted = COUNTED_FLOPS_PER_ELEM[w.key]
    roof = {"bound": "fp64", "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_tf / fp64_peak,
            "traffic": traffic, "peak_source": "fp64 matmul micro-benchmark in this run (nominal %.0f)" % FP64_PEAK_NOMINAL_TFLOPS,
            "flops_per_element": flops, "flops_per_element_source": "SURVEY.md §8d (the contract's per-unit figure)",
            "flops_per_element_counted": counted,
            "frac_executed": (counted["executed"] * w.ne / kern_s / 1e12 / fp64_peak) if counted["executed"] else None,
            "note": "binding roof is the FP64 pipe (AI ~25 flop/B, SURVEY.md §8d); frac_executed counts what the kernel "
                    "issues (tensor-core tile padding included)"}
    roof_hbm = {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                "traffic": traffic, "peak_source": hbm_src, "bytes_per_element": bpe}
    return roof, roof_hbm


def count_flops():
    """Re-count the restatement's operations per element (small meshes; the count per element is size-independent
    up to the Dirichlet rows)."""
    from goma_b200.mesh import box_mesh
    from oracle import port
    from goma_b200.state import make_state

    out = {}
    for cfg, (key, elem, _) in CONFIGS.items():
        p = problem_maker(cfg)(box_mesh(elem, (8,) * 3 if elem == "HEX8" else (4,) * 3))
        st = make_state(p, seed=1)
        h, U = (p.global_h_elem_siz(), p.global_velocity_norm(st["x"])) if p.pspg else (0.0, 0.0)
        f, tot = port.port_flops(p, st, h_elem_avg=h, U_norm=U)
        out[key] = {"flops_per_element": f, "ops": tot, "elements": p.mesh.num_elems}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--edge", dest="n", type=int, default=None,
                    help="elements per direction (default: the size the config is named on: c2 100, c3 126, c4 32, c5 100)")
    ap.add_argument("--energy", action="store_true", help="config C3 physics (NS + energy) instead of C2")
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "c5"],
                    help="c2 = headline (default); c3 = NS + energy; c4 = ALE; c5 = hex8 PSPG + T + 2 species")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = `edge`^3 elements per GPU (default); strong = ONE `edge`^3 mesh split over the GPUs")
    ap.add_argument("--partition", default="brick", choices=["brick", "slab"],
                    help="--scaling strong: 2x2x2 bricks (default at 2, 4, 8 GPUs) or x-slabs")
    ap.add_argument("--scatter", type=int, default=2,
                    help="0 fp64 atomics, 1 coloured load+add+store, 2 coloured first-touch stores (default)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true",
                    help="N = 1, headline config: skip the device-timed C3 / C5 / C4 lines under `configs`")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--host-stream-chunks", type=int, default=8,
                    help="N = 1 e2e leg: goma_gpu_problem.host_stream_chunks of the host-buffer context (0 = one sweep, one copy)")
    ap.add_argument("--count-flops", action="store_true", help="re-count the restatement's flops per element and exit")
    args = ap.parse_args()
    if args.count_flops:
        count_flops()
        return
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.energy:
        args.config = "c3"
    args.energy = args.config == "c3"
    if args.n is None:
        args.n = CONFIGS[args.config][2]
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

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the GPU fill has no CPU fallback")
    ge.build(quiet=True)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
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

    def timed_steps(w, steps, warmup, sampler=None):
        """`warmup` untimed + `steps` timed device-resident fills: (s/step wall max over ranks, device ms/step, launches)."""
        w.mf.matrix_fill_full(w.x, assemble_jacobian=False)  # state into HBM once (one residual-only host call)
        for _ in range(warmup):
            w.step(dist, dev)
        if sampler:
            sampler.mark()
        barrier()
        t0 = time.perf_counter()
        kernel_ms, launches = 0.0, 0
        for _ in range(steps):
            w.step(dist, dev)
            ms, nl = w.mf.last_stats()
            kernel_ms += ms
            launches += nl
        barrier()
        step_s = max_over_ranks((time.perf_counter() - t0) / steps)
        return step_s, max_over_ranks(kernel_ms / steps), launches

    from goma_b200.matrix_fill import MatrixFill, device_view

    w = Workload(args.config, args.n, rank, world, local_rank, args.scaling, args.scatter, partition=args.partition)
    mf, sub, problem = w.mf, w.sub, w.problem
    ne, n_unk, nnz, ne_owned, x = w.ne, w.n_unk, w.nnz, w.ne_owned, w.x
    setup = mf.setup_stats()
    bufs = mf.device_buffers()
    d_x = device_view(bufs.d_x, n_unk, dev)

    halo = None
    if sub is not None:
        # ghost refresh = goma_gpu_exchange_dof: one kernel pulling the ghost values out of the neighbours'
        # HBM over NVLink.  Checked once against the torch.distributed (NCCL send/recv) restatement.
        from goma_b200.dp_comm import exchange_dof, setup_peer_exchange

        setup_peer_exchange(mf, sub)
        mf.matrix_fill_full(x, assemble_jacobian=False)  # state into HBM
        ref = d_x.clone()
        exchange_dof(ref, sub)
        torch.cuda.synchronize(dev)
        mf.exchange_dof(0)
        torch.cuda.synchronize(dev)
        if not bool(torch.equal(ref, d_x)):
            raise SystemExit(f"rank {rank}: peer-memory exchange_dof differs from the NCCL send/recv result")
        halo = ("goma_gpu_exchange_dof (one pull kernel over NVLink peer memory) on its own stream before every fill, "
                "overlapped with the interior element classes; checked == NCCL send/recv")

    sampler = ClockSampler(local_rank)
    sampler.start()
    step_s, dev_ms, launches = timed_steps(w, args.steps, args.warmup, sampler)
    clocks = sampler.stop()
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
        # w = A v of the Newton line search (mm_sol_nonlinear.c:442-449) on the device-resident matrix: columns from
        # the node-level neighbour lists, the values streamed once
        vv = torch.ones(n_unk, dtype=torch.float64, device=dev)
        ts = []
        for _ in range(3):
            ev0.record(st_lib)
            mf.matvec(vv)
            ev1.record(st_lib)
            torch.cuda.synchronize(dev)
            ts.append(ev0.elapsed_time(ev1))
        post["matvec_ms"] = min(ts[1:])
        post["matvec_GB/s"] = 8.0 * nnz / post["matvec_ms"] / 1e6
        post["matvec_note"] = "goma_gpu_matvec: one read of the values (8 B per entry), no column-index array"
        del vv
        # CSR hand-off to a GPU solver (§8f-2): structure once, values re-gathered after every fill
        free_b, _ = torch.cuda.mem_get_info(dev)
        if 12.0 * (nnz + n_unk) > 0.9 * free_b:  # colind + values
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

        def host_steps(m):
            m.fill_raw(ptrs)  # warm-up
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                m.fill_raw(ptrs)
            barrier()
            return max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)

        e2e_s = host_steps(mf)
        streamed = None
        if world == 1 and args.host_stream_chunks > 1:
            # the host-buffer configuration of the library (goma_gpu_problem.host_stream_chunks): finished rows leave
            # for the host under the assembly of the later chunks.  Its own context: the class order is fixed at init.
            single_s, single_sum = e2e_s, float(ha.numpy()[: n_unk].sum())
            mf.close()
            torch.cuda.empty_cache()
            mf = MatrixFill(problem, device=local_rank, host_stream_chunks=args.host_stream_chunks)
            mf.set_option("scatter", args.scatter)
            bufs = mf.device_buffers()
            ha.numpy()[: n_unk] = 0.0
            e2e_s = host_steps(mf)
            streamed = {"host_stream_chunks": args.host_stream_chunks, "single_copy_ms_per_step": single_s * 1e3,
                        "diag_checksum": [single_sum, float(ha.numpy()[: n_unk].sum())]}
        # the host-copy roof beside it: what one D2H of the matrix alone takes on this link
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d_a = device_view(bufs.d_a, nnz + 1, dev)
        ev0.record()
        ha.copy_(d_a, non_blocking=True)
        ev1.record()
        torch.cuda.synchronize(dev)
        d2h_ms = max_over_ranks(ev0.elapsed_time(ev1))
        e2e = {"value": total_elems / e2e_s, "unit": "elements/s", "h2d_bytes_per_step": 8 * n_unk,
               "d2h_bytes_per_step": 8 * (nnz + 1) + 8 * n_unk, "ms_per_step": e2e_s * 1e3,
               "host_buffers": "pinned" if pinned else "pageable", "steps": args.e2e_steps,
               "resid_checksum": float(hr.numpy().sum()), "streaming": streamed,
               "host_copy_roof": {"d2h_matrix_ms": d2h_ms, "GB/s": 8e-6 * (nnz + 1) / d2h_ms,
                                  "elements_per_s_at_roof": total_elems / (d2h_ms * 1e-3),
                                  "note": "one cudaMemcpy D2H of the MSR values alone, all ranks at once (max over ranks): "
                                          "the floor of any host-buffer step"}}
        del hx, ha, hr, d_a

    line = None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    traffic_tbl = {}
    try:
        traffic_tbl = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except OSError:
        pass

    def traffic_of(wk, nlaunch, nsteps):  # measured DRAM bytes per launch (ncu capture summarised under profiles/)
        if wk.key in traffic_tbl and nlaunch:
            return traffic_tbl[wk.key]["bytes_per_element"] * wk.ne * nsteps / nlaunch
        return None

    if rank == 0:
        fp64_peak = max(fp64_peak_tflops(dev), 1e-9)
        roof, roof_hbm = roofline_of(w, dev_ms, fp64_peak, hbm_peak, hbm_src, traffic_of(w, launches, args.steps))
        line = {
            "metric": "jacobian_residual_elements_per_s", "value": value, "unit": "elements/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
            "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.config, args.n, args.scaling, world, args.partition), "elements_per_gpu": ne_owned,
                       "neighbors": (len(sub.neighbors) if sub is not None else 0),
                       "elements_assembled_per_gpu": ne, "unknowns_per_gpu": n_unk,
                       "halo": halo, "numa_binding": numa,
                       "nnz_per_gpu": nnz, "scatter": ["fp64 atomics", "coloured load+add+store", "coloured first-touch stores"][args.scatter],
                       "l2": "inputs larger than L2 (MSR values %.1f GB per GPU rewritten every step)" % (8e-9 * nnz),
                       "setup_s": round(setup["total_s"], 2),
                       "setup": {"MatrixFill_constructor_s (python struct marshalling + goma_gpu_fill_init)": round(w.init_s, 3),
                                 **{"goma_gpu_fill_init_" + k: round(v, 3) for k, v in setup.items()},
                                 "synthetic_mesh_and_state_s (python harness, not the product)": round(w.mesh_gen_s, 2)}},
            "clocks": clocks, "gpu_launches": launches, "device_ms_per_step": dev_ms,
            "roofline": roof, "roofline_hbm": roof_hbm,
        }
        if post:
            post["hbm_frac"] = post["GB/s"] / hbm_peak
            line["post_fill"] = post
        if e2e:
            line["e2e"] = e2e
    mf.close()
    del mf, d_x
    torch.cuda.empty_cache()

    # ---- the other named configs, device-timed, one after the other on the freed GPU (N = 1, headline run only)
    if world == 1 and args.config == "c2" and not args.no_extra_configs and rank == 0:
        extra = {}
        for name, cfg, layout in (("c3", "c3", "msr"), ("c5", "c5", "msr"), ("c4", "c4", "msr"), ("c3_csr_layout", "c3", "csr")):
            try:
                wk = Workload(cfg, CONFIGS[cfg][2], 0, 1, local_rank, "weak", args.scatter, layout=layout)
                s_s, d_ms, nl = timed_steps(wk, 3, 3)
                r, rh = roofline_of(wk, d_ms, fp64_peak, hbm_peak, hbm_src, traffic_of(wk, nl, 3))
                extra[name] = {"workload": workload_name(cfg, wk.n), "value": wk.ne_owned / s_s, "unit": "elements/s",
                               "ms_per_step": s_s * 1e3, "device_ms_per_step": d_ms, "steps": 3, "warmup": 3,
                               "gpu_launches": nl, "elements": wk.ne, "unknowns": wk.n_unk, "nnz": wk.nnz,
                               "setup_s": round(wk.mf.setup_stats()["total_s"], 2), "roofline": r, "roofline_hbm": rh}
                if layout == "csr":
                    # the single-copy solver hand-off: the fill scatters straight into the CSR values
                    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    st_lib = torch.cuda.ExternalStream(wk.mf.device_buffers().stream, device=dev)
                    ts = []
                    for _ in range(3):
                        ev0.record(st_lib)
                        wk.mf.row_sum_scale(want_scale=False)
                        ev1.record(st_lib)
                        torch.cuda.synchronize(dev)
                        ts.append(ev0.elapsed_time(ev1))
                    rowptr, values = wk.mf.csr_rows()
                    extra[name]["layout"] = ("CSR of the owned rows assembled in place (diagonal at its sorted position): rowptr + values "
                                             "handed to a GPU solver with no second copy of the matrix")
                    extra[name]["csr_nnz"] = int(values.numel())
                    extra[name]["row_sum_scale_ms"] = min(ts[1:])
                    extra[name]["row_sum_scale_GB/s"] = 16.0 * values.numel() / min(ts[1:]) / 1e6
                    extra[name]["row_sum_scale_hbm_frac"] = extra[name]["row_sum_scale_GB/s"] / hbm_peak
                    del rowptr, values
                wk.mf.close()
                del wk
                torch.cuda.empty_cache()
            except Exception as ex:  # an extra line must never take the headline down with it
                extra[cfg] = {"error": str(ex)[:300]}
        line["configs"] = extra

    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(energy=args.energy)
            except Exception as ex:  # the baseline must never take the GPU number down with it
                line["cpu_baseline"] = {"value": None, "unit": "elements/s", "cores": 0, "kind": "reference",
                                        "sample": f"failed: {ex}"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
