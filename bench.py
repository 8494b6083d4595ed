#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on its own configuration.

metric   GCG solve seconds (lower is better): wall time of one whole block-GCG solve --
         ops->EigenSolver, i.e. InitializeX .. last Rayleigh-Ritz, matrix build/upload excluded
         (reference test/test_eig_sol_gcg.c:88-144 "Time is") -- of the P1-FEM stiffness/mass
         pencil A x = lambda B x on the Kuhn triangulation, n = m^3 (default m = 200, n = 8.0 M,
         ~15 nnz/row), nev = 200 (nevMax 400, block_size 40), reference default tolerances.
step     one whole solve of that pencil from srand(0).
value    seconds per solve, inputs (A, B) resident in HBM and the four GCG workspaces created
         beforehand (the reference's driver creates gcg_mv_ws[0..3] before its clock starts,
         test/test_eig_sol_gcg.c:57-68,88), CUDA events on the library stream.
e2e      the same solve through the C-ABI with HOST buffers: CCS arrays of A and B uploaded
         from page-locked host memory and turned into device matrices, workspaces allocated,
         solve, eigenvalues + the nev converged eigenvectors copied back to the host, workspaces
         freed -- all inside the timed region.
roofline the dominant kernel class of the timed region (device time from CUDA events around
         every launch of the class, algorithmic bytes/flops of SURVEY.md 8d / DESIGN.md).
cpu_baseline / --impl reference
         the UNMODIFIED reference (oracle/_ref: CCS + OpenMP GCG, OpenBLAS) on the box's host
         cores, on a bounded sample of the same workload (same pencil family, nev, block size
         and tolerances on a smaller lattice, solved to convergence), scaled linearly in n and
         in the outer-iteration count to the full workload -- see `sample` in the JSON line.

Multi-GPU (torchrun, one rank per GPU): the SAME pencil is split into 1-D row blocks over the N
GPUs (strong scaling): halo rows of SpMM by ncclSend/ncclRecv, Gram blocks / dots / CG scalars by
ncclAllReduce, projected problem replicated.  value = seconds per solve, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# outer iterations the full workload needs (measured on B200 with this repo, recorded by
# bench.py itself into profiles/bench_iters.json); the reference arm scales its bounded
# sample with it.  Iteration-count parity with the reference is a tested property (tests/).
ITERS_FILE = ROOT / "profiles" / "bench_iters.json"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index),
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except ValueError:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------- reference
def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def iters_full(m: int, nev: int, fallback: int) -> tuple[int, str]:
    try:
        d = json.loads(ITERS_FILE.read_text())
        return int(d[f"m{m}_nev{nev}"]["num_iter"]), f"{ITERS_FILE.name} (measured on B200 by this bench)"
    except Exception:
        return fallback, "fallback (no recorded B200 run for this size)"


def reference_sample(m_full: int, nev: int, m_ref: int, it_ref: int, threads: int) -> dict:
    """One bounded sample of the reference's own GCG (oracle/_ref, unmodified sources):
    P1-FEM pencil at m_ref^3, same nev / nevMax / block_size / tolerances, it_ref outer
    iterations from srand(0) (InitializeX included)."""
    from gcge_b200 import problems as P
    from oracle import ref
    ref.set_threads(threads)
    pen = P.p1_fem_kuhn(m_ref)
    r = ref.gcg_solve(pen.A, pen.B, nev=nev, max_iter=it_ref, want_evec=False)
    return {"seconds": r["seconds"], "num_iter": r["num_iter"], "n": pen.A.ncols, "m": m_ref}


def scale_reference(sample: dict, m_full: int, nev: int) -> tuple[float, str]:
    n_full = m_full ** 3
    itf, src = iters_full(m_full, nev, fallback=100)
    per_row_iter = sample["seconds"] / sample["n"] / max(sample["num_iter"], 1)
    value = per_row_iter * n_full * itf
    desc = (f"reference GCG (oracle/_ref, CCS+OpenMP, unmodified) solving p1_fem_kuhn m={sample['m']} (n={sample['n']}), "
            f"nev={nev}, to convergence: {sample['num_iter']} outer iterations, {sample['seconds']:.2f} s incl. "
            f"InitializeX; scaled linearly x(n {n_full}/{sample['n']}) x(outer iterations {itf}/{sample['num_iter']}; "
            f"{itf} = full solve, {src}).  The sample fits the host caches, so this flatters the CPU.")
    return value, desc


def run_reference(a) -> int:
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    from oracle import ref
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libgcge_ref.so was not built "
                          "(needs /root/reference at build time)"}))
        return 0
    cores = host_cores()
    vals, last = [], None
    for i in range(a.warmup + a.steps):
        s = reference_sample(a.m, a.nev, a.ref_m, a.ref_iters, cores)
        v, desc = scale_reference(s, a.m, a.nev)
        if i >= a.warmup:
            vals.append(v); last = desc
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": "gcg_solve_seconds", "value": value, "unit": "s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": value * 1e3, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(a, 1),
            "cpu_baseline": {"value": value, "unit": "s", "cores": cores, "kind": "reference", "sample": last},
            "e2e": {"value": value, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def workload_config(a, n_gpus: int) -> dict:
    return {"workload": f"3D P1-FEM stiffness/mass pencil A x = lambda B x (Kuhn triangulation), n = {a.m}^3 = "
                        f"{a.m ** 3}, nev = {a.nev} (nevMax {2 * a.nev}, block_size {a.nev // 5 if a.nev >= 30 else a.nev}), "
                        f"B-orthogonal GCG with BlockPCG, tol = (1e-1, 1e-8), srand(0)",
            "generator": "gcge_b200.problems.p1_fem_kuhn", "m": a.m, "nev": a.nev,
            "parallelism": "single GPU" if n_gpus == 1 else f"1-D row blocks over {n_gpus} GPUs (NCCL halo exchange + allreduce, replicated Rayleigh-Ritz)",
            "l2": "inputs >> L2 (matrices 2.9 GB, multi-vectors 64 GB at m=200); no flush needed"}


def residual_check(api, A, B, evec, ev, npairs, prm, n, width=40) -> dict:
    """Residual norms of the first npairs returned eigenpairs, width columns at a time."""
    ta, tb = api.MultiVec(n, width), api.MultiVec(n, width)
    res = np.zeros(npairs)
    bnorm = np.zeros(npairs)
    one = np.array([1.0])
    try:
        for c0 in range(0, npairs, width):
            k = min(width, npairs - c0)
            api.mat_dot_multivec(A, evec, ta, (c0, 0), (c0 + k, k))
            api.mat_dot_multivec(B, evec, tb, (c0, 0), (c0 + k, k))
            d = np.zeros(k)
            api.multivec_inner_prod("D", evec, tb, (c0, 0), (c0 + k, k), d, 1)          # x^T B x
            bnorm[c0:c0 + k] = d
            coef = np.asfortranarray(np.diag(-np.asarray(ev[c0:c0 + k], dtype=np.float64)))
            api.multivec_linear_comb(tb, ta, (0, 0), (k, k), coef, k, one, 0)           # A x - lambda B x
            api.multivec_inner_prod("D", ta, ta, (0, 0), (k, k), d, 1)
            res[c0:c0 + k] = np.sqrt(d)
    finally:
        ta.close(); tb.close()
    lam = np.abs(np.asarray(ev[:npairs], dtype=np.float64))
    tol0, tol1 = float(prm.tol[0]), float(prm.tol[1])
    ok = (res <= 1.001 * tol0) & ((res <= 1.001 * lam * tol1) | (lam <= tol1))
    return {"pairs_checked": int(npairs), "pairs_meeting_reference_tolerance": int(ok.sum()),
            "max_residual": float(res.max()), "max_residual_over_lambda": float((res / np.maximum(lam, 1e-300)).max()),
            "max_abs_xBx_minus_1": float(np.abs(bnorm - 1.0).max()), "tol": [tol0, tol1]}


# ------------------------------------------------------------------------------------- ours
def run_b200(a) -> int:
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from gcge_b200 import api, problems as P
    api.init(local)
    if world > 1:
        api.comm_init_from_torch()

    t0 = time.time()
    pen = P.p1_fem_kuhn(a.m)
    n, nnz = pen.A.ncols, pen.A.nnz
    host_arrays = [pen.A.j_col, pen.A.i_row, pen.A.data, pen.B.j_col, pen.B.i_row, pen.B.data]
    for h in host_arrays:
        api.host_register(h)
    t_gen = time.time() - t0
    A, B = api.Mat(pen.A), api.Mat(pen.B)
    prm = api.default_params(a.nev)
    evec = api.MultiVec(n, prm.nevMax)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        api.sync()

    def solve(Am, Bm, ws=None):
        return api.gcg_solve(Am, Bm, nev=a.nev, evec=evec, seed=0, ws=ws, numIterMax=a.max_iter if a.max_iter > 0 else 500)

    # the four GCG workspaces are created before the clock starts, as in the reference's driver
    # (test/test_eig_sol_gcg.c:57-68 allocates gcg_mv_ws[0..3] ahead of `time_start`, :88)
    ws = api.gcg_workspace(n, prm)
    for _ in range(a.warmup):
        out = solve(A, B, ws)
    # ---- timed region: K solves, matrices resident ---------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    api.prof_enable(True)
    l0 = api.kernel_launches()
    barrier()
    api.timer_start()
    w0 = time.time()
    for _ in range(a.steps):
        out = solve(A, B, ws)
    ms = api.timer_stop()
    barrier()
    wall = time.time() - w0
    launches = api.kernel_launches() - l0
    prof = api.prof_report()
    api.prof_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    sec = ms / 1e3 / a.steps
    if dist is not None:
        t = torch.tensor([sec], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    stats = out["stats"]

    # ---- parity at full size, outside the timed region: every returned pair against the reference's
    # own acceptance test (src/ops_eig_sol_gcg.c:229-252): ||A x - lambda B x||_2 <= tol[0] and
    # <= |lambda| tol[1], recomputed from scratch with the library's SpMM / LinearComb / dot kernels
    parity = None
    try:
        parity = residual_check(api, A, B, evec, out["eval"], min(a.nev, int(out["nev_conv"])), prm, n)
    except Exception as exc:                                  # never lose the bench line over the check
        parity = {"error": str(exc)[:200]}

    # ---- e2e: host CCS arrays -> upload -> workspaces -> solve -> eigenpairs back on the host
    A.close(); B.close()
    for w in ws:
        w.close()
    nev_out = a.nev
    _, nloc = evec.local_range()
    host_vec = np.zeros((nloc, nev_out), order="F")          # this rank's rows of the eigenvectors
    api.host_register(host_vec)
    e2e_s = []
    for i in range(a.e2e_steps):
        barrier()
        w0 = time.time()
        A2, B2 = api.Mat(pen.A), api.Mat(pen.B)
        o2 = solve(A2, B2)
        evec.numpy_local(0, nev_out, out=host_vec)
        ev_host = o2["eval"][:nev_out].copy()
        barrier()
        e2e_s.append(time.time() - w0)
        A2.close(); B2.close()
    e2e = float(np.mean(e2e_s)) if e2e_s else None
    if dist is not None and e2e is not None:
        t = torch.tensor([e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = float(t.item())
    h2d = sum(h.nbytes for h in host_arrays) * world      # every rank receives the whole CCS and cuts out its slab on the device
    d2h = 8 * n * nev_out + 8 * prm.nevMax            # all ranks together

    if rank != 0:
        api.comm_finalize()
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel class ---------------------------------------------
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback"
    tot_ms = sum(v["ms"] for v in prof.values()) or 1.0
    classes = {k: {"ms": round(v["ms"] / a.steps, 3), "calls": v["calls"] // a.steps,
                   "share": round(v["ms"] / tot_ms, 4),
                   "GBs": round(v["bytes"] / v["ms"] / 1e6, 1) if v["ms"] > 0 else None,
                   "TFs": round(v["flops"] / v["ms"] / 1e9, 2) if v["ms"] > 0 else None,
                   "gap_before_ms": round(v.get("gap_before_ms", 0.0) / a.steps, 3)}
               for k, v in prof.items()}
    top = max(prof, key=lambda k: prof[k]["ms"])
    tv = prof[top]
    if top in ("gram", "lincomb"):
        dmma_peak = api.measure_dmma_peak()
        roof = {"kernel": top, "bound": "tensor", "achieved": tv["flops"] / tv["ms"] / 1e9, "peak": dmma_peak,
                "unit": "TFLOP/s", "peak_source": "b200_measure_dmma_peak (FP64 DMMA m8n8k4 issue ceiling, measured in this run)"}
    else:
        roof = {"kernel": top, "bound": "hbm", "achieved": tv["bytes"] / tv["ms"] / 1e6, "peak": hbm_peak, "unit": "GB/s",
                "peak_source": hbm_src}
    roof["frac"] = roof["achieved"] / roof["peak"]
    # DRAM traffic per launch of the class's main kernel from the committed ncu --set full capture
    # (profiles/ncu_traffic.json), only when this run is the workload the capture was taken on
    roof["traffic"] = None
    try:
        tr = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())
        if tr.get("workload", {}).get("m") == a.m and world == 1 and top in tr:
            roof["traffic"] = tr[top]["traffic_bytes_per_launch"]
            roof["traffic_source"] = f'{tr[top]["kernel"]}: {tr[top]["source"]}'
    except Exception:
        pass
    roof["launches"] = tv["calls"] // a.steps
    roof["avg_launch_ms"] = tv["ms"] / max(tv["calls"], 1)
    roof["algorithmic_per_launch"] = (tv["flops"] if roof["bound"] == "tensor" else tv["bytes"]) / max(tv["calls"], 1)
    # secondary rooflines: always report SpMM (HBM) and the DMMA contractions
    spmm = prof.get("spmm")
    extra = {}
    if spmm and spmm["ms"] > 0:
        extra["spmm_GBs"] = round(spmm["bytes"] / spmm["ms"] / 1e6, 1)
        extra["spmm_frac_hbm"] = round(spmm["bytes"] / spmm["ms"] / 1e6 / hbm_peak, 4)

    # ---- CPU baseline: the reference itself on the host cores, bounded sample -----------------
    cpu = None
    if world == 1 and not a.no_cpu:
        try:
            from oracle import ref
            if ref.available():
                cores = host_cores()
                s = reference_sample(a.m, a.nev, a.ref_m, a.ref_iters, cores)
                try:
                    ITERS_FILE.parent.mkdir(exist_ok=True)
                    d = json.loads(ITERS_FILE.read_text()) if ITERS_FILE.exists() else {}
                    if a.max_iter <= 0:
                        d[f"m{a.m}_nev{a.nev}"] = {"num_iter": int(out["num_iter"]), "nev_conv": int(out["nev_conv"])}
                        ITERS_FILE.write_text(json.dumps(d, indent=1, sort_keys=True) + "\n")
                except Exception:
                    pass
                v, desc = scale_reference(s, a.m, a.nev)
                cpu = {"value": v, "unit": "s", "cores": cores, "kind": "reference", "sample": desc}
            else:
                cpu = {"value": None, "unit": "s", "cores": 0, "kind": "reference",
                       "sample": "oracle/_ref not present on this box"}
        except Exception as e:  # the baseline must never take the GPU number down with it
            cpu = {"value": None, "unit": "s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}

    line = {"metric": "gcg_solve_seconds", "value": sec, "unit": "s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(a, world),
            "parity_at_full_size": parity,
            "result": {"num_iter": int(out["num_iter"]), "nev_conv": int(out["nev_conv"]),
                       "eval_first": float(out["eval"][0]), "eval_nev": float(out["eval"][a.nev - 1]),
                       "wall_s_per_step": wall / a.steps},
            "phases_s": {k: round(stats[k], 3) for k in ("initX", "checkconv", "compP", "compRR", "rr_eig", "compRV",
                                                          "compW", "linsol", "compX")},
            "kernel_classes": classes,
            "roofline": roof, **extra,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e, "unit": "s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "setup_s": {"generate_pencil": round(t_gen, 2)}}
    print(json.dumps(line))
    api.comm_finalize()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--m", type=int, default=200, help="lattice size: n = m^3 unknowns (200 -> 8.0 M)")
    ap.add_argument("--nev", type=int, default=200)
    ap.add_argument("--max-iter", type=int, default=0, help="cap on outer iterations (0 = reference default 500)")
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--ref-m", type=int, default=24, help="lattice size of the reference's bounded sample")
    ap.add_argument("--ref-iters", type=int, default=500, help="outer-iteration cap of the reference's bounded sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    a = ap.parse_args()
    if a.impl == "reference":
        return run_reference(a)
    return run_b200(a)


if __name__ == "__main__":
    sys.exit(main())
