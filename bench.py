#!/usr/bin/env python
"""bench.py — simplex pivots/sec on the 50k x 50k dense LP (BASELINE.json config 3), 1..8 B200.

A "step" is one simplex pivot (one successful Solver::pivot, solver.rs:1023) of the full revised-simplex loop:
pricing scan, FTRAN, Harris ratio test, BTRAN, price-out, x_B / reduced-cost / steepest-edge updates (second FTRAN,
second BTRAN, second price-out) and the eta push or refactorization.  `value` is timed with CUDA events on the engine's
stream, `e2e` with the host clock around the same host-driven loop through the C ABI (every per-pivot host<->device
copy inside).  The matrix A is state, like weights: it is streamed from pinned HOST buffers before the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W]          this repo's engine
  python bench.py --impl reference ...                         the reference's CPU algorithm (oracle port, 1 thread)
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "simplex_pivots_per_sec"
UNIT = "pivots/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--m", "--rows", dest="m", type=int, default=50000)  # under torchrun use --rows / --cols (its parser grabs --m)
    ap.add_argument("--n", "--cols", dest="n", type=int, default=50000)
    ap.add_argument("--kind", type=int, default=0, help="synthetic LP family (0 = dense_pos, see DESIGN.md)")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cpu-baseline-seconds", type=float, default=30.0,
                    help="bounded CPU sample for the cpu_baseline object (0 disables)")
    ap.add_argument("--ref-budget-seconds", type=float, default=150.0, help="time cap of the --impl reference arm")
    ap.add_argument("--workload", default="dense", choices=["dense", "netlib_like", "sparse_pos"],
                    help="dense: BASELINE config 3 / 5 (the default, what the driver runs); netlib_like / sparse_pos: BASELINE "
                         "config 4, a sparse LP that enters as free-format MPS text (use with --rows 100000 --cols 100000)")
    ap.add_argument("--no-extras", dest="extras", action="store_false",
                    help="default run only: skip the short config-5 and config-4 runs reported under `extra`")
    ap.add_argument("--refactor-factor", type=float, default=1.0,
                    help="refactorize when eta nnz >= factor * lu nnz; 1 = the reference's rule (solver.rs:1096-1097)")
    ap.add_argument("--col-nnz", type=float, default=100.0, help="mean entries per column of the sparse workloads (0.1 % of 100k)")
    return ap.parse_args()


def bench_config(a):
    """The `config` object: identical in both arms (the driver compares them)."""
    if a.workload != "dense":
        return {"workload": workload_name(a), "m": a.m, "n": a.n, "col_nnz": a.col_nnz, "seed": a.seed,
                "l2": "flush_not_needed: every pivot reads the whole CSC copy (12 nnz bytes, about the L2 size) between two "
                      "passes over other data; see roofline.traffic"}
    return {"workload": workload_name(a), "m": a.m, "n": a.n, "kind": a.kind, "seed": a.seed,
            "l2": "inputs_exceed_l2 (8*m*n bytes of A are read per pivot, split over the GPUs)"}


def workload_name(a):
    if a.workload != "dense":
        return (f"{a.workload} {a.m}x{a.n}, {a.col_nnz:g} entries per column, seed {a.seed}, through free-format MPS text "
                f"(BASELINE config 4: sparse price-out)")
    which = "config 5: pricing column-sharded across the GPUs" if a.n == 4 * a.m else "config 3: full pivot loop with eta updates"
    fam = {0: "dense_pos", 1: "dense_box", 2: "dense_cover", 3: "dense_mixed"}[a.kind]
    return f"{fam} {a.m}x{a.n} seed {a.seed} (BASELINE {which})"


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        ms = os.environ.get("MLP_BENCH_CLOCK_SAMPLE_MS", "100")  # 0: no sampling (A/B of the sampler's own interference)
        if ms == "0":
            self.p = None
            return
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", ms], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------- CPU arms
def cpu_port_run(a, warmup, steps, budget_s, threads_for_gen):
    """Times the oracle port (single thread: the reference has no threads anywhere) on the same LP.
    Returns (pivots timed, seconds, note)."""
    import oracle
    m = a.m
    note = ""
    try:
        s = oracle.DenseSolver.synth(a.kind, m, a.n, a.seed, threads=threads_for_gen)
    except (MemoryError, RuntimeError) as exc:  # host RAM too small for the dense 8*m*n bytes
        m = max(1000, a.m // 4)
        note = f" (host could not hold {a.m}x{a.n}: {type(exc).__name__}; rows cut to {m})"
        s = oracle.DenseSolver.synth(a.kind, m, a.n, a.seed, threads=threads_for_gen)
    s.set_record_trace(True)  # the oracle's pivots of this very LP are the parity reference of the bench line
    if warmup > 0:
        s.continue_solve(warmup)
    done_p, sec = 0, 0.0
    while done_p < steps and sec < budget_s:
        fin, dt = s.continue_timed(1)
        sec += dt
        done_p += 1
        if fin:
            break
    ties = {"tied_pivots": s.tied_pivots, "near_tie_pivots": s.near_tie_pivots, "tie_events": s.tie_events,
            "sel_near_tie_pivots": s.sel_near_tie_pivots}
    return done_p, sec, m, note, s.trace().copy(), ties


def parity_against(trace_gpu, trace_cpu, ties, m_used, m):
    """Pivot-for-pivot comparison of the engine's trace with the oracle's (reference tie rule) on the same LP: phase, entering
    variable, its position, leaving row, leaving variable must be equal; objective after each pivot within 1e-8 relative."""
    if m_used != m:
        return {"pivots_compared": 0, "note": "oracle ran a smaller LP (host memory)"}
    k = int(min(trace_gpu.shape[0], trace_cpu.shape[0]))
    if k == 0:
        return {"pivots_compared": 0}
    same = np.all(trace_gpu[:k, :5] == trace_cpu[:k, :5], axis=1)
    first = -1 if same.all() else int(np.argmin(same))
    upto = k if first < 0 else first
    og, oc = trace_gpu[:upto, 7], trace_cpu[:upto, 7]
    rel = float(np.max(np.abs(og - oc) / np.maximum(1.0, np.abs(oc)))) if upto else None
    return {"pivots_compared": k, "first_divergence": first, "oracle_tie_events": ties["tie_events"],
            "oracle_tied_pivots": ties["tied_pivots"], "oracle_near_tie_pivots": ties["near_tie_pivots"],
            "oracle_selection_near_ties": ties.get("sel_near_tie_pivots"),
            "obj_rel_diff": rel, "tolerance": 1e-8, "oracle": "oracle/ C++ port, reference tie rule, same LP from the slack basis"}


def sparse_text(a):
    from minilp_b200 import synth
    return getattr(synth, a.workload)(a.m, a.n, a.col_nnz, a.seed)


def cpu_port_run_sparse(a, text, d, warmup, steps, budget_s):
    """The oracle's faithful sparse solver (CSR + CSC with usize indices, sparse LU, hyper-sparse solves) on the same MPS
    text, one thread.  Returns (pivots timed, seconds, trace, ties, setup seconds)."""
    import oracle
    t0 = time.perf_counter()
    ref = oracle.MpsFile.parse(text, d).problem.init_only()
    setup = time.perf_counter() - t0
    if warmup > 0:
        ref.continue_solve(warmup)
    done_p, sec = 0, 0.0
    while done_p < steps and sec < budget_s:
        k = min(25, steps - done_p)
        fin, dt = ref.continue_timed(k)
        sec += dt
        done_p = ref.pivots_done - warmup
        if fin:
            break
    ties = {"tied_pivots": ref.tied_pivots, "near_tie_pivots": ref.near_tie_pivots, "tie_events": ref.tie_events,
            "sel_near_tie_pivots": ref.sel_near_tie_pivots}
    return done_p, sec, ref.trace().copy(), ties, setup


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    wu = a.warmup
    if a.workload != "dense":
        text, d = sparse_text(a)
        piv, sec, _, _, setup_s = cpu_port_run_sparse(a, text, d, wu, a.steps, a.ref_budget_seconds)
        val = piv / sec if sec > 0 else 0.0
        sample = (f"{piv} consecutive pivots after {wu} warm-up pivots of the same LP parsed from the same MPS text, time-capped "
                  f"at {a.ref_budget_seconds:.0f}s; MPS parse + try_new ({setup_s:.1f}s) excluded")
        emit({"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": piv, "warmup": wu,
              "ms_per_step": 1000.0 * sec / max(piv, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
              "dtype": "f64", "data": "synthetic", "config": bench_config(a),
              "run_detail": {"note": "reference = C++ port of minilp's Rust solver (oracle/), no Rust toolchain in the image"},
              "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                               "host_cores_available": cores},
              "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})
        return
    piv, sec, m_used, note, _, _ = cpu_port_run(a, wu, a.steps, a.ref_budget_seconds, cores)
    val = piv / sec if sec > 0 else 0.0
    sample = (f"{piv} consecutive pivots after {wu} warm-up pivot(s) from the slack basis of the same {m_used}x{a.n} LP, "
              f"time-capped at {a.ref_budget_seconds:.0f}s{note}; matrix generation and try_new excluded")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": piv, "warmup": wu,
        "ms_per_step": 1000.0 * sec / max(piv, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": bench_config(a),
        "run_detail": {"m_used": m_used,
                       "note": "reference = C++ port of minilp's Rust solver (oracle/), no Rust toolchain in the image"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                         "host_cores_available": cores},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------------------------- our arm
def build_solver(a, device, rank=0, world=1, comm=None):
    """Stream this shard's column block of A from pinned host buffers (generated on the host cores) into the engine, then
    Solver::try_new.  world == 1: the block is all of A."""
    import torch

    import minilp_b200 as mb
    t0 = time.perf_counter()
    d, obj, mins, maxs, ops, rhs = mb.synth_vectors(a.kind, a.m, a.n, a.seed)
    s = mb.Solver(a.m, a.n, device, rank, world, comm)
    c0, c1 = s.engine.col_begin, s.engine.col_end
    nloc = c1 - c0
    rows_per_chunk = max(1, min(a.m, (256 << 20) // (8 * nloc)))
    bufs = [torch.empty(rows_per_chunk * nloc, dtype=torch.float64).pin_memory().numpy().reshape(rows_per_chunk, nloc)
            for _ in range(2)]
    threads = max(1, (os.cpu_count() or 1) // world)
    gen_s = up_s = 0.0
    for i, r0 in enumerate(range(0, a.m, rows_per_chunk)):
        nr = min(rows_per_chunk, a.m - r0)
        buf = bufs[i % 2]
        t1 = time.perf_counter()
        mb.synth_block(a.kind, a.m, a.n, a.seed, r0, nr, c0, nloc, threads, out=buf)
        t2 = time.perf_counter()
        s.upload_local_rows(r0, buf[:nr])
        t3 = time.perf_counter()
        gen_s += t2 - t1
        up_s += t3 - t2
    obj_int = -obj if d == mb.OptimizationDirection.Maximize else obj
    t4 = time.perf_counter()
    s.init(obj_int, mins, maxs, ops, rhs)
    s.engine.sync()
    t5 = time.perf_counter()
    setup = {"generate_host_s": round(gen_s, 3), "h2d_upload_s": round(up_s, 3), "h2d_upload_bytes": 8 * a.m * nloc,
             "try_new_s": round(t5 - t4, 3), "total_s": round(t5 - t0, 3)}
    return s, setup


def load_traffic(a, nloc):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "price_traffic.json")
    try:
        t = json.load(open(p))
        if t.get("m") == a.m and t.get("n") == nloc:
            return t.get("dram_bytes_per_launch")
    except Exception:
        pass
    return None


class Ctx:
    """One process per GPU (torchrun): NCCL process group for the bench's own barriers / max-over-ranks, and a fresh NCCL
    unique id per engine (an id serves one ncclCommInitRank round)."""

    def __init__(self):
        import torch

        import minilp_b200 as mb
        self.torch, self.mb = torch, mb
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if mb.device_count() < 1:
            raise RuntimeError("bench.py needs a CUDA device: minilp_b200 has no CPU fallback")
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            torch.cuda.set_device(self.local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist

    def fresh_comm(self):
        if self.dist is None:
            return None
        torch = self.torch
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if self.rank == 0:
            idt = torch.tensor(list(self.mb.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        self.dist.broadcast(idt, 0)
        return bytes(idt.cpu().tolist())

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def measure_sparse(a, ctx):
    """BASELINE config 4: MPS text -> native reader -> CSR -> device (CSC built there) -> dual simplex loop.  Returns the JSON
    line on rank 0, None elsewhere."""
    import minilp_b200 as mb
    from minilp_b200 import mps
    world, rank, local, barrier, max_over_ranks = ctx.world, ctx.rank, ctx.local, ctx.barrier, ctx.max_over_ranks
    comm = ctx.fresh_comm()
    t0 = time.perf_counter()
    text, d = sparse_text(a)
    t1 = time.perf_counter()
    p = mps.MpsFile.parse(text, d).problem
    t2 = time.perf_counter()
    rp, ci, va, ops, rhs = p.to_csr()
    m, n, nnz = len(ops), len(p.obj_coeffs), len(va)
    s = mb.Solver(m, n, local, rank, world, comm, csr=(rp, ci, va))
    t3 = time.perf_counter()
    s.init(np.array(p.obj_coeffs), np.array(p.var_mins), np.array(p.var_maxs), ops, rhs)
    s.engine.sync()
    t4 = time.perf_counter()
    setup = {"generate_text_s": round(t1 - t0, 3), "mps_bytes": len(text), "mps_parse_s": round(t2 - t1, 3),
             "create_upload_transpose_s": round(t3 - t2, 3), "try_new_s": round(t4 - t3, 3),
             "ingest_s (parse + upload + try_new)": round(t4 - t1, 3)}
    e = s.engine
    s.set_record_trace(True)
    s.set_refactor_factor(a.refactor_factor)
    if a.warmup > 0:
        s.run(a.warmup)
    c0 = e.counters()
    p0 = s.pivots_done
    e.profile_enable(True)
    sampler = ClockSampler(local) if rank == 0 else None
    time.sleep(0.3)
    e.sync()
    barrier()
    w0 = time.perf_counter()
    e.event_mark(0)
    done = s.run(a.steps)
    e.event_mark(1)
    e.sync()
    barrier()
    w1 = max_over_ranks(time.perf_counter() - w0) + w0
    dev_ms = max_over_ranks(e.event_elapsed_ms(0, 1))
    clocks = sampler.stop() if sampler else None
    e.profile_enable(False)
    prof = e.profile()
    c1 = e.counters()
    steps = s.pivots_done - p0
    _, refac_s = s.timers()
    if rank != 0:
        s.close()
        return None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # the dual loop prices out the tableau row (slot rho); the primal loop (sparse_pos) adds the N^T v product (slot v)
    lau = prof["price_rho_launches"] + prof["price_v_launches"]
    pms = prof["price_rho_ms"] + prof["price_v_ms"]
    pby = prof["price_rho_bytes"] + prof["price_v_bytes"]
    ach = pby / (pms * 1e-3) / 1e9 if pms > 0 else 0.0
    traffic = None
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "price_csc_traffic.json")))
        if t.get("nnz") == nnz:
            traffic = t.get("dram_bytes_per_launch")
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": steps / (dev_ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": a.warmup,
        "ms_per_step": dev_ms / max(steps, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": bench_config(a),
        "run_detail": {"rows": m, "cols": n, "nnz": nnz,
                       "parallelism": (f"price-out and per-variable arrays column-sharded over {world} GPUs ({e.n} columns each), "
                                       f"matrix and basis replicated, one candidate exchange per pivot ({e.exchange_kind()})")
                       if world > 1 else "single GPU", "pivots_before_timed_region": p0, "optimal_reached": bool(done),
                       "k_structural_end": c1["k_structural"], "eta_count_end": c1["eta_count"],
                       "refactor_rule": f"eta nnz >= {a.refactor_factor:g} * lu nnz (1 = the reference's, solver.rs:1096-1097)",
                       "refactors_in_region": c1["refactors"] - c0["refactors"],
                       "of_them_product_form_refreshes": c1["refreshes"] - c0["refreshes"],
                       "refreshes_redone_as_true_factorizations (accuracy probe)": c1["refresh_rejects"] - c0["refresh_rejects"],
                       "refactor_wall_s": refac_s,
                       "refactor_share_of_wall": refac_s / max(w1 - w0, 1e-9), "setup": setup,
                       "objective_after": s.cur_obj_val, "l2": "the CSC copy (12 nnz bytes) is about the size of the 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": steps / (w1 - w0), "unit": UNIT, "h2d_bytes_per_step": (c1["h2d_bytes"] - c0["h2d_bytes"]) / max(steps, 1),
                "d2h_bytes_per_step": (c1["d2h_bytes"] - c0["d2h_bytes"]) / max(steps, 1), "wall_s": w1 - w0},
        "gpu_launches": (c1["kernel_launches"] - c0["kernel_launches"]) * world,
        "roofline": {"bound": "hbm", "kernel": "k_price_csc_seg + k_price_csc_fin: price-out over the CSC copy (calc_row_coeffs "
                                               "solver.rs:685-692; N^T v 1117-1132 in the primal loop)",
                     "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                     "algorithmic_bytes_per_launch": pby / max(lau, 1), "formula": "12 nnz + 8 m + 8 (n + m)",
                     "launches_timed": lau, "avg_launch_ms": pms / max(lau, 1), "share_of_step_time": pms / dev_ms if dev_ms else 0,
                     "note": "12 nnz = 117 MB sits at the edge of the 126 MB L2: a launch that finds the matrix there runs above "
                             "the HBM figure; traffic (ncu dram bytes per launch, profiles/) says how much really came from HBM"},
    }
    if a.cpu_baseline_seconds > 0 and world == 1:
        piv, sec, tr_cpu, ties, setup_s = cpu_port_run_sparse(a, text, d, a.warmup, 100000, a.cpu_baseline_seconds)
        line["parity"] = parity_against(s.trace(), tr_cpu, ties, m, m)
        line["parity"]["engine_ties"] = s.tie_stats()
        line["cpu_baseline"] = {"value": piv / sec if sec > 0 else 0.0, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": (f"pivots {a.warmup + 1}..{a.warmup + piv} of the same LP from the same MPS text ({sec:.1f}s of "
                                           f"single-thread CPU work after {a.warmup} untimed pivots; parse + try_new {setup_s:.1f}s excluded)"),
                                "host_cores_available": os.cpu_count() or 1}
    s.close()
    return line


def measure_dense(a, ctx):
    """One process per GPU.  world > 1 (torchrun): the SAME LP is column-sharded over the ranks (strong scaling); every rank
    runs the identical host control loop, the one exchange step per pivot goes over NVLink inside the engine.  Returns the
    JSON line on rank 0, None elsewhere."""
    world, rank, local, barrier, max_over_ranks = ctx.world, ctx.rank, ctx.local, ctx.barrier, ctx.max_over_ranks
    s, setup = build_solver(a, local, rank, world, ctx.fresh_comm())
    e = s.engine
    nloc = e.n
    s.set_record_trace(True)
    if a.warmup > 0:
        s.run(a.warmup)
    c0 = e.counters()
    p0 = s.pivots_done
    e.profile_enable(True)
    sampler = ClockSampler(local) if rank == 0 else None
    time.sleep(0.3)
    e.sync()
    barrier()
    t0 = time.perf_counter()
    e.event_mark(0)
    done = s.run(a.steps)
    e.event_mark(1)
    e.sync()
    barrier()
    t1 = time.perf_counter()
    dev_ms = max_over_ranks(e.event_elapsed_ms(0, 1))
    wall = max_over_ranks(t1 - t0)
    clocks = sampler.stop() if sampler else None
    e.profile_enable(False)
    prof = e.profile()
    c1 = e.counters()
    steps = s.pivots_done - p0
    if steps != a.steps and rank == 0:
        print(f"# note: optimum reached after {steps} timed pivots (asked for {a.steps})", file=sys.stderr)
    value = steps / (dev_ms / 1e3)
    e2e = steps / wall
    run_s, refac_s = s.timers()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    nv = max(prof["price_v_launches"], 1)
    ach = prof["price_v_bytes"] / nv / (prof["price_v_ms"] / nv * 1e-3) / 1e9 if prof["price_v_ms"] > 0 else 0.0
    iso_ms, iso_bytes = e.bench_price_dense(5)
    barrier()
    if rank != 0:
        s.close()
        return None
    roofline = {
        "bound": "hbm", "kernel": "k_price_partial_tma (bulk-copy ring; chunk partials reduced inside k_update_select): N^T v of "
                                   "update_primal_sq_norms, solver.rs:1117-1132",
        "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": load_traffic(a, nloc),
        "traffic_source": "static: DRAM bytes per launch from the committed ncu --set full capture (profiles/price_traffic.json), "
                          "reported when its shape equals this run's; not measured in this run",
        "peak_source": peak_src, "launches_timed": prof["price_v_launches"],
        "avg_launch_ms": prof["price_v_ms"] / nv, "algorithmic_bytes_per_launch": prof["price_v_bytes"] / nv,
        "share_of_step_time": prof["price_v_ms"] / dev_ms,
        "isolated_dense_GBps": iso_bytes / (iso_ms * 1e-3) / 1e9,
        "price_rho": {"launches": prof["price_rho_launches"], "ms_total": prof["price_rho_ms"],
                      "bytes_total": prof["price_rho_bytes"]},
    }
    if world > 1:
        roofline["note"] = f"per GPU (rank 0): its {a.m} x {nloc} column block"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": a.warmup,
        "ms_per_step": dev_ms / max(steps, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": bench_config(a),
        "run_detail": {"a_bytes_per_gpu_per_pivot": 8 * a.m * nloc,
                   "parallelism": (f"columns sharded over {world} GPUs ({nloc} each), basis replicated, one candidate exchange "
                                   f"per pivot ({e.exchange_kind()})") if world > 1 else "single GPU",
                   "pivots_before_timed_region": p0, "optimal_reached": bool(done),
                   "k_structural_end": c1["k_structural"], "eta_count_end": c1["eta_count"],
                   "refactors_in_region": c1["refactors"] - c0["refactors"], "setup": setup,
                   "objective_after": s.cur_obj_val,
                   "engine_switches": {k: os.environ[k] for k in ("MLP_FUSED", "MLP_FUSED_MAX", "MLP_LANE1_LDG", "MLP_OVERLAP",
                                                                  "MLP_PRICE_TMA", "MLP_P2P") if k in os.environ}},
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": UNIT,
                "h2d_bytes_per_step": (c1["h2d_bytes"] - c0["h2d_bytes"]) / max(steps, 1),
                "d2h_bytes_per_step": (c1["d2h_bytes"] - c0["d2h_bytes"]) / max(steps, 1),
                "wall_s": wall, "refactor_wall_s": refac_s,
                # the same pivots with the one-time upload of A (pinned host -> HBM, before the loop) charged to them
                "value_incl_matrix_upload": steps / (wall + setup["h2d_upload_s"]),
                "matrix_upload_bytes": setup["h2d_upload_bytes"]},
        "gpu_launches": (c1["kernel_launches"] - c0["kernel_launches"]) * world,
        "roofline": roofline,
    }
    if a.cpu_baseline_seconds > 0 and world == 1:
        cores = os.cpu_count() or 1
        # the very first pivot of the port is ~7x slower than the following ones (first touch of its work vectors and of the
        # 20 GB matrix): it is run untimed, as the --impl reference arm does
        piv, sec, m_used, note, tr_cpu, ties = cpu_port_run(a, 1, 1000, a.cpu_baseline_seconds, cores)
        line["parity"] = parity_against(s.trace(), tr_cpu, ties, m_used, a.m)
        line["parity"]["engine_ties"] = s.tie_stats()
        line["cpu_baseline"] = {
            "value": piv / sec if sec > 0 else 0.0, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": (f"pivots 2..{piv + 1} of the same {m_used}x{a.n} LP after one untimed pivot ({sec:.1f}s of single-thread "
                       f"CPU work; the reference is single-threaded){note}"), "host_cores_available": cores}
    s.close()
    return line


def compact_line(line):
    """What an `extra` entry keeps of a full bench line."""
    keep = ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "config", "e2e", "gpu_launches", "parity", "cpu_baseline")
    out = {k: line[k] for k in keep if k in line}
    r = line.get("roofline", {})
    out["roofline"] = {k: r.get(k) for k in ("bound", "achieved", "peak", "unit", "frac", "avg_launch_ms", "algorithmic_bytes_per_launch",
                                             "share_of_step_time", "launches_timed")}
    d = line.get("run_detail", {})
    out["run_detail"] = {k: d[k] for k in ("parallelism", "k_structural_end", "eta_count_end", "refactors_in_region", "of_them_product_form_refreshes", "refactor_share_of_wall",
                                           "setup", "objective_after", "nnz") if k in d}
    return out


def run_ours(a):
    """The driver's line is BASELINE config 3 (or whatever --workload / --rows / --cols / --kind name).  With the default
    workload it also carries `extra`: short runs of the other two benchmark configurations on the same GPUs, so that they are
    measured wherever the headline is — config 5 (dense 50k x 200k, the column-sharding config) and config 4 (netlib_like
    100k x 100k through MPS)."""
    import copy
    ctx = Ctx()
    line = measure_dense(a, ctx) if a.workload == "dense" else measure_sparse(a, ctx)
    default_run = a.workload == "dense" and (a.m, a.n, a.kind, a.seed) == (50000, 50000, 0, 1)
    if a.extras and default_run:
        extra = {}
        c5 = copy.copy(a)
        c5.n, c5.steps, c5.warmup, c5.cpu_baseline_seconds = 200000, min(a.steps, 20), 3, 0.0
        c4 = copy.copy(a)
        c4.workload, c4.m, c4.n, c4.steps, c4.warmup = "netlib_like", 100000, 100000, 2000, 20
        c4.cpu_baseline_seconds = min(a.cpu_baseline_seconds, 8.0)
        for name, args, fn in (("config5_dense_50000x200000", c5, measure_dense), ("config4_netlib_like_100000x100000", c4, measure_sparse)):
            try:
                t0 = time.perf_counter()
                got = fn(args, ctx)
                if got is not None:
                    extra[name] = compact_line(got)
                    extra[name]["bench_seconds"] = round(time.perf_counter() - t0, 1)
            except Exception as exc:  # an extra must never cost the headline line
                extra[name] = {"error": repr(exc)}
        if line is not None:
            line["extra"] = extra
    if line is not None:
        emit(line)
    ctx.close()


def emit(line):
    """The ONE JSON line goes to the process's real stdout; everything else printed while the bench runs (NCCL's version
    banner, library chatter) was redirected to stderr by main()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    a = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
