#!/usr/bin/env python
"""bench.py - placements/sec of the per-query placement hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--queries Q]

A step = one pass of the hot path (pp_place: query prep -> prefilter over every edge -> ranking ->
baseball/Brent -> keep/refine -> weights) over one batch of Q synthetic queries on config C2 of
BASELINE.json (16S shape: 800-taxon tree, 1500 columns, 10k reads, GTR+G4).  With N
ranks every rank places its own Q-query shard against the replicated reference GLVs (weak scaling,
no collective on the data path); value = all queries placed by all ranks / max-over-ranks time.

  value     device-resident: query characters already in HBM, results left in HBM (pp_place_device)
  e2e       the public C-ABI call with HOST buffers (pp_place): pinned host chars -> device, results
            device -> host, inside the timed region
  roofline  the dominant kernel's achieved FP64 rate against the DFMA peak measured in the same run
            (MEASURED_PEAKS.json has no FP64 figure), plus its compulsory-HBM-traffic rate against
            MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the CPU oracle port (oracle/) on all host cores over a bounded sample (rank 0, N=1)

--impl reference times the CPU oracle port alone (the reference binary cannot be built here: no
OCaml toolchain) on the same config, each step a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "placements/sec"
WORKLOADS = {
    # name: (taxa, seed, description, pplacer flags, -p?, prefs overrides)
    "c2": (800, 2, "C2: 16S shape, 800-taxon tree (E=1597), 1500 columns, GTR+G4, synthetic reads 250-1500 columns",
           "pplacer defaults (keep_at_most 7, max_strikes 6, max_pitches 40)", False, {}),
    "c4": (10_000, 4, "C4: 10k-taxon synthetic 16S tree (E=19997, mean branch length 0.002), 1500 columns, GTR+G4, synthetic reads 250-1500 columns",
           "pplacer defaults (keep_at_most 7, max_strikes 6, max_pitches 40)", False, {}),
    # robustness check, not a BASELINE config: C4 with C2's branch lengths - the reference itself degenerates
    # here; shows what the prefilter does when most site maxima sit far below 1
    "c4_long": (10_000, 4, "C4 with mean branch length 0.05 (tree length ~1000: the reference's exponent pull stops, DESIGN.md 5b)",
                "pplacer defaults", False, {}),
    # donovan-test ar53 (the reference's own regression set): 3 bundled genomes + synthetic genomes made from reference
    # rows (5 % substitutions), SURVEY.md 8d "Q=256 synthetic genomes".  Needs the git-ignored fixture tests/golden/_ar53.
    "c3": (6968, 3, "C3: donovan-test ar53 GTDB archaeal protein set (WAG + Price-CAT 20, 6968 taxa, E=13933, 8035 columns, 54 GB of GLVs), bundled + synthetic genomes",
           "pplacer defaults (keep_at_most 7, max_strikes 6, max_pitches 40)", False, {}),
    "c5": (800, 2, "C5: C2 shape in posterior-probability mode",
           "-p --keep-at-most 40 (calc_pp, keep_factor 0.001, max_strikes 20)", True, {"keep_at_most": 40}),
}
WORKLOAD = WORKLOADS["c2"][2]


def env_int(k, d):
    try:
        return int(os.environ.get(k, d))
    except ValueError:
        return d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
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
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh), "MEASURED_PEAKS.json"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def build_workload(name: str, Q: int, rank: int):
    from pplacer_b200 import synth
    from pplacer_b200.reference import build_problem
    if name == "c3":
        from pplacer_b200 import configs, models
        if not configs.ar53_available():
            raise SystemExit("bench.py --workload c3 needs tests/golden/_ar53 (tests/golden/make_fixtures.py /root/reference --ar53)")
        import numpy as np
        base = configs.ar53()
        names = [n[0] for n in base.query_names]
        # synthetic genomes: reference rows with 5 % substitutions (full length), seeded per rank
        rng = np.random.default_rng(3000 + rank)
        z = np.load(os.path.join(configs.GOLDEN, "_ar53", "msa.npz"))
        rows = rng.choice(len(z["names"]), size=max(Q - 3, 0), replace=False)
        aln = [(str(z["names"][i]), z["chars"][i].tobytes().decode("ascii")) for i in rows]
        extra = [(n, s) for n, s in synth.perturbed_queries(aln, models.PROT_CODE, 3000 + rank) if n.endswith("_0")]
        return configs.ar53(extra_queries=extra)
    taxa, seed = WORKLOADS[name][0], WORKLOADS[name][1]
    from pplacer_b200.configs import C4_MEAN_BL
    s = synth.simulate_reference(taxa, 1500, seed=seed, alpha=0.5, mean_bl=C4_MEAN_BL if name == "c4" else 0.05)
    queries = synth.simulate_queries(s, Q, seed=seed + 1000 * rank)
    return build_problem(s.tree, s.model, s.ref_align, queries)


def build_c2(Q: int, rank: int):
    return build_workload("c2", Q, rank)


def oracle_rate(prob, n_threads: int, target_s: float, pp: bool = False, **prefs_kw):
    """Times the CPU oracle port on a bounded sample; returns (placements/s, n_sample, seconds)."""
    from oracle import binding as ob
    from pplacer_b200 import reference
    o = ob.Oracle(prob)
    o.precompute()
    if pp:   # prefs.ml:198-203
        prefs_kw = dict(dict(calc_pp=1, keep_at_most=20, keep_factor=0.001, max_strikes=20), **prefs_kw)
    pf = ob.default_prefs(prior_mean=reference.flat_exp_prior_mean(prob.tree), **prefs_kw)
    Q = prob.query_chars.shape[0]
    n0 = min(Q, 2 * n_threads)
    t = time.perf_counter()
    o.place(prob.query_chars[:n0], pf, n_threads=n_threads)
    dt0 = max(time.perf_counter() - t, 1e-6)
    n = int(min(Q, max(n0, (n0 / dt0) * target_s)))
    t = time.perf_counter()
    _, n_out, st, _ = o.place(prob.query_chars[:n], pf, n_threads=n_threads)
    dt = time.perf_counter() - t
    o.close()
    return n / dt, n, dt


def run_reference(args):
    """--impl reference: the reference's CPU algorithm for this path (oracle port; all host threads)."""
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    if rank != 0:
        return
    from oracle import binding as ob
    from pplacer_b200 import reference
    cores = os.cpu_count() or 1
    prob = build_c2(min(args.queries if args.queries > 0 else 10_000, 4096), 0)
    o = ob.Oracle(prob)
    o.precompute()
    pf = ob.default_prefs(prior_mean=reference.flat_exp_prior_mean(prob.tree))
    Q = prob.query_chars.shape[0]
    # calibrate the per-step sample to ~args.ref_step_seconds of CPU work
    n0 = min(Q, 2 * cores)
    t = time.perf_counter(); o.place(prob.query_chars[:n0], pf, n_threads=cores); dt0 = max(time.perf_counter() - t, 1e-6)
    n = int(min(Q, max(n0, n0 / dt0 * args.ref_step_seconds)))
    for _ in range(max(args.warmup - 1, 0)):
        o.place(prob.query_chars[:n], pf, n_threads=cores)
    t = time.perf_counter()
    for k in range(args.steps):
        lo = (k * n) % max(Q - n + 1, 1)
        o.place(prob.query_chars[lo:lo + n], pf, n_threads=cores)
    dt = time.perf_counter() - t
    value = n * args.steps / dt
    sample = "%d of the workload's queries per step (oracle port of the reference algorithm, %d threads over a shared query counter = pplacer -j %d)" % (n, cores, cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "placements/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "queries_per_step": n, "flags": "pplacer defaults"},
        "cpu_baseline": {"value": value, "unit": "placements/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "placements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from pplacer_b200 import _lib, seqio
    from pplacer_b200.placer import Placer

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout at the first communicator; the contract is ONE JSON
        # line on stdout, so stdout points at stderr while the communicator comes up
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    wl = WORKLOADS[args.workload]
    Q = args.queries if args.queries > 0 else {"c5": 2000, "c3": 256}.get(args.workload, 10_000)
    strong = args.scaling == "strong"
    if strong:
        # ONE fixed query set (the same on every rank: seeded), contiguous shards (pplacer_b200/shard.py) - what
        # Multiprocessing's fork-per-job does to a query file (pplacer_run.ml:618-649)
        from pplacer_b200 import shard
        Q_total = args.total_queries if args.total_queries > 0 else Q
        prob = build_workload(args.workload, Q_total, 0)
        q_lo, q_hi = shard.shard_range(Q_total, rank, world)
        all_chars = prob.query_chars
        prob.query_chars = np.ascontiguousarray(all_chars[q_lo:q_hi])
    else:
        prob = build_workload(args.workload, Q, rank)
        Q_total = prob.query_chars.shape[0] * world
        q_lo = rank * prob.query_chars.shape[0]
    Q = prob.query_chars.shape[0]
    p = Placer(prob, device=local)
    t0 = time.perf_counter(); p.precompute(); torch.cuda.synchronize(); precompute_ms = (time.perf_counter() - t0) * 1e3
    prefs = p.prefs(pp=wl[4], **wl[5])
    K = prefs.keep_at_most
    stream = torch.cuda.Stream(device=dev)
    p.set_stream(stream.cuda_stream)
    p.set_prefilter_kernel({"default": 0, "direct": 1, "table": 2, "staged": 4, "dmma": 5, "logtable": 6, "qlane": 7, "pairs": 8}[args.prefilter])
    p.set_three_tax_engine({"default": 0, "block": 1, "warp": 2, "group": 3, "eigen": 4}[args.tt_engine], args.tt_blocks_per_sm)
    n_figs = p.set_figs(cutoff=args.fig_cutoff) if args.fig_cutoff > 0 else 0
    fp64_peak = p.measure_fp64_peak()

    # ---- device-resident arm -------------------------------------------------------------------
    d_chars = torch.from_numpy(prob.query_chars).to(dev)
    d_out = torch.zeros(Q * K * 56, dtype=torch.uint8, device=dev)
    d_n = torch.zeros(Q, dtype=torch.int32, device=dev)
    d_st = torch.zeros(Q, dtype=torch.int32, device=dev)

    def step_device():
        p.place_device(d_chars.data_ptr(), Q, d_out.data_ptr(), d_n.data_ptr(), d_st.data_ptr(), prefs)

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    barrier()
    p.kernel_launches(reset=True); p.phase_times_ms(reset=True); p.three_tax_work(reset=True)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            step_device()
        ev1.record(stream)
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = p.kernel_launches()
    phases = p.phase_times_ms() / args.steps       # prefilter, rank, optimise(+refine), pp
    tt_work = p.three_tax_work() / args.steps      # pendant calls, distal calls, pendant call-sites, distal call-sites
    total_q = sum_over_ranks(float(Q)) * args.steps
    value = total_q / (ms * 1e-3)
    st_host = d_st.cpu().numpy()
    n_host = d_n.cpu().numpy()

    # ---- end-to-end arm: the public call with pinned HOST buffers -------------------------------
    h_chars = torch.from_numpy(prob.query_chars).pin_memory()
    chars_np = h_chars.numpy()
    for _ in range(min(args.warmup, 2)):
        p.place(chars_np, prefs)
    from pplacer_b200 import shard as shard_mod
    gathered = None
    gather_ms = 0.0
    if world > 1:   # warm the gather's communicator and buffers
        shard_mod.gather_results(*p.place(chars_np, prefs)[:3], Q_total, rank, world, device=dev, reuse=True)
    barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            out, n_out, st, _ = p.place(chars_np, prefs)
            if world > 1:
                # the path's one collective: the final gather of the result records on rank 0 (NCCL gather of the
                # pp_result bytes, torch.distributed), inside the end-to-end region
                tg = time.perf_counter()
                gathered = shard_mod.gather_results(out, n_out, st, Q_total, rank, world, device=dev, reuse=True)
                gather_ms += (time.perf_counter() - tg) * 1e3
        ev1.record(stream)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max_over_ranks(max(ev0.elapsed_time(ev1), 0.0))
    e2e_value = total_q / (e2e_ms * 1e-3)
    gather_ms = max_over_ranks(gather_ms) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    assert np.array_equal(n_out, n_host) and np.array_equal(st, st_host), "device and host entry points disagree"
    gather_check = None
    if world > 1 and rank == 0:
        # sharded == single GPU: rank 0 recomputes a sample of the LAST rank's queries on its own GPU and compares the
        # gathered records byte for byte (placements do not depend on the batch they are computed in)
        g_out, g_n, g_st = gathered
        assert g_out.shape[0] == Q_total
        assert np.array_equal(g_out[q_lo:q_lo + Q].tobytes(), out.tobytes()), "rank 0's own shard changed in the gather"
        n_chk = min(256, Q)
        if strong:
            lo = shard_mod.shard_range(Q_total, world - 1, world)[0]
            chk_chars = np.ascontiguousarray(all_chars[lo:lo + n_chk])
        else:
            lo = (world - 1) * Q
            chk_chars = np.ascontiguousarray(build_workload(args.workload, Q, world - 1).query_chars[:n_chk])
        c_out, c_n, c_st, _ = p.place(chk_chars, prefs)
        same = (c_out.tobytes() == g_out[lo:lo + n_chk].tobytes() and np.array_equal(c_n, g_n[lo:lo + n_chk])
                and np.array_equal(c_st, g_st[lo:lo + n_chk]))
        assert same, "gathered records of rank %d differ from a single-GPU recomputation" % (world - 1)
        gather_check = {"collective": "torch.distributed gather over NCCL (pp_result bytes + counts + status) to rank 0, every step",
                        "bytes_per_rank_per_step": int(Q * (K * 56 + 8)), "ms_per_step_max_over_ranks": gather_ms,
                        "verified": "%d gathered records of rank %d byte-equal to a single-GPU recomputation on rank 0" % (n_chk, world - 1)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -------------------------------------------------------
    peaks, peaks_src = measured_peaks()
    inf = seqio.informative(prob.query_chars)
    sum_nq = float(inf.sum())
    R, S, E = p.R, p.S, p.n_nodes - 1
    share = {k: float(v) for k, v in zip(("prefilter", "rank", "optimise", "pp"), phases)}
    pre_flops = 2.0 * (S if p.gcat else R * S) * E * sum_nq   # 2RS flops per (query, edge, informative site); Gcat: one category per site
    pre_ms = share["prefilter"]
    pre_tflops = pre_flops / (pre_ms * 1e-3) / 1e12 if pre_ms > 0 else 0.0
    used = p.prefilter_kernel_used()
    kname = {1: "prefilter_direct_kernel", 2: "prefilter_table_kernel", 4: "prefilter_staged_kernel<MMA=false>", 5: "prefilter_staged_kernel<MMA=true>", 6: "prefilter_logtable_kernel", 7: "prefilter_qlane_kernel", 8: "prefilter_qlane_kernel<NC=36> (pair-of-sites table)"}.get(used, "?")
    # compulsory HBM bytes: one pass over the prefilter copy of the supernodes + query codes + scores
    pre_bytes = (8.0 * (4 if used == 2 else R * S) + 4.0) * p.L * (E + 1) + Q * p.L + 8.0 * Q * (E + 1)
    if args.workload == "c3" and used == 1:
        kname = "prefilter_direct_kernel<S=20, Gcat>"
    roofline_prefilter = {
        "kernel": kname, "bound": "fp64", "achieved": pre_tflops, "peak": fp64_peak,
        "unit": "TFLOP/s", "frac": pre_tflops / fp64_peak if fp64_peak else None,
        "peak_source": "DFMA micro-benchmark in this run (pp_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 figure",
        "algorithmic_flops_per_launch": pre_flops, "ms_per_launch": pre_ms,
        "hbm": {"achieved": pre_bytes / (pre_ms * 1e-3) / 1e9 if pre_ms > 0 else 0.0, "peak": peaks.get("hbm_gbs"),
                "unit": "GB/s", "peak_source": peaks_src, "bytes_per_launch": pre_bytes,
                "note": "compulsory bytes: one pass over the supernodes (or their four-base table) + query codes + scores"},
        "traffic": None,
    }
    props = torch.cuda.get_device_properties(dev)
    sm_mhz = (clocks or {}).get("sm_max_mhz") or getattr(props, "clock_rate", 0) / 1e3 or None
    peak_theory = props.multi_processor_count * 64 * 2 * float(sm_mhz) * 1e6 / 1e12 if sm_mhz else None   # 64 DFMA lanes per SM per clock
    units = E * sum_nq                                   # (query, edge, informative site) units per launch
    if used == 2:
        # The table form does NOT execute 2RS flops per unit (the four-base table turns a unit into a select and a
        # multiply), so it has no place on the FP64 roofline: it is reported in its own unit against its own limiter,
        # instruction issue (ncu, profiles/: issue-active and FP64-pipe fractions of the committed capture).
        roofline_prefilter.update({
            "bound": "issue", "achieved": units / (pre_ms * 1e-3) / 1e9 if pre_ms > 0 else 0.0, "unit": "G units/s",
            "peak": None, "frac": None, "peak_source": None,
            "ncu": {"issue_active": 0.71, "fp64_pipe": 0.06, "thread_inst_per_unit": 25,
                    "source": "profiles/r2v_full_ncu.txt (same kernel, same geometry)"},
            "algorithmic_tflops_equivalent": pre_tflops,
            "note": "table form: one select + one multiply per (query, edge, site) unit; issue-bound, not FP64-bound - no FP64 "
                    "fraction is claimed for it; roofline.prefilter_direct carries the FMA form the FP64 roofline is defined on",
        })
    if used in (7, 8):
        # query-lane log table: one shared-memory load (a broadcast request: <= 6 distinct words) and one DADD per
        # (query, edge, site) unit, no FMAs - its roof is the shared-memory pipe, one warp request per clock per SM
        # a warp-wide 8-byte shared-memory load is two 128-byte wavefronts, and the pipe moves one wavefront per clock
        # per SM: 16 units per clock per SM (ncu r5b / r5o: l1tex throughput 88 % at 0.78 of this figure)
        lsu_peak = props.multi_processor_count * (32.0 if used == 8 else 16.0) * float(sm_mhz) * 1e6 / 1e9 if sm_mhz else None   # pair table: one word per TWO units
        ach = units / (pre_ms * 1e-3) / 1e9 if pre_ms > 0 else 0.0
        roofline_prefilter.update({
            "bound": "shared-memory bandwidth", "achieved": ach, "unit": "G units/s", "peak": lsu_peak,
            "frac": ach / lsu_peak if lsu_peak else None,
            "peak_source": "SMs x 16 units per clock x SM clock: one 128-byte shared-memory wavefront per clock per SM, two wavefronts per warp-wide 8-byte load (one unit per lane); the DADD pipe allows four times that",
            "algorithmic_tflops_equivalent": pre_tflops, "units_per_launch": units,
            "note": "query-lane log-table form (prefilter_qlane.cuh): lane = query, warp = 32 queries x 32 edges, table tile streamed by TMA bulk copies; "
                    "no FP64 fraction is claimed for it; roofline.prefilter_direct carries the FMA form the FP64 roofline is defined on",
        })
        ncols = (18 if used == 8 else 6) if S == 4 else S + 2   # columns per site
        roofline_prefilter["kernel"] = kname if S == 4 else "prefilter_qlane_kernel<NC=%d>" % ncols
        roofline_prefilter["hbm"]["bytes_per_launch"] = 8.0 * ncols * p.L * (E + 1) + Q * p.L + 8.0 * Q * (E + 1)
        roofline_prefilter["hbm"]["achieved"] = roofline_prefilter["hbm"]["bytes_per_launch"] / (pre_ms * 1e-3) / 1e9 if pre_ms > 0 else 0.0
        roofline_prefilter["hbm"]["note"] = "compulsory bytes: one pass over the log table (6 columns DNA, 22 columns proteins) + query codes + scores"
    # the graded prefilter form: K = R*S DFMAs per unit exactly as ten_masked_logdot_c writes them (linear_c.c:465),
    # timed on the same batch outside the step loop (pp_prefilter_time: CUDA events on the launching stream)
    roofline_direct = None
    other_forms = {}
    if not args.no_direct_prefilter and world == 1 and S == 4 and not p.gcat and R == 4:
        try:
            forms = {"direct": (1, "prefilter_direct_kernel"), "staged": (4, "prefilter_staged_kernel<MMA=false> (TMA bulk copies + mbarrier ring, DFMA)"),
                     "dmma": (5, "prefilter_staged_kernel<MMA=true> (TMA bulk copies + mbarrier ring, DMMA m8n8k4)")}
            times = {k: p.prefilter_time_ms(prob.query_chars, kernel=v[0], reps=2) for k, v in forms.items()}
            best = min(times, key=times.get)
            d_ms = times[best]
            d_tf = pre_flops / (d_ms * 1e-3) / 1e12
            other_forms = {k: {"kernel": forms[k][1], "ms_per_launch": t, "frac": pre_flops / (t * 1e-3) / 1e12 / fp64_peak} for k, t in times.items()}
            roofline_direct = {
                "kernel": forms[best][1], "forms": other_forms, "bound": "fp64", "achieved": d_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": d_tf / fp64_peak if fp64_peak else None, "peak_theoretical": peak_theory,
                "frac_of_theoretical": d_tf / peak_theory if peak_theory else None,
                "peak_source": "DFMA micro-benchmark in this run (pp_measure_fp64_peak); peak_theoretical = SMs x 64 x 2 x SM clock",
                "algorithmic_flops_per_launch": pre_flops, "ms_per_launch": d_ms, "units_per_launch": units,
                "hbm": {"achieved": ((8.0 * R * S + 4.0) * p.L * (E + 1) + Q * p.L + 8.0 * Q * (E + 1)) / (d_ms * 1e-3) / 1e9,
                        "peak": peaks.get("hbm_gbs"), "unit": "GB/s", "peak_source": peaks_src},
                "note": "one launch over the step's whole batch, not part of the timed step (the step uses a table form)",
            }
        except Exception as ex:   # e.g. a model without a direct/table split
            roofline_direct = {"error": str(ex)}
    # second stage: unit = one objective evaluation on n_q sites.  A distal move re-evolves the
    # distal and proximal GLVs (2 x 2RS^2 flops per site) and takes log_like3 (4RS); a pendant move is a
    # P-column lookup plus log_like3.  Compulsory bytes: the d/p gathers 2*8*R*S per site of a distal
    # evaluation, the cached (statd*prox*dist) product 8*R*S per site of a pendant evaluation.
    opt_ms = share["optimise"]
    tt_flops = (4.0 * R * S * S + 4.0 * R * S) * tt_work[3] + 4.0 * R * S * tt_work[2]
    tt_bytes = 16.0 * R * S * tt_work[3] + 8.0 * R * S * tt_work[2]
    tt_tflops = tt_flops / (opt_ms * 1e-3) / 1e12 if opt_ms > 0 else 0.0
    engine = p.three_tax_engine_used()
    roofline_tt = {
        "kernel": {2: "tt_warp_kernel", 3: "tt_group_kernel", 4: "tt_eigen_kernel"}.get(engine, "three_tax_kernel"), "bound": "fp64", "achieved": tt_tflops, "peak": fp64_peak,
        "unit": "TFLOP/s", "frac": tt_tflops / fp64_peak if fp64_peak else None,
        "peak_source": "DFMA micro-benchmark in this run (pp_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 figure",
        "algorithmic_flops_per_step": tt_flops, "ms_per_step": opt_ms,
        "objective_evaluations_per_step": {"pendant": float(tt_work[0]), "distal": float(tt_work[1])},
        "evaluations_per_s": (tt_work[0] + tt_work[1]) / (opt_ms * 1e-3) if opt_ms > 0 else 0.0,
        "hbm": {"achieved": tt_bytes / (opt_ms * 1e-3) / 1e9 if opt_ms > 0 else 0.0, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                "peak_source": peaks_src, "bytes_per_step": tt_bytes,
                "note": "algorithmic bytes of the d/p gathers and the cached product; most of them are served by L2"},
        "traffic": None,
        "note": "all second-stage launches of a step (baseball waves + refinement), timed together with their bookkeeping kernels",
    }
    if engine == 1 or tt_work[:2].sum() == 0:
        roofline_tt["note"] += "; work counters exist for the warp engine only"
    # measured DRAM traffic of the same kernels from a committed ncu capture (profiles/traffic.json), C2 only
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tr = json.load(fh).get(args.workload, {})
        if roofline_prefilter["kernel"] in tr:
            roofline_prefilter["traffic"] = tr[roofline_prefilter["kernel"]]["dram_bytes"]
            roofline_prefilter["traffic_note"] = "dram bytes per launch, ncu capture %s (profiles/traffic.json)" % tr[roofline_prefilter["kernel"]].get("capture", "r2f")
        if roofline_tt["kernel"] in tr and Q == 10_000:
            roofline_tt["traffic"] = tr[roofline_tt["kernel"]]["dram_bytes"]
            roofline_tt["traffic_note"] = "dram bytes per step (all second-stage launches), ncu capture %s (profiles/traffic.json)" % tr[roofline_tt["kernel"]].get("capture", "r2f")
    except Exception:
        pass
    if args.workload == "c3" and used == 1:
        # few queries against 17.9 GB of supernodes: the prefilter is bound by HBM - report it on that roof
        hb = roofline_prefilter["hbm"]
        n_rows = (Q + 63) // 64   # the supernodes stream once per 64-query block row (8 queries per warp x 8 warps)
        hb["bytes_per_launch"] = (8.0 * (p.S if p.gcat else R * S) + 4.0) * p.L * (E + 1) * n_rows + Q * p.L + 8.0 * Q * (E + 1)
        hb["achieved"] = hb["bytes_per_launch"] / (pre_ms * 1e-3) / 1e9 if pre_ms > 0 else 0.0
        hb["note"] = "one pass over the supernodes (K=20 planes + exponents) per 64-query block row + query codes + scores"
        roofline_prefilter.update({"bound": "hbm", "achieved": hb["achieved"], "peak": hb["peak"], "unit": "GB/s",
                                   "frac": hb["achieved"] / hb["peak"] if hb["peak"] else None, "peak_source": hb["peak_source"],
                                   "fp64": {"achieved": pre_tflops, "peak": fp64_peak, "unit": "TFLOP/s"}})
    dominant_tt = opt_ms >= pre_ms
    roofline = dict(roofline_tt if dominant_tt else roofline_prefilter)
    roofline["share_of_step_ms"] = share
    roofline["step_ms"] = ms / args.steps
    roofline["other_kernel"] = roofline_prefilter if dominant_tt else roofline_tt
    roofline["peak_theoretical"] = peak_theory
    if roofline_direct is not None:
        roofline["prefilter_direct"] = roofline_direct

    # ---- CPU baseline on a bounded sample ------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        rate, n_s, dt = oracle_rate(prob, cores, args.cpu_seconds, pp=wl[4], **wl[5])
        cpu = {"value": rate, "unit": "placements/s", "cores": cores, "kind": "port",
               "sample": "first %d queries of the workload, %.1f s, oracle port with %d threads" % (n_s, dt, cores)}

    line = {
        "metric": METRIC, "value": value, "unit": "placements/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl[2], "queries_per_gpu_per_step": Q, "queries_per_step_all_gpus": int(Q_total), "flags": wl[3],
                   "sharding": ("one fixed query set sharded over ranks (strong scaling)" if strong else "every rank places its own queries (weak scaling)")
                               + ", reference GLVs replicated, no data-path collective; results gathered on rank 0 inside e2e",
                   "gather": gather_check, "fig_cutoff": args.fig_cutoff, "n_figs": n_figs,
                   "l2": "inputs larger than L2: supernodes %.0f MB + darr/parr %.0f MB per pass" % (
                       8.0 * R * S * p.L * (E + 1) / 1e6, 16.0 * R * S * p.L * (E + 1) / 1e6),
                   "placed_ok_frac": float((st_host == 0).mean()), "mean_kept": float(n_host.mean())},
        "e2e": {"value": e2e_value, "unit": "placements/s", "h2d_bytes_per_step": int(Q * p.L),
                "d2h_bytes_per_step": int(Q * K * 56 + 8 * Q), "ms_per_step": e2e_ms / args.steps, "wall_ms_per_step": wall_ms / args.steps},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "precompute_ms": precompute_ms, "fp64_peak_tflops_measured": fp64_peak,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--queries", type=int, default=0, help="queries per GPU per step (default: C2/C4 10k, C5 2k)")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS), help="BASELINE.json config (the bench line is C2)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU-baseline sample budget")
    ap.add_argument("--ref-step-seconds", type=float, default=6.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fig-cutoff", type=float, default=0.0, help="rank through figs (pplacer --fig-cutoff); 0 = the reference's default, disabled")
    ap.add_argument("--no-direct-prefilter", action="store_true", help="skip the extra launch that times the direct-FMA prefilter form")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = every rank places its own --queries; strong = one fixed set of --total-queries sharded over the ranks")
    ap.add_argument("--total-queries", type=int, default=0, help="--scaling strong: size of the fixed query set (default: --queries)")
    ap.add_argument("--prefilter", default="default", choices=["default", "direct", "table", "staged", "dmma", "logtable", "qlane", "pairs"],
                    help="prefilter kernel: library default, the FMA form, or the four-base table form")
    ap.add_argument("--tt-engine", default="default", choices=["default", "block", "warp", "group", "eigen"],
                    help="second-stage engine: library default (group), block-per-candidate, warp-per-candidate, block + shared-memory w")
    ap.add_argument("--tt-blocks-per-sm", type=int, default=0, help="cap on resident blocks of the warp engine (0 = max)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
