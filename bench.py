#!/usr/bin/env python
"""bench.py -- the reference's headline metric on the B200 engine.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W   (the reference's CPU path, oracle port, host cores)

One "step" = one Lanczos iteration (x += H y on the fly + the recurrence sweeps) on config 3 of BASELINE.json:
HubbardOneOrbital 4x4, 8 up 8 down, dim 165 636 900, fp64.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Lanczos s/iteration (16-site Hubbard 4x4, 8 up 8 down, on-the-fly SpMV, fp64)"
UNIT = "s/iteration"
BYTES_PER_ROW_SPMV = 24.0    # read y, read x, write x
BYTES_PER_ROW_ITER = 48.0    # fused Lanczos iteration
KERNELS = {"c3": "k_dblock (two-pass block down sweep) + k_sweep_up_packed (one launch each)", "c3small": "k_dblock + k_sweep_up_packed",
           "c4": "k_sweep_down_lean + k_sweep_up_packed + k_sweep_twospin_tab", "c2": "k_spmv_heis (one launch)"}


def metric_name(workload_name, desc):
    return METRIC if workload_name == "c3" else "Lanczos s/iteration (%s, on-the-fly SpMV, fp64)" % desc


def workload(name):
    from lanczosplusplus_b200 import geometry as geo
    if name == "c3":
        return dict(model=0, nsite=16, nup=8, ndown=8, orbitals=1, hop=geo.square(4, 4, -1.0), U=np.full(16, 4.0),
                    V=np.zeros(16)), "HubbardOneOrbital 4x4 PBC t=-1 U=4, 8 up 8 down, dim 165636900"
    if name == "c3small":   # 12-site stand-in for quick checks (not a bench line)
        return dict(model=0, nsite=12, nup=6, ndown=6, orbitals=1, hop=geo.square(4, 3, -1.0), U=np.full(12, 4.0),
                    V=np.zeros(12)), "HubbardOneOrbital 4x3 PBC t=-1 U=4, 6 up 6 down, dim 853776"
    if name == "c4":
        hop = geo.with_orbitals(geo.square(2, 4, -1.0, periodic_x=False, periodic_y=True), 2, 1.0, 0.5)
        return dict(model=1, nsite=8, nup=6, ndown=6, orbitals=2, hop=hop, U=np.array([4.0, 3.0, -0.8, -0.4]),
                    V=np.zeros(32), D=np.array([0.0])), "FeAsBasedSc 2x4 two orbitals, 6 up 6 down, dim 64128064"
    if name == "c2":
        J = geo.chain(24, 1.0, True)
        return dict(model=2, nsite=24, nup=12, ndown=0, orbitals=1, hop=J, jzz=J), "Heisenberg 24-ring Sz=0, dim 2704156"
    raise SystemExit("unknown workload " + name)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None
        self.skip = 0

    def wait_ready(self, timeout=3.0):
        """nvidia-smi needs a few hundred ms before its first line; a 100 ms timed region could end without a sample.  Wait for
        the first line and drop what was sampled before the load starts."""
        if self.p is None:
            return
        t0 = time.time()
        while time.time() - t0 < timeout:
            try:
                if os.path.getsize(self.f.name) > 0:
                    break
            except OSError:
                break
            time.sleep(0.01)
        try:
            with open(self.f.name) as g:
                self.skip = len(g.read().splitlines())
        except OSError:
            self.skip = 0

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().strip().splitlines()[self.skip:]:
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for nme, v in zip(names, c[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        os.unlink(self.f.name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(workload, world):
    """DRAM bytes (read + write) per SpMV from the committed `ncu --set full` capture of the two sweep kernels
    (profiles/traffic_r02.json, written from profiles/prof_sweeps_r02_raw.csv); None when no capture applies."""
    p = os.path.join(ROOT, "profiles", "traffic_r02.json")
    if workload != "c3" or world != 1 or not os.path.exists(p):
        return None, None
    t = json.load(open(p))
    return float(t["spmv_traffic_bytes"]), t["kernels"]


def cpu_sample(case, budget_rows, faithful, steps, warmup):
    """Bounded-sample timing of the reference's CPU path (oracle port) on the host cores: rows [r0, r0+n) of x += H y
    plus the three PsimagLite sweeps on n elements, scaled to the full dimension."""
    from oracle import oracle as orc
    from lanczosplusplus_b200 import geometry as geo
    orc.build()
    orc.set_num_threads(host_cores())
    m = orc.OracleModel(case["model"], case["nsite"], case["nup"], case["ndown"], case["orbitals"], hop=case.get("hop"),
                        jzz=case.get("jzz"), U=case.get("U"), V=case.get("V"), D=case.get("D"),
                        fast_rank=0 if faithful else 1)
    rows = m.rows()
    n = int(min(rows, budget_rows))
    r0 = (rows - n) // 2
    y = geo.splitmix64_vector(rows, 42)
    x = np.zeros(n)
    ys = y[r0:r0 + n].copy()
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        m.matvec_range(x, y, r0, r0 + n, faithful=faithful)
        t1 = time.perf_counter()
        orc.lanczos_sweeps(x, ys)
        t2 = time.perf_counter()
        if it >= warmup:
            times.append((t1 - t0, t2 - t1))
        ys[:] = y[r0:r0 + n]
        x[:] = 0
    mv = float(np.median([t[0] for t in times])) * rows / n
    sw = float(np.median([t[1] for t in times])) * rows / n
    return dict(s_per_iter=mv + sw, spmv_s=mv, sweeps_s=sw, cores=orc.num_threads(), rows=rows, sample_rows=n)


def compiled_reference_check(case):
    """The reference's own HubbardHelper / FeBasedSc product (oracle/_ref: its model headers compiled against the PsimagLite
    shim) next to the port on a sector of the SAME lattice small enough to run whole (the reference has no row-range entry
    point): rows/s of both.  Returns None when the prebuilt library did not travel with the snapshot."""
    try:
        from oracle import oracle as orc, reference as ref
        from lanczosplusplus_b200 import geometry as geo
        if case["model"] == 2 or ref.build() is None:
            return None
        small = dict(case, ndown=2)
        nt = orc.num_threads()
        ref.set_threads(nt)
        r = ref.ReferenceModel(small["model"], small["nsite"], small["nup"], small["ndown"], small["orbitals"],
                               hop=small.get("hop"), U=small.get("U"), V=small.get("V"), D=small.get("D"))
        o = orc.OracleModel(small["model"], small["nsite"], small["nup"], small["ndown"], small["orbitals"],
                            hop=small.get("hop"), U=small.get("U"), V=small.get("V"), D=small.get("D"), u3_all_pairs=0,
                            fast_rank=0)
        n = r.rows()
        if n > 4_000_000:
            return None
        y = geo.splitmix64_vector(n, 42)
        out = {}
        xs = []
        for name, f in (("reference", lambda x: r.matvec(x, y)), ("port", lambda x: o.matvec(x, y, faithful=True))):
            x = np.zeros(n)
            f(x)
            x[:] = 0
            t0 = time.perf_counter()
            f(x)
            out[name + "_rows_per_s"] = n / (time.perf_counter() - t0)
            xs.append(x)
        out.update(rows=n, cores=nt, sector="%d up %d down of the same lattice" % (small["nup"], small["ndown"]),
                   max_abs_diff=float(np.abs(xs[0] - xs[1]).max()))
        return out
    except Exception as exc:      # the check is informative only
        return {"unavailable": str(exc)[:200]}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def checks_vs_fixtures(workload_name, a, b):
    """alpha/beta of the seeded decomposition against (1) the oracle pin at the named size (tools/make_fullsize_pins.py) and
    (2) the committed 1-GPU coefficients (tests/golden/bench_ab_1gpu.json, written by `bench.py --write-ab-fixture`)."""
    out = {"alpha0": float(a[0]), "beta0": float(b[0]), "n": int(min(20, len(a))),
           "ab_checksum": float(np.sum(a[:20]) + np.sum(b[:20]))}
    pins = os.path.join(ROOT, "tests", "golden", "fullsize_pins.json")
    if os.path.exists(pins):
        p = json.load(open(pins)).get(workload_name)
        if p and "alpha0" in p:
            out["oracle_pin"] = {"alpha0_rel_err": abs(a[0] - p["alpha0"]) / abs(p["alpha0"]),
                                 "beta0_rel_err": abs(b[0] - p["beta0"]) / abs(p["beta0"])}
        elif p and "alpha" in p:
            n = min(len(p["alpha"]), len(a))
            out["oracle_pin"] = {"alpha_rel_err": float(np.max(np.abs(np.array(p["alpha"][:n]) - a[:n]) / np.abs(p["alpha"][:n]))),
                                 "beta_rel_err": float(np.max(np.abs(np.array(p["beta"][:n]) - b[:n]) / np.abs(p["beta"][:n])))}
    fx = os.path.join(ROOT, "tests", "golden", "bench_ab_1gpu.json")
    if os.path.exists(fx):
        f = json.load(open(fx)).get(workload_name)
        if f:
            n = min(len(f["alpha"]), len(a), 20)
            fa, fb = np.array(f["alpha"][:n]), np.array(f["beta"][:n])
            out["vs_1gpu"] = {"n": n, "max_rel_diff": float(max(np.max(np.abs(a[:n] - fa) / np.abs(fa)), np.max(np.abs(b[:n] - fb) / np.abs(fb))))}
    return out


def run_reference(args, case, desc):
    """The reference's CPU implementation of the path (oracle port; oracle/_ref has no row-range entry point) on ALL host cores,
    whatever OMP_NUM_THREADS says (torchrun exports 1).  Every step is a bounded row sample, scaled to the full dimension."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    orc.set_num_threads(host_cores())
    r = cpu_sample(case, args.cpu_rows, True, max(1, args.steps), max(0, args.warmup))
    tuned = cpu_sample(case, args.cpu_rows, False, 3, 1)
    frac = r["sample_rows"] / r["rows"]
    line = {"impl": "reference", "metric": metric_name(args.workload, desc), "value": r["s_per_iter"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["s_per_iter"] * 1e3,
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "rows": r["rows"]},
            "extrapolated": frac < 1.0, "sample_fraction": frac,
            "cpu_baseline": {"value": r["s_per_iter"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                             "sample": "rows [%d..+%d) of x+=Hy (faithful HubbardHelper.h:105-134 port: serial diagonal pass, "
                                       "SparseRow per row, OpenMP over rows) + 3 PsimagLite sweeps, median of %d steps after %d "
                                       "warm-up, scaled by %d/%d" % ((r["rows"] - r["sample_rows"]) // 2, r["sample_rows"],
                                                                     max(1, args.steps), max(0, args.warmup), r["rows"], r["sample_rows"])},
            "cpu_baseline_tuned": {"value": tuned["s_per_iter"], "unit": UNIT, "cores": tuned["cores"], "kind": "port",
                                   "sample": "same rows, tuned port: diagonal computed inline, no per-row allocation, table rank "
                                             "(3 steps after 1 warm-up)"},
            "spmv_gbs": BYTES_PER_ROW_SPMV * r["rows"] / r["spmv_s"] / 1e9,
            "e2e": {"value": r["s_per_iter"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "compiled_reference_check": compiled_reference_check(case)}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--kernel", type=int, default=0)
    ap.add_argument("--cpu-rows", type=int, default=5_000_000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-energy", action="store_true")
    ap.add_argument("--write-ab-fixture", action="store_true", help="1 GPU: store alpha/beta of the seeded run in tests/golden")
    args = ap.parse_args()
    case, desc = workload(args.workload)
    if args.impl == "reference":
        return run_reference(args, case, desc)
    args.warmup = max(args.warmup, 3)

    import torch
    import lanczosplusplus_b200 as lpp
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lpp.build()
    eng = lpp.InternalProductCuda(case["model"], case["nsite"], case["nup"], case["ndown"], case["orbitals"],
                                  hop=case.get("hop"), jzz=case.get("jzz"), U=case.get("U"), V=case.get("V"),
                                  D=case.get("D"), device=local, rank=rank, nranks=world, kernel=args.kernel)
    if world > 1:
        from lanczosplusplus_b200 import distributed as lppdist
        lppdist.attach(eng, dist)   # NCCL id broadcast + CUDA IPC handles of the column shards (peer-memory exchange)
    rows = eng.rows()
    _, nloc = eng.local_rows()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def maxranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    # --- kernel-level: x += H y alone (inputs resident in HBM; vectors 1.3 GB >> 126 MB L2, so no flush needed)
    barrier()
    spmv_ms, spmv_launches = eng.bench_spmv(args.steps, args.warmup)
    spmv_ms = maxranks(spmv_ms)
    # --- the step: one Lanczos iteration, device resident, CUDA events on the engine's stream, max over ranks
    if rank == 0:
        sampler.start()
        sampler.wait_ready()
    barrier()
    iter_ms, launches = eng.bench_lanczos(args.steps, args.warmup)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    iter_ms = maxranks(iter_ms)
    # --- e2e: the reference-facing C-ABI call (lpp_lanczos_decomposition) with a HOST initial vector in pinned
    # memory and host alpha/beta: the H2D copy and the per-step D2H scalar reads are inside the timed region
    init = torch.empty(rows, dtype=torch.float64, pin_memory=True)
    init_np = init.numpy()
    f0, _ = eng.local_rows()
    from lanczosplusplus_b200 import geometry as geo
    init_np[f0:f0 + nloc] = geo.splitmix64_vector(nloc, 1234, offset=f0)
    solver = lpp.LanczosSolver(eng, lpp.ParametersForSolver(steps=args.steps, eps=0.0))
    solver.decomposition(init_np)                       # warm (allocations, tables)
    barrier()
    t0 = time.perf_counter()
    a, b, _ = solver.decomposition(init_np)
    torch.cuda.synchronize()
    e2e_s = maxranks((time.perf_counter() - t0) / len(a))
    energy = None
    if not args.no_energy:
        gs = lpp.LanczosSolver(eng, lpp.ParametersForSolver(steps=300, eps=1e-12))
        energy, _, aa, _ = gs.computeOneState(None, want_vector=False)
    if rank != 0:
        return
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    if args.write_ab_fixture and world == 1:
        fx = os.path.join(ROOT, "tests", "golden", "bench_ab_1gpu.json")
        cur = json.load(open(fx)) if os.path.exists(fx) else {}
        cur[args.workload] = {"alpha": [float(t) for t in a[:20]], "beta": [float(t) for t in b[:20]], "seed": 1234,
                              "how": "bench.py --write-ab-fixture on 1 B200 (lpp_lanczos_decomposition, eps 0)"}
        json.dump(cur, open(fx, "w"), indent=1)
    peak, peak_src = measured_peak()
    spmv_bytes = BYTES_PER_ROW_SPMV * rows / world           # per GPU, per launch of the SpMV
    achieved = spmv_bytes / (spmv_ms * 1e-3) / 1e9
    traffic, traffic_kernels = measured_traffic(args.workload, world)
    line = {"metric": metric_name(args.workload, desc), "value": iter_ms * 1e-3, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": iter_ms, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "rows": rows},
            "run": {"kernel": args.kernel, "l2": "vectors (%.2f GB) %s L2; no flush" % (rows * 8e-9, "exceed" if rows * 8 > 2.5e8 else "FIT in"),
                    "sharding": ("two-layout: up sweep on row shards, down sweep on column shards, re-layout by peer-memory kernels "
                                 "over NVLink (CUDA IPC); NCCL for the scalar all-reduces") if world > 1 else "none"},
            "spmv_ms": spmv_ms, "spmv_gbs": BYTES_PER_ROW_SPMV * rows / (spmv_ms * 1e-3) / 1e9,
            "iter_gbs": BYTES_PER_ROW_ITER * rows / (iter_ms * 1e-3) / 1e9,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "kernel": "x += H y = %s, 24 B/row algorithmic, per GPU; duration = CUDA events around the launches "
                                   "on the engine's stream" % KERNELS.get(args.workload, "engine AUTO kernels"),
                         "traffic_source": "profiles/traffic_r02.json (ncu --set full, dram read+write of both sweeps: r02/prof_dblock_2cta_raw.csv + prof_sweeps_r02_raw.csv)"
                                           if traffic else None,
                         "ncu_kernels": traffic_kernels},
            "gpu_launches": int(launches), "clocks": clocks,
            "e2e": {"value": e2e_s, "unit": UNIT, "h2d_bytes_per_step": nloc * 8.0 / len(a), "d2h_bytes_per_step": 16},
            "energy": energy, "parity": checks_vs_fixtures(args.workload, a, b)}
    if world == 1 and not args.no_cpu:
        r = cpu_sample(case, args.cpu_rows, True, 2, 1)
        line["cpu_baseline"] = {"value": r["s_per_iter"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                "sample": "%d of %d rows of x+=Hy (faithful port) + 3 sweeps, scaled" % (r["sample_rows"], r["rows"])}
    print(json.dumps(line), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
