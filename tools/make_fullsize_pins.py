#!/usr/bin/env python
"""tools/make_fullsize_pins.py -- pins of the BASELINE.json configs at their NAMED size, computed by the oracle
(oracle/lanczos_oracle.c, the CPU restatement of HubbardHelper.h:105-134 / FeBasedSc.h:66-105 / Heisenberg.h:80-114).

For config 3 (4x4 Hubbard, dim 165 636 900) and config 4 (FeAs 2x4, dim 64 128 064) one full-size oracle mat-vec gives the first
Lanczos coefficients of the seeded initial vector v0 = splitmix64(seed 1234) / |.|:
    alpha_0 = <v0 | H v0>,   beta_0 = | H v0 - alpha_0 v0 |
and a per-chunk checksum of H v0 (sum and sum of squares over 64 equal row chunks), so the GPU test can localise a wrong
panel.  Config 2 (Heisenberg 24) is small enough for the whole oracle recurrence: alpha_0..alpha_9, beta_0..beta_9.

Writes tests/golden/fullsize_pins.json.  Run on the CPU container: python tools/make_fullsize_pins.py [c2 c3 c4]
(c3 takes a few minutes on 8 cores; memory 2 x 1.3 GB).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lanczosplusplus_b200 import geometry as geo   # noqa: E402
from oracle import oracle as orc                    # noqa: E402
import bench                                        # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "fullsize_pins.json")
NCHUNK = 64


def pin_first_step(name):
    case, desc = bench.workload(name)
    m = orc.OracleModel(case["model"], case["nsite"], case["nup"], case["ndown"], case["orbitals"], hop=case.get("hop"),
                        jzz=case.get("jzz"), U=case.get("U"), V=case.get("V"), D=case.get("D"), fast_rank=1)
    n = m.rows()
    v = geo.splitmix64_vector(n, 1234)
    v /= np.sqrt(np.dot(v, v))
    bounds = [(n * i) // NCHUNK for i in range(NCHUNK + 1)]
    sums, sq = [], []
    alpha = 0.0
    hv_chunks = []
    t0 = time.time()
    for i in range(NCHUNK):
        r0, r1 = bounds[i], bounds[i + 1]
        x = np.zeros(r1 - r0)
        m.matvec_range(x, v, r0, r1, faithful=False)
        alpha += float(np.dot(x, v[r0:r1]))
        sums.append(float(x.sum()))
        sq.append(float(np.dot(x, x)))
        hv_chunks.append(x)
        print("%s chunk %d/%d  %.0f s" % (name, i + 1, NCHUNK, time.time() - t0), flush=True)
    b2 = 0.0
    for i in range(NCHUNK):
        r0, r1 = bounds[i], bounds[i + 1]
        w = hv_chunks[i] - alpha * v[r0:r1]
        b2 += float(np.dot(w, w))
    return {"workload": desc, "rows": int(n), "seed": 1234, "alpha0": alpha, "beta0": float(np.sqrt(b2)), "chunks": NCHUNK,
            "chunk_sum": sums, "chunk_sumsq": sq, "oracle": "oracle/lanczos_oracle.c orc_matvec_range, fast_rank=1, faithful=0"}


def pin_c2():
    case, desc = bench.workload("c2")
    m = orc.OracleModel(case["model"], case["nsite"], case["nup"], case["ndown"], case["orbitals"], hop=case.get("hop"),
                        jzz=case.get("jzz"), fast_rank=1)
    n = m.rows()
    v = geo.splitmix64_vector(n, 1234)
    a, b = m.decomposition(v, steps=10, eps=0.0)
    return {"workload": desc, "rows": int(n), "seed": 1234, "alpha": [float(t) for t in a], "beta": [float(t) for t in b],
            "oracle": "oracle/lanczos_oracle.c orc_lanczos_decomposition, 10 steps, eps 0"}


def main():
    which = sys.argv[1:] or ["c2", "c3", "c4"]
    pins = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for w in which:
        pins[w] = pin_c2() if w == "c2" else pin_first_step(w)
        json.dump(pins, open(OUT, "w"), indent=1)
        print("wrote", w, {k: pins[w][k] for k in pins[w] if k in ("alpha0", "beta0", "rows")}, flush=True)


if __name__ == "__main__":
    main()
