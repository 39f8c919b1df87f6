"""Host logic of the block down sweep (lanczosplusplus_b200/csrc/lpp_dblock_kernel.cuh): the multi-pass plan on the CPU.

tests/dblock_plan_check.cu builds the plan for small Hubbard-type bases (tori, open chains, a non-bipartite 3x3 torus, with and
without site potentials) and for the 4x4 half-filled basis of config 3 in both layouts (two CTAs per SM / one), and walks its
tables on the host the way k_dblock does; every hop has to be applied exactly once and the
result has to equal the plain ELL application of D + T_dn (HubbardHelper.h:105-134).  No device call: nvcc is only the compiler
of the header.
"""
import os
import re
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "dblock_plan_check.cu")
EXE = os.path.join(HERE, "_dblock_plan_check")
DEPS = [SRC, os.path.join(HERE, "..", "lanczosplusplus_b200", "csrc", "lpp_dblock_kernel.cuh")]


def test_block_plan_applies_every_hop_once():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in DEPS):
        subprocess.check_call([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", EXE, SRC])
    r = subprocess.run([EXE], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout
    lines = [l for l in r.stdout.splitlines() if "states" in l]
    assert len(lines) == 12
    assert "2 CTA/SM" in lines[7] and "max block 464" in lines[7]      # config 3 fits two CTAs per SM
    assert "1 CTA/SM" in lines[8] and "max block 924" in lines[8]
    assert all("3 passes" in l for l in lines[9:12])                         # the forced three-pass plans
    for l in lines:
        m = re.search(r"hops (\d+) walked (\d+), max diff (\S+) (\w+)", l)
        assert m and m.group(1) == m.group(2) and float(m.group(3)) <= 1e-13 and m.group(4) == "ok", l
