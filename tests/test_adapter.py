"""include/InternalProductCuda.h at the reference's own template slot.

oracle/_ref/adapter_check (tests/adapter_check.cpp) is the adapter compiled against the reference's ModelBase, DefaultSymmetry,
InternalProductOnTheFly, InternalProductStored and the four model headers (PsimagLite replaced by oracle/psimag_shim).  It is
built where /root/reference exists and travels to the GPU box as a binary.
"""
import os
import re
import subprocess

import pytest

reference = pytest.importorskip("oracle.reference")


def _exe(lpp):
    lpp.build()
    if reference.build() is None or not os.path.exists(reference.ADAPTER_CHECK):
        pytest.skip("oracle/_ref/adapter_check is not built and /root/reference is absent")
    return reference.ADAPTER_CHECK


def test_adapter_compiles_and_reference_products_agree(lpp):
    """CPU part: the header compiles at Engine's InternalProductTemplate slot; the reference's own two products agree."""
    r = subprocess.run([_exe(lpp), "--no-cuda"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout
    for name in ("HubbardOneBand", "FeAsBasedSc"):
        d = float(re.search(name + r" rows=\d+ otf_vs_stored=(\S+)", r.stdout).group(1))
        assert 0 <= d <= 1e-13


@pytest.mark.gpu
def test_adapter_matches_reference_internal_products(lpp):
    """x += H y through InternalProductCuda equals InternalProductStored / InternalProductOnTheFly on the same model objects."""
    r = subprocess.run([_exe(lpp)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout
    for name in ("HubbardOneBand", "FeAsBasedSc", "Heisenberg", "Tj1Orbital"):
        d = float(re.search(name + r" rows=\d+ otf_vs_stored=\S+ cuda_vs_stored=(\S+)", r.stdout).group(1))
        assert 0 <= d <= 1e-12, (name, d)
        # decomposition() / groundState() / statesBelow() of the adapter (the device-resident Krylov loop, reached through the
        # tag dispatch of integration/engine_cuda.patch) against a host Lanczos through the reference's InternalProductStored
        m = re.search(name + r" krylov steps=\d+ ab_rel_diff=(\S+) energy_host=(\S+) energy_cuda=(\S+) residual=(\S+) (\w+)", r.stdout)
        assert m and m.group(5) == "ok", r.stdout
        assert float(m.group(1)) <= 1e-10 and abs(float(m.group(2)) - float(m.group(3))) <= 1e-9 * max(1.0, abs(float(m.group(2))))
