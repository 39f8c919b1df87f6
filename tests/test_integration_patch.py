"""integration/engine_cuda.patch: the reference-side change that routes Engine's two solver call sites (Engine.h:472-478,
Engine.h:609-626) to the device-resident Krylov loop when the product is InternalProductCuda, selects the product from
`SolverOptions=` (LanczosDriver1.h:222-238) and registers the token (InputCheck.h:157-162).

The executables of the reference cannot be compiled here (PsimagLite is absent), so the test checks that the patch applies
cleanly to the files as they lie in /root/reference, and that what it calls exists in include/InternalProductCuda.h with the
signatures used; tests/adapter_check.cpp compiles and RUNS the same tag dispatch on the GPU (tests/test_adapter.py)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATCH = os.path.join(ROOT, "integration", "engine_cuda.patch")
REF = "/root/reference"


def test_patch_names_what_the_adapter_provides():
    patch = open(PATCH).read()
    header = open(os.path.join(ROOT, "include", "InternalProductCuda.h")).read()
    for name in ("KrylovPlacement", "KrylovOnHostTag", "KrylovOnDeviceTag", "CudaTopology", "setFromGpusLabel"):
        assert name in patch and name in header, name
    # the member functions the patched Engine calls on the product
    assert re.search(r"matrix\.decomposition\(modifVector, ab, params\.steps, params\.tolerance, params\.minSteps\)", patch)
    assert re.search(r"void decomposition\(const VectorType& init,\s+TridiagonalMatrixType& ab,\s+SizeType steps,\s+RealType eps,\s+SizeType minSteps\) const", header)
    assert re.search(r"hamiltonian\.statesBelow\(eigs, zs, initial, excitedPlusOne, params\.steps, params\.tolerance, params\.minSteps\)", patch)
    assert re.search(r"void statesBelow\(VectorRealType& eigs,\s+VectorVectorType& zs,\s+const VectorType& init,\s+SizeType excitedPlusOne,", header)
    assert 'registerOpts.push_back("InternalProductCuda");' in patch


@pytest.mark.skipif(not os.path.isdir(REF) or shutil.which("patch") is None, reason="needs /root/reference and patch(1)")
def test_patch_applies_to_the_reference(tmp_path):
    for f in ("Engine.h", "LanczosDriver1.h", "InputCheck.h"):
        dst = tmp_path / "src" / "Engine"
        dst.mkdir(parents=True, exist_ok=True)
        shutil.copy(os.path.join(REF, "src", "Engine", f), dst / f)
    r = subprocess.run(["patch", "-p1", "--dry-run", "-i", PATCH], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run(["patch", "-p1", "-i", PATCH], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    eng = (tmp_path / "src" / "Engine" / "Engine.h").read_text()
    assert "decompositionDispatch(matrix, params, modifVector, ab" in eng and "statesBelowDispatch(hamiltonian" in eng
    drv = (tmp_path / "src" / "Engine" / "LanczosDriver1.h").read_text()
    assert 'tmp.find("InternalProductCuda")' in drv and "LanczosPlusPlus::InternalProductCuda>(model" in drv
