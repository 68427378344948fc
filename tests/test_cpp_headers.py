"""The header-only concept classes (include/saddlepoint_b200/*.hpp) compile against the C ABI
(CPU tier) and reproduce the reference benchmark through them on the GPU (GPU tier)."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = "/tmp/spb_rosenbrock_gpu"
EXE_GN = "/tmp/spb_gn_legacy_gpu"
EXE_WC = "/tmp/spb_wrapper_concept_gpu"


def _compile(spb):
    spb.capi.load()
    lib = os.path.join(ROOT, "saddlepoint_b200", "lib")
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "rosenbrock_main.cpp"), "-L" + lib, "-lsaddlepoint_b200", "-Wl,-rpath," + lib, "-o", EXE]
    subprocess.check_call(cmd)


def _compile_gn(spb):
    spb.capi.load()
    lib = os.path.join(ROOT, "saddlepoint_b200", "lib")
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "gauss_newton_legacy.cpp"), "-L" + lib, "-lsaddlepoint_b200", "-Wl,-rpath," + lib, "-o", EXE_GN]
    subprocess.check_call(cmd)


def _compile_wc(spb):
    spb.capi.load()
    lib = os.path.join(ROOT, "saddlepoint_b200", "lib")
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "wrapper_concept.cpp"), "-L" + lib, "-lsaddlepoint_b200", "-Wl,-rpath," + lib, "-o", EXE_WC]
    subprocess.check_call(cmd)


def test_headers_compile_and_link(spb):
    _compile(spb)
    assert os.path.exists(EXE)
    _compile_gn(spb)
    assert os.path.exists(EXE_GN)
    _compile_wc(spb)
    assert os.path.exists(EXE_WC)


@pytest.mark.gpu
@pytest.mark.parametrize("solver", [0, 1])
def test_linear_solver_concept_overloads_and_failures(spb, solver):
    """EigenSolverWrapper.h:26-78 surface: solve(Vector), solve(Matrix), public solver / A; an asymmetric
    pattern is refused; a Jacobian pattern that drifts inside LMSolver::solve (same nnz!) and a throwing
    traits class surface as C++ exceptions after the C call returned, and the handle stays usable."""
    _compile_wc(spb)
    out = subprocess.run([EXE_WC, str(solver)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "RESULT failures=0" in out.stdout
    assert "LM mode=1 threw=1 what=Jacobian pattern changed inside LMSolver::solve" in out.stdout
    assert "LM mode=2 threw=1 what=traits failure" in out.stdout
    assert "ASYMMETRIC refused=1" in out.stdout


@pytest.mark.gpu
def test_check_traits_header(spb):
    _compile(spb)
    out = subprocess.run([EXE, "200", "1", "check"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    mres = re.search(r"CHECK max=(\S+) discrepancies=(\d+) colors=(\d+) evaluations=(\d+)", out.stdout)
    assert mres, out.stdout[-500:]
    assert float(mres.group(1)) < 10e-7 and int(mres.group(2)) == 0 and int(mres.group(3)) == 2 and int(mres.group(4)) == 5
    assert "WARNING: FE gradient checking, reducing performance!!!" in out.stdout
    assert "Maximum gradient difference" in out.stdout and "RESULT ret=" in out.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("n,solver", [(1000, 0), (1000, 1), (30000, 2)])
def test_legacy_triplet_concept_gauss_newton(spb, n, solver):
    """GNSolver.h:131,169,175 call sequence (analyze(HRows,HCols,true) / factorize(HVals,true) / solve)
    on the Broyden tridiagonal system; the root is checked against an independent Newton solve."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    _compile_gn(spb)
    out = subprocess.run([EXE_GN, str(n), str(solver)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    mres = re.search(r"RESULT iterations=(\d+) energy=(\S+) x0=(\S+) xmid=(\S+) xlast=(\S+)", out.stdout)
    assert mres, out.stdout[-500:]
    x = -np.ones(n)
    for _ in range(50):
        lo = np.concatenate([[0.0], x[:-1]])
        hi = np.concatenate([x[1:], [0.0]])
        f = (3 - 2 * x) * x - lo - 2 * hi + 1
        if f @ f < 1e-28:
            break
        Jm = sp.diags([-np.ones(n - 1), 3 - 4 * x, -2 * np.ones(n - 1)], [-1, 0, 1], format="csc")
        x = x + spla.spsolve(Jm, -f)
    assert float(mres.group(2)) < 1e-20 and int(mres.group(1)) < 20
    for g, want in ((3, x[0]), (4, x[n // 2]), (5, x[-1])):
        assert abs(float(mres.group(g)) - want) <= 1e-10 * abs(want)


@pytest.mark.gpu
@pytest.mark.parametrize("blocks,solver", [(1, 0), (1, 1), (500, 0)])
def test_reference_rosenbrock_main_through_cpp_headers(spb, orc, blocks, solver):
    _compile(spb)
    out = subprocess.run([EXE, str(blocks), str(solver)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    mres = re.search(r"RESULT ret=(\d) currIter=(\d+) factorizations=(\d+) energy=(\S+) lambda=(\S+) x0=(\S+) x1=(\S+)", out.stdout)
    assert mres, out.stdout[-500:]
    ref = orc.lm_solve("rosenbrock", iparam0=blocks, maxIterations=1000)
    assert int(mres.group(1)) == int(ref.ret)
    assert int(mres.group(2)) == ref.currIter and int(mres.group(3)) == ref.factorizations
    assert abs(float(mres.group(6)) - ref.x[0]) <= 1e-10 * abs(ref.x[0]) + 1e-12
    assert abs(float(mres.group(7)) - ref.x[1]) <= 1e-10 * abs(ref.x[1]) + 1e-12
    assert float(mres.group(5)) == ref.currLambda
    assert "Beginning Optimization" in out.stdout
