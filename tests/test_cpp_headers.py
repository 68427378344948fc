"""The header-only concept classes (include/saddlepoint_b200/*.hpp) compile against the C ABI
(CPU tier) and reproduce the reference benchmark through them on the GPU (GPU tier)."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = "/tmp/spb_rosenbrock_gpu"


def _compile(spb):
    spb.capi.load()
    lib = os.path.join(ROOT, "saddlepoint_b200", "lib")
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "rosenbrock_main.cpp"), "-L" + lib, "-lsaddlepoint_b200", "-Wl,-rpath," + lib, "-o", EXE]
    subprocess.check_call(cmd)


def test_headers_compile_and_link(spb):
    _compile(spb)
    assert os.path.exists(EXE)


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
