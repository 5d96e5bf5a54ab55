"""The reference's OWN pytest suite (parament/test/test_numerics.py, test_error.py, test_wrapper.py), run
unmodified against our libparament.so through the unmodified pyparament wrapper (staged by oracle/build_ref.sh
into oracle/_ref/pyparament; nothing of /root/reference is read at run time)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WRAPPER = os.path.join(ROOT, "oracle", "_ref", "pyparament")


def test_reference_test_suite_passes_on_our_library(tmp_path):
    if not os.path.isdir(os.path.join(WRAPPER, "parament", "test")):
        pytest.skip("reference wrapper not staged (run oracle/build_ref.sh where /root/reference exists)")
    env = dict(os.environ, PARAMENT_LIB_DIR=os.path.join(ROOT, "parament_b200", "lib"), PYTHONPATH=WRAPPER)
    # np.float was removed in numpy 1.24; the wrapper (parament.py:271) still calls it for fp32 contexts.
    code = ("import numpy as np; np.float = float; import sys, pytest; "
            f"sys.exit(pytest.main(['-q', '-p', 'no:cacheprovider', r'{os.path.join(WRAPPER, 'parament', 'test')}']))")
    r = subprocess.run([sys.executable, "-c", code], env=env, cwd=str(tmp_path), capture_output=True, text=True)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "passed" in r.stdout and "failed" not in r.stdout


def test_unchanged_wrapper_in_process_matches_mirror(tmp_path):
    """Same inputs through the reference wrapper and through parament_b200.Parament give identical bits."""
    if not os.path.isdir(os.path.join(WRAPPER, "parament")):
        pytest.skip("reference wrapper not staged")
    env = dict(os.environ, PARAMENT_LIB_DIR=os.path.join(ROOT, "parament_b200", "lib"), PYTHONPATH=WRAPPER + os.pathsep + ROOT)
    code = """
import numpy as np; np.float = float
import parament, parament_b200
from workloads import make_workload
w = make_workload("C2", pts=4001)
with parament.Parament() as ctx:
    ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=False, quadrature_mode="simpson")
    a = ctx.equiprop(w.dt, *w.carr)
with parament_b200.Parament() as ctx:
    ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=False, quadrature_mode="simpson")
    b = ctx.equiprop(w.dt, *w.carr)
assert a.dtype == np.complex64 and np.array_equal(a, b), np.abs(a - b).max()
with parament.Parament(precision="fp64") as ctx:
    ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=True, quadrature_mode="simpson")
    a = ctx.equiprop(w.dt, *w.carr)
with parament_b200.Parament("fp64") as ctx:
    ctx.set_hamiltonian(w.H0, *w.H1, use_magnus=True, quadrature_mode="simpson")
    b = ctx.equiprop(w.dt, *w.carr)
assert np.array_equal(a, b)
parament.device_info()
print("ok")
"""
    r = subprocess.run([sys.executable, "-c", code], env=env, cwd=str(tmp_path), capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
