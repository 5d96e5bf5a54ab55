"""`pip install .` gives a working `import parament` (the reference's package name) and `pytest --pyargs parament`
collects / runs its acceptance tests.  The install goes into a temporary --target directory, offline, without build isolation."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def installed(tmp_path_factory):
    target = tmp_path_factory.mktemp("site")
    r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-deps", "--no-build-isolation", "--quiet",
                        "--target", str(target), ROOT], capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    return str(target)


def _run(installed, code_or_args, module=False):
    env = {k: v for k, v in os.environ.items() if k not in ("PYTHONPATH", "PARAMENT_LIB_DIR")}
    env["PYTHONPATH"] = installed
    cmd = [sys.executable] + (["-m"] + code_or_args if module else ["-c", code_or_args])
    return subprocess.run(cmd, env=env, cwd=installed, capture_output=True, text=True, timeout=600)


def test_install_layout(installed):
    assert os.path.exists(os.path.join(installed, "parament_b200", "lib", "libparament.so"))
    assert os.path.exists(os.path.join(installed, "parament", "__init__.py"))
    assert os.path.exists(os.path.join(installed, "parament", "test", "test_acceptance.py"))


def test_import_parament_binds_the_installed_library(installed):
    r = _run(installed, "import parament, parament_b200._lib as l; print(l.library_path()); print(parament.Parament.__module__)")
    assert r.returncode == 0, r.stderr[-2000:]
    assert installed in r.stdout and "parament_b200" in r.stdout


def test_pytest_pyargs_collects(installed):
    r = _run(installed, ["pytest", "--pyargs", "parament", "--collect-only", "-q", "-p", "no:cacheprovider"], module=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "test_acceptance.py" in r.stdout and "test_expm_scipy_random" in r.stdout


@pytest.mark.gpu
def test_pytest_pyargs_passes_on_a_gpu(installed):
    r = _run(installed, ["pytest", "--pyargs", "parament", "-q", "-p", "no:cacheprovider"], module=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "passed" in r.stdout and "failed" not in r.stdout and "skipped" not in r.stdout
