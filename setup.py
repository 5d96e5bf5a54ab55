"""pip-installable build (the reference's src/setup.py:28-77 shells out to nvcc the same way).

    pip install --no-build-isolation .          # builds parament_b200/lib/libparament.so for sm_100a with nvcc
    NVCC=/path/to/nvcc pip install .            # honour a specific compiler (the reference honours NVCC_ARGS, setup.py:43)
    pytest --pyargs parament                    # the reference's CI command (README.md:83-87) on the installed package

Installs two packages: `parament_b200` (the library + its ctypes interface) and `parament`, the reference's package name, as a
thin alias with its own acceptance tests, so that `import parament` code and `pytest --pyargs parament` work after the install.
"""
import os
import subprocess

from setuptools import setup
from setuptools.command.build_py import build_py


class BuildWithCuda(build_py):
    def run(self):
        here = os.path.dirname(os.path.abspath(__file__))
        env = dict(os.environ)
        subprocess.check_call(["make", "-C", os.path.join(here, "parament_b200", "csrc"), "-j4"], env=env)
        super().run()


setup(
    name="parament-b200",
    version="0.2.0",
    description="B200-native Parament_equiprop: drop-in libparament.so (FP64 / TF32 tensor pipes) and its ctypes interface",
    packages=["parament_b200", "parament", "parament.test"],
    package_dir={"parament": "packaging/parament"},
    package_data={"parament_b200": ["lib/libparament.so"]},
    cmdclass={"build_py": BuildWithCuda},
    python_requires=">=3.9",
    install_requires=["numpy"],
)
