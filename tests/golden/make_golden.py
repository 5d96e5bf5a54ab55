"""Generate tests/golden/reference_tests.npz: the expected values of the reference's own known-answer tests.

The reference's test-suite has no stored vectors; every expected value is a scipy expression evaluated at
test time (test_numerics.py:42,63,82,110-116) plus one printed vector in a docstring (parament.py:61-68).
This script evaluates exactly those expressions on the seeded inputs of cases.py and stores them, so that the
oracle and the CUDA library are checked against frozen numbers.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import scipy.linalg

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from cases import reference_test_cases  # noqa: E402


def expected(case):
    name = case["name"]
    H0 = case["H0"]
    if "expm_scipy_random" in name:
        return scipy.linalg.expm(-1j * case["dt"] * H0)                      # test_numerics.py:42,82
    if "expm_debug" in name:
        return scipy.linalg.expm(case["H1"][0])                              # test_numerics.py:63 (m = H1, H0 = i m)
    if "multi_fields" in name:
        ref = np.eye(H0.shape[0], dtype=np.complex128)
        for i in range(case["carr"].shape[1]):                               # test_numerics.py:111-116
            X = H0 + case["carr"][0, i] * case["H1"][0] + case["carr"][1, i] * case["H1"][1]
            ref = scipy.linalg.expm(-1j * X * case["dt"]) @ ref
        return ref
    if name == "docstring_kat":                                              # parament.py:61-68, printed digits
        return np.array([[0.54030234 - 0.84147096j, 0], [0, 0.54030234 + 0.84147096j]])
    raise KeyError(name)


if __name__ == "__main__":
    store = {c["name"]: expected(c) for c in reference_test_cases()}
    np.savez_compressed(os.path.join(HERE, "reference_tests.npz"), **store)
    print("wrote", len(store), "vectors")
