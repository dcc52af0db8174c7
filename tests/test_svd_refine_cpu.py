"""CPU: the GEMM-based SVD refinement planned for two-site splits of numerically singular blocks (DESIGN section 9,
tools/svd_refine_prototype.py; a prototype, not on the product path) -- from the factors a polar-decomposition driver
returns for such a matrix (the exact SVD of A + E, |E| = 3e-10 |A|), two refinement passes must reach LAPACK's singular
values, a reconstruction error and isometry defects of order 1e-14, for random, rank-deficient, graded and
degenerate spectra."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import svd_refine_prototype as proto   # noqa: E402


@pytest.mark.parametrize("n", [48, 130])
def test_refinement_reaches_working_precision(n):
    rng = np.random.default_rng(n)
    for name, a in proto.test_matrices(n, rng).items():
        u, s, v = proto.polar_like_start(a, 3e-10, rng)
        before = proto.errors(a, u, s, v)
        assert before["reconstruction"] > 1e-11, name            # the start really is what the driver's fallback sees
        u, s, v = proto.refine_svd(a, u, v)
        after = proto.errors(a, u, s, v)
        assert after["sigma"] < 1e-13 and after["reconstruction"] < 1e-13, (name, after)
        assert after["u isometry"] < 1e-13 and after["v isometry"] < 1e-13, (name, after)


def test_cluster_detection():
    s = np.array([1.0, 0.9, 0.9 - 1e-9, 0.5, 1e-12, 1e-14, 0.0])
    assert proto.clusters_of(s, 1e-5) == [(1, 3), (4, 7)]
    assert proto.clusters_of(np.array([3.0, 2.0, 1.0]), 1e-5) == []
