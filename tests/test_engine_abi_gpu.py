"""The engine ABI below the host solver: a warm start from a basis of STRUCTURAL columns through mlp_engine_init_state —
the device LU of the basis core, FTRAN and BTRAN against numpy on the reference's own LU test matrix (lu.rs:480-552), and
the reference's two singular matrices (lu.rs:555-609), which must come back as MLP_SINGULAR (the reference panics at
solver.rs:316 / 1301) instead of producing numbers."""
import ctypes as C

import numpy as np
import pytest

import minilp_b200 as mb
from minilp_b200 import _lib
from minilp_b200._lib import InitState, pd, pi64, pu8
from minilp_b200.api import Engine

pytestmark = pytest.mark.gpu


def _p(a, t=pd):
    return a.ctypes.data_as(t)


def warm_start(a, basic, pse):
    """Engine over the dense matrix `a` with the given basic variables (all structural here); the non-basic ones sit at 0.
    Returns (status, engine wrapper or None, keep-alive list)."""
    m, n = a.shape
    L = _lib.lib()
    h = C.c_void_p()
    assert L.mlp_engine_create_dense(0, m, n, C.byref(h)) == 0
    a = np.ascontiguousarray(a, dtype=np.float64)
    assert L.mlp_engine_upload_rows(h, 0, m, _p(a)) == 0
    nb = np.array([v for v in range(n + m) if v not in set(basic)], dtype=np.int64)
    assert nb.shape[0] == n
    keep = dict(lo=np.zeros(n + m), hi=np.full(n + m, np.inf), obj=np.zeros(n + m), rhs=np.zeros(m), nb=nb, nbv=np.zeros(n),
                nbd=np.ones(n), nbs=np.full(n, _lib.AT_MIN, dtype=np.uint8), bv=np.array(basic, dtype=np.int64), bx=np.zeros(m),
                blo=np.zeros(m), bhi=np.full(m, np.inf))
    st = InitState(_p(keep["lo"]), _p(keep["hi"]), _p(keep["obj"]), _p(keep["rhs"]), _p(nb, pi64), _p(keep["nbv"]),
                   _p(keep["nbd"]), _p(keep["nbs"], pu8), None, _p(keep["bv"], pi64), _p(keep["bx"]), _p(keep["blo"]),
                   _p(keep["bhi"]), None, int(pse), 1)
    rc = L.mlp_engine_init_state(h, C.byref(st))
    if rc != 0:
        msg = L.mlp_last_error().decode()
        L.mlp_engine_destroy(h)
        return rc, msg
    return 0, (Engine(h, m, n), h)


LU_SIMPLE = np.array([[2.0, 2.0, 123.0, 0.0], [0.0, 0.0, 456.0, 1.0], [3.0, 4.0, 789.0, 1.0]])  # lu.rs:482-486


@pytest.mark.parametrize("pse", [0, 1])
def test_warm_start_from_structural_basis_solves_like_numpy(pse):
    a, basic = LU_SIMPLE, [1, 0, 3]  # the basis columns of lu.rs:487
    rc, (e, h) = warm_start(a, basic, pse)
    assert rc == 0
    m, n = a.shape
    full = np.hstack([a, np.eye(m)])
    B = full[:, basic]
    assert np.allclose(np.linalg.solve(B, [6.0, 3.0, 13.0]), [1.0, 2.0, 3.0])  # lu.rs:527-531, the same system
    for var in (2, 4, 5, 6):  # the non-basic column of A and the three slack columns
        e.ftran_col(var)
        assert np.allclose(e.download(5), np.linalg.solve(B, full[:, var]), rtol=1e-12, atol=1e-12), var
    nbmask = np.ones(n + m, dtype=bool)
    nbmask[basic] = False
    for r in range(m):
        e.calc_row_coeffs(r)
        rho = np.linalg.solve(B.T, np.eye(m)[r])
        assert np.allclose(e.download(6), rho, rtol=1e-12, atol=1e-12)
        assert np.allclose(e.download(7), np.where(nbmask, full.T @ rho, 0.0), rtol=1e-12, atol=1e-10)
    assert e.counters()["k_structural"] == 3
    _lib.lib().mlp_engine_destroy(h)


@pytest.mark.parametrize("rows", [
    [[1.0, 0.0, 0.0], [1.0, 2.0, 3.0], [0.0, 0.0, 0.0]],   # lu.rs:555-609, first matrix: structurally singular (an empty row)
    [[1.0, 0.0, 0.0], [1.0, 2.0, 3.0], [2.0, 2.0, 3.0]],   # second matrix: numerically singular (row 2 = row 0 + row 1)
])
def test_singular_basis_is_reported_not_factorized(rows):
    rc, msg = warm_start(np.array(rows), [0, 1, 2], 0)
    assert rc == _lib.MLP_SINGULAR, (rc, msg)
    assert "singular" in msg
