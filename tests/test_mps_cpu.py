"""Host-side MPS reader (minilp_b200/mps.py, mirror of mps.rs:39-329) against the oracle's restated parser and the
reference's own fixture (mps.rs:437-462, committed as tests/golden/testprob.mps).  No GPU."""
import os

import numpy as np
import pytest

import oracle
from minilp_b200 import mps, synth
from minilp_b200.api import ComparisonOp, OptimizationDirection

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "testprob.mps")


def same_problem(mf, of):
    obj, mins, maxs, rp, ci, va, ops, rhs = of.problem.export()
    rp2, ci2, va2, ops2, rhs2 = mf.problem.to_csr()
    for a, b in ((rp, rp2), (ci, ci2), (va, va2), (ops, ops2), (rhs, rhs2), (obj, np.array(mf.problem.obj_coeffs)),
                 (mins, np.array(mf.problem.var_mins)), (maxs, np.array(mf.problem.var_maxs))):
        assert np.array_equal(a, b)
    assert mf.variables == of.variables and mf.problem_name == of.problem_name


def test_reference_fixture_parses_like_the_reference():
    text = open(GOLD).read()
    mf = mps.MpsFile.parse(text, OptimizationDirection.Minimize)
    assert mf.problem_name == "TESTPROB"
    assert mf.variables == {"XONE": 0, "YTWO": 1, "ZTHREE": 2}
    p = mf.problem
    # mps.rs:437-462: bounds UP XONE 4, LO YTWO -1 / UP YTWO 1, ZTHREE default; LIM1 <= 5, LIM2 >= 10, MYEQN = 7
    assert list(zip(p.var_mins, p.var_maxs)) == [(0.0, 4.0), (-1.0, 1.0), (0.0, float("inf"))]
    assert p.obj_coeffs == [1.0, 4.0, 9.0]
    assert p.constraints == [([(0, 1.0), (1, 1.0)], ComparisonOp.Le, 5.0), ([(0, 1.0), (2, 1.0)], ComparisonOp.Ge, 10.0),
                             ([(1, -1.0), (2, 1.0)], ComparisonOp.Eq, 7.0)]
    same_problem(mf, oracle.MpsFile.parse(text, OptimizationDirection.Minimize))


@pytest.mark.parametrize("gen,args", [(synth.netlib_like, (120, 150, 5.0, 1)), (synth.sparse_pos, (80, 120, 5.0, 2)),
                                      (synth.netlib_like, (700, 500, 7.0, 3))])
def test_generated_mps_parses_like_the_oracle(gen, args):
    text, d = gen(*args)
    same_problem(mps.MpsFile.parse(text, d), oracle.MpsFile.parse(text, d))


HEAD = "NAME T\nROWS\n N COST\n L R1\n G R2\nCOLUMNS\n X COST 1 R1 1\n X R2 1\n Y R1 1\nRHS\n RHS R1 4 R2 1\n"


def test_sections_and_first_vector_rules():
    text = (HEAD + " RHS2 R1 99\nRANGES\n RNG R1 2.5\n RNG2 R2 7\nBOUNDS\n UP BND X -3\n FR BND Y\n UP BND2 Y 1\nENDATA\n")
    mf = mps.MpsFile.parse(text, OptimizationDirection.Minimize)
    p = mf.problem
    assert p.var_mins == [float("-inf"), float("-inf")] and p.var_maxs == [-3.0, float("inf")]  # mps.rs:299, FR
    # R1 (L, rhs 4, range 2.5) -> two rows [1.5, 4]; second RHS / RANGES / BOUNDS vectors ignored (193-198, 223-228, 253-258)
    assert p.constraints == [([(0, 1.0), (1, 1.0)], ComparisonOp.Ge, 1.5), ([(0, 1.0), (1, 1.0)], ComparisonOp.Le, 4.0),
                             ([(0, 1.0)], ComparisonOp.Ge, 1.0)]
    same_problem(mf, oracle.MpsFile.parse(text, OptimizationDirection.Minimize))


@pytest.mark.parametrize("text,msg", [
    ("ROWS\n", "line 1: expected NAME section"),
    ("NAME T\nCOLUMNS\n", "line 2: expected ROWS section"),
    ("NAME T\nROWS\n L R1\nCOLUMNS\n", "objective function name not declared"),
    ("NAME T\nROWS\n N C\n Q R1\n", "unexpected row type Q"),
    ("NAME T\nROWS\n N C\n L R1\n L R1\n", "row R1 already declared"),
    (HEAD.replace(" Y R1 1\n", " Y R9 1\n"), "unknown constraint: R9"),
    (HEAD.replace(" Y R1 1\n", " Y R1 1\n X R1 2\n"), "variable X already declared"),
    (HEAD + " RHS COST 3\n", "setting objective in RHS section is not supported"),
    (HEAD + "BOUNDS\n MI BND X 0\nENDATA\n", "bound type MI is not supported"),
    (HEAD + "BOUNDS\n UP BND Z 0\nENDATA\n", "unknown variable: Z"),
    (HEAD.replace("R1 4", "R1 4x"), "couldn't parse float from string: `4x`"),
    (HEAD, "expected ENDATA section"),
    (HEAD.replace(" Y R1 1\n", " Y R1\n"), "unexpected end of line"),
])
def test_syntax_errors_carry_the_reference_message(text, msg):
    with pytest.raises(mps.MpsError) as ei:
        mps.MpsFile.parse(text, OptimizationDirection.Minimize)
    assert msg in str(ei.value)
    with pytest.raises(oracle.MpsError):
        oracle.MpsFile.parse(text, OptimizationDirection.Minimize)
