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


PARSERS = pytest.mark.parametrize("parse", [mps.MpsFile.parse, mps.MpsFile.parse_python], ids=["native", "python"])


@PARSERS
def test_reference_fixture_parses_like_the_reference(parse):
    text = open(GOLD).read()
    mf = parse(text, OptimizationDirection.Minimize)
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
@PARSERS
def test_generated_mps_parses_like_the_oracle(gen, args, parse):
    text, d = gen(*args)
    same_problem(parse(text, d), oracle.MpsFile.parse(text, d))


HEAD = "NAME T\nROWS\n N COST\n L R1\n G R2\nCOLUMNS\n X COST 1 R1 1\n X R2 1\n Y R1 1\nRHS\n RHS R1 4 R2 1\n"


@PARSERS
def test_sections_and_first_vector_rules(parse):
    text = (HEAD + " RHS2 R1 99\nRANGES\n RNG R1 2.5\n RNG2 R2 7\nBOUNDS\n UP BND X -3\n FR BND Y\n UP BND2 Y 1\nENDATA\n")
    mf = parse(text, OptimizationDirection.Minimize)
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
@PARSERS
def test_syntax_errors_carry_the_reference_message(text, msg, parse):
    with pytest.raises(mps.MpsError) as ei:
        parse(text, OptimizationDirection.Minimize)
    assert msg in str(ei.value)
    with pytest.raises(oracle.MpsError):
        oracle.MpsFile.parse(text, OptimizationDirection.Minimize)


def test_native_reader_number_forms_and_duplicates():
    """f64::from_str forms (mps.rs:395-402) and CsVec::new's rejection of a repeated variable (lib.rs:247-249)."""
    text = HEAD.replace("R1 4 R2 1", "R1 +4.5e0 R2 .5") + "BOUNDS\n UP BND X inf\n LO BND Y -1E-2\nENDATA\n"
    a, b = mps.MpsFile.parse(text, OptimizationDirection.Minimize), mps.MpsFile.parse_python(text, OptimizationDirection.Minimize)
    assert a.problem.to_csr()[4].tolist() == b.problem.to_csr()[4].tolist() == [4.5, 0.5]
    assert a.problem.var_maxs == b.problem.var_maxs == [float("inf"), float("inf")] and a.problem.var_mins == [0.0, -0.01]
    same_problem(a, oracle.MpsFile.parse(text, OptimizationDirection.Minimize))
    dup = HEAD.replace(" X R2 1\n", " X R2 1\n X R2 2\n") + "ENDATA\n"
    for parse in (mps.MpsFile.parse, mps.MpsFile.parse_python):
        with pytest.raises(ValueError):
            parse(dup, OptimizationDirection.Minimize)
    for bad in ("1_000", "0x10", "1e", "--1", "+-1", ""):
        with pytest.raises(mps.MpsError):
            mps.MpsFile.parse(HEAD.replace("R1 4", "R1 " + bad if bad else "R1"), OptimizationDirection.Minimize)


def test_native_reader_is_fast_and_equal_on_a_large_file():
    """Config-4-shaped text (scaled down for the CPU suite): native == python restatement, natively in well under a second."""
    import time
    text, d = synth.netlib_like(6000, 6000, 20.0, 2)
    t0 = time.perf_counter()
    a = mps.MpsFile.parse(text, d)
    t1 = time.perf_counter()
    b = mps.MpsFile.parse_python(text, d)
    t2 = time.perf_counter()
    for x, y in zip(a.problem.to_csr(), b.problem.to_csr()):
        assert np.array_equal(x, y)
    assert a.variables == b.variables and a.problem.obj_coeffs == b.problem.obj_coeffs
    assert a.problem.var_mins == b.problem.var_mins and a.problem.var_maxs == b.problem.var_maxs
    assert (t1 - t0) < 0.5 * (t2 - t1), (t1 - t0, t2 - t1)


def test_native_reader_parallel_path_equals_serial_and_reports_serial_line_numbers():
    """Files above 8 MB are tokenised by several host threads; result and error messages must be the serial reader's."""
    import os
    text, d = synth.netlib_like(21000, 21000, 21.0, 4)
    assert len(text) > (9 << 20)
    a = mps.MpsFile.parse(text, d)
    os.environ["MLP_MPS_THREADS"] = "1"
    try:
        b = mps.MpsFile.parse(text, d)
    finally:
        del os.environ["MLP_MPS_THREADS"]
    for x, y in zip(a.problem.to_csr(), b.problem.to_csr()):
        assert np.array_equal(x, y)
    assert a.variables == b.variables and a.problem.obj_coeffs == b.problem.obj_coeffs
    same_problem(a, oracle.MpsFile.parse(text, d))
    n_lines = text.count("\n")
    with pytest.raises(mps.MpsError) as ei:
        mps.MpsFile.parse(text.replace("ENDATA", "BOGUS"), d)
    assert str(ei.value) == f"line {n_lines}: expected ENDATA section"
    # an error in the middle of COLUMNS: the fast path gives up and the serial loop reports it with its line number
    lines = text.split("\n")
    k = len(lines) // 2
    assert lines[k].startswith(" X")
    name = lines[k].split()[0]
    lines[k] = f" {name} NOSUCHROW 1.0"
    with pytest.raises(mps.MpsError) as ei:
        mps.MpsFile.parse("\n".join(lines), d)
    assert str(ei.value) == f"line {k + 1}: unknown constraint: NOSUCHROW"
    # a variable whose lines are not contiguous is a re-declaration (mps.rs: "variable ... already declared")
    lines = text.split("\n")
    first_col = next(i for i, l in enumerate(lines) if l.startswith(" X0 "))
    moved = lines.pop(first_col)
    lines.insert(k, moved)
    with pytest.raises(mps.MpsError) as ei:
        mps.MpsFile.parse("\n".join(lines), d)
    assert "variable X0 already declared" in str(ei.value) or "already declared" in str(ei.value)
