"""CPU-only checks of the product: the C-ABI library loads and exports every declared symbol, fails loudly without a
device, host-side sharding logic, generator parity with the oracle, and a world_size-2 gloo run of the pricing
arg-reduce."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import minilp_b200 as mb
from minilp_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "minilp_b200.h")).read()
    declared = set(re.findall(r"\b(mlp_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = _lib.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert mb.device_count() == 0
    with pytest.raises(mb.api.NoDevice):
        mb.Solver(4, 4)


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_generators_agree_with_oracle(kind):
    import oracle
    lp = mb.synth_dense(kind, 41, 29, 5, threads=2)
    d, a, obj, mins, maxs, ops, rhs = oracle.synth_dense(kind, 41, 29, 5)
    assert d == lp.direction
    for x, y in ((a, lp.a), (obj, lp.obj), (mins, lp.mins), (maxs, lp.maxs), (ops, lp.ops), (rhs, lp.rhs)):
        assert np.array_equal(x, y)
    # row-range streaming gives the same rows
    part = mb.synth_rows(kind, 41, 29, 5, 10, 7)
    assert np.array_equal(part, lp.a[10:17])


def test_shard_range_partitions_columns():
    for n in (1, 15, 16, 17, 1000, 50000, 200000):
        for world in (1, 2, 4, 8):
            prev = 0
            for r in range(world):
                b, e = mb.shard_range(n, world, r)
                assert b == prev and e >= b
                prev = e
            assert prev == n


def test_reduce_candidates_rules():
    assert mb.reduce_candidates([1.0, 3.0, 2.0], [0, 1, 2], [5, 6, 7]) == 1
    assert mb.reduce_candidates([3.0, 3.0], [9, 4], [1, 2]) == 1          # tie -> lowest position (solver.rs:719)
    assert mb.reduce_candidates([9.0, 1.0], [0, 1], [-1, 2]) == 1         # var < 0: no candidate
    assert mb.reduce_candidates([0.0, 0.0], [0, 1], [-1, -1]) == -1


def test_gloo_world2_pricing_argreduce():
    """N>1 host path on CPU: two ranks each price their own column block (numpy stands in for the device scan), all-gather
    the (score, pos, var) triples over gloo and apply mlp_reduce_candidates; every rank must pick the single-rank winner."""
    script = os.path.join(ROOT, "tests", "_gloo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29531", script],
                         capture_output=True, text=True, env=env, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("GLOO_OK") == 2, out.stdout + out.stderr


def test_tuning_knobs_match_the_header():
    """api.Engine.TUNE mirrors the MLP_TUNE_* enum of include/minilp_b200.h."""
    import re
    from minilp_b200.api import Engine
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "minilp_b200.h")).read()
    enum = dict((k.lower(), int(v)) for k, v in re.findall(r"MLP_TUNE_([A-Z0-9_]+)\s*=\s*(\d+)", hdr))
    assert enum == Engine.TUNE


def test_trivial_solution_has_the_incremental_methods():
    """lib.rs:368-423 on a problem without constraints (no engine is created: CPU only)."""
    import numpy as np
    import pytest
    import minilp_b200 as mb
    p = mb.Problem(mb.OptimizationDirection.Maximize)
    a = p.add_var(1.0, (0.0, 4.0))
    b = p.add_var(-1.0, (1.0, np.inf))
    sol = p.solve()
    assert (sol[a], sol[b], sol.objective()) == (4.0, 1.0, 3.0)
    assert dict(sol) == {0: 4.0, 1: 1.0}
    f = sol.fix_var(a, 2.0)
    assert (f[a], f.objective()) == (2.0, 1.0) and sol[a] == 4.0
    with pytest.raises(mb.Infeasible):
        sol.fix_var(a, 5.0)
    u, was = f.unfix_var(a)
    assert was and u[a] == 4.0
    assert sol.unfix_var(b) == (sol, False)
    assert f.clone()[a] == 2.0
    with pytest.raises(ValueError):
        sol.add_gomory_cut(a)
    with pytest.raises(mb.Infeasible):
        sol.add_constraint([], mb.ComparisonOp.Ge, 1.0)  # solver.rs:558-570


def test_rust_binding_is_generated_from_the_header():
    """bindings/minilp_b200.rs is regenerated from include/minilp_b200.h and must be current;
    it names every symbol the ctypes table — and with it the .so — knows."""
    import importlib.util
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_rust_bindings", os.path.join(root, "scripts", "gen_rust_bindings.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    text, names = gen.generate()
    assert open(os.path.join(root, "bindings", "minilp_b200.rs")).read() == text, "run python scripts/gen_rust_bindings.py"
    from minilp_b200 import _lib
    assert set(names) == set(_lib.SIGNATURES)
    assert set(re.findall(r"pub fn (\w+)\(", text)) == set(_lib.SIGNATURES)
