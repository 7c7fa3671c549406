"""Free-format MPS reader: host-side mirror of the reference's `MpsFile::parse` (src/mps.rs:39-329).

MPS parsing stays on the host (BASELINE north_star); it feeds `Problem`, whose `solve()` hands a CSR matrix to the device
engine.  Behaviour follows the reference line by line: whitespace tokenisation, `*` comment lines, sections
NAME / ROWS / COLUMNS / RHS / [RANGES] / [BOUNDS] / ENDATA, first N row = objective (further N rows are ignored free rows),
only the FIRST RHS / RANGES / BOUNDS vector is used (mps.rs:193-198, 223-228, 253-258), bound types LO / UP / FX / FR only
(282), a negative UP bound without LO gives (-inf, max] (299), ranged rows become two constraints (306-321).  Syntax
errors raise MpsError carrying the reference's "line N: ..." message (io::ErrorKind::InvalidData there).
"""
import io

from .api import ComparisonOp, Problem

INF = float("inf")


class MpsError(ValueError):
    """io::Error of kind InvalidData in the reference (mps.rs:341-346)."""


class _Lines:
    """mps.rs:332-358: skips comment (`*`) and blank lines, strips trailing whitespace, keeps the 1-based line number."""

    def __init__(self, text):
        self._f = io.StringIO(text, newline="\n")  # BufRead::read_line: lines end at \n only
        self.cur = ""
        self.idx = 0

    def to_next(self):
        while True:
            self.idx += 1
            line = self._f.readline()
            if line == "":
                self.cur = ""
                return
            if line.startswith("*"):
                continue
            t = line.rstrip()
            if t:
                self.cur = t
                return

    def err(self, msg):
        return MpsError(f"line {self.idx}: {msg}")


class _Tokens:
    def __init__(self, lines):
        self.line_idx = lines.idx
        self.it = iter(lines.cur.split())

    def next(self):
        t = next(self.it, None)
        if t is None:
            raise MpsError(f"line {self.line_idx}: unexpected end of line")
        return t

    def opt(self):
        return next(self.it, None)


def _f64(tok, line_idx):
    try:
        if "_" in tok:  # Python's float() accepts digit separators, Rust's f64::from_str does not
            raise ValueError
        return float(tok)
    except ValueError:
        raise MpsError(f"line {line_idx}: couldn't parse float from string: `{tok}`") from None


def _kv_pairs(tokens):
    """mps.rs:404-431: one or two (name, value) pairs per line."""
    k1 = tokens.next()
    v1 = _f64(tokens.next(), tokens.line_idx)
    k2 = tokens.opt()
    if k2 is None:
        return [(k1, v1)]
    v2 = _f64(tokens.next(), tokens.line_idx)
    return [(k1, v1), (k2, v2)]


class MpsFile:
    """mps.rs:8-16: problem_name, variables (name -> variable index), problem."""

    def __init__(self, problem_name, variables, problem):
        self.problem_name, self.variables, self.problem = problem_name, variables, problem

    @classmethod
    def parse(cls, text, direction):
        """The product path: the native reader (csrc/mps_reader.cpp behind mlp_mps_parse — one pass over the buffer, flat
        arrays out; ~100x the pure-Python restatement below, which is kept as `parse_python` and cross-checked in the tests)."""
        import ctypes as C

        import numpy as np

        from . import _lib
        if hasattr(text, "read"):
            text = text.read()
        raw = text.encode() if isinstance(text, str) else bytes(text)
        L = _lib.lib()
        h = C.c_void_p()
        st = L.mlp_mps_parse(raw, len(raw), C.byref(h))
        if st != 0:
            msg = L.mlp_last_error().decode()
            if msg.startswith("line "):
                raise MpsError(msg)
            raise ValueError(msg)  # CsVec::new's panic on a repeated variable (lib.rs:247-249)
        try:
            nv, nc, nnz, nb = (C.c_int64() for _ in range(4))
            L.mlp_mps_sizes(h, C.byref(nv), C.byref(nc), C.byref(nnz), C.byref(nb))
            nv, nc, nnz, nb = nv.value, nc.value, nnz.value, nb.value
            obj, mins, maxs = np.empty(nv), np.empty(nv), np.empty(nv)
            row_ptr, col_idx, vals = np.empty(nc + 1, dtype=np.int64), np.empty(nnz, dtype=np.int32), np.empty(nnz)
            ops, rhs = np.empty(nc, dtype=np.int32), np.empty(nc)
            blob, off = C.create_string_buffer(max(nb, 1)), np.empty(nv + 1, dtype=np.int64)
            pd, pi64, pi32 = _lib.pd, _lib.pi64, _lib.pi32
            L.mlp_mps_export(h, obj.ctypes.data_as(pd), mins.ctypes.data_as(pd), maxs.ctypes.data_as(pd),
                             row_ptr.ctypes.data_as(pi64), col_idx.ctypes.data_as(pi32), vals.ctypes.data_as(pd),
                             ops.ctypes.data_as(pi32), rhs.ctypes.data_as(pd), blob, off.ctypes.data_as(pi64))
            name = L.mlp_mps_name(h).decode()
        finally:
            L.mlp_mps_free(h)
        names = blob.raw[:nb]
        o = off.tolist()
        variables = {names[o[j]:o[j + 1]].decode(): j for j in range(nv)}
        return cls(name, variables, Problem.from_arrays(direction, obj, mins, maxs, row_ptr, col_idx, vals, ops, rhs))

    @classmethod
    def parse_python(cls, text, direction):
        """Pure-Python restatement of mps.rs:39-329, line by line (reference for the native reader's behaviour)."""
        if hasattr(text, "read"):
            text = text.read()
        lines = _Lines(text)
        lines.to_next()
        tk = _Tokens(lines)
        if tk.next() != "NAME":
            raise lines.err("expected NAME section")
        problem_name = tk.opt() or ""

        obj_name = None
        free_rows = set()
        constraints = []  # [lhs list, op, rhs, range]
        cidx = {}
        lines.to_next()
        if lines.cur != "ROWS":
            raise lines.err("expected ROWS section")
        while True:
            lines.to_next()
            if not lines.cur.startswith(" "):
                break
            tk = _Tokens(lines)
            row_type, name = tk.next(), tk.next()
            if row_type == "N":
                if obj_name is None:
                    obj_name = name
                else:
                    free_rows.add(name)
                continue
            op = {"L": ComparisonOp.Le, "G": ComparisonOp.Ge, "E": ComparisonOp.Eq}.get(row_type)
            if op is None:
                raise lines.err(f"unexpected row type {row_type}")
            if name in cidx:
                raise lines.err(f"row {name} already declared")
            cidx[name] = len(constraints)
            constraints.append([[], op, 0.0, 0.0])
        if obj_name is None:
            raise lines.err("objective function name not declared")

        var_defs = []  # [min|None, max|None, obj]
        vidx = {}
        if lines.cur != "COLUMNS":
            raise lines.err("expected COLUMNS section")
        cur_var, cur_name, cur_def = 0, "", [None, None, 0.0]
        while True:
            lines.to_next()
            if not lines.cur.startswith(" "):
                break
            tk = _Tokens(lines)
            name = tk.next()
            if name != cur_name:
                if name in vidx:
                    raise lines.err(f"variable {name} already declared")
                if cur_name:
                    vidx[cur_name] = cur_var
                    var_defs.append(cur_def)
                    cur_def = [None, None, 0.0]
                    cur_var += 1
                cur_name = name
            for key, val in _kv_pairs(tk):
                if key == obj_name:
                    cur_def[2] = val
                elif key in cidx:
                    constraints[cidx[key]][0].append((cur_var, val))
                elif key not in free_rows:
                    raise lines.err(f"unknown constraint: {key}")
        if cur_name:
            vidx[cur_name] = cur_var
            var_defs.append(cur_def)

        if lines.cur != "RHS":
            raise lines.err("expected RHS section")
        vec = None
        while True:
            lines.to_next()
            if not lines.cur.startswith(" "):
                break
            tk = _Tokens(lines)
            vn = tk.next()
            if vec is None:
                vec = vn
            elif vec != vn:
                continue  # only the first RHS vector
            for key, val in _kv_pairs(tk):
                if key == obj_name:
                    raise lines.err("setting objective in RHS section is not supported")
                if key not in cidx:
                    raise lines.err(f"unknown constraint: {key}")
                constraints[cidx[key]][2] = val

        if lines.cur == "RANGES":
            vec = None
            while True:
                lines.to_next()
                if not lines.cur.startswith(" "):
                    break
                tk = _Tokens(lines)
                vn = tk.next()
                if vec is None:
                    vec = vn
                elif vec != vn:
                    continue
                for key, val in _kv_pairs(tk):
                    if key not in cidx:
                        raise lines.err(f"unknown constraint: {key}")
                    constraints[cidx[key]][3] = val

        if lines.cur == "BOUNDS":
            vec = None
            while True:
                lines.to_next()
                if not lines.cur.startswith(" "):
                    break
                tk = _Tokens(lines)
                btype = tk.next()
                vn = tk.next()
                if vec is None:
                    vec = vn
                elif vec != vn:
                    continue
                vname = tk.next()
                if vname not in vidx:
                    raise lines.err(f"unknown variable: {vname}")
                vd = var_defs[vidx[vname]]
                if btype == "FR":
                    vd[0], vd[1] = -INF, INF
                    continue
                val = _f64(tk.next(), lines.idx)
                if btype == "LO":
                    vd[0] = val
                elif btype == "UP":
                    vd[1] = val
                elif btype == "FX":
                    vd[0] = vd[1] = val
                else:
                    raise lines.err(f"bound type {btype} is not supported")

        if lines.cur != "ENDATA":
            raise lines.err("expected ENDATA section")

        problem = Problem(direction)
        for mn, mx, oc in var_defs:  # mps.rs:294-303
            if mn is not None and mx is not None:
                b = (mn, mx)
            elif mn is not None:
                b = (mn, INF)
            elif mx is not None:
                b = (-INF, mx) if mx < 0.0 else (0.0, mx)
            else:
                b = (0.0, INF)
            problem.add_var(oc, b)
        for lhs, op, rhs, rng in constraints:  # mps.rs:305-322
            if rng == 0.0:
                problem.add_constraint(lhs, op, rhs)
                continue
            if op == ComparisonOp.Ge:
                lo, hi = rhs, rhs + abs(rng)
            elif op == ComparisonOp.Le:
                lo, hi = rhs - abs(rng), rhs
            elif rng > 0.0:
                lo, hi = rhs, rhs + rng
            else:
                lo, hi = rhs + rng, rhs
            problem.add_constraint(list(lhs), ComparisonOp.Ge, lo)
            problem.add_constraint(lhs, ComparisonOp.Le, hi)
        return cls(problem_name, vidx, problem)
