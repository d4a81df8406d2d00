"""Runs straight-line GLSL functions of the reference's shader text on the CPU, in float32, to produce golden vectors.

The reference cannot execute here (no Vulkan device, no GLSL compiler), but its pure shading functions are plain expression
code. `load_functions(path, names)` reads the shader file where it lies under /root/reference, turns each named function into a
Python function by a mechanical rewrite (declarations lose their type, swizzles become slices, the GLSL built-ins map to float32
numpy helpers with the same operation order) and returns them. Nothing of the shader text is stored in this repository: the
generator (tools/make_golden.py) commits only the numeric input / output vectors.

Supported subset: functions whose body consists of declarations (with or without initialiser), assignments, compound assignments,
`++`, `?:`, `if` / `else` (single statements or blocks, nested), counted `for (uint i = a; cond; i++)` loops, `while` loops, `continue`,
`break` and returns; scalar / vec3 / vec4 float arithmetic and 32-bit unsigned arithmetic; struct parameters and buffer blocks as
attribute access; calls of functions with `out` / `inout` parameters; the built-ins listed in BUILTINS.
GLSL value semantics are kept: every assignment and every `in` parameter copies its value (structs deeply), compound assignments
never update an array in place.
"""
import copy
import re
import types

import numpy as np

F = np.float32


def _vec(n):
    def ctor(*a):
        parts = []
        for x in a:
            parts.extend(np.asarray(x, F).reshape(-1).tolist())
        if len(parts) == 1:
            parts = parts * n
        assert len(parts) == n, (n, a)
        return np.array(parts, F)
    return ctor


def _dot(a, b):
    a, b = np.asarray(a, F), np.asarray(b, F)
    s = F(a[0] * b[0])
    for k in range(1, a.shape[0]):
        s = F(s + F(a[k] * b[k]))
    return s


def _length(a):
    return F(np.sqrt(_dot(a, a)))


BUILTINS = {
    "vec2": _vec(2), "vec3": _vec(3), "vec4": _vec(4),
    "dot": _dot, "length": _length,
    # normalize() is implementation-defined in GLSL; GPUs evaluate x * inversesqrt(dot(x, x)), which is also the oracle's contract
    "normalize": lambda a: (np.asarray(a, F) * F(F(1.0) / F(np.sqrt(_dot(a, a))))).astype(F),
    "max": lambda a, b: np.maximum(np.asarray(a, F), np.asarray(b, F)).astype(F)[()],
    "min": lambda a, b: np.minimum(np.asarray(a, F), np.asarray(b, F)).astype(F)[()],
    "clamp": lambda x, lo, hi: np.minimum(np.maximum(np.asarray(x, F), F(lo)), F(hi)).astype(F)[()],
    "pow": lambda a, b: np.power(np.asarray(a, F), F(b)).astype(F)[()],
    "mix": lambda a, b, t: (np.asarray(a, F) * (F(1.0) - F(t)) + np.asarray(b, F) * F(t)).astype(F),
    "sqrt": lambda a: np.sqrt(np.asarray(a, F)).astype(F)[()],
    "abs": lambda a: np.abs(np.asarray(a, F)).astype(F)[()],
}
_TYPES = r"(?:void|float|int|uint|bool|vec[234]|ivec[234]|uvec[234]|mat[34])"
_SWIZZLE = {"xyz": "[:3]", "rgb": "[:3]", "xy": "[:2]", "xz": "[[0, 2]]", "yz": "[1:3]", "x": "[0]", "y": "[1]", "z": "[2]", "w": "[3]", "r": "[0]", "g": "[1]", "b": "[2]"}


class U32(int):
    """GLSL uint: 32-bit wrap-around on every operation the hash functions use"""
    def _w(self, v): return U32(int(v) & 0xFFFFFFFF)
    def __add__(self, o): return self._w(int(self) + int(o))
    def __xor__(self, o): return self._w(int(self) ^ int(o))
    def __and__(self, o): return self._w(int(self) & int(o))
    def __or__(self, o): return self._w(int(self) | int(o))
    def __lshift__(self, o): return self._w(int(self) << int(o))
    def __rshift__(self, o): return self._w(int(self) >> int(o))
    __iadd__, __ixor__, __iand__, __ior__ = __add__, __xor__, __and__, __or__


def _is_float(o):
    return isinstance(o, (float, np.floating, np.ndarray))


def _mixed(cls):
    """int OP float converts the integer to float32 first (GLSL implicit conversion); numpy would promote an int SUBCLASS to float64"""
    for op in ("add", "sub", "mul", "truediv"):
        ident = getattr(cls, f"__{op}__")
        def fwd(self, o, _op=op, _ident=ident):
            if _is_float(o):
                return getattr(F(int(self)), f"__{_op}__")(o if isinstance(o, np.ndarray) else F(o))
            return _ident(self, o)
        def rev(self, o, _op=op):
            return getattr(F(o) if not isinstance(o, np.ndarray) else o, f"__{_op}__")(F(int(self)))
        setattr(cls, f"__{op}__", fwd)
        setattr(cls, f"__r{op}__", rev)
    for op in ("lt", "le", "gt", "ge"):
        ident = getattr(int, f"__{op}__")
        def cmp(self, o, _op=op, _ident=ident):
            if _is_float(o):
                return bool(getattr(F(int(self)), f"__{_op}__")(F(o)))
            return _ident(self, o)
        setattr(cls, f"__{op}__", cmp)
    cls.__array_priority__ = 100  # numpy scalars / arrays defer to these reflected operators
    return cls


class I32(int):
    """GLSL int: 32-bit two's-complement wrap-around; mixed with a uint it converts to uint (GLSL 4.60 section 4.1.10)"""
    @staticmethod
    def _w(v):
        v = int(v) & 0xFFFFFFFF
        return I32(v - (1 << 32) if v & 0x80000000 else v)
    def __mul__(self, o): return U32(int(self) & 0xFFFFFFFF) * o if isinstance(o, U32) else I32._w(int(self) * int(o))
    def __xor__(self, o): return U32(int(self) & 0xFFFFFFFF) ^ o if isinstance(o, U32) else I32._w(int(self) ^ int(o))
    def __mod__(self, o):
        assert isinstance(o, U32), "only int % uint is needed"
        return U32((int(self) & 0xFFFFFFFF) % int(o))


class IVec(tuple):
    """ivecN: component-wise addition; ivec * vec converts to float32 first (numpy would promote the integers to float64)"""
    __array_ufunc__ = None  # float-vector OP ivec goes through the reflected operators below
    def __add__(self, o): return IVec(I32._w(int(a) + int(b)) for a, b in zip(self, o))
    def _f(self): return np.array([F(int(c)) for c in self], F)
    def __mul__(self, o): return (self._f() * np.asarray(o, F)).astype(F)
    def __rmul__(self, o): return (np.asarray(o, F) * self._f()).astype(F)


def _ivec(n):
    def ctor(*a):
        parts = []
        for x in a:
            parts.extend(np.asarray(x).reshape(-1).tolist())
        if len(parts) == 1:
            parts = parts * n
        if len(a) == 1 and len(parts) > n:
            parts = parts[:n]  # ivec2(uvec3): the leading components
        assert len(parts) == n
        return IVec(I32._w(int(np.trunc(v))) for v in parts)  # float -> int conversion truncates
    return ctor


class Mat:
    """matN, column-major like GLSL: m[c] is column c. Products accumulate left to right in float32; inverse() goes through
    float64 and is rounded once (the transposed inverse of the model matrix is a per-instance constant in this repository)"""
    def __init__(self, cols): self.c = np.asarray(cols, F)  # c[col, row]
    def __mul__(self, o):
        if isinstance(o, Mat):
            return Mat([self * o.c[k] for k in range(o.c.shape[0])])
        v = np.asarray(o, F)
        acc = (self.c[0] * v[0]).astype(F)
        for k in range(1, self.c.shape[0]):
            acc = (acc + (self.c[k] * v[k]).astype(F)).astype(F)
        return acc


U32.__truediv__ = lambda self, o: U32(int(self) // int(o))  # uint / uint is integral (floats are handled by _mixed)
U32.__sub__ = lambda self, o: U32((int(self) - int(o)) & 0xFFFFFFFF)
I32.__add__ = lambda self, o: I32._w(int(self) + int(o))
I32.__sub__ = lambda self, o: I32._w(int(self) - int(o))
I32.__truediv__ = lambda self, o: I32._w(int(int(self) / int(o)))  # truncating division
_mixed(U32)
_mixed(I32)


U32.__mod__ = lambda self, o: U32(int(self) % int(o))
BUILTINS.update({
    "ivec2": _ivec(2), "ivec3": _ivec(3),
    "mat3": lambda m: Mat(m.c[:3, :3]),
    "transpose": lambda m: Mat(m.c.T),
    "reflect": lambda i, n: (np.asarray(i, F) - (F(F(2.0) * _dot(n, i)) * np.asarray(n, F)).astype(F)).astype(F),
    "bool": lambda a: bool(a), "all": lambda b: bool(np.all(b)),
    "any": lambda b: bool(np.any(b)), "lessThan": lambda a, b: np.asarray(a, F) < np.asarray(b, F),
    "greaterThan": lambda a, b: np.asarray(a, F) > np.asarray(b, F),
    "determinant": lambda m: F(np.linalg.det(m.c.astype(np.float64))),
    "inverse": lambda m: Mat(np.linalg.inv(m.c.astype(np.float64).T).T.astype(F)),
    "cross": lambda a, b: np.array([F(F(a[1] * b[2]) - F(b[1] * a[2])), F(F(a[2] * b[0]) - F(b[2] * a[0])), F(F(a[0] * b[1]) - F(b[0] * a[1]))], F),
    "distance": lambda a, b: _length(np.asarray(a, F) - np.asarray(b, F)),
    # the numerics contract of this repository for transcendental functions: float64 libm, rounded once (oracle/snk_math.h)
    "tan": lambda a: F(np.tan(np.float64(F(a)))),
    "sin": lambda a: F(np.sin(np.float64(F(a)))), "cos": lambda a: F(np.cos(np.float64(F(a)))),
    "uint": lambda a: U32(int(a)) if isinstance(a, int) else U32(int(np.trunc(np.float64(a)))),  # float -> uint truncates
    "float": lambda a: F(a),
    "floatBitsToUint": lambda f: U32(int(np.asarray(f, F).view(np.uint32))),
    "uintBitsToFloat": lambda u: np.asarray(int(u) & 0xFFFFFFFF, np.uint32).view(F)[()],
    "fract": lambda a: (np.asarray(a, F) - np.floor(np.asarray(a, F))).astype(F)[()],
})


def _expr(e):
    # unsigned literals (10u, 0x007FFFFFu) become wrap-around integers
    e = re.sub(r"\b(0x[0-9A-Fa-f]+|\d+)u\b", r"U32(\1)", e)
    e = re.sub(r"\b(\d+\.\d*|\.\d+)f\b", r"\1", e)  # 0.5f
    # float literals are binary32 in GLSL (matters when the other operand is an integer: i * 20382997.34741)
    e = re.sub(r"(?<![\w.])(\d+\.\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?)(?![\w.])", r"F(\1)", e)
    # swizzles after an identifier, `)` or `]` (never inside a float literal)
    return re.sub(r"(?<=[A-Za-z_0-9\)\]])\.(xyz|rgb|xy|xz|yz|x|y|z|w|r|g|b)\b(?!\s*\()", lambda m: _SWIZZLE[m.group(1)], e)


def _val(x):
    """GLSL assigns by value: structs (namespaces) are copied deeply, vectors / matrices are copied"""
    if isinstance(x, types.SimpleNamespace):
        return copy.deepcopy(x)
    if isinstance(x, np.ndarray):
        return x.copy()
    if isinstance(x, Mat):
        return Mat(x.c.copy())
    return x


# struct definitions seen so far: name -> [(member type, member name)]; `PixelReservoir r;` creates the nested members
_STRUCTS = {}


def new_struct(name):
    ns = types.SimpleNamespace()
    for ty, member in _STRUCTS.get(name, []):
        setattr(ns, member, new_struct(ty) if ty in _STRUCTS else None)
    return ns


# signatures of the functions seen so far: name -> [(qualifier, is_basic_type)] per parameter (call sites with out / inout arguments)
_SIGS = {}


def _top_level(st, tokens):
    """index and token of the first occurrence of one of `tokens` outside parentheses / brackets, else (-1, None)"""
    depth = 0
    i = 0
    while i < len(st):
        ch = st[i]
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        elif depth == 0:
            for t in tokens:
                if st.startswith(t, i):
                    return i, t
        i += 1
    return -1, None


def _split_args(argstr):
    out, depth, cur = [], 0, ""
    for ch in argstr:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _ternary(e):
    """c ? a : b (right associative) -> Python conditional expression; the pieces go through _expr / _cond"""
    i, _ = _top_level(e, ["?"])
    if i < 0:
        return _expr(e)
    cond, rest = e[:i], e[i + 1:]
    depth, j, nest = 0, 0, 0
    while j < len(rest):
        ch = rest[j]
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        elif depth == 0 and ch == "?":
            nest += 1
        elif depth == 0 and ch == ":":
            if nest == 0:
                break
            nest -= 1
        j += 1
    return f"(({_ternary(rest[:j])}) if ({_cond(cond)}) else ({_ternary(rest[j + 1:])}))"


def _assign_index(st):
    """index of the top-level plain `=` of an assignment (not ==, <=, >=, !=), else -1"""
    depth = 0
    for i, ch in enumerate(st):
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        elif ch == "=" and depth == 0:
            if st[i + 1:i + 2] == "=" or st[i - 1:i] in ("=", "<", ">", "!", "+", "-", "*", "/", "^", "&", "|"):
                continue
            return i
    return -1


def _call_with_outs(st):
    """`f(a, b, c);` where f has out / inout scalar or vector parameters -> `b, c = f(a, c)` (as _translate returns them)"""
    m = re.match(r"^(\w+)\((.*)\)$", st)
    if not m or m.group(1) not in _SIGS:
        return None
    sig = _SIGS[m.group(1)]
    args = _split_args(m.group(2))
    if len(args) != len(sig) or not any(q == "out" or (q == "inout" and basic) for q, basic in sig):
        return None
    outs = [a for a, (q, basic) in zip(args, sig) if q == "out" or (q == "inout" and basic)]
    ins = [a for a, (q, basic) in zip(args, sig) if q != "out"]
    lhs = ", ".join(_expr(o) for o in outs)
    return f"{lhs}{',' if len(outs) == 1 else ''} = {m.group(1)}({', '.join(_ternary(a) for a in ins)})"


def _stmt(st):
    if re.match(r"(return|else|if)\b", st):
        m = re.match(r"return\s+(.*)$", st)
        return f"return {_ternary(m.group(1))}" if m else _expr(st)
    if st in ("continue", "break"):
        return st
    st = re.sub(rf"^(?:const\s+)?{_TYPES}\s+(\w+)$", r"\1 = None", st)       # `vec3 nt;`
    st = re.sub(rf"^(?:const\s+)?{_TYPES}\s+(?=\w+\s*=)", "", st)
    st = re.sub(r"^(?:const\s+)?[A-Za-z_]\w*\s+(?=[A-Za-z_]\w*\s*=[^=])", "", st)  # declaration with a struct type: `InstanceData instance = ...`
    st = re.sub(r"^([A-Z]\w*)\s+([A-Za-z_]\w*)$", r'\2 = new_struct("\1")', st)  # `TriangleData triangle;`
    st = re.sub(r"^([\w\.\[\]]+)\+\+$", r"\1 += 1", st)
    st = re.sub(r"^([\w\.\[\]]+)--$", r"\1 -= 1", st)
    call = _call_with_outs(st)
    if call:
        return call
    i, op = _top_level(st, ["<<=", ">>=", "+=", "-=", "*=", "/=", "^=", "&=", "|="])
    if i >= 0 and _assign_index(st) < 0:
        lhs, rhs = st[:i].strip(), st[i + len(op):].strip()
        return f"{_expr(lhs)} = {_expr(lhs)} {op[:-1]} ({_ternary(rhs)})"  # never in place: `xy += v` must not write through a swizzle view
    i = _assign_index(st)
    if i >= 0:
        return f"{_expr(st[:i].strip())} = _val({_ternary(st[i + 1:].strip())})"
    return _expr(st)


def _cond(c):
    c = c.replace("&&", " and ").replace("||", " or ")
    c = re.sub(r"!(?!=)", " not ", c)
    return _expr(c)


def _split_paren(st):
    """'if (cond) rest' -> (cond, rest)"""
    i = st.index("(")
    depth, j = 1, i + 1
    while depth:
        depth += {"(": 1, ")": -1}.get(st[j], 0)
        j += 1
    return st[i + 1:j - 1], st[j:].strip()


def _signature(params):
    sig = []
    for p in params.split(","):
        toks = p.strip().split()
        if not toks:
            continue
        q = toks[0] if toks[0] in ("in", "out", "inout") else "in"
        ty = toks[1] if toks[0] in ("in", "out", "inout", "const") else toks[0]
        sig.append((q, bool(re.fullmatch(_TYPES, ty))))
    return sig


def _translate(name, params, body, copy_params=True):
    args, outs, copies = [], [], []
    for p in params.split(","):
        p = p.strip()
        if p:
            toks = p.split()
            if toks[0] == "out":
                outs.append(toks[-1])  # `out` parameters are returned (as a tuple, in declaration order); `inout` structs are mutated
            else:
                args.append(toks[-1])
                if toks[0] == "inout" and re.fullmatch(_TYPES, toks[1]):
                    outs.append(toks[-1])  # scalars / vectors passed `inout` come back in the returned tuple as well
                elif toks[0] != "inout":
                    copies.append(toks[-1])  # `in` parameters are copies of the caller's values
    lines = [f"def {name}({', '.join(args)}):"]
    for c in copies if copy_params else []:
        lines.append(f"    {c} = _val({c})")
    body = re.sub(r"//[^\n]*", "", body)
    body = re.sub(r"\btrue\b", "True", re.sub(r"\bfalse\b", "False", body))
    indent, cur, depth, emitted, epilogue = 1, "", 0, [0], []

    def emit(text):
        lines.append("    " * indent + text)
        emitted[-1] += 1

    for ch in body:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if depth == 0 and ch in ";{}":
            st = " ".join(cur.split())
            cur = ""
            if ch == ";":
                if not st:
                    continue
                assert not re.match(r"(for|while|switch|do)\b", st), f"{name}: loops are outside the supported subset: {st}"
                prefix = ""
                if re.match(r"else\b", st):
                    st = st[4:].strip()
                    prefix = "el" if st.startswith("if") else "else: "
                if st.startswith("if"):  # single-statement form: if (cond) statement;
                    cond, st = _split_paren(st)
                    prefix += f"if {_cond(cond)}: "
                emit(prefix + _stmt(st))
            elif ch == "{":
                assert not re.match(r"(switch|do)\b", st), f"{name}: unsupported control flow: {st}"
                if st.startswith("while"):
                    emit(f"while {_cond(_split_paren(st)[0])}:")
                    epilogue.append(None)
                    indent += 1
                    emitted.append(0)
                    cur = ""
                    continue
                m_for = re.match(r"for\s*\(\s*(?:uint|int)\s+(\w+)\s*=\s*([^;]+);\s*([^;]+);\s*(\w+)\+\+\s*\)$", st)
                if st.startswith("for"):
                    assert m_for and m_for.group(1) == m_for.group(4), f"{name}: only `for (uint i = a; cond; i++)` is supported: {st}"
                    emit(f"{m_for.group(1)} = {_expr(m_for.group(2).strip())}")
                    emit(f"while {_cond(m_for.group(3).strip())}:")
                    indent += 1
                    lines.append("    " * indent + "try:")  # the increment runs on `continue` as well
                    epilogue.append(f"{m_for.group(1)} += 1")
                    indent += 1
                    emitted.append(0)
                    cur = ""
                    continue
                epilogue.append(None)
                if re.match(r"else\s+if\b", st):
                    emit(f"elif {_cond(_split_paren(st)[0])}:")
                elif st.startswith("if"):
                    emit(f"if {_cond(_split_paren(st)[0])}:")
                elif st == "else":
                    emit("else:")
                else:
                    assert not st, f"{name}: unsupported block header: {st}"
                    emit("if True:")
                indent += 1
                emitted.append(0)
            else:
                assert not st, f"{name}: statement without ';' before '}}': {st}"
                ep = epilogue.pop()
                if emitted[-1] == 0:
                    lines.append("    " * indent + "pass")
                emitted.pop()
                indent -= 1
                if ep:  # close the `try:` of a counted loop
                    lines.append("    " * indent + "finally:")
                    lines.append("    " * (indent + 1) + ep)
                    indent -= 1
        else:
            cur += ch
    if outs:
        lines.append("    return " + ", ".join(outs))
    return "\n".join(lines)


def load_functions(path, names, defines=None, env=None):
    """-> dict name -> python function, plus the `defines` found in the file (object-like macros with a numeric value)"""
    text = _read_shader(path)
    env = dict(BUILTINS) if env is None else env
    env["U32"] = U32
    env["F"] = F
    env["namespace"] = namespace
    env["_val"] = _val
    env["new_struct"] = new_struct
    for m in re.finditer(r"^\s*struct\s+(\w+)\s*\{([^}]*)\}", text, re.M):
        body = re.sub(r"//[^\n]*", "", m.group(2))
        _STRUCTS.setdefault(m.group(1), [tuple(d.split()[-2:]) for d in body.split(";") if len(d.split()) >= 2])
    for m in re.finditer(rf"^\s*(?:{_TYPES}|[A-Z]\w*)\s+(\w+)\s*\(([^)]*)\)\s*\{{", text, re.M):
        _SIGS.setdefault(m.group(1), _signature(m.group(2)))
    for m in re.finditer(r"^#define\s+(\w+)\s+([0-9.]+)f?\s*$", text, re.M):
        env[m.group(1)] = F(float(m.group(2)))
    for m in re.finditer(r"^\s*const\s+float\s+(\w+)\s*=\s*([0-9.]+)f?\s*;", text, re.M):  # file-scope float constants
        env.setdefault(m.group(1), F(float(m.group(2))))
    env.update(defines or {})
    for name in names:
        m = re.search(rf"^\s*(?:{_TYPES}|[A-Z]\w*)\s+{name}\s*\(([^)]*)\)\s*\{{", text, re.M)
        assert m, f"{name} not found in {path}"
        depth, i = 1, m.end()
        while depth:
            depth += {"{": 1, "}": -1}.get(text[i], 0)
            i += 1
        exec(_translate(name, m.group(1), text[m.end():i - 1]), env)  # noqa: S102 - the input is the reference shader, read-only
    return env


def _read_shader(path):
    text = open(path, encoding="utf-8", errors="replace").read()
    return re.sub(r"\bas\b", "accel", text)  # the acceleration-structure binding is called `as`, a Python keyword


def load_block(path, start_regex, name, params, prologue="", result="", env=None, defines=None, occurrence=1):
    """a brace-delimited block INSIDE a function (the text from the match of `start_regex`, which must end at its `{`, to the
    matching `}`) as a function `name(params)`: `prologue` statements run first, `result` is returned"""
    text = _read_shader(path)
    ms = list(re.finditer(start_regex, text))
    assert len(ms) >= occurrence and text[ms[occurrence - 1].end() - 1] == "{", f"block not found: {start_regex}"
    m = ms[occurrence - 1]
    depth, i = 1, m.end()
    while depth:
        depth += {"{": 1, "}": -1}.get(text[i], 0)
        i += 1
    env = dict(BUILTINS) if env is None else env
    env["U32"], env["F"], env["namespace"], env["_val"], env["new_struct"] = U32, F, namespace, _val, new_struct
    env.update(defines or {})
    src = _translate(name, params, prologue + text[m.start():i], copy_params=False)
    if result:
        src += f"\n    return {result}"
    exec(src, env)  # noqa: S102
    return env, text.count("\n", 0, m.start()) + 1, text.count("\n", 0, i) + 1


def load_lines(path, first_line, last_line, must_start_with, name, params, result="", env=None, defines=None):
    """lines first_line..last_line (1-based, inclusive) of a function body as a function `name(params)`"""
    lines = open(path, encoding="utf-8", errors="replace").read().split("\n")
    assert lines[first_line - 1].strip().startswith(must_start_with), lines[first_line - 1]
    env = dict(BUILTINS) if env is None else env
    env["U32"], env["F"], env["namespace"], env["_val"], env["new_struct"] = U32, F, namespace, _val, new_struct
    env.update(defines or {})
    src = _translate(name, params, "\n".join(lines[first_line - 1:last_line]), copy_params=False)
    if result:
        src += f"\n    return {result}"
    exec(src, env)  # noqa: S102
    return env


def namespace(**kw):
    return types.SimpleNamespace(**kw)
