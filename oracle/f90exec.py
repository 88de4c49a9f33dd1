"""f90exec -- ORACLE-SIDE TEST INFRASTRUCTURE (never imported by the product).

Executes the reference's OWN Fortran source text: a line-by-line translator from the restricted Fortran 90 subset the NEMO
tracer, boundary-condition and decomposition routines are written in to Python, which is then exec'ed on numpy arrays.
Understood: SUBROUTINE / FUNCTION, declarations (explicit-shape locals, ALLOCATABLE + ALLOCATE, OPTIONAL dummies, module variables),
DO (with step) / IF - ELSE IF - ELSE (block and one-line) / SELECT CASE, assignments to scalars, elements, sections and whole arrays,
array-valued expressions, integer division, MAX / MIN / ABS / SIGN / MOD / SIZE / MAXVAL / MINVAL / COUNT / REAL / CMPLX / AIMAG,
CALL with keyword arguments, scalar INTENT(out / inout) arguments (handed back by position), message buffers passed by their first
element, RETURN / STOP; I/O, NAMELIST and FORMAT statements are skipped.  `cpp` is a small C preprocessor (conditionals, object- and
function-like macros) for the generic .h90 files.  There is no Fortran compiler in this image or on the GPU box (oracle/_ref_recipe/),
so this is the nearest thing to "the reference run here": the arithmetic expressions, loop bounds and statement order come
from the reference's file, character for character, not from anybody's reading of it.  It is used to pin oracle/*.c
(tests/test_cpu_reference_exec.py) and to generate tests/golden/ref_exec_*.npz (oracle/_ref_recipe/make_ref_exec_golden.py).

What the translation guarantees, and what it assumes:
  * REAL(wp) arithmetic is IEEE binary64, one rounding per operation, no contraction (Python floats / numpy float64): what the
    reference's arch file asks of gfortran (-O3, no -ffast-math; arch/arch-linux_gfortran.fcm:48).
  * Real literals without a kind suffix are double precision, as under the arch file's -fdefault-real-8 (`literal_report` lists
    the ones that would differ in single precision, so the assumption is visible).
  * Operators keep Fortran's precedence and left-to-right association (Python's are the same for + - * / and **).
  * MAX / MIN follow gfortran's expansion (first argument, replaced by a later one that compares greater / smaller).  The
    intrinsic SIGN transfers the IEEE sign bit (F2003, gfortran) -- but the reference's gfortran arch file defines
    key_nosignedzero (arch/arch-linux_gfortran.fcm:46), under which lib_fortran.F90:334-351 overrides SIGN by "pb >= 0 -> |pa|":
    ref_exec loads that function from the reference's text and binds SIGN to it; `sign_mode="ieee"` keeps the intrinsic.
  * Local arrays start as NaN: a read of an undefined element that reaches the result would show.
  * Index expressions are shifted to 0-based at translation time; an out-of-range 0 index (Fortran reads the element before the
    array, e.g. pt_in(ji,jj,ikb-1) on land columns with mbkt = 1) wraps to the last element here -- such values must not reach
    the result either way, and the comparison with the oracle (which skips the read) would show if they did.
  * CALLs to routines of the same file run their translation; other CALLs (lbc_lnk*, trd_tra, dia_*) and module variables are
    whatever the caller puts in the namespace.
Nothing of the reference is copied into the repository: the translation lives in memory only."""
import math
import re

import numpy as np

INTRINSICS = {"max": "f_max", "min": "f_min", "abs": "f_abs", "sign": "f_sign", "real": "f_real", "mod": "f_mod", "int": "f_int",
              "nint": "f_nint", "sqrt": "f_sqrt", "sum": "f_sum", "maxval": "f_maxval", "minval": "f_minval", "present": "f_present",
              "trim": "f_trim", "size": "f_size", "cmplx": "f_cmplx", "aimag": "f_aimag", "count": "f_count",
              "maxloc": "f_maxloc", "minloc": "f_minloc", "isnan": "f_isnan", "nint": "f_nint"}


# ---- run-time support (the namespace translated code runs in) ---------------------------------------------------------------
def f_max(a, *rest):
    r = a
    for b in rest:
        if isinstance(r, np.ndarray) or isinstance(b, np.ndarray):
            r = np.where(np.asarray(b) > np.asarray(r), b, r)
        elif b > r:
            r = b
    return r


def f_min(a, *rest):
    r = a
    for b in rest:
        if isinstance(r, np.ndarray) or isinstance(b, np.ndarray):
            r = np.where(np.asarray(b) < np.asarray(r), b, r)
        elif b < r:
            r = b
    return r


def f_abs(a):
    return np.abs(a) if isinstance(a, np.ndarray) else abs(a)


def f_sign(a, b):
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        return np.copysign(np.abs(a), b)
    return math.copysign(abs(a), b)


def f_real(a, kind=None):
    if isinstance(a, complex):
        return a.real                                     # REAL( z ): the first component of the (sum, error) pair of DDPDD
    return a.astype(np.float64) if isinstance(a, np.ndarray) else float(a)


def f_mod(a, p):
    return int(math.fmod(a, p)) if isinstance(a, (int, np.integer)) and isinstance(p, (int, np.integer)) else math.fmod(a, p)


def f_int(a):
    return int(a)


def f_nint(a):
    return int(math.floor(a + 0.5)) if a >= 0 else -int(math.floor(-a + 0.5))


def f_div(a, b):
    """Fortran `/`: truncating for two integers, IEEE otherwise (x / 0 is Inf or NaN as in Fortran, not a Python exception)"""
    if isinstance(a, (int, np.integer)) and isinstance(b, (int, np.integer)):
        q = abs(int(a)) // abs(int(b))
        return q if (a >= 0) == (b >= 0) else -q
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        return a / b
    with np.errstate(all="ignore"):
        return np.float64(a) / np.float64(b)


def _reduce(fn, empty, a, dim, mask):
    """MAXVAL / MINVAL( a [, DIM=dim] [, MASK=mask] ): the value over an empty set is -HUGE / +HUGE"""
    a = np.asarray(a)
    if mask is not None:
        a = np.where(mask, a, empty)
    if a.size == 0:
        return empty
    if a.dtype.kind == "f" and np.isnan(a).any():          # gfortran: NaN operands are skipped unless every operand is one
        if np.isnan(a).all():
            return np.float64(np.nan)
        a = np.where(np.isnan(a), empty, a)
    return fn(a) if dim is None else fn(a, axis=int(dim) - 1)


def _loc(a, mask, want_max):
    """MAXLOC / MINLOC( a [, MASK=mask] ): 1-based indices of the FIRST extremum in array element order (0, ... for an empty set);
    NaN operands are skipped, as gfortran does"""
    a = np.asarray(a, dtype=np.float64)
    ok = ~np.isnan(a) if mask is None else (np.asarray(mask) & ~np.isnan(a))
    if not ok.any():
        return np.zeros(a.ndim, dtype=np.int64)
    v = np.where(ok, a, -np.inf if want_max else np.inf)
    flat = v.ravel(order="F")
    k = int(np.argmax(flat) if want_max else np.argmin(flat))
    return np.array(np.unravel_index(k, a.shape, order="F"), dtype=np.int64) + 1


def f_ref(a, idx):
    """the storage of the Fortran-ordered array `a` from element `idx` (0-based) on, as a flat view: what a routine receives
    when it is handed an array element as the start of a buffer"""
    assert a.flags["F_CONTIGUOUS"]
    return a.reshape(-1, order="F")[int(np.ravel_multi_index(tuple(int(i) for i in idx), a.shape, order="F")):]


def f_alloc(shape, integer=False):
    shape = tuple(int(n) for n in shape)
    return np.zeros(shape, dtype=np.int32, order="F") if integer else np.full(shape, np.nan, order="F")


RUNTIME = dict(wp=8, f_ref=f_ref, f_max=f_max, f_min=f_min, f_abs=f_abs, f_sign=f_sign, f_real=f_real, f_mod=f_mod, f_int=f_int, f_nint=f_nint,
               f_sqrt=math.sqrt, f_sum=np.sum, f_maxval=lambda a, dim=None, mask=None: _reduce(np.max, -np.finfo(np.float64).max, a, dim, mask),
               f_minval=lambda a, dim=None, mask=None: _reduce(np.min, np.finfo(np.float64).max, a, dim, mask), f_div=f_div, f_alloc=f_alloc, np=np,
               f_trim=lambda s: s.rstrip(), f_present=lambda a: a is not None,
               f_size=lambda a, dim=None: a.size if dim is None else a.shape[dim - 1],
               f_cmplx=lambda re_, im=0.0, kind=None: complex(float(re_), float(im)), f_aimag=lambda z: z.imag,
               f_count=lambda a: int(np.count_nonzero(a)), f_isnan=lambda x: bool(np.isnan(x)),
               f_maxloc=lambda a, mask=None: _loc(a, mask, True), f_minloc=lambda a, mask=None: _loc(a, mask, False))


# ---- source preparation --------------------------------------------------------------------------------------------------------
def _strip_comment(line):
    out, q = [], None
    for ch in line:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            out.append(ch)
        elif ch == "!":
            break
        else:
            out.append(ch)
    return "".join(out).rstrip()


def _lower_outside_strings(s):
    out, q = [], None
    for ch in s:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        else:
            if ch in "'\"":
                q = ch
            out.append(ch.lower())
    return "".join(out)


def _split_top(s, sep):
    """split at `sep` outside parentheses and string literals"""
    parts, depth, q, cur = [], 0, None, []
    for ch in s:
        if q:
            cur.append(ch)
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch
        elif ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == sep and depth == 0:
            parts.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    parts.append("".join(cur))
    return parts


def statements(text, defines=None, defined=()):
    """free-form source -> list of lower-cased statements: comments dropped, continuations joined, `;` split, cpp macros of
    `defines` substituted as whole words, `#if defined KEY` / `#if ! defined KEY` / `#else` / `#endif` blocks resolved
    against the set `defined` (other cpp lines -- include, define, undef -- are dropped)"""
    out, cur, take = [], "", [True]
    for raw in text.split("\n"):
        if raw.lstrip().startswith("#"):
            d = raw.lstrip()[1:].strip()
            m = re.match(r"if\s*(!)?\s*defined\s*\(?\s*(\w+)\s*\)?\s*$", d) or re.match(r"if(n)?def\s+(\w+)\s*$", d)
            if m:
                take.append((m.group(2) in defined) != bool(m.group(1)))
            elif d.startswith("if"):
                raise SyntaxError("cpp condition not understood: " + raw)
            elif d.startswith("else"):
                take[-1] = not take[-1]
            elif d.startswith("endif"):
                take.pop()
            continue
        if not all(take):
            continue
        line = _strip_comment(raw).strip()
        if not line:
            continue
        if line.startswith("&"):
            line = line[1:].lstrip()
        if line.endswith("&"):
            cur += line[:-1].rstrip() + " "
            continue
        cur += line
        for st in _split_top(cur, ";"):
            st = _lower_outside_strings(st.strip())
            for k, v in (defines or {}).items():
                st = re.sub(r"\b%s\b" % re.escape(k), v, st)
            if st:
                out.append(st)
        cur = ""
    return out


def cpp(text, defined=(), macros=None):
    """A small C preprocessor for the reference's generic .h90 files: `#if defined KEY` / `#if ! defined KEY` / `#else` / `#endif`
    (nested), `#define NAME body`, `#define NAME(a,b) body`, `#undef`, and expansion of those macros in the code lines (repeated
    until nothing changes, as cpp rescans).  `defined`: keys defined from outside; `macros`: {name: body} object-like macros defined
    from outside (e.g. ROUTINE_NFD).  `#include` lines are dropped.  Returns the preprocessed text (no `#` lines left)."""
    obj = dict(macros or {})
    defined = set(defined) | set(obj)
    fun = {}
    out, take = [], [True]

    def expand(line):
        for _ in range(20):
            before = line
            for name, (params, body) in fun.items():
                pos = 0
                while True:
                    m = re.search(r"\b%s\s*\(" % re.escape(name), line[pos:])
                    if not m:
                        break
                    start = pos + m.start()
                    j = pos + m.end() - 1
                    k = Translator._match_paren(line, j)
                    args = [a.strip() for a in _split_top(line[j + 1:k], ",")]
                    rep = body
                    for prm, arg in zip(params, args):
                        rep = re.sub(r"\b%s\b" % re.escape(prm), lambda _m, a=arg: a, rep)
                    line = line[:start] + rep + line[k + 1:]
                    pos = start + len(rep)
            for name, body in obj.items():
                line = re.sub(r"\b%s\b" % re.escape(name), lambda _m, b=body: b, line)
            if line == before:
                break
        return line

    for raw in text.split("\n"):
        st = raw.lstrip()
        if st.startswith("#"):
            d = st[1:].strip()
            m = re.match(r"if\s*(!)?\s*defined\s*\(?\s*(\w+)\s*\)?\s*$", d) or re.match(r"if(n)?def\s+(\w+)\s*$", d)
            if m:
                take.append(all(take) and ((m.group(2) in defined) != bool(m.group(1))))
            elif d.startswith("if"):
                raise SyntaxError("cpp condition not understood: " + raw)
            elif d.startswith("else"):
                take[-1] = all(take[:-1]) and not take[-1]
            elif d.startswith("endif"):
                take.pop()
            elif all(take) and d.startswith("define"):
                m = re.match(r"define\s+(\w+)\(([^)]*)\)\s*(.*)$", d)
                if m:
                    fun[m.group(1)] = ([a.strip() for a in m.group(2).split(",")], m.group(3).strip())
                else:
                    m = re.match(r"define\s+(\w+)\s*(.*)$", d)
                    obj[m.group(1)] = m.group(2).strip()                 # an empty body (e.g. LBC_ARG) expands to nothing
                    defined.add(m.group(1))
            elif all(take) and d.startswith("undef"):
                nm = d.split()[1]
                fun.pop(nm, None); obj.pop(nm, None); defined.discard(nm)
            continue
        if all(take):
            out.append(expand(raw))
    return "\n".join(out)


def cpp_defines(text, defined=()):
    """`#  define name value` lines of a small header, honouring one level of #if defined KEY / #else / #endif"""
    out, take = {}, [True]
    for raw in text.split("\n"):
        s = raw.strip()
        if not s.startswith("#"):
            continue
        s = s[1:].strip()
        if s.startswith("if defined"):
            take.append(s.split()[2] in defined)
        elif s.startswith("else"):
            take[-1] = not take[-1]
        elif s.startswith("endif"):
            take.pop()
        elif s.startswith("define") and all(take):
            _, name, val = s.split(None, 2)
            out[name.lower()] = val.strip().lower()
    return out


def literal_report(text):
    """real literals without a kind suffix whose single- and double-precision values differ (they are doubles under
    -fdefault-real-8; listed so that the assumption is visible)"""
    bad = set()
    for st in statements(text):
        st = re.sub(r"'[^']*'|\"[^\"]*\"", "", st)
        for m in re.finditer(r"(?<![\w.])(\d+\.\d*(?:e[+-]?\d+)?|\d*\.\d+(?:e[+-]?\d+)?|\d+e[+-]?\d+)(?![\w.])", st):
            if st[m.end():m.end() + 3] == "_wp":
                continue
            v = float(m.group(1))
            if float(np.float32(v)) != v:
                bad.add(m.group(1))
    return sorted(bad)


# ---- the translator ------------------------------------------------------------------------------------------------------------
class Translator:
    #: external routines that receive a message buffer by its FIRST ELEMENT (sequence association), e.g.
    #: CALL mppsend( 2, zt3we(1,1,1,1,1,1), imigr, noea, ml_req1 ): the element is passed as a flat view of the array from there on
    BYREF_CALLEES = {"mppsend": (1,), "mpprecv": (1,), "mpi_allgather": (0, 3)}        # name -> positions of the buffer arguments

    def __init__(self, arrays=(), int_arrays=()):
        self.global_arrays = set(a.lower() for a in arrays) | set(a.lower() for a in int_arrays)
        self.global_int_arrays = set(a.lower() for a in int_arrays)
        self.module_names = set()                         # variables declared at module level: `global` inside the routines

    def scan_module_level(self, sts):
        """names (and which of them are arrays) declared between MODULE and CONTAINS"""
        for st in sts:
            if st == "contains" or re.match(r"(subroutine|function)\s+\w+", st):
                break
            if re.match(r"(integer|real|logical|character)", st) and "::" in st:
                left, right = st.split("::", 1)
                is_dim = "dimension" in left
                for ent in _split_top(right, ","):
                    m = re.match(r"\s*([a-z_]\w*)\s*(\(.*\))?", ent)
                    self.module_names.add(m.group(1))
                    if is_dim or m.group(2):
                        self.global_arrays.add(m.group(1))

    # -- expressions --
    def expr(self, s, arrays):
        s = s.strip()
        out, i, n = [], 0, len(s)
        while i < n:
            ch = s[i]
            if ch in "'\"":
                j = s.index(ch, i + 1)
                out.append(repr(s[i + 1:j]))
                i = j + 1
            elif ch == ".":
                m = re.match(r"\.(and|or|not|true|false|eq|ne|lt|le|gt|ge|eqv|neqv)\.", s[i:])
                if m:
                    out.append({"and": " and ", "or": " or ", "not": " not ", "true": "True", "false": "False", "eq": "==", "ne": "!=",
                                "lt": "<", "le": "<=", "gt": ">", "ge": ">=", "eqv": "==", "neqv": "!="}[m.group(1)])
                    i += m.end()
                else:
                    m = re.match(r"\.\d+(?:[ed][+-]?\d+)?(?:_wp)?", s[i:])
                    if not m:
                        raise SyntaxError("unexpected '.' in " + s)
                    out.append(self._number(m.group(0)))
                    i += m.end()
            elif ch.isdigit():
                m = re.match(r"\d+(?:\.(?!(?:and|or|not|eq|ne|lt|le|gt|ge|eqv|neqv)\.)\d*)?(?:[ed][+-]?\d+)?(?:_wp)?", s[i:])
                out.append(self._number(m.group(0)))
                i += m.end()
            elif ch.isalpha() or ch == "_":
                m = re.match(r"[a-z_]\w*(?:%\w+)*", s[i:])
                name = m.group(0).replace("%", ".")
                i += m.end()
                j = i
                while j < n and s[j] == " ":
                    j += 1
                if j < n and s[j] == "(":
                    k = self._match_paren(s, j)
                    args = [a.strip() for a in _split_top(s[j + 1:k], ",")]
                    i = k + 1
                    if name in arrays or (name not in INTRINSICS and any(len(_split_top(a, ":")) > 1 for a in args)):
                        out.append("%s[%s]" % (name, self.index(args, arrays)))      # (a section can only be an array's)
                    elif name in INTRINSICS:
                        conv = []
                        for a in args:
                            km = re.match(r"(dim|mask|kind)\s*=\s*(?!=)(.*)$", a)
                            conv.append("%s=%s" % (km.group(1), self.expr(km.group(2), arrays)) if km else self.expr(a, arrays))
                        out.append("%s(%s)" % (INTRINSICS[name], ", ".join(conv)))
                    else:
                        out.append("%s(%s)" % (name, ", ".join(self.expr(a, arrays) for a in args if a)))
                else:
                    out.append(name)
            elif s.startswith("(/", i):                                  # array constructor (/ a, b, ... /)
                k = s.index("/)", i)
                out.append("np.array([%s])" % ", ".join(self.expr(a, arrays) for a in _split_top(s[i + 2:k], ",")))
                i = k + 2
            elif ch == "(":
                k = self._match_paren(s, i)
                out.append("(" + self.expr(s[i + 1:k], arrays) + ")")
                i = k + 1
            elif s.startswith("/=", i):
                out.append("!=")
                i += 2
            elif s.startswith("==", i) or s.startswith("<=", i) or s.startswith(">=", i) or s.startswith("**", i) or s.startswith("//", i):
                out.append("+" if s.startswith("//", i) else s[i:i + 2])
                i += 2
            elif ch == "/":
                # Fortran integer division truncates: route every division through f_div (IEEE for reals)
                left = self._pop_operand(out)
                right, i = self._next_operand(s, i + 1, arrays)
                out.append("f_div(%s, %s)" % (left, right))
            else:
                out.append(ch)
                i += 1
        return "".join(out)

    @staticmethod
    def _number(tok):
        tok = tok.replace("_wp", "")
        return tok.replace("d", "e")

    @staticmethod
    def _match_paren(s, j):
        depth, q = 0, None
        for k in range(j, len(s)):
            ch = s[k]
            if q:
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
            elif ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
                if depth == 0:
                    return k
        raise SyntaxError("unbalanced parenthesis in " + s)

    @staticmethod
    def _pop_operand(out):
        """remove and return the operand (with the * and ** chain that binds it to the left of a `/`) at the end of `out`"""
        # `/` and `*` associate left to right: a * b / c = (a * b) / c, so everything back to the previous + or - (top level) is taken
        toks = []
        while out:
            t = out[-1]
            if t.strip() in ("+", "-", ",", "==", "!=", "<", ">", "<=", ">=", "=") or t in (" and ", " or ", " not "):
                # a sign right after an operator (or at the very start) belongs to the operand: -a / b = (-a) / b, same bits
                break
            toks.append(out.pop())
        left = "".join(reversed(toks)).strip()
        if not left:
            raise SyntaxError("division without a left operand")
        return left

    def _next_operand(self, s, i, arrays):
        """the primary (with its ** chain) right of a `/`: returns (python text, next index)"""
        n = len(s)
        while i < n and s[i] == " ":
            i += 1
        start = i
        if i < n and s[i] in "+-":
            i += 1
        while True:
            while i < n and s[i] == " ":
                i += 1
            if i < n and s[i] == "(":
                i = self._match_paren(s, i) + 1
            else:
                m = re.match(r"\d+(?:\.\d*)?(?:[ed][+-]?\d+)?(?:_wp)?|\.\d+(?:[ed][+-]?\d+)?(?:_wp)?|[a-z_]\w*(?:%\w+)*", s[i:])
                if not m:
                    raise SyntaxError("operand expected in " + s[start:])
                i += m.end()
                j = i
                while j < n and s[j] == " ":
                    j += 1
                if j < n and s[j] == "(" and re.match(r"[a-z_]", m.group(0)):
                    i = self._match_paren(s, j) + 1
            j = i
            while j < n and s[j] == " ":
                j += 1
            if s.startswith("**", j):
                i = j + 2
                continue
            break
        return self.expr(s[start:i], arrays), i

    def index(self, args, arrays):
        out = []
        for a in args:
            parts = _split_top(a, ":")
            if len(parts) == 1:
                out.append("(%s)-1" % self.expr(a, arrays))
            else:
                lo = "(%s)-1" % self.expr(parts[0], arrays) if parts[0].strip() else ""
                hi = "(%s)" % self.expr(parts[1], arrays) if parts[1].strip() else ""
                if len(parts) == 3:
                    raise SyntaxError("strided sections are not supported: " + a)
                out.append("%s:%s" % (lo, hi))
        return ", ".join(out)

    # -- declarations --
    def declaration(self, st, dummies, arrays, body, ind):
        is_result_of = lambda n: n == getattr(self, "function_name", None)      # noqa: E731
        left, right = st.split("::", 1)
        attrs = [a.strip() for a in _split_top(left, ",")]
        integer = attrs[0].startswith("integer") or attrs[0].startswith("logical")
        dim = next((a for a in attrs if a.startswith("dimension")), None)
        allocatable = any(a.startswith("allocatable") for a in attrs)
        optional = any(a.startswith("optional") for a in attrs)
        for ent in _split_top(right, ","):
            ent = ent.strip()
            init = None
            if "=" in ent and not ent.startswith("("):
                ent, init = [x.strip() for x in ent.split("=", 1)]
            m = re.match(r"([a-z_]\w*)\s*(\(.*\))?$", ent)
            name, own = m.group(1), m.group(2)
            self.local_names.add(name)
            if integer:
                self.int_names.add(name)
            if optional:
                self.optional.add(name)
            shape = own or (dim[dim.index("("):] if dim else None)
            if shape:
                arrays.add(name)
                if name not in dummies and not allocatable and ":" not in shape:
                    dims = [self.expr(d, arrays) for d in _split_top(shape[1:-1], ",")]
                    body.append("%s%s = f_alloc((%s,), %s)" % (ind, name, ", ".join(dims), integer))
            elif init is not None:
                body.append("%s%s = %s" % (ind, name, self.expr(init, arrays)))
            elif name not in dummies and name not in self.module_names and not is_result_of(name):
                # an undefined local scalar: it may legally be handed to a routine that sets it (e.g. an MPI request handle)
                kind = attrs[0]
                body.append("%s%s = %s" % (ind, name, "0" if kind.startswith("integer") else "False" if kind.startswith("logical")
                                           else "''" if kind.startswith("character") else "complex(float('nan'), float('nan'))"
                                           if kind.startswith("complex") else "float('nan')"))

    def outs(self, arrays):
        """what a translated SUBROUTINE returns: its scalar dummy arguments by name (INTENT(out) scalars cannot be passed back
        through Python arguments; arrays are modified in place)"""
        return "{%s}" % ", ".join("'%s': %s, %d: %s" % (d, d, i, d) for i, d in enumerate(self.dummies) if d not in arrays)

    # -- one subroutine --
    def subroutine(self, sts):
        head = re.match(r"(subroutine|function)\s+(\w+)\s*(?:\((.*)\))?\s*$", sts[0])
        is_function = head.group(1) == "function"
        self.function_name = head.group(2) if is_function else None
        name, dummies = head.group(2), [a.strip() for a in (head.group(3) or "").split(",") if a.strip()]
        arrays = set(self.global_arrays)
        self.optional = set()
        self.dummies = dummies
        self.local_names = set(dummies)
        self.int_names = set()
        body, depth, sel = [], 1, []
        ind = lambda: "    " * depth                      # noqa: E731
        mod = sorted(self.module_names - set(dummies))
        if mod:
            body.append("    global " + ", ".join(mod))
        for st in sts[1:]:
            if re.match(r"end\s*(subroutine|function)", st):
                break
            if re.match(r"(integer|real|logical|character|complex|type\s*\()", st) and "::" in st:
                self.declaration(st, dummies, arrays, body, ind())
                continue
            if re.match(r"(use |implicit |intent|external )", st):
                continue
            depth = self.statement(st, arrays, body, depth, sel)
        if is_function:
            body.append("    return " + name)                          # the result variable carries the function's name
        else:
            body.append("    return " + self.outs(arrays))
        src = "def %s(%s):\n" % (name, ", ".join(d + "=None" if d in self.optional else d for d in dummies)) + \
            ("\n".join(body) if body else "    pass") + "\n"
        if is_function:                                                # ... and must not shadow the function inside its own body
            src = re.sub(r"\b%s\b(?!\()" % name, "_res_" + name, src).replace("def _res_%s(" % name, "def %s(" % name)
        return name, src

    def statement(self, st, arrays, body, depth, sel):
        ind = "    " * depth
        st = re.sub(r"^\d+\s+", "", st)                   # a statement label (only ever the target of an I/O ERR= here)
        m = re.match(r"do\s+(\w+)\s*=\s*(.*)$", st)
        if m:
            parts = _split_top(m.group(2), ",")
            a, b = self.expr(parts[0], arrays), self.expr(parts[1], arrays)
            if len(parts) == 3:
                c = self.expr(parts[2], arrays)
                body.append("%sfor %s in range(%s, (%s) + (1 if (%s) > 0 else -1), %s):" % (ind, m.group(1), a, b, c, c))
            else:
                body.append("%sfor %s in range(%s, (%s) + 1):" % (ind, m.group(1), a, b))
            return depth + 1
        if re.match(r"end\s*do$", st) or re.match(r"end\s*if$", st):
            if body[-1].rstrip().endswith(":"):
                body.append(ind + "pass")
            return depth - 1
        m = re.match(r"(else\s*)?if\s*\(", st)
        if m:
            j = st.index("(", m.start())
            k = self._match_paren(st, j)
            cond, rest = self.expr(st[j + 1:k], arrays), st[k + 1:].strip()
            if m.group(1):
                if body[-1].rstrip().endswith(":"):
                    body.append(ind + "pass")
                body.append("%selif %s:" % ("    " * (depth - 1), cond))
                return depth
            if rest == "then":
                body.append("%sif %s:" % (ind, cond))
                return depth + 1
            body.append("%sif %s:" % (ind, cond))
            d2 = self.statement(rest, arrays, body, depth + 1, sel)
            assert d2 == depth + 1, "one-line IF with a block statement: " + st
            return depth
        if st == "else":
            if body[-1].rstrip().endswith(":"):
                body.append(ind + "pass")
            body.append("    " * (depth - 1) + "else:")
            return depth
        m = re.match(r"select\s*case\s*\((.*)\)$", st)
        if m:
            sel.append([len(sel), True])
            body.append("%s_sel%d = %s" % (ind, sel[-1][0], self.expr(m.group(1), arrays)))
            body.append("%sif False:" % ind)                           # the CASEs hang off it as elif
            body.append("%s    pass" % ind)
            return depth + 1
        m = re.match(r"case\s*\((.*)\)$", st)
        if m:
            vals = ", ".join(self.expr(v, arrays) for v in _split_top(m.group(1), ","))
            if body[-1].rstrip().endswith(":"):
                body.append(ind + "pass")
            body.append("%selif _sel%d in (%s,):" % ("    " * (depth - 1), sel[-1][0], vals))
            return depth
        if re.match(r"case\s+default$", st):
            if body[-1].rstrip().endswith(":"):
                body.append(ind + "pass")
            body.append("    " * (depth - 1) + "else:")
            return depth
        if re.match(r"end\s*select$", st):
            if body[-1].rstrip().endswith(":"):
                body.append(ind + "pass")
            sel.pop()
            return depth - 1
        m = re.match(r"call\s+(\w+)\s*(\(.*\))?$", st)
        if m:
            args = _split_top(m.group(2)[1:-1], ",") if m.group(2) else []
            conv = []
            for pos, a in enumerate(args):
                a = a.strip()
                km = re.match(r"(\w+)\s*=\s*(?!=)(.*)$", a)
                em = re.fullmatch(r"([a-z_]\w*)\s*\((.*)\)", a)
                if pos in self.BYREF_CALLEES.get(m.group(1), ()) and em and em.group(1) in arrays and ":" not in em.group(2):
                    idx = ", ".join("(%s)-1" % self.expr(x, arrays) for x in _split_top(em.group(2), ","))
                    conv.append("f_ref(%s, (%s,))" % (em.group(1), idx))
                    continue
                conv.append("%s=%s" % (km.group(1), self.expr(km.group(2), arrays)) if km else self.expr(a, arrays))
            body.append("%s_r = %s(%s)" % (ind, m.group(1), ", ".join(conv)))
            # scalar INTENT(out / inout) dummies come back by position (see outs): copy them into the caller's own scalars
            back = [(i, a.strip()) for i, a in enumerate(args) if re.fullmatch(r"[a-z_]\w*", a.strip()) and a.strip() not in arrays
                    and (a.strip() in self.local_names or a.strip() in self.module_names)]
            if back:
                body.append("%sif isinstance(_r, dict):" % ind)
                for i, a in back:
                    body.append("%s    %s = _r.get(%d, %s)" % (ind, a, i, a))
            return depth
        m = re.match(r"allocate\s*\((.*)\)$", st)
        if m:
            for ent in _split_top(m.group(1), ","):
                em = re.match(r"\s*(\w+)\s*\((.*)\)\s*$", ent)
                if em:
                    arrays.add(em.group(1))
                    dims = [self.expr(d, arrays) for d in _split_top(em.group(2), ",")]
                    body.append("%s%s = f_alloc((%s,), %s)" % (ind, em.group(1), ", ".join(dims), em.group(1) in self.int_names or em.group(1) in self.global_int_arrays))
            return depth
        if re.match(r"(deallocate|write|print|format|namelist|rewind|read|open|close)\b", st) or st == "continue":
            body.append(ind + "pass")
            return depth
        if st == "return":
            body.append(ind + "return " + self.outs(arrays))
            return depth
        if st == "stop" or st.startswith("stop "):
            body.append(ind + "raise RuntimeError('STOP')")
            return depth
        # assignment
        parts = self._split_assignment(st)
        if not parts:
            raise SyntaxError("statement not understood: " + st)
        lhs, rhs = parts
        lm = re.match(r"([a-z_]\w*)\s*(\(.*\))?$", lhs.strip())
        if not lm:
            raise SyntaxError("left-hand side not understood: " + st)
        name, sub = lm.group(1), lm.group(2)
        r = self.expr(rhs, arrays)
        if name in arrays:
            idx = self.index([a.strip() for a in _split_top(sub[1:-1], ",")], arrays) if sub else "..."
            body.append("%s%s[%s] = %s" % (ind, name, idx, r))
        else:
            assert not sub, "subscripted scalar: " + st
            body.append("%s%s = %s" % (ind, name, r))
        return depth

    @staticmethod
    def _split_assignment(st):
        depth, q = 0, None
        for i, ch in enumerate(st):
            if q:
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
            elif ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "=" and depth == 0 and st[i + 1:i + 2] != "=" and st[i - 1] not in "=/<>":
                return st[:i], st[i + 1:]
        return None

    # -- a whole file --
    def module(self, text, defines=None, only=None, defined=()):
        """{subroutine name: python source} for the subroutines of a module file (all, or those named in `only`)"""
        sts = statements(text, defines, defined)
        self.scan_module_level(sts)
        out, i = {}, 0
        while i < len(sts):
            if re.match(r"(subroutine|function)\s+\w+", sts[i]):
                j = next(k for k in range(i, len(sts)) if re.match(r"end\s*(subroutine|function)", sts[k]))
                nm = re.match(r"(?:subroutine|function)\s+(\w+)", sts[i]).group(1)
                if only is None or nm in only:
                    name, src = self.subroutine(sts[i:j + 1])
                    out[name] = src
                i = j + 1
            else:
                i += 1
        return out


def module_parameters(text, namespace, defines=None, defined=()):
    """scalar module variables declared with an initialiser before CONTAINS (e.g. `REAL(wp) :: r1_6 = 1._wp / 6._wp`) -> namespace"""
    tr = Translator()
    for k, v in RUNTIME.items():
        namespace.setdefault(k, v)
    for st in statements(text, defines, defined):
        if st == "contains" or re.match(r"(subroutine|function)\s+\w+", st):
            break
        if re.match(r"(integer|real|logical)", st) and "::" in st:
            for ent in _split_top(st.split("::", 1)[1], ","):
                if "=" in ent and "(" not in ent.split("=", 1)[0]:
                    name, init = [x.strip() for x in ent.split("=", 1)]
                    namespace[name] = eval(tr.expr(init, set()), namespace)


def load(text, namespace, arrays=(), int_arrays=(), defines=None, only=None, defined=(), module_vars=()):
    """translate the subroutines of `text` and define them in `namespace` (which must hold the module variables and the
    external routines they use); returns {name: python source}"""
    tr = Translator(arrays, int_arrays)
    tr.module_names.update(v.lower() for v in module_vars)             # variables of USEd modules the routines assign (e.g. r2dt of dom_oce)
    srcs = tr.module(text, defines, only, defined)
    for k, v in RUNTIME.items():
        namespace.setdefault(k, v)
    for name, src in srcs.items():
        exec(compile(src, "<f90exec:%s>" % name, "exec"), namespace)
    return srcs
