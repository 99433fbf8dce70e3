"""
Python energy functions on the GPU: trace -> differentiate -> CUDA source.

The reference accepts ANY Python callable as an energy and differentiates it numerically on the host, 2 * dim
callbacks per Langevin step (tsu/core.py:82-98, 100-162).  A Python callable cannot run inside a kernel, and this
package has no CPU fallback.  What can be done without one: call the function ONCE on symbolic inputs, record the
arithmetic it performs as an expression graph, differentiate the graph analytically (reverse mode) and emit the
gradient as a CUDA device function that NVRTC compiles into the fused Langevin kernel (csrc/langevin_body.cuh).

    energy_fn(x)  with  x = object ndarray of Sym      ->  Sym (the energy)  ->  grad  ->  source text

Everything NumPy does through Python's operators or through the object-dtype loops of its ufuncs works unchanged:
    + - * / ** unary minus, abs, np.sum / mean / dot / @, np.exp / log / sqrt / sin / cos / tanh / square / power,
    indexing, slicing, broadcasting against constant arrays, Python loops over centres, closures over data.
What cannot be traced raises TraceError (the caller turns it into SamplingError): branches on the value of x
(`if x[0] > 0`, np.maximum, np.where, np.abs(x) > c), float(...) / int(...) of a traced value, calls into compiled code.
The traced graph is checked numerically against the callable before it is used.
"""

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np


class TraceError(Exception):
    pass


class Sym:
    """node of the expression graph (interned per Tracer: equal sub-expressions are one node)"""

    __slots__ = ("tr", "op", "args", "val", "idx")
    __array_priority__ = 1000.0

    def __init__(self, tr, op, args=(), val=None):
        self.tr, self.op, self.args, self.val = tr, op, args, val
        self.idx = len(tr.nodes)
        tr.nodes.append(self)

    # -- arithmetic ---------------------------------------------------------------------------------------
    def _lift(self, other):
        if isinstance(other, Sym):
            return other
        if isinstance(other, (int, float, np.integer, np.floating)):
            return self.tr.const(float(other))
        if isinstance(other, np.ndarray) and other.ndim == 0:
            return self._lift(other.item())
        return None

    def _bin(self, op, other, swap=False):
        if isinstance(other, np.ndarray) and other.ndim > 0:  # broadcast over a constant / symbolic array
            f = (lambda a: self.tr.binary(op, self._lift(a), self)) if swap else (lambda a: self.tr.binary(op, self, self._lift(a)))
            out = np.empty(other.shape, dtype=object)
            for i, a in np.ndenumerate(other):
                out[i] = f(a)
            return out
        o = self._lift(other)
        if o is None:
            return NotImplemented
        return self.tr.binary(op, o, self) if swap else self.tr.binary(op, self, o)

    def __add__(self, o): return self._bin("add", o)
    def __radd__(self, o): return self._bin("add", o, True)
    def __sub__(self, o): return self._bin("sub", o)
    def __rsub__(self, o): return self._bin("sub", o, True)
    def __mul__(self, o): return self._bin("mul", o)
    def __rmul__(self, o): return self._bin("mul", o, True)
    def __truediv__(self, o): return self._bin("div", o)
    def __rtruediv__(self, o): return self._bin("div", o, True)
    def __pow__(self, o): return self._bin("pow", o)
    def __rpow__(self, o): return self._bin("pow", o, True)
    def __neg__(self): return self.tr.unary("neg", self)
    def __pos__(self): return self
    def __abs__(self): return self.tr.unary("abs", self)

    # object-dtype ufunc loops call these methods (np.exp(obj_array) -> element.exp())
    def exp(self): return self.tr.unary("exp", self)
    def log(self): return self.tr.unary("log", self)
    def sqrt(self): return self.tr.unary("sqrt", self)
    def sin(self): return self.tr.unary("sin", self)
    def cos(self): return self.tr.unary("cos", self)
    def tanh(self): return self.tr.unary("tanh", self)
    def square(self): return self.tr.binary("mul", self, self)
    def conjugate(self): return self
    conj = conjugate

    @property
    def real(self): return self

    @property
    def imag(self): return self.tr.const(0.0)

    # -- what cannot be traced ------------------------------------------------------------------------------
    def _no(self, what):
        raise TraceError(f"the energy function {what} of a traced value: its control flow or result depends on x in a way "
                         "that cannot be recorded")

    def __bool__(self): self._no("takes the truth value")
    def __float__(self): self._no("takes float()")
    def __int__(self): self._no("takes int()")
    def __lt__(self, o): self._no("compares (<)")
    def __le__(self, o): self._no("compares (<=)")
    def __gt__(self, o): self._no("compares (>)")
    def __ge__(self, o): self._no("compares (>=)")
    __hash__ = object.__hash__


class Tracer:
    def __init__(self, dim: int):
        self.dim = dim
        self.nodes: List[Sym] = []
        self.intern: Dict[tuple, Sym] = {}
        self.inputs = [self._mk("x", (), i) for i in range(dim)]

    def _mk(self, op, args, val=None):
        key = (op, tuple(a.idx for a in args), val)
        n = self.intern.get(key)
        if n is None:
            n = Sym(self, op, tuple(args), val)
            self.intern[key] = n
        return n

    def const(self, v: float) -> Sym:
        v = float(v)
        if not math.isfinite(v):
            raise TraceError("a non-finite constant appears in the energy")
        return self._mk("const", (), v)

    def unary(self, op, a: Sym) -> Sym:
        if a.op == "const":
            f = {"neg": lambda t: -t, "abs": abs, "exp": math.exp, "log": math.log, "sqrt": math.sqrt, "sin": math.sin,
                 "cos": math.cos, "tanh": math.tanh, "sign": lambda t: (t > 0) - (t < 0)}[op]
            return self.const(f(a.val))
        if op == "neg" and a.op == "neg":
            return a.args[0]
        return self._mk(op, (a,))

    def binary(self, op, a: Sym, b: Sym) -> Sym:
        if a is None or b is None:
            raise TraceError("unsupported operand in the energy function")
        ca, cb = a.op == "const", b.op == "const"
        if ca and cb:
            f = {"add": lambda s, t: s + t, "sub": lambda s, t: s - t, "mul": lambda s, t: s * t, "div": lambda s, t: s / t,
                 "pow": lambda s, t: s ** t}[op]
            return self.const(f(a.val, b.val))
        if op == "add":
            if ca and a.val == 0.0: return b
            if cb and b.val == 0.0: return a
        elif op == "sub":
            if cb and b.val == 0.0: return a
            if ca and a.val == 0.0: return self.unary("neg", b)
        elif op == "mul":
            if (ca and a.val == 0.0) or (cb and b.val == 0.0): return self.const(0.0)
            if ca and a.val == 1.0: return b
            if cb and b.val == 1.0: return a
            if ca and a.val == -1.0: return self.unary("neg", b)
            if cb and b.val == -1.0: return self.unary("neg", a)
        elif op == "div":
            if cb and b.val == 1.0: return a
            if cb: return self.binary("mul", a, self.const(1.0 / b.val))
        elif op == "pow":
            if cb and b.val == 1.0: return a
            if cb and b.val == 0.0: return self.const(1.0)
            if cb and b.val == 2.0: return self.binary("mul", a, a)
            if cb and b.val == 3.0: return self.binary("mul", self.binary("mul", a, a), a)
            if cb and b.val == 0.5: return self.unary("sqrt", a)
        if op in ("add", "mul") and a.idx > b.idx:
            a, b = b, a  # canonical order of commutative operands
        return self._mk(op, (a, b))

    # -- reverse-mode differentiation ------------------------------------------------------------------------
    def gradient(self, out: Sym) -> List[Sym]:
        adj: Dict[int, Sym] = {out.idx: self.const(1.0)}
        order = self._reachable(out)
        add, mul = (lambda s, t: self.binary("add", s, t)), (lambda s, t: self.binary("mul", s, t))

        def acc(node, v):
            adj[node.idx] = add(adj[node.idx], v) if node.idx in adj else v

        for n in reversed(order):
            g = adj.get(n.idx)
            if g is None or n.op in ("x", "const"):
                continue
            a = n.args[0]
            b = n.args[1] if len(n.args) > 1 else None
            if n.op == "add":
                acc(a, g); acc(b, g)
            elif n.op == "sub":
                acc(a, g); acc(b, self.unary("neg", g))
            elif n.op == "mul":
                acc(a, mul(g, b)); acc(b, mul(g, a))
            elif n.op == "div":
                acc(a, self.binary("div", g, b))
                acc(b, self.unary("neg", self.binary("div", mul(g, n), b)))
            elif n.op == "neg":
                acc(a, self.unary("neg", g))
            elif n.op == "pow":
                if b.op == "const":
                    acc(a, mul(g, mul(self.const(b.val), self.binary("pow", a, self.const(b.val - 1.0)))))
                else:
                    acc(a, mul(g, mul(b, self.binary("pow", a, self.binary("sub", b, self.const(1.0))))))
                    acc(b, mul(g, mul(n, self.unary("log", a))))
            elif n.op == "exp":
                acc(a, mul(g, n))
            elif n.op == "log":
                acc(a, self.binary("div", g, a))
            elif n.op == "sqrt":
                acc(a, self.binary("div", mul(g, self.const(0.5)), n))
            elif n.op == "sin":
                acc(a, mul(g, self.unary("cos", a)))
            elif n.op == "cos":
                acc(a, self.unary("neg", mul(g, self.unary("sin", a))))
            elif n.op == "tanh":
                acc(a, mul(g, self.binary("sub", self.const(1.0), mul(n, n))))
            elif n.op == "abs":
                acc(a, mul(g, self.unary("sign", a)))
            elif n.op == "sign":
                pass
            else:
                raise TraceError(f"no derivative rule for {n.op}")
        zero = self.const(0.0)
        return [adj.get(x.idx, zero) for x in self.inputs]

    def _reachable(self, *outs: Sym) -> List[Sym]:
        seen, order = set(), []
        stack = [(o, False) for o in outs]
        while stack:
            n, done = stack.pop()
            if done:
                order.append(n)
                continue
            if n.idx in seen:
                continue
            seen.add(n.idx)
            stack.append((n, True))
            for a in n.args:
                if a.idx not in seen:
                    stack.append((a, False))
        return order

    # -- evaluation (numerical check of the trace) and code generation -----------------------------------------
    def evaluate(self, outs: Sequence[Sym], x: np.ndarray) -> List[float]:
        val: Dict[int, float] = {}
        for n in self._reachable(*outs):
            a = [val[t.idx] for t in n.args]
            if n.op == "x": v = float(x[n.val])
            elif n.op == "const": v = n.val
            elif n.op == "add": v = a[0] + a[1]
            elif n.op == "sub": v = a[0] - a[1]
            elif n.op == "mul": v = a[0] * a[1]
            elif n.op == "div": v = a[0] / a[1]
            elif n.op == "pow": v = a[0] ** a[1]
            elif n.op == "neg": v = -a[0]
            elif n.op == "abs": v = abs(a[0])
            elif n.op == "sign": v = float((a[0] > 0) - (a[0] < 0))
            else: v = getattr(math, n.op)(a[0])
            val[n.idx] = v
        return [val[o.idx] for o in outs]

    def cuda_source(self, grads: Sequence[Sym]) -> str:
        """`template <typename real> __device__ void tsu_user_grad(const real* x, real* g)`: straight-line code"""
        lines = []
        name: Dict[int, str] = {}
        for n in self._reachable(*grads):
            a = [name[t.idx] for t in n.args]
            if n.op == "x":
                name[n.idx] = f"x[{n.val}]"
                continue
            if n.op == "const":
                name[n.idx] = f"(real){n.val!r}"
                continue
            v = f"v{n.idx}"
            if n.op in ("add", "sub", "mul", "div"):
                e = f"{a[0]} {'+-*/'[('add', 'sub', 'mul', 'div').index(n.op)]} {a[1]}"
            elif n.op == "pow": e = f"pow({a[0]}, (real){a[1]})" if not a[1].startswith("(real)") else f"pow({a[0]}, {a[1]})"
            elif n.op == "neg": e = f"-{a[0]}"
            elif n.op == "abs": e = f"fabs({a[0]})"
            elif n.op == "sign": e = f"(real)(({a[0]} > (real)0) - ({a[0]} < (real)0))"
            else: e = f"{n.op}({a[0]})"
            lines.append(f"  const real {v} = {e};")
            name[n.idx] = v
        for i, gnode in enumerate(grads):
            lines.append(f"  g[{i}] = {name[gnode.idx]};")
        return ("template <typename real>\n__device__ __forceinline__ void tsu_user_grad(const real* x, real* g) {\n"
                + "\n".join(lines) + "\n}\n")


def _scalar(v):
    """the Sym a traced function returned (possibly wrapped in a 0-d / 1-element object array)"""
    if isinstance(v, np.ndarray):
        if v.size != 1:
            raise TraceError(f"the energy function returned an array of shape {v.shape}, not a scalar")
        v = v.reshape(-1)[0]
    return v


def trace_energy(energy_fn: Callable, dim: int, probe_points: Optional[np.ndarray] = None, rtol: float = 1e-9):
    """(Tracer, energy node, gradient nodes) of energy_fn on a dim-vector, or TraceError.

    The traced graph is evaluated at `probe_points` and compared with the callable itself: a function whose result
    does not follow from the recorded arithmetic (hidden state, data-dependent branches that happened not to raise) is
    rejected."""
    tr = Tracer(dim)
    x = np.empty(dim, dtype=object)
    for i, s in enumerate(tr.inputs):
        x[i] = s
    try:
        out = _scalar(energy_fn(x))
    except TraceError:
        raise
    except Exception as exc:  # NumPy raising on object arrays, attribute errors of unsupported functions, ...
        raise TraceError(f"the energy function could not be evaluated on symbolic inputs ({type(exc).__name__}: {exc})")
    if isinstance(out, (int, float, np.integer, np.floating)):
        out = tr.const(float(out))
    if not isinstance(out, Sym):
        raise TraceError(f"the energy function returned {type(out).__name__} on symbolic inputs")
    if probe_points is not None:
        checked = 0
        for p in np.atleast_2d(probe_points):
            try:
                want = float(energy_fn(np.asarray(p, dtype=np.float64)))
                got = tr.evaluate([out], p)[0]
            except (ValueError, OverflowError, ZeroDivisionError, FloatingPointError):
                continue  # outside the function's domain (log of a negative number, ...): nothing to compare here
            if not (math.isfinite(want) and math.isfinite(got)):
                continue
            if not (abs(got - want) <= rtol * max(1.0, abs(want))):
                raise TraceError(f"the traced expression gives {got!r} where the function gives {want!r}")
            checked += 1
        if checked == 0:
            raise TraceError("the energy function is not finite at any of the probe points around x_init")
    grads = tr.gradient(out)
    return tr, out, grads
