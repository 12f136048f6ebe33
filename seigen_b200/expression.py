"""C-string ``Expression`` objects, evaluated with NumPy at nodal coordinates.

The reference scripts describe initial conditions, sources and sponges as Firedrake (2017) ``Expression``s:
C snippets over ``x[0..d)`` with keyword parameters that stay mutable attributes, e.g.
``Expression((("x[0] >= 44.5 && x[0] <= 45.5 ? (-1.0 + 2*a*pow(t - 0.3, 2))*exp(-a*pow(t - 0.3, 2)) : 0.0", "0.0"), ...), a=159.42, t=0.0)``
(``tests/explosive_source/explosive_source_lf4.py:36-38``), whose ``.t`` is updated every step
(``seigen/elastic.py:287``).  Firedrake compiles such strings to C; here a small recursive-descent parser
turns each one into a closure over NumPy arrays (no ``eval`` of user text).
"""
from __future__ import annotations

import math
import re

import numpy as np

__all__ = ["Expression", "ExpressionSyntaxError"]


class ExpressionSyntaxError(ValueError):
    pass


_TOKEN = re.compile(r"\s*(?:(\d+\.\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|\d+(?:[eE][-+]?\d+)?)|([A-Za-z_]\w*)|"
                    r"(&&|\|\||<=|>=|==|!=|[-+*/()<>?:,\[\]!]))")

_FUNCS = {
    "sin": np.sin, "cos": np.cos, "tan": np.tan, "asin": np.arcsin, "acos": np.arccos, "atan": np.arctan,
    "sinh": np.sinh, "cosh": np.cosh, "tanh": np.tanh, "exp": np.exp, "log": np.log, "sqrt": np.sqrt,
    "fabs": np.abs, "abs": np.abs, "floor": np.floor, "ceil": np.ceil,
    "pow": np.power, "atan2": np.arctan2, "fmin": np.minimum, "fmax": np.maximum, "min": np.minimum,
    "max": np.maximum,
}
_CONSTS = {"pi": math.pi, "M_PI": math.pi, "e": math.e, "DOLFIN_PI": math.pi}


def _tokenise(text):
    pos, out = 0, []
    text = text.strip()
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m or m.end() == pos:
            raise ExpressionSyntaxError(f"cannot parse {text!r} at position {pos}")
        num, ident, op = m.groups()
        if num is not None:
            out.append(("num", float(num)))
        elif ident is not None:
            out.append(("id", ident))
        else:
            out.append(("op", op))
        pos = m.end()
    out.append(("end", None))
    return out


class _Parser:
    """expr := or ('?' expr ':' expr)? ; usual C precedence below.  Produces closures f(env) -> ndarray|float."""

    def __init__(self, text):
        self.text = text
        self.toks = _tokenise(text)
        self.i = 0

    def peek(self):
        return self.toks[self.i]

    def take(self, kind=None, val=None):
        t = self.toks[self.i]
        if (kind and t[0] != kind) or (val is not None and t[1] != val):
            raise ExpressionSyntaxError(f"unexpected {t[1]!r} in {self.text!r}")
        self.i += 1
        return t

    def accept(self, val):
        t = self.toks[self.i]
        if t[0] == "op" and t[1] == val:
            self.i += 1
            return True
        return False

    def parse(self):
        f = self.ternary()
        self.take("end")
        return f

    def ternary(self):
        c = self.lor()
        if self.accept("?"):
            a = self.ternary()
            self.take("op", ":")
            b = self.ternary()
            return lambda env: np.where(np.asarray(c(env)) != 0, a(env), b(env))
        return c

    def lor(self):
        a = self.land()
        while self.accept("||"):
            b = self.land()
            a = (lambda a, b: lambda env: (np.asarray(a(env)) != 0) | (np.asarray(b(env)) != 0))(a, b)
        return a

    def land(self):
        a = self.cmp()
        while self.accept("&&"):
            b = self.cmp()
            a = (lambda a, b: lambda env: (np.asarray(a(env)) != 0) & (np.asarray(b(env)) != 0))(a, b)
        return a

    def cmp(self):
        a = self.add()
        ops = {"<": np.less, ">": np.greater, "<=": np.less_equal, ">=": np.greater_equal,
               "==": np.equal, "!=": np.not_equal}
        while self.peek()[0] == "op" and self.peek()[1] in ops:
            op = ops[self.take()[1]]
            b = self.add()
            a = (lambda a, b, op: lambda env: op(a(env), b(env)))(a, b, op)
        return a

    def add(self):
        a = self.mul()
        while self.peek()[0] == "op" and self.peek()[1] in "+-":
            o = self.take()[1]
            b = self.mul()
            if o == "+":
                a = (lambda a, b: lambda env: a(env) + b(env))(a, b)
            else:
                a = (lambda a, b: lambda env: a(env) - b(env))(a, b)
        return a

    def mul(self):
        a = self.unary()
        while self.peek()[0] == "op" and self.peek()[1] in "*/":
            o = self.take()[1]
            b = self.unary()
            if o == "*":
                a = (lambda a, b: lambda env: a(env) * b(env))(a, b)
            else:
                a = (lambda a, b: lambda env: a(env) / b(env))(a, b)
        return a

    def unary(self):
        if self.accept("-"):
            a = self.unary()
            return lambda env: -a(env)
        if self.accept("+"):
            return self.unary()
        if self.accept("!"):
            a = self.unary()
            return lambda env: np.asarray(a(env)) == 0
        return self.primary()

    def primary(self):
        kind, val = self.peek()
        if kind == "num":
            self.take()
            return lambda env, v=val: v
        if kind == "op" and val == "(":
            self.take()
            a = self.ternary()
            self.take("op", ")")
            return a
        if kind == "id":
            self.take()
            if self.accept("("):
                args = []
                if not self.accept(")"):
                    args.append(self.ternary())
                    while self.accept(","):
                        args.append(self.ternary())
                    self.take("op", ")")
                if val not in _FUNCS:
                    raise ExpressionSyntaxError(f"unknown function {val!r} in {self.text!r}")
                fn = _FUNCS[val]
                return lambda env, fn=fn, args=tuple(args): fn(*[np.asarray(a(env), dtype=float) for a in args])
            if val == "x" and self.accept("["):
                idx = self.take("num")[1]
                self.take("op", "]")
                k = int(idx)
                return lambda env, k=k: env["x"][..., k]
            return lambda env, name=val: env["params"][name] if name in env["params"] else _CONSTS[name]
        raise ExpressionSyntaxError(f"unexpected {val!r} in {self.text!r}")


def _compile_tree(code):
    if isinstance(code, (tuple, list)):
        return tuple(_compile_tree(c) for c in code)
    if isinstance(code, (int, float)):
        return (lambda env, v=float(code): v)
    return _Parser(str(code)).parse()


def _shape_of(tree):
    if isinstance(tree, tuple):
        inner = _shape_of(tree[0])
        return (len(tree),) + inner
    return ()


class Expression:
    """``Expression(code, **params)``: ``code`` a C string, a tuple of them (vector) or a tuple of tuples (tensor).

    Keyword parameters become mutable attributes (``expr.t = 0.3``), as with the Firedrake original."""

    def __init__(self, code=None, **kwargs):
        if code is None:
            raise ValueError("Expression needs code")
        object.__setattr__(self, "_params", {k: float(v) for k, v in kwargs.items()})
        object.__setattr__(self, "code", code)
        object.__setattr__(self, "_tree", _compile_tree(code))
        object.__setattr__(self, "_shape", _shape_of(self._tree))
        # fail early on unknown identifiers
        self.evaluate(np.zeros((1, 3)))

    def __getattr__(self, name):
        params = object.__getattribute__(self, "_params")
        if name in params:
            return params[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in self._params:
            self._params[name] = float(value)
        else:
            object.__setattr__(self, name, value)

    def value_shape(self):
        return self._shape

    def code_key(self):
        """Hashable form of the source text (cache key for precomputed tables)."""
        def flat(c):
            return tuple(flat(k) for k in c) if isinstance(c, (tuple, list)) else str(c)
        return flat(self.code)

    @property
    def user_parameters(self):
        return dict(self._params)

    def evaluate(self, x, **override):
        """Values at points ``x`` (n, d) -> (n,) + value_shape, float64."""
        x = np.asarray(x, dtype=np.float64)
        params = dict(self._params)
        params.update(override)
        env = {"x": x, "params": params}
        n = x.shape[0]

        def ev(tree):
            if isinstance(tree, tuple):
                return np.stack([ev(t) for t in tree], axis=1)
            try:
                v = tree(env)
            except KeyError as exc:
                raise ExpressionSyntaxError(f"unknown identifier {exc.args[0]!r} in expression") from None
            return np.broadcast_to(np.asarray(v, dtype=np.float64), (n,)).copy()

        return ev(self._tree)
