"""The few UFL constructs the reference's scripts use *outside* ElasticLF4 -- the L2-projection error norm of
tests/eigenmode/eigenmode_2d.py:49-63 / eigenmode_3d.py:53-67:

    HU = VectorFunctionSpace(mesh, "DG", 6); temp = Function(HU)
    G = inner(TestFunction(HU), TrialFunction(HU))*dx - inner(TestFunction(HU), abs(u1 - uexact))*dx
    solve(lhs(G) == rhs(G), temp);  u_error = norm(temp)

This is not a form compiler.  Expressions are kept as small trees; ``solve`` recognises exactly one problem -- a mass
matrix on a DG space against ``inner(test, g)*dx`` -- and carries it out cell by cell (the mass matrix of a DG space
is block diagonal, so the "solve" is an L2 projection of ``g`` onto each cell's polynomials), with a collapsed Gauss
rule of degree ``deg(test) + deg(g)`` as UFL's degree estimation would choose.  Host-side NumPy, post-processing only;
anything else raises ``NotImplementedError`` rather than guessing.
"""
from __future__ import annotations

import numpy as np

__all__ = ["TestFunction", "TrialFunction", "inner", "dx", "lhs", "rhs", "solve", "FieldExpr"]


class FieldExpr:
    """A pointwise expression of Functions: f, a +- b, c * a, abs(a)."""

    def __init__(self, op, args):
        self.op, self.args = op, args

    # -- algebra
    def __sub__(self, other):
        return FieldExpr("sum", (as_field(self), as_field(other), -1.0))

    def __add__(self, other):
        return FieldExpr("sum", (as_field(self), as_field(other), 1.0))

    def __rsub__(self, other):
        return FieldExpr("sum", (as_field(other), as_field(self), -1.0))

    __radd__ = __add__

    def __neg__(self):
        return FieldExpr("scale", (-1.0, self))

    def __mul__(self, c):
        if not np.isscalar(c):
            return NotImplemented
        return FieldExpr("scale", (float(c), self))

    __rmul__ = __mul__

    def __abs__(self):
        return FieldExpr("abs", (self,))

    # -- evaluation
    def functions(self):
        if self.op == "function":
            return [self.args[0]]
        out = []
        for a in self.args:
            if isinstance(a, FieldExpr):
                out += a.functions()
        return out

    def degree(self):
        return max(f.function_space().degree for f in self.functions())

    def shape(self):
        return self.functions()[0].function_space().shape

    def evaluate(self, xq):
        """Values at reference points xq (nq, d) of every owned cell: (E, nq) + shape."""
        if self.op == "function":
            f = self.args[0]
            fs = f.function_space()
            phi = fs.elem.tabulate(xq)                                         # (nq, nd)
            vals = f.dat.data.reshape((fs.plan.n_owned, fs.elem.nd) + fs.shape)
            return np.einsum("qb,eb...->eq...", phi, vals)
        if self.op == "sum":
            a, b, sign = self.args
            return a.evaluate(xq) + sign * b.evaluate(xq)
        if self.op == "scale":
            return self.args[0] * self.args[1].evaluate(xq)
        if self.op == "abs":
            return np.abs(self.args[0].evaluate(xq))
        raise NotImplementedError(self.op)


def as_field(x):
    from .compat import Function
    if isinstance(x, FieldExpr):
        return x
    if isinstance(x, Function):
        return FieldExpr("function", (x,))
    raise NotImplementedError(f"cannot use {type(x).__name__} in a pointwise expression")


class Argument:
    def __init__(self, space, number):
        self.space, self.number = space, number

    def function_space(self):
        return self.space


def TestFunction(space):
    return Argument(space, 0)


def TrialFunction(space):
    return Argument(space, 1)


class Inner:
    def __init__(self, a, b):
        self.a, self.b = a, b

    def __mul__(self, measure):
        if not isinstance(measure, Measure):
            return NotImplemented
        if measure.kind != "dx":
            raise NotImplementedError("only cell integrals (dx) are supported here")
        return Form([(1.0, self)])


def inner(a, b):
    from .compat import Function
    wrap = lambda x: as_field(x) if isinstance(x, (Function, FieldExpr)) else x
    return Inner(wrap(a), wrap(b))


class Measure:
    def __init__(self, kind):
        self.kind = kind


dx = Measure("dx")


class Form:
    def __init__(self, terms):
        self.terms = list(terms)                       # [(sign, Inner)]

    def __add__(self, other):
        return Form(self.terms + other.terms)

    def __sub__(self, other):
        return Form(self.terms + [(-s, t) for s, t in other.terms])

    def __neg__(self):
        return Form([(-s, t) for s, t in self.terms])

    def __eq__(self, other):
        return Equation(self, other)

    __hash__ = None

    @staticmethod
    def _arity(term):
        return sum(isinstance(x, Argument) for x in (term.a, term.b))


class Equation:
    def __init__(self, a, L):
        self.a, self.L = a, L


def lhs(form):
    """The bilinear part (test and trial function)."""
    return Form([(s, t) for s, t in form.terms if Form._arity(t) == 2])


def rhs(form):
    """Minus the linear part, as ufl.rhs."""
    return Form([(-s, t) for s, t in form.terms if Form._arity(t) == 1])


def solve(equation, u, **kwargs):
    """``solve(lhs(G) == rhs(G), temp)`` for G = inner(test, trial)*dx - inner(test, g)*dx on a DG space."""
    from .compat import _quadrature
    if not isinstance(equation, Equation):
        raise NotImplementedError("solve(a == L, u) expected")
    a, L = equation.a, equation.L
    space = u.function_space()
    ok = (len(a.terms) == 1 and a.terms[0][0] > 0 and isinstance(a.terms[0][1].a, Argument)
          and isinstance(a.terms[0][1].b, Argument)
          and {a.terms[0][1].a.number, a.terms[0][1].b.number} == {0, 1}
          and a.terms[0][1].a.space is space and a.terms[0][1].b.space is space)
    if not ok:
        raise NotImplementedError("only the mass matrix inner(test, trial)*dx of the solution's own DG space is supported")
    scale = 1.0 / a.terms[0][0]
    el = space.elem
    d = space.mesh().dim
    E = space.plan.n_owned
    out = np.zeros((E, el.nd) + space.shape)
    for sign, term in L.terms:
        v, g = (term.a, term.b) if isinstance(term.a, Argument) else (term.b, term.a)
        if not (isinstance(v, Argument) and v.number == 0 and v.space is space and isinstance(g, FieldExpr)):
            raise NotImplementedError("right-hand side must be inner(test, <expression of Functions>)*dx")
        if tuple(g.shape()) != tuple(space.shape):
            raise ValueError("inner(): shapes of the test function and the expression differ")
        xq, wq = _quadrature(d, el.degree + g.degree())
        phi = el.tabulate(xq)                                                  # (nq, nd)
        gq = g.evaluate(xq)                                                    # (E, nq) + shape
        out += sign * scale * np.einsum("q,qa,eq...->ea...", wq, phi, gq)
    # block-diagonal mass: |detJ| cancels between the two sides; the reference element's measure is 1/d!
    u.dat.data[...] = np.einsum("ab,eb...->ea...", el.Minv, out).reshape(u.dat.data.shape)
    return u
