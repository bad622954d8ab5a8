#!/usr/bin/env python
"""Generate the Gauss-Kronrod-Patterson 10/21/43/87-point tables used by QUADPACK's QNG
(and hence GSL's gsl_integration_qng, which pplacer calls at pplacer_src/integration.ml:27-31).

GSL is an un-vendored dependency of the reference, so the
tables are regenerated here from their mathematical definition with mpmath:
  * the 10-point Gauss-Legendre rule,
  * its Kronrod extension to 21 points,
  * the Patterson extensions to 43 and 87 points.
An extension of a rule with nodes {x_i}, i=1..n, adds the n+1 roots of the polynomial E of degree
n+1 with  integral_{-1}^{1} P(x) E(x) x^k dx = 0  for k = 0..n  (P = prod (x - x_i)); weights are
the interpolatory weights on the union of nodes.

Array naming and ordering follow QUADPACK qng.f / GSL integration/qng.h (positive abscissae,
descending):
  x1[5]  abscissae common to the 10-, 21-, 43-, 87-point rules      w10[5]
  x2[5]  abscissae added by the 21-point rule                        w21a[5] (x1), w21b[6] (x2, centre)
  x3[11] abscissae added by the 43-point rule                        w43a[10] (x1,x2), w43b[12] (x3, centre)
  x4[22] abscissae added by the 87-point rule                        w87a[21] (x1,x2,x3), w87b[23] (x4, centre)

Writes the same header to oracle/qng_tables.h (oracle) and pplacer_b200/csrc/qng_tables.h (product).
"""
import os
import sys
from mpmath import mp, mpf, legendre, findroot, polyval

mp.dps = 120


def poly_mul(a, b):
    """coefficients lowest degree first"""
    out = [mpf(0)] * (len(a) + len(b) - 1)
    for i, ai in enumerate(a):
        if ai == 0:
            continue
        for j, bj in enumerate(b):
            out[i + j] += ai * bj
    return out


def poly_from_roots(roots):
    p = [mpf(1)]
    for r in roots:
        p = poly_mul(p, [-r, mpf(1)])
    return p


def poly_int_m1_1(p):
    """integral over [-1,1] of sum p[k] x^k"""
    s = mpf(0)
    for k, c in enumerate(p):
        if k % 2 == 0:
            s += 2 * c / (k + 1)
    return s


def poly_eval(p, x):
    return polyval(list(reversed(p)), x)


def poly_deriv(p):
    return [k * c for k, c in enumerate(p)][1:]


def poly_div_linear(p, r):
    """p(x) / (x - r) assuming r is a root; synthetic division; lowest-first in/out"""
    n = len(p) - 1
    q = [mpf(0)] * n
    carry = mpf(0)
    for k in range(n, 0, -1):
        carry = p[k] + carry * r
        q[k - 1] = carry
    return q


def extend(nodes):
    """Given all nodes of the current symmetric rule, return the n+1 new nodes."""
    n = len(nodes)
    P = poly_from_roots(nodes)
    m = n + 1
    # moments mu[j] = int P x^j, j = 0..(m+n)
    mu = [poly_int_m1_1(poly_mul(P, [mpf(0)] * j + [mpf(1)])) for j in range(m + n + 1)]
    # E = x^m + sum_{j<m} c_j x^j ;  sum_j c_j mu[j+k] = -mu[m+k], k = 0..n
    A = mp.matrix(m, m)
    b = mp.matrix(m, 1)
    for k in range(m):
        for j in range(m):
            A[k, j] = mu[j + k]
        b[k] = -mu[m + k]
    # parity: P has parity n, E must have parity m => half the unknowns vanish and half the
    # equations are 0 = 0; solve the reduced system to keep it non-singular.
    idx_unknown = [j for j in range(m) if (m - j) % 2 == 0]
    idx_eq = [k for k in range(m) if (n + m + k) % 2 == 0][: len(idx_unknown)]
    A2 = mp.matrix(len(idx_eq), len(idx_unknown))
    b2 = mp.matrix(len(idx_eq), 1)
    for a, k in enumerate(idx_eq):
        for c, j in enumerate(idx_unknown):
            A2[a, c] = A[k, j]
        b2[a] = b[k]
    sol = mp.lu_solve(A2, b2)
    E = [mpf(0)] * (m + 1)
    E[m] = mpf(1)
    for c, j in enumerate(idx_unknown):
        E[j] = sol[c]
    # roots interlace with existing nodes (and the end points)
    pts = sorted([mpf(-1)] + list(nodes) + [mpf(1)])
    new = []
    for lo, hi in zip(pts[:-1], pts[1:]):
        flo, fhi = poly_eval(E, lo), poly_eval(E, hi)
        assert flo * fhi < 0, "no sign change: extension roots do not interlace"
        r = findroot(lambda x: poly_eval(E, x), (lo, hi), solver="anderson", tol=mpf(10) ** (-100))
        new.append(r)
    assert len(new) == m
    return new


def weights(nodes):
    N = poly_from_roots(nodes)
    dN = poly_deriv(N)
    ws = []
    for x in nodes:
        q = poly_div_linear(N, x)
        ws.append(poly_int_m1_1(q) / poly_eval(dN, x))
    return ws


def main():
    # 10-point Gauss-Legendre
    import numpy as np
    g10 = []
    for guess in np.polynomial.legendre.leggauss(10)[0]:
        g10.append(findroot(lambda x: legendre(10, x), mpf(float(guess)), tol=mpf(10) ** (-100)))
    assert all(abs(a - b) > mpf("0.01") for a, b in zip(sorted(g10)[:-1], sorted(g10)[1:]))
    g10 = sorted(g10)
    w10_all = dict(zip(g10, weights(g10)))
    k21_new = extend(g10)
    n21 = sorted(g10 + k21_new)
    w21_all = dict(zip(n21, weights(n21)))
    p43_new = extend(n21)
    n43 = sorted(n21 + p43_new)
    w43_all = dict(zip(n43, weights(n43)))
    p87_new = extend(n43)
    n87 = sorted(n43 + p87_new)
    w87_all = dict(zip(n87, weights(n87)))

    tiny = mpf(10) ** (-60)

    def pos_desc(xs):
        return sorted([x for x in xs if x > tiny], reverse=True)

    def centre(xs):
        c = [x for x in xs if abs(x) <= tiny]
        assert len(c) == 1
        return c[0]

    x1 = pos_desc(g10)
    x2 = pos_desc(k21_new)
    x3 = pos_desc(p43_new)
    x4 = pos_desc(p87_new)
    assert (len(x1), len(x2), len(x3), len(x4)) == (5, 5, 11, 22)
    tab = {
        "x1": x1, "w10": [w10_all[x] for x in x1],
        "x2": x2, "w21a": [w21_all[x] for x in x1],
        "w21b": [w21_all[x] for x in x2] + [w21_all[centre(k21_new)]],
        "x3": x3, "w43a": [w43_all[x] for x in x1 + x2],
        "w43b": [w43_all[x] for x in x3] + [w43_all[centre(n43)]],
        "x4": x4, "w87a": [w87_all[x] for x in x1 + x2 + x3],
        "w87b": [w87_all[x] for x in x4] + [w87_all[centre(n87)]],
    }
    # sanity: each rule integrates x^k exactly up to its degree of exactness; weights sum to 2
    for name, nodes, wall, deg in (("g10", g10, w10_all, 19), ("k21", n21, w21_all, 31),
                                   ("p43", n43, w43_all, 64), ("p87", n87, w87_all, 130)):
        assert abs(sum(wall.values()) - 2) < mpf(10) ** (-80), name
        for k in (0, 2, deg - 1 if (deg - 1) % 2 == 0 else deg - 2):
            exact = mpf(2) / (k + 1)
            got = sum(w * x ** k for x, w in wall.items())
            assert abs(got - exact) < mpf(10) ** (-60), (name, k, got - exact)
        assert all(w > 0 for w in wall.values()), name

    lines = [
        "/* GENERATED by oracle/gen_qng_tables.py (mpmath, %d digits) - do not edit.",
        " * Gauss-Kronrod-Patterson 10/21/43/87-point abscissae and weights, in the array layout of",
        " * QUADPACK QNG / GSL gsl_integration_qng (the routine pplacer reaches through",
        " * pplacer_src/integration.ml:27-31). Positive abscissae in descending order. */",
        "#ifndef PPB_QNG_TABLES_H",
        "#define PPB_QNG_TABLES_H",
        "#ifndef PPB_QNG_CONST",
        "#define PPB_QNG_CONST static const",
        "#endif",
    ]
    lines[0] = lines[0] % mp.dps
    for name in ("x1", "w10", "x2", "w21a", "w21b", "x3", "w43a", "w43b", "x4", "w87a", "w87b"):
        vals = tab[name]
        lines.append("PPB_QNG_CONST double ppb_qng_%s[%d] = {" % (name, len(vals)))
        for v in vals:
            lines.append("  %s," % mp.nstr(v, 25, min_fixed=-1, max_fixed=1))
        lines.append("};")
    lines.append("#endif")
    text = "\n".join(lines) + "\n"
    here = os.path.dirname(os.path.abspath(__file__))
    for out in (os.path.join(here, "qng_tables.h"),
                os.path.join(here, "..", "pplacer_b200", "csrc", "qng_tables.h")):
        with open(out, "w") as f:
            f.write(text)
        print("wrote", out)
    print("x1[0] =", mp.nstr(x1[0], 20), " w21b[5] =", mp.nstr(tab["w21b"][5], 20),
          " w43b[11] =", mp.nstr(tab["w43b"][11], 20), " w87b[22] =", mp.nstr(tab["w87b"][22], 20))


if __name__ == "__main__":
    sys.exit(main())
