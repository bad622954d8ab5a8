#!/usr/bin/env python3
"""How often is a Brent iterate a golden-section step, and how often would the NEXT iterate be one of the two
golden-section points that are known before f(u) is (one per outcome of f(u) <= f(z))?

Builds an instrumented COPY of the oracle under /tmp (oracle/pplacer_oracle.c is not touched), places a C2-shape
sample with it and prints the counts per objective (pendant / distal) and round (first round tol 0.01 / refinement
tol 1e-5).  Evidence for DESIGN.md 3.4 and profiles/README.md r5e (speculative evaluation: measured slower on the GPU).
usage: python tools/brent_trace.py [n_queries]"""
import ctypes as C
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TMP = "/tmp/ppb_brent_trace"


def patch(src: str) -> str:
    def rep(s, old, new):
        assert old in s, old[:40]
        return s.replace(old, new, 1)
    s = src
    s = rep(s, "int orc_brent(orc_fn f, void *ud, double raw_start,",
            "long long g_tr[512]; static __thread int g_kind;\nlong long *orc_trace(void){return g_tr;}\nint orc_brent(orc_fn f, void *ud, double raw_start,")
    s = rep(s, "return orc_brent(opt_pend_fun, tt,", "g_kind = 0; return orc_brent(opt_pend_fun, tt,")
    s = rep(s, "return orc_brent(opt_dist_fun, tt,", "g_kind = 1; return orc_brent(opt_dist_fun, tt,")
    s = rep(s, "      double p = 0, q = 0, r = 0;\n", "      double p = 0, q = 0, r = 0; int golden_taken = 0, forced = 0;\n")
    s = rep(s, "        if ((u - x_left) < t2 || (x_right - u) < t2) dd = (z < midpoint) ? tol : -tol;\n      } else {",
            "        if ((u - x_left) < t2 || (x_right - u) < t2) { dd = (z < midpoint) ? tol : -tol; forced = 1; }\n      } else {\n        golden_taken = 1;")
    s = rep(s, "      if (fabs(dd) >= tol) u = z + dd; else u = z + ((dd > 0) ? tol : -tol);\n      e = ee; d = dd;",
            """      int kind = golden_taken ? 0 : (forced ? 2 : 1);  /* 0 golden, 1 parabolic, 2 parabolic forced to tol, 3 tiny -> tol */
      if (fabs(dd) >= tol) u = z + dd; else { u = z + ((dd > 0) ? tol : -tol); kind = 3; }
      { int tb = (tolerance < 1e-3 ? 64 : 0) + 128 * g_kind; __sync_fetch_and_add(&g_tr[tb + kind], 1);
        int hit = 0; for (int i = 0; i < n_spec; ++i) if (spec[i] == u) hit = 1;
        if (step > 1) __sync_fetch_and_add(&g_tr[tb + 4 + (hit ? 0 : 1)], 1);
        if (!(hit && !prev_skipped)) { __sync_fetch_and_add(&g_tr[tb + 8], 1); prev_skipped = 0; } else prev_skipped = 1; }
      e = ee; d = dd;""")
    s = rep(s, "      f_u = callf(&w, u); if (!isfinite(f_u)) return ORC_GSL_EBADFUNC;",
            """      n_spec = 0;
      for (int oc = 0; oc < 2; ++oc) {  /* outcome f_u <= f_z: u becomes the minimum; else u becomes a bracket end */
        double zl = x_left, zr = x_right, zz;
        if (oc == 0) { zz = u; if (u < z) zr = z; else zl = z; } else { zz = z; if (u < z) zl = u; else zr = u; }
        double mid = 0.5 * (zl + zr), tl = GSL_SQRT_DBL_EPSILON * fabs(zz);
        double ee2 = (zz < mid) ? zr - zz : -(zz - zl), dd2 = golden * ee2;
        spec[n_spec++] = (fabs(dd2) >= tl) ? zz + dd2 : zz + ((dd2 > 0) ? tl : -tl);
      }
      f_u = callf(&w, u); if (!isfinite(f_u)) return ORC_GSL_EBADFUNC;""")
    s = rep(s, "  for (int step = 1;; step++) {\n    if (step > max_iters) return ORC_EXCEEDED_MAX_ITER;",
            "  double spec[4]; int n_spec = 0, prev_skipped = 0;\n  for (int step = 1;; step++) {\n    if (step > max_iters) return ORC_EXCEEDED_MAX_ITER;")
    return s


def main():
    nq = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    os.makedirs(TMP, exist_ok=True)
    for f in ("pplacer_oracle.h", "qng_tables.h"):
        shutil.copy(os.path.join(ROOT, "oracle", f), TMP)
    with open(os.path.join(ROOT, "oracle", "pplacer_oracle.c")) as fh:
        src = patch(fh.read())
    with open(os.path.join(TMP, "pplacer_oracle.c"), "w") as fh:
        fh.write(src)
    so = os.path.join(TMP, "liboracle_trace.so")
    subprocess.run(["gcc", "-O2", "-fPIC", "-std=gnu11", "-ffp-contract=off", "-w", "-shared", "-o", so,
                    os.path.join(TMP, "pplacer_oracle.c"), "-lm", "-lpthread"], check=True)
    from oracle import binding as ob
    ob.ORACLE_SO = so
    from pplacer_b200 import configs, reference
    prob = configs.c2(Q=nq)
    o = ob.Oracle(prob)
    L = ob.lib()
    L.orc_trace.restype = C.POINTER(C.c_longlong)
    o.place(prob.query_chars, ob.default_prefs(prior_mean=reference.flat_exp_prior_mean(prob.tree)), n_threads=os.cpu_count() or 1)
    t = L.orc_trace()
    names = ("golden", "parabolic", "parabolic->tol", "tiny->tol")
    for tb, lab in ((0, "pendant, first round"), (64, "pendant, refinement"), (128, "distal, first round"), (192, "distal, refinement")):
        tot = sum(t[tb + k] for k in range(4))
        if not tot:
            continue
        print("%-22s iterates %6d  %s | next point predicted: %d of %d | passes with 2 speculative points: %d (%.0f %%)" % (
            lab, tot, " ".join("%s %.3f" % (names[k], t[tb + k] / tot) for k in range(4)), t[tb + 4], t[tb + 4] + t[tb + 5],
            t[tb + 8], 100.0 * t[tb + 8] / tot))


if __name__ == "__main__":
    main()
