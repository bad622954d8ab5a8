"""Helpers shared by the tests that use the reference's OWN linear_c.c (oracle/_ref/liblinear_ref.so): Diagd.multi_exp,
Model.evolve_into and the log-dots computed by the reference's C functions only."""
import numpy as np

from pplacer_b200 import models

LN2 = float(np.log(2.0))
_AMBIG_NUC = {"R": "AG", "Y": "CT", "M": "AC", "K": "GT", "S": "CG", "W": "AT", "H": "ACT", "B": "CGT", "V": "ACG", "D": "AGT"}
_AMBIG_AA = {"B": "DN", "J": "IL", "Z": "EQ"}


def indicator(chars, S):
    """lv_of_nuc / lv_of_aa (nuc_models.ml:28-54, prot_models.ml:5-50): [L][S] zero-one likelihood vectors"""
    code = models.NUC_CODE if S == 4 else models.PROT_CODE
    amb = _AMBIG_NUC if S == 4 else _AMBIG_AA
    out = np.zeros((len(chars), S))
    for i, ch in enumerate(bytes(chars).decode("ascii")):
        if ch in code:
            out[i, code.index(ch)] = 1.0
        elif ch in amb:
            for c in amb[ch]:
                out[i, code.index(c)] = 1.0
        else:
            assert ch in "N-?X", ch
            out[i, :] = 1.0
    return out


class RefC:
    """Diagd.multi_exp + Model.evolve_into + the log-dots, computed by the reference's C functions only."""

    def __init__(self, ob, R, num, gcat, site_cat):
        self.ob, self.R, self.num, self.gcat, self.site_cat = ob, R, num, gcat, site_cat
        self.S = len(num["statd"])

    def pmats(self, bl):
        """diagd.ml:79-98: one dediagonalize per rate with exp(lambda * rate * bl)"""
        n = self.num
        out = np.zeros((len(n["rates"]), self.S, self.S))
        for r, rate in enumerate(n["rates"]):
            ex = np.exp(n["lam"] * (rate * bl))
            self.R.ref_dediagonalize(self.ob.dp(out[r]), self.ob.dp(np.ascontiguousarray(n["x"])), self.ob.dp(ex),
                                     self.ob.dp(np.ascontiguousarray(n["xit"])), self.S)
        return out

    def evolve(self, a, bl):
        """gmix_model.ml:308-315 (one gemmish per rate) / gcat_model.ml (per-site category matrix)"""
        P = self.pmats(bl)
        if not self.gcat:
            out = np.zeros_like(a)
            for r in range(a.shape[0]):
                self.R.ref_gemmish(self.ob.dp(out[r]), self.ob.dp(P[r]), self.ob.dp(np.ascontiguousarray(a[r])), self.S, a.shape[1])
            return out
        out = np.zeros_like(a)
        for cat in range(P.shape[0]):           # gemmish on the sites of one category at a time
            idx = np.nonzero(self.site_cat == cat)[0]
            if len(idx) == 0:
                continue
            src = np.ascontiguousarray(a[idx])
            dst = np.zeros_like(src)
            self.R.ref_gemmish(self.ob.dp(dst), self.ob.dp(P[cat]), self.ob.dp(src), self.S, len(idx))
            out[idx] = dst
        return out

    def masked_logdot(self, x, y, mask):
        m = mask.astype(np.uint16)
        if self.gcat:
            return self.R.ref_mat_masked_logdot(self.ob.dp(x), self.ob.dp(y), m.ctypes.data_as(self.ob.c_u16_p), x.shape[0], self.S)
        util = np.zeros(x.shape[1])
        return self.R.ref_ten_masked_logdot(self.ob.dp(x), self.ob.dp(y), m.ctypes.data_as(self.ob.c_u16_p), self.ob.dp(util),
                                            x.shape[0], x.shape[1], self.S)

    def log_like3(self, x, y, z):
        st = np.ascontiguousarray(self.num["statd"])
        if self.gcat:
            return self.R.ref_mat_log_like3(self.ob.dp(st), self.ob.dp(x), self.ob.dp(y), self.ob.dp(z), x.shape[0], self.S)
        util = np.zeros(x.shape[1])
        return self.R.ref_ten_log_like3(self.ob.dp(st), self.ob.dp(x), self.ob.dp(y), self.ob.dp(z), self.ob.dp(util),
                                        x.shape[0], x.shape[1], self.S)


def sites(a, idx, gcat):
    return np.ascontiguousarray(a[idx] if gcat else a[:, idx])
