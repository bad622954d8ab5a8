"""Substitution-model inputs for the placement path (host side).

Only *inputs* are assembled here (exchangeabilities, stationary-frequency source, rate-category
description); the numerics — discrete Gamma rates (gamma.ml:52-70), Q = B.diag(pi)
diagonalisation (diagd.ml:123-172), empirical frequencies (alignment.ml:186-217) — are done by
the C++ host code inside libpplacer_b200.so (pplacer_b200/csrc/model.cpp), and independently by
the oracle.

Follows pplacer_src/gmix_model.ml:386-413 (init_of_json, ras_model "gamma"),
pplacer_src/gcat_model.ml:453-478 (ras_model "Price-CAT", 1-based SiteCategories),
pplacer_src/gstar_support.ml:81-98 (GTR always uses empirical frequencies; proteins use the
model's unless empirical_frequencies), pplacer_src/nuc_models.ml:78-84 (ac ag at cg ct gt order),
pplacer_src/prot_models.ml:52-68 (lower-triangular PAML tables).
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

NUC_CODE = "ACGT"
PROT_CODE = "ARNDCQEGHILKMFPSTWYV"
NUC_ALPHABET = "ACGTRYMKSWHBVDN-?X"          # nuc_models.ml:28-45
PROT_ALPHABET = PROT_CODE + "BJZX-?"          # prot_models.ml:5,40-44


@dataclass
class ModelSpec:
    n_states: int                 # 4 | 20
    model_class: int              # 0 Gmix (gamma mixture) | 1 Gcat (Price-CAT)
    subs_model: str               # GTR | LG | WAG | JTT
    exch: np.ndarray              # float64[S,S] symmetric, zero diagonal
    model_statd: Optional[np.ndarray]   # float64[S] for protein models, None for GTR
    empirical_freqs: bool         # take pi from the (un-premasked) reference alignment
    n_rates: int
    gamma_alpha: float            # Gmix only
    rates: Optional[np.ndarray]   # Gcat: CAT rates; Gmix: None (computed) or explicit
    site_cat: Optional[np.ndarray]  # Gcat: int32[L_raw], zero-based

    def alphabet(self) -> str:
        return NUC_ALPHABET if self.n_states == 4 else PROT_ALPHABET


def gtr_exch(ac, ag, at, cg, ct, gt) -> np.ndarray:
    """Nuc_models.b_of_trans_vector (nuc_models.ml:78-84)."""
    m = np.zeros((4, 4))
    for (i, j), v in zip(((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)), (ac, ag, at, cg, ct, gt)):
        m[i, j] = m[j, i] = v
    return m


def load_prot_model(name: str):
    """Prot_models.trans_and_statd_of_model_name (prot_models.ml:160-163); parseProtModel :55-66."""
    name = name.upper()
    if name not in ("LG", "WAG", "JTT"):
        raise ValueError("unknown model: " + name)
    with open(os.path.join(_DATA, name + ".dat")) as fh:
        rows = [ln.split() for ln in fh if ln.strip()]
    tri, freq = rows[:19], rows[19]
    m = np.zeros((20, 20))
    for i in range(20):
        for j in range(20):
            if i != j:
                m[i, j] = float(tri[max(i, j) - 1][min(i, j)])
    return m, np.array([float(x) for x in freq])


def model_from_json(obj: dict) -> ModelSpec:
    subs = obj["subs_model"]
    ras = obj["ras_model"]
    emp = bool(obj["empirical_frequencies"])
    if subs == "GTR":
        sr = obj.get("subs_rates")
        if sr is None:
            raise ValueError("GTR specified but no substitution rates given.")
        exch = gtr_exch(*(float(sr[k]) for k in ("ac", "ag", "at", "cg", "ct", "gt")))
        S, statd, emp = 4, None, True   # gstar_support.ml:83-89: GTR => empirical pi
    else:
        exch, statd = load_prot_model(subs)
        S = 20
    if ras == "gamma":
        g = obj["gamma"]
        return ModelSpec(S, 0, subs, exch, statd, emp, int(g["n_cats"]), float(g["alpha"]), None, None)
    if ras == "Price-CAT":
        pc = obj["Price-CAT"]
        rates = np.array([float(x) for x in pc["Rates"]])
        cats = np.array([int(x) - 1 for x in pc["SiteCategories"]], dtype=np.int32)  # 1-indexed in file
        return ModelSpec(S, 1, subs, exch, statd, emp, len(rates), 0.0, rates, cats)
    raise ValueError("unknown ras_model " + ras)


def model_from_json_file(path: str) -> ModelSpec:
    with open(path) as fh:
        return model_from_json(json.load(fh))
