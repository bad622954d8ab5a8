"""Model inputs from a tree-builder's statistics file (the deprecated `tree_stats` entry of a reference package,
refpkg.ml:51-59 -> Gmix_model.init_of_stats_fname, gmix_model.ml:363-384): PhyML v3 statistics, RAxML v7 info files
(plain and "-f e" re-estimated) and the JSON form, restating pplacer_src/parse_stats.ml:84-349.  All of them describe a
Gamma-mixture model: (model name, Gamma categories, alpha, GTR exchangeabilities or None for proteins).
"""
from __future__ import annotations

import json
import re
from typing import List, Optional, Tuple

import numpy as np

from . import models


class StatsParsingError(ValueError):
    pass


def _float(s: str) -> float:
    t = re.sub(r"[ \n\t\r]*", "", s)
    try:
        return float(t)
    except ValueError:
        raise StatsParsingError("float_of_string failed on: '%s'" % t)


def _find_beginning(pred, lines: List[str]):
    """File_parsing.find_beginning: (first line matching, the lines after it)."""
    for i, ln in enumerate(lines):
        m = pred(ln)
        if m:
            return m, lines[i + 1:]
    raise LookupError


def _extract(beginning: str, s: str) -> float:
    if not s.startswith(beginning):   # assert_and_extract_float, parse_stats.ml:58-66
        raise StatsParsingError("assert_and_extract_float: didn't match template!")
    return _float(s[len(beginning):])


def _parse_phyml(lines: List[str]):
    """parse_phyml_stats, parse_stats.ml:223-276."""
    try:
        m, after_m = _find_beginning(lambda l: re.match(r". Model of .* substitution:[ \t]+([A-Z]+)", l), lines)
    except LookupError:
        raise StatsParsingError("couldn't find model line!")
    name = m.group(1)
    n_cat, alpha = 1, 1.0          # prefs.ml defaults gamma_n_cat 1, gamma_alpha 1.
    g, after_g = _find_beginning(lambda l: re.match(r". Discrete gamma model:[ \t]*(.*)", l), after_m)
    if g.group(1).strip() == "Yes":
        if len(after_g) < 2:
            raise StatsParsingError("not enough lines after gamma!")
        mc = re.match(r"  - Number of categories:[ \t]*([0-9]*)", after_g[0])
        ma = re.match(r"  - Gamma shape parameter:[ \t]*([0-9\.]*)", after_g[1])
        if mc:
            n_cat = int(mc.group(1))
        if ma:
            alpha = _float(ma.group(1))
    try:
        _, nuc = _find_beginning(lambda l: re.match(r". Nucleotides frequencies:", l), after_g)
    except LookupError:
        if name == "GTR":
            raise StatsParsingError("GTR model specified but I couldn't parse the frequency and rate information.")
        return name, n_cat, alpha, None
    if name != "GTR":
        raise StatsParsingError("model must be GTR if we are using nucleotides")
    after_freq = nuc[1:][4:]                 # blank line, four frequency lines
    rate_lines = after_freq[3:][:6]          # blank, header, blank
    keys = ["  A <-> C", "  A <-> G", "  A <-> T", "  C <-> G", "  C <-> T", "  G <-> T"]
    return name, n_cat, alpha, [_extract(k, l) for k, l in zip(keys, rate_lines)]


def _partitions(lines: List[str]) -> List[List[str]]:
    """File_parsing.partition_list on '^Partition:' lines."""
    parts, cur = [], []
    for ln in lines:
        if ln.startswith("Partition:"):
            parts.append(cur)
            cur = []
        else:
            cur.append(ln)
    parts.append(cur)
    return parts


def _parse_raxml(lines: List[str]):
    """parse_raxml_info, parse_stats.ml:84-211: plain 7.x info file or the re-estimated (-f e) form."""
    parts = _partitions(lines)
    if len(parts) < 2:
        raise StatsParsingError("couldn't find a partition line")
    if len(parts) > 2:
        raise StatsParsingError("too many partitions. Only one is allowed.")
    body = parts[1]
    if any(re.match(r".*raxmlHPC.*-f e (.*)", l) for l in lines):
        try:
            dt, rest = _find_beginning(lambda l: re.match(r"^DataType: (.*)", l), body)
        except LookupError:
            raise StatsParsingError("couldn't find data type line")
        if not rest:
            raise StatsParsingError("unexpected end of file after subs matrix line")
        sm = re.match(r"^Substitution Matrix: (.*)", rest[0])
        if not sm:
            raise StatsParsingError("couldn't match substitution matrix line")
        try:
            al, final = _find_beginning(lambda l: re.match(r"^alpha: (.*)", l), rest)
        except LookupError:
            raise StatsParsingError("couldn't find alpha line")
        name, alpha = sm.group(1).strip(), _float(al.group(1))
        kind = dt.group(1).strip()
        if kind == "AA":
            return name, 4, alpha, None
        if kind != "DNA":
            raise StatsParsingError("data type unknown: " + kind)
        if not final:
            raise StatsParsingError("unexpected end of file after alpha line")
        keys = ["rate A <-> C:", "rate A <-> G:", "rate A <-> T:", "rate C <-> G:", "rate C <-> T:", "rate G <-> T:"]
        return name, 4, alpha, [_extract(k, l) for k, l in zip(keys, final[1:7])]
    try:
        sm, rest = _find_beginning(lambda l: re.match(r"^Substitution Matrix: (.*)", l), body)
    except LookupError:
        raise StatsParsingError("couldn't find substitution matrix line")
    try:
        al, _ = _find_beginning(lambda l: re.match(r".*alpha\[0\]: ([^ ]*) (.*)", l), rest)
    except LookupError:
        raise StatsParsingError("couldn't find alpha line")
    name, alpha, info = sm.group(1).strip(), _float(al.group(1)), al.group(2)
    rm = re.match(r"^rates\[0\] ac ag at cg ct gt: (.*)", info)
    if rm:
        if name != "GTR":
            raise StatsParsingError("have rates but model is not GTR! GTR is only allowed nucleotide model.")
        return name, 4, alpha, [_float(x) for x in rm.group(1).split(" ") if x]
    if name == "GTR":
        raise StatsParsingError("GTR model specified, but no rates found.")
    return name, 4, alpha, None


def _parse_json(path: str):
    """parse_json, parse_stats.ml:287-324."""
    with open(path) as fh:
        root = json.load(fh)
    name = root["subs_model"]
    n_cat, alpha = 1, 1.0
    ras = root.get("ras_model")
    if ras == "gamma":
        n_cat, alpha = int(root["gamma"]["n_cats"]), float(root["gamma"]["alpha"])
    elif ras is not None:
        raise StatsParsingError("unknown ras_model " + str(ras))
    dt = root["datatype"]
    if dt == "DNA":
        if name != "GTR":
            raise StatsParsingError("unsupported model " + name)
        return name, n_cat, alpha, [float(root["subs_rates"][k]) for k in ("ac", "ag", "at", "cg", "ct", "gt")]
    if dt == "AA":
        if name != "WAG":
            raise StatsParsingError("unsupported model " + name)
        return name, n_cat, alpha, None
    raise StatsParsingError("unknown datatype " + dt)


def parse_stats(path: str) -> Tuple[str, int, float, Optional[List[float]]]:
    """Parse_stats.parse_stats (parse_stats.ml:330-349) -> (model name, gamma categories, alpha, GTR rates | None)."""
    if path.endswith(".json"):
        return _parse_json(path)
    with open(path) as fh:
        lines = fh.read().split("\n")
    for ln in lines:
        if re.match(r".* RAxML version ([^ ]+)", ln):
            return _parse_raxml(lines)
    for ln in lines:
        if re.match(r"[ \t]*---  PhyML v([^ ]+)", ln):
            return _parse_phyml(lines)
    raise StatsParsingError("is this a RAxML v7 info or PHYML v3 statistics file? The header didn't match.")


def model_from_stats_file(path: str, empirical_freqs: bool = True) -> models.ModelSpec:
    """Gmix_model.init_of_stats_fname (gmix_model.ml:363-384): always a Gamma mixture; GTR takes its stationary
    distribution from the alignment (gstar_support.ml:83-89), proteins from the model unless empirical_freqs."""
    name, n_cat, alpha, rates = parse_stats(path)
    if rates is not None:
        return models.ModelSpec(4, 0, "GTR", models.gtr_exch(*rates), None, True, n_cat, alpha, None, None)
    exch, statd = models.load_prot_model(name)
    return models.ModelSpec(20, 0, name.upper(), exch, statd, empirical_freqs, n_cat, alpha, None, None)
