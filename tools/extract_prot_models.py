#!/usr/bin/env python
"""One-off provenance script: write the published LG / WAG / JTT exchangeability matrices and
stationary frequencies in PAML .dat layout (19 lower-triangle rows, blank line, 20 frequencies;
amino-acid order ARNDCQEGHILKMFPSTWYV) into pplacer_b200/data/.  The numbers are the published
constants (Le & Gascuel 2008; Whelan & Goldman 2001; Jones, Taylor & Thornton 1992) as pplacer
carries them at pplacer_src/prot_models.ml:71-158; only numeric tokens are read."""
import re, sys, os
src = open(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/pplacer_src/prot_models.ml").read()
out_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pplacer_b200", "data")
os.makedirs(out_dir, exist_ok=True)
for name in ("wag", "lg", "jtt"):
    m = re.search(r"let %s_trans =\s*parseProtModel \[\|(.*?)\|\]" % name, src, re.S)
    rows = re.findall(r'"([^"]*)"', m.group(1))
    assert len(rows) == 19 and all(len(r.split()) == i + 1 for i, r in enumerate(rows)), name
    f = re.search(r'let %s_statd = parseFreq "([^"]*)"' % name, src).group(1).split()
    assert len(f) == 20
    with open(os.path.join(out_dir, name.upper() + ".dat"), "w") as fh:
        for r in rows:
            fh.write(" ".join(r.split()) + "\n")
        fh.write("\n" + " ".join(f) + "\n")
    print(name, "ok")
