"""FASTA input and alignment helpers (host side; behaviour of pplacer_src/fasta.ml:7-58 and
common_src/alignment.ml:60-76,235).  Names are the first whitespace-delimited token after '>';
';' comment lines are skipped; sequence chunks are concatenated; .gz is read transparently
(the ar53 regression queries ship gzipped)."""
from __future__ import annotations

import gzip
from typing import List, Tuple

import numpy as np


def read_fasta(path: str, upper: bool = True) -> List[Tuple[str, str]]:
    opener = gzip.open if path.endswith(".gz") else open
    out: List[Tuple[str, str]] = []
    name = None
    chunks: List[str] = []
    with opener(path, "rt") as fh:
        for line in fh:
            line = line.strip()
            if not line or line.startswith(";"):
                continue
            if line.startswith(">"):
                if name is not None:
                    out.append((name, "".join(chunks)))
                toks = line[1:].split()
                if not toks:
                    raise ValueError("empty sequence name in " + path)
                name, chunks = toks[0], []
            else:
                if name is None:
                    raise ValueError("sequence before name in " + path)
                chunks.append("".join(line.split()))
    if name is not None:
        out.append((name, "".join(chunks)))
    if upper:
        out = [(n, s.upper()) for n, s in out]
    return out


def to_char_matrix(seqs: List[str]) -> np.ndarray:
    """N strings of equal length -> uint8[N, L]."""
    if not seqs:
        return np.zeros((0, 0), dtype=np.uint8)
    L = len(seqs[0])
    for s in seqs:
        if len(s) != L:
            raise ValueError("sequences are not all the same length (%d vs %d)" % (len(s), L))
    return np.frombuffer("".join(seqs).encode("ascii"), dtype=np.uint8).reshape(len(seqs), L).copy()


def informative(chars: np.ndarray) -> np.ndarray:
    """Alignment.informative (alignment.ml:235): everything except '?' and '-'."""
    return (chars != ord("-")) & (chars != ord("?"))


def read_stockholm(path: str, upper: bool = True) -> List[Tuple[str, str]]:
    """Stockholm 1.0 alignment (pplacer_src/stockholm.ml:9-90): '# STOCKHOLM 1.0' header, '#=' markup ignored, sequence
    lines 'name  chunk' concatenated per name (interleaved blocks), '.' gaps become '-', '//' footer; every sequence
    must have the same length.  The reference collects the names in a string map, so they come back SORTED."""
    opener = gzip.open if path.endswith(".gz") else open
    seqs = {}
    state = 0   # 0 needs header, 1 needs footer, 2 found footer
    with opener(path, "rt") as fh:
        for raw in fh:
            line = raw.strip()
            if not line:
                continue
            if line == "# STOCKHOLM 1.0":
                if state != 0:
                    raise ValueError("stockholm: header unexpected")
                state = 1
            elif state == 0:
                raise ValueError("stockholm: something before header")
            elif line == "//":
                if state != 1:
                    raise ValueError("stockholm: footer unexpected")
                state = 2
            elif state == 2:
                raise ValueError("stockholm: something after footer")
            elif line.startswith("#"):
                continue        # "#=GF", "#=GC", "#=GS", "#=GR" markup
            else:
                toks = line.split()
                if len(toks) != 2:
                    raise ValueError("stockholm: malformed sequence line: " + line[:40])
                seqs.setdefault(toks[0], []).append(toks[1].replace(".", "-"))
    if state != 2:
        raise ValueError("stockholm: no footer")
    out = [(n, "".join(chunks)) for n, chunks in sorted(seqs.items())]
    for n, s in out:
        if len(s) != len(out[0][1]):
            raise ValueError("%s (length %d) doesn't match the previous sequence lengths (%d)" % (n, len(s), len(out[0][1])))
    if upper:
        out = [(n, s.upper()) for n, s in out]
    return out
