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
