"""Newick trees with pplacer's node numbering.

Mirrors the behaviour (not the code) of the reference parser:
  * pplacer_src/newick_lexer.mll:19-36 — unquoted labels exclude blanks ()[]{}':;, ; quoted labels
    are '...' with \\' and \\\\ escapes and MAY contain ':' ';' ',' etc.
  * pplacer_src/newick_parser.mly:49-62 — every leaf / internal node takes the next integer when it
    is reduced, i.e. ids are assigned left-to-right in POST-ORDER; these ids are the jplace edge
    numbers and the indices of darr/parr/snodes.
  * `{n}` after a node overrides its id (jplace-style edge numbering, newick_parser.mly:89-95,
    :121-128); `[x]` is an edge label (ignored here except in legacy form).
  * pplacer_src/newick_gtree.ml:145-157 add_zero_root_bl — a root without a branch length gets 0.
  * pplacer_src/newick_bark.ml:69-74 — branch lengths print as %g; edge numbers as {n}.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

_UNQUOTED_STOP = set(" \t\n()[]{}':;,")


@dataclass
class Tree:
    n_nodes: int
    root: int
    children: List[List[int]]          # ordered as in the Newick string
    bl: np.ndarray                      # float64[n_nodes]; NaN where absent (root -> 0 after add_zero_root_bl)
    node_label: List[Optional[str]]     # leaf names / internal labels
    has_bl: np.ndarray = field(default=None)

    # ---- derived views -------------------------------------------------------------------
    def parent(self) -> np.ndarray:
        p = np.full(self.n_nodes, -1, dtype=np.int32)
        for i, ch in enumerate(self.children):
            for c in ch:
                p[c] = i
        return p

    def leaves(self) -> List[int]:
        return [i for i in range(self.n_nodes) if not self.children[i]]

    def nonroot_ids(self) -> List[int]:
        """Stree.nonroot_node_ids (pplacer_src/stree.ml:33-35): ascending ids without the root."""
        return [i for i in range(self.n_nodes) if i != self.root]

    def csr(self):
        ptr = np.zeros(self.n_nodes + 1, dtype=np.int32)
        idx: List[int] = []
        for i, ch in enumerate(self.children):
            idx.extend(ch)
            ptr[i + 1] = len(idx)
        return ptr, np.asarray(idx, dtype=np.int32)

    def n_edges(self) -> int:
        return self.n_nodes - 1

    def tree_length(self) -> float:
        """Gtree.tree_length: sum of branch lengths over non-root edges."""
        return float(sum(self.bl[i] for i in self.nonroot_ids()))

    def postorder(self) -> List[int]:
        out: List[int] = []
        stack = [(self.root, 0)]
        while stack:
            node, k = stack.pop()
            if k < len(self.children[node]):
                stack.append((node, k + 1))
                stack.append((self.children[node][k], 0))
            else:
                out.append(node)
        return out

    def levels(self):
        """depth of each node (root = 0)."""
        d = np.zeros(self.n_nodes, dtype=np.int32)
        for node in reversed(self.postorder()):
            for c in self.children[node]:
                d[c] = d[node] + 1
        return d

    def to_newick(self, with_node_numbers: bool = True) -> str:
        """Newick_gtree.to_string ~with_node_numbers:true (pplacer_src/newick_gtree.ml:55-69):
        label, ':%g' branch length, '{id}' edge number."""

        def fmt_label(s: str) -> str:
            if s == "" or any(ch in _UNQUOTED_STOP for ch in s):
                return "'" + s.replace("\\", "\\\\").replace("'", "\\'") + "'"
            return s

        def rec(i: int, is_root: bool) -> str:
            s = ""
            if self.children[i]:
                s += "(" + ",".join(rec(c, False) for c in self.children[i]) + ")"
            if self.node_label[i] is not None:
                s += fmt_label(self.node_label[i])
            if self.has_bl[i] and not is_root:
                s += ":" + ("%g" % self.bl[i])
            if with_node_numbers:
                s += "{%d}" % i
            return s

        import sys
        lim = sys.getrecursionlimit()
        sys.setrecursionlimit(max(lim, 4 * self.n_nodes + 1000))
        try:
            return rec(self.root, True) + ";"
        finally:
            sys.setrecursionlimit(lim)


class NewickError(ValueError):
    pass


def _tokenize(s: str):
    i, n = 0, len(s)
    while i < n:
        c = s[i]
        if c in " \t\n\r":
            i += 1
        elif c in "():;,[]{}":
            yield c, None
            i += 1
        elif c == "'":
            j = i + 1
            buf = []
            while True:
                if j >= n:
                    raise NewickError("unterminated quoted label")
                if s[j] == "\\" and j + 1 < n and s[j + 1] in "'\\":
                    buf.append(s[j + 1])
                    j += 2
                elif s[j] == "'":
                    break
                else:
                    buf.append(s[j])
                    j += 1
            yield "LABEL", "".join(buf)
            i = j + 1
        else:
            j = i
            while j < n and s[j] not in _UNQUOTED_STOP:
                j += 1
            yield "LABEL", s[i:j]
            i = j


def parse(s: str, add_zero_root_bl: bool = True) -> Tree:
    toks = list(_tokenize(s))
    pos = 0
    counter = [-1]
    children: dict = {}
    bl: dict = {}
    label: dict = {}
    forced_id: dict = {}

    def peek():
        return toks[pos][0] if pos < len(toks) else None

    def take(kind=None):
        nonlocal pos
        if pos >= len(toks):
            raise NewickError("unexpected end of tree")
        t = toks[pos]
        if kind is not None and t[0] != kind:
            raise NewickError("expected %s, got %s" % (kind, t[0]))
        pos += 1
        return t

    # iterative parser (the ar53 tree is deep): explicit stack of open child lists
    stack: List[List[int]] = []
    cur_node: Optional[int] = None

    def new_node(kids):
        counter[0] += 1
        i = counter[0]
        children[i] = kids
        return i

    def read_suffix(i):
        # [LABEL] [: LABEL] [{LABEL}] [[LABEL]]
        if peek() == "LABEL":
            label[i] = take()[1]
        if peek() == ":":
            take()
            v = take("LABEL")[1]
            try:
                bl[i] = float(v)
            except ValueError:
                raise NewickError("branch lengths must be floats")
        if peek() == "{":
            take()
            forced_id[i] = int(take("LABEL")[1])
            take("}")
        if peek() == "[":
            take()
            take("LABEL")
            take("]")

    root = None
    while True:
        t = peek()
        if t == "(":
            take()
            stack.append([])
            continue
        if t in ("LABEL", ":", "{", "["):
            node = new_node([])
            read_suffix(node)
        elif t in (",", ")") or t is None or t == ";":
            node = new_node([])  # empty subtree -> leaf (newick_parser.mly:131)
        else:
            raise NewickError("syntax error parsing")
        # node complete: attach / close groups
        while True:
            if not stack:
                root = node
                break
            stack[-1].append(node)
            t = peek()
            if t == ",":
                take()
                node = None
                break
            if t == ")":
                take()
                kids = stack.pop()
                node = new_node(kids)
                read_suffix(node)
                continue
            raise NewickError("syntax error parsing")
        if root is not None:
            break
    if peek() == ";":
        take()
    if pos != len(toks):
        raise NewickError("trailing characters after tree")

    n = counter[0] + 1
    if forced_id:
        if len(forced_id) != n:
            raise NewickError("tree is only partially numbered with {n}")
        remap = [forced_id[i] for i in range(n)]
        if sorted(remap) != list(range(n)):
            raise NewickError("{n} edge numbers are not a permutation of 0..n-1")
    else:
        remap = list(range(n))
    ch = [None] * n
    blv = np.full(n, np.nan)
    has = np.zeros(n, dtype=bool)
    lab: List[Optional[str]] = [None] * n
    for i in range(n):
        ch[remap[i]] = [remap[c] for c in children[i]]
        if i in bl:
            blv[remap[i]] = bl[i]
            has[remap[i]] = True
        lab[remap[i]] = label.get(i)
    r = remap[root]
    if add_zero_root_bl and not has[r]:
        blv[r] = 0.0
    return Tree(n_nodes=n, root=r, children=ch, bl=blv, node_label=lab, has_bl=has)


def parse_file(path: str, **kw) -> Tree:
    """Newick_gtree.of_file: a single tree on a single (non-empty) line."""
    with open(path) as f:
        lines = [ln for ln in (x.strip() for x in f) if ln]
    if not lines:
        raise NewickError("empty file in " + path)
    if len(lines) != 1:
        raise NewickError("expected a single tree on a single line in " + path)
    return parse(lines[0], **kw)
