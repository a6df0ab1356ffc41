"""Tree / alignment encoders: the data-dict side of the likelihood boundary.

Host-side mirror of the reference's encoders (paths relative to /root/reference):

* ``setup_indexes``      phylostan/utils.py:59-72   tips 1..S in taxon-namespace order,
                                                    internals S+1.. in post-order, root 2S-1
* ``get_peeling_order``  phylostan/utils.py:75-81   rows [child1, child2, parent], post-order
* ``get_preorder``       phylostan/utils.py:84-90   rows [node, parent], pre-order
* ``get_lowers``         phylostan/utils.py:93-104
* ``compress_patterns``  phylostan/utils.py:156-190 (get_dna_leaves_partials_compressed)
* ``unrooted_swap``      phylostan/phylostan.py:264-267

The reference leans on dendropy for parsing; dendropy is not part of this path, so a small
Newick / FASTA / NEXUS reader lives here.  The boundary handed to the GPU library is exactly
the reference's Stan data dict (``peel``, ``tipdata``, ``weights``, ``map``), with ``tipdata``
packed as one 4-bit state mask per (tip, pattern) instead of four ints.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np

DNA_ORDER = "acgt"  # phylostan/utils.py:180


# --------------------------------------------------------------------------- tree

@dataclass(eq=False)
class Node:
    label: Optional[str] = None
    edge_length: Optional[float] = None
    children: List["Node"] = field(default_factory=list)
    parent: Optional["Node"] = None
    index: int = -1
    date: float = 0.0

    def is_leaf(self) -> bool:
        return not self.children


class Tree:
    """Rooted tree; taxon namespace = tip labels in order of first appearance in the Newick
    string (what dendropy builds when the tree file is read before the alignment,
    phylostan/phylostan.py:165-185)."""

    def __init__(self, root: Node):
        self.root = root
        self.taxa: List[str] = [n.label for n in self.preorder() if n.is_leaf()]

    def postorder(self) -> Iterator[Node]:
        stack: List[Tuple[Node, int]] = [(self.root, 0)]
        while stack:
            node, i = stack.pop()
            if i < len(node.children):
                stack.append((node, i + 1))
                stack.append((node.children[i], 0))
            else:
                yield node

    def preorder(self) -> Iterator[Node]:
        stack = [self.root]
        while stack:
            node = stack.pop()
            yield node
            stack.extend(reversed(node.children))

    def leaves(self) -> Iterator[Node]:
        return (n for n in self.preorder() if n.is_leaf())

    @property
    def n_tips(self) -> int:
        return len(self.taxa)

    def resolve_polytomies(self) -> None:
        """Make the tree strictly bifurcating (phylostan/phylostan.py:175 calls dendropy's
        ``resolve_polytomies``).  Extra children are folded left-to-right under new
        zero-length internal nodes; dendropy's exact attachment order cannot be checked
        here (no dendropy) and does not change the likelihood."""
        for node in list(self.postorder()):
            while len(node.children) > 2:
                a, b = node.children[0], node.children[1]
                new = Node(edge_length=0.0 if a.edge_length is not None else None, children=[a, b], parent=node)
                a.parent = b.parent = new
                node.children = [new] + node.children[2:]


def parse_newick(text: str) -> Tree:
    """Minimal Newick reader: labels (optionally quoted), branch lengths, [comments]."""
    text = re.sub(r"\[[^\]]*\]", "", text.strip())
    if text.endswith(";"):
        text = text[:-1]
    pos = 0

    def parse_label() -> Optional[str]:
        nonlocal pos
        if pos < len(text) and text[pos] == "'":
            end = text.index("'", pos + 1)
            lab = text[pos + 1:end]
            pos = end + 1
            return lab
        start = pos
        while pos < len(text) and text[pos] not in ",():;":
            pos += 1
        lab = text[start:pos].strip()
        return lab or None

    def parse_length() -> Optional[float]:
        nonlocal pos
        if pos < len(text) and text[pos] == ":":
            pos += 1
            start = pos
            while pos < len(text) and text[pos] not in ",();":
                pos += 1
            return float(text[start:pos])
        return None

    def parse_node(parent: Optional[Node]) -> Node:
        nonlocal pos
        node = Node(parent=parent)
        if text[pos] == "(":
            pos += 1
            while True:
                node.children.append(parse_node(node))
                if text[pos] == ",":
                    pos += 1
                    continue
                if text[pos] == ")":
                    pos += 1
                    break
                raise ValueError(f"bad Newick at {pos}: {text[pos:pos + 20]!r}")
        node.label = parse_label()
        node.edge_length = parse_length()
        return node

    return Tree(parse_node(None))


def read_tree(path: str, offset: int = 0) -> Tree:
    """Newick file (one tree per line) or NEXUS trees block (phylostan/phylostan.py:167-173)."""
    with open(path) as fp:
        content = fp.read()
    if content.lstrip().upper().startswith("#NEXUS"):
        translate: Dict[str, str] = {}
        m = re.search(r"translate(.*?);", content, flags=re.I | re.S)
        if m:
            for item in m.group(1).split(","):
                parts = item.split()
                if len(parts) >= 2:
                    translate[parts[0]] = parts[1].strip("'")
        trees = re.findall(r"^\s*tree\s+[^=]+=\s*(?:\[[^\]]*\]\s*)?(.*?;)", content, flags=re.I | re.M | re.S)
        tree = parse_newick(trees[offset])
        if translate:
            for leaf in tree.leaves():
                leaf.label = translate.get(leaf.label, leaf.label)
            tree.taxa = [n.label for n in tree.leaves()]
        return tree
    lines = [ln for ln in content.splitlines() if ln.strip()]
    return parse_newick(lines[offset])


# --------------------------------------------------------------------------- alignment

def read_fasta(path: str) -> Dict[str, str]:
    seqs: Dict[str, List[str]] = {}
    name = None
    with open(path) as fp:
        for line in fp:
            line = line.strip()
            if not line:
                continue
            if line.startswith(">"):
                name = line[1:].strip()
                seqs[name] = []
            else:
                seqs[name].append(line)
    return {k: "".join(v) for k, v in seqs.items()}


def read_nexus_matrix(path: str) -> Dict[str, str]:
    """Sequential or interleaved MATRIX of the first DATA/CHARACTERS block."""
    with open(path) as fp:
        content = fp.read()
    content = re.sub(r"\[[^\]]*\]", "", content)
    m = re.search(r"\bmatrix\b(.*?);", content, flags=re.I | re.S)
    if not m:
        raise ValueError("no MATRIX block in " + path)
    seqs: Dict[str, List[str]] = {}
    for line in m.group(1).splitlines():
        parts = line.split()
        if len(parts) < 2:
            continue
        name = parts[0].strip("'")
        seqs.setdefault(name, []).append("".join(parts[1:]))
    return {k: "".join(v) for k, v in seqs.items()}


def read_alignment(path: str) -> Dict[str, str]:
    """FASTA if the file starts with '>' else NEXUS (phylostan/phylostan.py:186-190)."""
    with open(path) as fp:
        first = fp.readline()
    return read_fasta(path) if first.startswith(">") else read_nexus_matrix(path)


# --------------------------------------------------------------------------- encoders

def setup_indexes(tree: Tree) -> None:
    """phylostan/utils.py:59-72."""
    taxa = {label: i for i, label in enumerate(tree.taxa)}
    s = len(tree.taxa) + 1
    for node in tree.postorder():
        if node.is_leaf():
            node.index = taxa[node.label] + 1
        else:
            node.index = s
            s += 1


def get_peeling_order(tree: Tree) -> np.ndarray:
    """phylostan/utils.py:75-81 -> int32 [S-1, 3], 1-based."""
    rows = [[c.index for c in n.children] + [n.index] for n in tree.postorder() if not n.is_leaf()]
    return np.asarray(rows, dtype=np.int32).reshape(-1, 3)


def get_preorder(tree: Tree) -> np.ndarray:
    """phylostan/utils.py:84-90 -> int32 [2S-1, 2] rows [node, parent] (root's parent 0)."""
    rows = [[tree.root.index, 0]]
    rows += [[n.index, n.parent.index] for n in tree.preorder() if n.parent is not None]
    return np.asarray(rows, dtype=np.int32)


def setup_dates(tree: Tree, dates: Optional[Dict[str, float]] = None, heterochronous: bool = False) -> Optional[float]:
    """phylostan/utils.py:5-57: sampling dates -> ``node.date`` = time before the most recent tip.

    ``dates``: taxon -> date (what the reference reads from its csv file); None with
    ``heterochronous=True`` takes the root-to-tip distances of the (time) tree, as ``get_dates`` does.
    Returns ``oldest`` (None for contemporaneous tips)."""
    if dates:
        heterochronous = True
    if not heterochronous:
        for node in tree.postorder():
            node.date = 0.0
        return None
    if not dates:
        dist = {id(tree.root): 0.0}
        dates = {}
        for node in tree.preorder():
            if node.parent is not None:
                dist[id(node)] = dist[id(node.parent)] + (node.edge_length or 0.0)
                if node.is_leaf():
                    dates[node.label] = dist[id(node)]
    hi, lo = max(dates.values()), min(dates.values())
    for node in tree.leaves():
        node.date = dates[node.label] if lo == 0 else hi - dates[node.label]      # utils.py:42-52
    return hi if lo == 0 else hi - lo


def get_lowers(tree: Tree) -> np.ndarray:
    """phylostan/utils.py:93-104: lower bound of each node height = max date below it."""
    ll: Dict[int, float] = {}
    lowers = np.zeros(sum(1 for _ in tree.postorder()))
    for node in tree.postorder():
        ll[id(node)] = node.date if node.is_leaf() else max(ll[id(c)] for c in node.children)
    for node in tree.preorder():
        lowers[node.index - 1] = ll[id(node)]
    return lowers


def unrooted_swap(peel: np.ndarray) -> np.ndarray:
    """phylostan/phylostan.py:264-267: for unconstrained trees the last peel row lists the
    smaller child index first; the second child (2S-2) then carries no branch."""
    peel = np.array(peel, dtype=np.int32, copy=True)
    if peel[-1, 0] > peel[-1, 1]:
        peel[-1, 0], peel[-1, 1] = peel[-1, 1], peel[-1, 0]
    return peel


def state_mask(ch: str) -> int:
    """One-hot A,C,G,T -> bit 0..3; every other symbol -> 1,1,1,1 (phylostan/utils.py:180-188)."""
    i = DNA_ORDER.find(ch.lower())
    return (1 << i) if i >= 0 and ch != "" else 0xF


def compress_patterns(seqs: Dict[str, str], taxa: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:
    """phylostan/utils.py:156-190.  Patterns are distinct raw columns (case-sensitive string of
    the column, first occurrence order); returns (tipmask uint8 [S, L], weights float64 [L])."""
    rows = [seqs[t] for t in taxa]
    n = len(rows[0])
    if any(len(r) != n for r in rows):
        raise ValueError("sequences differ in length")
    count: Dict[str, int] = {}
    keep: List[int] = []
    for i in range(n):
        pat = "".join(r[i] for r in rows).upper()  # dendropy symbols are case-normalised
        if pat in count:
            count[pat] += 1
        else:
            count[pat] = 1
            keep.append(i)
    weights = np.asarray([count["".join(r[i] for r in rows).upper()] for i in keep], dtype=np.float64)
    tipmask = np.empty((len(rows), len(keep)), dtype=np.uint8)
    for s, r in enumerate(rows):
        tipmask[s] = [state_mask(r[i]) for i in keep]
    return tipmask, weights


def tipdata_to_mask(tipdata: np.ndarray) -> np.ndarray:
    """Reference layout ``tipdata[S, L, 4]`` in {0,1} -> uint8 mask [S, L]."""
    td = np.asarray(tipdata)
    return ((td[..., 0] != 0) * 1 + (td[..., 1] != 0) * 2 + (td[..., 2] != 0) * 4 + (td[..., 3] != 0) * 8).astype(np.uint8)


def mask_to_tipdata(tipmask: np.ndarray) -> np.ndarray:
    m = np.asarray(tipmask, dtype=np.uint8)
    return np.stack([(m >> s) & 1 for s in range(4)], axis=-1).astype(np.int64)


def branch_lengths(tree: Tree, rooted: bool = True) -> np.ndarray:
    """Branch length above node k at position k-1 (``blens`` convention,
    phylostan/generate_script.py:660-679).  Unrooted: the root edge of child 2S-2 is merged
    into its sibling's (phylostan/phylostan.py:232-240 does the same for the geo model)."""
    S = tree.n_tips
    out = np.zeros(2 * S - 1)
    for node in tree.postorder():
        out[node.index - 1] = node.edge_length if node.edge_length is not None else 0.0
    if rooted:
        return out[:2 * S - 2]
    c = sorted(ch.index for ch in tree.root.children)
    out[c[0] - 1] += out[c[1] - 1]
    return out[:2 * S - 3]


@dataclass
class PhyloData:
    """The slice of the reference's Stan data dict that the likelihood consumes."""
    S: int
    L: int
    peel: np.ndarray      # int32 [S-1, 3]
    tipmask: np.ndarray   # uint8 [S, L]
    weights: np.ndarray   # float64 [L]
    map: np.ndarray       # int32 [2S-1, 2]
    taxa: List[str]
    rooted: bool


def encode(tree: Tree, seqs: Dict[str, str], rooted: bool = True) -> PhyloData:
    """Everything ``run()`` puts in the data dict for the likelihood
    (phylostan/phylostan.py:175-204,255-267)."""
    tree.resolve_polytomies()
    setup_indexes(tree)
    if set(tree.taxa) != set(seqs):
        raise ValueError("taxon names in trees and alignment are different")  # phylostan.py:194-196
    peel = get_peeling_order(tree)
    if not rooted:
        peel = unrooted_swap(peel)
    tipmask, weights = compress_patterns(seqs, tree.taxa)
    return PhyloData(S=tree.n_tips, L=int(weights.size), peel=peel, tipmask=tipmask, weights=weights,
                     map=get_preorder(tree), taxa=list(tree.taxa), rooted=rooted)


def weibull_rates(wshape: float, C: int) -> np.ndarray:
    """Discretised Weibull site rates, phylostan/generate_script.py:267-278."""
    i = np.arange(C)
    rs = (-np.log(1.0 - (2.0 * i + 1.0) / (2.0 * C))) ** (1.0 / wshape)
    return rs / (rs.sum() / C)
