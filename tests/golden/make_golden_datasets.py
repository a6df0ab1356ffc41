#!/usr/bin/env python3
"""Data-dict fixtures produced by the REFERENCE's own encoders (authoring container only).

Runs /root/reference/phylostan/utils.py -- setup_indexes, get_peeling_order, get_preorder,
get_dna_leaves_partials_compressed -- unmodified on fluA, DS1 (tree 0) and HCV.  utils.py expects
dendropy objects; dendropy is not installed, so the files are parsed with this repo's reader and
wrapped in duck-typed stand-ins exposing just the attributes utils.py touches.  The resulting
arrays (the reference's ``peel``, ``map``, ``tipdata``, ``weights``) are committed as
tests/golden/<name>.npz and the repo's own encoders are tested against them.

What this cannot pin (no dendropy here): taxon-namespace iteration order of a
DnaCharacterMatrix and resolve_polytomies' attachment order (DS1's trifurcating root).

    python tests/golden/make_golden_datasets.py
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from phylostan_b200 import encode as E  # noqa: E402

REF = "/root/reference"
np.int = int  # utils.py:157,177 use the alias removed in numpy>=1.24


class _Taxon:
    def __init__(self, label):
        self.label = label

    def __str__(self):
        return "'%s'" % self.label


class _Annotations:
    def add_bound_attribute(self, name):
        pass


class _Edge:
    def __init__(self, length):
        self.length = length


class _Node:
    def __init__(self, src, parent, taxa):
        self.parent_node = parent
        self.edge_length = src.edge_length
        self.edge = _Edge(src.edge_length)
        self.annotations = _Annotations()
        self.taxon = taxa[src.label] if src.is_leaf() else None
        self._children = [_Node(c, self, taxa) for c in src.children]

    def is_leaf(self):
        return not self._children

    def child_node_iter(self):
        return iter(self._children)

    def child_nodes(self):
        return list(self._children)


class _Tree:
    def __init__(self, tree):
        self.taxon_namespace = [_Taxon(t) for t in tree.taxa]
        taxa = {t.label: t for t in self.taxon_namespace}
        self.seed_node = _Node(tree.root, None, taxa)

    def postorder_node_iter(self):
        def rec(n):
            for c in n._children:
                yield from rec(c)
            yield n
        return rec(self.seed_node)

    def preorder_node_iter(self):
        def rec(n):
            yield n
            for c in n._children:
                yield from rec(c)
        return rec(self.seed_node)

    def leaf_node_iter(self):
        return (n for n in self.preorder_node_iter() if n.is_leaf())


class _Seq:
    def __init__(self, s):
        self._s = s.upper()  # dendropy state symbols are canonical upper case

    def __getitem__(self, i):
        return self._s[i]

    def symbols_as_string(self):
        return self._s


class _Matrix:
    """Iterates taxa in namespace order, like dendropy's CharacterMatrix.__iter__."""

    def __init__(self, seqs, taxa):
        self._taxa = list(taxa)
        self._seqs = {t: _Seq(seqs[t]) for t in taxa}
        self.sequence_size = len(next(iter(seqs.values())))

    def __iter__(self):
        return iter(self._taxa)

    def __getitem__(self, name):
        return self._seqs[name]

    def __len__(self):
        return len(self._taxa)


def load_utils():
    spec = importlib.util.spec_from_file_location("ref_utils", os.path.join(REF, "phylostan", "utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    U = load_utils()
    sets = {
        "fluA": ("examples/fluA/fluA.tree", "examples/fluA/fluA.fa", True),
        "DS1": ("examples/DS1/DS1.trees", "examples/DS1/DS1.nex", False),
        "HCV": ("examples/HCV/HCV.tree", "examples/HCV/HCV.nexus", True),
    }
    for name, (tf, af, rooted) in sets.items():
        tree = E.read_tree(os.path.join(REF, tf))
        tree.resolve_polytomies()
        seqs = E.read_alignment(os.path.join(REF, af))
        rt = _Tree(tree)
        U.setup_indexes(rt)
        peel = np.asarray(U.get_peeling_order(rt), dtype=np.int32)
        pre = np.asarray(U.get_preorder(rt), dtype=np.int32)
        if not rooted:  # phylostan/phylostan.py:264-267
            last = peel[-1].copy()
            if last[0] > last[1]:
                peel[-1] = [last[1], last[0], last[2]]
        tipdata, weights = U.get_dna_leaves_partials_compressed(_Matrix(seqs, tree.taxa))
        blens = np.zeros(2 * len(tree.taxa) - 1)
        for n in rt.postorder_node_iter():
            blens[n.index - 1] = n.edge_length if n.edge_length is not None else np.nan
        out = os.path.join(HERE, name + ".npz")
        np.savez_compressed(out, peel=peel, map=pre, tipmask=E.tipdata_to_mask(tipdata),
                            weights=np.asarray(weights, dtype=np.float64), tree_blens=blens,
                            rooted=np.asarray(rooted), taxa=np.asarray(tree.taxa))
        print(name, "S", len(tree.taxa), "L", len(weights), "sites", int(np.sum(weights)), "->", out,
              os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
