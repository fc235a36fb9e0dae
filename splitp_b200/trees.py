"""Minimal rooted-tree bookkeeping for the benchmarks and tests.

The reference keeps trees in networkx behind a Newick parser (splitp/phylogeny.py, parsers/newick.py,
both out of scope: SURVEY.md section 2 rows 12-13).  The hot path only needs `.taxa`, `.get_taxa()`,
the edge list (for the site simulator) and the tree's true splits, so this module builds the same
balanced topology and leaf names as `splitp.trees.balanced_newick_tree` (trees.py:6-29: two halves
per node, the larger half first on a left child; leaves named 0-9A-Z in left-to-right order for
n <= 36, t0.. otherwise) directly as a parent array.
"""
from __future__ import annotations

import numpy as np


class Tree:
    """parent[i] = parent node of node i (-1 for the root); leaves carry names; every non-root node has
    the branch length of the edge above it."""

    def __init__(self, parent, names, branch_length):
        self.parent = list(parent)
        self.names = dict(names)  # node index -> leaf name
        self.branch_length = list(branch_length)
        self.taxa = sorted(self.names.values())  # phylogeny.py:49: taxa are sorted

    def get_taxa(self):
        return [self.names[i] for i in sorted(self.names)]

    def children(self, node):
        return [i for i, p in enumerate(self.parent) if p == node]

    def leaves_below(self, node):
        if node in self.names:
            return [self.names[node]]
        out = []
        for c in self.children(node):
            out += self.leaves_below(c)
        return out

    def splits(self, include_trivial=False, as_strings=False):
        """True splits of the tree (phylogeny.py:149-170): taxa[0] on the left, sides in taxa order."""
        seen = set()
        pos = {t: i for i, t in enumerate(self.taxa)}
        for node in range(len(self.parent)):
            below = set(self.leaves_below(node))
            left = tuple(sorted(below, key=pos.__getitem__))
            right = tuple(t for t in self.taxa if t not in below)
            if self.taxa[0] not in left:
                left, right = right, left
            if (include_trivial or (len(left) > 1 and len(right) > 1)) and (left, right) not in seen:
                seen.add((left, right))
                yield f'{"".join(left)}|{"".join(right)}' if as_strings else (left, right)


def as_tree(obj):
    """`Tree` as is; the reference's `Phylogeny` (phylogeny.py:13-56: a networkx DiGraph in `.networkx_graph`, edges
    parent -> child, the branch above a node in its `branch_length` attribute, leaves named by taxon, `.taxa` the sort
    order) converted to the parent-array form the simulator walks."""
    if hasattr(obj, "parent") and hasattr(obj, "names"):
        return obj
    graph = getattr(obj, "networkx_graph", None)
    if graph is None:
        raise TypeError("expected a splitp_b200.trees.Tree or a Phylogeny-like object with .networkx_graph")
    roots = [v for v, d in graph.in_degree() if d == 0]
    if len(roots) != 1:
        raise ValueError("the tree must have exactly one root")
    parent, names, bls, order = [-1], {}, [0.0], [roots[0]]
    index = {roots[0]: 0}
    for v in order:  # breadth first: parents precede their children in the numbering
        kids = list(graph.successors(v))
        if not kids:
            names[index[v]] = str(v)
        for c in kids:
            index[c] = len(parent)
            parent.append(index[v])
            bl = graph.nodes[c].get("branch_length", 0.0)
            bls.append(float(bl) if bl is not None else 0.0)
            order.append(c)
    tree = Tree(parent, names, bls)
    taxa = getattr(obj, "taxa", None)
    if taxa is not None:
        tree.taxa = [str(t) for t in taxa]
    return tree


def _leaf_name(i, num_taxa):
    return str(np.base_repr(i, base=max(i + 1, 2))) if num_taxa <= 36 else f"t{i}"


def balanced_tree(num_taxa, branch_length=None):
    """Same topology / names as splitp.trees.balanced_newick_tree(num_taxa, branch_length)."""
    if num_taxa % 2 != 0:
        raise ValueError(f"There is no balanced tree on {num_taxa} taxa. Please specify an even number.")
    parent, names, bls = [-1], {}, [0.0]
    counter = [0]
    bl = 0.0 if branch_length is None else float(branch_length)

    def grow(node, size, left):
        if size == 1:
            names[node] = _leaf_name(counter[0], num_taxa)
            counter[0] += 1
            return
        big = size // 2 + (size % 2 if left else 0)
        for part, is_left in ((big, True), (size - big, False)):
            parent.append(node)
            bls.append(bl)
            grow(len(parent) - 1, part, is_left)

    if num_taxa == 2:
        for _ in range(2):
            parent.append(0)
            bls.append(bl)
            grow(len(parent) - 1, 1, True)
    else:
        for is_left in (True, False):
            parent.append(0)
            bls.append(bl)
            grow(len(parent) - 1, num_taxa // 2, is_left)
    return Tree(parent, names, bls)
