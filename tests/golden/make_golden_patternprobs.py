"""Fixtures behind the statistical tests of the alignment simulators, from the UNMODIFIED reference:

  * exact site-pattern probabilities of small Jukes-Cantor trees: simulation.get_pattern_probabilities
    (splitp/simulation.py:71-157);
  * a seeded sample of the reference's own simulator on a 4-taxon GTR tree: simulation.generate_alignment
    (splitp/simulation.py:9-56).  For a non-symmetric transition matrix the reference draws the child state with
    random.choices(weights = COLUMN `parent state` of expm(t Q)) (simulation.py:18-19), which normalises the column --
    so for GTR the sampled distribution is NOT the one get_pattern_probabilities tabulates (that table does not even sum
    to 1); the simulators follow evolve_pattern, hence a sample of it as the fixture.

    PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference python tests/golden/make_golden_patternprobs.py
"""
import json
import os
import random

from splitp import model as M
from splitp import simulation, trees

HERE = os.path.dirname(os.path.abspath(__file__))
exact = []
for n, bl in ((4, 0.05), (4, 0.3), (6, 0.1)):
    tree = trees.balanced_newick_tree(n, bl)
    probs = simulation.get_pattern_probabilities(tree, M.GTR.JukesCantor(1 / 2))
    assert abs(sum(probs.values()) - 1.0) < 1e-12
    exact.append({"n": n, "branch_length": bl, "model": "JC", "taxa": list(tree.taxa), "newick": tree.newick_string,
                  "patterns": list(probs.keys()), "probs": [float(v) for v in probs.values()]})
    print("exact", n, bl, len(probs))
sampled = []
for n, bl, N in ((4, 0.3, 40_000),):
    tree = trees.balanced_newick_tree(n, bl)
    mdl = M.GTR("gtr", ("A", "C", "G", "T"), [0.1, 0.2, 0.3, 0.4], [1, 2, 3, 4, 5, 6])
    random.seed(0)
    aln = simulation.generate_alignment(tree, mdl, N)
    sampled.append({"n": n, "branch_length": bl, "model": "GTR", "sites": N, "seed": 0, "taxa": list(tree.taxa),
                    "patterns": list(aln.keys()), "counts": [int(round(v * N)) for v in aln.values()]})
    print("sampled", n, bl, len(aln))
with open(os.path.join(HERE, "golden_patternprobs.json"), "w") as f:
    json.dump({"exact": exact, "sampled": sampled, "source": "js51/SplitP v0.3.2 (unmodified): simulation.get_pattern_probabilities / "
               "generate_alignment"}, f)
