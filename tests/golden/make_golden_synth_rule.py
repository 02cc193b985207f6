"""Mint ``synth_overlap_rule.npz`` by running THE REFERENCE'S OWN ``bounding_boxes_overlap`` (multi_mnist.py:77-87).
``multi_mnist.py`` imports TensorFlow at module level, so the module cannot be imported here; the function is pure
Python, so its definition is taken from the file where it lies (``ast``; nothing is copied into the repo) and executed
on seeded boxes.  Run from the repo root in the authoring container: ``python tests/golden/make_golden_synth_rule.py``."""
import ast
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/multi_mnist.py"
tree = ast.parse(open(REF).read())
fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "bounding_boxes_overlap")
ns = {}
exec(compile(ast.Module(body=[fn], type_ignores=[]), REF, "exec"), ns)
ref_fn = ns["bounding_boxes_overlap"]

rng = np.random.default_rng(0)
N = 4000
new = np.stack([rng.integers(-2, 50, N), rng.integers(-2, 50, N), rng.integers(1, 26, N), rng.integers(1, 26, N)], 1)   # x, y, w, h
old = np.stack([rng.integers(0, 50, N), rng.integers(0, 50, N), rng.integers(1, 26, N), rng.integers(1, 26, N)], 1)
gap = rng.integers(0, 4, N)
# a share of near-touching pairs: the rule's boundaries
k = N // 3
new[:k, 0] = old[:k, 0] + old[:k, 2] + rng.integers(-2, 3, k) + gap[:k]
new[k:2 * k, 0] = old[k:2 * k, 0] - new[k:2 * k, 2] - gap[k:2 * k] + rng.integers(-2, 3, k)
out = np.array([ref_fn(int(a[0]), int(a[1]), int(a[2]), int(a[3]), [int(b[0]), int(b[1])], [int(b[2]), int(b[3])], int(g))
                for a, b, g in zip(new, old, gap)], bool)
# and lists of several placed boxes (the function returns on the first hit)
multi_new = np.stack([rng.integers(0, 40, 500), rng.integers(0, 40, 500), rng.integers(5, 16, 500), rng.integers(5, 16, 500)], 1)
multi_old = np.stack([rng.integers(0, 40, (500, 3)), rng.integers(0, 40, (500, 3)), rng.integers(5, 16, (500, 3)), rng.integers(5, 16, (500, 3))], 2)
multi_out = np.array([ref_fn(int(a[0]), int(a[1]), int(a[2]), int(a[3]), [int(v) for q in o for v in q[:2]], [int(v) for q in o for v in q[2:]], 1)
                      for a, o in zip(multi_new, multi_old)], bool)
np.savez_compressed(os.path.join(HERE, "synth_overlap_rule.npz"), new=new, old=old, gap=gap, overlap=out, multi_new=multi_new,
                    multi_old=multi_old, multi_overlap=multi_out)
print("pairs:", N, "overlap share", out.mean(), "| lists:", len(multi_out), "overlap share", multi_out.mean())
