"""Golden vectors for result post-processing, produced by RUNNING THE REFERENCE's
utils.transitive_reduction (utils.py:120-129) and utils.hamming_distance (utils.py:148-149).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_reduction.py   # needs /root/reference

Cases: closed DAGs (the use in main.output_handling, main.py:50), unclosed DAGs, digraphs with cycles and digraphs
with self loops (where the reference's in-place update order is visible), S from 1 to 70 (one to three bit words)."""
import os, sys
import numpy as np

sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True
import utils as rutils  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def closure(a):
    r = a.astype(bool).copy()
    for k in range(r.shape[0]):
        r |= np.outer(r[:, k], r[k, :])
    return r.astype(np.int64)


rng = np.random.default_rng(2024)
out = {}
sizes = [1, 2, 3, 5, 8, 13, 20, 31, 32, 33, 40, 50, 64, 65, 70]
n = 0
for S in sizes:
    for kind in range(4):
        a = (rng.random((S, S)) < rng.choice([0.1, 0.25, 0.5])).astype(np.int64)
        if kind == 0:
            perm = rng.permutation(S)
            a = closure(np.triu(a, 1))[np.ix_(perm, perm)]      # closed DAG, nodes relabelled
        elif kind == 1:
            a = np.triu(a, 1)                                    # unclosed DAG
        elif kind == 2:
            np.fill_diagonal(a, 0)                               # digraph with cycles, no self loops
        b = (rng.random((S, S)) < 0.3).astype(np.int64)
        out[f"in{n}"] = a.astype(np.uint8)
        out[f"red{n}"] = rutils.transitive_reduction(a).astype(np.uint8)
        out[f"ham{n}"] = np.int64(rutils.hamming_distance(a, b))
        out[f"other{n}"] = b.astype(np.uint8)
        n += 1
out["n_cases"] = np.int64(n)
np.savez_compressed(os.path.join(HERE, "reduction.npz"), **out)
print("wrote", n, "cases")
