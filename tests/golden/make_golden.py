"""Generate golden fixtures by RUNNING THE REFERENCE (only possible where /root/reference
is mounted).  Output: tests/golden/networks.npz + tests/golden/networks.json.

    PYTHONDONTWRITEBYTECODE=1 WANDB_MODE=disabled python tests/golden/make_golden.py

What is recorded (all produced by reference code, nothing by this repo's code):
  * per network N=0..19 (DAGs/networks/networkN/networkN.csv, networkN_red.csv):
      adjacency (closed and reduced), end nodes, errors, observed knock-down matrix
      (NEM.observed_knockdown_mat, MT19937 seed 42), NEM.U, NEM.A/B, NEM.obs_ll, NEM.real_ll,
      utils.ancestor(reduced) (float input)
  * the obs_ll values the reference's author logged
      (wandb/run-20240131_220740-wokieg7a/files/output.log) for cross-checking
  * reference scores of random order-compatible closed DAGs on a few networks, driven
      exactly as nem.py:132-142 drives get_reduced_score_tables/compute_ll_ratios/calculate_ll
  * utils.initial_order_guess, and NEMOrderMCMC.get_permissible_parents for one order
"""
import io, json, os, re, sys, contextlib
import numpy as np

REF = "/root/reference"
sys.path.insert(0, REF)
os.environ.setdefault("WANDB_MODE", "disabled")
sys.dont_write_bytecode = True
import nem as rnem      # noqa: E402
import utils as rutils  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def ref_score(model, adj, score_tables):
    """nem.py:132-142 verbatim call sequence for a binary graph."""
    S = model.num_s
    parents_list = [[] for _ in range(S)]
    weights = [[] for _ in range(S)]
    n_parents = np.zeros(S, dtype=int)
    for i in range(S):
        for j in range(S):
            if adj[i][j] == 1 and i != j:
                parents_list[j].append(i)
                n_parents[j] += 1
                weights[j].append(1.0)
    red = model.get_reduced_score_tables(score_tables, parents_list)
    _, ll = model.calculate_ll(model.compute_ll_ratios(weights, red, n_parents, model.U))
    return float(ll)


def main():
    logged = {}
    log = open(f"{REF}/wandb/run-20240131_220740-wokieg7a/files/output.log").read().splitlines()
    k = 0
    for line in log:
        m = re.match(r"Comparison: real_ll: (\S+) - inferred_ll", line)
        if m:
            logged[k] = float(m.group(1)); k += 1
    arrays, meta = {}, {"logged_obs_ll": logged, "networks": {}}
    os.chdir(REF)
    rng = np.random.default_rng(20261018)
    for n in range(20):
        adj, end_nodes, errors, S, E = rutils.read_csv_to_adj(f"DAGs/networks/network{n}/network{n}.csv")
        red, *_ = rutils.read_csv_to_adj(f"DAGs/networks/network{n}/network{n}_red.csv")
        with contextlib.redirect_stdout(io.StringIO()):
            model = rnem.NEM(adj.copy(), end_nodes, errors, S, E, seed=42)
        anc = rutils.ancestor(red.astype(float))
        arrays[f"adj{n}"] = adj.astype(np.uint8)
        arrays[f"red{n}"] = red.astype(np.uint8)
        arrays[f"anc_red{n}"] = anc.astype(np.uint8)
        arrays[f"end{n}"] = end_nodes.astype(np.int32)
        arrays[f"D{n}"] = model.observed_knockdown_mat.astype(np.uint8)
        arrays[f"real{n}"] = model.real_knockdown_mat.astype(np.uint8)
        arrays[f"U{n}"] = model.U
        arrays[f"order0_{n}"] = rutils.initial_order_guess(model.observed_knockdown_mat).astype(np.int32)
        info = dict(S=int(S), E=int(E), alpha=float(errors[0]), beta=float(errors[1]),
                    A=float(model.A), B=float(model.B), obs_ll=float(model.obs_ll),
                    real_ll=float(model.real_ll))
        # random order-compatible closed DAGs scored by the reference (a few networks, kept small)
        if n in (0, 5, 11, 12, 19):
            tabs = model.get_score_tables(model.observed_knockdown_mat)
            graphs, scores = [], []
            for _ in range(6):
                order = rng.permutation(S)
                pos = np.empty(S, dtype=int); pos[order] = np.arange(S)
                g = ((pos[:, None] < pos[None, :]) & (rng.random((S, S)) < 0.25)).astype(float)
                g = rutils.ancestor(g)
                graphs.append(g.astype(np.uint8)); scores.append(ref_score(model, g, tabs))
            # the empty graph and the full order graph as edge cases
            g0 = np.zeros((S, S), dtype=int)
            graphs.append(g0.astype(np.uint8)); scores.append(ref_score(model, g0, tabs))
            order = rng.permutation(S); pos = np.empty(S, dtype=int); pos[order] = np.arange(S)
            g1 = (pos[:, None] < pos[None, :]).astype(int)
            graphs.append(g1.astype(np.uint8)); scores.append(ref_score(model, g1, tabs))
            arrays[f"rg{n}"] = np.stack(graphs)
            info["rg_scores"] = scores
        meta["networks"][str(n)] = info
        print(n, S, E, info["obs_ll"], logged.get(n), flush=True)
    # knock-down known-answer test, tests/utils.tests.py:11-27 (run through the reference)
    s_mat = [[0,1,1,0,1,0],[0,0,1,0,1,0],[0,0,0,0,1,0],[0,0,1,0,1,0],[0,0,0,0,0,0],[0,0,0,0,1,0]]
    arrays["kat_smat"] = np.array(s_mat, dtype=np.uint8)
    arrays["kat_earr"] = np.array([0,1,2,3,4,5,0], dtype=np.int32)
    arrays["kat_kd"] = rutils.create_real_knockdown_mat(s_mat, [0,1,2,3,4,5,0]).astype(np.uint8)
    # dense closure case: float ancestor on the full upper-triangular DAG at S=70 (Q8)
    tri = np.triu(np.ones((70, 70)), 1)
    arrays["anc_tri70"] = rutils.ancestor(tri).astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "networks.npz"), **arrays)
    with open(os.path.join(HERE, "networks.json"), "w") as f:
        json.dump(meta, f, indent=1)


if __name__ == "__main__":
    main()
