"""Golden fixtures for the per-order edge-weight fit, produced by RUNNING THE REFERENCE
(only possible where /root/reference is mounted):

    PYTHONDONTWRITEBYTECODE=1 WANDB_MODE=disabled python tests/golden/make_golden_fit.py

For a few shipped networks and orders it drives NEMOrderMCMC.get_optimal_weights(init=True, max_iter=1)
(nem_order_mcmc.py:172-208; the per-edge scipy L-BFGS-B fit of :160-170 with the objective of :18-23) exactly as
NEMOrderMCMC.method does at :258/:277 and records, all from reference code:
  weights before the pass, posterior attachment weights, ancestor_x, the c vectors' minimisers (weights after
  the pass), the thresholded DAG / NEM and the returned dag_ll.
Output: tests/golden/fit.npz.  The scipy version that produced it is stored too (the fit is an iterative
optimiser stopped at tol=0.01, so the exact iterates depend on it - SURVEY 8c: "parity unpinned" upstream)."""
import contextlib, io, os, sys
import numpy as np

REF = "/root/reference"
sys.path.insert(0, REF)
os.environ.setdefault("WANDB_MODE", "disabled")
sys.dont_write_bytecode = True
import scipy  # noqa: E402
import nem as rnem  # noqa: E402
import utils as rutils  # noqa: E402
import nem_order_mcmc as rmc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    os.chdir(REF)
    out = {"scipy_version": np.array(scipy.__version__)}
    rng = np.random.default_rng(7)
    cases = []
    for n in (12, 0, 14):
        adj, end_nodes, errors, S, E = rutils.read_csv_to_adj(f"DAGs/networks/network{n}/network{n}.csv")
        with contextlib.redirect_stdout(io.StringIO()):
            model = rnem.NEM(adj.copy(), end_nodes, errors, S, E, seed=42)
        for rep in range(2):
            order = rutils.initial_order_guess(model.observed_knockdown_mat) if rep == 0 else rng.permutation(S)
            with contextlib.redirect_stdout(io.StringIO()):
                mc = rmc.NEMOrderMCMC(model, np.array(order))
                # the state method() has at :277: weights reset to init_value 0.5 on every permissible pair
                mc.reset(np.array(order), init=True)
                w_before = mc.parent_weights.copy()
                dag_ll = mc.get_optimal_weights(init=True, use_nem=False)
                w_after = mc.parent_weights.copy()
                dag, _ = mc.create_dag(mc.parent_weights)
                nemg, _ = mc.create_nem(mc.parent_weights)
                # second pass from the fitted weights (non-uniform start)
                dag_ll2 = mc.get_optimal_weights(init=True, use_nem=True)
                w_after2 = mc.parent_weights.copy()
            tag = f"n{n}_r{rep}"
            cases.append(tag)
            out[f"D_{tag}"] = model.observed_knockdown_mat.astype(np.uint8)
            out[f"AB_{tag}"] = np.array([model.A, model.B])
            out[f"order_{tag}"] = np.asarray(order, dtype=np.int32)
            out[f"w0_{tag}"] = w_before
            out[f"w1_{tag}"] = w_after
            out[f"w2_{tag}"] = w_after2
            out[f"dag_{tag}"] = np.asarray(dag, dtype=np.uint8)
            out[f"nem_{tag}"] = np.asarray(nemg, dtype=np.uint8)
            out[f"ll_{tag}"] = np.array([dag_ll, dag_ll2])
            print(tag, "S", S, "E", E, "dag_ll", dag_ll, dag_ll2, "edges", int(np.sum(dag)), "w range", w_after[w_after > 0].min(), w_after.max())
    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "fit.npz"), **out)


if __name__ == "__main__":
    main()
