"""CPU-only tests: the C-ABI library loads and exports every declared symbol, host-side logic
(bit packing, drop-in host utilities, chain sharding + collectives over gloo, world size 2)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import nem_b200
from nem_b200 import _lib, engine, parallel, utils
from oracle import nem_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    _lib.build()
    L = _lib.lib()
    header = open(os.path.join(ROOT, "include", "nemb200.h")).read()
    declared = set(re.findall(r"\b(nemb200_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.nemb200_version() == 1
    assert L.nemb200_dataset_bytes(50, 10000) > 10000 * 128
    assert L.nemb200_dataset_bytes(129, 10) == 0  # S > 128 unsupported


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu():
    with pytest.raises(_lib.NemB200Error):
        nem_b200.Dataset(np.zeros((5, 8), dtype=np.uint8), -2.0, -2.0)
    with pytest.raises(_lib.NemB200Error):
        utils.ancestor(np.zeros((4, 4)))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "nem-mcmc-optimization_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle uses the same function", ""), f


@pytest.mark.parametrize("S", [1, 5, 31, 32, 33, 64, 100, 128])
def test_pack_unpack_roundtrip(S):
    rng = np.random.default_rng(S)
    a = (rng.random((7, S, S)) < 0.3).astype(np.uint8)
    bits = engine.pack_adj(a)
    assert bits.shape == (7, S, (S + 31) // 32) and bits.dtype == np.uint32
    assert np.array_equal(engine.unpack_adj(bits, S), a)
    i, j = S - 1, S // 2
    a[:] = 0; a[0, i, j] = 1
    assert engine.pack_adj(a)[0, i, j // 32] == np.uint32(1) << np.uint32(j % 32)


def test_host_utils_match_reference_semantics(golden):
    arr, meta = golden
    assert np.array_equal(utils.create_real_knockdown_mat(arr["kat_smat"], arr["kat_earr"]), arr["kat_kd"])
    for n in (0, 7, 12):
        info = meta["networks"][str(n)]
        real = utils.create_real_knockdown_mat(arr[f"adj{n}"], arr[f"end{n}"])
        assert np.array_equal(real, arr[f"real{n}"])
        assert np.array_equal(utils.create_observed_knockdown_mat(real, info["alpha"], info["beta"]), arr[f"D{n}"])
        assert np.array_equal(utils.initial_order_guess(arr[f"D{n}"]), arr[f"order0_{n}"])
    assert utils.hamming_distance(np.eye(3), np.zeros((3, 3))) == 3


def test_egene_storage_order_is_a_tile_homogeneous_permutation():
    """Dataset(sort_egenes=True): a permutation; every 128-E-gene tile holds a contiguous range of the sort by number
    of ones (homogeneous tiles), tiles alternate light / heavy, the ragged tail stays last."""
    rng = np.random.default_rng(0)
    for S, E in [(5, 100), (50, 1000), (100, 3000), (70, 128 * 7), (64, 129)]:
        d = (rng.random((S, E)) < rng.random(E)[None, :]).astype(np.uint8)
        order = engine.egene_storage_order(d)
        assert sorted(order.tolist()) == list(range(E))
        ones = d.sum(axis=0)[order]
        nfull = E // 128
        tiles = ones[:nfull * 128].reshape(nfull, 128) if nfull else np.zeros((0, 128))
        assert all((np.diff(t) >= 0).all() for t in tiles)                       # sorted inside a tile
        if nfull > 2:
            lo, hi = tiles.min(axis=1), tiles.max(axis=1)
            ranked = np.argsort(lo, kind="stable")
            assert all(hi[ranked[i]] <= lo[ranked[i + 1]] for i in range(nfull - 1))   # tiles do not overlap
            assert lo[0] == ones.min() and hi[1] == ones[:nfull * 128].max()    # lightest first, heaviest second
        if E % 128 and nfull:
            assert ones[nfull * 128:].min() >= tiles.max()                       # the ragged tail is the heaviest


def test_dot_text_matches_reference_format():
    """DAGs/dot.py:28-36: 'digraph {', one '    i -> j;' line per nonzero entry in row-major order, '}'."""
    a = np.array([[0, 1, 1], [0, 0, 2], [0, 0, 0]])
    assert utils.dot_string(a) == "digraph {\n    0 -> 1;\n    0 -> 2;\n    1 -> 2;\n}"
    assert utils.dot_string(np.zeros((2, 2))) == "digraph {\n}"


def test_read_csv(tmp_path, monkeypatch):
    p = tmp_path / "net.csv"
    p.write_text("3,4\n0,1\n1,2\n0,2\n0,1,2,2\n0.05,0.1\n")
    monkeypatch.chdir(tmp_path)
    adj, end_nodes, errors, S, E = utils.read_csv_to_adj("net.csv")
    assert (S, E) == (3, 4) and adj.tolist() == [[0, 1, 1], [0, 0, 1], [0, 0, 0]]
    assert end_nodes.tolist() == [0, 1, 2, 2] and errors.tolist() == [0.05, 0.1]


def test_start_orders_and_shards():
    from nem_b200.nem_order_mcmc import start_orders
    o = start_orders(np.arange(10)[::-1], 5, seed=3)
    assert np.array_equal(o[0], np.arange(10)[::-1])
    assert all(sorted(r) == list(range(10)) for r in o) and not np.array_equal(o[1], o[2])
    assert np.array_equal(o, start_orders(np.arange(10)[::-1], 5, seed=3))
    for n, w in [(4096, 8), (10, 4), (3, 8), (1024, 1)]:
        r = [parallel.shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n and all(r[k][1] == r[k + 1][0] for k in range(w - 1))
        assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


_GLOO_WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from nem_b200 import parallel
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
lo, hi = parallel.shard_range(7)
assert (lo, hi) == ((0, 4) if rank == 0 else (4, 7))
S, W = 5, 1
counts = torch.full((S, S), rank + 1, dtype=torch.int64)
parallel.allreduce_edge_counts(counts)
assert int(counts[0, 0]) == 3
n = hi - lo
best = torch.tensor([-10.0 - c for c in range(lo, hi)], dtype=torch.float64)
if rank == 1: best[1] = -1.0          # global best lives on rank 1, global chain id 5
orders = torch.arange(lo, hi, dtype=torch.int32)[:, None].repeat(1, S)
bits = torch.arange(lo, hi, dtype=torch.int32)[:, None, None].repeat(1, S, W)
score, chain, order, b = parallel.global_best(best, orders, bits, lo)
assert score == -1.0 and chain == 5 and int(order[0]) == 5 and int(b[0, 0]) == 5
assert parallel.allreduce_sum_scalar(float(rank + 1), "cpu") == 3.0
# SURVEY 8e: the data matrix travels once, rank 0 -> all (bit-packed)
rng = np.random.default_rng(5)
D0 = (rng.random((7, 45)) < 0.4).astype(np.uint8)
d, A, B = parallel.broadcast_dataset(D0 if rank == 0 else None, -2.89 if rank == 0 else None, -2.25 if rank == 0 else None, src=0, device="cpu")
assert d.shape == (7, 45) and np.array_equal(d, D0) and (A, B) == (-2.89, -2.25)
dist.destroy_process_group()
print("ok", rank)
"""


def test_gloo_world_size_2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"ok {r}" in o, o


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` (the CPU arm the driver launches) prints one JSON line with the contract keys."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-budget-s", "3"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "nem_graph_scores_per_s" and d["value"] > 0
    from oracle import ref_stage
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_stage.staged() else "port") and d["e2e"]["h2d_bytes_per_step"] == 0
    # same workload description as the GPU arm, no extrapolated step time
    assert d["config"]["S"] == 50 and d["config"]["E"] == 10000 and d["config"]["graphs_per_step_per_gpu"] == 1024
    assert abs(d["ms_per_step"] * 1e-3 * d["value"] - d["config"].get("reference_step_sample", 1024)) < 1e-6 * 1024


def test_reference_driver_matches_port_and_goldens():
    """oracle/_ref (the reference's own nem.py / utils.py, staged by oracle/ref_stage.py) driven as nem.py:132-142
    gives the same scores as the numpy restatement; skipped when the reference is not staged."""
    import pytest
    from oracle import ref_stage, nem_oracle as orc
    if not ref_stage.staged():
        pytest.skip("oracle/_ref not staged")
    from oracle import ref_driver
    _, _, D = orc.synth_dataset(12, 150, seed=3)
    A, B = orc.nem_constants(0.05, 0.1)
    rs = ref_driver.RefScorer(D, A, B)
    rng = np.random.default_rng(0)
    for _ in range(5):
        g = np.triu(rng.random((12, 12)) < 0.3, 1).astype(np.uint8)
        a, b = rs.score(g), orc.score_matrix(orc.closure_bool(g), D, A, B)
        assert abs(a - b) <= 1e-11 * abs(b)


def test_score_kernel_partition_invariants():
    """The flat (N tile, E tile) partition of the score kernel (csrc/score_tc.cu: Walker / k_finalize): host-only
    self-test over many problem sizes - every tile owned once, dense slots, at most 24 per N tile."""
    import nem_b200
    L = nem_b200._lib.lib()
    rng = np.random.default_rng(0)
    cases = [(1, 1, 148), (1, 79, 148), (256, 79, 148), (128, 79, 128), (205, 79, 148), (1024, 40, 148), (8192, 40, 148),
             (2, 782, 148), (16384, 782, 132), (3, 5, 7), (1000, 1, 148), (1, 1000, 148)]
    cases += [(int(rng.integers(1, 3000)), int(rng.integers(1, 900)), int(rng.integers(1, 160))) for _ in range(60)]
    for nt, nm, c in cases:
        assert L.nemb200_selftest_partition(nt, nm, c) == 0, (nt, nm, c)
    assert L.nemb200_selftest_partition(0, 1, 1) < 0


def test_hamming_distance_on_uint8_graphs():
    """the package hands out uint8 graphs; 0 - 1 must not wrap (utils.py:148-149 has int matrices)."""
    import nem_b200
    a = np.zeros((5, 5), dtype=np.uint8)
    b = np.zeros((5, 5), dtype=np.uint8)
    a[0, 1] = a[2, 3] = 1
    b[2, 3] = b[4, 0] = b[1, 1] = 1
    assert nem_b200.utils.hamming_distance(a, b) == 3 == nem_b200.utils.hamming_distance(b.astype(int), a.astype(int))
