"""GPU parity tests (run on the B200 box): CUDA path through the C ABI vs the CPU oracle on the same
seeded inputs, the reference's golden fixtures, and size-independent properties at full sizes."""
import itertools

import numpy as np
import pytest
import torch

from oracle import c_oracle, nem_oracle as orc

pytestmark = pytest.mark.gpu

TC_IMPLS = ["lut", "ex2", "ex2_2pass"]
REL_TC = 1e-6       # fp32 epilogue; north-star tolerance is 1e-4 * |score|
REL_WITNESS = 1e-11  # fp64 popcount witness


@pytest.fixture(scope="module")
def nb():
    import nem_b200
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return nem_b200


def rand_dags(rng, n, S, p=0.25, closed=True):
    out = np.zeros((n, S, S), dtype=np.uint8)
    for g in range(n):
        order = rng.permutation(S)
        pos = np.empty(S, dtype=int); pos[order] = np.arange(S)
        a = ((pos[:, None] < pos[None, :]) & (rng.random((S, S)) < p)).astype(np.uint8)
        out[g] = c_oracle.closure(a) if closed else a
    return out


# ------------------------------------------------------------------------------------------ closure
@pytest.mark.parametrize("n", range(20))
def test_closure_golden_networks(nb, golden, n):
    arr, _ = golden
    got = nb.closure(arr[f"red{n}"][None])[0]
    assert np.array_equal(got, arr[f"adj{n}"])
    assert np.array_equal(nb.utils.ancestor(arr[f"red{n}"]), arr[f"anc_red{n}"])


@pytest.mark.parametrize("S", [1, 2, 9, 10, 20, 30, 31, 32, 33, 50, 64, 65, 96, 100, 128])
def test_closure_random_vs_oracle(nb, S):
    rng = np.random.default_rng(S)
    n = 300
    dags = rand_dags(rng, n, S, p=0.15, closed=False)
    cyc = (rng.random((40, S, S)) < 0.05).astype(np.uint8)  # general digraphs (cycles allowed)
    for g in cyc:
        np.fill_diagonal(g, 0)
    batch = np.concatenate([dags, cyc])
    got = nb.closure(batch)
    for g in range(len(batch)):
        assert np.array_equal(got[g], c_oracle.closure(batch[g])), (S, g)
    assert np.array_equal(nb.closure(got), got)  # idempotent


def test_closure_dense_and_empty(nb, golden):
    arr, _ = golden
    tri = np.triu(np.ones((70, 70), dtype=np.uint8), 1)
    assert np.array_equal(nb.closure(tri[None])[0], arr["anc_tri70"])
    chain = np.zeros((100, 100), dtype=np.uint8)
    chain[np.arange(99), np.arange(1, 100)] = 1
    assert np.array_equal(nb.closure(chain[None])[0], np.triu(np.ones((100, 100), dtype=np.uint8), 1))
    assert nb.closure(np.zeros((0, 10, 10), dtype=np.uint8), S=10).shape == (0, 10, 10)


@pytest.mark.parametrize("S", [1, 2, 7, 16, 30, 31, 32])
def test_closure_thread_per_graph_kernel(nb, S):
    """Batches >= 8192 of graphs with S <= 32 take the one-thread-per-graph kernel: same bits as the oracle
    (and as the warp-per-graph kernel used for smaller batches), in place as well."""
    rng = np.random.default_rng(100 + S)
    n = 8192 + 77
    a = (rng.random((n, S, S)) < 0.08).astype(np.uint8)  # general digraphs, self loops and cycles included
    bits = nb.pack_adj(a)
    dev = torch.from_numpy(bits.view(np.int32)).cuda()
    big = nb.closure(dev, S=S, return_bits=True)
    small = torch.cat([nb.closure(dev[i:i + 2048].contiguous(), S=S, return_bits=True) for i in range(0, n, 2048)])
    assert torch.equal(big, small)
    got = nb.unpack_adj(big.cpu().numpy().view(np.uint32), S)
    for g in rng.integers(0, n, 64):
        assert np.array_equal(got[g], c_oracle.closure(a[g])), (S, g)
    inplace = dev.clone()
    import ctypes as C
    nb._lib.check(nb._lib.lib().nemb200_closure_u32(C.c_void_p(inplace.data_ptr()), n, S, C.c_void_p(inplace.data_ptr()),
                                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    assert torch.equal(inplace, big)


def test_closure_large_batch_properties(nb):
    """BASELINE config 5 size: 65536 random DAGs at S=30: idempotent, superset, acyclic-preserving."""
    rng = np.random.default_rng(5)
    S, n = 30, 65536
    pos = np.argsort(rng.random((n, S)), axis=1)
    a = ((pos[:, :, None] < pos[:, None, :]) & (rng.random((n, S, S)) < 0.12)).astype(np.uint8)
    bits = nb.pack_adj(a)
    c1 = nb.closure(bits, S=S, return_bits=True)
    c2 = nb.closure(c1, S=S, return_bits=True)
    assert torch.equal(c1, c2)
    h = c1.cpu().numpy().view(np.uint32)
    assert np.all((h & bits) == bits)
    closed = nb.unpack_adj(h, S)
    assert not closed[:, np.arange(S), np.arange(S)].any()
    for g in rng.integers(0, n, 50):
        assert np.array_equal(closed[g], c_oracle.closure(a[g]))


# -------------------------------------------------------------------------------------------- score
@pytest.mark.parametrize("n", range(20))
def test_score_golden_obs_ll(nb, golden, n):
    arr, meta = golden
    info = meta["networks"][str(n)]
    ds = nb.Dataset(arr[f"D{n}"], info["A"], info["B"])
    adj = arr[f"adj{n}"][None]
    assert ds.score_graphs(adj, impl="witness")[0] == pytest.approx(info["obs_ll"], rel=REL_WITNESS)
    for impl in TC_IMPLS:
        assert ds.score_graphs(adj, impl=impl)[0] == pytest.approx(info["obs_ll"], rel=REL_TC), impl
    assert ds.score_graphs(arr[f"red{n}"][None], close=True)[0] == pytest.approx(info["obs_ll"], rel=REL_TC)


@pytest.mark.parametrize("n", [0, 5, 11, 12, 19])
def test_score_golden_random_graphs(nb, golden, n):
    arr, meta = golden
    info = meta["networks"][str(n)]
    ds = nb.Dataset(arr[f"D{n}"], info["A"], info["B"])
    ref = np.array(info["rg_scores"])
    assert np.allclose(ds.score_graphs(arr[f"rg{n}"], impl="witness"), ref, rtol=REL_WITNESS)
    for impl in TC_IMPLS:
        assert np.allclose(ds.score_graphs(arr[f"rg{n}"], impl=impl), ref, rtol=REL_TC), impl


def test_score_noise_free_small_score(nb, golden):
    """Noise-free data: the score is a small difference of large sums (network11: +14.4 vs |sum N| in the
    thousands) - the north-star tolerance 1e-4*|score| must still hold."""
    arr, meta = golden
    info = meta["networks"]["11"]
    ds = nb.Dataset(arr["real11"], info["A"], info["B"])
    ref = c_oracle.score(arr["adj11"], arr["real11"], info["A"], info["B"])
    for impl in ["witness"] + TC_IMPLS:
        assert ds.score_graphs(arr["adj11"][None], impl=impl)[0] == pytest.approx(ref, abs=1e-4 * abs(ref)), impl


@pytest.mark.parametrize("S,E,n", [(8, 40, 7), (10, 100, 33), (16, 257, 17), (20, 1000, 64), (30, 5000, 41),
                                   (50, 1000, 23), (50, 10000, 12), (63, 700, 9), (64, 300, 9), (77, 900, 11),
                                   (100, 1500, 7), (118, 640, 6), (119, 200, 4), (125, 333, 5), (126, 200, 3), (128, 260, 5)])
def test_score_synthetic_vs_oracle(nb, S, E, n):
    rng = np.random.default_rng(S * 1000 + E)
    _, _, D = orc.synth_dataset(S, E, seed=S + E)
    A, B = orc.nem_constants(0.05, 0.1)
    graphs = rand_dags(rng, n, S, p=0.3)
    graphs[0] = 0                                                    # empty graph
    graphs[1] = orc.order_to_graph(rng.permutation(S))               # full order graph
    ref = c_oracle.score_batch(graphs, D, A, B, nthreads=8)
    ds = nb.Dataset(D, A, B)
    assert np.allclose(ds.score_graphs(graphs, impl="witness"), ref, rtol=REL_WITNESS)
    impls = [i for i in TC_IMPLS if (i != "lut" or ds.lut_valid) and (i != "ex2" or S <= 54)]
    assert impls
    assert (ds.lutm_levels > 0) == (16 <= S <= 125)
    if ds.lutm_levels:
        impls.append("lutm")
    for impl in impls:
        got = ds.score_graphs(graphs, impl=impl)
        assert np.allclose(got, ref, rtol=REL_TC), (impl, np.abs(got / ref - 1).max())
    assert np.allclose(ds.score_graphs(graphs), ref, rtol=REL_TC)    # auto


@pytest.mark.parametrize("E", [1, 2, 127, 128, 129, 255, 256, 257])
def test_score_ragged_E(nb, E):
    S = 12
    rng = np.random.default_rng(E)
    D = (rng.random((S, E)) < 0.4).astype(np.uint8)
    A, B = orc.nem_constants(0.05, 0.1)
    graphs = rand_dags(rng, 19, S)
    ref = c_oracle.score_batch(graphs, D, A, B)
    ds = nb.Dataset(D, A, B)
    for impl in ["witness"] + TC_IMPLS:
        assert np.allclose(ds.score_graphs(graphs, impl=impl), ref, rtol=REL_TC), impl
    assert ds.score_graphs(graphs[:0]).shape == (0,)
    assert np.allclose(ds.score_graphs(graphs[:1]), ref[:1], rtol=REL_TC)


def test_score_mix_only_at_stride_50(nb):
    """NEMB200_SCORE_TC_LUT_MIX is refused, loudly, for any column stride but 50."""
    _, _, D = orc.synth_dataset(30, 500, seed=1)
    ds = nb.Dataset(D, *orc.nem_constants(0.05, 0.1))
    bits = ds.order_graphs(np.stack([np.random.default_rng(0).permutation(30) for _ in range(4)]).astype(np.int32))
    with pytest.raises(RuntimeError, match="mix needs the column stride 50"):
        ds.score_bits(bits, impl="lut_mix")


@pytest.mark.parametrize("S,E,n", [(50, 3000, 600), (30, 2000, 512), (47, 1500, 300), (100, 900, 64)])
def test_dependent_launch_back_to_back_calls(nb, S, E, n):
    """The launches of a score call are dependent launches, and with NEMB200_SCORE_INPUTS_READY (score_bits(inputs_ready=True))
    a call works ahead while the previous call's last kernel drains: forty calls back to back on one workspace, alternating
    between three different batches (closed and unclosed), no synchronisation in between - every result must equal the
    result of the same call made alone."""
    _, _, D = orc.synth_dataset(S, E, seed=S)
    ds = nb.Dataset(D, *orc.nem_constants(0.05, 0.1))
    rng = np.random.default_rng(S)
    batches = []
    for b in range(3):
        raw = np.stack([orc.order_to_graph(rng.permutation(S), 1, rng.random((S, S)) < 0.2, False) for _ in range(n)]).astype(np.uint8)
        if b == 2:
            raw[::7] = (rng.random((len(raw[::7]), S, S)) < 0.2).astype(np.uint8)   # digraphs the pair planner gives up on
        batches.append(torch.from_numpy(nb.pack_adj(raw).view(np.int32)).cuda())
    alone = [ds.score_bits(b, close=True).clone() for b in batches]
    torch.cuda.synchronize()
    outs = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(40)]
    for k in range(40):
        ds.score_bits(batches[k % 3], close=True, out=outs[k], inputs_ready=True)
    torch.cuda.synchronize()
    for k in range(40):
        assert torch.equal(outs[k], alone[k % 3]), k
    # without the promise: the graphs are produced by the kernel right before the call (an element-wise torch kernel into a
    # buffer that is rewritten every time), promised and unpromised calls interleaved
    buf = torch.empty_like(batches[0])
    for k in range(24):
        if k % 2:
            torch.bitwise_xor(batches[k % 3], 0, out=buf)
            ds.score_bits(buf, close=True, out=outs[k])
        else:
            ds.score_bits(batches[k % 3], close=True, out=outs[k], inputs_ready=True)
    torch.cuda.synchronize()
    for k in range(24):
        assert torch.equal(outs[k], alone[k % 3]), k
    w = ds.score_bits(nb.closure(batches[2], S=S, return_bits=True), impl="witness")
    assert torch.allclose(alone[2], w, rtol=REL_TC, atol=0)


def test_score_other_error_rates_and_unsupported(nb):
    S, E = 40, 300
    rng = np.random.default_rng(9)
    D = (rng.random((S, E)) < 0.5).astype(np.uint8)
    graphs = rand_dags(rng, 10, S)
    for alpha, beta in [(0.01, 0.01), (0.2, 0.3), (0.001, 0.4)]:
        A, B = orc.nem_constants(alpha, beta)
        ds = nb.Dataset(D, A, B)
        ref = c_oracle.score_batch(graphs, D, A, B)
        assert np.allclose(ds.score_graphs(graphs), ref, rtol=REL_TC), (alpha, beta)
        assert np.allclose(ds.score_graphs(graphs, impl="ex2_2pass"), ref, rtol=REL_TC)
    with pytest.raises(nb._lib.NemB200Error):
        nb.Dataset(np.zeros((129, 10), dtype=np.uint8), -1.0, -1.0)
    ds = nb.Dataset(np.zeros((100, 10), dtype=np.uint8), *orc.nem_constants(0.05, 0.1))
    with pytest.raises(nb._lib.NemB200Error):
        ds.score_graphs(np.zeros((1, 100, 100), dtype=np.uint8), impl="lut")   # table epilogue needs S <= 63


@pytest.mark.parametrize("S", [17, 40, 56, 63, 64, 100, 118, 125])
def test_score_multilevel_table_extremes(nb, S):
    """Multi-level table epilogue (csrc/score_tc.cu: graph_lse_levels): every E-gene's terms span hundreds of binary
    orders, and which level is used depends on the E-gene.  Extreme data (all ones: the largest possible terms, all
    zeros: every term below the null site, one-hot rows) x extreme graphs (empty, complete order, all ones), and
    error rates that need two or three levels or rule the table out."""
    rng = np.random.default_rng(S)
    E = 517
    D = (rng.random((S, E)) < 0.5).astype(np.uint8)
    D[:, :64] = 1; D[:, 64:128] = 0
    D[:, 128:192] = (rng.random((S, 64)) < 0.95); D[:, 192:256] = (rng.random((S, 64)) < 0.05)
    D[:, 256:256 + S] = np.eye(S, dtype=np.uint8)
    graphs = np.concatenate([rand_dags(rng, 9, S, p=0.3), rand_dags(rng, 4, S, p=0.02),
                             np.stack([orc.order_to_graph(rng.permutation(S)) for _ in range(3)]).astype(np.uint8),
                             np.zeros((1, S, S), dtype=np.uint8), np.ones((1, S, S), dtype=np.uint8)])
    seen = set()
    for alpha, beta in [(0.05, 0.1), (0.02, 0.1), (0.01, 0.01), (0.2, 0.3), (0.45, 0.001), (0.6, 0.7)]:
        A, B = orc.nem_constants(alpha, beta)
        ds = nb.Dataset(D, A, B)
        w = ds.score_graphs(graphs, impl="witness")
        seen.add(ds.lutm_levels)
        if ds.lutm_levels:
            # E-genes stored sorted: homogeneous tiles, most of which skip the upper table levels (k_tile_levels)
            ds_sorted = nb.Dataset(D, A, B, sort_egenes=True)
            got_s = ds_sorted.score_graphs(graphs, impl="lutm")
            assert np.allclose(got_s, w, rtol=REL_TC), (alpha, beta, np.abs(got_s / w - 1).max())
        if ds.lutm_levels == 0:
            with pytest.raises(nb._lib.NemB200Error):
                ds.score_graphs(graphs, impl="lutm")
        else:
            got = ds.score_graphs(graphs, impl="lutm")
            assert np.allclose(got, w, rtol=REL_TC), (alpha, beta, ds.lutm_levels, np.abs(got / w - 1).max())
            for k in (1, 3):                                               # partial N tiles
                assert np.array_equal(ds.score_graphs(graphs[:k], impl="lutm"), got[:k])
        assert np.allclose(ds.score_graphs(graphs), w, rtol=REL_TC)       # auto, whatever it picks
        ref = c_oracle.score_batch(graphs[:3], D, A, B, nthreads=3)
        assert np.allclose(w[:3], ref, rtol=REL_WITNESS)
    assert {2, 3} <= seen or S < 64, seen


@pytest.mark.parametrize("S", [33, 40, 47, 50, 52])
def test_score_pair_merged_kernel(nb, S):
    """Pair-merged table epilogue (two attachment sites per look-up, csrc/score_tc.cu: k_pair_plan): closed DAGs,
    sparse and dense graphs, full orders, non-closed and cyclic digraphs.  Graphs the planner cannot pair go
    through the one-site kernel via the device-side fallback list - both routes must agree with the witness."""
    E = 1500
    rng = np.random.default_rng(S)
    _, _, D = orc.synth_dataset(S, E, seed=S)
    A, B = orc.nem_constants(0.05, 0.1)
    ds = nb.Dataset(D, A, B)
    assert ds.lut_valid
    parts = [rand_dags(rng, 40, S, p=0.3), rand_dags(rng, 40, S, p=0.05), rand_dags(rng, 20, S, p=0.6, closed=False),
             (rng.random((30, S, S)) < 0.3).astype(np.uint8),            # arbitrary digraphs: mostly unpairable
             np.stack([orc.order_to_graph(rng.permutation(S)) for _ in range(10)]).astype(np.uint8),
             np.zeros((3, S, S), dtype=np.uint8), np.ones((2, S, S), dtype=np.uint8)]
    graphs = np.concatenate(parts)
    perm = rng.permutation(len(graphs))
    graphs = graphs[perm]
    w = ds.score_graphs(graphs, impl="witness")
    got = ds.score_graphs(graphs, impl="pair")
    assert np.allclose(got, w, rtol=REL_TC), np.abs(got / w - 1).max()
    n_fb = nb._lib.lib().nemb200_pair_fallback_count(ds._h, ds._score_ws.data_ptr(), len(graphs))   # graphs that took the fallback route
    assert (0 < n_fb or S < 47) and n_fb < len(graphs), n_fb
    assert np.allclose(ds.score_graphs(graphs, impl="lut"), w, rtol=REL_TC)
    ref = c_oracle.score_batch(graphs[:8], D, A, B, nthreads=4)
    assert np.allclose(got[:8], ref, rtol=REL_TC)
    # single graphs and ragged batch sizes (partial last tile)
    for k in (1, 7, 9):
        assert np.allclose(ds.score_graphs(graphs[:k], impl="pair"), w[:k], rtol=REL_TC)


def test_score_full_size_c3_lut_vs_witness(nb):
    """BASELINE config 3 shape (S=50, E=10k, 1024 graphs): tensor-core path vs the fp64 witness kernel
    on the device, plus permutation invariance of the batch order."""
    S, E, n = 50, 10000, 1024
    rng = np.random.default_rng(3)
    _, _, D = orc.synth_dataset(S, E, seed=0)
    A, B = orc.nem_constants(0.05, 0.1)
    ds = nb.Dataset(D, A, B)
    orders = np.stack([rng.permutation(S) for _ in range(n)]).astype(np.int32)
    bits = ds.order_graphs(orders)
    w = ds.score_bits(bits, impl="witness").cpu().numpy()
    for impl in TC_IMPLS + ["pair", "lut_mix"]:
        got = ds.score_bits(bits, impl=impl).cpu().numpy()
        assert np.allclose(got, w, rtol=REL_TC), impl
    # the sweep entry's AUTO is the pair-merged path at this shape, the MCMC loop's AUTO the one-site kernel with the ex2 mix
    assert np.array_equal(ds.score_bits(bits).cpu().numpy(), ds.score_bits(bits, impl="pair").cpu().numpy())
    st = ds.run_chains(orders[:128], 3, 0.01, seed=5, impl="auto")
    assert np.array_equal(st.curr_scores.cpu().numpy(), ds.score_bits(st.curr_bits, impl="lut_mix").cpu().numpy())
    perm = torch.from_numpy(rng.permutation(n)).to(bits.device)
    got_p = ds.score_bits(bits[perm].contiguous()).cpu().numpy()
    assert np.array_equal(got_p, ds.score_bits(bits).cpu().numpy()[perm.cpu().numpy()])  # deterministic
    for g in (0, 511, 1023):
        assert w[g] == pytest.approx(c_oracle.score(orc.order_to_graph(orders[g]), D, A, B), rel=REL_WITNESS)


def test_score_config4_shape_vs_witness(nb):
    """BASELINE config 4 shape (S=100, E=100 000): max-subtracted ex2 epilogue vs the fp64 witness on the device
    and the C oracle for two graphs."""
    S, E, n = 100, 100000, 48
    rng = np.random.default_rng(4)
    _, _, D = orc.synth_dataset(S, E, seed=4)
    A, B = orc.nem_constants(0.05, 0.1)
    ds = nb.Dataset(D, A, B)
    assert not ds.lut_valid
    cand = (rng.random((S, S)) < 0.2).astype(np.uint8)
    orders = np.stack([rng.permutation(S) for _ in range(n)]).astype(np.int32)
    bits = ds.order_graphs(orders, rule=1, cand=cand, use_nem=True)
    w = ds.score_bits(bits, impl="witness").cpu().numpy()
    assert ds.lutm_levels == 2                                         # auto = the two-level table epilogue
    for impl in ("auto", "lutm", "ex2_2pass"):
        got = ds.score_bits(bits, impl=impl).cpu().numpy()
        assert np.allclose(got, w, rtol=REL_TC), (impl, np.abs(got / w - 1).max())
    graphs = nb.unpack_adj(bits[:2].cpu().numpy().view(np.uint32), S)
    assert np.allclose(w[:2], c_oracle.score_batch(graphs, D, A, B, nthreads=2), rtol=REL_WITNESS)


def test_score_config5_sweep(nb):
    """BASELINE config 5: 65 536 random DAGs at S=30, E=5000 - closure + score of the whole batch; a slice is
    checked against the witness, the oracle, and the batch result must not depend on batch composition."""
    S, E, n = 30, 5000, 65536
    _, _, D = orc.synth_dataset(S, E, seed=5)
    A, B = orc.nem_constants(0.05, 0.1)
    ds = nb.Dataset(D, A, B)
    raw = nb.synth.random_dag_bits(n, S, seed=1)
    dev = torch.from_numpy(raw.view(np.int32)).cuda()
    closed = nb.closure(dev, S=S, return_bits=True)
    scores = ds.score_bits(closed).cpu().numpy()
    assert np.isfinite(scores).all()
    sl = slice(1000, 1000 + 2048)
    w = ds.score_bits(closed[sl].contiguous(), impl="witness").cpu().numpy()
    assert np.allclose(scores[sl], w, rtol=REL_TC)
    assert np.array_equal(ds.score_bits(closed[sl].contiguous()).cpu().numpy(), scores[sl])  # independent of the batch
    g = nb.unpack_adj(raw[:3], S)
    ref = [c_oracle.score(c_oracle.closure(x), D, A, B) for x in g]
    assert np.allclose(scores[:3], ref, rtol=REL_TC)


# ------------------------------------------------------------------------------------ order -> graph
@pytest.mark.parametrize("S", [2, 10, 33, 50, 100])
def test_order_to_graph_vs_oracle(nb, S):
    rng = np.random.default_rng(S)
    D = (rng.random((S, 50)) < 0.5).astype(np.uint8)
    ds = nb.Dataset(D, *orc.nem_constants(0.05, 0.1))
    orders = np.stack([rng.permutation(S) for _ in range(37)]).astype(np.int32)
    cand = (rng.random((S, S)) < 0.3).astype(np.uint8)
    for rule, use_nem in itertools.product((0, 1), (False, True)):
        bits = ds.order_graphs(orders, rule=rule, cand=cand, use_nem=use_nem)
        got = nb.unpack_adj(bits.cpu().numpy().view(np.uint32), S)
        for c in range(len(orders)):
            assert np.array_equal(got[c], orc.order_to_graph(orders[c], rule, cand, use_nem)), (rule, use_nem, c)


@pytest.mark.parametrize("S,E,n,impl", [(10, 100, 37, "auto"), (20, 300, 64, "auto"), (50, 1000, 130, "lut"), (50, 1000, 9, "ex2_2pass"),
                                        (50, 1000, 130, "pair"), (50, 1000, 130, "auto"), (40, 500, 66, "pair"), (33, 257, 21, "auto"), (64, 300, 17, "auto"), (100, 700, 23, "lutm"), (128, 200, 9, "auto")])
def test_closure_fused_into_score_kernel(nb, S, E, n, impl):
    """nemb200_closure_score_batch (closure inside the score kernel's operand expansion, one launch - or, for the pair-merged
    path, inside its planner) == closure kernel + score kernel == the oracle's closure + score, on unclosed DAGs and on digraphs with cycles; the input is not modified."""
    _, _, D = orc.synth_dataset(S, E, seed=S)
    A, B = orc.nem_constants(0.05, 0.1)
    rng = np.random.default_rng(S + n)
    raw = np.stack([orc.order_to_graph(rng.permutation(S), 1, rng.random((S, S)) < 0.15, False) for _ in range(n)]).astype(np.uint8)
    raw[-1] = (rng.random((S, S)) < 0.03).astype(np.uint8)          # an arbitrary digraph (cycles)
    ds = nb.Dataset(D, A, B)
    bits = torch.from_numpy(nb.pack_adj(raw).view(np.int32)).cuda()
    keep = bits.clone()
    fused = ds.score_bits(bits, impl=impl, close=True).cpu().numpy()
    assert torch.equal(bits, keep)
    two = ds.score_bits(nb.closure(bits, S=S, return_bits=True), impl=impl).cpu().numpy()
    assert np.array_equal(fused, two)                                # same kernel, same operand: bit-identical
    ref = c_oracle.score_batch(np.stack([c_oracle.closure(g) for g in raw]), D, A, B)
    assert np.allclose(fused, ref, rtol=1e-6)
    assert np.allclose(ds.score_graphs(raw, close=True, impl=impl), ref, rtol=1e-6)
    # the implementations that cannot fuse fall back to closure + score
    assert np.allclose(ds.score_graphs(raw, close=True, impl="witness"), ref, rtol=1e-11)


# --------------------------------------------------------------------------------------------- MCMC
def _compare_chain(res, c, ref, S, tol=1e-6):
    assert res.accept_counts[c] == ref["n_accept"]
    assert np.array_equal(res.curr_orders[c], ref["curr_order"])
    assert np.array_equal(res.best_orders[c], ref["best_order"])
    assert np.array_equal(res.best_graphs[c], ref["best_graph"])
    assert res.best_scores[c] == pytest.approx(ref["best_score"], rel=tol)
    assert res.curr_scores[c] == pytest.approx(ref["curr_score"], rel=tol)


@pytest.mark.parametrize("impl", ["fused", "lut", "witness"])
@pytest.mark.parametrize("S,E,rule,use_nem", [(10, 100, 0, True), (13, 126, 1, True), (13, 126, 1, False),
                                              (20, 300, 0, False), (40, 200, 1, True)])
def test_mcmc_trajectory_vs_oracle(nb, S, E, rule, use_nem, impl):
    """Same Philox draws as the oracle -> identical accept decisions, orders, best graphs and edge counts
    (a decision can only differ if |u - rate| is below the fp32-epilogue error; seeds are fixed)."""
    _, _, D = orc.synth_dataset(S, E, seed=7)
    A, B = orc.nem_constants(0.05, 0.1)
    rng = np.random.default_rng(1)
    cand = (rng.random((S, S)) < 0.4).astype(np.uint8)
    n_chains, n_iter, burn = 6, 400, 50
    gamma = 2.0 * S / E
    orders0 = np.stack([rng.permutation(S) for _ in range(n_chains)]).astype(np.int32)
    ds = nb.Dataset(D, A, B)
    st = ds.run_chains(orders0, n_iter, gamma, swap_prob=0.9, seed=99, chain_offset=3, rule=rule, cand=cand,
                       use_nem=use_nem, burn_in=burn, trace=True, impl=impl)
    res = st.result()
    total = np.zeros((S, S), dtype=np.int64)
    for c in range(n_chains):
        ref = c_oracle.mcmc_chain(D, A, B, orders0[c], n_iter, gamma, swap_prob=0.9, seed=99, chain=3 + c, rule=rule,
                                  cand=cand, use_nem=use_nem, burn_in=burn, keep_scores=True)
        _compare_chain(res, c, ref, S)
        assert np.allclose(res.score_trace[:, 0, c], ref["all_score_list"][1:], rtol=1e-6)
        assert st.init_scores[c].item() == pytest.approx(ref["all_score_list"][0], rel=1e-6)
        total += ref["edge_counts"]
    assert np.array_equal(res.edge_counts, total)
    assert 0 < res.accept_counts.sum() < n_chains * n_iter


@pytest.mark.parametrize("S,E,impl", [(70, 300, "auto"), (100, 260, "lutm"), (100, 260, "ex2_2pass"), (58, 200, "auto")])
def test_mcmc_trajectory_large_S_vs_oracle(nb, S, E, impl):
    """The device-resident loop around the S > 54 score paths (multi-level table epilogue = what `auto` picks, and the
    max-subtracted ex2 epilogue): same accept decisions, orders and edge counts as the oracle chain."""
    _, _, D = orc.synth_dataset(S, E, seed=S)
    A, B = orc.nem_constants(0.05, 0.1)
    rng = np.random.default_rng(S + 1)
    cand = (rng.random((S, S)) < 0.2).astype(np.uint8)
    n_chains, n_iter, burn = 5, 120, 20
    gamma = 2.0 * S / E
    orders0 = np.stack([rng.permutation(S) for _ in range(n_chains)]).astype(np.int32)
    ds = nb.Dataset(D, A, B, sort_egenes=True)
    assert not ds.lut_valid and ds.lutm_levels >= 2
    st = ds.run_chains(orders0, n_iter, gamma, swap_prob=0.9, seed=5, chain_offset=11, rule=1, cand=cand, use_nem=True,
                       burn_in=burn, trace=True, impl=impl)
    res = st.result()
    total = np.zeros((S, S), dtype=np.int64)
    for c in range(n_chains):
        ref = c_oracle.mcmc_chain(D, A, B, orders0[c], n_iter, gamma, swap_prob=0.9, seed=5, chain=11 + c, rule=1,
                                  cand=cand, use_nem=True, burn_in=burn, keep_scores=True)
        _compare_chain(res, c, ref, S)
        assert np.allclose(res.score_trace[:, 0, c], ref["all_score_list"][1:], rtol=1e-6)
        total += ref["edge_counts"]
    assert np.array_equal(res.edge_counts, total)
    assert 0 < res.accept_counts.sum() < n_chains * n_iter


def _oracle_chains(D, A, B, orders0, chains, n_iter, gamma, **kw):
    """c_oracle.mcmc_chain for several chains at once (ctypes releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(8, len(chains))) as ex:
        futs = {c: ex.submit(c_oracle.mcmc_chain, D, A, B, orders0[c], n_iter, gamma, chain=c, **kw) for c in chains}
        return {c: f.result() for c, f in futs.items()}


@pytest.mark.parametrize("name,S,E,n_chains,n_iter,check,sort,impl", [
    # BASELINE config 2 exactly: every chain against the oracle (the persistent fused kernel is what `auto` picks)
    ("config2", 20, 1000, 64, 100, list(range(64)), False, "auto"),
    ("config2-loop", 20, 1000, 64, 40, [0, 1, 31, 32, 63], False, "lut"),
    # BASELINE config 3: 1024 chains so that the two-group CUDA-graph loop runs; chains of both groups are checked
    ("config3", 50, 10000, 1024, 64, [0, 1, 255, 511, 512, 513, 800, 1023], True, "auto"),
    # BASELINE config 4's one-GPU share: S=100, E=100 000, 512 chains, sorted E-genes, multi-level table epilogue
    ("config4-shape", 100, 100000, 512, 32, [0, 255, 256, 511], True, "auto"),
])
def test_mcmc_baseline_shapes_vs_oracle(nb, name, S, E, n_chains, n_iter, check, sort, impl):
    """The device-resident order-MCMC loop at the BASELINE shapes: accept counts, current and best orders,
    best graphs, scores and per-chain posterior edge counts equal the oracle's restatement of
    nem_order_mcmc.py:272-302 driven by the same Philox stream (MASKED rule + closure, i.e. use_nem=True :283-284)."""
    _, _, D = orc.synth_dataset(S, E, seed=0)
    A, B = orc.nem_constants(0.05, 0.1)
    rng = np.random.default_rng(7)
    cand = (rng.random((S, S)) < (0.2 if S >= 100 else 0.3)).astype(np.uint8)
    orders0 = np.stack([np.random.default_rng(1000 + c).permutation(S) for c in range(n_chains)]).astype(np.int32)
    gamma = 2.0 * S / E
    ds = nb.Dataset(D, A, B, sort_egenes=sort)
    if S >= 100:
        assert ds.lutm_levels >= 2
    st = nb.ChainState(ds, orders0, gamma, edge_counts=True, seed=1234, rule=nb.RULE_MASKED, cand=cand, use_nem=True)
    st.run(n_iter, swap_prob=0.95, impl=impl)
    res = st.result()
    per_chain = st.edge_counts.cpu().numpy().astype(np.int64).reshape(n_chains, S, S)
    refs = _oracle_chains(D, A, B, orders0, check, n_iter, gamma, swap_prob=0.95, seed=1234, rule=1, cand=cand, use_nem=True)
    for c in check:
        _compare_chain(res, c, refs[c], S)
        assert np.array_equal(per_chain[c], refs[c]["edge_counts"]), (name, c)
    assert res.edge_counts.sum() == per_chain.sum() and 0 < res.accept_counts.sum() < n_chains * n_iter


def test_mcmc_with_pair_merged_score(nb):
    """The device loop with the pair-merged score kernel (impl="pair") reproduces the oracle chain draw for draw."""
    S, E = 40, 200
    _, _, D = orc.synth_dataset(S, E, seed=7)
    A, B = orc.nem_constants(0.05, 0.1)
    rng = np.random.default_rng(1)
    cand = (rng.random((S, S)) < 0.4).astype(np.uint8)
    orders0 = np.stack([rng.permutation(S) for _ in range(5)]).astype(np.int32)
    ds = nb.Dataset(D, A, B)
    res = ds.run_chains(orders0, 300, 2.0 * S / E, swap_prob=0.9, seed=99, chain_offset=3, rule=1, cand=cand, use_nem=True,
                        burn_in=50, impl="pair").result()
    for c in range(5):
        ref = c_oracle.mcmc_chain(D, A, B, orders0[c], 300, 2.0 * S / E, swap_prob=0.9, seed=99, chain=3 + c, rule=1,
                                  cand=cand, use_nem=True, burn_in=50)
        _compare_chain(res, c, ref, S)


@pytest.mark.parametrize("impl", ["fused", "lut"])
def test_mcmc_resume_equals_single_run(nb, impl):
    S, E = 12, 150
    _, _, D = orc.synth_dataset(S, E, seed=2)
    A, B = orc.nem_constants(0.05, 0.1)
    ds = nb.Dataset(D, A, B)
    orders0 = np.stack([np.random.default_rng(c).permutation(S) for c in range(5)]).astype(np.int32)
    a = ds.run_chains(orders0, 300, 0.2, seed=5, burn_in=100, impl=impl).result()
    st = ds.run_chains(orders0, 120, 0.2, seed=5, burn_in=100, impl=impl)
    st.run(37, impl=impl); st.run(143, impl=impl)
    b = st.result()
    for f in ("best_scores", "best_orders", "curr_orders", "accept_counts", "edge_counts", "curr_scores"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f


def test_mcmc_boltzmann_distribution(nb):
    """S=4: visit frequencies over many device chains follow exp(gamma*score)."""
    S, E = 4, 24
    _, _, D = orc.synth_dataset(S, E, seed=1)
    A, B = orc.nem_constants(0.05, 0.1)
    gamma = 0.35
    perms = list(itertools.permutations(range(S)))
    sc = np.array([orc.score_counts(orc.order_to_graph(np.array(o)), D, A, B) for o in perms])
    p = np.exp(gamma * (sc - sc.max())); p /= p.sum()
    ds = nb.Dataset(D, A, B)
    n = 8192
    res = ds.run_chains(np.tile(np.arange(S, dtype=np.int32), (n, 1)), 60, gamma, swap_prob=0.5, seed=11).result()
    idx = {o: k for k, o in enumerate(perms)}
    freq = np.bincount([idx[tuple(int(x) for x in o)] for o in res.curr_orders], minlength=len(perms)) / n
    assert np.abs(freq - p).max() < 0.02


def test_replica_exchange_properties(nb):
    S, E = 10, 100
    _, _, D = orc.synth_dataset(S, E, seed=4)
    ds = nb.Dataset(D, *orc.nem_constants(0.05, 0.1))
    ladder, n = 10, 40
    gam = np.tile((1 + 0.2 * np.arange(ladder)) * S / E, n // ladder)
    st = ds.run_chains(np.stack([np.random.default_rng(c).permutation(S) for c in range(n)]).astype(np.int32), 50, gam)
    before = st.curr_scores.clone()
    orders_before = st.orders.clone()
    nsw = int(st.replica_exchange(ladder, 0, 0).item())
    after = st.curr_scores
    # scores are permuted only inside pairs (2k, 2k+1); orders travel with their scores
    for k in range(0, n, 2):
        pair_b, pair_a = before[k:k + 2], after[k:k + 2]
        same = torch.equal(pair_a, pair_b)
        assert same or torch.equal(pair_a, pair_b.flip(0))
        if not same:
            assert torch.equal(st.orders[k], orders_before[k + 1]) and torch.equal(st.orders[k + 1], orders_before[k])
    assert nsw == sum(int(not torch.equal(after[k:k + 2], before[k:k + 2])) for k in range(0, n, 2)) or nsw >= 0
    sc = ds.score_bits(st.curr_bits).cpu().numpy()
    # graphs travelled with the scores too (the chains ran in the fused kernel, this is the tensor-core kernel: two
    # fp32 summation orders, the suite's 1e-6 tolerance; different graphs differ in the first digits)
    assert np.allclose(sc, after.cpu().numpy(), rtol=1e-6)


# ------------------------------------------------------------------------------- drop-in classes
@pytest.mark.parametrize("n", [0, 11, 12])
def test_nem_class_golden(nb, golden, n):
    arr, meta = golden
    info = meta["networks"][str(n)]
    model = nb.NEM(arr[f"adj{n}"].astype(int), arr[f"end{n}"], [info["alpha"], info["beta"]], info["S"], info["E"])
    assert np.array_equal(model.observed_knockdown_mat, arr[f"D{n}"])
    assert np.allclose(model.U, arr[f"U{n}"], rtol=1e-13, atol=1e-12)
    assert model.obs_ll == pytest.approx(info["obs_ll"], rel=1e-12)
    assert model.real_ll == pytest.approx(info["real_ll"], abs=1e-9)
    assert model.A == info["A"] and model.B == info["B"]
    # reference call sequence nem.py:132-142 through the mirrored methods
    S = info["S"]
    adj = arr[f"adj{n}"].astype(int); np.fill_diagonal(adj, 0)
    parents = [[i for i in range(S) if adj[i][j] == 1] for j in range(S)]
    w = [[1.0] * len(p) for p in parents]
    tabs = model.get_score_tables(model.observed_knockdown_mat)
    red = model.get_reduced_score_tables(tabs, parents)
    ow, ll = model.calculate_ll(model.compute_ll_ratios(w, red, [len(p) for p in parents], model.U))
    assert ll == pytest.approx(info["obs_ll"], rel=1e-12)
    assert ow.shape == (S + 1, info["E"]) and np.allclose(ow.sum(axis=0), 1.0)
    assert model.score_graphs(adj[None])[0] == pytest.approx(info["obs_ll"], rel=REL_TC)


def test_mcmc_class_reference_smoke(nb):
    """The reference's own smoke test (tests/utils.tests.py:30-49) through the drop-in classes."""
    s_mat = np.array([[0, 1, 1, 0, 1], [0, 0, 1, 0, 1], [0, 0, 0, 0, 1], [0, 0, 1, 0, 1], [0, 0, 0, 0, 0]])
    e_arr = [0, 1, 2, 3, 0, 1, 4, 2]
    model = nb.NEM(adj_matrix=s_mat, end_nodes=e_arr, errors=[0.05, 0.1], num_s=5, num_e=8)
    mc = nb.NEMOrderMCMC(model, perm_order=np.array([4, 3, 2, 1, 0]))
    score, best_dag = mc.method(n_iterations=110, gamma=2.0 * 5 / 8, seed=42, verbose=False)
    assert best_dag.shape == (5, 5) and np.isfinite(score)
    assert len(mc.all_score_list) == 111 and len(mc.curr_score_list) == 111
    assert score == max(mc.best_score_list) and score >= mc.all_score_list[0]
    assert score == pytest.approx(c_oracle.score(best_dag.astype(np.uint8), model.observed_knockdown_mat.astype(np.uint8),
                                                 model.A, model.B), rel=REL_TC)
    ref = c_oracle.mcmc_chain(model.observed_knockdown_mat, model.A, model.B, [4, 3, 2, 1, 0], 110, 1.25, seed=42)
    assert score == pytest.approx(ref["best_score"], rel=REL_TC)
    assert np.array_equal(mc.best_order, ref["best_order"])
    # multi-chain + masked rule + replica exchange run and return consistent results
    s2, d2 = mc.method(n_iterations=64, gamma=1.25, seed=1, verbose=False, n_chains=32, rule="masked",
                       cand=np.triu(np.ones((5, 5)), 1) + np.tril(np.ones((5, 5)), -1), use_nem=True)
    assert s2 == pytest.approx(model.score_graphs(d2[None])[0], rel=REL_TC)
    assert mc.edge_freq.shape == (5, 5) and mc.edge_freq.max() <= 1.0 + 1e-12
    s3, rep = nb.replica_exchange_method(model, 6, 20, np.array([4, 3, 2, 1, 0]))
    assert np.isfinite(s3) and rep.best_dag.shape == (5, 5)


def test_mcmc_cell_ratio_methods(nb, golden):
    """compute_cell_ratios / calculate_ll of the driver (nem_order_mcmc.py:79-93) incl. the expit quirk."""
    from scipy.special import expit
    arr, meta = golden
    info = meta["networks"]["12"]
    model = nb.NEM(arr["adj12"].astype(int), arr["end12"], [info["alpha"], info["beta"]], info["S"], info["E"])
    order = arr["order0_12"]
    mc = nb.NEMOrderMCMC(model, order)
    S = info["S"]
    rng = np.random.default_rng(0)
    w = rng.normal(size=(S, S))
    T, N, U = orc.score_tables(arr["D12"], info["A"], info["B"])
    for use_expit in (True, False):
        ww = expit(w) if use_expit else np.clip(w, 0.01, 0.99)
        ref = U.copy()
        for i in range(S):
            for j in mc.parents_list[i]:
                ref[i] += np.log(1.0 - ww[i][j] + ww[i][j] * np.exp(T[j]))
        got = mc.compute_cell_ratios(w if use_expit else ww, expit=use_expit)
        assert np.allclose(got, ref, rtol=1e-12, atol=1e-12)
        mc.cell_ratios = got
        ow, ll = mc.calculate_ll()
        assert ll == pytest.approx(float(np.logaddexp.reduce(ref, axis=0).sum()), rel=1e-12)
    assert nb.utils.compute_ll(ref) == pytest.approx(ll, rel=1e-12)


@pytest.mark.parametrize("S", [121, 128])
def test_cell_ratios_and_nem_class_at_max_S(nb, S):
    """compute_ll_ratios / compute_cell_ratios for S = 121..128 (the kernel stages its tables per chunk of
    child rows, the whole S x S tables would not fit a CTA): drop-in NEM() construction and general weights."""
    rng = np.random.default_rng(S)
    E = 150
    adj = c_oracle.closure(nb.synth.random_dag(S, rng))
    ends = rng.integers(0, S, E)
    model = nb.NEM(adj.astype(int), ends, [0.05, 0.1], S, E)
    A, B = orc.nem_constants(0.05, 0.1)
    T, N, U = orc.score_tables(model.observed_knockdown_mat, A, B)
    assert np.allclose(model.U, U, rtol=1e-12, atol=1e-12)
    assert model.obs_ll == pytest.approx(orc.score_matrix(adj, model.observed_knockdown_mat, A, B), rel=1e-11)
    order = rng.permutation(S)
    mc = nb.NEMOrderMCMC(model, order)
    w = rng.normal(size=(S, S))
    ref = U.copy()
    from scipy.special import expit
    ww = expit(w)
    for i in range(S):
        for j in mc.parents_list[i]:
            ref[i] += np.log(1.0 - ww[i][j] + ww[i][j] * np.exp(T[j]))
    got = mc.compute_cell_ratios(w, expit=True)
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-11)


def test_network_csv_round_trip(nb, tmp_path):
    """utils.write_network_csv writes what DAGs/rnd_dag_gen.py:86-103 writes (closed + reduced files) and
    utils.read_csv_to_adj (:96-118) reads it back; the reference's own shipped pair reproduces."""
    rng = np.random.default_rng(4)
    S, E = 17, 40
    raw = nb.synth.random_dag(S, rng)
    ends = rng.integers(0, S, E)
    f, fr = tmp_path / "network99.csv", tmp_path / "network99_red.csv"
    closed = nb.utils.write_network_csv(raw, ends, [0.05, 0.1], str(f), str(fr))
    assert np.array_equal(closed, c_oracle.closure(raw))
    adj, e2, err, s2, ne2 = nb.utils.read_csv_to_adj(str(f))
    assert (s2, ne2) == (S, E) and np.array_equal(adj, closed) and np.array_equal(e2, ends) and list(err) == [0.05, 0.1]
    red, _, _, _, _ = nb.utils.read_csv_to_adj(str(fr))
    assert np.array_equal(red, orc.transitive_reduction(closed)) and np.array_equal(c_oracle.closure(red), closed)
    assert f.read_bytes().count(b"\r\n") == closed.sum() + 3


@pytest.mark.parametrize("S", [5, 20, 64, 128])
def test_ancestor_x_inverse_vs_numpy(nb, S):
    """nemb200_ancestor_x = clip(inv(I - expit_parent_weights(W)) - I, 0, 1) (nem_order_mcmc.py:185, :152-157): the
    in-kernel Gauss-Jordan against numpy's LU inverse, for weights that respect an order (every pivot 1) and for a
    general matrix whose inversion needs row exchanges."""
    import ctypes as C
    from scipy.special import expit
    L = nb._lib.lib()
    rng = np.random.default_rng(S)
    order = rng.permutation(S)
    pos = np.argsort(order)
    perm = (pos[None, :] < pos[:, None]).astype(np.uint8)        # k permissible parent of i iff k comes earlier
    for case in ("order", "general"):
        W = rng.normal(size=(S, S))
        if case == "order":
            W = W * perm
            pm = perm
        else:
            pm = (rng.random((S, S)) < 0.5).astype(np.uint8)
            W[np.arange(S), np.arange(S)] = 1.0 + rng.random(S)  # I - W: small / negative diagonal -> pivoting
            W *= 0.9
        M = np.where(pm != 0, expit(W), W)
        ref = np.clip(np.linalg.inv(np.eye(S) - M) - np.eye(S), 0.0, 1.0)
        wt = torch.from_numpy(W).cuda()
        pt = torch.from_numpy(pm).cuda()
        out = torch.empty((S, S), dtype=torch.float64, device="cuda")
        sing = torch.zeros(1, dtype=torch.int32, device="cuda")
        nb._lib.check(L.nemb200_ancestor_x(C.c_void_p(wt.data_ptr()), C.c_void_p(pt.data_ptr()), S, C.c_void_p(out.data_ptr()),
                                           C.c_void_p(sing.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        assert int(sing.item()) == 0
        cond = np.linalg.cond(np.eye(S) - M)
        assert np.allclose(out.cpu().numpy(), ref, rtol=0, atol=1e-12 * max(1.0, cond)), (case, cond)


def test_device_synthesis_vs_oracle(nb):
    """nemb200_synth_knockdown is bit-exact against the oracle restatement of utils.py:15-35 driven
    by the same Philox stream, and has the flip rates alpha / beta at a large shape."""
    rng = np.random.default_rng(8)
    S, E = 13, 300
    adj = c_oracle.closure(nb.synth.random_dag(S, rng))
    ends = rng.integers(0, S, E)
    obs, real = nb.synth.device_knockdown(adj, ends, 0.05, 0.1, seed=77, want_real=True)
    r_ref, o_ref = orc.philox_observed(adj, ends, 0.05, 0.1, 77)
    assert np.array_equal(real.cpu().numpy(), r_ref) and np.array_equal(obs.cpu().numpy(), o_ref)
    S, E = 100, 100000
    adj = c_oracle.closure(nb.synth.random_dag(S, rng))
    ends = rng.integers(0, S, E)
    obs, real = nb.synth.device_knockdown(adj, ends, 0.05, 0.1, seed=5, want_real=True)
    assert np.array_equal(real.cpu().numpy(), orc.create_real_knockdown_mat(adj, ends).astype(np.uint8))
    r, o = real.bool(), obs.bool()
    fp = (o & ~r).sum().item() / max(1, (~r).sum().item())
    fn = (~o & r).sum().item() / max(1, r.sum().item())
    assert abs(fp - 0.05) < 0.002 and abs(fn - 0.1) < 0.004
    ds = nb.Dataset(obs, *orc.nem_constants(0.05, 0.1))   # device tensor accepted as data
    assert np.isfinite(ds.score_graphs(adj[None].astype(np.uint8))[0])


def test_c_abi_error_paths(nb):
    """Status codes + nemb200_last_error for bad arguments (nothing is launched)."""
    import ctypes as C
    L = nb._lib.lib()
    S, E = 20, 100
    D = (np.random.default_rng(0).random((S, E)) < 0.5).astype(np.uint8)
    ds = nb.Dataset(D, *orc.nem_constants(0.05, 0.1))
    bits = ds.order_graphs(np.arange(S)[None])
    out = torch.empty(1, dtype=torch.float64, device="cuda")
    ws = torch.empty(16, dtype=torch.uint8, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = L.nemb200_score_batch(ds._h, C.c_void_p(bits.data_ptr()), 1, C.c_void_p(out.data_ptr()), C.c_void_p(ws.data_ptr()), 16, 0, st)
    assert rc < 0 and b"workspace" in L.nemb200_last_error()
    rc = L.nemb200_score_batch(ds._h, None, 1, C.c_void_p(out.data_ptr()), C.c_void_p(ws.data_ptr()), 16, 0, st)
    assert rc < 0 and b"null" in L.nemb200_last_error()
    rc = L.nemb200_closure_u32(C.c_void_p(bits.data_ptr()), 1, 129, C.c_void_p(bits.data_ptr()), st)
    assert rc < 0 and b"S must be" in L.nemb200_last_error()
    rc = L.nemb200_order_to_graph(C.c_void_p(bits.data_ptr()), 1, S, 1, None, 1, C.c_void_p(bits.data_ptr()), st)
    assert rc < 0 and b"candidate" in L.nemb200_last_error()
    rc = L.nemb200_transitive_reduction_u32(C.c_void_p(bits.data_ptr()), 1, 0, C.c_void_p(bits.data_ptr()), st)
    assert rc < 0 and b"S must be" in L.nemb200_last_error()
    rc = L.nemb200_transitive_reduction_u32(None, 1, S, C.c_void_p(bits.data_ptr()), st)
    assert rc < 0 and b"bad arguments" in L.nemb200_last_error()
    assert L.nemb200_transitive_reduction_u32(None, 0, S, None, st) == 0                     # empty batch: nothing to do
    rc = L.nemb200_score_batch(ds._h, C.c_void_p(bits.data_ptr()), 1, C.c_void_p(out.data_ptr()), C.c_void_p(ds._score_workspace(1)[0].data_ptr()),
                               ds._score_workspace(1)[1], nb.engine.SCORE_IMPLS["ex2_2pass"], st)
    assert rc == 0
    big = nb.Dataset(np.zeros((126, 10), dtype=np.uint8), *orc.nem_constants(0.05, 0.1))       # S > 125: no multi-level table
    assert big.lutm_levels == 0
    with pytest.raises(nb._lib.NemB200Error, match="multi-level"):
        big.score_graphs(np.zeros((1, 126, 126), dtype=np.uint8), impl="lutm")
    with pytest.raises(nb._lib.NemB200Error):
        ds.score_graphs(np.zeros((1, S, S), dtype=np.uint8), impl=17)
    with pytest.raises(ValueError):
        ds.score_graphs(np.zeros((1, S + 1, S + 1), dtype=np.uint8))
    with pytest.raises(ValueError):
        ds.run_chains(np.zeros((2, S), dtype=np.int32), 1, 0.1)   # not permutations


def test_score_batch_split_over_launch_limit(nb):
    """Batches beyond one launch's 32-bit tile index are split by Dataset.score_bits: same scores as one launch."""
    S, E, n = 20, 300, 3000
    rng = np.random.default_rng(5)
    _, _, D = orc.synth_dataset(S, E, seed=5)
    ds = nb.Dataset(D, *orc.nem_constants(0.05, 0.1))
    bits = ds.order_graphs(np.stack([rng.permutation(S) for _ in range(n)]).astype(np.int32))
    whole = ds.score_bits(bits).cpu().numpy()
    assert ds.max_graphs_per_launch() >= 1 << 20          # E = 300: the limit is far away
    ds.max_graphs_per_launch = lambda: 1024               # force three launches (1024 + 1024 + 952)
    assert np.array_equal(ds.score_bits(bits).cpu().numpy(), whole)
    assert np.allclose(ds.score_bits(bits, impl="witness").cpu().numpy(), whole, rtol=REL_TC)
    big = nb.Dataset(np.zeros((100, 2_000_000), dtype=np.uint8)[:, :1], *orc.nem_constants(0.05, 0.1))
    big.n_mtiles = 15626                                   # as for E = 2 000 000: about 2400 graphs per launch
    assert 1024 <= big.max_graphs_per_launch() <= 4096 and big.max_graphs_per_launch() % 1024 == 0


def test_sorted_egenes_and_multichain_loop_path(nb):
    """Dataset(sort_egenes=True) is score-invariant; the CUDA-graph loop around the tensor-core score and the
    persistent kernel (which takes more chains than fit on the GPU at once in waves) agree chain by chain (same
    global chain ids)."""
    S, E = 16, 500
    _, _, D = orc.synth_dataset(S, E, seed=9)
    A, B = orc.nem_constants(0.05, 0.1)
    rng = np.random.default_rng(2)
    graphs = rand_dags(rng, 33, S)
    a = nb.Dataset(D, A, B).score_graphs(graphs)
    b = nb.Dataset(D, A, B, sort_egenes=True).score_graphs(graphs)
    assert np.allclose(a, b, rtol=1e-12)
    n = 320
    orders0 = np.stack([rng.permutation(S) for _ in range(n)]).astype(np.int32)
    ds = nb.Dataset(D, A, B)
    loop = ds.run_chains(orders0, 64, 0.1, seed=3, impl="lut").result()  # the three-kernel loop
    fused = ds.run_chains(orders0[:64], 64, 0.1, seed=3).result()         # auto -> persistent kernel
    assert np.array_equal(loop.curr_orders[:64], fused.curr_orders)
    assert np.array_equal(loop.accept_counts[:64], fused.accept_counts)
    assert np.allclose(loop.best_scores[:64], fused.best_scores, rtol=1e-12)
    waves = ds.run_chains(orders0, 64, 0.1, seed=3).result()              # auto: 320 chains > 296 resident CTAs, still fused
    assert np.array_equal(waves.curr_orders, loop.curr_orders)
    assert np.array_equal(waves.accept_counts, loop.accept_counts)
    assert np.array_equal(waves.edge_counts, loop.edge_counts)


def test_mcmc_two_group_overlap_equals_separate_runs(nb):
    """>= 512 chains run as two chain groups captured as parallel branches of one CUDA graph (csrc/mcmc.cu);
    chains keep their global Philox id, so the result equals two independent runs of the halves (and the
    oracle for a few chains), including the score trace, edge counts and a resumed continuation."""
    S, E = 20, 700
    _, _, D = orc.synth_dataset(S, E, seed=11)
    A, B = orc.nem_constants(0.05, 0.1)
    rng = np.random.default_rng(5)
    cand = (rng.random((S, S)) < 0.4).astype(np.uint8)
    n, n_iter, burn, gamma = 602, 77, 10, 2.0 * S / E
    orders0 = np.stack([rng.permutation(S) for _ in range(n)]).astype(np.int32)
    ds = nb.Dataset(D, A, B)
    kw = dict(swap_prob=0.9, seed=21, rule=nb.RULE_MASKED, cand=cand, use_nem=True, burn_in=burn, trace=True)
    st = ds.run_chains(orders0, n_iter, gamma, chain_offset=5, **kw)
    both = st.result()
    h = 300
    lo = ds.run_chains(orders0[:h], n_iter, gamma, chain_offset=5, **kw).result()
    hi = ds.run_chains(orders0[h:], n_iter, gamma, chain_offset=5 + h, **kw).result()
    for f in ("best_scores", "best_orders", "curr_orders", "accept_counts", "curr_scores", "best_graphs"):
        assert np.array_equal(getattr(both, f), np.concatenate([getattr(lo, f), getattr(hi, f)])), f
    assert np.array_equal(both.score_trace, np.concatenate([lo.score_trace, hi.score_trace], axis=2))
    assert np.array_equal(both.edge_counts, lo.edge_counts + hi.edge_counts)
    for c in (0, 299, 300, 301, 601):
        ref = c_oracle.mcmc_chain(D, A, B, orders0[c], n_iter, gamma, swap_prob=0.9, seed=21, chain=5 + c, rule=1,
                                  cand=cand, use_nem=True, burn_in=burn, keep_scores=True)
        _compare_chain(both, c, ref, S)
    # resumed continuation of the two-group run == one longer run
    st.run(40, swap_prob=0.9)
    long = ds.run_chains(orders0, n_iter + 40, gamma, chain_offset=5, **kw).result()
    cont = st.result()
    for f in ("best_scores", "curr_orders", "accept_counts", "edge_counts"):
        assert np.array_equal(getattr(cont, f), getattr(long, f)), f


@pytest.mark.parametrize("faithful", [True, False])
@pytest.mark.parametrize("S,E,rule,use_nem", [(8, 60, 0, False), (13, 126, 1, True)])
def test_replica_exchange_vs_oracle(nb, S, E, rule, use_nem, faithful):
    """Replica exchange swap for swap against the oracle's statement-for-statement restatement of
    nem_order_mcmc.py:316-363 (faithful=True: scores[i] = best_score of the round, whole replicas swapped, overall
    best over rounds; faithful=False: the documented default, delta from the current scores), same Philox stream."""
    _, _, D = orc.synth_dataset(S, E, seed=2)
    A, B = orc.nem_constants(0.05, 0.1)
    rng = np.random.default_rng(3)
    cand = (rng.random((S, S)) < 0.4).astype(np.uint8)
    n_rep, n_exchange, n_iter, seed = 10, 8, 25, 77
    orders0 = np.stack([rng.permutation(S) for _ in range(n_rep)]).astype(np.int32)
    ref = orc.replica_exchange_method(D, A, B, orders0, n_exchange, n_iter, n_replicas=n_rep, seed=seed, rule=rule, cand=cand,
                                      use_nem=use_nem, swap_prob=0.95, use_best=faithful)
    ds = nb.Dataset(D, A, B)
    gammas = (1.0 + 0.2 * np.arange(n_rep)) * S / E
    st = nb.ChainState(ds, orders0, gammas, edge_counts=False, seed=seed, rule=rule, cand=cand, use_nem=use_nem)
    total, overall = 0, -np.inf
    for r in range(n_exchange):
        if faithful:
            st.restart_best()
        st.run(n_iter, swap_prob=0.95)
        if faithful:
            overall = max(overall, float(st.best_scores.max().item()))
        nsw = int(st.replica_exchange(n_rep, r & 1, r, faithful=faithful).item())
        ex, ref_orders, ref_scores = ref["trace"][r]
        assert nsw == ex, (r, nsw, ex)
        assert np.array_equal(st.orders.cpu().numpy(), np.stack(ref_orders)), r
        got_scores = (st.best_scores if faithful else st.curr_scores).cpu().numpy()
        assert np.allclose(got_scores, ref_scores, rtol=1e-6), r
        total += nsw
    assert total == ref["n_exchanges"] and total > 0
    if faithful:
        assert overall == pytest.approx(ref["best_score"], rel=1e-6)
    # the drop-in wrapper runs the same rounds
    class _N:
        pass
    nm = _N()
    nm.num_s, nm.num_e, nm.dataset, nm.device = S, E, ds, ds.device
    nm.U = np.zeros((S + 1, E))
    import nem_b200.nem_order_mcmc as nom
    saved = nom.start_orders
    nom.start_orders = lambda perm, n, sd: orders0.copy()
    try:
        score, best = nom.replica_exchange_method(nm, n_exchange, n_iter, orders0[0], n_replicas=n_rep, seed=seed,
                                                  rule={0: "full", 1: "masked"}[rule], cand=cand, use_nem=use_nem, faithful=faithful)
    finally:
        nom.start_orders = saved
    assert best.n_exchanges == ref["n_exchanges"]
    if faithful:
        assert score == pytest.approx(ref["best_score"], rel=1e-6) and np.array_equal(best.best_dag, ref["best_graph"])


def test_replica_exchange_swap_rate(nb):
    """Swap decisions follow u < exp(-delta) (nem_order_mcmc.py:333-337): with equal gammas delta = 0, so every
    pair swaps; with a wide ladder and spread-out scores fewer do."""
    S, E = 10, 100
    _, _, D = orc.synth_dataset(S, E, seed=4)
    ds = nb.Dataset(D, *orc.nem_constants(0.05, 0.1))
    n, ladder = 64, 8
    orders = np.stack([np.random.default_rng(c).permutation(S) for c in range(n)]).astype(np.int32)
    st = ds.run_chains(orders, 8, 0.3)
    assert int(st.replica_exchange(ladder, 0, 0).item()) == n // 2        # (0,1),(2,3).. all swap
    assert int(st.replica_exchange(ladder, 1, 1).item()) == (n // ladder) * (ladder // 2 - 1)  # (1,2),(3,4),(5,6) per ladder
    gam = np.tile(np.linspace(0.05, 5.0, ladder), n // ladder)
    st2 = ds.run_chains(orders, 8, gam)
    s = st2.curr_scores.cpu().numpy(); g = gam
    expect = sum(min(1.0, np.exp(-((g[i] - g[i + 1]) * (s[i + 1] - s[i])))) for i in range(0, n, 2))
    got = int(st2.replica_exchange(ladder, 0, 0).item())
    assert abs(got - expect) <= max(6.0, 4 * np.sqrt(expect))


# ---------------------------------------------------------------------------- per-order weight fit (f-1)
def test_fit_edges_vs_exact_oracle(nb):
    """nemb200_fit_edges (one CTA per (node, parent) pair) against oracle.fit_pass(minimiser="exact") on the
    reference's golden fit inputs: same minimisers, same thresholded DAG, same soft log-likelihood."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "fit.npz"))

    class _Nem:  # the attributes NEMOrderMCMC reads
        pass
    for tag in [str(t) for t in z["cases"]]:
        D, (A, B), order = z[f"D_{tag}"], z[f"AB_{tag}"], z[f"order_{tag}"]
        S, E = D.shape
        nm = _Nem()
        nm.num_s, nm.num_e = S, E
        nm.dataset = nb.Dataset(D, float(A), float(B))
        nm.device = nm.dataset.device
        _, _, U = orc.score_tables(D, A, B)
        nm.U = U
        mc = nb.NEMOrderMCMC(nm, order)
        for w_in in (z[f"w0_{tag}"], z[f"w1_{tag}"]):
            mc.get_permissible_parents(order, init=True)
            mc.parent_weights = np.array(w_in, copy=True)
            got, ll = mc.fit_weights()
            ref, ll_ref = orc.fit_pass(D, A, B, order, w_in, "exact")
            assert ll == pytest.approx(ll_ref, rel=1e-11)
            perm = w_in != 0
            close = np.abs(got - ref)[perm] < 1e-6
            assert close.mean() >= 0.98, (tag, close.mean())       # ties between equal minima may pick the other one
            decided = np.abs(ref - 0.5) > 1e-6
            assert np.array_equal((got > 0.5)[perm & decided & (np.abs(got - ref) < 1e-6)],
                                  (ref > 0.5)[perm & decided & (np.abs(got - ref) < 1e-6)])
            assert np.array_equal(got[~perm], w_in[~perm])          # masked-out pairs are copied through
        # the drop-in method: one pass + threshold + score, as the reference's driver calls it (:277)
        mc.get_permissible_parents(order, init=True)
        dag_ll = mc.get_optimal_weights(init=True, use_nem=False)
        ref_w, _ = orc.fit_pass(D, A, B, order, z[f"w0_{tag}"], "exact")
        dagw = 1 * (ref_w > 0.5)
        T, _, U = orc.score_tables(D, A, B)
        cell = U.copy()
        ex = 1.0 / (1.0 + np.exp(-dagw.astype(float)))
        pos = np.empty(S, dtype=int); pos[order] = np.arange(S)
        for i in range(S):
            for k in order[:pos[i]]:
                cell[i] += np.log(1.0 - ex[i][k] + ex[i][k] * np.exp(T[k]))     # :79-87 with the expit quirk
        assert dag_ll == pytest.approx(float(np.logaddexp.reduce(cell, axis=0).sum()), rel=1e-10)


def test_mcmc_method_with_fitted_weights(nb):
    """The reference's own smoke test (tests/utils.tests.py:30-49: 5-node DAG, E=8, errors (0.05, 0.1), order
    [4,3,2,1,0], 110 iterations, gamma = 2S/E, no assertion upstream) through method(rule="fit"): the reference-
    faithful loop with the GPU edge-weight fit.  Checks the bookkeeping contract of :303-309."""
    adj = np.array([[0, 1, 1, 1, 1], [0, 0, 1, 1, 1], [0, 0, 0, 1, 1], [0, 0, 0, 0, 1], [0, 0, 0, 0, 0]])
    end_nodes = np.array([0, 1, 2, 3, 4, 0, 1, 2])
    model = nb.NEM(adj, end_nodes, (0.05, 0.1), 5, 8, seed=42)
    mc = nb.NEMOrderMCMC(model, np.array([4, 3, 2, 1, 0]))
    for rule in ("fit_host", "fit"):         # the host-driven loop and the device-resident loop (NEMB200_RULE_FIT)
        mc = nb.NEMOrderMCMC(model, np.array([4, 3, 2, 1, 0]))
        best_score, best_dag = mc.method(n_iterations=110, gamma=2 * 5 / 8, seed=42, verbose=False, rule=rule)
        assert best_dag.shape == (5, 5) and set(np.unique(best_dag)) <= {0, 1}
        assert len(mc.all_score_list) == 111 and len(mc.curr_score_list) == 111
        assert best_score == max(mc.best_score_list) and best_score >= mc.all_score_list[0]
        assert sorted(mc.best_order.tolist()) == [0, 1, 2, 3, 4]
        pos = np.empty(5, dtype=int); pos[mc.best_order] = np.arange(5)
        a, b = np.nonzero(best_dag)
        assert np.all(pos[a] < pos[b])       # only order-compatible edges (SURVEY Q7 fixed)


@pytest.mark.parametrize("S,E,use_nem", [(6, 40, False), (9, 80, True), (13, 126, False)])
def test_mcmc_fit_rule_device_loop_vs_oracle(nb, S, E, use_nem):
    """rule FIT inside the device-resident loop (nemb200_mcmc_run, NEMB200_RULE_FIT): per iteration and chain the
    reference's reset -> edge-weight fits -> threshold (-> closure) -> soft score (nem_order_mcmc.py:276-286, :172-208)
    with no host round trip.  Accept decisions, orders, thresholded graphs, scores and edge counts equal the oracle
    chain that restates the same steps with the exact minimiser and the same Philox stream."""
    _, _, D = orc.synth_dataset(S, E, seed=S + 3)
    A, B = orc.nem_constants(0.05, 0.1)
    rng = np.random.default_rng(S)
    n_chains, n_iter = 5, 24
    gamma = 2.0 * S / E
    orders0 = np.stack([rng.permutation(S) for _ in range(n_chains)]).astype(np.int32)
    ds = nb.Dataset(D, A, B)
    st = nb.ChainState(ds, orders0, gamma, edge_counts=True, seed=31, chain_offset=2, rule=nb.RULE_FIT, use_nem=use_nem)
    st.run(n_iter, swap_prob=0.9, trace=True)
    res = st.result()
    per_chain = st.edge_counts.cpu().numpy().astype(np.int64).reshape(n_chains, S, S)
    for c in range(n_chains):
        ref = orc.mcmc_chain_fit(D, A, B, orders0[c], n_iter, gamma, swap_prob=0.9, seed=31, chain=2 + c, use_nem=use_nem)
        assert np.allclose(res.score_trace[:, 0, c], ref["all_score_list"][1:], rtol=1e-9), c
        _compare_chain(res, c, ref, S, tol=1e-9)
        assert np.array_equal(per_chain[c], ref["edge_counts"])
    assert 0 < res.accept_counts.sum() < n_chains * n_iter
    # resume == single run
    st2 = nb.ChainState(ds, orders0, gamma, edge_counts=True, seed=31, chain_offset=2, rule=nb.RULE_FIT, use_nem=use_nem)
    st2.run(10, swap_prob=0.9); st2.run(n_iter - 10, swap_prob=0.9)
    assert torch.equal(st2.orders, st.orders) and torch.equal(st2.accept, st.accept)


@pytest.mark.parametrize("S,E,n,n_batches,group", [(20, 900, 96, 5, 4), (50, 2000, 300, 11, 4), (50, 2000, 300, 7, 1), (47, 1500, 128, 9, 3)])
def test_score_stream_pipelined_host_batches(nb, S, E, n, n_batches, group):
    """Dataset.score_stream (groups of host->device copies double-buffered against closure + score, the calls of a group
    chained by dependent launches) returns, batch for batch, exactly what the synchronous API returns."""
    _, _, D = orc.synth_dataset(S, E, seed=3)
    A, B = orc.nem_constants(0.05, 0.1)
    ds = nb.Dataset(D, A, B)
    rng = np.random.default_rng(0)
    batches = [rand_dags(rng, n, S, p=0.2, closed=False) for _ in range(n_batches)]
    batches[1][::5] = (rng.random((len(batches[1][::5]), S, S)) < 0.2).astype(np.uint8)   # digraphs (list-mode graphs at S >= 44)
    host = [torch.from_numpy(nb.pack_adj(b).view(np.int32)).pin_memory() for b in batches]
    out = torch.zeros((n_batches, n), dtype=torch.float64).pin_memory()
    ds.score_stream(host, out, group=group)
    torch.cuda.synchronize()
    for k in range(n_batches):
        assert np.array_equal(out[k].numpy(), ds.score_graphs(batches[k], close=True)), k
    ds.score_stream([], out)  # empty sequence is a no-op


# ------------------------------------------------------------------------------------------ f-4: post-processing
def test_transitive_reduction_golden_and_oracle(nb):
    """nemb200_transitive_reduction_u32 vs the reference's own outputs (tests/golden/reduction.npz: closed DAGs,
    unclosed DAGs, cyclic digraphs, digraphs with self loops, S = 1..70) and vs the oracle on larger batches."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "reduction.npz"))
    for c in range(int(z["n_cases"])):
        a = z[f"in{c}"]
        got = nb.transitive_reduction(a[None])[0]
        assert np.array_equal(got, z[f"red{c}"]), (c, a.shape)
        assert np.array_equal(nb.utils.transitive_reduction(a.astype(int)), z[f"red{c}"].astype(int))
    rng = np.random.default_rng(8)
    for S in (7, 32, 50, 97, 128):
        graphs = np.concatenate([rand_dags(rng, 20, S, p=0.2), rand_dags(rng, 10, S, p=0.1, closed=False),
                                 (rng.random((10, S, S)) < 0.05).astype(np.uint8), np.ones((1, S, S), dtype=np.uint8)])
        got = nb.transitive_reduction(graphs)
        for g in range(len(graphs)):
            assert np.array_equal(got[g], orc.transitive_reduction_rows(graphs[g])), (S, g)
        # closed DAG: the reduction is the Hasse diagram - closing it again gives the graph back, and it is minimal
        closed = graphs[:20]
        red = got[:20]
        assert np.array_equal(nb.closure(red), closed)
        assert (red <= closed).all()
    # in place, on device bit rows
    bits = torch.from_numpy(nb.pack_adj(graphs).view(np.int32)).cuda()
    out = nb.transitive_reduction(bits, S=128, return_bits=True)
    assert np.array_equal(nb.unpack_adj(out.cpu().numpy().view(np.uint32), 128), got)


def test_output_handling_writes_dot_files(nb, golden, tmp_path):
    """main.output_handling (main.py:44-55) minus graphviz: closed + reduced DOT files of the best graph."""
    arr, meta = golden
    adj, red = arr["adj11"], arr["red11"]   # the network's closed form and its reduced form (networks/network11*.csv)
    res = nb.utils.output_handling(red, str(tmp_path / "output"), true_adj=adj)
    assert np.array_equal(res["closed"], arr["anc_red11"])       # utils.ancestor(reduced), produced by the reference
    assert np.array_equal(res["reduced"], orc.transitive_reduction_rows((np.asarray(red) != 0).astype(int)))
    for pth, mat in zip(res["paths"], (res["closed"], res["reduced"])):
        text = open(pth).read()
        assert text.startswith("digraph {") and text.endswith("}")
        assert text.count("->") == int(np.count_nonzero(mat))
    assert res["hamming_closed"] == 0 and res["hamming_reduced"] == 0
