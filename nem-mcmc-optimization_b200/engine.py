"""Torch-side plumbing of the B200 NEM hot path: device memory, streams and the batched entry points
(`Dataset.score_graphs`, `closure`, `Dataset.run_chains`) that the drop-in `NEM` / `NEMOrderMCMC`
classes call.  PyTorch only owns tensors and streams; all compute is in libnemb200.so.

Bit-packed graph layout (include/nemb200.h): uint32 [n][S][W], bit b of word w of row a set iff
edge a -> 32*w+b, i.e. the reference's ``dag`` orientation (nem_order_mcmc.py:213).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib

RULE_FULL, RULE_MASKED, RULE_FIT = 0, 1, 2  # FIT: nemb200_mcmc_run only (the reference's per-order edge-weight fit)
SCORE_INPUTS_READY = 0x100  # nemb200.h: flag OR-ed into impl (the graphs are not written by the stream's previous kernel)
SCORE_IMPLS = {"auto": 0, "witness": 1, "lut": 2, "ex2": 3, "ex2_2pass": 4, "fused": 5, "pair": 6, "lutm": 7, "lut_mix": 8}


def _require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.NemB200Error("nem_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def words(S: int) -> int:
    return (S + 31) // 32


def pack_adj(adj) -> np.ndarray:
    """(n,S,S) or (S,S) 0/1 array -> uint32 bit rows (n,S,W) (host, numpy)."""
    a = np.asarray(adj)
    single = a.ndim == 2
    if single:
        a = a[None]
    n, S, S2 = a.shape
    assert S == S2
    W = words(S)
    buf = np.zeros((n, S, W * 32), dtype=np.uint8)
    buf[:, :, :S] = a != 0
    bits = np.packbits(buf, axis=-1, bitorder="little").view("<u4").reshape(n, S, W)
    return bits[0] if single else bits


def unpack_adj(bits, S: int) -> np.ndarray:
    """uint32 bit rows (n,S,W) -> (n,S,S) uint8."""
    b = np.ascontiguousarray(np.asarray(bits).astype("<u4", copy=False))
    single = b.ndim == 2
    if single:
        b = b[None]
    n = b.shape[0]
    u8 = b.view(np.uint8).reshape(n, S, 4 * words(S))
    out = np.unpackbits(u8, axis=-1, bitorder="little")[:, :, :S]
    return out[0] if single else out


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream_ptr(stream: torch.cuda.Stream):
    return C.c_void_p(stream.cuda_stream)


def _bits_to_device(graphs, S, device) -> torch.Tensor:
    """Accept (n,S,S) 0/1 arrays, (n,S,W) uint32 bit rows (numpy) or int32 CUDA tensors of bit rows."""
    if isinstance(graphs, torch.Tensor):
        t = graphs
        if t.dtype not in (torch.int32, torch.uint32):
            raise TypeError("device graphs must be bit rows stored as int32")
        return t.to(device).contiguous().view(torch.int32)
    a = np.asarray(graphs)
    if a.dtype == np.uint32:  # uint32 arrays are always bit rows (n,S,W); adjacency uses any other dtype
        bits = a if a.ndim == 3 else a[None]
        if bits.shape[1:] != (S, words(S)):
            raise ValueError(f"bit rows must have shape (n,{S},{words(S)}), got {bits.shape}")
    else:
        if a.ndim == 2:
            a = a[None]
        if a.shape[1:] != (S, S):
            raise ValueError(f"adjacency must have shape (n,{S},{S}), got {a.shape}")
        bits = pack_adj(a)
    return torch.from_numpy(np.ascontiguousarray(bits).view(np.int32)).to(device, non_blocking=False)


def egene_storage_order(d_host: np.ndarray) -> np.ndarray:
    """Storage order of the E-genes (columns of the 0/1 matrix D) for `Dataset(sort_egenes=True)`; host logic.
    The score is a sum over E-genes, so their order is free.  They are sorted by number of ones (it bounds every
    count p, so the 32 E-genes of a warp hit neighbouring table entries: conflict-free look-ups), then by the pattern
    itself (row 0 most significant; compared as big-endian 64-bit words: S/64 sort keys instead of S).  Whole tiles
    of 128 E-genes are then dealt from both ends (lightest, heaviest, second lightest, ...): tiles stay homogeneous -
    the multi-level table epilogue skips table levels a tile's E-genes cannot need - while every contiguous run of
    tiles, which is what one CTA of the score kernel gets, has the same mix of cheap and expensive tiles.  The
    ragged last tile stays last."""
    packed = np.packbits(d_host, axis=0)                                  # row 0 = MSB of byte 0
    pad = (-packed.shape[0]) % 8
    if pad:
        packed = np.concatenate([packed, np.zeros((pad, packed.shape[1]), dtype=np.uint8)])
    words = np.ascontiguousarray(packed.T).view(">u8")                    # [E][ceil(S/64)]
    keys = [words[:, w] for w in range(words.shape[1] - 1, -1, -1)] + [d_host.sum(axis=0)]
    order = np.lexsort(keys)
    nfull = len(order) // 128
    if nfull > 2:
        deal = np.empty(nfull, dtype=np.int64)
        deal[0::2] = np.arange((nfull + 1) // 2)
        deal[1::2] = nfull - 1 - np.arange(nfull // 2)
        order = np.concatenate([order[:nfull * 128].reshape(nfull, 128)[deal].ravel(), order[nfull * 128:]])
    return order


def closure(graphs, S: int | None = None, device=None, return_bits: bool = False):
    """Transitive closure of a batch of graphs on the GPU (replaces utils.ancestor, utils.py:37-54)."""
    device = _require_cuda(device)
    if S is None:
        S = np.asarray(graphs).shape[-2]
    bits = _bits_to_device(graphs, S, device)
    n = bits.shape[0]
    out = torch.empty_like(bits)
    st = torch.cuda.current_stream(device)
    _lib.check(_lib.lib().nemb200_closure_u32(_ptr(bits), n, S, _ptr(out), _stream_ptr(st)))
    if return_bits:
        return out
    return unpack_adj(out.cpu().numpy().view(np.uint32), S)


def transitive_reduction(graphs, S: int | None = None, device=None, return_bits: bool = False):
    """Transitive reduction of a batch of graphs on the GPU with the reference's exact in-place semantics
    (utils.transitive_reduction, utils.py:120-129): e.g. the best graph of every chain of a run."""
    device = _require_cuda(device)
    if S is None:
        S = np.asarray(graphs).shape[-2]
    bits = _bits_to_device(graphs, S, device)
    out = torch.empty_like(bits)
    st = torch.cuda.current_stream(device)
    _lib.check(_lib.lib().nemb200_transitive_reduction_u32(_ptr(bits), bits.shape[0], S, _ptr(out), _stream_ptr(st)))
    if return_bits:
        return out
    return unpack_adj(out.cpu().numpy().view(np.uint32), S)


@dataclass
class ChainResult:
    best_scores: np.ndarray      # (n,)
    best_orders: np.ndarray      # (n,S)
    best_graphs: np.ndarray      # (n,S,S) uint8, [a][b]=1 iff a->b
    curr_scores: np.ndarray
    curr_orders: np.ndarray
    accept_counts: np.ndarray    # (n,)
    edge_counts: np.ndarray | None   # (S,S) int64 summed over chains (post burn-in)
    score_trace: np.ndarray | None   # (n_iters, 2, n): proposed score, current score after the decision
    n_iters: int


class Dataset:
    """Device-resident observed knock-down matrix D (S x E, 0/1) + the error constants A, B (nem.py:17-18).
    Replaces the S x S x E score tables of the reference (nem.py:36-64)."""

    def __init__(self, D, A: float, B: float, device=None, sort_egenes: bool = False):
        """sort_egenes: store the E-genes (columns of D) sorted by knock-down pattern.  The score is a sum
        over E-genes, so the order is free; similar neighbours make the 32 E-genes of a warp hit the same
        exp-table entries (fewer shared-memory bank conflicts).  `self.egene_order[k]` is the original index
        of stored column k (used to un-permute per-E-gene outputs)."""
        self.device = _require_cuda(device)
        L = _lib.lib()
        if isinstance(D, torch.Tensor) and D.is_cuda and not sort_egenes:
            d = (D != 0).to(torch.uint8).to(self.device).contiguous()  # already on the device (e.g. synth.device_knockdown)
            self.egene_order = np.arange(d.shape[1])
        else:
            if isinstance(D, torch.Tensor):
                D = D.detach().cpu().numpy()
            d_host = np.ascontiguousarray((np.asarray(D) != 0).astype(np.uint8))
            if sort_egenes and d_host.shape[1] > 1:
                self.egene_order = egene_storage_order(d_host)
                d_host = np.ascontiguousarray(d_host[:, self.egene_order])
            else:
                self.egene_order = np.arange(d_host.shape[1])
            d = torch.from_numpy(d_host).to(self.device)
        self.S, self.E = int(d.shape[0]), int(d.shape[1])
        self.A, self.B = float(A), float(B)
        nbytes = L.nemb200_dataset_bytes(self.S, self.E)
        if nbytes == 0:
            raise _lib.NemB200Error(f"unsupported dataset shape S={self.S}, E={self.E} (1 <= S <= 128)")
        # over-allocate so the workspace can be 1024-byte aligned
        self._ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)
        off = (-self._ws.data_ptr()) % 1024
        self._ws_ptr = self._ws.data_ptr() + off
        self._stream = torch.cuda.Stream(self.device)  # non-default stream: CUDA-graph capture needs one
        h = C.c_void_p()
        st = torch.cuda.current_stream(self.device)
        _lib.check(L.nemb200_dataset_create(_ptr(d), self.S, self.E, self.A, self.B, C.c_void_p(self._ws_ptr),
                                            nbytes, _stream_ptr(st), C.byref(h)))
        self._h = h
        st.synchronize()  # d (the uint8 copy) may be freed after this
        info = (C.c_int64 * 10)()
        _lib.check(L.nemb200_dataset_info(self._h, info, 10))
        self.W, self.graphs_per_tile, self.lut_valid, self.scale_log2, self.n_mtiles = (
            int(info[2]), int(info[4]), bool(info[5]), int(info[6]), int(info[7]))
        self.lutm_levels = int(info[8])  # levels of the multi-level exp table (16 <= S <= 125), 0 = not usable
        self.pair_ok = bool(info[9])     # pair-merged score path available (33 <= S <= 52, single-scale table valid)
        self._score_ws = None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.lib().nemb200_dataset_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- batched score -------------------------------------------------------------------------
    def _score_workspace(self, n, impl="auto"):
        code = SCORE_IMPLS[impl] if isinstance(impl, str) else int(impl)
        need = _lib.lib().nemb200_score_workspace_bytes_impl(self._h, n, code)
        if self._score_ws is None or self._score_ws.numel() < need:
            self._score_ws = torch.empty(max(need, 1), dtype=torch.uint8, device=self.device)
        return self._score_ws, need

    def max_graphs_per_launch(self) -> int:
        """Largest batch one score launch takes: the kernel's flat tile index is 32-bit (tiles x CTAs < 4e9; the
        library refuses larger launches), so very large batches at large E are split here.  Where AUTO is the pair-merged
        path (9.7 KB of plan + tables per graph in the workspace) a launch takes at most 65 536 graphs."""
        per_tile = 2                                   # fewest graphs per MMA N tile of any score kernel
        n_tiles = int(3.0e9 // (self.n_mtiles * 160))  # 160 > the SM count
        cap = max(per_tile * n_tiles // 1024 * 1024, 1024)
        return min(cap, 65536) if (self.pair_ok and 44 <= self.S <= 50) else cap

    def _fusable(self, code) -> bool:
        """closure fused into the score (nemb200_closure_score_batch): every tensor-core kernel (the pair-merged path closes
        in its planner, the others in the score kernel's operand expansion)."""
        return self.S >= 8 and code != SCORE_IMPLS["witness"]

    def score_bits(self, bits: torch.Tensor, impl="auto", out: torch.Tensor | None = None, stream=None,
                   close: bool = False, inputs_ready: bool = False) -> torch.Tensor:
        """Score bit-packed graphs already on the device; returns a float64 CUDA tensor (no sync).
        close=True: the graphs are not transitively closed yet; the score kernel closes them while it expands its
        operand, or the pair planner does (`bits` is left untouched) - or, for the witness kernel and S < 8, a
        closed copy is made first.  inputs_ready=True: `bits` is not written by the kernel that precedes this call on
        the stream (NEMB200_SCORE_INPUTS_READY: the call may then work ahead while that kernel drains)."""
        n = int(bits.shape[0])
        if out is None:
            out = torch.empty(n, dtype=torch.float64, device=self.device)
        st = stream or torch.cuda.current_stream(self.device)
        code = SCORE_IMPLS[impl] if isinstance(impl, str) else int(impl)
        step = self.max_graphs_per_launch()
        ws, need = self._score_workspace(min(n, step), impl)
        L = _lib.lib()
        fused = close and self._fusable(code)
        if close and not fused and n:
            closed = torch.empty_like(bits)
            _lib.check(L.nemb200_closure_u32(_ptr(bits), n, self.S, _ptr(closed), _stream_ptr(st)))
            bits, inputs_ready = closed, False         # (the closure kernel just wrote them)
        fn = L.nemb200_closure_score_batch if fused else L.nemb200_score_batch
        flag = SCORE_INPUTS_READY if inputs_ready else 0
        for lo in range(0, max(n, 1), step):           # one launch unless the batch exceeds the launch limit
            m = min(step, n - lo)
            _lib.check(fn(self._h, _ptr(bits[lo:lo + m]) if n else _ptr(bits), m, _ptr(out[lo:lo + m]) if n else _ptr(out),
                          _ptr(ws), ws.numel(), code | flag, _stream_ptr(st)))
        return out

    def score_graphs(self, graphs, impl="auto", close: bool = False) -> np.ndarray:
        """NEM log-likelihood ratio of each graph (host in, host out): the reference's
        compute_ll_ratios + calculate_ll for binary weights (nem.py:72-86 as driven at :132-142).
        graphs: (n,S,S) 0/1 adjacency ([a][b]=1 iff a->b) or (n,S,W) uint32 bit rows.
        close=True applies the transitive closure first (create_nem, nem_order_mcmc.py:216-221)."""
        bits = _bits_to_device(graphs, self.S, self.device)
        return self.score_bits(bits, impl, close=close).cpu().numpy()

    def score_stream(self, host_batches, host_out, close: bool = True, impl="auto", group: int = 8):
        """Score a sequence of equally-sized batches of bit-packed graphs that live in PINNED host memory
        (`host_batches[k]`: int32 tensor (n,S,W)), writing float64 scores into the pinned `host_out[k]` (n,).
        Batches travel in groups of up to `group` (1, 2, 4, ... at the start): the host->device copies of group g+1 run on a copy stream while group g is
        closed and scored, and the scores of group g-1 go back on a third stream - two sets of device buffers, no host
        synchronisation inside (the caller synchronises once at the end).  Inside a group the score calls follow each
        other on the compute stream with nothing in between, so each call's planner works while the previous call's
        score kernel drains (dependent launches, NEMB200_SCORE_INPUTS_READY: the graphs come from the copy stream, joined by
        one event per group).  This is the end-to-end form of closure + score_batch."""
        if not len(host_batches):
            return
        dev = self.device
        n = int(host_batches[0].shape[0])
        group = max(1, int(group))
        if getattr(self, "_pipe", None) is None or self._pipe["n"] != n or self._pipe["group"] != group:
            self._pipe = {"n": n, "group": group, "copy": torch.cuda.Stream(dev), "comp": torch.cuda.Stream(dev), "back": torch.cuda.Stream(dev),
                          "bits": [torch.empty((group, n, self.S, self.W), dtype=torch.int32, device=dev) for _ in range(2)],
                          "scores": [torch.empty((group, n), dtype=torch.float64, device=dev) for _ in range(2)],
                          "ready": [torch.cuda.Event() for _ in range(2)], "free": [torch.cuda.Event() for _ in range(2)],
                          "drained": [torch.cuda.Event() for _ in range(2)],
                          "ws": torch.empty(max(_lib.lib().nemb200_score_workspace_bytes_impl(self._h, n, SCORE_IMPLS["pair"]), 1),
                                            dtype=torch.uint8, device=dev)}
        P = self._pipe
        L = _lib.lib()
        cur = torch.cuda.current_stream(dev)
        P["copy"].wait_stream(cur)
        P["comp"].wait_stream(cur)
        P["back"].wait_stream(cur)
        code = SCORE_IMPLS[impl] if isinstance(impl, str) else int(impl)
        fused = close and self._fusable(code)
        fn = L.nemb200_closure_score_batch if fused else L.nemb200_score_batch   # fused: closure inside the score call
        # group sizes ramp up 1, 2, 4, ... to `group`: the first kernels start after ONE copy, not after a whole group's
        bounds, lo, size = [], 0, 1
        while lo < len(host_batches):
            bounds.append((lo, min(lo + size, len(host_batches))))
            lo, size = lo + size, min(2 * size, group)
        for gi, (lo, hi) in enumerate(bounds):
            b = gi & 1
            members = list(range(lo, hi))
            with torch.cuda.stream(P["copy"]):
                if gi >= 2:
                    P["copy"].wait_event(P["free"][b])      # group gi-2 no longer reads these buffers
                for j, k in enumerate(members):
                    P["bits"][b][j].copy_(host_batches[k], non_blocking=True)
                    if close and not fused:  # (witness kernel, S < 8: a separate closure, beside the score kernels of the previous group)
                        _lib.check(L.nemb200_closure_u32(_ptr(P["bits"][b][j]), n, self.S, _ptr(P["bits"][b][j]), _stream_ptr(P["copy"])))
                P["ready"][b].record(P["copy"])
            with torch.cuda.stream(P["comp"]):
                P["comp"].wait_event(P["ready"][b])
                if gi >= 2:
                    P["comp"].wait_event(P["drained"][b])   # the scores of group gi-2 have left these buffers
                for j, k in enumerate(members):
                    _lib.check(fn(self._h, _ptr(P["bits"][b][j]), n, _ptr(P["scores"][b][j]), _ptr(P["ws"]), P["ws"].numel(),
                                  code | SCORE_INPUTS_READY, _stream_ptr(P["comp"])))
                P["free"][b].record(P["comp"])
            with torch.cuda.stream(P["back"]):               # scores go back on their own stream, off the kernels' path
                P["back"].wait_event(P["free"][b])
                for j, k in enumerate(members):
                    host_out[k].copy_(P["scores"][b][j], non_blocking=True)
                P["drained"][b].record(P["back"])
        cur.wait_stream(P["comp"])
        cur.wait_stream(P["back"])

    def order_graphs(self, orders, rule=RULE_FULL, cand=None, use_nem=True) -> torch.Tensor:
        """Bit rows of the graph each order induces (device tensor)."""
        o = torch.as_tensor(np.ascontiguousarray(np.asarray(orders, dtype=np.int32))).to(self.device)
        if o.ndim == 1:
            o = o[None]
        n = o.shape[0]
        cand_t = None if cand is None else _bits_to_device(np.asarray(cand)[None], self.S, self.device)[0].contiguous()
        out = torch.empty((n, self.S, self.W), dtype=torch.int32, device=self.device)
        st = torch.cuda.current_stream(self.device)
        _lib.check(_lib.lib().nemb200_order_to_graph(_ptr(o), n, self.S, int(rule), _ptr(cand_t), int(bool(use_nem)),
                                                     _ptr(out), _stream_ptr(st)))
        return out

    # ---- order-MCMC --------------------------------------------------------------------------------
    def run_chains(self, orders0, n_iters: int, gamma, swap_prob: float = 0.95, seed: int = 1234,
                   chain_offset: int = 0, rule: int = RULE_FULL, cand=None, use_nem: bool = True,
                   burn_in: int = 0, edge_counts: bool = True, trace: bool = False, impl="auto",
                   state: "ChainState | None" = None) -> "ChainState":
        """Run independent order-MCMC chains on the device (the loop of NEMOrderMCMC.method,
        nem_order_mcmc.py:272-302).  orders0: (n,S) start orders; gamma: scalar or (n,).
        Returns the device-resident ChainState (call .result() for host arrays)."""
        if state is None:
            state = ChainState(self, orders0, gamma, edge_counts, chain_offset=chain_offset, seed=seed,
                               rule=rule, cand=cand, use_nem=use_nem, burn_in=burn_in)
        state.run(n_iters, swap_prob=swap_prob, trace=trace, impl=impl)
        return state


class ChainState:
    """All per-chain MCMC state, resident in HBM (caller-owned buffers of nemb200_mcmc_run)."""

    def __init__(self, ds: Dataset, orders0, gamma, edge_counts=True, chain_offset=0, seed=1234, rule=RULE_FULL,
                 cand=None, use_nem=True, burn_in=0):
        self.ds = ds
        dev = ds.device
        o = np.ascontiguousarray(np.asarray(orders0, dtype=np.int32))
        if o.ndim == 1:
            o = o[None]
        n, S = o.shape
        if S != ds.S:
            raise ValueError(f"orders have {S} entries, dataset has S={ds.S}")
        if not (np.sort(o, axis=1) == np.arange(S, dtype=np.int32)).all():   # every row: the kernels index by label
            raise ValueError("each start order must be a permutation of range(S)")
        self.n, self.S, self.W = n, S, ds.W
        self.orders = torch.from_numpy(o).to(dev)
        g = np.broadcast_to(np.asarray(gamma, dtype=np.float64), (n,)).copy()
        self.gamma = torch.from_numpy(g).to(dev)
        self.curr_scores = torch.zeros(n, dtype=torch.float64, device=dev)
        self.curr_bits = torch.zeros((n, S, self.W), dtype=torch.int32, device=dev)
        self.best_scores = torch.zeros(n, dtype=torch.float64, device=dev)
        self.best_orders = torch.zeros((n, S), dtype=torch.int32, device=dev)
        self.best_bits = torch.zeros((n, S, self.W), dtype=torch.int32, device=dev)
        self.accept = torch.zeros(n, dtype=torch.int64, device=dev)
        self.edge_counts = torch.zeros((n, S * S), dtype=torch.int32, device=dev) if edge_counts else None
        self.cand = None if cand is None else _bits_to_device(np.asarray(cand)[None], S, dev)[0].contiguous()
        self.rule, self.use_nem, self.burn_in = int(rule), bool(use_nem), int(burn_in)
        self.seed, self.chain_offset = int(seed), int(chain_offset)
        self.iters_done = 0
        self.started = False
        self.trace = None
        need = _lib.lib().nemb200_mcmc_workspace_bytes_rule(ds._h, n, int(rule))
        self._ws = torch.empty(need + 256, dtype=torch.uint8, device=dev)
        self._ws_off = (-self._ws.data_ptr()) % 256
        self._ws_bytes = need

    def _call(self, n_iters, swap_prob, tr, impl):
        ds = self.ds
        _lib.check(_lib.lib().nemb200_mcmc_run(
            ds._h, self.n, int(n_iters), self.iters_done, self.burn_in, _ptr(self.gamma), float(swap_prob),
            C.c_uint64(self.seed), C.c_uint32(self.chain_offset), self.rule, _ptr(self.cand), int(self.use_nem),
            int(self.started), SCORE_IMPLS[impl] if isinstance(impl, str) else int(impl),
            _ptr(self.orders), _ptr(self.curr_scores), _ptr(self.curr_bits), _ptr(self.best_scores),
            _ptr(self.best_orders), _ptr(self.best_bits), _ptr(self.accept), _ptr(self.edge_counts), _ptr(tr),
            C.c_void_p(self._ws.data_ptr() + self._ws_off), self._ws_bytes, _stream_ptr(ds._stream)))
        self.started = True
        self.iters_done += int(n_iters)

    def run(self, n_iters: int, swap_prob=0.95, trace=False, impl="auto"):
        """Advance every chain by n_iters iterations (asynchronous; no host sync)."""
        ds = self.ds
        tr = torch.empty((n_iters, 2, self.n), dtype=torch.float64, device=ds.device) if trace and n_iters > 0 else None
        st = ds._stream
        st.wait_stream(torch.cuda.current_stream(ds.device))
        with torch.cuda.stream(st):
            if not self.started:  # method(): score of the start orders first (nem_order_mcmc.py:258)
                self._call(0, swap_prob, None, impl)
                self.init_scores = self.curr_scores.clone()
            self._call(n_iters, swap_prob, tr, impl)
        torch.cuda.current_stream(ds.device).wait_stream(st)
        self.trace = tr
        return self

    def replica_exchange(self, ladder: int, phase: int, round_idx: int, faithful: bool = False) -> torch.Tensor:
        """Neighbour swaps inside ladders of `ladder` consecutive chains (nem_order_mcmc.py:326-340).
        Default: delta (:333) from the chains' current scores (standard parallel tempering).  faithful=True: the
        reference's own criterion - scores[i] = replicas[i].best_score (:321), whole replicas (current AND best
        state) change places (:335); use `restart_best()` before each round to get method()'s per-call best (:259-260)."""
        ds = self.ds
        nsw = torch.zeros(1, dtype=torch.int64, device=ds.device)
        st = torch.cuda.current_stream(ds.device)
        L = _lib.lib()
        if faithful:
            _lib.check(L.nemb200_replica_exchange_best(self.n, int(ladder), int(phase), int(round_idx), C.c_uint64(self.seed),
                                                       self.S, _ptr(self.gamma), _ptr(self.orders), _ptr(self.curr_scores),
                                                       _ptr(self.curr_bits), _ptr(self.best_scores), _ptr(self.best_orders),
                                                       _ptr(self.best_bits), _ptr(nsw), _stream_ptr(st)))
        else:
            _lib.check(L.nemb200_replica_exchange(self.n, int(ladder), int(phase), int(round_idx), C.c_uint64(self.seed),
                                                  self.S, _ptr(self.gamma), _ptr(self.orders), _ptr(self.curr_scores),
                                                  _ptr(self.curr_bits), _ptr(nsw), _stream_ptr(st)))
        return nsw

    def restart_best(self):
        """method() starts every call with best_score = curr_score, best_dag / best_order = the current ones
        (nem_order_mcmc.py:259-265): device copies, no host sync.  The first run() must have scored the start orders."""
        if not self.started:
            self.run(0)
        self.best_scores.copy_(self.curr_scores)
        self.best_orders.copy_(self.orders)
        self.best_bits.copy_(self.curr_bits)

    def edge_count_matrix(self) -> torch.Tensor:
        """int64 (S,S) device tensor: per-chain posterior edge counts summed over this GPU's chains."""
        ds = self.ds
        out = torch.zeros(self.S * self.S, dtype=torch.int64, device=ds.device)
        if self.edge_counts is not None:
            st = torch.cuda.current_stream(ds.device)
            _lib.check(_lib.lib().nemb200_edge_counts_reduce(_ptr(self.edge_counts), self.n, self.S, _ptr(out),
                                                             _stream_ptr(st)))
        return out.view(self.S, self.S)

    def result(self) -> ChainResult:
        S = self.S
        ec = self.edge_count_matrix().cpu().numpy() if self.edge_counts is not None else None
        return ChainResult(
            best_scores=self.best_scores.cpu().numpy(), best_orders=self.best_orders.cpu().numpy(),
            best_graphs=unpack_adj(self.best_bits.cpu().numpy().view(np.uint32), S),
            curr_scores=self.curr_scores.cpu().numpy(), curr_orders=self.orders.cpu().numpy(),
            accept_counts=self.accept.cpu().numpy(), edge_counts=ec,
            score_trace=None if self.trace is None else self.trace.cpu().numpy(), n_iters=self.iters_done)
