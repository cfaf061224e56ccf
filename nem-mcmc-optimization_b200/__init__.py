"""nem_b200 - B200-native (sm_100a) NEM scoring inside order-MCMC.

Drop-in mirrors of the reference's flat modules (MrGreyPanda/NEM-MCMC-optimization):
    nem_b200.nem.NEM                      <- nem.py
    nem_b200.nem_order_mcmc.NEMOrderMCMC  <- nem_order_mcmc.py
    nem_b200.utils                        <- utils.py
plus the batched device entry points in nem_b200.engine (Dataset.score_graphs, closure, run_chains).
All compute runs in hand-written CUDA behind the C ABI of include/nemb200.h; there is no CPU fallback.
"""
from . import _lib, engine, utils, nem, nem_order_mcmc, parallel, synth  # noqa: F401
from .engine import Dataset, ChainState, ChainResult, closure, transitive_reduction, pack_adj, unpack_adj, RULE_FULL, RULE_MASKED, RULE_FIT  # noqa: F401
from .nem import NEM  # noqa: F401
from .nem_order_mcmc import NEMOrderMCMC, replica_exchange_method  # noqa: F401

__all__ = ["Dataset", "ChainState", "ChainResult", "closure", "transitive_reduction", "pack_adj", "unpack_adj", "NEM", "NEMOrderMCMC",
           "replica_exchange_method", "RULE_FULL", "RULE_MASKED", "RULE_FIT"]
