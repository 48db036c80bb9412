"""peps_b200: B200-native walker-batched VMC sampling hot path of QuantumLiquids/PEPS.

Python here is host-side plumbing over the C ABI (include/peps_b200.h); all arithmetic runs in the hand-written
sm_100a kernels of peps_b200/csrc/backend_cuda.cu. See DESIGN.md.
"""
from .api import (BMPSTruncateParams, MonteCarloParams, SplitIndexTPS, FermionSplitIndexTPS, TableModel, Configuration, SquareSpinOneHalfXXZModelOBC,
                  SquareSpinOneHalfJ1J2XXZModelOBC, TransverseFieldIsingSquareOBC, MCUpdateSquareNNExchange,
                  MCUpdateSquareNNFullSpaceUpdate, MCUpdateSquareTNN3SiteExchange, WalkerBatch, MCEnergyGradEvaluator, MCPEPSMeasurer, PepsError)

__all__ = ["BMPSTruncateParams", "MonteCarloParams", "SplitIndexTPS", "FermionSplitIndexTPS", "TableModel", "Configuration",
           "SquareSpinOneHalfXXZModelOBC", "SquareSpinOneHalfJ1J2XXZModelOBC", "TransverseFieldIsingSquareOBC", "MCUpdateSquareNNExchange",
           "MCUpdateSquareNNFullSpaceUpdate", "MCUpdateSquareTNN3SiteExchange", "WalkerBatch", "MCEnergyGradEvaluator", "MCPEPSMeasurer",
           "PepsError"]
