"""minilp_b200 — B200-native revised-simplex pivot engine behind ztlpn/minilp's Problem/Solution API.

Python here is only the host-side mirror of the reference's public interface (lib.rs:61-464) over the
C ABI of libminilp_b200.so; all bulk arithmetic runs in hand-written sm_100a CUDA kernels
(minilp_b200/csrc/engine.cu).  There is no CPU fallback: without the built library or without a CUDA
device the calls raise.
"""
from .api import (ComparisonOp, DenseLP, Engine, Error, Infeasible, LocalGroup, OptimizationDirection, Problem, Solution,
                  Solver, Unbounded, device_count, nccl_unique_id, reduce_candidates, shard_range, synth_block, synth_dense,
                  synth_rows, synth_vectors)

__all__ = ["ComparisonOp", "DenseLP", "Engine", "Error", "Infeasible", "LocalGroup", "OptimizationDirection", "Problem",
           "Solution", "Solver", "Unbounded", "device_count", "nccl_unique_id", "reduce_candidates", "shard_range",
           "synth_block", "synth_dense", "synth_rows", "synth_vectors"]
