#!/bin/bash
# r02 session 28: the bit-identity test of the one-round-trip dual iteration (last seconds of the GPU budget)
mkdir -p gpurun_out/r02s28
timeout 40 python -m pytest "tests/test_parity_gpu.py::test_one_round_trip_dual_iteration_is_bit_identical" -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r02s28/test.log
