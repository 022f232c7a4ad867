"""CPU oracle for the xrt_b200 hot path -- TEST INFRASTRUCTURE ONLY.

A numpy restatement of the reference's numpy path (kklmn/xrt @ v2.0.0b1), function by function, each citing
the reference file:line it follows.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import this package; the product (`xrt_b200/`) never does and
fails loudly when its CUDA library is missing.

Pinning: the reference publishes no asserted numbers for this path, so the oracle is
pinned against outputs of the reference itself, run in the build container by `tools/gen_golden.py` and
committed under `tests/golden/` (tests/test_oracle_golden.py replays them).
"""
