"""Accumulator lane of output row i of a tcgen05 MMA with M = 64 (cta_group::1): 16 rows per 32-lane quadrant.
Measured by tests/test_umma.py::test_umma_dw_probe; see csrc/umma.cuh."""


def dw_lane(i: int, D: int) -> int:
    return 32 * (i // 16) + i % 16
