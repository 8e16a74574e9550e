"""CPU: host-side arithmetic of the step pipeline (no kernel is launched)."""
import pytest


@pytest.mark.parametrize("n", [0, 1, 2047, 2048, 2049, 3 * 32 * 32 * 1024, 64 * 128 * 64 * 2048])
@pytest.mark.parametrize("shares", [(0, 0, 0, 0), (0, 0, 0, 0.5), (0.15, 0.05, 0.1, 0.1), (0.3, 0.2, 0.25, 0.25),
                                    (0.001, 0.5, 0, 0), (0.9, 0.9, 0.9, 0.9), (-1, 0, 0, 2.0)])
@pytest.mark.parametrize("chains", [True, False])
def test_fill_partition_tiles_the_buffer(n, shares, chains):
    """The carried zero fill: the five carriers' ranges must tile the gradient buffer exactly --
    a gap would leave garbage in the gradient, an overlap would be written twice -- with every
    inner boundary on an 8 KB page (the carriers copy whole zero pages)."""
    from coarse3d_b200.pipeline import fill_partition
    parts = fill_partition(n, shares, chains)
    assert len(parts) == 5
    lo = 0
    for j, (a, b) in enumerate(parts):
        assert a == lo and a <= b <= n
        if j < 4:
            assert b % 2048 == 0 or b == n
            want = max(shares[j], 0.0) if (chains or j == 0) else 0.0
            assert b - a <= int(n * want) + 1e-9 or b == n
        lo = b
    assert lo == n
    if not chains:
        assert all(a == b for a, b in parts[1:4])


def test_step_class_keeps_the_methods_the_bench_uses():
    """bench.py and the tools drive HotPathStep through these names (a refactor once dropped one)."""
    import re
    import os
    from coarse3d_b200.pipeline import HotPathStep
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    used = set(re.findall(r"\bstep\.([a-z_]+)\(", open(os.path.join(root, "bench.py")).read()))
    assert {"run", "capture", "step", "algorithmic_bytes", "fill_bytes_by_carrier"} <= used
    for name in used:
        assert hasattr(HotPathStep, name), name
