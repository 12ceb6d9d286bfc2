"""GPU variant of test_remove_edges.py::test_remove_other_row_then_add (written after round 1's GPU budget was spent: sorts behind the
established suite)."""
import pytest

from test_remove_edges import _graph_sim, _check_reversed_cycle


@pytest.mark.gpu
@pytest.mark.parametrize("ET", ["EdgeD", "EdgeS", "EdgeT"])
def test_remove_other_row_then_add_gpu(cuda, ET):
    _check_reversed_cycle(_graph_sim(cuda, ET, "cycle"), ET)
