"""First GPU gate: the C-ABI path against the oracle on small seeded inputs (details in test_gpu_dsm/ortho)."""
import pytest


@pytest.mark.gpu
def test_smoke_matches_oracle():
    import __graft_entry__ as g
    g.smoke()
