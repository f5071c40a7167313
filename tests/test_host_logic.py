"""Host-side helpers (no GPU, no kernels): sharding of sweeps, four-step line splits and the digit-transposed
order, separability detection — property tests."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from spinor_gpe_b200._separable import split_separable
from spinor_gpe_b200.slab import MAX_LINE, digit_order, four_step_split
from spinor_gpe_b200.sweep import shard


@given(n=st.integers(0, 300), world=st.integers(1, 16))
@settings(max_examples=60, deadline=None)
def test_shards_cover_every_trajectory_once_and_are_balanced(n, world):
    parts = [shard(n, r, world) for r in range(world)]
    assert sorted(i for p in parts for i in p) == list(range(n))
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


@pytest.mark.parametrize('n', [32, 1024, 4096, 8192, 16384, 32768, 65536])
def test_four_step_split_fits_the_kernels(n):
    n1 = four_step_split(n)
    if n <= MAX_LINE:
        assert n1 == 1
        np.testing.assert_array_equal(digit_order(n, n1), np.arange(n))
        return
    n2 = n // n1
    assert n1 * n2 == n and 32 <= n1 <= 1024 and 32 <= n2 <= MAX_LINE       # sgpe_plan_create_lines' limits
    nat = digit_order(n, n1)
    assert sorted(nat) == list(range(n))                                   # a permutation
    # position k1 * n2 + k2 holds frequency k1 + n1 * k2
    k1, k2 = np.divmod(np.arange(n), n2)
    np.testing.assert_array_equal(nat, k1 + n1 * k2)
    assert four_step_split(n, forced=64) == 64


@given(seed=st.integers(0, 2 ** 16), ny=st.sampled_from([4, 8, 32]), nx=st.sampled_from([4, 16, 32]),
       scale=st.floats(1e-3, 1e6))
@settings(max_examples=40, deadline=None)
def test_separable_grids_are_split_exactly_and_others_rejected(seed, ny, nx, scale):
    rng = np.random.default_rng(seed)
    gx, gy = rng.normal(size=(2, nx)) * scale, rng.normal(size=(2, ny)) * scale
    grid = gx[:, None, :] + gy[:, :, None]
    out = split_separable(grid)
    assert out is not None
    np.testing.assert_allclose(out[0][:, None, :] + out[1][:, :, None], grid, rtol=0, atol=1e-12 * scale * 10)
    assert (out[1].min(axis=1) == 0).all()                                 # anchored at the smallest row: gy >= 0
    bumped = grid.copy()
    bumped[0, ny // 2, nx // 2] += 1e-6 * scale                              # one pixel off: not separable any more
    assert split_separable(bumped) is None


def test_all_zero_grid_is_separable():
    out = split_separable(np.zeros((2, 8, 16)))                            # trap switched off (time of flight)
    assert out is not None and not out[0].any() and not out[1].any()
