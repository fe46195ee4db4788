"""GPU-vs-oracle parity through the C ABI (the parity tests proper). Protocol: tests/parity.py."""
import numpy as np
import pytest

import parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(parity.CASES))
def test_case_bit_exact(backend, name):
    """Face-number sequences, world exit directions and weights bit-exact against the oracle replay of
    the engine's own roots; image within IMG_RTOL of the oracle accumulation of the same exits."""
    res = parity.run_case(parity.CASES[name], n_rays=30000, seed=42, backend=backend)
    assert res["exits"] > 0
    assert res["paths_equal"], res
    assert res["dirs_bit_equal"], res
    assert res["weights_bit_equal"], res
    assert res["meta_equal"], res
    assert res["stats_ok"], res
    assert res["image_ok"], res
    assert res["landed_rel_err"] < 1e-5, res


def test_seed_and_tile_invariance(backend):
    """Counter-based RNG: the same seed gives the same exits whatever the tile size; another seed differs."""
    a = parity.run_case(parity.CASES["column_config2"], n_rays=20000, seed=7, backend=backend, tile_rays=4096)
    b = parity.run_case(parity.CASES["column_config2"], n_rays=20000, seed=7, backend=backend, tile_rays=1 << 22)
    c = parity.run_case(parity.CASES["column_config2"], n_rays=20000, seed=8, backend=backend, tile_rays=1 << 22)
    assert a["paths_equal"] and b["paths_equal"] and c["paths_equal"]
    assert a["exits"] == b["exits"]
    assert abs(a["landed_gpu"] - b["landed_gpu"]) <= 1e-5 * abs(a["landed_gpu"])
    assert a["exits"] != c["exits"] or a["landed_gpu"] != c["landed_gpu"]
