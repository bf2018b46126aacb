"""Random configurations of the two consumers of the mask tiles (SummaryOutput rows, the three-layer overlay) against the
NumPy oracles: a slice of the soak of tools/fuzz_consumers.py (15,000 configurations, profiles/fuzz_consumers_r02.txt)
that runs with every GPU test pass.  Frame sizes, instance counts, box widths on the lane-layout boundaries of the box
reduction, tile densities, semantic-map kinds and frame dtypes are drawn per seed."""
import importlib.util
import os

import pytest

pytestmark = pytest.mark.gpu

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fuzz():
    spec = importlib.util.spec_from_file_location("fuzz_consumers", os.path.join(_ROOT, "tools", "fuzz_consumers.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("first", [0, 40, 80])
def test_consumers_random(first):
    fc = _fuzz()
    for seed in range(first, first + 40):
        assert fc.run(seed) == [], "seed %d" % seed
