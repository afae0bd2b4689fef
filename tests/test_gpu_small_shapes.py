"""GPU tier: every kernel on small, ragged shapes (path counts that are not multiples of the tile, vector or warp
size; 3..12 steps; both storage types; barrier / exercise-step / SVD / batch / exposures variants).  compute-sanitizer
is not available on the GPU pool, so this walk -- with every result checked finite and shapes checked -- plus the oracle
parity tests is what guards the indexing."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_all_kernels_on_ragged_small_shapes(amc):
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import sanitize_small
    sanitize_small.main()
