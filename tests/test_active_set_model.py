"""CPU statement of the line solver's active set (tests/models/active_set_model.py): skipping the columns whose right-hand side
and whose neighbours' iterates are still exactly zero never changes an iterate, and saves work when saltation is patchy."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "models"))
import active_set_model as am  # noqa: E402


def test_skipped_updates_are_no_ops():
    for fo in (None, am.patchy):
        sy = am.build(70, forcing=fo)
        k0, x0, d0 = am.solve(sy, False)
        k1, x1, d1 = am.solve(sy, True)
        assert k0 is not None and k0 == k1
        assert np.array_equal(x0, x1)
        assert all(a <= b for a, b in zip(d1, d0)) and d1[0] < d0[0]
        # the live set only grows, one ring of faces per colour pass
        assert all(a <= b for a, b in zip(d1, d1[1:]))
