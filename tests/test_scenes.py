"""CPU: the synthetic scene generators that bench.py and the strip tests rely on."""
import math

import numpy as np

from ferox_b200 import capi, scenes


def test_strip_pile_strips_tile_the_world_without_overlap():
    n_strips, per = 4, 5000
    specs = [scenes.strip_pile(per, k, n_strips, storeys=2) for k in range(n_strips)]
    tables = [[(s.kind, s.params, s.density, s.friction, s.restitution) for s in sp.shapes] for sp, _, _ in specs]
    assert all(t == tables[0] for t in tables)                  # ghost rows carry shape-table indices
    prev_hi = -math.inf
    for k, (sp, lo, hi) in enumerate(specs):
        dyn = sp.body_type == capi.BODY_DYNAMIC
        assert dyn.sum() == per
        x = sp.pos[dyn, 0]
        assert (x >= (lo if k else 0.0)).all() and (x < hi).all()
        assert lo == prev_hi or k == 0
        prev_hi = hi
        # one floor segment per storey, reaching 64 units past both cuts; end walls only at the ends
        stat = ~dyn
        floors = (sp.body_shape[stat] == 0).sum()
        walls = (sp.body_shape[stat] == 1).sum()
        assert floors == 2 and walls == (k == 0) + (k == n_strips - 1)
        assert sp.body_shape[dyn].min() >= 2 and sp.body_shape[dyn].max() <= 8
    assert specs[0][1] == -math.inf and specs[-1][2] == math.inf
    # different strips draw different bodies
    assert not np.array_equal(specs[0][0].angle[-100:], specs[1][0].angle[-100:])


def test_mixed_pile_depth_is_bounded():
    for n in (3000, 65536):
        sp = scenes.mixed_pile(n)
        dyn = sp.body_type == capi.BODY_DYNAMIC
        rows = (sp.pos[dyn, 1].max() - sp.pos[dyn, 1].min()) / 1.7
        assert rows <= 97, rows                                 # deeper piles diverge in the reference's own arithmetic


def test_column_collapse_defaults_are_stable_and_jitter_is_a_parameter():
    a = scenes.column_collapse(4096, cols=64)
    b = scenes.column_collapse(4096, cols=64)
    assert np.array_equal(a.pos, b.pos)
    c = scenes.column_collapse(4096, cols=64, pitch=1.2, jitter=0.2)
    d = np.abs(c.pos[2:, 0] - (0.5 + 1.2 * (np.arange(4096) % 64)))
    assert d.max() <= 0.1 + 1e-5 and d.max() > 0.05
