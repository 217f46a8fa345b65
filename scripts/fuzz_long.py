"""Extended randomised parity sweep (GPU path vs oracle), beyond the cases in tests/test_fuzz_gpu.py.
Usage: python scripts/fuzz_long.py [trials] [first_seed]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from resolv_b200 import _abi, space_for_world, worlds  # noqa: E402
from tests.test_fuzz_gpu import _lattice_world  # noqa: E402
from tests.util import OracleBatch, bits, compare_frame, device_cells  # noqa: E402

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 200
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
t0 = time.time()
tot = dict(worlds=0, frames=0, candidates=0, hits=0, rays=0, ray_hits=0)
for t in range(trials):
    rs = np.random.default_rng(seed0 + t)
    kind = t % 4
    if kind == 0:
        w = _lattice_world(rs, int(rs.integers(50, 600)), int(rs.choice([128, 256, 300])), int(rs.choice([8, 16, 24])), int(rs.choice([4, 8])))
    elif kind == 1:
        w = worlds.make_mixed(int(rs.integers(2, 3000)), seed=int(rs.integers(1, 1 << 30)), space=int(rs.choice([64, 257, 1024, 2000])),
                              cell=int(rs.choice([7, 16, 32, 50])), filtered=bool(rs.integers(0, 2)))
        w.pos = w.pos * float(rs.uniform(0.3, 1.5)) - float(rs.uniform(0, 0.3)) * w.space_w   # crowded and/or partly off-grid
        w.pos0 = w.pos.copy()
        w.finalize()
    elif kind == 2:
        w = worlds.make_rects(int(rs.integers(10, 4000)), seed=int(rs.integers(1, 1 << 30)), space=int(rs.choice([512, 1500])), cell=int(rs.choice([16, 32])))
    else:
        w = worlds.make_platformer_batch(int(rs.integers(1, 5)), seed=int(rs.integers(1, 1 << 30)))
    if kind in (0, 1) and rs.integers(0, 2):
        w.tester[:] = rs.integers(0, 2, w.n)
    if kind == 0 and rs.integers(0, 2):
        w.use_any[:] = rs.integers(0, 2, w.n)
        w.any_mask[:] = rs.integers(0, 8, w.n)
        w.none_mask[:] = rs.integers(0, 8, w.n) & rs.integers(0, 8, w.n)
    ob = OracleBatch(w)
    ds = space_for_world(w, max_rays=256)
    try:
        margin = int(rs.integers(0, 5)) if kind != 3 else w.margin
        frames = int(rs.integers(1, 5))
        for f in range(frames):
            if f:
                worlds.advance(w, f)
                if w.vel is None:
                    w.pos = w.pos + rs.uniform(-3, 3, w.pos.shape)
                ob = OracleBatch(w)              # fresh Space: the definition of the device's generation rank
                ds.set_positions(w.pos)
            c, exact = compare_frame(ds, ob, margin)
            assert exact, (t, f)
            tot["frames"] += 1; tot["candidates"] += int(c.n_candidates); tot["hits"] += int(c.n_hits)
        assert device_cells(ds, c.n_entries) == ob.cells(), t
        if w.n_spaces == 1:
            rays = worlds.make_rays(200, int(rs.integers(1, 1 << 20)), w.space_w, w.space_h, 4, max(w.space_w / 2, 8))
            rc = ds.linetest(rays, flags=_abi.FRAME_DOWNLOAD)
            hits, con, first, count = ds.ray_results()
            fc, oh = ob.spaces[0].linetest(rays, mode="rect")
            o = np.lexsort((oh.b, oh.a)); d = np.lexsort((hits["b"], hits["ray"]))
            assert np.array_equal(hits["ray"][d], oh.a[o]) and np.array_equal(hits["b"][d], oh.b[o]), t
            assert np.array_equal(bits(hits["mtv"][d]), bits(oh.mtv[o])) and np.array_equal(bits(hits["key"][d]), bits(oh.dist[o])), t
            tot["rays"] += len(rays); tot["ray_hits"] += len(o)
    finally:
        ds.close()
    tot["worlds"] += 1
print(tot, f"{time.time() - t0:.0f} s", "ALL BIT-IDENTICAL")
