"""Parity of the batched LineTest kernel (through the C ABI) against the CPU oracle's LineTest (utils.go:276-370).

RECT iterator: results must equal the oracle run with the reference-expressible TestAgainst
space.FilterCells(bounds of the cast line).FilterShapes() — same sets, same generation order, bit-identical values.
DDA iterator: every reported set must be bit-identical to the brute-force oracle's set for that (ray, shape); any
brute-force set the traversal does not report must be a "phantom" of the +1 fudge (its shape is registered in no
cell that the exact supercover of the cast line touches), which is checked with exact rational arithmetic.
"""
from fractions import Fraction

import numpy as np
import pytest

from resolv_b200 import _abi, space_for_world, worlds
from tests.util import OracleBatch, bits

pytestmark = pytest.mark.gpu


def _canon_dev(hits, con):
    o = np.lexsort((hits["b"], hits["ray"]))
    h = hits[o]
    n = h["count_rank"] & 0xff
    c = np.concatenate([np.concatenate([con["p"][f:f + k], con["n"][f:f + k]], -1)
                        for f, k in zip(h["first_contact"], n)]) if len(h) else np.zeros((0, 4))
    return h, n, c


def _canon_ora(oh):
    o = np.lexsort((oh.b, oh.a))
    c = np.concatenate([oh.contacts[f:f + k] for f, k in zip(oh.first[o], oh.count[o])]) if len(o) else np.zeros((0, 4))
    return o, c


def _setup(n_shapes=20000, space=9280, n_rays=20000, filtered=False, seed=7):
    w = worlds.make_mixed(n_shapes, space=space, cell=32, filtered=False)
    ds = space_for_world(w)
    ds.frame(1, _abi.FRAME_GRID_ONLY)
    ob = OracleBatch(w)
    rays = worlds.make_rays(n_rays, seed, w.space_w, w.space_h)
    kw = {}
    if filtered:
        kw = dict(use_any=np.ones(n_rays, np.uint8), any_mask=np.full(n_rays, 0b0110, np.uint64),
                  none_mask=np.full(n_rays, 0b1000, np.uint64))
    return w, ds, ob.spaces[0], rays, kw


@pytest.mark.parametrize("filtered", [False, True])
def test_rect_iterator_matches_reference_filtercells(filtered):
    w, ds, sp, rays, kw = _setup(filtered=filtered)
    try:
        c = ds.linetest(rays, **kw)
        hits, con, first, count = ds.ray_results()
        fc, oh = sp.linetest(rays, mode="rect", **kw)
        assert c.n_candidates == fc.candidates
        assert c.n_hits == fc.hits == len(hits) and c.n_contacts == fc.contacts
        assert c.n_hits > 1000
        # grouping: records of ray i are hits[first[i] : first[i] + count[i]], in generation order
        assert int(count.sum()) == len(hits)
        cnt = count.astype(np.int64)
        idx = np.repeat(first.astype(np.int64), cnt) + (np.arange(int(cnt.sum())) - np.repeat(np.cumsum(cnt) - cnt, cnt))
        assert np.array_equal(hits["ray"][idx], np.repeat(np.arange(len(rays)), cnt))
        rk = (hits["count_rank"][idx] >> 8).astype(np.int64)
        same = hits["ray"][idx][1:] == hits["ray"][idx][:-1]
        assert np.all(rk[1:][same] > rk[:-1][same]), "records of a ray are not in generation order"
        h, n, dc = _canon_dev(hits, con)
        o, oc = _canon_ora(oh)
        assert np.array_equal(h["ray"], oh.a[o]) and np.array_equal(h["b"], oh.b[o])
        assert np.array_equal(n, oh.count[o])
        for name, dev, ora in (("contacts", dc, oc), ("mtv", h["mtv"], oh.mtv[o]), ("center", h["center"], oh.center[o]),
                               ("key", h["key"], oh.dist[o])):
            assert np.all(np.abs(dev - ora) <= 1e-9), name
            assert np.array_equal(bits(dev), bits(ora)), f"{name} within 1e-9 but not bit-identical"
        # callback order: ascending key, ties by generation rank == the oracle's callback index
        rank = h["count_rank"] >> 8
        order = np.lexsort((rank, h["key"], h["ray"]))
        cb = np.zeros(len(h), dtype=np.int64)
        pos = np.arange(len(h))
        starts = np.r_[0, np.nonzero(np.diff(h["ray"][order]))[0] + 1]
        cb[order] = pos - np.repeat(starts, np.diff(np.r_[starts, len(h)]))
        assert np.array_equal(cb, oh.order[o])
    finally:
        ds.close()


def test_rect_iterator_is_brute_force_restricted_to_the_shapes_under_the_rect():
    """Where the RECT iterator stands against TestAgainst = space.FilterShapes() (every shape, the brute force of SURVEY
    §3.4): every set it reports is the brute-force set of that (ray, shape), bit for bit, and every brute-force set it
    does not report belongs to a shape that is registered in NO cell under the cast line's bounds — a contact the "+1"
    fudge of IntersectionPointsLine (convexPolygon.go:721-726) invents for a far-away, nearly parallel edge, which no
    cell-based TestAgainst of the reference would be handed either. Exact integer check, no tolerance."""
    w, ds, sp, rays, kw = _setup(n_rays=30000)
    try:
        ds.linetest(rays)
        hits, con, first, count = ds.ray_results()
        fc, oh = sp.linetest(rays, mode="all")
        h, n, dc = _canon_dev(hits, con)
        o, oc = _canon_ora(oh)
        dev_keys = h["ray"].astype(np.int64) * w.n + h["b"]
        ora_keys = oh.a[o].astype(np.int64) * w.n + oh.b[o]
        assert np.isin(dev_keys, ora_keys).all(), "RECT reported a set brute force does not have"
        sel = np.isin(ora_keys, dev_keys)
        osel = o[sel]
        ocs = np.concatenate([oh.contacts[f:f + k] for f, k in zip(oh.first[osel], oh.count[osel])])
        assert np.array_equal(n, oh.count[osel]) and np.array_equal(bits(dc), bits(ocs))
        assert np.array_equal(bits(h["mtv"]), bits(oh.mtv[osel])) and np.array_equal(bits(h["key"]), bits(oh.dist[osel]))
        gw, gh = -(-w.space_w // w.cell_w), -(-w.space_h // w.cell_h)
        for m in o[~sel]:
            r, b = int(oh.a[m]), int(oh.b[m])
            s_, e_ = rays[r, :2], rays[r, 2:]
            vu = (e_ - s_) / np.hypot(*(e_ - s_))
            st = s_ - 0.01 * vu
            # cell rect of the cast line's bounds (utils.go:90-98), clipped to the grid, against the shape's cell rect
            qx0, qx1 = sorted((int(np.floor(st[0] / w.cell_w)), int(np.floor(e_[0] / w.cell_w))))
            qy0, qy1 = sorted((int(np.floor(st[1] / w.cell_h)), int(np.floor(e_[1] / w.cell_h))))
            bx0, by0, bx1, by1 = [int(v) for v in sp.cell_rect(b)]
            shared = (max(qx0, bx0, 0) <= min(qx1, bx1, gw - 1)) and (max(qy0, by0, 0) <= min(qy1, by1, gh - 1))
            assert not shared, f"ray {r}: shape {b} shares a cell with the cast line's bounds but RECT did not report its set"
    finally:
        ds.close()


def _supercover_touches(rays_i, cell_rect, cw, ch):
    """Exact: does the closed segment (pulled-back start -> end) touch any cell of the inclusive rect?"""
    (x0, y0, x1, y1) = [Fraction(float(v)) for v in rays_i]
    cx0, cy0, cx1, cy1 = [int(v) for v in cell_rect]
    bx0, by0, bx1, by1 = Fraction(cx0 * cw), Fraction(cy0 * ch), Fraction((cx1 + 1) * cw), Fraction((cy1 + 1) * ch)
    t0, t1 = Fraction(0), Fraction(1)
    for p, d, lo, hi in ((x0, x1 - x0, bx0, bx1), (y0, y1 - y0, by0, by1)):
        if d == 0:
            if p < lo or p > hi:
                return False
        else:
            u0, u1 = (lo - p) / d, (hi - p) / d
            if u0 > u1:
                u0, u1 = u1, u0
            t0, t1 = max(t0, u0), min(t1, u1)
    return t0 <= t1


def _segment_touches_box(seg, box, grow):
    """Exact: does the closed segment touch the box (minx, miny, maxx, maxy) grown by `grow` on every side?"""
    (x0, y0, x1, y1) = [Fraction(float(v)) for v in seg]
    g = Fraction(grow)
    bx0, by0, bx1, by1 = Fraction(float(box[0])) - g, Fraction(float(box[1])) - g, Fraction(float(box[2])) + g, Fraction(float(box[3])) + g
    t0, t1 = Fraction(0), Fraction(1)
    for p, d, lo, hi in ((x0, x1 - x0, bx0, bx1), (y0, y1 - y0, by0, by1)):
        if d == 0:
            if p < lo or p > hi:
                return False
        else:
            u0, u1 = (lo - p) / d, (hi - p) / d
            if u0 > u1:
                u0, u1 = u1, u0
            t0, t1 = max(t0, u0), min(t1, u1)
    return t0 <= t1


def test_dda_iterator_bit_identical_sets_and_only_phantoms_missing():
    w, ds, sp, rays, kw = _setup(n_rays=30000)
    try:
        c = ds.linetest(rays, flags=_abi.FRAME_DOWNLOAD | _abi.LINE_DDA)
        hits, con, first, count = ds.ray_results()
        fc, oh = sp.linetest(rays, mode="all")
        h, n, dc = _canon_dev(hits, con)
        o, oc = _canon_ora(oh)
        dev_keys = h["ray"].astype(np.int64) * w.n + h["b"]
        ora_keys = oh.a[o].astype(np.int64) * w.n + oh.b[o]
        assert len(np.unique(dev_keys)) == len(dev_keys), "a shape was tested twice for one ray (dedupe broken)"
        assert np.isin(dev_keys, ora_keys).all(), "DDA reported a set brute force does not have"
        sel = np.isin(ora_keys, dev_keys)
        # bit-identical values for every reported set
        osel = o[sel]
        ocs = np.concatenate([oh.contacts[f:f + k] for f, k in zip(oh.first[osel], oh.count[osel])])
        assert np.array_equal(n, oh.count[osel])
        assert np.array_equal(bits(dc), bits(ocs))
        assert np.array_equal(bits(h["mtv"]), bits(oh.mtv[osel])) and np.array_equal(bits(h["center"]), bits(oh.center[osel]))
        assert np.array_equal(bits(h["key"]), bits(oh.dist[osel]))
        # anything missing must be a phantom of the "+1" fudge: the cast line does not come within 0.9 world units of
        # the shape's Bounds (the device gate is 1.0), or the shape is in no cell the exact supercover of the line touches
        missing = o[~sel]
        assert len(missing) <= 0.01 * len(o), (len(missing), len(o))
        for m in missing:
            r, b = int(oh.a[m]), int(oh.b[m])
            s, e = rays[r, :2], rays[r, 2:]
            vu = (e - s) / np.hypot(*(e - s))
            st = s - 0.01 * vu
            seg = (st[0], st[1], e[0], e[1])
            assert (not _segment_touches_box(seg, sp.bounds(b), 0.9)
                    or not _supercover_touches(seg, sp.cell_rect(b), w.cell_w, w.cell_h)), \
                f"ray {r} misses shape {b} although the cast line touches its bounds"
        assert c.n_candidates < fc.candidates
    finally:
        ds.close()


def test_kat5_through_the_device_both_iterators():
    rows = [dict(kind=1, pos=(10.0, 5.0), radius=0.0, verts=worlds.rect_verts(4.0, 4.0), tags=1)]
    w = worlds._from_rows(rows, (64, 64), (16, 16), "kat5").finalize()
    ds = space_for_world(w)
    try:
        ds.frame(1, _abi.FRAME_GRID_ONLY)
        for fl in (0, _abi.LINE_DDA):
            c = ds.linetest(np.array([[0.0, 5.0, 20.0, 5.0]]), flags=_abi.FRAME_DOWNLOAD | fl)
            hits, con, first, count = ds.ray_results()
            assert c.n_hits == 1 and (hits["count_rank"][0] & 0xff) == 2
            assert np.abs(con["p"] - np.array([[7.75, 5.0], [12.250000000000002, 5.0]])).max() < 1e-12
            assert np.abs(hits["mtv"][0] - np.array([7.74, 0.0])).max() < 1e-12
            assert np.abs(hits["center"][0] - np.array([10.0, 5.0])).max() < 1e-12
    finally:
        ds.close()


def test_rays_outside_and_degenerate():
    w, ds, sp, _, _ = _setup(n_shapes=3000, space=1024, n_rays=8)
    try:
        rays = np.array([[-500.0, -500.0, -100.0, -100.0],     # entirely outside the grid
                         [-200.0, 300.0, 1500.0, 310.0],       # crosses the whole Space
                         [100.0, 100.0, 100.0, 100.0],         # zero length (Unit leaves (0,0) alone)
                         [512.0, -300.0, 512.0, 1400.0],       # exactly on a cell boundary column
                         [0.0, 0.0, 1024.0, 1024.0],           # through cell corners
                         [1000.0, 1000.0, 10.0, 20.0]])        # backwards
        for fl in (0, _abi.LINE_DDA):
            c = ds.linetest(rays, flags=_abi.FRAME_DOWNLOAD | fl)
            hits, con, first, count = ds.ray_results()
            fc, oh = sp.linetest(rays, mode="rect" if fl == 0 else "all")
            dk = set(zip(hits["ray"].tolist(), hits["b"].tolist()))
            ok = set(zip(oh.a.tolist(), oh.b.tolist()))
            if fl == 0:
                assert dk == ok
            else:
                assert dk <= ok and len(ok - dk) <= 2
    finally:
        ds.close()


def test_platformer_probes_on_a_batch_of_spaces():
    """Config 4: four probes per dynamic body, ByTags(Platform), 6 independent Spaces in one rsv_space (the ray's Space
    rides in the mask word). RECT iterator == the oracle's FilterCells iterator, per Space, bit for bit."""
    from tests.util import OracleBatch
    w = worlds.make_platformer_batch(6)
    rays, use_any, any_mask, sidx = worlds.make_platformer_probes(w)
    ds = space_for_world(w, max_rays=len(rays))
    ob = OracleBatch(w)
    try:
        ds.frame(w.margin, _abi.FRAME_GRID_ONLY)
        for fl, mode in ((0, "rect"), (_abi.LINE_DDA, "all")):
            ds.linetest(rays, use_any=use_any, any_mask=any_mask, flags=_abi.FRAME_DOWNLOAD | fl, space_idx=sidx)
            hits, con, first, count = ds.ray_results()
            n_dev = 0
            for s_, (sp, idx) in enumerate(zip(ob.spaces, ob.glob)):
                sel = np.nonzero(sidx == s_)[0]
                fc, oh = sp.linetest(rays[sel], use_any=use_any[sel], any_mask=any_mask[sel], mode=mode)
                want = {}
                for j in range(len(oh.a)):
                    c0, k = int(oh.first[j]), int(oh.count[j])
                    want[(int(sel[oh.a[j]]), int(idx[oh.b[j]]))] = (bits(oh.mtv[j]).tolist(), bits(oh.contacts[c0:c0 + k]).tolist())
                got = {}
                for r in sel:
                    for h in hits[first[r]:first[r] + count[r]]:
                        k = int(h["count_rank"] & 0xff)
                        c0 = int(h["first_contact"])
                        pts = np.concatenate([con["p"][c0:c0 + k], con["n"][c0:c0 + k]], -1)
                        got[(int(r), int(h["b"]))] = (bits(h["mtv"]).tolist(), bits(pts).tolist())
                n_dev += len(got)
                if mode == "rect":
                    assert got == want
                else:   # DDA reports a subset of brute force with identical values; what is missing are fudge phantoms
                    assert all(want[k] == v for k, v in got.items())
                    loc = {int(g): j for j, g in enumerate(idx)}
                    for (r, b) in set(want) - set(got):
                        s0, e0 = rays[r, :2], rays[r, 2:]
                        vu = (e0 - s0) / np.hypot(*(e0 - s0))
                        st = s0 - 0.01 * vu
                        seg = (st[0], st[1], e0[0], e0[1])
                        assert (not _segment_touches_box(seg, sp.bounds(loc[b]), 0.9)
                                or not _supercover_touches(seg, sp.cell_rect(loc[b]), w.cell_w, w.cell_h)), (r, b)
                    assert len(got) >= 0.9 * len(want)
            assert n_dev > 0
    finally:
        ds.close()


@pytest.mark.parametrize("dda", [False, True])
def test_rays_against_large_polygons_polylines_and_lines(dda):
    """Shapes on both sides of the hit-line mask of the count pass (rsv_lines.cuh: polygons of up to 15 lines carry a
    mask, larger ones a count): 3..40-gons, open polylines of 2..20 points, two-point lines, circles — and rays long
    enough to cut several lines of one shape."""
    rs = np.random.default_rng(23)
    rows = []
    for i in range(500):
        x, y = rs.uniform(0, 600), rs.uniform(0, 600)
        k = i % 4
        if k == 0:
            n = int(rs.integers(3, 41))
            ang = np.sort(rs.uniform(0, 2 * np.pi, n))
            r = rs.uniform(8, 40)
            rows.append(dict(kind=1, pos=(x, y), radius=0.0, verts=np.stack([np.cos(ang) * r, np.sin(ang) * r], -1), tags=1))
        elif k == 1:
            n = int(rs.integers(2, 21))
            pts = np.cumsum(rs.uniform(-12, 12, (n, 2)), 0)
            rows.append(dict(kind=1, pos=(x, y), radius=0.0, closed=0, verts=pts - pts.mean(0), tags=2))
        elif k == 2:
            d = rs.uniform(-30, 30, 2)
            rows.append(dict(kind=1, pos=(x, y), radius=0.0, verts=np.array([-d, d]), tags=1))
        else:
            rows.append(dict(kind=0, pos=(x, y), radius=rs.uniform(3, 25), tags=2))
    w = worlds._from_rows(rows, (600, 600), (32, 32), "raymix").finalize()
    ds = space_for_world(w, max_rays=4000)
    try:
        ds.frame(1, _abi.FRAME_GRID_ONLY)
        sp = OracleBatch(w).spaces[0]
        rays = worlds.make_rays(4000, 31, w.space_w, w.space_h, min_len=8.0, max_len=300.0)
        c = ds.linetest(rays, flags=_abi.FRAME_DOWNLOAD | (_abi.LINE_DDA if dda else 0))
        hits, con, first, count = ds.ray_results()
        fc, oh = sp.linetest(rays, mode="rect" if not dda else "all")
        h, n, dc = _canon_dev(hits, con)
        o, oc = _canon_ora(oh)
        if not dda:
            assert c.n_candidates == fc.candidates and c.n_hits == fc.hits and c.n_contacts == fc.contacts
            sel = np.arange(len(o))
        else:
            # every reported set is one of the brute-force sets (missing ones are phantoms of the +1 fudge: the DDA test above)
            key_o = oh.a[o].astype(np.int64) * w.n + oh.b[o].astype(np.int64)
            key_d = h["ray"].astype(np.int64) * w.n + h["b"].astype(np.int64)
            sel = np.searchsorted(key_o, key_d)
            assert np.all(sel < len(key_o)) and np.array_equal(key_o[sel], key_d)
            assert len(key_d) > 0.9 * len(key_o)
        assert np.array_equal(h["ray"], oh.a[o][sel]) and np.array_equal(h["b"], oh.b[o][sel])
        assert np.array_equal(n, oh.count[o][sel])
        starts = (np.cumsum(oh.count[o]) - oh.count[o])[sel]
        oc_sel = np.concatenate([oc[s:s + k] for s, k in zip(starts, oh.count[o][sel])]) if len(sel) else np.zeros((0, 4))
        for name, dev, ora in (("contacts", dc, oc_sel), ("mtv", h["mtv"], oh.mtv[o][sel]), ("center", h["center"], oh.center[o][sel]),
                               ("key", h["key"], oh.dist[o][sel])):
            assert np.array_equal(bits(dev), bits(ora)), name
        assert int((n > 2).sum()) > 20, "no ray cut more than two lines of one shape"
        big = w.nv[h["b"]] > 16
        assert int(big.sum()) > 50, "no hits on polygons beyond the mask width"
    finally:
        ds.close()
