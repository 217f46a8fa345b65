"""The C++ host layer (resolv_b200/csrc/host/resolv.hpp: Space / Shape / IntersectionTest / LineTest / ShapeLineTest
with callbacks, mirroring the Go API) is driven by tests/host/test_host.cpp on the GPU; its transcript is compared,
bit for bit, with the same scenario played through the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

from oracle.oracle import OracleSpace

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOLID, PLATFORM, PLAYER = 1, 2, 4


def _build_and_run(tmp_path):
    exe = str(tmp_path / "test_host")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I/usr/local/cuda/include", os.path.join(ROOT, "tests/host/test_host.cpp"),
                           "-o", exe, "-L" + os.path.join(ROOT, "resolv_b200"), "-lresolv_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "resolv_b200")])
    out = subprocess.check_output([exe], text=True)
    recs = []
    for ln in out.splitlines():
        f = ln.split()
        if len(f) <= 3:
            recs.append((f[0], [int(x) for x in f[1:]] if f[0] not in ("MVP", "PWP") else [float.fromhex(x) for x in f[1:]]))
        else:
            recs.append((f[0], int(f[1]), int(f[2]), int(f[3]), [float.fromhex(x) for x in f[4:]]))
    return recs


def _world():
    sp = OracleSpace(640, 360, 16, 16)
    sp.add_circle(128, 128, 32, SOLID | PLATFORM)
    sp.add_rectangle_top_left(0, 0, 640, 8, SOLID | PLATFORM)
    sp.add_rectangle_top_left(64, 200, 300, 8, SOLID | PLATFORM)
    sp.add_rectangle_top_left(512, 96, 32, 200, SOLID | PLATFORM)
    sp.add_rectangle_top_left(0, 352, 640, 8, SOLID | PLATFORM)
    sp.add_rectangle_top_left(400, 200, 32, 16, PLATFORM)
    sp.add_polygon(180, 175, [-24, 8, 8, -8, 48, -8, 80, 8], True, PLATFORM | SOLID)
    player = sp.add_rectangle(192, 128, 16, 12, PLAYER)
    ball = sp.add_circle(150, 150, 8, PLAYER)
    return sp, player, ball


def _sets(hits, sel, with_center=False):
    out = []
    for i in np.nonzero(sel)[0][np.argsort(hits.order[sel], kind="stable")]:
        f, k = int(hits.first[i]), int(hits.count[i])
        c = hits.center[i].tolist() if with_center and hits.center is not None else [0.0, 0.0]
        out.append((int(hits.b[i]), k, hits.mtv[i].tolist() + c + hits.contacts[f:f + k].reshape(-1).tolist()))
    return out


def _same(dev, ora, what):
    assert len(dev) == len(ora), (what, len(dev), len(ora))
    for d, o in zip(dev, ora):
        assert d[0] == o[0] and d[1] == o[1], (what, d[:2], o[:2])
        assert np.array_equal(np.array(d[2]).view(np.uint64), np.array(o[2]).view(np.uint64)), (what, d, o)


def test_host_layer_transcript_equals_oracle(tmp_path):
    recs = _build_and_run(tmp_path)
    sp, player, ball = _world()
    n = sp.num_shapes
    only = lambda i: np.eye(n, dtype=np.uint8)[i]
    px = [192, 100, 150, 520, 200, 405, 170]
    py = [128, 196, 140, 150, 346, 204, 172]
    # 1. player.IntersectionTest(SelectTouchingCells(4).ByTags(SolidWall))
    for k in range(7):
        sp.set_position(player, px[k], py[k])
        fc, cands, hits = sp.frame(4, tester=only(player), use_any=np.ones(n, np.uint8), any_mask=np.full(n, SOLID, np.uint64))
        ora = _sets(hits, hits.a == player)
        dev = [(r[2], r[3], r[4]) for r in recs if r[0] == "IT" and r[1] == k]
        _same(dev, ora, f"IT{k}")
        assert [r[1] for r in recs if r[0] == "ITR" and r[1][0] == k][0][1] == int(len(ora) > 0)
        assert [r[1] for r in recs if r[0] == "ITE" and r[1][0] == k][0][1] == int(len(ora) > 0)   # early exit after 1 callback
    assert sum(1 for r in recs if r[0] == "IT") >= 3
    # 2. ball.IntersectionTest(SelectTouchingCells(1).Not(player))
    for k in range(3):
        sp.set_position(ball, 150 + 8 * k, 150 + 3 * k)
        fc, cands, hits = sp.frame(1, tester=only(ball))
        ora = [s for s in _sets(hits, hits.a == ball) if s[0] != player]
        dev = [(r[2], r[3], r[4]) for r in recs if r[0] == "IC" and r[1] == k]
        _same(dev, ora, f"IC{k}")
    # 3. LineTest(Start, End, space.FilterShapes().ByTags(Platform))
    fc, lh = sp.linetest([[20, 30, 600, 330]], use_any=[1], any_mask=[PLATFORM])
    ora = _sets(lh, lh.a == 0, with_center=True)
    dev = [(r[2], r[3], r[4]) for r in recs if r[0] == "LT"]
    _same(dev, ora, "LT")
    assert len(ora) >= 2
    # 4. player.ShapeLineTest(Vector{0,12}, SelectTouchingCells(4).ByTags(Platform))
    sp.set_position(player, 200, 188)
    sh = sp.shape_linetest(player, (0.0, 12.0), margin=4, use_any=True, any_mask=PLATFORM)
    ora = _sets(sh, sh.a == player)
    dev = [(r[2], r[3], r[4][:2] + [0.0, 0.0] + r[4][4:]) for r in recs if r[0] == "SL"]   # merged sets: Center is the first ray's
    _same(dev, ora, "SL")
    assert len(ora) >= 1
    # 5. a callback that moves the tester: the reference loop (shape.go:297-312) replayed with oracle primitives
    sp.set_position(player, 100, 196)
    fc, cands, hits = sp.frame(4, tester=only(player), use_any=np.ones(n, np.uint8), any_mask=np.full(n, SOLID, np.uint64))
    order = np.argsort(cands["dist"], kind="stable")
    ora = []
    for b in cands["b"][order]:
        mtv, con = sp.intersection(player, int(b))
        if len(con):
            ora.append((int(b), len(con), mtv.tolist() + [0.0, 0.0] + con.reshape(-1).tolist()))
            sp.move(player, mtv[0], mtv[1])
    dev = [(r[2], r[3], r[4]) for r in recs if r[0] == "MV"]
    _same(dev, ora, "MV")
    mvp = [r[1] for r in recs if r[0] == "MVP"][0]
    assert np.array_equal(np.array(mvp).view(np.uint64), sp.position(player).view(np.uint64))
    # 6. Remove(shape 2) == the shape is in no cell any more
    sp.set_position(2, 1e7, 1e7)
    sp.set_position(player, 100, 196)
    fc, cands, hits = sp.frame(4, tester=only(player))
    ora = _sets(hits, hits.a == player)
    dev = [(r[2], r[3], r[4]) for r in recs if r[0] == "RM"]
    _same(dev, ora, "RM")
    # the frames above were incremental: shapes that stood still were flagged static by the host layer on its own
    si = [r[1] for r in recs if r[0] == "SI"]
    assert len(si) == 2 and si[0][0] > 0 and si[0][1] < n, si


def _replay(sp, tester, cand, move):
    """The reference loop (shape.go:291-312) with oracle primitives: candidates `cand` (generation order) sorted by
    distance with the oracle's restatement of sort.Slice, Intersection() evaluated live, optional MoveVec(MTV)."""
    from oracle.oracle import go_sort
    p = sp.position(tester)
    d = np.array([float((sp.position(int(b))[0] - p[0]) ** 2 + (sp.position(int(b))[1] - p[1]) ** 2) for b in cand])
    out = []
    for k in go_sort(d):
        b = int(cand[k])
        mtv, con = sp.intersection(tester, b)
        if len(con):
            out.append((b, len(con), mtv.tolist() + [0.0, 0.0] + con.reshape(-1).tolist()))
            if move:
                sp.move(tester, mtv[0], mtv[1])
    return out


def test_host_layer_moves_whole_space_filters_and_go_sort(tmp_path):
    recs = _build_and_run(tmp_path)
    sp, player, ball = _world()
    sp.set_position(2, 1e7, 1e7)                      # (scenario 6 removed shape 2)
    block = sp.add_rectangle(600, 66, 16, 8, SOLID)
    wall = sp.add_rectangle(606, 46, 8, 11, SOLID)
    n = sp.num_shapes
    only = lambda i: np.eye(sp.num_shapes, dtype=np.uint8)[i]
    # 7. pushed into the wall by the first MTV: the wall is reported in the same call
    sp.set_position(player, 600, 59)
    fc, cands, hits = sp.frame(2, tester=only(player), use_any=np.ones(n, np.uint8), any_mask=np.full(n, SOLID, np.uint64))
    assert not any(int(b) == wall for b in hits.b), "the wall must not touch the player before the move"
    ora = _replay(sp, player, cands["b"], move=True)
    assert [o[0] for o in ora] == [block, wall]
    _same([(r[2], r[3], r[4]) for r in recs if r[0] == "PW"], ora, "PW")
    pwp = [r[1] for r in recs if r[0] == "PWP"][0]
    assert np.array_equal(np.array(pwp).view(np.uint64), sp.position(player).view(np.uint64))
    # 8. whole-Space iterator (Add order), then ByDistance
    sp.set_position(player, 100, 196)
    solid = [i for i in range(sp.num_shapes) if i != player and i in (0, 1, 2, 3, 4, 6, block, wall)]
    _same([(r[2], r[3], r[4]) for r in recs if r[0] == "WS"], _replay(sp, player, solid, move=False), "WS")
    ring = []
    for i in range(sp.num_shapes):
        q = sp.position(i)
        dd = float(np.sqrt((q[0] - 100.0) ** 2 + (q[1] - 196.0) ** 2))
        if i != player and 50 < dd < 200:
            ring.append(i)
    _same([(r[2], r[3], r[4]) for r in recs if r[0] == "BD"], _replay(sp, player, ring, move=False), "BD")
    # 9. 24 candidates in two rings of exactly equal distances: callback order == the oracle's (pdqsort beyond 12 elements)
    hub = sp.add_circle(300, 100, 5.5, PLAYER)
    off = [(3, 4), (-3, 4), (3, -4), (-3, -4), (4, 3), (-4, 3), (4, -3), (-4, -3), (5, 0), (-5, 0), (0, 5), (0, -5),
           (5, 5), (-5, 5), (5, -5), (-5, -5), (1, 7), (-1, 7), (1, -7), (-1, -7), (7, 1), (-7, 1), (7, -1), (-7, -1)]
    for ox, oy in off:
        sp.add_circle(300 + ox, 100 + oy, 2, PLATFORM)
    n = sp.num_shapes
    fc, cands, hits = sp.frame(1, tester=only(hub), use_any=np.ones(n, np.uint8), any_mask=np.full(n, PLATFORM, np.uint64))
    assert fc.candidates == 24 and fc.hits == 24
    ora = _sets(hits, hits.a == hub)
    dev = [(r[2], r[3], r[4]) for r in recs if r[0] == "PQ"]
    _same(dev, ora, "PQ")
    assert [d[0] for d in dev] != sorted(d[0] for d in dev[:12]) + sorted(d[0] for d in dev[12:]), "order should not be the stable one"
    # 10. the Lines quirk: {3} == {0} (only line 0 is looked at), a four-element slice == every line == no Lines
    lq = lambda v: [(r[1] - v * 10000, r[2], r[3], r[4]) for r in recs if r[0] == "LQ" and r[1] // 10000 == v]
    assert lq(0) == lq(1) == []          # line 0 is the top edge: it does not face the cast direction, no rays, no sets
    sp.set_position(player, 416, 190)
    sh = sp.shape_linetest(player, (0.0, 12.0), margin=4, use_any=True, any_mask=PLATFORM)
    ora = _sets(sh, sh.a == player)
    _same([(r[1], r[2], r[3][:2] + [0.0, 0.0] + r[3][4:]) for r in lq(2)], [(o[0], o[1], o[2]) for o in ora], "LQ all lines")
    lqr = {r[1][0]: r[1][1] for r in recs if r[0] == "LQR"}
    assert lqr == {0: 0, 1: 0, 2: 1} and len(ora) >= 1
    # 11. a set of 16 contacts with sixteen equal sort keys: the host layer's port of Go's sort must leave them in the
    # oracle's order (the device delivers sets of more than 12 contacts in generation order)
    octagon = sp.add_polygon(500, 40, [20, 0, 14, 14, 0, 20, -14, 14, -20, 0, -14, -14, 0, -20, 14, -14], True, PLAYER)
    sp.add_circle(500, 40, 19, PLATFORM)
    n = sp.num_shapes
    fc, cands, hits = sp.frame(1, tester=only(octagon), use_any=np.ones(n, np.uint8), any_mask=np.full(n, PLATFORM, np.uint64))
    ora = _sets(hits, hits.a == octagon)
    assert len(ora) == 1 and ora[0][1] == 16
    _same([(r[2], r[3], r[4]) for r in recs if r[0] == "MC"], ora, "MC")
