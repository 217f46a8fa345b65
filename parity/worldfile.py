"""Text world / dump formats shared by the Go harness (parity/go/dump) and the CPU oracle.

Purpose: pin the oracle (and through it the CUDA path) against the REAL reference. This image has no Go toolchain, so
the pieces are shipped ready to run elsewhere:

    python parity/worldfile.py export            # writes parity/worlds/*.world and parity/expected/*.dump (oracle)
    cd parity/go && go run ./dump ../worlds/X.world > ../dumps/X.dump      # on a machine with Go >= 1.20
    diff parity/expected/X.dump parity/dumps/X.dump                        # or: pytest tests/test_go_dump.py

Every float is written as the 16-hex-digit IEEE-754 bit pattern, so the comparison is bit-exact and plain `diff` works.

World file
    RESOLVWORLD 1
    name <id>
    space <space_w> <space_h> <cell_w> <cell_h>
    margin <m>
    shapes <N>
    S <C|P> <tester> <closed> <useAny> <tags> <anyMask> <noneMask> <pos0x> <pos0y> <posx> <posy> <radius> <nv> [<vx> <vy>]*nv
    frames <F>                       # F-1 blocks follow, one per frame >= 1
    pos <f>                          # then N lines "<x> <y>": SetPosition on every shape in index order
    rays <M>
    R <x1> <y1> <x2> <y2>

Dump file (what the reference produces for that world; the oracle writes the same text)
    RESOLVDUMP 1 <id>
    frame <f>
    cell <cy> <cx> <idx>...                          non-empty cells, row-major, in stored order
    test <A> <n> <B>...                              candidates of tester A in generation order (filters applied)
    hit <A> <k> <B> <mtvx> <mtvy> <n> [<px> <py> <nx> <ny>]*n     k-th OnIntersect callback of A's IntersectionTest
    ray <r> <mode>                                   LineTest r with TestAgainst = all | rect (see oracle.linetest)
    rhit <r> <k> <B> <mtvx> <mtvy> <cx> <cy> <n> [<px> <py> <nx> <ny>]*n
"""
from __future__ import annotations

import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def fhex(x: float) -> str:
    return "%016x" % struct.unpack("<Q", struct.pack("<d", float(x)))[0]


def unhex(s: str) -> float:
    return struct.unpack("<d", struct.pack("<Q", int(s, 16)))[0]


def write_world(path, w, frame_pos, rays):
    """w: worlds.World (single Space); frame_pos: list of f64[N,2] for frames 1..; rays: f64[M,4]."""
    assert w.n_spaces == 1
    out = ["RESOLVWORLD 1", f"name {w.name}", f"space {w.space_w} {w.space_h} {w.cell_w} {w.cell_h}", f"margin {w.margin}",
           f"shapes {w.n}"]
    for i in range(w.n):
        nv = int(w.nv[i])
        v = w.verts[int(w.vert_off[i]):int(w.vert_off[i]) + nv].reshape(-1)
        f = ["S", "CP"[int(w.kind[i])], str(int(w.tester[i])), str(int(w.closed[i])), str(int(w.use_any[i])),
             "%x" % int(w.tags[i]), "%x" % int(w.any_mask[i]), "%x" % int(w.none_mask[i]),
             fhex(w.pos0[i, 0]), fhex(w.pos0[i, 1]), fhex(w.pos[i, 0]), fhex(w.pos[i, 1]), fhex(w.radius[i]), str(nv)]
        f += [fhex(x) for x in v]
        out.append(" ".join(f))
    out.append(f"frames {len(frame_pos) + 1}")
    for f, p in enumerate(frame_pos, start=1):
        out.append(f"pos {f}")
        out += [f"{fhex(x)} {fhex(y)}" for x, y in p]
    out.append(f"rays {len(rays)}")
    out += ["R " + " ".join(fhex(x) for x in r) for r in rays]
    with open(path, "w") as fh:
        fh.write("\n".join(out) + "\n")


def read_world(path):
    """Returns (worlds.World, frame_pos list, rays array)."""
    from resolv_b200 import worlds
    with open(path) as fh:
        lines = [l.split() for l in fh.read().splitlines() if l and not l.startswith("#")]
    assert lines[0] == ["RESOLVWORLD", "1"], "not a world file"
    it = iter(lines[1:])
    name = next(it)[1]
    sw, sh, cw, ch = map(int, next(it)[1:])
    margin = int(next(it)[1])
    n = int(next(it)[1])
    kind = np.zeros(n, np.uint8); tester = np.zeros(n, np.uint8); closed = np.zeros(n, np.uint8); use_any = np.zeros(n, np.uint8)
    tags = np.zeros(n, np.uint64); anym = np.zeros(n, np.uint64); nonem = np.zeros(n, np.uint64)
    pos0 = np.zeros((n, 2)); pos = np.zeros((n, 2)); radius = np.zeros(n); nv = np.zeros(n, np.uint32)
    verts = []
    for i in range(n):
        t = next(it)
        assert t[0] == "S"
        kind[i] = "CP".index(t[1]); tester[i] = int(t[2]); closed[i] = int(t[3]); use_any[i] = int(t[4])
        tags[i] = int(t[5], 16); anym[i] = int(t[6], 16); nonem[i] = int(t[7], 16)
        pos0[i] = unhex(t[8]), unhex(t[9]); pos[i] = unhex(t[10]), unhex(t[11]); radius[i] = unhex(t[12]); nv[i] = int(t[13])
        verts += [unhex(x) for x in t[14:14 + 2 * int(t[13])]]
    frames = int(next(it)[1])
    frame_pos = []
    for f in range(1, frames):
        assert next(it) == ["pos", str(f)]
        frame_pos.append(np.array([[unhex(a), unhex(b)] for a, b in (next(it) for _ in range(n))]).reshape(n, 2))
    m = int(next(it)[1])
    rays = np.array([[unhex(x) for x in next(it)[1:5]] for _ in range(m)]).reshape(m, 4)
    w = worlds.World(space_w=sw, space_h=sh, cell_w=cw, cell_h=ch, kind=kind, pos=pos, pos0=pos0, radius=radius, tags=tags,
                     nv=nv, vert_off=(np.cumsum(nv) - nv).astype(np.uint32), verts=np.array(verts).reshape(-1, 2), closed=closed,
                     tester=tester, use_any=use_any, any_mask=anym, none_mask=nonem, space_idx=np.zeros(n, np.uint32),
                     margin=margin, name=name).finalize()
    return w, frame_pos, rays


def oracle_dump(w, frame_pos, rays) -> str:
    """The dump text the CPU oracle produces for a world (same text the Go harness prints for the reference)."""
    from tests.util import OracleBatch
    ob = OracleBatch(w)
    sp = ob.spaces[0]
    W = sp.width_cells
    out = [f"RESOLVDUMP 1 {w.name}"]
    for f in range(len(frame_pos) + 1):
        if f:
            ob.set_positions(frame_pos[f - 1])
        out.append(f"frame {f}")
        counts, entries = sp.cells()
        e = 0
        for c, k in enumerate(counts):
            if k:
                out.append(f"cell {c // W} {c % W} " + " ".join(str(int(x)) for x in entries[e:e + k]))
            e += k
        tot, cands, hits = ob.frame(w.margin)
        # candidates arrive grouped by tester in generation order; hits grouped by tester in callback order
        ca, cb = cands["a"], cands["b"]
        cstart = {int(a): i for i, a in reversed(list(enumerate(ca)))}
        ccount = np.bincount(ca, minlength=w.n) if len(ca) else np.zeros(w.n, np.int64)
        hstart = {int(a): i for i, a in reversed(list(enumerate(hits["a"])))}
        hcount = np.bincount(hits["a"], minlength=w.n) if len(hits["a"]) else np.zeros(w.n, np.int64)
        for a in range(w.n):
            if not w.tester[a]:
                continue
            k = int(ccount[a])
            s = cstart.get(a, 0)
            assert k == 0 or np.all(cands["rank"][s:s + k] == np.arange(k))
            out.append(f"test {a} {k}" + "".join(f" {int(b)}" for b in cb[s:s + k]))
            s, k = hstart.get(a, 0), int(hcount[a])
            for j in range(s, s + k):
                n, first = int(hits["count"][j]), int(hits["first"][j])
                con = hits["contacts"][first:first + n].reshape(-1)
                out.append(f"hit {a} {int(hits['order'][j])} {int(hits['b'][j])} {fhex(hits['mtv'][j, 0])} {fhex(hits['mtv'][j, 1])} {n}"
                           + "".join(" " + fhex(x) for x in con))
    for mode in ("all", "rect"):
        if not len(rays):
            break
        fc, h = sp.linetest(rays, mode=mode)
        start = {int(a): i for i, a in reversed(list(enumerate(h.a)))}
        cnt = np.bincount(h.a, minlength=len(rays)) if len(h.a) else np.zeros(len(rays), np.int64)
        for r in range(len(rays)):
            out.append(f"ray {r} {mode}")
            s = start.get(r, 0)
            for j in range(s, s + int(cnt[r])):
                n, first = int(h.count[j]), int(h.first[j])
                con = h.contacts[first:first + n].reshape(-1)
                out.append(f"rhit {r} {int(h.order[j])} {int(h.b[j])} {fhex(h.mtv[j, 0])} {fhex(h.mtv[j, 1])} "
                           f"{fhex(h.center[j, 0])} {fhex(h.center[j, 1])} {n}" + "".join(" " + fhex(x) for x in con))
    return "\n".join(out) + "\n"


def cases():
    """The exported parity worlds: small versions of BASELINE configs 1-3 (+ motion), the KAT pairs and edge cases."""
    from resolv_b200 import worlds
    out = []

    def add(w, frames, n_rays, seed):
        pos = []
        for f in range(1, frames):
            worlds.advance(w, f)
            pos.append(w.pos.copy())
        rays = worlds.make_rays(n_rays, seed, w.space_w, w.space_h, 16, min(w.space_w, w.space_h) / 2) if n_rays else np.zeros((0, 4))
        # frame 0 positions: advance() mutated w.pos, so rebuild the frame-0 state
        return pos, rays

    for make, frames, n_rays, seed in (
            (lambda: worlds.make_bouncer(294), 3, 64, 101),
            (lambda: worlds.make_rects(1200, space=2304, cell=32), 3, 64, 102),
            (lambda: worlds.make_mixed(1200, space=2304, cell=32), 3, 64, 103),
            (lambda: worlds.make_mixed(600, space=1024, cell=32, filtered=False), 2, 32, 104),   # crowded
    ):
        w = make()
        p0 = w.pos.copy()
        pos, rays = add(w, frames, n_rays, seed)
        w.pos = p0
        out.append((w, pos, rays))
    # config 4: one platformer Space (tiles + bodies, SelectTouchingCells(4).ByTags(SolidWall)) with the bodies' probe rays
    w = worlds.make_platformer_batch(1)
    w.name = "cfg4_platformer_1"
    p0 = w.pos.copy()
    pos, _ = add(w, 3, 0, 0)
    w.pos = p0
    out.append((w, pos, worlds.make_platformer_probes(w)[0]))
    # shape zoo: 3..40-gons, open polylines, two-point lines, circles, part of them off the grid; long rays
    out.append(_zoo())
    return out


def _zoo():
    from resolv_b200 import worlds
    rs = np.random.default_rng(23)
    rows = []
    for i in range(240):
        x, y = rs.uniform(-40, 640), rs.uniform(-40, 640)
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
    w = worlds._from_rows(rows, (600, 600), (32, 32), "zoo_240").finalize()
    w.vel = rs.uniform(-6, 6, (w.n, 2))
    p0 = w.pos.copy()
    pos = []
    for f in range(1, 3):
        worlds.advance(w, f)
        pos.append(w.pos.copy())
    w.pos = p0
    return w, pos, worlds.make_rays(48, 105, w.space_w, w.space_h, 8.0, 300.0)


def export():
    os.makedirs(os.path.join(HERE, "worlds"), exist_ok=True)
    os.makedirs(os.path.join(HERE, "expected"), exist_ok=True)
    for w, pos, rays in cases():
        wp = os.path.join(HERE, "worlds", w.name + ".world")
        write_world(wp, w, pos, rays)
        w2, pos2, rays2 = read_world(wp)           # the dump is made from what the Go side will read
        txt = oracle_dump(w2, pos2, rays2)
        with open(os.path.join(HERE, "expected", w.name + ".dump"), "w") as fh:
            fh.write(txt)
        print(w.name, "shapes", w.n, "frames", len(pos) + 1, "rays", len(rays), "dump lines", txt.count("\n"))


def export_bench(n: int):
    """A bench.py-sized world (configs[1] at n shapes, the workload's density) for parity/go/bench."""
    from resolv_b200 import worlds
    side = int(round(65536 * (n / 1_000_000) ** 0.5 / 32)) * 32
    w = worlds.make_rects(n, space=side, cell=32)
    w.name = f"bench_cfg2_{n}"
    path = os.path.join(HERE, "worlds", w.name + ".world")
    write_world(path, w, [], np.zeros((0, 4)))
    print(path)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "export":
        export()
    elif len(sys.argv) > 2 and sys.argv[1] == "bench":
        export_bench(int(sys.argv[2]))
    else:
        print(__doc__)
