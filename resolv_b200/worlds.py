"""Deterministic synthetic worlds for the five BASELINE.json configs.

Counter-based SplitMix64 keyed by (seed, stream, index) so that every consumer (oracle, CUDA path, CPU baseline)
sees bit-identical inputs. Worlds are plain SoA numpy arrays in the layout of include/resolv_b200.h; the
host-side quantities the reference computes at shape-edit time (ConvexPolygon.updateBounds,
convexPolygon.go:163-197) are computed here with the reference's float64 operation order.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)
_GOLD = np.uint64(0x9E3779B97F4A7C15)

TAG_SOLID_WALL = 1   # examples/common.go style tags: first NewTag -> 1, second -> 2, ...
TAG_PLATFORM = 2
TAG_PLAYER = 4
TAG_BOUNCER = 8


def _mix(z):
    with np.errstate(over="ignore"):
        z = (z + _GOLD) & _M
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
        return z ^ (z >> np.uint64(31))


def rand_u64(seed: int, stream: int, n: int, start: int = 0) -> np.ndarray:
    """n SplitMix64 outputs for counters start..start+n-1 of (seed, stream)."""
    with np.errstate(over="ignore"):
        key = _mix(np.uint64(seed) * np.uint64(0x100000001B3) + np.uint64(stream))
        idx = np.arange(start, start + n, dtype=np.uint64)
        return _mix(key + idx * _GOLD)


def rand_u01(seed: int, stream: int, n: int, start: int = 0) -> np.ndarray:
    """Uniform doubles in [0,1): (z >> 11) * 2^-53."""
    return (rand_u64(seed, stream, n, start) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


@dataclass
class World:
    """SoA description of one Space (or a batch of identical Spaces) and its shapes."""
    space_w: int
    space_h: int
    cell_w: int
    cell_h: int
    kind: np.ndarray        # u8[N]   0 circle, 1 polygon
    pos: np.ndarray         # f64[N,2] current position
    pos0: np.ndarray        # f64[N,2] position at the last updateBounds (polygons)
    radius: np.ndarray      # f64[N]
    tags: np.ndarray        # u64[N]
    nv: np.ndarray          # u32[N]  vertex count (0 for circles)
    vert_off: np.ndarray    # u32[N]
    verts: np.ndarray       # f64[V,2] scale/rotation applied, position not applied
    closed: np.ndarray      # u8[N]
    tester: np.ndarray      # u8[N]
    use_any: np.ndarray     # u8[N]
    any_mask: np.ndarray    # u64[N]
    none_mask: np.ndarray   # u64[N]
    space_idx: np.ndarray   # u32[N]
    n_spaces: int = 1
    margin: int = 1
    name: str = ""
    vel: np.ndarray | None = None   # optional per-shape velocity for motion
    local_aabb: np.ndarray = field(default=None)  # f64[N,4], filled by finalize()

    @property
    def n(self) -> int:
        return len(self.kind)

    def finalize(self) -> "World":
        self.local_aabb = compute_local_aabb(self)
        return self

    def meta(self) -> np.ndarray:
        """RSV_BUF_META words (include/resolv_b200.h)."""
        m = (self.kind.astype(np.uint32) & 1) | (self.closed.astype(np.uint32) << 1) | (self.tester.astype(np.uint32) << 2) \
            | (self.use_any.astype(np.uint32) << 3) | (self.nv.astype(np.uint32) << 8)
        return m.astype(np.uint32)


def compute_local_aabb(w: World) -> np.ndarray:
    """ConvexPolygon.updateBounds (convexPolygon.go:163-197): AABB of Transformed() at pos0, then Move(-pos0);
    circles: (-r,-r,+r,+r) so that local + position == Circle.Bounds (circle.go:33-39) exactly."""
    out = np.zeros((w.n, 4))
    circ = w.kind == 0
    out[circ, 0] = -w.radius[circ]
    out[circ, 1] = -w.radius[circ]
    out[circ, 2] = w.radius[circ]
    out[circ, 3] = w.radius[circ]
    poly = np.nonzero(w.kind == 1)[0]
    if len(poly):
        nvp = w.nv[poly].astype(np.int64)
        starts = np.cumsum(nvp) - nvp
        owner = np.repeat(poly, nvp)
        vidx = np.repeat(w.vert_off[poly].astype(np.int64), nvp) + (np.arange(int(nvp.sum())) - np.repeat(starts, nvp))
        t = w.verts[vidx] + w.pos0[owner]                       # Transformed(): p + position
        mn = np.minimum.reduceat(t, starts, axis=0)
        mx = np.maximum.reduceat(t, starts, axis=0)
        out[poly, 0:2] = mn + (-w.pos0[poly])                  # Bounds.Move(-pos.X, -pos.Y)
        out[poly, 2:4] = mx + (-w.pos0[poly])
    return out


def _empty(n, space, cell, name, margin=1, n_spaces=1):
    z = lambda dt: np.zeros(n, dtype=dt)
    return World(space_w=space[0], space_h=space[1], cell_w=cell[0], cell_h=cell[1], kind=z(np.uint8),
                 pos=np.zeros((n, 2)), pos0=np.zeros((n, 2)), radius=z(np.float64), tags=z(np.uint64), nv=z(np.uint32),
                 vert_off=z(np.uint32), verts=np.zeros((0, 2)), closed=np.ones(n, np.uint8), tester=np.ones(n, np.uint8),
                 use_any=z(np.uint8), any_mask=z(np.uint64), none_mask=z(np.uint64), space_idx=z(np.uint32),
                 n_spaces=n_spaces, margin=margin, name=name)


def rect_verts(w, h):
    """NewRectangle's point list (convexPolygon.go:616-632): clockwise from top-left, origin at the centre."""
    hw, hh = np.asarray(w) / 2, np.asarray(h) / 2
    return np.stack([np.stack([-hw, -hh], -1), np.stack([hw, -hh], -1), np.stack([hw, hh], -1), np.stack([-hw, hh], -1)], -2)


def make_rects(n: int, seed: int = 20262, space: int = 65536, cell: int = 32, lo: float = 8.0, hi: float = 32.0) -> World:
    """Config 2: n axis-aligned NewRectangle(x,y,w,h), w,h ~ U[lo,hi), x,y ~ U[0,space), tags = 1, no filter."""
    w = _empty(n, (space, space), (cell, cell), f"cfg2_rects_{n}")
    w.kind[:] = 1
    w.pos[:, 0] = rand_u01(seed, 0, n) * space
    w.pos[:, 1] = rand_u01(seed, 1, n) * space
    ww = lo + rand_u01(seed, 2, n) * (hi - lo)
    hh = lo + rand_u01(seed, 3, n) * (hi - lo)
    w.pos0 = w.pos.copy()
    w.nv[:] = 4
    w.vert_off = (np.arange(n, dtype=np.uint32) * 4).astype(np.uint32)
    w.verts = rect_verts(ww, hh).reshape(-1, 2)
    w.tags[:] = 1
    w.vel = np.stack([rand_u01(seed, 4, n) * 32 - 16, rand_u01(seed, 5, n) * 32 - 16], -1)
    return w.finalize()


def make_mixed(n: int, seed: int = 20263, space: int = 65536, cell: int = 32, filtered: bool = True) -> World:
    """Config 3 / 5: 50 % circles r ~ U[4,16); 50 % convex polygons with nv ~ U{3..8} vertices at sorted random
    angles on a circle of radius U[4,16) (clockwise on screen), pre-rotated on the host (rotation = 0);
    tags = 1 << U{0..3}; every tester filters ByTags(random non-empty subset of the 4 bits)."""
    w = _empty(n, (space, space), (cell, cell), f"cfg3_mixed_{n}")
    is_poly = rand_u64(seed, 0, n) & np.uint64(1)
    w.kind = is_poly.astype(np.uint8)
    w.pos[:, 0] = rand_u01(seed, 1, n) * space
    w.pos[:, 1] = rand_u01(seed, 2, n) * space
    w.pos0 = w.pos.copy()
    rad = 4.0 + rand_u01(seed, 3, n) * 12.0
    w.radius = np.where(w.kind == 0, rad, 0.0)
    nv = np.where(w.kind == 1, 3 + (rand_u64(seed, 4, n) % np.uint64(6)).astype(np.uint32), 0).astype(np.uint32)
    w.nv = nv
    w.vert_off = (np.cumsum(nv) - nv).astype(np.uint32)
    V = int(nv.sum())
    owner = np.repeat(np.arange(n), nv)
    ang = rand_u01(seed, 5, V) * (2 * np.pi)
    # sort angles ascending within each polygon: increasing screen angle (y down) == clockwise on screen
    order = np.lexsort((ang, owner))
    ang = ang[order]
    w.verts = np.stack([np.cos(ang) * rad[owner], np.sin(ang) * rad[owner]], -1)
    w.tags = (np.uint64(1) << (rand_u64(seed, 6, n) % np.uint64(4))).astype(np.uint64)
    if filtered:
        w.use_any[:] = 1
        w.any_mask = (np.uint64(1) + rand_u64(seed, 7, n) % np.uint64(15)).astype(np.uint64)
    w.vel = np.stack([rand_u01(seed, 8, n) * 32 - 16, rand_u01(seed, 9, n) * 32 - 16], -1)
    return w.finalize()


CFG5_SPACE = 262144
CFG5_BANDS = 8


def make_cfg5_band(band: int, n_total: int = 16_000_000, seed: int = 20265) -> World:
    """Config 5 is ONE world of n_total mixed shapes (as config 3) in a 262144^2 Space, defined as 8 horizontal bands of
    n_total/8 shapes each (band b covers y in [b, b+1) * 32768) so that any strip partition into 1/2/4/8 ranks sees the
    SAME world: rank g of G holds bands [g*8/G, (g+1)*8/G). Slot order = band order."""
    n = n_total // CFG5_BANDS
    bh = CFG5_SPACE // CFG5_BANDS
    w = make_mixed(n, seed=seed + 101 * band, space=CFG5_SPACE)
    w.pos[:, 1] = rand_u01(seed + 101 * band, 2, n) * bh + band * float(bh)
    w.pos0 = w.pos.copy()
    w.name = f"cfg5_band{band}"
    return w.finalize()


def concat(ws, space_h=None, name="concat") -> World:
    """Concatenate Worlds of one Space (slot order = list order)."""
    w0 = ws[0]
    cat = np.concatenate
    voff, acc = [], 0
    for w in ws:
        voff.append(w.vert_off.astype(np.int64) + acc)
        acc += len(w.verts)
    out = World(space_w=w0.space_w, space_h=space_h or w0.space_h, cell_w=w0.cell_w, cell_h=w0.cell_h,
                kind=cat([w.kind for w in ws]), pos=cat([w.pos for w in ws]), pos0=cat([w.pos0 for w in ws]),
                radius=cat([w.radius for w in ws]), tags=cat([w.tags for w in ws]), nv=cat([w.nv for w in ws]),
                vert_off=cat(voff).astype(np.uint32), verts=cat([w.verts for w in ws]),
                closed=cat([w.closed for w in ws]), tester=cat([w.tester for w in ws]),
                use_any=cat([w.use_any for w in ws]), any_mask=cat([w.any_mask for w in ws]),
                none_mask=cat([w.none_mask for w in ws]), space_idx=cat([w.space_idx for w in ws]),
                n_spaces=1, margin=w0.margin, name=name,
                vel=cat([w.vel for w in ws]) if w0.vel is not None else None)
    out.local_aabb = cat([w.local_aabb for w in ws])
    return out


def make_cfg5(n_total: int = 16_000_000, seed: int = 20265) -> World:
    """The whole config-5 world (single-GPU run / oracle checks at reduced n_total)."""
    return concat([make_cfg5_band(b, n_total, seed) for b in range(CFG5_BANDS)], name=f"cfg5_{n_total}")


def _append_top_left_rects(rows, x, y, ww, hh, tags):
    """NewRectangleFromTopLeft (convexPolygon.go:638-643): built at (x,y), then Move(w/2, h/2)."""
    rows.append(dict(kind=1, pos0=(x, y), pos=(x + ww / 2, y + hh / 2), radius=0.0, verts=rect_verts(ww, hh), tags=tags))


def _from_rows(rows, space, cell, name, margin=1):
    n = len(rows)
    w = _empty(n, space, cell, name, margin)
    verts = []
    off = 0
    for i, r in enumerate(rows):
        w.kind[i] = r["kind"]
        w.pos[i] = r["pos"]
        w.pos0[i] = r.get("pos0", r["pos"])
        w.radius[i] = r["radius"]
        w.tags[i] = r["tags"]
        w.tester[i] = r.get("tester", 1)
        w.closed[i] = r.get("closed", 1)
        if r["kind"] == 1:
            v = np.asarray(r["verts"], dtype=np.float64).reshape(-1, 2)
            w.nv[i] = len(v)
            w.vert_off[i] = off
            off += len(v)
            verts.append(v)
    w.verts = np.concatenate(verts) if verts else np.zeros((0, 2))
    return w


def make_bouncer(n_dynamic: int = 294, seed: int = 20261, space=(640, 480), cell=(16, 16)) -> World:
    """Config 1: the Bouncer demo world headless (examples/worldBouncer.go:44-54,124-135): six wall rectangles
    tagged SolidWall, then n_dynamic bouncers — circles r = 8 and 16x16 rectangles alternating — tagged Bouncer,
    every shape tested with SelectTouchingCells(1).FilterShapes() (no tag filter)."""
    W, H = space
    rows = []
    for (x, y, ww, hh) in ((0, 0, W, 16), (0, H - 16, W, 16), (0, 16, 16, H - 16), (W - 16, 16, 16, H - 16),
                           (64, 128, 16, 200), (120, 300, 200, 8)):
        _append_top_left_rects(rows, float(x), float(y), float(ww), float(hh), TAG_SOLID_WALL)
        rows[-1]["tester"] = 0
    px = 24.0 + rand_u01(seed, 0, n_dynamic) * (W - 48.0)
    py = 24.0 + rand_u01(seed, 1, n_dynamic) * (H - 48.0)
    for i in range(n_dynamic):
        if i % 2 == 0:
            rows.append(dict(kind=0, pos=(px[i], py[i]), radius=8.0, tags=TAG_BOUNCER))
        else:
            rows.append(dict(kind=1, pos=(px[i], py[i]), radius=0.0, verts=rect_verts(16.0, 16.0), tags=TAG_BOUNCER))
    w = _from_rows(rows, space, cell, f"cfg1_bouncer_{n_dynamic}")
    vel = np.zeros((w.n, 2))
    vel[6:, 0] = rand_u01(seed, 2, n_dynamic) * 2 - 1          # examples/worldBouncer.go:128
    w.vel = vel
    return w.finalize()


def make_platformer_batch(n_spaces: int, seed: int = 20264, n_tiles: int = 500, n_bodies: int = 32) -> World:
    """Config 4: n_spaces independent 640x360 Spaces with 16x16 cells (examples/worldPlatformer.go:34). Each has
    ~n_tiles static 16x16 tiles (border + random platforms; SolidWall|Platform or Platform only) and n_bodies
    dynamic 16x12 bodies (examples/worldPlatformer.go:146) that test with SelectTouchingCells(4).FilterShapes()
    .ByTags(SolidWall) (examples/worldPlatformer.go:210,255-258)."""
    W, H, C = 640, 360, 16
    gw, gh = W // C, (H + C - 1) // C  # 40 x 23 (last row is half a tile tall in the reference grid)
    border = [(x, 0) for x in range(gw)] + [(x, gh - 2) for x in range(gw)] + \
             [(0, y) for y in range(1, gh - 2)] + [(gw - 1, y) for y in range(1, gh - 2)]
    n_rand = max(n_tiles - len(border), 0)
    per = len(border) + n_rand + n_bodies
    n = per * n_spaces
    w = _empty(n, (W, H), (C, C), f"cfg4_platformer_{n_spaces}", margin=4, n_spaces=n_spaces)
    w.kind[:] = 1
    w.nv[:] = 4
    w.vert_off = (np.arange(n, dtype=np.uint32) * 4).astype(np.uint32)
    sp = np.repeat(np.arange(n_spaces, dtype=np.uint32), per)
    w.space_idx = sp
    local = np.tile(np.arange(per), n_spaces)
    is_border = local < len(border)
    is_rand = (local >= len(border)) & (local < len(border) + n_rand)
    is_body = local >= len(border) + n_rand
    bx = np.array([b[0] for b in border], dtype=np.float64)
    by = np.array([b[1] for b in border], dtype=np.float64)
    tx = np.zeros(n)
    ty = np.zeros(n)
    tx[is_border] = np.tile(bx, n_spaces) * C
    ty[is_border] = np.tile(by, n_spaces) * C
    k = int(is_rand.sum())
    tx[is_rand] = (1 + (rand_u64(seed, 0, k) % np.uint64(gw - 2)).astype(np.float64)) * C
    ty[is_rand] = (1 + (rand_u64(seed, 1, k) % np.uint64(gh - 3)).astype(np.float64)) * C
    ww = np.where(is_body, 16.0, 16.0)
    hh = np.where(is_body, 12.0, 16.0)
    # tiles: NewRectangleFromTopLeft(tx, ty, 16, 16); bodies: NewRectangle(x, y, 16, 12)
    kb = int(is_body.sum())
    cx = np.where(is_body, 0.0, tx)
    cy = np.where(is_body, 0.0, ty)
    cx[is_body] = 24.0 + rand_u01(seed, 2, kb) * (W - 48.0)
    cy[is_body] = 24.0 + rand_u01(seed, 3, kb) * (H - 64.0)
    w.pos0 = np.stack([cx, cy], -1)
    w.pos = np.where(is_body[:, None], w.pos0, w.pos0 + np.stack([ww / 2, hh / 2], -1))
    w.verts = rect_verts(ww, hh).reshape(-1, 2)
    solid = is_border | (is_rand & ((rand_u64(seed, 4, n) & np.uint64(1)) == 0))
    w.tags = np.where(is_body, np.uint64(TAG_PLAYER),
                      np.where(solid, np.uint64(TAG_SOLID_WALL | TAG_PLATFORM), np.uint64(TAG_PLATFORM))).astype(np.uint64)
    w.tester = is_body.astype(np.uint8)
    w.use_any = is_body.astype(np.uint8)
    w.any_mask = np.where(is_body, np.uint64(TAG_SOLID_WALL), np.uint64(0)).astype(np.uint64)
    vel = np.zeros((n, 2))
    vel[is_body, 0] = rand_u01(seed, 5, kb) * 8 - 4
    vel[is_body, 1] = rand_u01(seed, 6, kb) * 8 - 4
    w.vel = vel
    return w.finalize()


def make_platformer_probes(w: World, reach: float = 16.0):
    """Config 4's LineTest probes (examples/worldPlatformer.go:221-246,291-296 in spirit): per dynamic body two rays
    straight down from its two bottom vertices and two rays along its facing from the leading edge's two vertices,
    each `reach` long, against ByTags(Platform). Returns (rays f64[4B,4], use_any u8, any_mask u64, space_idx u32)."""
    body = np.nonzero(w.tester != 0)[0]
    p = w.pos[body]
    hw, hh = 8.0, 6.0
    face = np.where((w.vel[body, 0] if w.vel is not None else np.ones(len(body))) < 0, -1.0, 1.0)
    rays = np.zeros((len(body), 4, 4))
    for k, sx in enumerate((-hw, hw)):                      # down from the bottom vertices
        rays[:, k, 0] = p[:, 0] + sx; rays[:, k, 1] = p[:, 1] + hh
        rays[:, k, 2] = p[:, 0] + sx; rays[:, k, 3] = p[:, 1] + hh + reach
    for k, sy in enumerate((-hh, hh)):                      # along the facing from the leading edge
        rays[:, 2 + k, 0] = p[:, 0] + face * hw; rays[:, 2 + k, 1] = p[:, 1] + sy
        rays[:, 2 + k, 2] = p[:, 0] + face * (hw + reach); rays[:, 2 + k, 3] = p[:, 1] + sy
    n = 4 * len(body)
    return (rays.reshape(n, 4), np.ones(n, np.uint8), np.full(n, TAG_PLATFORM, np.uint64),
            np.repeat(w.space_idx[body], 4).astype(np.uint32))


def advance(w: World, frame: int) -> None:
    """Per-frame motion of SURVEY.md §8d: pos += vel, reflected at the Space borders (in place)."""
    if w.vel is None:
        return
    p = w.pos + w.vel
    for ax, lim in ((0, float(w.space_w)), (1, float(w.space_h))):
        lo = p[:, ax] < 0
        hi = p[:, ax] >= lim
        p[lo, ax] = -p[lo, ax]
        p[hi, ax] = 2 * lim - p[hi, ax] - 1e-9
        w.vel[lo | hi, ax] *= -1
    w.pos = p


def make_rays(n: int, seed: int, space_w: float, space_h: float, min_len: float = 16.0, max_len: float = 512.0) -> np.ndarray:
    """Config 5 ray sweep: start uniform in the Space, direction uniform, length U[min_len,max_len). (n,4)."""
    sx = rand_u01(seed, 100, n) * space_w
    sy = rand_u01(seed, 101, n) * space_h
    th = rand_u01(seed, 102, n) * (2 * np.pi)
    ln = min_len + rand_u01(seed, 103, n) * (max_len - min_len)
    return np.stack([sx, sy, sx + np.cos(th) * ln, sy + np.sin(th) * ln], -1)
