"""ctypes binding for the CPU oracle (oracle/resolv_oracle.cpp).

TEST INFRASTRUCTURE ONLY — PARITY UNPINNED (see the header of resolv_oracle.cpp).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; nothing under resolv_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc only, no GPU needed)."""
    src = os.path.join(_HERE, "resolv_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class FrameCounts(C.Structure):
    _fields_ = [
        ("testers", C.c_longlong), ("candidates", C.c_longlong), ("hits", C.c_longlong), ("contacts", C.c_longlong),
        ("mtv_sum_x", C.c_double), ("mtv_sum_y", C.c_double), ("seconds", C.c_double),
    ]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    vp, i32, u64, f64 = C.c_void_p, C.c_int, C.c_uint64, C.c_double
    P = C.POINTER
    L.orc_space_new.restype = vp
    L.orc_space_new.argtypes = [i32, i32, i32, i32]
    L.orc_space_free.argtypes = [vp]
    for n in ("orc_width_cells", "orc_height_cells", "orc_num_shapes"):
        getattr(L, n).restype = i32
        getattr(L, n).argtypes = [vp]
    L.orc_add_circle.restype = i32
    L.orc_add_circle.argtypes = [vp, f64, f64, f64, u64]
    L.orc_add_polygon.restype = i32
    L.orc_add_polygon.argtypes = [vp, f64, f64, P(f64), i32, i32, u64]
    L.orc_add_rectangle.restype = i32
    L.orc_add_rectangle.argtypes = [vp, f64, f64, f64, f64, u64]
    L.orc_add_rectangle_top_left.restype = i32
    L.orc_add_rectangle_top_left.argtypes = [vp, f64, f64, f64, f64, u64]
    L.orc_add_line.restype = i32
    L.orc_add_line.argtypes = [vp, f64, f64, f64, f64, u64]
    L.orc_add_bulk.argtypes = [vp, i32, P(C.c_uint8), P(f64), P(f64), P(f64), P(C.c_uint32), P(C.c_uint32), P(f64),
                               P(C.c_uint8), P(u64)]
    L.orc_set_position.argtypes = [vp, i32, f64, f64]
    L.orc_move.argtypes = [vp, i32, f64, f64]
    L.orc_set_tags.argtypes = [vp, i32, u64]
    L.orc_set_positions.argtypes = [vp, P(f64)]
    L.orc_get_position.argtypes = [vp, i32, P(f64)]
    L.orc_get_local_bounds.argtypes = [vp, i32, P(f64)]
    L.orc_get_bounds.argtypes = [vp, i32, P(f64)]
    L.orc_get_cell_rect.argtypes = [vp, i32, P(C.c_longlong)]
    L.orc_cells_dump.restype = C.c_longlong
    L.orc_cells_dump.argtypes = [vp, P(C.c_uint32), P(C.c_uint32), C.c_longlong]
    L.orc_intersection.restype = i32
    L.orc_intersection.argtypes = [vp, i32, i32, P(f64), P(f64), i32]
    L.orc_frame.argtypes = [vp, i32, P(C.c_uint8), P(C.c_uint8), P(u64), P(u64), i32, i32, P(FrameCounts)]
    L.orc_linetest.argtypes = [vp, P(f64), i32, P(C.c_uint8), P(u64), P(u64), i32, i32, P(FrameCounts)]
    L.orc_linetest_ex.argtypes = [vp, P(f64), i32, P(C.c_uint8), P(u64), P(u64), i32, i32, P(FrameCounts), i32,
                                  P(C.c_uint32), P(u64)]
    L.orc_shape_linetest.restype = i32
    L.orc_shape_linetest.argtypes = [vp, i32, f64, f64, f64, f64, i32, i32, i32, u64, u64]
    for n in ("orc_num_candidates", "orc_num_hits", "orc_num_contacts"):
        getattr(L, n).restype = C.c_longlong
        getattr(L, n).argtypes = [vp]
    for n in ("orc_cand_a", "orc_cand_b", "orc_cand_rank", "orc_hit_a", "orc_hit_b", "orc_hit_first", "orc_hit_count",
              "orc_hit_order"):
        getattr(L, n).restype = P(C.c_uint32)
        getattr(L, n).argtypes = [vp]
    for n in ("orc_cand_dist", "orc_hit_mtv", "orc_hit_dist", "orc_contacts", "orc_hit_center"):
        getattr(L, n).restype = P(f64)
        getattr(L, n).argtypes = [vp]
    L.orc_go_sort_f64.argtypes = [P(f64), i32, P(i32)]
    L.orc_bouncer_run.argtypes = [vp, P(i32), i32, P(f64), i32, i32, P(FrameCounts)]
    L.orc_hardware_threads.restype = i32
    _lib = L
    return L


def _ptr(a, ty):
    return None if a is None else a.ctypes.data_as(C.POINTER(ty))


def _arr(p, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(p, shape=(n,)).astype(dtype, copy=True)


@dataclass
class HitResult:
    """Canonical dump of a frame / line test: one row per non-empty IntersectionSet, in callback order."""
    a: np.ndarray        # tester shape index (or ray index)
    b: np.ndarray        # OtherShape index
    order: np.ndarray    # callback index within the tester
    first: np.ndarray    # first contact row
    count: np.ndarray    # number of contacts
    mtv: np.ndarray      # (H,2)
    dist: np.ndarray     # sort key the reference ordered this set by
    contacts: np.ndarray  # (K,4) = point.x, point.y, normal.x, normal.y
    center: np.ndarray | None = None


class OracleSpace:
    """resolv.Space restated on the CPU. Shapes are addressed by insertion index."""

    def __init__(self, space_w: int, space_h: int, cell_w: int, cell_h: int):
        self.L = lib()
        self.h = self.L.orc_space_new(space_w, space_h, cell_w, cell_h)
        self.cell_w, self.cell_h = cell_w, cell_h

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_space_free(self.h)
            self.h = None

    # --- construction (mirrors NewCircle / NewConvexPolygon / NewRectangle* / NewLine + Space.Add)
    def add_circle(self, x, y, r, tags=0):
        return self.L.orc_add_circle(self.h, x, y, r, tags)

    def add_polygon(self, x, y, pts, closed=True, tags=0):
        pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1)
        return self.L.orc_add_polygon(self.h, x, y, _ptr(pts, C.c_double), len(pts) // 2, int(closed), tags)

    def add_rectangle(self, x, y, w, h, tags=0):
        return self.L.orc_add_rectangle(self.h, x, y, w, h, tags)

    def add_rectangle_top_left(self, x, y, w, h, tags=0):
        return self.L.orc_add_rectangle_top_left(self.h, x, y, w, h, tags)

    def add_line(self, x1, y1, x2, y2, tags=0):
        return self.L.orc_add_line(self.h, x1, y1, x2, y2, tags)

    def add_bulk(self, kind, pos0, pos, radius, vert_off, nv, verts, closed, tags):
        """Adds len(kind) shapes in index order (see orc_add_bulk)."""
        c = np.ascontiguousarray
        kind, closed = c(kind, dtype=np.uint8), c(closed, dtype=np.uint8)
        pos0, pos, radius, verts = (c(a, dtype=np.float64) for a in (pos0, pos, radius, verts))
        vert_off, nv, tags = c(vert_off, dtype=np.uint32), c(nv, dtype=np.uint32), c(tags, dtype=np.uint64)
        if verts.size == 0:
            verts = np.zeros(2)
        self.L.orc_add_bulk(self.h, len(kind), _ptr(kind, C.c_uint8), _ptr(pos0, C.c_double), _ptr(pos, C.c_double),
                            _ptr(radius, C.c_double), _ptr(vert_off, C.c_uint32), _ptr(nv, C.c_uint32),
                            _ptr(verts, C.c_double), _ptr(closed, C.c_uint8), _ptr(tags, C.c_uint64))

    def set_position(self, i, x, y):
        self.L.orc_set_position(self.h, i, x, y)

    def move(self, i, dx, dy):
        self.L.orc_move(self.h, i, dx, dy)

    def set_positions(self, xy):
        xy = np.ascontiguousarray(xy, dtype=np.float64)
        assert xy.size == 2 * self.num_shapes
        self.L.orc_set_positions(self.h, _ptr(xy, C.c_double))

    # --- inspection
    @property
    def num_shapes(self):
        return self.L.orc_num_shapes(self.h)

    @property
    def width_cells(self):
        return self.L.orc_width_cells(self.h)

    @property
    def height_cells(self):
        return self.L.orc_height_cells(self.h)

    def local_bounds(self, i):
        out = np.zeros(4)
        self.L.orc_get_local_bounds(self.h, i, _ptr(out, C.c_double))
        return out

    def bounds(self, i):
        out = np.zeros(4)
        self.L.orc_get_bounds(self.h, i, _ptr(out, C.c_double))
        return out

    def position(self, i):
        out = np.zeros(2)
        self.L.orc_get_position(self.h, i, _ptr(out, C.c_double))
        return out

    def cell_rect(self, i):
        out = np.zeros(4, dtype=np.int64)
        self.L.orc_get_cell_rect(self.h, i, _ptr(out, C.c_longlong))
        return out

    def cells(self):
        """(counts[H*W], entries[R]) — entries in the reference's in-cell order, cells row-major."""
        n = self.width_cells * self.height_cells
        counts = np.zeros(n, dtype=np.uint32)
        R = self.L.orc_cells_dump(self.h, _ptr(counts, C.c_uint32), None, 0)
        entries = np.zeros(max(R, 1), dtype=np.uint32)
        self.L.orc_cells_dump(self.h, _ptr(counts, C.c_uint32), _ptr(entries, C.c_uint32), R)
        return counts, entries[:R]

    def intersection(self, a, b, cap=64):
        """A.Intersection(B) -> (mtv[2], contacts[n,4])."""
        mtv = np.zeros(2)
        con = np.zeros(4 * cap)
        n = self.L.orc_intersection(self.h, a, b, _ptr(mtv, C.c_double), _ptr(con, C.c_double), cap)
        return mtv, con.reshape(-1, 4)[:n].copy()

    # --- the two-phase frame
    def _hits(self, with_center=False):
        L, h = self.L, self.h
        H, K = L.orc_num_hits(h), L.orc_num_contacts(h)
        res = HitResult(
            a=_arr(L.orc_hit_a(h), H, np.uint32), b=_arr(L.orc_hit_b(h), H, np.uint32),
            order=_arr(L.orc_hit_order(h), H, np.uint32), first=_arr(L.orc_hit_first(h), H, np.uint32),
            count=_arr(L.orc_hit_count(h), H, np.uint32), mtv=_arr(L.orc_hit_mtv(h), 2 * H, np.float64).reshape(-1, 2),
            dist=_arr(L.orc_hit_dist(h), H, np.float64), contacts=_arr(L.orc_contacts(h), 4 * K, np.float64).reshape(-1, 4))
        if with_center:
            res.center = _arr(L.orc_hit_center(h), 2 * H, np.float64).reshape(-1, 2)
        return res

    def frame(self, margin=1, tester=None, use_any=None, any_mask=None, none_mask=None, record=True, threads=1):
        """Every tester runs IntersectionTest against SelectTouchingCells(margin).FilterShapes()[.ByTags][.NotByTags]
        with OnIntersect returning true. Returns (counts, candidates, hits)."""
        n = self.num_shapes
        tester = None if tester is None else np.ascontiguousarray(tester, dtype=np.uint8)
        use_any = None if use_any is None else np.ascontiguousarray(use_any, dtype=np.uint8)
        any_mask = np.zeros(n, dtype=np.uint64) if any_mask is None else np.ascontiguousarray(any_mask, dtype=np.uint64)
        none_mask = None if none_mask is None else np.ascontiguousarray(none_mask, dtype=np.uint64)
        fc = FrameCounts()
        self.L.orc_frame(self.h, margin, _ptr(tester, C.c_uint8), _ptr(use_any, C.c_uint8), _ptr(any_mask, C.c_uint64),
                         _ptr(none_mask, C.c_uint64), int(record), threads, C.byref(fc))
        if not record:
            return fc, None, None
        L, h = self.L, self.h
        P = L.orc_num_candidates(h)
        cands = dict(a=_arr(L.orc_cand_a(h), P, np.uint32), b=_arr(L.orc_cand_b(h), P, np.uint32),
                     rank=_arr(L.orc_cand_rank(h), P, np.uint32), dist=_arr(L.orc_cand_dist(h), P, np.float64))
        return fc, cands, self._hits()

    def bouncer_run(self, dyn, vel, frames, sequential=True):
        """Config 1: the Bouncer demo's update loop (examples/worldBouncer.go:139-174) for `frames` frames over the shapes
        `dyn` with velocities `vel` (updated in place). sequential=False: the two-phase / frozen form."""
        dyn = np.ascontiguousarray(dyn, dtype=np.int32)
        assert vel.dtype == np.float64 and vel.flags.c_contiguous and vel.size == 2 * len(dyn)
        fc = FrameCounts()
        self.L.orc_bouncer_run(self.h, _ptr(dyn, C.c_int), len(dyn), _ptr(vel, C.c_double), frames, int(sequential), C.byref(fc))
        return fc

    def linetest(self, rays, use_any=None, any_mask=None, none_mask=None, record=True, threads=1, mode="all",
                 candidates=None):
        """Batched LineTest. mode "all": TestAgainst = space.FilterShapes() (brute force, the reference has no line
        broadphase); "rect": space.FilterCells(bounds of the cast line).FilterShapes(); "list": an explicit
        ShapeCollection per ray (candidates = list of index arrays, iterated in the given order)."""
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 4)
        n = len(rays)
        use_any = None if use_any is None else np.ascontiguousarray(use_any, dtype=np.uint8)
        any_mask = np.zeros(n, dtype=np.uint64) if any_mask is None else np.ascontiguousarray(any_mask, dtype=np.uint64)
        none_mask = None if none_mask is None else np.ascontiguousarray(none_mask, dtype=np.uint64)
        m = {"all": 0, "rect": 1, "list": 2}[mode]
        cand = off = None
        if m == 2:
            off = np.zeros(n + 1, dtype=np.uint64)
            off[1:] = np.cumsum([len(c) for c in candidates])
            cand = np.ascontiguousarray(np.concatenate([np.asarray(c, dtype=np.uint32) for c in candidates] + [np.zeros(0, np.uint32)]),
                                        dtype=np.uint32)
            if cand.size == 0:
                cand = np.zeros(1, dtype=np.uint32)
        fc = FrameCounts()
        self.L.orc_linetest_ex(self.h, _ptr(rays, C.c_double), n, _ptr(use_any, C.c_uint8), _ptr(any_mask, C.c_uint64),
                               _ptr(none_mask, C.c_uint64), int(record), threads, C.byref(fc), m, _ptr(cand, C.c_uint32),
                               _ptr(off, C.c_uint64))
        return fc, (self._hits(with_center=True) if record else None)

    def shape_linetest(self, shape, vec, start_offset=(0.0, 0.0), include_all=False, margin=1, use_any=False,
                       any_mask=0, none_mask=0):
        self.L.orc_shape_linetest(self.h, shape, vec[0], vec[1], start_offset[0], start_offset[1], int(include_all),
                                  margin, int(use_any), any_mask, none_mask)
        return self._hits(with_center=True)


def go_sort(keys):
    """sort.Slice restatement: returns the permutation Go's pdqsort_func would produce for ascending keys."""
    keys = np.ascontiguousarray(keys, dtype=np.float64)
    perm = np.zeros(len(keys), dtype=np.int32)
    lib().orc_go_sort_f64(_ptr(keys, C.c_double), len(keys), _ptr(perm, C.c_int))
    return perm


def hardware_threads():
    return lib().orc_hardware_threads()
