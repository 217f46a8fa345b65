"""ctypes binding of libresolv_b200.so (include/resolv_b200.h) — the same C ABI the Go/cgo host layer binds.

There is no CPU fallback: importing works anywhere (so the symbol table can be checked without a GPU), but
DeviceSpace() raises unless the CUDA library is built and a device is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RSV_LIB_PATH") or os.path.join(_HERE, "libresolv_b200.so")   # (RSV_LIB_PATH: tuning experiments)

RSV_OK, RSV_ERR_BAD_ARG, RSV_ERR_OOM, RSV_ERR_CUDA, RSV_ERR_CAPACITY, RSV_ERR_NO_DEVICE = 0, -1, -2, -3, -4, -5
(BUF_POS, BUF_LOCAL_AABB, BUF_RADIUS, BUF_TAGS, BUF_META, BUF_VERT_OFF, BUF_VERTS, BUF_QMASK, BUF_SPACE_IDX) = range(9)
BUF_DTYPES = {BUF_POS: np.float64, BUF_LOCAL_AABB: np.float64, BUF_RADIUS: np.float64, BUF_TAGS: np.uint64,
              BUF_META: np.uint32, BUF_VERT_OFF: np.uint32, BUF_VERTS: np.float64, BUF_QMASK: np.uint64,
              BUF_SPACE_IDX: np.uint32}
META_POLYGON, META_CLOSED, META_TESTER, META_USE_ANY, META_ABSENT, META_NV_SHIFT = 1, 2, 4, 8, 16, 8
META_STATIC = 0x20
FRAME_DOWNLOAD, FRAME_ALL_CANDIDATES, FRAME_TIMING, FRAME_NO_TAG_FILTERS, FRAME_TESTER_TABLE = 1, 2, 4, 8, 16
FRAME_INCREMENTAL = 0x40
FRAME_CANDIDATE_WALK = 0x80
HIT_NOT_GATED = 0x80
KERNEL_NAMES = ["clear", "bounds", "scan", "scatter", "cellsort", "pairs", "narrow"]

# every symbol include/resolv_b200.h declares (tests check the built library exports all of them)
ABI_SYMBOLS = [
    "rsv_abi_version", "rsv_device_count", "rsv_create", "rsv_destroy", "rsv_last_error", "rsv_grid_dims",
    "rsv_reserve", "rsv_staging", "rsv_set_counts", "rsv_commit", "rsv_device_buffer", "rsv_frame",
    "rsv_frame_launch", "rsv_frame_finish", "rsv_results_view", "rsv_timing", "rsv_sync", "rsv_timer_start",
    "rsv_timer_stop", "rsv_debug_cells", "rsv_debug_cell_rects", "rsv_linetest", "rsv_linetest_launch",
    "rsv_linetest_finish", "rsv_ray_staging", "rsv_ray_reserve", "rsv_ray_results_view", "rsv_tester_table",
    "rsv_set_stream", "rsv_strip_setup", "rsv_halo_pack", "rsv_halo_apply", "rsv_staging_page", "rsv_commit_pos_page",
    "rsv_halo_mailbox", "rsv_ipc_export", "rsv_ipc_open", "rsv_ipc_close", "rsv_halo_connect", "rsv_halo_exchange",
    "rsv_static_invalidate", "rsv_static_info", "rsv_pos_pages", "rsv_pos_page_upload", "rsv_pos_page_select",
    "rsv_strip_frame_launch", "rsv_pos_list_staging", "rsv_commit_pos_list", "rsv_pair_staging", "rsv_test_pairs",
]


class Config(C.Structure):
    _fields_ = [("space_w", C.c_int32), ("space_h", C.c_int32), ("cell_w", C.c_int32), ("cell_h", C.c_int32),
                ("n_spaces", C.c_uint32), ("max_shapes", C.c_uint32), ("max_verts", C.c_uint32),
                ("max_entries", C.c_uint32), ("max_pairs", C.c_uint32), ("max_contacts", C.c_uint32),
                ("max_rays", C.c_uint32), ("device", C.c_int32), ("row_lo", C.c_int32), ("row_hi", C.c_int32),
                ("flags", C.c_uint32)]


class Counts(C.Structure):
    _fields_ = [("n_shapes", C.c_uint64), ("n_testers", C.c_uint64), ("n_entries", C.c_uint64),
                ("n_candidates", C.c_uint64), ("n_gated", C.c_uint64), ("n_hits", C.c_uint64),
                ("n_contacts", C.c_uint64), ("overflow", C.c_uint32), ("reserved", C.c_uint32)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_ if k != "reserved"}


class Results(C.Structure):
    _fields_ = [("hits", C.c_void_p), ("n_hit_records", C.c_uint64), ("contacts", C.c_void_p), ("n_contacts", C.c_uint64)]


class TimingInfo(C.Structure):
    _fields_ = [("kernel_us", C.c_float * 8), ("frame_us", C.c_float)]


class RayCounts(C.Structure):
    _fields_ = [("n_rays", C.c_uint64), ("n_candidates", C.c_uint64), ("n_hits", C.c_uint64), ("n_contacts", C.c_uint64),
                ("overflow", C.c_uint32), ("reserved", C.c_uint32), ("kernel_us", C.c_float), ("reserved2", C.c_float)]

    def as_dict(self):
        return {k: (float(getattr(self, k)) if k == "kernel_us" else int(getattr(self, k)))
                for k, _ in self._fields_ if not k.startswith("reserved")}


class RayResults(C.Structure):
    _fields_ = [("hits", C.c_void_p), ("n_hits", C.c_uint64), ("contacts", C.c_void_p), ("n_contacts", C.c_uint64),
                ("ray_first", C.c_void_p), ("ray_count", C.c_void_p), ("n_rays", C.c_uint64)]


LINE_DDA, LINE_RESIDENT, LINE_HOME_ONLY, FRAME_GRID_ONLY = 0x100, 0x200, 0x400, 0x20
HIT_DTYPE = np.dtype([("a", np.uint32), ("b", np.uint32), ("first_contact", np.uint32), ("count_rank", np.uint32),
                      ("mtv", np.float64, (2,))])
CONTACT_DTYPE = np.dtype([("p", np.float64, (2,)), ("n", np.float64, (2,))])
RAY_HIT_DTYPE = np.dtype([("ray", np.uint32), ("b", np.uint32), ("first_contact", np.uint32), ("count_rank", np.uint32),
                          ("mtv", np.float64, (2,)), ("center", np.float64, (2,)), ("key", np.float64), ("pad", np.float64)])


class RsvError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"resolv_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Loads the CUDA library; raises (loudly) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(make -C resolv_b200/csrc). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, u32, u64 = C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64
    L.rsv_abi_version.restype = i32
    L.rsv_device_count.restype = i32
    L.rsv_create.restype = i32
    L.rsv_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.rsv_destroy.restype = i32
    L.rsv_destroy.argtypes = [vp]
    L.rsv_last_error.restype = C.c_char_p
    L.rsv_last_error.argtypes = [vp]
    L.rsv_grid_dims.restype = i32
    L.rsv_grid_dims.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
    L.rsv_reserve.restype = i32
    L.rsv_reserve.argtypes = [vp, u32, u32, u32]
    L.rsv_staging.restype = vp
    L.rsv_staging.argtypes = [vp, i32, C.POINTER(C.c_size_t)]
    L.rsv_device_buffer.restype = vp
    L.rsv_device_buffer.argtypes = [vp, i32]
    L.rsv_set_counts.restype = i32
    L.rsv_set_counts.argtypes = [vp, u32, u32]
    L.rsv_commit.restype = i32
    L.rsv_commit.argtypes = [vp, u32, u32, u32]
    L.rsv_frame.restype = i32
    L.rsv_frame.argtypes = [vp, i32, u32, C.POINTER(Counts)]
    L.rsv_frame_launch.restype = i32
    L.rsv_frame_launch.argtypes = [vp, i32, u32]
    L.rsv_frame_finish.restype = i32
    L.rsv_frame_finish.argtypes = [vp, C.POINTER(Counts)]
    L.rsv_results_view.restype = i32
    L.rsv_results_view.argtypes = [vp, C.POINTER(Results)]
    L.rsv_timing.restype = i32
    L.rsv_timing.argtypes = [vp, C.POINTER(TimingInfo)]
    L.rsv_sync.restype = i32
    L.rsv_sync.argtypes = [vp]
    L.rsv_timer_start.restype = i32
    L.rsv_timer_start.argtypes = [vp]
    L.rsv_timer_stop.restype = i32
    L.rsv_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    L.rsv_staging_page.restype = vp
    L.rsv_staging_page.argtypes = [vp, i32, i32, C.POINTER(C.c_size_t)]
    L.rsv_commit_pos_page.restype = i32
    L.rsv_commit_pos_page.argtypes = [vp, i32, u32, u32]
    L.rsv_pos_pages.restype = i32
    L.rsv_pos_pages.argtypes = [vp, i32]
    L.rsv_pos_page_upload.restype = i32
    L.rsv_pos_page_upload.argtypes = [vp, i32, i32, u32, u32]
    L.rsv_pos_page_select.restype = i32
    L.rsv_pos_page_select.argtypes = [vp, i32]
    L.rsv_pair_staging.restype = vp
    L.rsv_pair_staging.argtypes = [vp, u32, C.POINTER(C.c_size_t)]
    L.rsv_test_pairs.restype = i32
    L.rsv_test_pairs.argtypes = [vp, u32, C.POINTER(Counts)]
    L.rsv_pos_list_staging.restype = vp
    L.rsv_pos_list_staging.argtypes = [vp, i32, C.POINTER(C.c_size_t)]
    L.rsv_commit_pos_list.restype = i32
    L.rsv_commit_pos_list.argtypes = [vp, u32]
    L.rsv_strip_frame_launch.restype = i32
    L.rsv_strip_frame_launch.argtypes = [vp, i32, u32, i32, i32, i32, i32, u32]
    L.rsv_tester_table.restype = i32
    L.rsv_tester_table.argtypes = [vp, vp, vp, u64]
    L.rsv_debug_cells.restype = i32
    L.rsv_debug_cells.argtypes = [vp, vp, u64, vp, u64]
    L.rsv_debug_cell_rects.restype = i32
    L.rsv_debug_cell_rects.argtypes = [vp, vp, u64]
    L.rsv_static_invalidate.restype = i32
    L.rsv_static_invalidate.argtypes = [vp]
    L.rsv_static_info.restype = i32
    L.rsv_static_info.argtypes = [vp, C.POINTER(u32), C.POINTER(u32)]
    if hasattr(L, "rsv_linetest"):
        L.rsv_ray_staging.restype = vp
        L.rsv_ray_staging.argtypes = [vp, i32, C.POINTER(C.c_size_t)]
        L.rsv_ray_reserve.restype = i32
        L.rsv_ray_reserve.argtypes = [vp, u32, u32]
        L.rsv_linetest.restype = i32
        L.rsv_linetest.argtypes = [vp, u32, u32, C.POINTER(RayCounts)]
        L.rsv_linetest_launch.restype = i32
        L.rsv_linetest_launch.argtypes = [vp, u32, u32]
        L.rsv_linetest_finish.restype = i32
        L.rsv_linetest_finish.argtypes = [vp, C.POINTER(RayCounts)]
        L.rsv_ray_results_view.restype = i32
        L.rsv_ray_results_view.argtypes = [vp, C.POINTER(RayResults)]
    _lib = L
    return L


def _view(ptr, nbytes, dtype):
    buf = (C.c_char * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype)


class DeviceSpace:
    """One rsv_space: owns the handle, exposes the pinned staging buffers as numpy views and runs frames."""

    def __init__(self, space_w, space_h, cell_w, cell_h, max_shapes, max_verts, n_spaces=1, device=0, row_lo=0,
                 row_hi=0, max_entries=0, max_pairs=0, max_contacts=0, max_rays=0):
        self.L = load()
        self.h = C.c_void_p()
        cfg = Config(space_w, space_h, cell_w, cell_h, n_spaces, max_shapes, max(max_verts, 1), max_entries, max_pairs,
                     max_contacts, max_rays, device, row_lo, row_hi, 0)
        rc = self.L.rsv_create(C.byref(cfg), C.byref(self.h))
        if rc != RSV_OK:
            msg = self.L.rsv_last_error(None).decode()
            self.h = None
            raise RsvError(rc, msg)
        self.cfg = cfg
        self.buf = {}
        for b, dt in BUF_DTYPES.items():
            nbytes = C.c_size_t()
            p = self.L.rsv_staging(self.h, b, C.byref(nbytes))
            self.buf[b] = _view(p, nbytes.value, dt)
        w, h = C.c_int32(), C.c_int32()
        self.L.rsv_grid_dims(self.h, C.byref(w), C.byref(h))
        self.grid_w, self.grid_h = w.value, h.value
        self.n_spaces = n_spaces
        self.rows = (row_hi - row_lo) if (row_lo or row_hi) else self.grid_h
        self.n_shapes = 0
        self.n_verts = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.rsv_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _check(self, rc):
        if rc != RSV_OK:
            raise RsvError(rc, self.L.rsv_last_error(self.h).decode())

    # ---- shape state
    def upload_world(self, w):
        """Writes a worlds.World into the staging buffers and commits everything."""
        n, nv = w.n, len(w.verts)
        self.n_shapes, self.n_verts = n, nv
        self._check(self.L.rsv_set_counts(self.h, n, nv))
        self.buf[BUF_POS][:2 * n] = w.pos.reshape(-1)
        self.buf[BUF_LOCAL_AABB][:4 * n] = w.local_aabb.reshape(-1)
        self.buf[BUF_RADIUS][:n] = w.radius
        self.buf[BUF_TAGS][:n] = w.tags
        self.buf[BUF_META][:n] = w.meta()
        self.buf[BUF_VERT_OFF][:n] = w.vert_off
        self.buf[BUF_VERTS][:2 * nv] = w.verts.reshape(-1)
        q = self.buf[BUF_QMASK][:2 * n].reshape(-1, 2)
        q[:, 0] = w.any_mask
        q[:, 1] = w.none_mask
        self.buf[BUF_SPACE_IDX][:n] = w.space_idx
        self._check(self.L.rsv_commit(self.h, (1 << 9) - 1, 0, n))
        self._check(self.L.rsv_sync(self.h))

    def set_positions(self, pos, commit=True):
        self.buf[BUF_POS][:2 * self.n_shapes] = np.asarray(pos, dtype=np.float64).reshape(-1)
        if commit:
            self.commit(1 << BUF_POS)

    def commit(self, mask, lo=0, hi=None):
        self._check(self.L.rsv_commit(self.h, mask, lo, self.n_shapes if hi is None else hi))

    def pos_page(self, page):
        """numpy view of position staging page 0 / 1 (double-buffered uploads)."""
        nb = C.c_size_t()
        p = self.L.rsv_staging_page(self.h, BUF_POS, page, C.byref(nb))
        if not p:
            raise RsvError(RSV_ERR_OOM, "rsv_staging_page failed")
        return _view(p, nb.value, np.float64)

    def commit_pos_page(self, page, lo=0, hi=None):
        self._check(self.L.rsv_commit_pos_page(self.h, page, lo, self.n_shapes if hi is None else hi))

    def device_pos_pages(self, positions):
        """Uploads every (N,2) array of `positions` into its own device position page (at most 4); frames then pick
        one with select_pos_page(i) — per-frame motion with all inputs resident in HBM."""
        self._check(self.L.rsv_pos_pages(self.h, len(positions)))
        for i, p in enumerate(positions):
            self.buf[BUF_POS][:2 * self.n_shapes] = np.asarray(p, dtype=np.float64).reshape(-1)
            self._check(self.L.rsv_pos_page_upload(self.h, i, 0, 0, self.n_shapes))
            self.sync()

    def test_pairs(self, pairs):
        """A.Intersection(B) for explicit ordered pairs [(a, b), ...] against the committed positions: returns
        (counts, hits, contacts), one hit record per pair in list order."""
        pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        nb = C.c_size_t()
        p = self.L.rsv_pair_staging(self.h, max(len(pairs), 1), C.byref(nb))
        if not p:
            raise RsvError(RSV_ERR_OOM, "rsv_pair_staging failed")
        _view(p, nb.value, np.uint32)[:2 * len(pairs)] = pairs.reshape(-1)
        c = Counts()
        self._check(self.L.rsv_test_pairs(self.h, len(pairs), C.byref(c)))
        hits, con = self.results()
        return c, hits, con

    def pos_list(self):
        """(slots u32[max_shapes], xy f64[2 * max_shapes]) pinned staging of the sparse position update."""
        nb = C.c_size_t()
        ps = self.L.rsv_pos_list_staging(self.h, 0, C.byref(nb))
        if not ps:
            raise RsvError(RSV_ERR_OOM, "rsv_pos_list_staging failed")
        slots = _view(ps, nb.value, np.uint32)
        px = self.L.rsv_pos_list_staging(self.h, 1, C.byref(nb))
        return slots, _view(px, nb.value, np.float64)

    def commit_pos_list(self, count):
        self._check(self.L.rsv_commit_pos_list(self.h, count))

    def pos_page_upload(self, device_page, staging_page, lo=0, hi=None):
        """Queues the upload of pinned staging page 0 / 1 into device position page `device_page` on the upload stream
        (overlaps the kernels of the frame in flight)."""
        self._check(self.L.rsv_pos_page_upload(self.h, device_page, staging_page, lo, self.n_shapes if hi is None else hi))

    def select_pos_page(self, page):
        self._check(self.L.rsv_pos_page_select(self.h, page))

    def sync(self):
        self._check(self.L.rsv_sync(self.h))

    # ---- incremental grid (RSV_META_STATIC + FRAME_INCREMENTAL)
    def set_static(self, mask):
        """Marks the shapes of the boolean `mask` RSV_META_STATIC (and clears the flag on the others)."""
        meta = self.buf[BUF_META][:self.n_shapes]
        meta[:] = np.where(np.asarray(mask, dtype=bool), meta | np.uint32(META_STATIC), meta & np.uint32(~META_STATIC & 0xffffffff))
        self.commit(1 << BUF_META)

    def static_invalidate(self):
        self._check(self.L.rsv_static_invalidate(self.h))

    def static_info(self):
        """(registrations held by the cached static grid, shapes re-registered per frame)."""
        a, b = C.c_uint32(), C.c_uint32()
        self._check(self.L.rsv_static_info(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def reserve(self, entries=0, pairs=0, contacts=0):
        self._check(self.L.rsv_reserve(self.h, entries, pairs, contacts))

    # ---- frames
    def frame(self, margin=1, flags=FRAME_DOWNLOAD, grow=False):
        """One full frame. grow=True implements the documented RSV_ERR_CAPACITY contract: reserve what the
        counters say is needed (plus 25 %) and run the frame again."""
        c = Counts()
        rc = self.L.rsv_frame(self.h, margin, flags, C.byref(c))
        # a later stage's need is only known once the earlier stage's buffer was large enough: up to 3 rounds
        for _ in range(3):
            if not (rc == RSV_ERR_CAPACITY and grow):
                break
            need = lambda v: int(v * 1.25) + 1024
            self.reserve(need(c.n_entries), need(c.n_gated), need(c.n_contacts))
            rc = self.L.rsv_frame(self.h, margin, flags, C.byref(c))
        self._check(rc)
        return c

    def frame_launch(self, margin=1, flags=0):
        self._check(self.L.rsv_frame_launch(self.h, margin, flags))

    def frame_finish(self):
        c = Counts()
        rc = self.L.rsv_frame_finish(self.h, C.byref(c))
        self._last_counts = c      # (filled even when the frame overflowed: the counters say what to reserve)
        self._check(rc)
        return c

    def last_counts(self):
        return self._last_counts

    def drain(self):
        """Finishes whatever frames are still in flight (at most two) and drops their results."""
        c = Counts()
        for _ in range(2):
            if self.L.rsv_frame_finish(self.h, C.byref(c)) == RSV_ERR_BAD_ARG:
                break

    def results(self):
        """(hits, contacts) structured numpy views of the pinned mirrors (valid until the next frame)."""
        r = Results()
        self._check(self.L.rsv_results_view(self.h, C.byref(r)))
        hits = _view(r.hits, int(r.n_hit_records) * HIT_DTYPE.itemsize, HIT_DTYPE) if r.n_hit_records else np.zeros(0, HIT_DTYPE)
        con = _view(r.contacts, int(r.n_contacts) * CONTACT_DTYPE.itemsize, CONTACT_DTYPE) if r.n_contacts else np.zeros(0, CONTACT_DTYPE)
        return hits, con

    def tester_table(self):
        """(first, count) per slot of the last frame run with FRAME_TESTER_TABLE."""
        first = np.zeros(self.n_shapes, dtype=np.uint32)
        count = np.zeros(self.n_shapes, dtype=np.uint32)
        self._check(self.L.rsv_tester_table(self.h, first.ctypes.data, count.ctypes.data, self.n_shapes))
        return first, count

    def timing(self):
        t = TimingInfo()
        self._check(self.L.rsv_timing(self.h, C.byref(t)))
        d = {KERNEL_NAMES[i]: float(t.kernel_us[i]) for i in range(len(KERNEL_NAMES))}
        d["frame"] = float(t.frame_us)
        return d

    def timer_start(self):
        self._check(self.L.rsv_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        self._check(self.L.rsv_timer_stop(self.h, C.byref(ms)))
        return float(ms.value)

    # ---- parity dumps
    def debug_cells(self, n_entries):
        n_cells = self.grid_w * self.rows * self.n_spaces
        end = np.zeros(n_cells, dtype=np.uint32)
        ent = np.zeros(max(int(n_entries), 1), dtype=np.uint32)
        self._check(self.L.rsv_debug_cells(self.h, end.ctypes.data, n_cells, ent.ctypes.data, len(ent)))
        return end, ent[:int(n_entries)]

    def debug_cell_rects(self):
        r = np.zeros((self.n_shapes, 4), dtype=np.int32)
        self._check(self.L.rsv_debug_cell_rects(self.h, r.ctypes.data, self.n_shapes))
        return r

    # ---- line tests
    def linetest(self, rays, use_any=None, any_mask=None, none_mask=None, flags=FRAME_DOWNLOAD, space_idx=None, grow=True):
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 4)
        n = len(rays)
        nb = C.c_size_t()
        p = self.L.rsv_ray_staging(self.h, 0, C.byref(nb))
        seg = _view(p, nb.value, np.float64)
        if 4 * n > len(seg):
            raise RsvError(RSV_ERR_BAD_ARG, f"{n} rays exceed max_rays")
        seg[:4 * n] = rays.reshape(-1)
        p = self.L.rsv_ray_staging(self.h, 1, C.byref(nb))
        msk = _view(p, nb.value, np.uint64)[:3 * n].reshape(-1, 3)
        msk[:, 0] = 0 if any_mask is None else any_mask
        msk[:, 1] = 0 if none_mask is None else none_mask
        word = np.zeros(n, dtype=np.uint64) if use_any is None else np.asarray(use_any, dtype=np.uint64) & np.uint64(1)
        if space_idx is not None:
            word = word | (np.asarray(space_idx, dtype=np.uint64) << np.uint64(32))
        msk[:, 2] = word
        c = RayCounts()
        rc = self.L.rsv_linetest(self.h, n, flags, C.byref(c))
        if rc == RSV_ERR_CAPACITY and grow:
            self._check(self.L.rsv_ray_reserve(self.h, int(c.n_hits * 1.25) + 1024, int(c.n_contacts * 1.25) + 1024))
            rc = self.L.rsv_linetest(self.h, n, flags, C.byref(c))
        self._check(rc)
        return c

    def ray_results(self):
        """(hits, contacts, ray_first, ray_count) views of the pinned mirrors of the last downloaded rsv_linetest."""
        r = RayResults()
        self._check(self.L.rsv_ray_results_view(self.h, C.byref(r)))
        hits = _view(r.hits, int(r.n_hits) * RAY_HIT_DTYPE.itemsize, RAY_HIT_DTYPE) if r.n_hits else np.zeros(0, RAY_HIT_DTYPE)
        con = _view(r.contacts, int(r.n_contacts) * CONTACT_DTYPE.itemsize, CONTACT_DTYPE) if r.n_contacts else np.zeros(0, CONTACT_DTYPE)
        first = _view(r.ray_first, int(r.n_rays) * 4, np.uint32) if r.n_rays else np.zeros(0, np.uint32)
        count = _view(r.ray_count, int(r.n_rays) * 4, np.uint32) if r.n_rays else np.zeros(0, np.uint32)
        return hits, con, first, count


def space_for_world(w, device=0, **kw) -> DeviceSpace:
    s = DeviceSpace(w.space_w, w.space_h, w.cell_w, w.cell_h, max_shapes=max(w.n, 1), max_verts=max(len(w.verts), 1),
                    n_spaces=w.n_spaces, device=device, **kw)
    s.upload_world(w)
    return s
