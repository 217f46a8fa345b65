/* resolv_b200.h — C ABI of libresolv_b200.so, the B200-native replacement for resolv's collision hot path.
 *
 * The reference (SolarLune/resolv v0.8.0, pure Go) has no FFI seam; the drop-in is a Go package with the same
 * import path and exported API whose Space / ShapeBase methods call these entry points through cgo
 * (INTEGRATION.md shows the binding). Each entry point cites the reference code it replaces (file:line relative
 * to the reference tree).
 *
 * Conventions
 *  - Every function returns int32_t: RSV_OK (0) or a negative RSV_ERR_* code. No exceptions, no abort, no
 *    callbacks into the host language. rsv_last_error() returns a human-readable message for the last failure.
 *  - All device memory and all pinned host staging / result buffers are owned by the library and freed by
 *    rsv_destroy(). The library never retains a caller pointer after a call returns (cgo pointer rules).
 *  - One CUDA stream per rsv_space. Calls on one space must be serialised by the caller; different spaces may
 *    be driven from different threads. Every call selects the space's device itself.
 *  - There is NO CPU fallback: rsv_create fails with RSV_ERR_NO_DEVICE when no CUDA device is usable.
 *  - Shapes are addressed by slot index i in [0, n_shapes); the host layer maps resolv shape IDs to slots.
 */
#ifndef RESOLV_B200_H
#define RESOLV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSV_ABI_VERSION 1

#define RSV_OK 0
#define RSV_ERR_BAD_ARG (-1)
#define RSV_ERR_OOM (-2)
#define RSV_ERR_CUDA (-3)
#define RSV_ERR_CAPACITY (-4) /* an output buffer overflowed; counts hold the required sizes: rsv_reserve + retry */
#define RSV_ERR_NO_DEVICE (-5)

typedef struct rsv_space rsv_space;

/* Replaces NewSpace(spaceWidth, spaceHeight, cellWidth, cellHeight) (space.go:16-30) and Space.Resize
 * (space.go:113-131): grid = ceil(space_w/cell_w) x ceil(space_h/cell_h) cells. n_spaces > 1 creates a batch
 * of independent, identically-sized Spaces in one device grid (RL-style env batches; shape i belongs to space
 * space_idx[i]). row_lo/row_hi select the strip of cell rows [row_lo,row_hi) held by this device for
 * multi-GPU strip partitioning (0,0 = the whole grid): cells outside the strip behave as if they held no
 * shapes, exactly like out-of-grid cells (space.go:135-142). */
typedef struct rsv_config {
    int32_t space_w, space_h;
    int32_t cell_w, cell_h;
    uint32_t n_spaces;
    uint32_t max_shapes;   /* capacity N (shape slots) */
    uint32_t max_verts;    /* capacity of the shared polygon vertex pool (sum of vertex counts) */
    uint32_t max_entries;  /* capacity R of (cell, shape) registrations; 0 = auto */
    uint32_t max_pairs;    /* capacity of the AABB-gated pair / hit-header buffer; 0 = auto */
    uint32_t max_contacts; /* capacity of the contact buffer; 0 = auto */
    uint32_t max_rays;     /* capacity of one rsv_linetest batch; 0 = auto */
    int32_t device;        /* CUDA device ordinal */
    int32_t row_lo, row_hi;
    uint32_t flags;        /* reserved, 0 */
} rsv_config;

/* Host-visible SoA staging buffers (library-owned pinned memory the host writes directly, then rsv_commit()s).
 * This is the device-resident restatement of ShapeBase / Circle / ConvexPolygon state (shape.go:54-62,
 * circle.go:4-7, convexPolygon.go:12-20). */
enum {
    RSV_BUF_POS = 0,        /* f64[2N]  ShapeBase.position */
    RSV_BUF_LOCAL_AABB = 1, /* f64[4N]  polygon: ConvexPolygon.bounds as cached by updateBounds (convexPolygon.go:163-197,
                                        minX,minY,maxX,maxY, position NOT applied); circle: (-r,-r,+r,+r) (circle.go:33-39) */
    RSV_BUF_RADIUS = 2,     /* f64[N]   Circle.radius (0 for polygons) */
    RSV_BUF_TAGS = 3,       /* u64[N]   *ShapeBase.tags (tags.go:8) */
    RSV_BUF_META = 4,       /* u32[N]   RSV_META_* bits | vertex count << 8 */
    RSV_BUF_VERT_OFF = 5,   /* u32[N]   first vertex of polygon i in RSV_BUF_VERTS */
    RSV_BUF_VERTS = 6,      /* f64[2*max_verts] polygon points with scale and rotation applied and position NOT applied:
                                        the value of p in ConvexPolygon.Transformed before "+ position" (convexPolygon.go:146-151) */
    RSV_BUF_QMASK = 7,      /* u64[2N]  per tester: (ByTags mask, OR of NotByTags masks) (shapefilter.go:47-61) */
    RSV_BUF_SPACE_IDX = 8,  /* u32[N]   which Space of the batch the shape was Add()ed to (0 when n_spaces == 1) */
    RSV_BUF_COUNT = 9
};

#define RSV_META_POLYGON 0x1u  /* 0 = Circle, 1 = ConvexPolygon */
#define RSV_META_CLOSED 0x2u   /* ConvexPolygon.Closed (convexPolygon.go:18) */
#define RSV_META_TESTER 0x4u   /* this shape runs IntersectionTest in rsv_frame */
#define RSV_META_USE_ANY 0x8u  /* tester's filter chain contains ByTags(qmask[2i]) */
#define RSV_META_ABSENT 0x10u  /* slot is free / shape was Space.Remove()d (space.go:50-66): in no cell, never a tester */
#define RSV_META_STATIC 0x20u  /* the host promises that this shape has not moved (position, local bounds, tags, Space)
                                  since the flag was committed: RSV_FRAME_INCREMENTAL keeps its cell registrations. The
                                  counterpart of the reference only touching the cells of the shape whose update()
                                  ran (shape.go:239-242); tiles and walls of a platformer never re-register. */
#define RSV_META_NV_SHIFT 8
#define RSV_MAX_POLY_VERTS 255

/* Per-frame counters (the SURVEY's N, R, P, H, K) and overflow information. */
typedef struct rsv_counts {
    uint64_t n_shapes;      /* N */
    uint64_t n_testers;
    uint64_t n_entries;     /* R: (cell, shape) registrations == sum of len(cell.Shapes) */
    uint64_t n_candidates;  /* P: ordered (tester, other) pairs after dedupe + tag filters == pair tests. Counted by frames
                               that run the literal candidate walk (RSV_FRAME_CANDIDATE_WALK, RSV_FRAME_ALL_CANDIDATES,
                               RSV_FRAME_INCREMENTAL); 0 for the default frame, which never enumerates them */
    uint64_t n_gated;       /* ordered pairs that pass the Bounds gate (utils.go:146-173) */
    uint64_t n_hits;        /* H: ordered pairs with a non-empty IntersectionSet */
    uint64_t n_contacts;    /* K */
    uint32_t overflow;      /* RSV_OVF_* bits */
    uint32_t reserved;
} rsv_counts;

#define RSV_OVF_ENTRIES 0x1u
#define RSV_OVF_PAIRS 0x2u
#define RSV_OVF_CONTACTS 0x4u
#define RSV_OVF_PAIR_CONTACTS 0x40u /* ONE pair produced more than 127 contacts (count_rank holds 7 bits): not a buffer size,
                                        rsv_reserve + retry cannot fix it */

/* One record per AABB-gated ordered pair, grouped by tester (all records of one tester are contiguous and in the
 * reference's candidate generation order: row-major cell walk, in-cell order, first occurrence; cell.go:86-121).
 * n_contacts == 0 means the IntersectionSet was empty (no callback). Replaces IntersectionSet (contactSet.go:15-20)
 * for shape-shape tests (Center is zero there). */
typedef struct rsv_hit {
    uint32_t a;             /* tester slot */
    uint32_t b;             /* IntersectionSet.OtherShape slot */
    uint32_t first_contact; /* index into the contact buffer */
    uint32_t count_rank;    /* n_contacts in bits 0..7, candidate generation rank in bits 8..31 (tie-break for the
                               distance sort of shape.go:291-293) */
    double mtv_x, mtv_y;    /* IntersectionSet.MTV */
} rsv_hit;

/* Replaces Intersection (contactSet.go:8-11). A set's contacts are in the reference's order: sorted by squared distance
 * to the tester's Center (convex-convex, shape.go:467-469) / to the circle's position (convex-circle, shape.go:385-387),
 * unsorted for circle-convex and circle-circle. EXCEPT sets of MORE THAN 12 contacts of the two sorted kinds: for those
 * sort.Slice is Go's pdqsort_func, whose result depends on its input order, and the device delivers that input — the
 * generation order — for the host to sort (go/resolv: sort.Slice itself; csrc/host/resolv.hpp: its port GoSortSlice). */
typedef struct rsv_contact {
    double px, py; /* Point */
    double nx, ny; /* Normal */
} rsv_contact;

/* Borrowed views of the pinned result mirrors, valid until the next rsv_frame / rsv_linetest on the space. */
typedef struct rsv_results {
    const rsv_hit* hits;
    uint64_t n_hit_records; /* == counts.n_gated (dense by gated pair; filter n_contacts > 0) */
    const rsv_contact* contacts;
    uint64_t n_contacts;
} rsv_results;

#define RSV_FRAME_DOWNLOAD 0x1u        /* copy hits + contacts to the pinned mirrors (rsv_results) */
#define RSV_FRAME_ALL_CANDIDATES 0x2u  /* parity/debug: emit a record for EVERY candidate pair, not only gated ones;
                                          records failing the Bounds gate carry RSV_HIT_NOT_GATED in count_rank bit 7 */
#define RSV_FRAME_TIMING 0x4u          /* record per-kernel CUDA events; read with rsv_timing */
#define RSV_FRAME_NO_TAG_FILTERS 0x8u  /* the host asserts that no tester's TestAgainst chain contains ByTags / NotByTags
                                          this frame: RSV_BUF_QMASK / RSV_BUF_TAGS are not read (META_USE_ANY ignored) */
#define RSV_FRAME_TESTER_TABLE 0x10u   /* also fill the per-tester (first record, record count) table for random access
                                          (default frames always fill it: their records are grouped through it) */
#define RSV_FRAME_CANDIDATE_WALK 0x80u /* run the reference's literal candidate walk (every tester visits every entry of its
                                          margin-grown cell rect) so that counts.n_candidates is produced. Without it (and
                                          without RSV_FRAME_ALL_CANDIDATES) a frame finds the Bounds-gated pairs cell by cell
                                          and derives their ranks afterwards: same records, n_candidates = 0 */
#define RSV_HIT_NOT_GATED 0x80u

/* Per-kernel device times of the last RSV_FRAME_TIMING frame, microseconds, CUDA events on the space's stream. */
enum {
    RSV_K_CLEAR = 0, RSV_K_BOUNDS = 1, RSV_K_SCAN = 2, RSV_K_SCATTER = 3, RSV_K_CELLSORT = 4,
    RSV_K_PAIRS = 5, RSV_K_NARROW = 6, RSV_K_COUNT = 7
};
typedef struct rsv_timing_info {
    float kernel_us[8];
    float frame_us; /* first kernel start to last kernel end */
} rsv_timing_info;

/* ---- lifetime ------------------------------------------------------------------------------------------- */
int32_t rsv_abi_version(void);
int32_t rsv_device_count(void);
/* NewSpace (space.go:16-30). */
int32_t rsv_create(const rsv_config* cfg, rsv_space** out);
int32_t rsv_destroy(rsv_space* s);
const char* rsv_last_error(const rsv_space* s); /* s may be NULL: error of the last failed rsv_create on this thread */
/* Grid dimensions in cells: Space.WidthInCells / HeightInCells (space.go:160-171). */
int32_t rsv_grid_dims(const rsv_space* s, int32_t* w_cells, int32_t* h_cells);
/* Grow output capacities after RSV_ERR_CAPACITY (0 = keep). */
int32_t rsv_reserve(rsv_space* s, uint32_t max_entries, uint32_t max_pairs, uint32_t max_contacts);

/* ---- shape state ---------------------------------------------------------------------------------------- */
/* Pinned staging buffer `which` (RSV_BUF_*); *bytes receives its size. */
void* rsv_staging(rsv_space* s, int32_t which, size_t* bytes);
/* Number of live slots / vertices (Space.Add / Remove bookkeeping, space.go:33-66, stays on the host). */
int32_t rsv_set_counts(rsv_space* s, uint32_t n_shapes, uint32_t n_verts);
/* Upload slots [lo,hi) of every buffer in `which_mask` (bit RSV_BUF_x); RSV_BUF_VERTS uploads the whole live
 * vertex pool. Replaces the write side of ShapeBase.update (shape.go:239-242): the host only marks dirty ranges. */
int32_t rsv_commit(rsv_space* s, uint32_t which_mask, uint32_t lo, uint32_t hi);
/* Double-buffered position staging for pipelined hosts: page 0 is rsv_staging(RSV_BUF_POS), page 1 a second pinned
 * buffer of the same size. The host writes frame f+1's positions into one page while frame f's upload (queued by
 * rsv_commit_pos_page) still reads the other. Only RSV_BUF_POS has pages. */
void* rsv_staging_page(rsv_space* s, int32_t which, int32_t page, size_t* bytes);
int32_t rsv_commit_pos_page(rsv_space* s, int32_t page, uint32_t lo, uint32_t hi);
/* Device-resident position pages. A host that already holds several position sets in HBM (frames prepared ahead, or a
 * device-side integrator writing the next positions while the current frame runs) switches between them without a
 * copy: rsv_pos_pages allocates pages 1..n_pages-1 (page 0 is RSV_BUF_POS's device buffer; new pages start as copies
 * of it), rsv_pos_page_upload fills device page `device_page` from pinned staging page 0 / 1, rsv_pos_page_select
 * makes every later frame, halo pack and line test read positions from that page. Uploads run on their own stream:
 * the upload of the NEXT frame's positions overlaps the kernels of the frame in flight (the library orders an upload
 * after the frames that still read the page and a frame after the upload of the page it selects). The write side of
 * ShapeBase.update (shape.go:239-242) for hosts that move shapes in bulk. */
int32_t rsv_pos_pages(rsv_space* s, int32_t n_pages);
int32_t rsv_pos_page_upload(rsv_space* s, int32_t device_page, int32_t staging_page, uint32_t lo, uint32_t hi);
int32_t rsv_pos_page_select(rsv_space* s, int32_t device_page);
/* Sparse position updates: rsv_pos_list_staging(which = 0) is a pinned u32[max_shapes] list of moved slots,
 * (which = 1) the matching f64[2 * max_shapes] new positions; rsv_commit_pos_list(count) uploads the first `count`
 * entries of both (20 bytes per moved shape instead of the whole position buffer) and scatters them into the position
 * page the next frame reads. The write side of ShapeBase.update (shape.go:239-242) when few shapes move; the host's
 * own copy in RSV_BUF_POS staging is not touched (keep it current if rsv_commit(RSV_BUF_POS) is used as well). */
void* rsv_pos_list_staging(rsv_space* s, int32_t which, size_t* bytes);
int32_t rsv_commit_pos_list(rsv_space* s, uint32_t count);
/* Device address of buffer `which` (for device-side producers such as the multi-GPU halo unpack). */
void* rsv_device_buffer(rsv_space* s, int32_t which);

/* ---- the frame ------------------------------------------------------------------------------------------ */
/* Bulk restatement of, for every tester A (RSV_META_TESTER):
 *   A.IntersectionTest(IntersectionTestSettings{TestAgainst: A.SelectTouchingCells(margin).FilterShapes()
 *       [.ByTags(any)][.NotByTags(none)], OnIntersect: <record>})                  (shape.go:220-237, 273-316)
 * against the frozen world: grid rebuild (shape.go:183-214, cell.go:19-38) -> candidate pairs (cell.go:86-121,
 * shapefilter.go:26-61) -> Bounds gate + SAT narrowphase (shape.go:318-479, convexPolygon.go:418-514,715-789).
 * OnIntersect is replayed by the host from rsv_results.
 * The records — one per ordered pair that passes the Bounds gate, with its index in possibleIntersections, contacts and
 * MTV, grouped by tester in generation order — do not depend on HOW the candidates are found: frames with
 * RSV_FRAME_ALL_CANDIDATES / RSV_FRAME_CANDIDATE_WALK (and incremental frames) run the reference's walk over every
 * tester's margin-grown cell rect; every other frame finds the gated pairs among the entries of one cell and derives
 * their indices afterwards (csrc/rsv_cellpairs.cuh): identical records, no candidate count, about 30 % less time. */
int32_t rsv_frame(rsv_space* s, int32_t margin, uint32_t flags, rsv_counts* out);
/* The same frame split in two so that a host can queue work without blocking: rsv_frame_launch enqueues every
 * kernel on the space's stream and returns; rsv_frame_finish waits for the OLDEST launched frame, fills the counters
 * and (RSV_FRAME_DOWNLOAD) the result mirrors. rsv_frame == launch + finish. Two frames may be in flight: results
 * live in two buffer sets and the download of frame f runs on a separate copy stream, so calling
 * rsv_frame_launch(f+1) before rsv_frame_finish(f) overlaps f's device-to-host copy with f+1's upload and kernels.
 * Launching a third frame without finishing drops the oldest frame's results. */
int32_t rsv_frame_launch(rsv_space* s, int32_t margin, uint32_t flags);
int32_t rsv_frame_finish(rsv_space* s, rsv_counts* out);
int32_t rsv_results_view(rsv_space* s, rsv_results* out);
/* Explicit pair tests: A.Intersection(B) (circle.go:69-82, convexPolygon.go:288-300 -> shape.go:318-479) for a list of
 * ordered pairs, against the positions committed so far. rsv_pair_staging returns a pinned u32[2 * max_pairs] list the
 * host fills with (tester slot, other slot); rsv_test_pairs(n) runs Bounds + the exact Bounds gate + the narrowphase
 * for the first n pairs and downloads one rsv_hit per pair IN LIST ORDER (rank = list index; RSV_HIT_NOT_GATED / no
 * contacts = empty set) into the result mirrors (rsv_results_view). The grid of the last frame is not touched. This is
 * what the host layer uses when OnIntersect moved a shape in the middle of an IntersectionTest: the reference tests
 * candidate after candidate against the CURRENT positions (shape.go:296-312). */
void* rsv_pair_staging(rsv_space* s, uint32_t max_pairs, size_t* bytes);
int32_t rsv_test_pairs(rsv_space* s, uint32_t n_pairs, rsv_counts* out);
/* Per tester slot: first record and number of records in the hit buffer of the last frame run with
 * RSV_FRAME_TESTER_TABLE (random access for a single Shape.IntersectionTest call); non-testers have count 0. */
int32_t rsv_tester_table(rsv_space* s, uint32_t* first, uint32_t* count, uint64_t n_shapes);
int32_t rsv_timing(rsv_space* s, rsv_timing_info* out);
/* Block until all work queued on the space's stream is done. */
int32_t rsv_sync(rsv_space* s);
/* CUDA-event stopwatch on the space's stream (the stream every kernel of this space is launched on):
 * rsv_timer_start records an event; rsv_timer_stop records a second one, waits for it and returns the elapsed
 * device time in milliseconds. */
int32_t rsv_timer_start(rsv_space* s);
int32_t rsv_timer_stop(rsv_space* s, float* ms);

/* ---- line tests ----------------------------------------------------------------------------------------- */
/* Batched LineTest (utils.go:276-370): for every ray i,
 *   LineTest(LineTestSettings{Start, End, TestAgainst: <iterator>[.ByTags(any)][.NotByTags(none)], OnIntersect: <record>})
 * against the grid built by the last rsv_frame (use RSV_FRAME_GRID_ONLY to build the grid without shape tests).
 * Iterator, default:  space.FilterCells(Bounds of the cast line).FilterShapes()  (space.go:181-195, cell.go:73-121),
 * the cast line being (Start - 0.01 * Unit(End - Start)) -> End (utils.go:278-282).
 * Iterator, RSV_LINE_DDA: the shapes registered in the cells a grid traversal of the cast line passes through (the
 * reference never shipped its cell-walking line query, space.go:197-253; see rsv_lines.cuh for the exact rule and
 * for the one documented deviation, "phantom" contacts of the +1 fudge on far-away near-parallel edges).
 * ShapeLineTest (convexPolygon.go:322-415) is built on this by the host layer: one ray per selected vertex. */
#define RSV_LINE_DDA 0x100u
#define RSV_LINE_RESIDENT 0x200u /* rays of the previous call are still on the device: skip their upload */
#define RSV_LINE_HOME_ONLY 0x400u /* strip partitioning: report a shape only on the device whose home rows hold its first cell row,
                                     so that the union of all devices' sets for a ray is the ray's result on the whole Space */
#define RSV_FRAME_GRID_ONLY 0x20u /* rsv_frame: rebuild the grid only (no candidate pairs, no narrowphase) */
/* Incremental grid. Replaces re-registering every shape every frame — which the reference never
 * does: only ShapeBase.update of a moved shape touches cells (shape.go:183-214,239-242, cell.go:19-38) — for worlds
 * that are mostly static. The registrations of all RSV_META_STATIC shapes are built once and cached on the device;
 * a frame re-registers only the other shapes and merges both into a grid that is bit-identical to a full rebuild
 * (so every result of the frame is, too). The cache is dropped by rsv_commit of META / LOCAL_AABB / TAGS /
 * SPACE_IDX, by rsv_set_counts with a new shape count, and by rsv_static_invalidate (call it after moving a static
 * shape); the next incremental frame rebuilds it (synchronously). Ignored (full rebuild) for strip-partitioned
 * spaces and when more than half of the shapes are dynamic. */
#define RSV_FRAME_INCREMENTAL 0x40u
int32_t rsv_static_invalidate(rsv_space* s);
/* Registrations held by the cached static grid and number of shapes re-registered per frame (0 / n_shapes while no
 * cache exists). */
int32_t rsv_static_info(rsv_space* s, uint32_t* static_entries, uint32_t* n_dynamic);
#define RSV_OVF_RAY_HITS 0x8u
#define RSV_OVF_RAY_CONTACTS 0x10u

/* One record per (ray, shape) with a non-empty contact set == one IntersectionSet of LineTest (contactSet.go:15-20).
 * All records of a ray are contiguous (ray_first / ray_count) and in candidate generation order; OnIntersect order
 * is ascending `key`, ties by generation rank (utils.go:355-357 with sort.Slice's insertion sort, n <= 12). */
typedef struct rsv_ray_hit {
    uint32_t ray;           /* index of the ray in the batch */
    uint32_t b;             /* IntersectionSet.OtherShape slot */
    uint32_t first_contact; /* index into the ray contact buffer; contacts sorted by distance to Start (utils.go:340-342) */
    uint32_t count_rank;    /* n_contacts in bits 0..7, candidate generation rank in bits 8..31 */
    double mtv_x, mtv_y;       /* IntersectionSet.MTV = contacts[0].Point - Start - 0.01 * vu (utils.go:344) */
    double center_x, center_y; /* IntersectionSet.Center (utils.go:332-337) */
    double key;                /* d2(contacts[0].Point, cast line start): the sort key of utils.go:355-357 */
    double pad;
} rsv_ray_hit;

typedef struct rsv_ray_counts {
    uint64_t n_rays;
    uint64_t n_candidates; /* (ray, shape) pairs tested after dedupe + tag filters */
    uint64_t n_hits;       /* IntersectionSets */
    uint64_t n_contacts;
    uint32_t overflow;     /* RSV_OVF_RAY_* */
    uint32_t reserved;
    float kernel_us;       /* device time of the line kernel (CUDA events on the space's stream) */
    float reserved2;
} rsv_ray_counts;

typedef struct rsv_ray_results {
    const rsv_ray_hit* hits;
    uint64_t n_hits;
    const rsv_contact* contacts;
    uint64_t n_contacts;
    const uint32_t* ray_first; /* per ray: first hit record */
    const uint32_t* ray_count; /* per ray: number of hit records */
    uint64_t n_rays;
} rsv_ray_results;

/* Pinned ray staging: which = 0: f64[4*max_rays] (Start.X, Start.Y, End.X, End.Y);
 * which = 1: u64[3*max_rays] (ByTags mask, OR of NotByTags masks, bit0 = chain has ByTags | space index << 32). */
void* rsv_ray_staging(rsv_space* s, int32_t which, size_t* bytes);
int32_t rsv_ray_reserve(rsv_space* s, uint32_t max_hits, uint32_t max_contacts);
/* flags: RSV_FRAME_DOWNLOAD | RSV_LINE_DDA | RSV_LINE_RESIDENT | RSV_LINE_HOME_ONLY */
int32_t rsv_linetest(rsv_space* s, uint32_t n_rays, uint32_t flags, rsv_ray_counts* out);
int32_t rsv_linetest_launch(rsv_space* s, uint32_t n_rays, uint32_t flags);
int32_t rsv_linetest_finish(rsv_space* s, rsv_ray_counts* out);
int32_t rsv_ray_results_view(rsv_space* s, rsv_ray_results* out);

/* ---- strip partitioning of one large Space across GPUs -------------------------------------------------- */
/* The reference has no distributed layer. One process per GPU creates an rsv_space whose stored rows
 * (rsv_config.row_lo/row_hi) are its strip plus a halo; static shape data of its own and its neighbours' slot
 * ranges is resident on every device; once per frame neighbours exchange (slot, position) records of the owned
 * shapes that reach into each other's stored rows (NCCL send/recv of the buffers filled by rsv_halo_pack), then
 * every device runs rsv_frame: a shape is a tester only on the device whose home rows contain its first cell row,
 * so each ordered pair is tested exactly once and results land where the tester lives. */
#define RSV_OVF_GHOSTS 0x20u /* more halo records than the ghost list or a mailbox buffer holds: results are incomplete;
                                grow max_ghosts / the mailbox capacity (rsv_strip_setup, rsv_halo_mailbox) and run the frame again */
/* Launch every kernel of this space on an externally owned CUDA stream (e.g. the stream the host framework
 * issues its NCCL calls on) so that pack -> exchange -> apply -> frame need no host synchronisation. NULL
 * restores the library's own stream. */
int32_t rsv_set_stream(rsv_space* s, void* cuda_stream);
/* own_lo/own_hi: slot range this device moves and packs; home_lo/home_hi: cell rows whose shapes are testers
 * here; max_ghosts: capacity of the per-frame halo list. */
int32_t rsv_strip_setup(rsv_space* s, uint32_t own_lo, uint32_t own_hi, int32_t home_lo, int32_t home_hi, uint32_t max_ghosts);
/* Fill device buffers of (1 + cap_records) 24-byte records {u32 slot, u32 pad, f64 x, f64 y} (record 0: count)
 * with the owned shapes whose first cell row <= up_row_max (for the upper neighbour) / last cell row >= dn_row_min
 * (for the lower neighbour). Either buffer may be NULL. */
int32_t rsv_halo_pack(rsv_space* s, int32_t up_row_max, int32_t dn_row_min, void* dev_send_up, void* dev_send_dn, uint32_t cap_records);
/* Apply received buffers: positions are written to slot + slot_off (sender-to-receiver slot translation) and the
 * slots become this frame's ghosts. Either buffer may be NULL. */
int32_t rsv_halo_apply(rsv_space* s, const void* dev_recv_up, int32_t slot_off_up, const void* dev_recv_dn, int32_t slot_off_dn,
                       uint32_t cap_records);

/* Peer-memory halo transport (NVLink / NVSwitch P2P): no collective library on the data path. Every device owns a
 * mailbox (receive buffers for both neighbours, double-buffered by epoch parity, plus flag words); a device pushes
 * its packed records straight into its neighbours' mailboxes with peer stores, raises their `ready` flag and waits
 * for its own with stream-ordered memory waits. Set-up, once:
 *   rsv_halo_mailbox (allocate, get the device pointer) -> rsv_ipc_export (64-byte CUDA IPC handle; ship it to the
 *   neighbour processes by any host channel) -> rsv_ipc_open on the neighbours' handles -> rsv_halo_connect.
 * Then rsv_halo_exchange once per frame on every device (same number of calls everywhere, like a collective),
 * followed by rsv_frame*. Spaces living in ONE process connect with each other's mailbox pointers directly. */
#define RSV_IPC_HANDLE_BYTES 64
int32_t rsv_halo_mailbox(rsv_space* s, uint32_t cap_records, void** dev_mailbox, size_t* bytes);
int32_t rsv_ipc_export(void* dev_ptr, unsigned char handle[RSV_IPC_HANDLE_BYTES]);
int32_t rsv_ipc_open(const unsigned char handle[RSV_IPC_HANDLE_BYTES], int32_t device, void** dev_ptr);
int32_t rsv_ipc_close(void* dev_ptr);
/* up_mailbox / dn_mailbox: the neighbours' mailboxes as mapped in this process (NULL: no neighbour on that side). */
int32_t rsv_halo_connect(rsv_space* s, void* up_mailbox, void* dn_mailbox);
/* rsv_halo_pack straight into the neighbours' mailboxes + publish + wait + rsv_halo_apply, all enqueued on the
 * space's stream (arguments as for those calls). phases: RSV_HALO_SEND | RSV_HALO_RECV for the whole exchange in one
 * call (one process per GPU). A process driving SEVERAL spaces on one device must enqueue the SEND halves of all of
 * them before any RECV half: streams of one process can share a hardware queue, and a memory wait queued ahead of
 * the store it waits for would never be satisfied. */
#define RSV_HALO_SEND 1u
#define RSV_HALO_RECV 2u
int32_t rsv_halo_exchange(rsv_space* s, int32_t up_row_max, int32_t dn_row_min, int32_t slot_off_up, int32_t slot_off_dn,
                          uint32_t phases);

/* The frame of a strip-partitioned space with the halo exchange folded into it (peer-memory transport; the space must
 * be connected with rsv_halo_connect). Replaces rsv_halo_exchange + rsv_frame_launch on every device of the chain:
 * the owned shapes are registered and, in the same pass, the ones reaching a neighbour's stored rows are stored
 * straight into that neighbour's mailbox (RSV_HALO_SEND half); the device then waits — on the device, epochs and flags
 * never touch the host — for both neighbours' batches, applies them, registers the ghosts and runs the rest of the
 * frame (RSV_HALO_RECV half). One fixed launch sequence per frame, replayed as one CUDA graph; finish with
 * rsv_frame_finish as usual. phases as for rsv_halo_exchange (a process driving several spaces on ONE device enqueues
 * all SEND halves before any RECV half). After rsv_halo_connect a space uses either this call or rsv_halo_exchange. */
int32_t rsv_strip_frame_launch(rsv_space* s, int32_t margin, uint32_t flags, int32_t up_row_max, int32_t dn_row_min,
                               int32_t slot_off_up, int32_t slot_off_dn, uint32_t phases);

/* ---- parity / debug dumps --------------------------------------------------------------------------------- */
/* Cell membership after the last frame: cell_end[c] (exclusive end offset of cell c in `entries`, cells row-major,
 * spaces concatenated; start = cell_end[c-1] or 0) and the shape slots per cell in ascending slot order. Either
 * pointer may be NULL. Mirrors Cell.Shapes (cell.go:4-7). */
int32_t rsv_debug_cells(rsv_space* s, uint32_t* cell_end, uint64_t n_cells, uint32_t* entries, uint64_t cap_entries);
/* Per-shape inclusive cell rect (cx,cy,ex,ey) = Bounds.toCellSpace (utils.go:90-98) after the last frame. */
int32_t rsv_debug_cell_rects(rsv_space* s, int32_t* rects4, uint64_t n_shapes);

#ifdef __cplusplus
}
#endif
#endif /* RESOLV_B200_H */
