package resolv

/*
#include "resolv_b200.h"
*/
import "C"

import (
	"math"
	"sync"
	"unsafe"
)

// Space — space.go:7-13. The cell grid itself lives on the device (rebuilt every frame by a counting sort of
// (cell, shape) registrations); the host keeps the shape list and the pinned SoA rows the device reads.
type Space struct {
	h                     *C.rsv_space
	mu                    sync.Mutex // one stream per rsv_space: calls are serialised (the reference is single-goroutine)
	shapes                ShapeCollection
	width, height         int // in cells
	cellWidth, cellHeight int
	nVerts                uint32
	maxShapes, maxVerts   uint32

	// views over the library-owned pinned staging (C memory, never Go heap)
	pos, laabb, radius, verts []float64
	tagsBuf                   []uint64
	meta, vertOff             []uint32

	// frame cache
	stale                 bool
	margin                int
	moveEpoch, frameEpoch uint64 // bumped by every move / edit; value at the last device frame
	testerFirst           []uint32
	testerCount           []uint32
	cellEnds, cellEntries []uint32 // snapshot of the device grid (Space.Cell / FilterCells), valid for cellEpoch
	cellEpoch             uint64

	// dirty tracking + static / dynamic classification for incremental frames (RSV_META_STATIC, RSV_FRAME_INCREMENTAL)
	fullCommit, metaDirty bool
	posLo, posHi          uint32
	moved                 []bool
	still                 []uint32
	frameNo               uint32
}

const (
	stillFrames  = 2 // Steps a shape must stand still before it is flagged static
	promoteEvery = 4 // promotions rewrite META (which drops the device's static cache): batch them
)

// NewSpace — space.go:16-30. Panics when no device space can be created: there is no CPU fallback.
func NewSpace(spaceWidth, spaceHeight, cellWidth, cellHeight int) *Space {
	sp := &Space{cellWidth: cellWidth, cellHeight: cellHeight, maxShapes: 1 << 16, maxVerts: 1 << 19, stale: true,
		fullCommit: true, margin: -1, posLo: math.MaxUint32}
	cfg := C.rsv_config{space_w: C.int32_t(spaceWidth), space_h: C.int32_t(spaceHeight), cell_w: C.int32_t(cellWidth),
		cell_h: C.int32_t(cellHeight), n_spaces: 1, max_shapes: C.uint32_t(sp.maxShapes), max_verts: C.uint32_t(sp.maxVerts)}
	if rc := C.rsv_create(&cfg, &sp.h); rc != C.RSV_OK {
		panic("resolv: " + C.GoString(C.rsv_last_error(nil)))
	}
	var w, h C.int32_t
	C.rsv_grid_dims(sp.h, &w, &h) // ceil(spaceWidth / cellWidth) x ceil(spaceHeight / cellHeight), space.go:23
	sp.width, sp.height = int(w), int(h)
	sp.pos = staging[float64](sp.h, C.RSV_BUF_POS)
	sp.laabb = staging[float64](sp.h, C.RSV_BUF_LOCAL_AABB)
	sp.radius = staging[float64](sp.h, C.RSV_BUF_RADIUS)
	sp.tagsBuf = staging[uint64](sp.h, C.RSV_BUF_TAGS)
	sp.meta = staging[uint32](sp.h, C.RSV_BUF_META)
	sp.vertOff = staging[uint32](sp.h, C.RSV_BUF_VERT_OFF)
	sp.verts = staging[float64](sp.h, C.RSV_BUF_VERTS)
	return sp
}

// Destroy releases the device space (additive: the reference relies on the garbage collector).
func (s *Space) Destroy() {
	if s.h != nil {
		C.rsv_destroy(s.h)
		s.h = nil
	}
}

// Add — space.go:33-46.
func (s *Space) Add(shapes ...IShape) {
	for _, shape := range shapes {
		if uint32(len(s.shapes)) >= s.maxShapes {
			panic("resolv: Space capacity exceeded")
		}
		b := shape.base()
		b.space, b.slot = s, uint32(len(s.shapes))
		s.shapes = append(s.shapes, shape)
		s.moved = append(s.moved, true)
		s.still = append(s.still, 0)
		shape.writeRow(s, b.slot, true)
	}
	C.rsv_set_counts(s.h, C.uint32_t(len(s.shapes)), C.uint32_t(s.nVerts))
	s.stale, s.fullCommit = true, true
	s.moveEpoch++
}

// Remove — space.go:50-66. The slot stays (flagged absent: in no cell, never a tester) so other slots keep their numbers.
func (s *Space) Remove(shapes ...IShape) {
	for _, shape := range shapes {
		b := shape.base()
		if b.space != s {
			continue
		}
		s.meta[b.slot] |= C.RSV_META_ABSENT
		b.space = nil
	}
	s.stale, s.fullCommit = true, true
	s.moveEpoch++
}

// RemoveAll — space.go:69-85.
func (s *Space) RemoveAll() { s.Remove(s.Shapes()...) }

// Shapes — space.go:88-90 (live shapes, Add order).
func (s *Space) Shapes() ShapeCollection {
	out := make(ShapeCollection, 0, len(s.shapes))
	for _, shape := range s.shapes {
		if shape.base().space == s {
			out = append(out, shape)
		}
	}
	return out
}

// ForEachShape — space.go:95-103.
func (s *Space) ForEachShape(forEach func(shape IShape, index, maxCount int) bool) {
	live := s.Shapes()
	for i, o := range live {
		if !forEach(o, i, len(live)) {
			break
		}
	}
}

// FilterShapes — space.go:106-110.
func (s *Space) FilterShapes() ShapeFilter { return ShapeFilter{all: s} }

// FilterCells — space.go:181-195.
func (s *Space) FilterCells(bounds Bounds) CellSelection {
	cx := func(v float64, c int) int { return int(math.Floor(v / float64(c))) } // utils.go:90-98
	return CellSelection{space: s, explicit: true, StartX: cx(bounds.Min.X, s.cellWidth), StartY: cx(bounds.Min.Y, s.cellHeight),
		EndX: cx(bounds.Max.X, s.cellWidth), EndY: cx(bounds.Max.Y, s.cellHeight)}
}

func (s *Space) Width() int         { return s.width * s.cellWidth }
func (s *Space) Height() int        { return s.height * s.cellHeight }
func (s *Space) WidthInCells() int  { return s.width }
func (s *Space) HeightInCells() int { return s.height }
func (s *Space) CellWidth() int     { return s.cellWidth }
func (s *Space) CellHeight() int    { return s.cellHeight }

// ---- dirty tracking ---------------------------------------------------------------------------------------------

func (s *Space) markMoved(b *ShapeBase) { // ShapeBase.update, shape.go:239-242
	i := b.slot
	s.pos[2*i], s.pos[2*i+1] = b.position.X, b.position.Y
	s.moved[i] = true
	if s.meta[i]&C.RSV_META_STATIC != 0 { // demotion is immediate
		s.meta[i] &^= C.RSV_META_STATIC
		s.metaDirty = true
	}
	if i < s.posLo {
		s.posLo = i
	}
	if i+1 > s.posHi {
		s.posHi = i + 1
	}
	s.stale = true
	s.moveEpoch++
}

func (s *Space) markEdited(b *ShapeBase) { // radius / points / rotation / scale changed: the whole row
	b.owner.writeRow(s, b.slot, false)
	s.moved[b.slot] = true
	s.stale, s.fullCommit = true, true
	s.moveEpoch++
}

func (s *Space) allocVerts(n uint32) uint32 {
	if s.nVerts+n > s.maxVerts {
		panic("resolv: vertex pool exhausted")
	}
	off := s.nVerts
	s.nVerts += n
	return off
}

// upload commits what changed since the last upload: whole rows after Add / edits, otherwise the dirty position range.
func (s *Space) upload(tagsDirty bool) {
	n := C.uint32_t(len(s.shapes))
	if s.fullCommit {
		C.rsv_set_counts(s.h, n, C.uint32_t(s.nVerts))
		C.rsv_commit(s.h, (1<<C.RSV_BUF_COUNT)-1, 0, n)
	} else {
		var mask C.uint32_t
		if s.metaDirty {
			mask |= 1 << C.RSV_BUF_META
		}
		if tagsDirty {
			mask |= 1 << C.RSV_BUF_TAGS
		}
		if mask != 0 {
			if s.posHi > s.posLo {
				mask |= 1 << C.RSV_BUF_POS
			}
			C.rsv_commit(s.h, mask, 0, n)
		} else if s.posHi > s.posLo {
			C.rsv_commit(s.h, 1<<C.RSV_BUF_POS, C.uint32_t(s.posLo), C.uint32_t(s.posHi))
		}
	}
	s.fullCommit, s.metaDirty = false, false
	s.posLo, s.posHi = math.MaxUint32, 0
}

// ---- the frame ---------------------------------------------------------------------------------------------------

// Step runs one device frame for all shapes (additive API; the reference has no frame concept): grid rebuild,
// candidate pairs of every shape against SelectTouchingCells(margin).FilterShapes() — EVERY candidate gets a record,
// the replay needs the whole list for the distance sort — and the narrowphase. Frames are incremental: shapes that
// stood still for stillFrames Steps keep their cell registrations on the device, like shapes whose update() never
// runs keep theirs in the reference (shape.go:239-242).
func (s *Space) Step(margin int) error {
	s.mu.Lock()
	defer s.mu.Unlock()
	n := len(s.shapes)
	pending, tagsDirty := 0, false
	for i := 0; i < n; i++ {
		if s.moved[i] {
			s.still[i] = 0
		} else if s.still[i] < math.MaxUint32 {
			s.still[i]++
		}
		s.moved[i] = false
		if s.still[i] >= stillFrames && s.meta[i]&(C.RSV_META_STATIC|C.RSV_META_ABSENT) == 0 {
			pending++
		}
		if t := uint64(*s.shapes[i].Tags()); s.tagsBuf[i] != t { // Tags() hands out a pointer: compare
			s.tagsBuf[i], tagsDirty = t, true
		}
	}
	if pending > 0 && (s.metaDirty || s.fullCommit || s.frameNo%promoteEvery == 0) {
		for i := 0; i < n; i++ {
			if s.still[i] >= stillFrames && s.meta[i]&C.RSV_META_ABSENT == 0 {
				s.meta[i] |= C.RSV_META_STATIC
			}
		}
		s.metaDirty = true
	}
	s.frameNo++
	s.upload(tagsDirty)
	const flags = C.RSV_FRAME_DOWNLOAD | C.RSV_FRAME_TESTER_TABLE | C.RSV_FRAME_NO_TAG_FILTERS | C.RSV_FRAME_INCREMENTAL |
		C.RSV_FRAME_ALL_CANDIDATES
	var c C.rsv_counts
	rc := C.rsv_frame(s.h, C.int32_t(margin), flags, &c)
	for tries := 0; rc == C.RSV_ERR_CAPACITY && tries < 3; tries++ { // grow and retry: the counters say what is needed
		C.rsv_reserve(s.h, C.uint32_t(c.n_entries*5/4+1024), C.uint32_t(c.n_gated*5/4+1024), C.uint32_t(c.n_contacts*5/4+1024))
		rc = C.rsv_frame(s.h, C.int32_t(margin), flags, &c)
	}
	if rc != C.RSV_OK {
		return rsvErr(s.h, rc)
	}
	if len(s.testerFirst) < n {
		s.testerFirst, s.testerCount = make([]uint32, n), make([]uint32, n)
	}
	if n > 0 {
		// (Go slices passed for the duration of the call only; they hold no Go pointers)
		C.rsv_tester_table(s.h, (*C.uint32_t)(unsafe.Pointer(&s.testerFirst[0])), (*C.uint32_t)(unsafe.Pointer(&s.testerCount[0])), C.uint64_t(n))
	}
	s.margin, s.stale, s.frameEpoch = margin, false, s.moveEpoch
	return nil
}

func (s *Space) ensureFrame(margin int) {
	if s.stale || margin != s.margin {
		if err := s.Step(margin); err != nil {
			panic(err)
		}
	}
}

// testPairs — A.Intersection(B) for explicit ordered pairs against the CURRENT positions (rsv_test_pairs): what the
// reference computes when IntersectionTest reaches a candidate after a callback has moved something (shape.go:298), and
// what Shape.Intersection / IsIntersecting compute. One set per pair, in order. Invalidates the frame's result views.
func (s *Space) testPairs(pairs [][2]IShape) []IntersectionSet {
	out := make([]IntersectionSet, len(pairs))
	if len(pairs) == 0 {
		return out
	}
	s.mu.Lock()
	defer s.mu.Unlock()
	s.upload(false)
	var nb C.size_t
	list := unsafe.Slice((*uint32)(C.rsv_pair_staging(s.h, C.uint32_t(len(pairs)), &nb)), 2*len(pairs))
	for k, p := range pairs {
		list[2*k], list[2*k+1] = p[0].base().slot, p[1].base().slot
	}
	var c C.rsv_counts
	rc := C.rsv_test_pairs(s.h, C.uint32_t(len(pairs)), &c)
	for tries := 0; rc == C.RSV_ERR_CAPACITY && tries < 3; tries++ {
		need := uint64(c.n_gated)
		if uint64(len(pairs)) > need {
			need = uint64(len(pairs))
		}
		C.rsv_reserve(s.h, 0, C.uint32_t(need*5/4+1024), C.uint32_t(c.n_contacts*5/4+1024))
		rc = C.rsv_test_pairs(s.h, C.uint32_t(len(pairs)), &c)
	}
	if rc != C.RSV_OK {
		panic(rsvErr(s.h, rc))
	}
	s.stale = true // the result mirrors now hold the pair records
	hits, contacts := resultsView(s.h)
	for k := range pairs {
		if k < len(hits) {
			out[k] = setFromHit(&hits[k], contacts, pairs[k][0], pairs[k][1])
		}
	}
	return out
}

// Cell — cell.go:4-7, space.go:135-142: a read-only snapshot of cell (cx, cy) of the device grid after the last frame
// (nil outside the grid). Shapes are in the cell's stored order (ascending Add order: the order of a freshly built
// reference Space).
type Cell struct {
	X, Y   int
	Shapes []IShape
}

func (s *Space) Cell(cx, cy int) *Cell {
	if cx < 0 || cy < 0 || cx >= s.width || cy >= s.height {
		return nil
	}
	s.snapshotCells()
	c := cy*s.width + cx
	lo := uint32(0)
	if c > 0 {
		lo = s.cellEnds[c-1]
	}
	cell := &Cell{X: cx, Y: cy}
	for _, slot := range s.cellEntries[lo:s.cellEnds[c]] {
		cell.Shapes = append(cell.Shapes, s.shapes[slot])
	}
	return cell
}

// snapshotCells reads the cell membership of the current positions back from the device (rsv_debug_cells), once per
// move epoch.
func (s *Space) snapshotCells() {
	if !s.stale && s.cellEpoch == s.frameEpoch && s.cellEnds != nil {
		return
	}
	s.ensureFrame(maxInt(s.margin, 1))
	nCells := s.width * s.height
	s.cellEnds = make([]uint32, nCells)
	C.rsv_debug_cells(s.h, (*C.uint32_t)(unsafe.Pointer(&s.cellEnds[0])), C.uint64_t(nCells), nil, 0)
	total := int(s.cellEnds[nCells-1])
	s.cellEntries = make([]uint32, total+1)
	if total > 0 {
		C.rsv_debug_cells(s.h, nil, 0, (*C.uint32_t)(unsafe.Pointer(&s.cellEntries[0])), C.uint64_t(total))
	}
	s.cellEntries = s.cellEntries[:total]
	s.cellEpoch = s.frameEpoch
}

// shapesInCells — the shapes registered in the cell rect, row-major, first occurrence only (cell.go:86-121), from the
// snapshot of the device grid. Only explicit Space.FilterCells selections come here.
func (s *Space) shapesInCells(x0, y0, x1, y1 int) ShapeCollection {
	s.snapshotCells()
	var out ShapeCollection
	seen := map[uint32]bool{}
	for y := maxInt(y0, 0); y <= y1 && y < s.height; y++ {
		for x := maxInt(x0, 0); x <= x1 && x < s.width; x++ {
			c := y*s.width + x
			lo := uint32(0)
			if c > 0 {
				lo = s.cellEnds[c-1]
			}
			for _, slot := range s.cellEntries[lo:s.cellEnds[c]] {
				if !seen[slot] {
					seen[slot] = true
					out = append(out, s.shapes[slot])
				}
			}
		}
	}
	return out
}

func maxInt(a, b int) int {
	if a > b {
		return a
	}
	return b
}
