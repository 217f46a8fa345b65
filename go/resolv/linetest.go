package resolv

/*
#include "resolv_b200.h"
*/
import "C"

import (
	"sort"
	"unsafe"
)

// LineTestSettings — utils.go:260-270.
type LineTestSettings struct {
	Start        Vector
	End          Vector
	TestAgainst  ShapeIterator
	OnIntersect  func(set IntersectionSet, index, max int) bool
	callingShape IShape
}

type lineCand struct {
	key float64
	set IntersectionSet
}

// LineTest — utils.go:276-370 through rsv_linetest: the device tests the ray against every shape registered in the
// cells under the cast line's bounds (== space.FilterCells(bounds).FilterShapes(), bit for bit and in the same
// generation order); the host keeps the sets whose shape the caller's iterator yields, sorts them with sort.Slice by
// the distance of their first contact (utils.go:355-357) and replays OnIntersect.
func LineTest(settings LineTestSettings) bool {
	if settings.TestAgainst == nil {
		return false
	}
	// which shapes the iterator yields (and which Space they live in)
	wanted := map[uint32]bool{}
	var sp *Space
	settings.TestAgainst.ForEach(func(other IShape) bool {
		if other != settings.callingShape {
			wanted[other.ID()] = true
			if sp == nil {
				sp = other.Space()
			}
		}
		return true
	})
	if sp == nil {
		return false
	}
	sp.mu.Lock()
	if sp.stale { // rays need the grid of the current positions, not the pair tests
		sp.upload(false)
		var c C.rsv_counts
		C.rsv_frame(sp.h, 1, C.RSV_FRAME_GRID_ONLY, &c)
	}
	var nb C.size_t
	seg := unsafe.Slice((*float64)(C.rsv_ray_staging(sp.h, 0, &nb)), 4)
	msk := unsafe.Slice((*uint64)(C.rsv_ray_staging(sp.h, 1, &nb)), 3)
	seg[0], seg[1], seg[2], seg[3] = settings.Start.X, settings.Start.Y, settings.End.X, settings.End.Y
	msk[0], msk[1], msk[2] = 0, 0, 0
	var rc C.rsv_ray_counts
	if st := C.rsv_linetest(sp.h, 1, C.RSV_FRAME_DOWNLOAD, &rc); st != C.RSV_OK {
		sp.mu.Unlock()
		panic(rsvErr(sp.h, st))
	}
	var rr C.rsv_ray_results
	C.rsv_ray_results_view(sp.h, &rr)
	var list []lineCand
	if rr.n_hits > 0 {
		hits := unsafe.Slice(rr.hits, int(rr.n_hits))
		var contacts []C.rsv_contact
		if rr.n_contacts > 0 {
			contacts = unsafe.Slice(rr.contacts, int(rr.n_contacts))
		}
		for i := range hits {
			h := &hits[i]
			other := sp.shapes[h.b]
			if !wanted[other.ID()] {
				continue
			}
			set := IntersectionSet{OtherShape: other, MTV: Vector{float64(h.mtv_x), float64(h.mtv_y)},
				Center: Vector{float64(h.center_x), float64(h.center_y)}}
			n := int(h.count_rank & 0xff)
			for _, k := range contacts[int(h.first_contact) : int(h.first_contact)+n] {
				set.Intersections = append(set.Intersections, Intersection{Point: Vector{float64(k.px), float64(k.py)},
					Normal: Vector{float64(k.nx), float64(k.ny)}})
			}
			list = append(list, lineCand{float64(h.key), set})
		}
	}
	sp.mu.Unlock()
	sort.Slice(list, func(i, j int) bool { return list[i].key < list[j].key })
	for i, c := range list {
		if settings.OnIntersect != nil && !settings.OnIntersect(c.set, i, len(list)) {
			break
		}
	}
	return len(list) > 0
}
