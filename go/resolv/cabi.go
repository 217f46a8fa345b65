package resolv

/*
#cgo CFLAGS:  -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../resolv_b200 -lresolv_b200 -Wl,-rpath,${SRCDIR}/../../resolv_b200
#include "resolv_b200.h"
*/
import "C"

import (
	"errors"
	"sort"
	"unsafe"
)

// rsvErr turns a status code of the C ABI into a Go error (the library keeps the message).
func rsvErr(h *C.rsv_space, rc C.int32_t) error {
	if rc == C.RSV_OK {
		return nil
	}
	return errors.New("resolv: " + C.GoString(C.rsv_last_error(h)))
}

// staging returns a Go slice over the library-owned PINNED host buffer `which` (RSV_BUF_*). The memory is C memory:
// no Go pointer is ever handed to the library (cgo pointer rules), the host writes rows in place and commits ranges.
func staging[T any](h *C.rsv_space, which C.int32_t) []T {
	var n C.size_t
	p := C.rsv_staging(h, which, &n)
	var z T
	return unsafe.Slice((*T)(p), int(n)/int(unsafe.Sizeof(z)))
}

// hitsView / contactsView: borrowed views of the pinned result mirrors, valid until the next frame / pair test.
func resultsView(h *C.rsv_space) ([]C.rsv_hit, []C.rsv_contact) {
	var r C.rsv_results
	C.rsv_results_view(h, &r)
	var hits []C.rsv_hit
	var con []C.rsv_contact
	if r.n_hit_records > 0 {
		hits = unsafe.Slice(r.hits, int(r.n_hit_records))
	}
	if r.n_contacts > 0 {
		con = unsafe.Slice(r.contacts, int(r.n_contacts))
	}
	return hits, con
}

// setFromHit builds one IntersectionSet from a device record. The device delivers the contacts in the reference's
// order, except sets of more than 12 contacts of the two kinds the reference sorts (convexConvexTest, shape.go:467-469;
// convexCircleTest, shape.go:385-387): sort.Slice is pdqsort there, whose result depends on its input order, so those
// arrive in generation order and are sorted here, by the very call the reference makes.
func setFromHit(h *C.rsv_hit, contacts []C.rsv_contact, tester, other IShape) IntersectionSet {
	set := IntersectionSet{}
	n := int(h.count_rank & 0x7f)
	if n == 0 {
		return set
	}
	set.OtherShape = other
	set.MTV = Vector{float64(h.mtv_x), float64(h.mtv_y)}
	for _, c := range contacts[int(h.first_contact) : int(h.first_contact)+n] {
		set.Intersections = append(set.Intersections, Intersection{
			Point:  Vector{float64(c.px), float64(c.py)},
			Normal: Vector{float64(c.nx), float64(c.ny)},
		})
	}
	if cp, ok := tester.(*ConvexPolygon); ok && n > 12 {
		center := cp.Center()
		if c, isCircle := other.(*Circle); isCircle {
			center = c.Position()
		}
		sort.Slice(set.Intersections, func(i, j int) bool {
			return set.Intersections[i].Point.DistanceSquared(center) < set.Intersections[j].Point.DistanceSquared(center)
		})
	}
	return set
}
