package resolv

/*
#include "resolv_b200.h"
*/
import "C"

import (
	"errors"
	"sort"
)

// ConvexPolygon — convexPolygon.go:12-20. Scale and rotation are applied on the host (Go's math.Sin / math.Cos are not
// reproducible on the device): the device receives Transformed() minus the position, and the cached local bounds.
type ConvexPolygon struct {
	ShapeBase
	scale    Vector
	rotation float64
	Points   []Vector
	Closed   bool
	bounds   Bounds // cached LOCAL bounds (convexPolygon.go:163-197): AABB of Transformed() at edit time, moved by -position
	vertOff  uint32 // first vertex in the shared vertex pool
	vertCap  uint32
}

// NewConvexPolygon — convexPolygon.go:28-46.
func NewConvexPolygon(x, y float64, points []float64) *ConvexPolygon {
	cp := &ConvexPolygon{ShapeBase: newShapeBase(x, y), scale: Vector{1, 1}, Closed: true}
	cp.owner = cp
	if len(points) > 0 {
		cp.AddPoints(points...)
	}
	return cp
}

// AddPoints — convexPolygon.go:92-107.
func (cp *ConvexPolygon) AddPoints(vertexPositions ...float64) error {
	if len(vertexPositions) == 0 {
		return errors.New("no vertices passed to AddPoints()")
	}
	if len(vertexPositions)%2 == 1 {
		return errors.New("odd number of float64 values passed to AddPoints(); it takes X and Y pairs")
	}
	if len(vertexPositions) < 4 {
		return errors.New("not enough vertices passed to AddPoints(); a ConvexPolygon needs at least two")
	}
	for i := 0; i < len(vertexPositions); i += 2 {
		cp.Points = append(cp.Points, Vector{vertexPositions[i], vertexPositions[i+1]})
	}
	cp.updateBounds()
	cp.edited()
	return nil
}

func (cp *ConvexPolygon) Clone() IShape { // convexPolygon.go:69-89
	pts := make([]float64, 0, 2*len(cp.Points))
	for _, p := range cp.Points {
		pts = append(pts, p.X, p.Y)
	}
	n := NewConvexPolygon(cp.position.X, cp.position.Y, pts)
	*n.tags = *cp.tags
	n.rotation, n.scale, n.Closed, n.data = cp.rotation, cp.scale, cp.Closed, cp.data
	n.updateBounds()
	return n
}

// transformedLocal — convexPolygon.go:146-150 without "+ position".
func (cp *ConvexPolygon) transformedLocal() []Vector {
	out := make([]Vector, len(cp.Points))
	for i, p := range cp.Points {
		q := Vector{p.X * cp.scale.X, p.Y * cp.scale.Y}
		if cp.rotation != 0 {
			q = q.Rotate(-cp.rotation)
		}
		out[i] = q
	}
	return out
}

// Transformed — convexPolygon.go:144-154.
func (cp *ConvexPolygon) Transformed() []Vector {
	out := cp.transformedLocal()
	for i := range out {
		out[i] = Vector{out[i].X + cp.position.X, out[i].Y + cp.position.Y}
	}
	return out
}

// updateBounds — convexPolygon.go:163-197 (the `else if` of the min/max scan is the reference's).
func (cp *ConvexPolygon) updateBounds() {
	t := cp.Transformed()
	if len(t) == 0 {
		return
	}
	tl, br := t[0], t[0]
	for _, p := range t {
		if p.X < tl.X {
			tl.X = p.X
		} else if p.X > br.X {
			br.X = p.X
		}
		if p.Y < tl.Y {
			tl.Y = p.Y
		} else if p.Y > br.Y {
			br.Y = p.Y
		}
	}
	cp.bounds = Bounds{tl, br}.Move(-cp.position.X, -cp.position.Y)
}

// Bounds — convexPolygon.go:158-161.
func (cp *ConvexPolygon) Bounds() Bounds { return cp.bounds.Move(cp.position.X, cp.position.Y) }
func (cp *ConvexPolygon) Center() Vector { return cp.Bounds().Center() } // convexPolygon.go:200-215

func (cp *ConvexPolygon) Rotation() float64 { return cp.rotation }
func (cp *ConvexPolygon) SetRotation(radians float64) {
	cp.rotation = radians
	cp.updateBounds()
	cp.edited()
}
func (cp *ConvexPolygon) Rotate(radians float64) { cp.SetRotation(cp.rotation + radians) }
func (cp *ConvexPolygon) Scale() Vector           { return cp.scale }
func (cp *ConvexPolygon) SetScale(x, y float64) {
	cp.scale = Vector{x, y}
	cp.updateBounds()
	cp.edited()
}

func (cp *ConvexPolygon) edited() {
	if cp.space != nil {
		cp.space.markEdited(&cp.ShapeBase)
	}
}

func (cp *ConvexPolygon) writeRow(sp *Space, slot uint32, isNew bool) {
	t := cp.transformedLocal()
	if isNew || uint32(len(t)) > cp.vertCap {
		cp.vertOff, cp.vertCap = sp.allocVerts(uint32(len(t))), uint32(len(t))
	}
	for k, v := range t {
		sp.verts[2*(int(cp.vertOff)+k)], sp.verts[2*(int(cp.vertOff)+k)+1] = v.X, v.Y
	}
	sp.pos[2*slot], sp.pos[2*slot+1] = cp.position.X, cp.position.Y
	sp.tagsBuf[slot] = uint64(*cp.tags)
	sp.radius[slot] = 0
	sp.laabb[4*slot], sp.laabb[4*slot+1], sp.laabb[4*slot+2], sp.laabb[4*slot+3] = cp.bounds.Min.X, cp.bounds.Min.Y, cp.bounds.Max.X, cp.bounds.Max.Y
	sp.vertOff[slot] = cp.vertOff
	m := uint32(C.RSV_META_POLYGON | C.RSV_META_TESTER)
	if cp.Closed {
		m |= C.RSV_META_CLOSED
	}
	sp.meta[slot] = m | uint32(len(t))<<C.RSV_META_NV_SHIFT
}

// NewRectangle — convexPolygon.go:616-632: clockwise from the top-left corner, origin at the centre.
func NewRectangle(x, y, w, h float64) *ConvexPolygon {
	hw, hh := w/2, h/2
	return NewConvexPolygon(x, y, []float64{-hw, -hh, hw, -hh, hw, hh, -hw, hh})
}

// NewRectangleFromTopLeft — convexPolygon.go:638-643.
func NewRectangleFromTopLeft(x, y, w, h float64) *ConvexPolygon {
	r := NewRectangle(x, y, w, h)
	r.Move(w/2, h/2)
	return r
}

// NewRectangleFromCorners — convexPolygon.go:647-667.
func NewRectangleFromCorners(x1, y1, x2, y2 float64) *ConvexPolygon {
	if x1 > x2 {
		x1, x2 = x2, x1
	}
	if y1 > y2 {
		y1, y2 = y2, y1
	}
	hw, hh := (x2-x1)/2, (y2-y1)/2
	return NewRectangle(x1+hw, y1+hh, x2-x1, y2-y1)
}

// NewLine — convexPolygon.go:671-684 (the two points end up mirrored about the midpoint; two points never close).
func NewLine(x1, y1, x2, y2 float64) *ConvexPolygon {
	cx, cy := x1+((x2-x1)/2), y1+((y2-y1)/2)
	return NewConvexPolygon(cx, cy, []float64{cx - x1, cy - y1, cx - x2, cy - y2})
}

// ShapeLineTestSettings — convexPolygon.go:304-315.
type ShapeLineTestSettings struct {
	StartOffset      Vector
	Vector           Vector
	TestAgainst      ShapeIterator
	OnIntersect      func(set IntersectionSet, index, count int) bool
	IncludeAllPoints bool
	Lines            []int
}

// ShapeLineTest — convexPolygon.go:322-415: one LineTest per selected vertex (all of them in ONE device batch), sets
// merged per OtherShape keeping the smallest MTV, callbacks in order of |MTV|^2. The reference iterates a Go map of
// vertices (random order); here vertices are taken in insertion order, which fixes the order of merged contacts.
func (cp *ConvexPolygon) ShapeLineTest(settings ShapeLineTestSettings) bool {
	var verts []Vector
	addv := func(v Vector) {
		for _, o := range verts {
			if o == v {
				return
			}
		}
		verts = append(verts, v)
	}
	t := cp.Transformed()
	if !settings.IncludeAllPoints {
		nl := len(t) - 1
		if cp.Closed && len(cp.Points) > 2 {
			nl = len(t)
		}
		for i := 0; i < nl; i++ {
			// `for lineIndex := range settings.Lines` ranges over the INDICES of the slice (convexPolygon.go:333-338):
			// line i is used iff i < len(Lines). Kept as in the reference.
			if len(settings.Lines) > 0 && i >= len(settings.Lines) {
				continue
			}
			start, end := t[i], t[(i+1)%len(t)]
			dir := end.Sub(start).Unit()
			normal := Vector{dir.Y, -dir.X}.Unit()
			if normal.Dot(settings.Vector) < 0.01 { // convexPolygon.go:347
				continue
			}
			v := dir.Scale(0.5).Invert() // convexPolygon.go:352
			addv(start.Sub(v))
			addv(end.Sub(v.Invert()))
		}
	} else {
		for _, p := range t {
			addv(p)
		}
	}
	vu := settings.Vector.Unit()
	var results []IntersectionSet
	for _, p := range verts {
		LineTest(LineTestSettings{
			Start:        p.Sub(vu.Add(settings.StartOffset)), // convexPolygon.go:368
			End:          p.Add(settings.Vector),
			TestAgainst:  settings.TestAgainst,
			callingShape: cp,
			OnIntersect: func(set IntersectionSet, index, max int) bool {
				for i := range results {
					if results[i].OtherShape == set.OtherShape {
						results[i].Intersections = append(results[i].Intersections, set.Intersections...)
						if set.MTV.MagnitudeSquared() < results[i].MTV.MagnitudeSquared() {
							results[i].MTV = set.MTV
						}
						return true
					}
				}
				results = append(results, set)
				return true
			},
		})
	}
	sort.Slice(results, func(i, j int) bool { return results[i].MTV.MagnitudeSquared() < results[j].MTV.MagnitudeSquared() })
	for i, r := range results {
		if settings.OnIntersect != nil && !settings.OnIntersect(r, i, len(results)) {
			break
		}
	}
	return len(results) > 0
}
