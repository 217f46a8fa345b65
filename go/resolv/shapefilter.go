package resolv

import (
	"reflect"
	"sort"
)

// ShapeIterator — shapefilter.go:10-14.
type ShapeIterator interface {
	ForEach(forEachFunc func(shape IShape) bool)
}

// CellSelection — cell.go:66-70. Cells live on the device: the selection is symbolic, either the cells around a shape
// grown by `margin` (SelectTouchingCells, shape.go:220-237) or an explicit cell rect (Space.FilterCells).
type CellSelection struct {
	space                      *Space
	owner                      IShape // excludeSelf
	margin                     int
	StartX, StartY, EndX, EndY int
	explicit                   bool
}

// FilterShapes — cell.go:73-83.
func (c CellSelection) FilterShapes() ShapeFilter { return ShapeFilter{sel: c} }

// ForEach — cell.go:86-121: the shapes registered in the selected cells, row-major, first occurrence only, without
// the owner. For a selection around a shape these are the records of that shape in the Space's cached device frame.
func (c CellSelection) ForEach(fn func(shape IShape) bool) {
	sp := c.space
	if sp == nil {
		return
	}
	if c.owner == nil {
		for _, s := range sp.shapesInCells(c.StartX, c.StartY, c.EndX, c.EndY) {
			fn(s)
		}
		return
	}
	sp.ensureFrame(c.margin)
	hits, _ := resultsView(sp.h)
	ob := c.owner.base()
	first, count := sp.testerFirst[ob.slot], sp.testerCount[ob.slot]
	others := make([]IShape, 0, count) // copied out first: fn may move shapes, which re-runs the frame
	for k := uint32(0); k < count; k++ {
		others = append(others, sp.shapes[hits[first+k].b])
	}
	// In the reference a false return only leaves the loop over the CURRENT cell's shapes (cell.go:108-110) and the
	// walk goes on with the next cell; the device list has no cell boundaries, so false simply moves on to the next shape.
	for _, s := range others {
		fn(s)
	}
}

// ShapeFilter — shapefilter.go:18-23.
type ShapeFilter struct {
	sel     CellSelection
	all     *Space // Space.FilterShapes(): every shape of the Space
	Filters []func(s IShape) bool
}

func (s ShapeFilter) accepts(shape IShape) bool {
	for _, f := range s.Filters {
		if !f(shape) {
			return false
		}
	}
	return true
}

// ForEach — shapefilter.go:26-43. As in the reference the iteration NEVER stops early: the inner callback returns true
// whatever forEachFunc returns (so First() yields the last match, shapefilter.go:125-134).
func (s ShapeFilter) ForEach(fn func(shape IShape) bool) {
	visit := func(shape IShape) bool {
		if s.accepts(shape) {
			fn(shape)
		}
		return true
	}
	if s.all != nil {
		s.all.ForEachShape(func(shape IShape, index, maxCount int) bool { return visit(shape) })
		return
	}
	s.sel.ForEach(visit)
}

func (s ShapeFilter) with(f func(IShape) bool) ShapeFilter {
	s.Filters = append(append([]func(IShape) bool(nil), s.Filters...), f)
	return s
}

func (s ShapeFilter) ByTags(tags Tags) ShapeFilter    { return s.with(func(o IShape) bool { return o.Tags().Has(tags) }) }
func (s ShapeFilter) NotByTags(tags Tags) ShapeFilter { return s.with(func(o IShape) bool { return !o.Tags().Has(tags) }) }
func (s ShapeFilter) ByFunc(f func(s IShape) bool) ShapeFilter { return s.with(f) }

// ByDistance — shapefilter.go:66-74.
func (s ShapeFilter) ByDistance(point Vector, min, max float64) ShapeFilter {
	return s.with(func(o IShape) bool {
		d := o.Position().Distance(point)
		return d > min && d < max
	})
}

// ByDataType — shapefilter.go:84-96.
func (s ShapeFilter) ByDataType(dataType reflect.Type) ShapeFilter {
	if dataType == nil {
		return s
	}
	return s.with(func(o IShape) bool {
		if o.Data() != nil {
			return reflect.TypeOf(o.Data()).Implements(dataType)
		}
		return false
	})
}

// Not — shapefilter.go:99-109.
func (s ShapeFilter) Not(shapes ...IShape) ShapeFilter {
	return s.with(func(o IShape) bool {
		for _, x := range shapes {
			if x == o {
				return false
			}
		}
		return true
	})
}

func (s ShapeFilter) Shapes() ShapeCollection {
	var out ShapeCollection
	s.ForEach(func(shape IShape) bool { out = append(out, shape); return true })
	return out
}

func (s ShapeFilter) First() IShape {
	var r IShape
	s.ForEach(func(shape IShape) bool { r = shape; return false })
	return r
}

func (s ShapeFilter) Last() IShape {
	var r IShape
	s.ForEach(func(shape IShape) bool { r = shape; return true })
	return r
}

func (s ShapeFilter) Count() int {
	n := 0
	s.ForEach(func(shape IShape) bool { n++; return true })
	return n
}

// ShapeCollection — shapefilter.go:178-215.
type ShapeCollection []IShape

func (s ShapeCollection) ForEach(fn func(shape IShape) bool) {
	for _, shape := range s {
		if !fn(shape) {
			return
		}
	}
}

func (s ShapeCollection) SortByDistance(point Vector) {
	sort.Slice(s, func(i, j int) bool {
		return s[i].Position().DistanceSquared(point) < s[j].Position().DistanceSquared(point)
	})
}
