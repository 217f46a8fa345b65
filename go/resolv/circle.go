package resolv

/*
#include "resolv_b200.h"
*/
import "C"

// Circle — circle.go:4-17.
type Circle struct {
	ShapeBase
	radius float64
}

func NewCircle(x, y, radius float64) *Circle {
	c := &Circle{ShapeBase: newShapeBase(x, y), radius: radius}
	c.owner = c
	return c
}

func (c *Circle) Clone() IShape { // circle.go:20-30
	n := NewCircle(c.position.X, c.position.Y, c.radius)
	*n.tags = *c.tags
	n.data = c.data
	return n
}

func (c *Circle) Radius() float64 { return c.radius }

func (c *Circle) SetRadius(radius float64) {
	c.radius = radius
	if c.space != nil {
		c.space.markEdited(&c.ShapeBase)
	}
}

// Bounds — circle.go:33-39.
func (c *Circle) Bounds() Bounds {
	return Bounds{Vector{c.position.X - c.radius, c.position.Y - c.radius}, Vector{c.position.X + c.radius, c.position.Y + c.radius}}
}

func (c *Circle) writeRow(sp *Space, slot uint32, isNew bool) {
	sp.pos[2*slot], sp.pos[2*slot+1] = c.position.X, c.position.Y
	sp.tagsBuf[slot] = uint64(*c.tags)
	sp.radius[slot] = c.radius
	// (-r, -r, +r, +r): local + position == Circle.Bounds exactly (x + (-r) == x - r in IEEE arithmetic)
	sp.laabb[4*slot], sp.laabb[4*slot+1], sp.laabb[4*slot+2], sp.laabb[4*slot+3] = -c.radius, -c.radius, c.radius, c.radius
	sp.meta[slot] = C.RSV_META_TESTER
}
