package resolv

import "math"

// Vector is the reference's 2D vector (vector.go:24-27); only the operations the hot path's host side needs are here,
// with the reference's float64 operation order.
type Vector struct {
	X, Y float64
}

func NewVector(x, y float64) Vector       { return Vector{x, y} }
func (v Vector) Add(o Vector) Vector       { return Vector{v.X + o.X, v.Y + o.Y} }
func (v Vector) Sub(o Vector) Vector       { return Vector{v.X - o.X, v.Y - o.Y} }
func (v Vector) Scale(s float64) Vector    { return Vector{v.X * s, v.Y * s} }
func (v Vector) Invert() Vector            { return Vector{-v.X, -v.Y} }
func (v Vector) Dot(o Vector) float64      { return v.X*o.X + v.Y*o.Y }
func (v Vector) Magnitude() float64        { return math.Sqrt(v.X*v.X + v.Y*v.Y) }
func (v Vector) MagnitudeSquared() float64 { return v.X*v.X + v.Y*v.Y }
func (v Vector) Distance(o Vector) float64 { return v.Sub(o).Magnitude() }
func (v Vector) DistanceSquared(o Vector) float64 {
	d := v.Sub(o)
	return d.X*d.X + d.Y*d.Y
}

// Unit — vector.go:167-175: vectors shorter than 1e-8 or of length exactly 1 are returned unchanged.
func (v Vector) Unit() Vector {
	l := v.Magnitude()
	if l < 1e-8 || l == 1 {
		return v
	}
	return Vector{v.X / l, v.Y / l}
}

// Rotate — vector.go:246-252.
func (v Vector) Rotate(angle float64) Vector {
	s, c := math.Sin(angle), math.Cos(angle)
	return Vector{v.X*c - v.Y*s, v.X*s + v.Y*c}
}

// Reflect — vector.go:197-200.
func (v Vector) Reflect(normal Vector) Vector {
	n := normal.Unit()
	return v.Sub(n.Scale(2 * n.Dot(v)))
}

// ClampMagnitude — vector.go:122-127.
func (v Vector) ClampMagnitude(max float64) Vector {
	if v.Magnitude() > max {
		return v.Unit().Scale(max)
	}
	return v
}

// Bounds — utils.go:85-88.
type Bounds struct {
	Min, Max Vector
}

func (b Bounds) Center() Vector { return b.Min.Add(b.Max.Sub(b.Min).Scale(0.5)) } // utils.go:101-103
func (b Bounds) Move(x, y float64) Bounds {
	return Bounds{Vector{b.Min.X + x, b.Min.Y + y}, Vector{b.Max.X + x, b.Max.Y + y}}
}
