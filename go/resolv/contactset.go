package resolv

// Intersection — contactSet.go:8-11.
type Intersection struct {
	Point  Vector
	Normal Vector
}

// IntersectionSet — contactSet.go:15-20. Center is only set by LineTest.
type IntersectionSet struct {
	Intersections []Intersection
	Center        Vector
	MTV           Vector
	OtherShape    IShape
}

func (is IntersectionSet) IsEmpty() bool { return len(is.Intersections) == 0 }
