package resolv

import "sort"

// IShape — shape.go:9-50. The unexported methods keep the interface closed to this package, as in the reference.
type IShape interface {
	ID() uint32
	Clone() IShape
	Tags() *Tags
	Data() any
	SetData(data any)

	Position() Vector
	SetPosition(x, y float64)
	SetPositionVec(vec Vector)
	Move(x, y float64)
	MoveVec(vec Vector)

	Bounds() Bounds
	Space() *Space

	SelectTouchingCells(margin int) CellSelection
	IntersectionTest(settings IntersectionTestSettings) bool
	IsIntersecting(other IShape) bool
	Intersection(other IShape) IntersectionSet
	DistanceSquaredTo(other IShape) float64

	base() *ShapeBase
	writeRow(sp *Space, slot uint32, isNew bool)
}

var globalShapeID uint32

// ShapeBase — shape.go:54-62. slot is the shape's row in the device-resident SoA.
type ShapeBase struct {
	position Vector
	space    *Space
	tags     *Tags
	data     any
	owner    IShape
	id       uint32
	slot     uint32
}

func newShapeBase(x, y float64) ShapeBase {
	t := Tags(0)
	id := globalShapeID
	globalShapeID++
	return ShapeBase{position: Vector{x, y}, tags: &t, id: id}
}

func (s *ShapeBase) base() *ShapeBase      { return s }
func (s *ShapeBase) ID() uint32            { return s.id }
func (s *ShapeBase) Data() any             { return s.data }
func (s *ShapeBase) SetData(data any)      { s.data = data }
func (s *ShapeBase) Tags() *Tags           { return s.tags }
func (s *ShapeBase) Position() Vector      { return s.position }
func (s *ShapeBase) Space() *Space         { return s.space }
func (s *ShapeBase) MoveVec(v Vector)      { s.Move(v.X, v.Y) }
func (s *ShapeBase) SetPositionVec(v Vector) { s.SetPosition(v.X, v.Y) }

// Move / SetPosition — shape.go:98-119. update() no longer touches cells: it writes 16 bytes of staging and marks the
// frame stale (shape.go:239-242 is replaced by the per-frame grid rebuild on the device).
func (s *ShapeBase) Move(x, y float64) {
	s.position.X += x
	s.position.Y += y
	s.update()
}

func (s *ShapeBase) SetPosition(x, y float64) {
	s.position.X = x
	s.position.Y = y
	s.update()
}

func (s *ShapeBase) update() {
	if s.space != nil {
		s.space.markMoved(s)
	}
}

func (s *ShapeBase) DistanceSquaredTo(other IShape) float64 { // shape.go:179-181
	return s.position.DistanceSquared(other.Position())
}

// SelectTouchingCells — shape.go:220-237. The selection is symbolic: (owner, margin) is all the device needs.
func (s *ShapeBase) SelectTouchingCells(margin int) CellSelection {
	return CellSelection{space: s.space, owner: s.owner, margin: margin}
}

// IntersectionTestSettings — shape.go:250-258.
type IntersectionTestSettings struct {
	TestAgainst ShapeIterator
	OnIntersect func(set IntersectionSet) bool
}

type possibleIntersection struct {
	Shape    IShape
	Distance float64
	known    bool            // set was computed at the move epoch `valid` below
	set      IntersectionSet // copied out of the result mirrors: callbacks may overwrite them
}

// IntersectionTest — shape.go:273-316. Candidates come from the device frame (ShapeFilter over this shape's cell
// selection: the tester's records, one per candidate, in generation order) or from any other ShapeIterator; they are
// sorted with sort.Slice exactly as the reference sorts them, and tested in that order: from the frame's records while
// nothing has moved since the frame, through Space.testPairs (rsv_test_pairs) against the CURRENT positions as soon
// as a callback has moved something — the reference evaluates Intersection() candidate by candidate (shape.go:298).
func (s *ShapeBase) IntersectionTest(settings IntersectionTestSettings) bool {
	sp := s.space
	if sp == nil || settings.TestAgainst == nil {
		return false
	}
	var list []possibleIntersection
	valid := ^uint64(0)
	if sf, ok := settings.TestAgainst.(ShapeFilter); ok && sf.sel.owner != nil && sf.sel.space == sp {
		sp.ensureFrame(sf.sel.margin)
		own := sf.sel.owner.ID() == s.id
		hits, contacts := resultsView(sp.h)
		ob := sf.sel.owner.base()
		first, count := sp.testerFirst[ob.slot], sp.testerCount[ob.slot]
		for k := uint32(0); k < count; k++ {
			h := &hits[first+k]
			other := sp.shapes[h.b]
			if other == s.owner || !sf.accepts(other) {
				continue
			}
			p := possibleIntersection{Shape: other, Distance: other.DistanceSquaredTo(s.owner), known: own}
			if own {
				p.set = setFromHit(h, contacts, s.owner, other)
			}
			list = append(list, p)
		}
		if own {
			valid = sp.frameEpoch
		}
	} else {
		settings.TestAgainst.ForEach(func(other IShape) bool {
			if other == s.owner {
				return true
			}
			list = append(list, possibleIntersection{Shape: other, Distance: other.DistanceSquaredTo(s.owner)})
			return true
		})
	}
	sort.Slice(list, func(i, j int) bool { return list[i].Distance < list[j].Distance }) // shape.go:291-293
	collided := false
	for idx := range list {
		if sp.moveEpoch != valid || !list[idx].known {
			rest := make([][2]IShape, 0, len(list)-idx)
			for k := idx; k < len(list); k++ {
				rest = append(rest, [2]IShape{s.owner, list[k].Shape})
			}
			sets := sp.testPairs(rest)
			for k := idx; k < len(list); k++ {
				list[k].set, list[k].known = sets[k-idx], true
			}
			valid = sp.moveEpoch
		}
		result := list[idx].set
		if !result.IsEmpty() {
			collided = true
			if settings.OnIntersect != nil {
				if !settings.OnIntersect(result) {
					break
				}
			} else {
				break
			}
		}
	}
	return collided
}

// Intersection — circle.go:69-82, convexPolygon.go:288-300: one ordered pair through rsv_test_pairs.
func (s *ShapeBase) Intersection(other IShape) IntersectionSet {
	if s.space == nil || other.Space() != s.space {
		return IntersectionSet{}
	}
	return s.space.testPairs([][2]IShape{{s.owner, other}})[0]
}

func (s *ShapeBase) IsIntersecting(other IShape) bool { return !s.Intersection(other).IsEmpty() } // shape.go:245-247
