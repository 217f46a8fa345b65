// Package world loads the text world files written by parity/worldfile.py into a real resolv.Space.
// The file format is specified there; every float travels as its IEEE-754 bit pattern (16 hex digits).
package world

import (
	"bufio"
	"fmt"
	"math"
	"os"
	"strconv"
	"strings"

	"github.com/solarlune/resolv"
)

// Shape is one shape of the world with its IntersectionTest settings.
type Shape struct {
	Shape    resolv.IShape
	Tester   bool
	UseAny   bool   // apply ByTags(AnyMask)
	AnyMask  uint64
	NoneMask uint64 // apply NotByTags(NoneMask) when non-zero
}

// World is a loaded world file.
type World struct {
	Name     string
	Space    *resolv.Space
	Margin   int
	Shapes   []Shape
	FramePos [][][2]float64 // positions of frames 1.. (SetPosition on every shape in index order)
	Rays     [][4]float64
	index    map[uint32]int
}

// Index maps a resolv shape back to its index in the world file.
func (w *World) Index(s resolv.IShape) int { return w.index[s.ID()] }

// Filter is the TestAgainst of tester a: SelectTouchingCells(margin).FilterShapes()[.ByTags][.NotByTags].
func (w *World) Filter(a int) resolv.ShapeFilter {
	rec := w.Shapes[a]
	filter := rec.Shape.SelectTouchingCells(w.Margin).FilterShapes()
	if rec.UseAny {
		filter = filter.ByTags(resolv.Tags(rec.AnyMask))
	}
	if rec.NoneMask != 0 {
		filter = filter.NotByTags(resolv.Tags(rec.NoneMask))
	}
	return filter
}

// SetFrame moves every shape to its position of frame f >= 1.
func (w *World) SetFrame(f int) {
	for i, p := range w.FramePos[f-1] {
		w.Shapes[i].Shape.SetPosition(p[0], p[1])
	}
}

type reader struct {
	sc   *bufio.Scanner
	path string
	err  error
}

func (r *reader) next() []string {
	for r.err == nil && r.sc.Scan() {
		t := strings.Fields(r.sc.Text())
		if len(t) > 0 && !strings.HasPrefix(t[0], "#") {
			return t
		}
	}
	if r.err == nil {
		r.err = fmt.Errorf("%s: unexpected end of file", r.path)
	}
	return make([]string, 32)
}

func (r *reader) expect(key string, fields int) []string {
	t := r.next()
	if r.err == nil && (t[0] != key || len(t) < fields) {
		r.err = fmt.Errorf("%s: expected %q with %d fields, got %q", r.path, key, fields, strings.Join(t, " "))
		return make([]string, 32)
	}
	return t
}

func (r *reader) f64(tok string) float64 {
	u, err := strconv.ParseUint(tok, 16, 64)
	if err != nil && r.err == nil {
		r.err = fmt.Errorf("%s: bad float bits %q", r.path, tok)
	}
	return math.Float64frombits(u)
}

func (r *reader) hex(tok string) uint64 {
	u, err := strconv.ParseUint(tok, 16, 64)
	if err != nil && r.err == nil {
		r.err = fmt.Errorf("%s: bad hex %q", r.path, tok)
	}
	return u
}

func (r *reader) int(tok string) int {
	v, err := strconv.Atoi(tok)
	if err != nil && r.err == nil {
		r.err = fmt.Errorf("%s: bad integer %q", r.path, tok)
	}
	return v
}

// Load reads a world file and builds the Space exactly as the CPU oracle does (oracle/resolv_oracle.cpp: orc_add_bulk):
// a polygon is created at pos0 (NewConvexPolygon -> AddPoints -> updateBounds caches its local AABB there), moved to
// its frame-0 position while still outside any Space, then added.
func Load(path string) (*World, error) {
	fh, err := os.Open(path)
	if err != nil {
		return nil, err
	}
	defer fh.Close()
	r := &reader{sc: bufio.NewScanner(fh), path: path}
	r.sc.Buffer(make([]byte, 1<<20), 1<<26)

	if t := r.expect("RESOLVWORLD", 2); r.err == nil && t[1] != "1" {
		return nil, fmt.Errorf("%s: unsupported world version %s", path, t[1])
	}
	w := &World{index: map[uint32]int{}}
	w.Name = r.expect("name", 2)[1]
	sp := r.expect("space", 5)
	if r.err != nil {
		return nil, r.err
	}
	w.Space = resolv.NewSpace(r.int(sp[1]), r.int(sp[2]), r.int(sp[3]), r.int(sp[4]))
	w.Margin = r.int(r.expect("margin", 2)[1])
	n := r.int(r.expect("shapes", 2)[1])
	if r.err != nil {
		return nil, r.err
	}
	w.Shapes = make([]Shape, n)
	for i := 0; i < n; i++ {
		t := r.expect("S", 14)
		if r.err != nil {
			return nil, r.err
		}
		pos0x, pos0y, posx, posy := r.f64(t[8]), r.f64(t[9]), r.f64(t[10]), r.f64(t[11])
		var s resolv.IShape
		switch t[1] {
		case "C":
			s = resolv.NewCircle(posx, posy, r.f64(t[12]))
		case "P":
			nv := r.int(t[13])
			if len(t) < 14+2*nv {
				return nil, fmt.Errorf("%s: shape %d: %d vertices announced, %d coordinates present", path, i, nv, len(t)-14)
			}
			pts := make([]float64, 2*nv)
			for k := range pts {
				pts[k] = r.f64(t[14+k])
			}
			p := resolv.NewConvexPolygon(pos0x, pos0y, pts)
			p.Closed = t[3] == "1"
			p.SetPosition(posx, posy)
			s = p
		default:
			return nil, fmt.Errorf("%s: shape %d: unknown kind %q", path, i, t[1])
		}
		s.Tags().Set(resolv.Tags(r.hex(t[5])))
		w.Space.Add(s)
		w.Shapes[i] = Shape{Shape: s, Tester: t[2] == "1", UseAny: t[4] == "1", AnyMask: r.hex(t[6]), NoneMask: r.hex(t[7])}
		w.index[s.ID()] = i
	}
	frames := r.int(r.expect("frames", 2)[1])
	for f := 1; f < frames && r.err == nil; f++ {
		if t := r.expect("pos", 2); r.int(t[1]) != f {
			return nil, fmt.Errorf("%s: frame block %s out of order", path, t[1])
		}
		pos := make([][2]float64, n)
		for i := range pos {
			t := r.next()
			pos[i] = [2]float64{r.f64(t[0]), r.f64(t[1])}
		}
		w.FramePos = append(w.FramePos, pos)
	}
	m := r.int(r.expect("rays", 2)[1])
	for k := 0; k < m && r.err == nil; k++ {
		t := r.expect("R", 5)
		w.Rays = append(w.Rays, [4]float64{r.f64(t[1]), r.f64(t[2]), r.f64(t[3]), r.f64(t[4])})
	}
	return w, r.err
}
