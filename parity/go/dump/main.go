// Command dump runs the REAL resolv package on a world file written by parity/worldfile.py and prints the dump text
// that tests/test_go_dump.py (or plain diff against parity/expected/*.dump) compares with the CPU oracle.
//
//	go run ./dump ../worlds/cfg1_bouncer_294.world > ../dumps/cfg1_bouncer_294.dump
//
// The harness is the two-phase frame of SURVEY.md §8(d): SetPosition on every shape in index order, then for every
// tester the candidate walk and IntersectionTest with a callback that always returns true.
package main

import (
	"bufio"
	"fmt"
	"math"
	"os"

	"github.com/solarlune/resolv"

	"resolvparity/world"
)

func hx(v float64) string { return fmt.Sprintf("%016x", math.Float64bits(v)) }

func printSet(out *bufio.Writer, w *world.World, set resolv.IntersectionSet, withCenter bool) {
	fmt.Fprintf(out, " %d %s %s", w.Index(set.OtherShape), hx(set.MTV.X), hx(set.MTV.Y))
	if withCenter {
		fmt.Fprintf(out, " %s %s", hx(set.Center.X), hx(set.Center.Y))
	}
	fmt.Fprintf(out, " %d", len(set.Intersections))
	for _, c := range set.Intersections {
		fmt.Fprintf(out, " %s %s %s %s", hx(c.Point.X), hx(c.Point.Y), hx(c.Normal.X), hx(c.Normal.Y))
	}
	fmt.Fprintln(out)
}

func main() {
	if len(os.Args) != 2 {
		fmt.Fprintln(os.Stderr, "usage: dump <file.world>")
		os.Exit(2)
	}
	w, err := world.Load(os.Args[1])
	if err != nil {
		fmt.Fprintln(os.Stderr, err)
		os.Exit(1)
	}
	out := bufio.NewWriterSize(os.Stdout, 1<<20)
	defer out.Flush()
	space := w.Space

	fmt.Fprintf(out, "RESOLVDUMP 1 %s\n", w.Name)
	for f := 0; f <= len(w.FramePos); f++ {
		if f > 0 {
			w.SetFrame(f)
		}
		fmt.Fprintf(out, "frame %d\n", f)
		for cy := 0; cy < space.HeightInCells(); cy++ {
			for cx := 0; cx < space.WidthInCells(); cx++ {
				c := space.Cell(cx, cy)
				if c == nil || len(c.Shapes) == 0 {
					continue
				}
				fmt.Fprintf(out, "cell %d %d", cy, cx)
				for _, s := range c.Shapes {
					fmt.Fprintf(out, " %d", w.Index(s))
				}
				fmt.Fprintln(out)
			}
		}
		for a := range w.Shapes {
			if !w.Shapes[a].Tester {
				continue
			}
			filter := w.Filter(a)
			var cands []int
			filter.ForEach(func(o resolv.IShape) bool {
				cands = append(cands, w.Index(o))
				return true
			})
			fmt.Fprintf(out, "test %d %d", a, len(cands))
			for _, b := range cands {
				fmt.Fprintf(out, " %d", b)
			}
			fmt.Fprintln(out)
			k := 0
			w.Shapes[a].Shape.IntersectionTest(resolv.IntersectionTestSettings{
				TestAgainst: filter,
				OnIntersect: func(set resolv.IntersectionSet) bool {
					fmt.Fprintf(out, "hit %d %d", a, k)
					printSet(out, w, set, false)
					k++
					return true
				},
			})
		}
	}

	for _, mode := range []string{"all", "rect"} {
		for r, ray := range w.Rays {
			start := resolv.NewVector(ray[0], ray[1])
			end := resolv.NewVector(ray[2], ray[3])
			var against resolv.ShapeIterator = space.FilterShapes()
			if mode == "rect" {
				// cells under the bounding box of the cast line (its start pulled back by castMargin, utils.go:278-280)
				vu := end.Sub(start).Unit()
				pulled := start.Sub(vu.Scale(0.01))
				against = space.FilterCells(resolv.Bounds{
					Min: resolv.NewVector(math.Min(pulled.X, end.X), math.Min(pulled.Y, end.Y)),
					Max: resolv.NewVector(math.Max(pulled.X, end.X), math.Max(pulled.Y, end.Y)),
				}).FilterShapes()
			}
			fmt.Fprintf(out, "ray %d %s\n", r, mode)
			ri := r
			resolv.LineTest(resolv.LineTestSettings{
				Start:       start,
				End:         end,
				TestAgainst: against,
				OnIntersect: func(set resolv.IntersectionSet, k, count int) bool {
					fmt.Fprintf(out, "rhit %d %d", ri, k)
					printSet(out, w, set, true)
					return true
				},
			})
		}
	}
}
