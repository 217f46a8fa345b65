// The true Go CPU baseline for bench.py's metric (SURVEY.md §8(d)): one frame = SetPosition on every shape + every
// tester's IntersectionTest with a callback returning true, on a world exported by
//
//	python parity/worldfile.py bench 250000        # writes parity/worlds/bench_cfg2_250000.world
//	RESOLV_WORLD=../../worlds/bench_cfg2_250000.world go test -bench Frame -benchtime 5x ./bench
//
// resolv keeps package-level scratch slices (shape.go:265, utils.go:255,272), so one process runs one goroutine.
package bench

import (
	"os"
	"testing"

	"github.com/solarlune/resolv"

	"resolvparity/world"
)

func BenchmarkFrame(b *testing.B) {
	path := os.Getenv("RESOLV_WORLD")
	if path == "" {
		b.Skip("set RESOLV_WORLD to a world file written by parity/worldfile.py")
	}
	w, err := world.Load(path)
	if err != nil {
		b.Fatal(err)
	}
	pairTests, hits := 0, 0
	for a := range w.Shapes { // candidates per frame, counted outside the timed region (the world does not move)
		if w.Shapes[a].Tester {
			w.Filter(a).ForEach(func(resolv.IShape) bool { pairTests++; return true })
		}
	}
	b.ResetTimer()
	for i := 0; i < b.N; i++ {
		for _, s := range w.Shapes {
			p := s.Shape.Position()
			s.Shape.SetPosition(p.X, p.Y) // the reference's update(): unregister + register every frame
		}
		for a := range w.Shapes {
			if !w.Shapes[a].Tester {
				continue
			}
			w.Shapes[a].Shape.IntersectionTest(resolv.IntersectionTestSettings{
				TestAgainst: w.Filter(a),
				OnIntersect: func(set resolv.IntersectionSet) bool { hits++; return true },
			})
		}
	}
	b.StopTimer()
	sec := b.Elapsed().Seconds()
	b.ReportMetric(float64(len(w.Shapes))*float64(b.N)/sec, "shapes/s")
	b.ReportMetric(float64(pairTests)*float64(b.N)/sec, "pair-tests/s")
	b.ReportMetric(float64(hits)/float64(b.N), "hits/frame")
}
