// Package resolv is the drop-in for github.com/solarlune/resolv whose collision hot path — the Space cell grid
// (broadphase), the ConvexPolygon / Circle pair tests (narrowphase) and LineTest — runs in libresolv_b200.so on a
// B200 through cgo (include/resolv_b200.h). The exported API is the reference's; shape editing, filters and the
// replay of OnIntersect callbacks stay in Go, everything per-frame and data-parallel runs on the device.
//
// Select it with a module replace, as the reference's own examples do (examples/go.mod:25):
//
//	replace github.com/solarlune/resolv => ../resolv-b200/go/resolv
//
// There is no CPU fallback: NewSpace panics when the library cannot create a device space.
//
// The reference has no frame concept: every Shape.IntersectionTest walks the cells around the shape. Here a Space
// caches the result of ONE device frame (every shape tested against SelectTouchingCells(margin).FilterShapes(), all
// candidates recorded); any move marks it stale and the next IntersectionTest re-runs it. A game that moves all its
// shapes and then tests them pays one frame per tick (Space.Step does that explicitly); a game that interleaves moves
// and tests shape by shape (examples/worldBouncer.go) still gets the reference's results, one frame per tested shape.
//
// This package cannot be compiled in the image it was written in (no Go toolchain there); its logic is the Go
// rendering of resolv_b200/csrc/host/resolv.hpp, which is compiled and tested against the oracle on the GPU
// (tests/test_host_gpu.py). parity/go holds the harness that diffs it against the real reference.
package resolv
