// rsv_math.cuh — float64 device primitives in the reference's exact operation order.
//
// Everything here must round exactly like Go's default amd64 code generation (no FMA contraction): the
// library is compiled with -fmad=false and, belt and braces, the products/sums that feed comparisons use the
// explicit round-to-nearest intrinsics. fp64 division and sqrt are IEEE-correct on sm_100a.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rsv {

struct V2 {
    double x, y;
};

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }

// Go math.Min / math.Max (NaN-propagating, Min(-0,+0) = -0, Max(-0,+0) = +0); used by
// Projection.Overlap (utils.go:75-77) and Bounds.Intersection (utils.go:160-164).
__device__ __forceinline__ double go_min(double x, double y) {
    if (isnan(x) || isnan(y)) {
        if ((isinf(x) && x < 0) || (isinf(y) && y < 0)) return -INFINITY;
        return NAN;
    }
    if (x == 0 && y == 0) return signbit(x) ? x : y;
    return x < y ? x : y;
}
__device__ __forceinline__ double go_max(double x, double y) {
    if (isnan(x) || isnan(y)) {
        if ((isinf(x) && x > 0) || (isinf(y) && y > 0)) return INFINITY;
        return NAN;
    }
    if (x == 0 && y == 0) return signbit(x) ? y : x;
    return x > y ? x : y;
}

__device__ __forceinline__ V2 vsub(V2 a, V2 b) { return V2{dsub(a.x, b.x), dsub(a.y, b.y)}; }       // vector.go:73-77
__device__ __forceinline__ V2 vadd(V2 a, V2 b) { return V2{dadd(a.x, b.x), dadd(a.y, b.y)}; }       // vector.go:54-58
__device__ __forceinline__ V2 vinv(V2 a) { return V2{-a.x, -a.y}; }                                 // vector.go:105-109
__device__ __forceinline__ V2 vscale(V2 a, double s) { return V2{dmul(a.x, s), dmul(a.y, s)}; }     // vector.go:273-277
__device__ __forceinline__ double vdot(V2 a, V2 b) { return dadd(dmul(a.x, b.x), dmul(a.y, b.y)); } // vector.go:287-289
__device__ __forceinline__ double vmag2(V2 a) { return dadd(dmul(a.x, a.x), dmul(a.y, a.y)); }      // vector.go:117-119
// Exact shortcut for axis-aligned vectors (rectangles, tiles, walls — most of a 2D game): when one component is
// +-0, x*x + 0 rounds to fl(x*x) and sqrt(fl(x*x)) == |x| exactly for every double whose square neither
// overflows nor falls into the subnormal range, so the IEEE sqrt (a ~20-instruction sequence on the fp64 pipe)
// can be skipped with bit-identical results. The tests run on the bit patterns (integer pipe): `axis_mag` returns
// true and |v| when v is axis-aligned with 2^-498 <= |v| < 2^498; everything else takes the generic path.
__device__ __forceinline__ bool axis_mag(V2 a, double& m, bool& along_x) {
    const uint32_t hx = (uint32_t)__double2hiint(a.x) & 0x7fffffffu, hy = (uint32_t)__double2hiint(a.y) & 0x7fffffffu;
    const bool yz = (hy | (uint32_t)__double2loint(a.y)) == 0u;   // a.y == +-0
    const bool xz = (hx | (uint32_t)__double2loint(a.x)) == 0u;
    along_x = yz;
    const uint32_t h = yz ? hx : hy;
    m = fabs(yz ? a.x : a.y);
    return (yz | xz) & (h - 0x20d00000u < 0x5f100000u - 0x20d00000u);
}

__device__ __forceinline__ double vmag(V2 a) {                                                      // vector.go:112-114
    double m;
    bool ax;
    if (axis_mag(a, m, ax)) return m;
    return sqrt(vmag2(a));
}
__device__ __forceinline__ double vdist2(V2 a, V2 b) { return vmag2(vsub(a, b)); }                  // vector.go:152-156
__device__ __forceinline__ V2 vunit(V2 a) {                                                         // vector.go:167-175
    double l;
    bool ax;
    if (axis_mag(a, l, ax)) {
        // l = |a| exactly; `l < 1e-8 || l == 1.0` on the bit pattern (l > 0 here, so the order of doubles is the order of bits)
        const unsigned long long lb = (unsigned long long)__double_as_longlong(l);
        if (lb < 0x3e45798ee2308c3aull || lb == 0x3ff0000000000000ull) return a;
        // x / |x| == +-1 and +-0 / l == +-0 exactly: no IEEE division needed
        return ax ? V2{copysign(1.0, a.x), a.y} : V2{a.x, copysign(1.0, a.y)};
    }
    l = sqrt(vmag2(a));
    if (l < 1e-8 || l == 1.0) return a;
    return V2{a.x / l, a.y / l};
}

// collidingLine.Normal (convexPolygon.go:705-712): Unit({v.Y, -v.X}) with v = Unit(End - Start).
__device__ __forceinline__ V2 edge_normal(V2 s, V2 e) {
    V2 v = vunit(vsub(e, s));
    return vunit(V2{v.y, -v.x});
}

// Bounds.IsIntersecting (utils.go:146-173) including the Bounds.IsEmpty typo (Max.Y - Min.X).
// a = (minx, miny, maxx, maxy) of the receiver, o = other. The predicate is symmetric in (a, o).
__device__ __forceinline__ bool bounds_gate(double aminx, double aminy, double amaxx, double amaxy, double ominx,
                                            double ominy, double omaxx, double omaxy) {
    if (omaxx < aminx || ominx > amaxx || omaxy < aminy || ominy > amaxy) return false;  // zero Bounds is "empty"
    double ovmaxx = go_min(omaxx, amaxx), ovminx = go_max(ominx, aminx);
    double ovmaxy = go_min(omaxy, amaxy);
    bool empty = (dsub(ovmaxx, ovminx) == 0) && (dsub(ovmaxy, ovminx) == 0);
    return !empty;
}

}  // namespace rsv
