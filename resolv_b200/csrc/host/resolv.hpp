// resolv.hpp — C++ host layer over libresolv_b200.so that mirrors resolv's Go API for the collision hot path.
//
// The reference is compiled Go and no Go toolchain exists in the build image, so the host side above the C ABI is
// written in C++ with the reference's names, argument meanings and error behaviour; the Go/cgo package a maintainer
// ships is the same logic. Shape editing, filters and callback replay live here (host); the grid
// broadphase, the pair tests and the line tests run in the CUDA library. There is no CPU fallback: NewSpace throws
// when the library cannot create a device space.
//
// Mirrored API (reference file:line):
//   NewSpace, Space.Add / Remove / Shapes / FilterShapes / FilterCells      space.go:16-66,100-110,181-195
//   NewCircle, NewConvexPolygon, NewRectangle, NewRectangleFromTopLeft, NewLine   circle.go:10-17, convexPolygon.go:28-46,616-684
//   Shape.Position / SetPosition / Move / Tags / Bounds / SelectTouchingCells     shape.go:98-119,220-237
//   CellSelection.FilterShapes, ShapeFilter.ByTags / NotByTags / ByFunc / Not / ForEach   cell.go:73-83, shapefilter.go:26-109
//   Shape.IntersectionTest(IntersectionTestSettings{TestAgainst, OnIntersect}), Intersection, IsIntersecting   shape.go:245-316
//   LineTest(LineTestSettings{Start, End, TestAgainst, OnIntersect})               utils.go:260-370
//   ConvexPolygon.ShapeLineTest(ShapeLineTestSettings{...})                        convexPolygon.go:304-415
//
// Semantics of the frame: a Space caches the results of one device frame (all shapes as testers, no device-side tag
// filter, margin of the last request). Any mutation marks it stale; the next IntersectionTest re-runs the frame.
// Filters are predicates on the other shape, so applying them to the compacted hit records on the host is
// equivalent to the reference applying them before the pair test. Callbacks run in the order Go's sort.Slice leaves
// the candidate list in (shape.go:291-293): detail::GoSortSlice below is pdqsort_func as shipped in the Go standard
// library (insertion sort up to 12 elements, pattern-defeating quicksort beyond), so exact distance ties among more
// than 12 candidates come out as the reference's do. Frames record EVERY candidate (RSV_FRAME_ALL_CANDIDATES): the
// sort needs the whole list, and when a callback moves a shape the remaining candidates are re-tested against the
// current positions (rsv_test_pairs), exactly as the reference evaluates Intersection() candidate by candidate.
#pragma once
#include <algorithm>
#include <any>
#include <typeinfo>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/resolv_b200.h"

namespace resolv {

struct Vector {  // vector.go:24-27
    double X = 0, Y = 0;
    Vector Add(Vector o) const { return {X + o.X, Y + o.Y}; }
    Vector Sub(Vector o) const { return {X - o.X, Y - o.Y}; }
    Vector Scale(double s) const { return {X * s, Y * s}; }
    Vector Invert() const { return {-X, -Y}; }
    double Dot(Vector o) const { return X * o.X + Y * o.Y; }
    double Magnitude() const { return std::sqrt(X * X + Y * Y); }
    double MagnitudeSquared() const { return X * X + Y * Y; }
    double DistanceSquared(Vector o) const { double x = X - o.X, y = Y - o.Y; return x * x + y * y; }
    Vector Unit() const {  // vector.go:167-175
        double l = Magnitude();
        if (l < 1e-8 || l == 1) return *this;
        return {X / l, Y / l};
    }
    Vector Rotate(double a) const { return {X * std::cos(a) - Y * std::sin(a), X * std::sin(a) + Y * std::cos(a)}; }  // vector.go:246-252
};

namespace detail {
// sort.Slice of the Go standard library (sort/zsortfunc.go, pdqsort_func) over an index range with caller-supplied
// less(i, j) / swap(i, j): same comparisons and swaps in the same order, hence the same permutation — including for
// equal keys, where the algorithm is not stable beyond 12 elements.
template <class Less, class Swap>
struct GoSorter {
    Less less;
    Swap swap;
    enum Hint { kUnknown, kIncreasing, kDecreasing };
    void insertionSort(int a, int b) {
        for (int i = a + 1; i < b; i++)
            for (int j = i; j > a && less(j, j - 1); j--) swap(j, j - 1);
    }
    void siftDown(int lo, int hi, int first) {
        int root = lo;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= hi) return;
            if (child + 1 < hi && less(first + child, first + child + 1)) child++;
            if (!less(first + root, first + child)) return;
            swap(first + root, first + child);
            root = child;
        }
    }
    void heapSort(int a, int b) {
        const int first = a, lo = 0, hi = b - a;
        for (int i = (hi - 1) / 2; i >= 0; i--) siftDown(i, hi, first);
        for (int i = hi - 1; i >= 0; i--) { swap(first, first + i); siftDown(lo, i, first); }
    }
    static int bitsLen(unsigned v) { int n = 0; while (v) { n++; v >>= 1; } return n; }
    void breakPatterns(int a, int b) {
        const int length = b - a;
        if (length >= 8) {
            uint64_t r = (uint64_t)length;                      // xorshift seeded with the length
            const unsigned modulus = 1u << bitsLen((unsigned)length);
            const int idx = a + (length / 4) * 2 - 1;
            for (int i = 0; i < 3; i++) {
                r ^= r << 13; r ^= r >> 7; r ^= r << 17;
                int other = (int)((unsigned)r & (modulus - 1));
                if (other >= length) other -= length;
                swap(idx - 1 + i, a + other);
            }
        }
    }
    void order2(int& a, int& b, int& swaps) { if (less(b, a)) { swaps++; std::swap(a, b); } }
    int median(int a, int b, int c, int& swaps) { order2(a, b, swaps); order2(b, c, swaps); order2(a, b, swaps); return b; }
    int medianAdjacent(int a, int& swaps) { return median(a - 1, a, a + 1, swaps); }
    int choosePivot(int a, int b, Hint& hint) {
        const int shortestNinther = 50, maxSwaps = 4 * 3;
        const int l = b - a;
        int swaps = 0, i = a + l / 4 * 1, j = a + l / 4 * 2, k = a + l / 4 * 3;
        if (l >= 8) {
            if (l >= shortestNinther) { i = medianAdjacent(i, swaps); j = medianAdjacent(j, swaps); k = medianAdjacent(k, swaps); }
            j = median(i, j, k, swaps);
        }
        hint = swaps == 0 ? kIncreasing : (swaps == maxSwaps ? kDecreasing : kUnknown);
        return j;
    }
    void reverseRange(int a, int b) { for (int i = a, j = b - 1; i < j; i++, j--) swap(i, j); }
    bool partialInsertionSort(int a, int b) {
        const int maxSteps = 5, shortestShifting = 50;
        int i = a + 1;
        for (int step = 0; step < maxSteps; step++) {
            while (i < b && !less(i, i - 1)) i++;
            if (i == b) return true;
            if (b - a < shortestShifting) return false;
            swap(i, i - 1);
            if (i - a >= 2)
                for (int j = i - 1; j >= 1; j--) { if (!less(j, j - 1)) break; swap(j, j - 1); }
            if (b - i >= 2)
                for (int j = i + 1; j < b; j++) { if (!less(j, j - 1)) break; swap(j, j - 1); }
        }
        return false;
    }
    int partitionEqual(int a, int b, int pivot) {
        swap(a, pivot);
        int i = a + 1, j = b - 1;
        for (;;) {
            while (i <= j && !less(a, i)) i++;
            while (i <= j && less(a, j)) j--;
            if (i > j) break;
            swap(i, j); i++; j--;
        }
        return i;
    }
    int partition(int a, int b, int pivot, bool& already) {
        swap(a, pivot);
        int i = a + 1, j = b - 1;
        while (i <= j && less(i, a)) i++;
        while (i <= j && !less(j, a)) j--;
        if (i > j) { swap(j, a); already = true; return j; }
        swap(i, j); i++; j--;
        for (;;) {
            while (i <= j && less(i, a)) i++;
            while (i <= j && !less(j, a)) j--;
            if (i > j) break;
            swap(i, j); i++; j--;
        }
        swap(j, a);
        already = false;
        return j;
    }
    void pdqsort(int a, int b, int limit) {
        const int maxInsertion = 12;
        bool wasBalanced = true, wasPartitioned = true;
        for (;;) {
            const int length = b - a;
            if (length <= maxInsertion) { insertionSort(a, b); return; }
            if (limit == 0) { heapSort(a, b); return; }
            if (!wasBalanced) { breakPatterns(a, b); limit--; }
            Hint hint;
            int pivot = choosePivot(a, b, hint);
            if (hint == kDecreasing) { reverseRange(a, b); pivot = (b - 1) - (pivot - a); hint = kIncreasing; }
            if (wasBalanced && wasPartitioned && hint == kIncreasing && partialInsertionSort(a, b)) return;
            if (a > 0 && !less(a - 1, pivot)) { a = partitionEqual(a, b, pivot); continue; }
            bool already = false;
            const int mid = partition(a, b, pivot, already);
            wasPartitioned = already;
            const int leftLen = mid - a, rightLen = b - mid, balanceThreshold = length / 8;
            if (leftLen < rightLen) { wasBalanced = leftLen >= balanceThreshold; pdqsort(a, mid, limit); a = mid + 1; }
            else { wasBalanced = rightLen >= balanceThreshold; pdqsort(mid + 1, b, limit); b = mid; }
        }
    }
};
// sort.Slice(v, func(i, j) bool { return key(v[i]) < key(v[j]) })
template <class T, class Key>
void GoSortSlice(std::vector<T>& v, Key key) {
    auto less = [&](int i, int j) { return key(v[(size_t)i]) < key(v[(size_t)j]); };
    auto swp = [&](int i, int j) { std::swap(v[(size_t)i], v[(size_t)j]); };
    GoSorter<decltype(less), decltype(swp)> s{less, swp};
    const int n = (int)v.size();
    s.pdqsort(0, n, GoSorter<decltype(less), decltype(swp)>::bitsLen((unsigned)n));
}
}  // namespace detail

using Tags = uint64_t;  // tags.go:8; Has == any bit in common (tags.go:29-31)
inline bool TagsHas(Tags t, Tags v) { return (t & v) > 0; }

struct Bounds { Vector Min, Max; };

struct Intersection { Vector Point, Normal; };  // contactSet.go:8-11
class Shape;
struct IntersectionSet {                         // contactSet.go:15-20
    std::vector<Intersection> Intersections;
    Vector Center, MTV;
    Shape* OtherShape = nullptr;
    bool IsEmpty() const { return Intersections.empty(); }
};

class Space;
class ShapeFilter;

struct IntersectionTestSettings {  // shape.go:250-258
    const ShapeFilter* TestAgainst = nullptr;
    std::function<bool(const IntersectionSet&)> OnIntersect;
};

class Shape {
public:
    virtual ~Shape() = default;
    uint32_t ID() const { return id_; }
    Tags& GetTags() { return tags_; }
    Vector Position() const { return pos_; }
    void SetPosition(double x, double y) { pos_ = {x, y}; updatePosition(); }   // shape.go:115-119
    void Move(double x, double y) { pos_.X += x; pos_.Y += y; updatePosition(); }  // shape.go:98-102
    void MoveVec(Vector v) { Move(v.X, v.Y); }
    Space* GetSpace() const { return space_; }
    void SetData(std::any d) { data_ = std::move(d); }      // shape.go:121-129
    const std::any& Data() const { return data_; }
    virtual Bounds GetBounds() const = 0;
    virtual bool IsCircle() const = 0;
    inline ShapeFilter SelectTouchingCells(int margin);                  // + .FilterShapes() folded in (cell.go:73-83)
    inline bool IntersectionTest(const IntersectionTestSettings& settings);
    inline IntersectionSet Intersection(Shape& other);                   // circle.go:69-82, convexPolygon.go:288-300
    bool IsIntersecting(Shape& other) { return !Intersection(other).IsEmpty(); }

protected:
    friend class Space;
    inline void update();          // shape.go:239-242: marks the staged row dirty (shape data changed)
    inline void updatePosition();  // same, when only the position changed: 16 bytes to upload, registrations of others kept
    Vector pos_;
    Tags tags_ = 0;
    std::any data_;
    uint32_t id_ = 0;
    Space* space_ = nullptr;
    uint32_t slot_ = 0;
};

class Circle : public Shape {  // circle.go:4-17
public:
    Circle(double x, double y, double r) : radius_(r) { pos_ = {x, y}; }
    double Radius() const { return radius_; }
    void SetRadius(double r) { radius_ = r; update(); }
    bool IsCircle() const override { return true; }
    Bounds GetBounds() const override { return {{pos_.X - radius_, pos_.Y - radius_}, {pos_.X + radius_, pos_.Y + radius_}}; }

private:
    double radius_;
};

struct ShapeLineTestSettings {  // convexPolygon.go:304-315
    Vector StartOffset, Vec;
    const ShapeFilter* TestAgainst = nullptr;
    std::function<bool(const IntersectionSet&, int, int)> OnIntersect;
    bool IncludeAllPoints = false;
    std::vector<int> Lines;   // convexPolygon.go:314, 329-340: matched against the INDICES of this slice (reference quirk kept)
};

class ConvexPolygon : public Shape {  // convexPolygon.go:12-46
public:
    std::vector<Vector> Points;
    bool Closed = true;
    ConvexPolygon(double x, double y, const std::vector<double>& pts) {
        pos_ = {x, y};
        if (pts.size() >= 4 && pts.size() % 2 == 0) {  // AddPoints' error cases leave the polygon empty (convexPolygon.go:92-99)
            for (size_t i = 0; i < pts.size(); i += 2) Points.push_back({pts[i], pts[i + 1]});
            updateBounds();
        }
    }
    bool IsCircle() const override { return false; }
    void SetRotation(double r) { rotation_ = r; updateBounds(); update(); }
    void SetScale(double x, double y) { scale_ = {x, y}; updateBounds(); update(); }
    std::vector<Vector> TransformedLocal() const {  // convexPolygon.go:146-150 without "+ position"
        std::vector<Vector> t;
        for (auto p : Points) {
            Vector q{p.X * scale_.X, p.Y * scale_.Y};
            if (rotation_ != 0) q = q.Rotate(-rotation_);
            t.push_back(q);
        }
        return t;
    }
    std::vector<Vector> Transformed() const {
        auto t = TransformedLocal();
        for (auto& p : t) p = {p.X + pos_.X, p.Y + pos_.Y};
        return t;
    }
    void updateBounds() {  // convexPolygon.go:163-197: cached LOCAL bounds
        auto t = Transformed();
        if (t.empty()) return;
        Vector tl = t[0], br = t[0];
        for (auto p : t) {
            if (p.X < tl.X) tl.X = p.X; else if (p.X > br.X) br.X = p.X;
            if (p.Y < tl.Y) tl.Y = p.Y; else if (p.Y > br.Y) br.Y = p.Y;
        }
        local_ = {{tl.X + -pos_.X, tl.Y + -pos_.Y}, {br.X + -pos_.X, br.Y + -pos_.Y}};
    }
    Bounds GetBounds() const override { return {{local_.Min.X + pos_.X, local_.Min.Y + pos_.Y}, {local_.Max.X + pos_.X, local_.Max.Y + pos_.Y}}; }
    const Bounds& LocalBounds() const { return local_; }
    inline bool ShapeLineTest(const ShapeLineTestSettings& settings);

private:
    Vector scale_{1, 1};
    double rotation_ = 0;
    Bounds local_;
};

inline std::unique_ptr<Circle> NewCircle(double x, double y, double r) { return std::make_unique<Circle>(x, y, r); }
inline std::unique_ptr<ConvexPolygon> NewConvexPolygon(double x, double y, const std::vector<double>& pts) {
    return std::make_unique<ConvexPolygon>(x, y, pts);
}
inline std::unique_ptr<ConvexPolygon> NewRectangle(double x, double y, double w, double h) {  // convexPolygon.go:616-632
    double hw = w / 2, hh = h / 2;
    return NewConvexPolygon(x, y, {-hw, -hh, hw, -hh, hw, hh, -hw, hh});
}
inline std::unique_ptr<ConvexPolygon> NewRectangleFromTopLeft(double x, double y, double w, double h) {  // convexPolygon.go:638-643
    auto r = NewRectangle(x, y, w, h);
    r->Move(w / 2, h / 2);
    return r;
}
inline std::unique_ptr<ConvexPolygon> NewLine(double x1, double y1, double x2, double y2) {  // convexPolygon.go:671-684
    double cx = x1 + ((x2 - x1) / 2), cy = y1 + ((y2 - y1) / 2);
    return NewConvexPolygon(cx, cy, {cx - x1, cy - y1, cx - x2, cy - y2});
}

// ShapeFilter over either a cell selection of one shape or the whole Space (shapefilter.go:17-109).
class ShapeFilter {
public:
    ShapeFilter ByTags(Tags t) const { auto f = *this; f.filters_.push_back([t](Shape& s) { return TagsHas(s.GetTags(), t); }); return f; }
    ShapeFilter NotByTags(Tags t) const { auto f = *this; f.filters_.push_back([t](Shape& s) { return !TagsHas(s.GetTags(), t); }); return f; }
    ShapeFilter ByFunc(std::function<bool(Shape&)> fn) const { auto f = *this; f.filters_.push_back(std::move(fn)); return f; }
    ShapeFilter Not(Shape* x) const { auto f = *this; f.filters_.push_back([x](Shape& s) { return &s != x; }); return f; }
    // shapefilter.go:66-74: strictly between min and max away from the point
    ShapeFilter ByDistance(Vector point, double min, double max) const {
        auto f = *this;
        f.filters_.push_back([=](Shape& s) { Vector p = s.Position(); double x = p.X - point.X, y = p.Y - point.Y; double d = std::sqrt(x * x + y * y); return d > min && d < max; });
        return f;
    }
    // shapefilter.go:84-96: shapes whose Data() is of the given type (Go tests interface implementation by reflection;
    // the C++ counterpart is the dynamic type held by the std::any). Shapes without data never pass.
    ShapeFilter ByDataType(const std::type_info& t) const {
        auto f = *this;
        const std::type_info* tp = &t;
        f.filters_.push_back([tp](Shape& s) { return s.Data().has_value() && s.Data().type() == *tp; });
        return f;
    }
    template <class T> ShapeFilter ByDataType() const { return ByDataType(typeid(T)); }
    bool Accepts(Shape& s) const {
        for (auto& f : filters_) if (!f(s)) return false;
        return true;
    }
    Space* space = nullptr;
    Shape* owner = nullptr;  // SelectTouchingCells: the tester (excludeSelf); nullptr: Space.FilterShapes()
    int margin = 1;

private:
    std::vector<std::function<bool(Shape&)>> filters_;
};

struct LineTestSettings {  // utils.go:260-270
    Vector Start, End;
    const ShapeFilter* TestAgainst = nullptr;
    std::function<bool(const IntersectionSet&, int, int)> OnIntersect;
};

class Space {
public:
    // NewSpace — space.go:16-30
    Space(int spaceWidth, int spaceHeight, int cellWidth, int cellHeight, uint32_t maxShapes = 1u << 16, uint32_t maxVerts = 1u << 19) {
        rsv_config cfg{};
        cfg.space_w = spaceWidth; cfg.space_h = spaceHeight; cfg.cell_w = cellWidth; cfg.cell_h = cellHeight;
        cfg.n_spaces = 1; cfg.max_shapes = maxShapes; cfg.max_verts = maxVerts;
        if (rsv_create(&cfg, &h_) != RSV_OK) throw std::runtime_error(std::string("resolv: ") + rsv_last_error(nullptr));
        size_t n;
        pos_ = (double*)rsv_staging(h_, RSV_BUF_POS, &n);
        laabb_ = (double*)rsv_staging(h_, RSV_BUF_LOCAL_AABB, &n);
        radius_ = (double*)rsv_staging(h_, RSV_BUF_RADIUS, &n);
        tags_ = (uint64_t*)rsv_staging(h_, RSV_BUF_TAGS, &n);
        meta_ = (uint32_t*)rsv_staging(h_, RSV_BUF_META, &n);
        voff_ = (uint32_t*)rsv_staging(h_, RSV_BUF_VERT_OFF, &n);
        verts_ = (double*)rsv_staging(h_, RSV_BUF_VERTS, &n);
        maxShapes_ = maxShapes; maxVerts_ = maxVerts;
    }
    ~Space() { rsv_destroy(h_); }
    Space(const Space&) = delete;

    // Add — space.go:33-46
    void Add(Shape* s) {
        if (shapes_.size() >= maxShapes_) throw std::runtime_error("resolv: Space capacity exceeded");
        s->space_ = this; s->id_ = nextId_++; s->slot_ = (uint32_t)shapes_.size();
        shapes_.push_back(s);
        moved_.push_back(1); still_.push_back(0);
        writeRow(s, true);
        rsv_set_counts(h_, (uint32_t)shapes_.size(), nVerts_);
        stale_ = true; fullCommit_ = true;
    }
    // Remove — space.go:50-66: the slot stays (flagged absent) so that other slots keep their numbers.
    void Remove(Shape* s) {
        if (s->space_ != this) return;
        meta_[s->slot_] |= RSV_META_ABSENT;
        removed_.push_back(s->slot_);
        stale_ = true; fullCommit_ = true;
    }
    std::vector<Shape*> Shapes() const {
        std::vector<Shape*> out;
        for (auto* s : shapes_) if (!(meta_[s->slot_] & RSV_META_ABSENT)) out.push_back(s);
        return out;
    }
    ShapeFilter FilterShapes() { ShapeFilter f; f.space = this; return f; }  // space.go:106-110

    // One device frame for all shapes (additive API). Frames are incremental (RSV_FRAME_INCREMENTAL):
    // shapes that stood still for kStillFrames Steps are flagged RSV_META_STATIC and keep their cell registrations on
    // the device, like shapes whose update() never runs keep theirs in the reference (shape.go:239-242).
    void Step(int margin) {
        prepare();
        // every candidate gets a record (rank = position in possibleIntersections before the sort, shape.go:277-289)
        const uint32_t flags = RSV_FRAME_DOWNLOAD | RSV_FRAME_TESTER_TABLE | RSV_FRAME_NO_TAG_FILTERS | RSV_FRAME_INCREMENTAL | RSV_FRAME_ALL_CANDIDATES;
        rsv_counts c{};
        int32_t rc = rsv_frame(h_, margin, flags, &c);
        for (int tries = 0; rc == RSV_ERR_CAPACITY && tries < 3; tries++) {
            rsv_reserve(h_, (uint32_t)(c.n_entries * 5 / 4 + 1024), (uint32_t)(c.n_gated * 5 / 4 + 1024), (uint32_t)(c.n_contacts * 5 / 4 + 1024));
            rc = rsv_frame(h_, margin, flags, &c);
        }
        if (rc != RSV_OK) throw std::runtime_error(std::string("resolv: ") + rsv_last_error(h_));
        first_.resize(shapes_.size()); count_.resize(shapes_.size());
        rsv_tester_table(h_, first_.data(), count_.data(), shapes_.size());
        rsv_results_view(h_, &res_);
        margin_ = margin; stale_ = false; counts_ = c; frameEpoch_ = moveEpoch_;
    }
    // (static registrations cached on the device, shapes re-registered per frame) — see rsv_static_info
    std::pair<uint32_t, uint32_t> StaticInfo() { uint32_t a = 0, b = 0; rsv_static_info(h_, &a, &b); return {a, b}; }
    const rsv_counts& LastCounts() const { return counts_; }
    rsv_space* Handle() { return h_; }

private:
    friend class Shape;
    friend class ConvexPolygon;
    friend bool LineTest(const LineTestSettings&);
    static constexpr uint32_t kStillFrames = 2, kPromoteEvery = 4;
    void markDirty(Shape* s) { writeRow(s, false); moved_[s->slot_] = 1; stale_ = true; moveEpoch_++; }
    void markMoved(Shape* s) {
        const uint32_t i = s->slot_;
        pos_[2 * i] = s->pos_.X; pos_[2 * i + 1] = s->pos_.Y;
        moved_[i] = 1;
        if (meta_[i] & RSV_META_STATIC) { meta_[i] &= ~RSV_META_STATIC; metaDirty_ = true; }   // demotion is immediate
        posLo_ = std::min(posLo_, i); posHi_ = std::max(posHi_, i + 1);
        stale_ = true; moveEpoch_++;
    }
    // Static / dynamic classification + upload of what changed; runs before every device frame.
    void prepare() {
        const uint32_t n = (uint32_t)shapes_.size();
        uint32_t pending = 0;
        bool tagsDirty = false;
        for (uint32_t i = 0; i < n; i++) {
            if (moved_[i]) still_[i] = 0;
            else if (still_[i] < 0xffffffffu) still_[i]++;
            moved_[i] = 0;
            if (still_[i] >= kStillFrames && !(meta_[i] & (RSV_META_STATIC | RSV_META_ABSENT))) pending++;
            if (tags_[i] != shapes_[i]->tags_) { tags_[i] = shapes_[i]->tags_; tagsDirty = true; }   // Tags() hands out a mutable reference
        }
        // promotions rewrite META (which drops the device's static cache): batch them
        if (pending && (metaDirty_ || fullCommit_ || frameNo_ % kPromoteEvery == 0)) {
            for (uint32_t i = 0; i < n; i++)
                if (still_[i] >= kStillFrames && !(meta_[i] & RSV_META_ABSENT)) meta_[i] |= RSV_META_STATIC;
            metaDirty_ = true;
        }
        frameNo_++;
        upload(tagsDirty);
    }
    // Uploads what changed since the last upload (whole rows after Add / shape edits, otherwise the dirty position range).
    void upload(bool tagsDirty = false) {
        const uint32_t n = (uint32_t)shapes_.size();
        if (fullCommit_) {
            rsv_commit(h_, (1u << RSV_BUF_COUNT) - 1, 0, n);
        } else {
            uint32_t mask = (metaDirty_ ? 1u << RSV_BUF_META : 0u) | (tagsDirty ? 1u << RSV_BUF_TAGS : 0u);
            if (mask) rsv_commit(h_, mask | (posHi_ > posLo_ ? 1u << RSV_BUF_POS : 0u), 0, n);
            else if (posHi_ > posLo_) rsv_commit(h_, 1u << RSV_BUF_POS, posLo_, posHi_);
        }
        fullCommit_ = false; metaDirty_ = false;
        posLo_ = 0xffffffffu; posHi_ = 0;
    }
    // A.Intersection(B) for explicit pairs against the CURRENT positions (rsv_test_pairs); one set per pair, in order.
    std::vector<IntersectionSet> testPairs(const std::vector<std::pair<Shape*, Shape*>>& pairs) {
        std::vector<IntersectionSet> out(pairs.size());
        if (pairs.empty()) return out;
        upload();
        size_t nb = 0;
        uint32_t* list = (uint32_t*)rsv_pair_staging(h_, (uint32_t)pairs.size(), &nb);
        if (!list) throw std::runtime_error("resolv: rsv_pair_staging failed");
        for (size_t k = 0; k < pairs.size(); k++) { list[2 * k] = pairs[k].first->slot_; list[2 * k + 1] = pairs[k].second->slot_; }
        rsv_counts c{};
        int32_t rc = rsv_test_pairs(h_, (uint32_t)pairs.size(), &c);
        for (int tries = 0; rc == RSV_ERR_CAPACITY && tries < 3; tries++) {
            rsv_reserve(h_, 0, (uint32_t)std::max<uint64_t>(c.n_gated, pairs.size()) * 5 / 4 + 1024, (uint32_t)(c.n_contacts * 5 / 4 + 1024));
            rc = rsv_test_pairs(h_, (uint32_t)pairs.size(), &c);
        }
        if (rc != RSV_OK) throw std::runtime_error(std::string("resolv: ") + rsv_last_error(h_));
        // (the result views now hold the pair records: the frame's views are gone, so is the tester table)
        stale_ = true;
        rsv_results r{};
        rsv_results_view(h_, &r);
        for (size_t k = 0; k < pairs.size() && k < r.n_hit_records; k++) fillSet(out[k], r.hits[k], r.contacts, pairs[k].first, pairs[k].second);
        return out;
    }
    static void fillSet(IntersectionSet& set, const rsv_hit& h, const rsv_contact* contacts, Shape* tester, Shape* other);
    void writeRow(Shape* s, bool isNew) {
        uint32_t i = s->slot_;
        pos_[2 * i] = s->pos_.X; pos_[2 * i + 1] = s->pos_.Y;
        tags_[i] = s->tags_;
        if (s->IsCircle()) {
            double r = static_cast<Circle*>(s)->Radius();
            radius_[i] = r;
            laabb_[4 * i] = -r; laabb_[4 * i + 1] = -r; laabb_[4 * i + 2] = r; laabb_[4 * i + 3] = r;
            meta_[i] = RSV_META_TESTER;
        } else {
            auto* p = static_cast<ConvexPolygon*>(s);
            auto t = p->TransformedLocal();
            if (isNew) { voff_[i] = nVerts_; nVerts_ += (uint32_t)t.size(); if (nVerts_ > maxVerts_) throw std::runtime_error("resolv: vertex pool exhausted"); }
            for (size_t k = 0; k < t.size(); k++) { verts_[2 * (voff_[i] + k)] = t[k].X; verts_[2 * (voff_[i] + k) + 1] = t[k].Y; }
            const Bounds& lb = p->LocalBounds();
            laabb_[4 * i] = lb.Min.X; laabb_[4 * i + 1] = lb.Min.Y; laabb_[4 * i + 2] = lb.Max.X; laabb_[4 * i + 3] = lb.Max.Y;
            radius_[i] = 0;
            meta_[i] = RSV_META_POLYGON | RSV_META_TESTER | (p->Closed ? RSV_META_CLOSED : 0) | ((uint32_t)t.size() << RSV_META_NV_SHIFT);
        }
        fullCommit_ = true;
    }
    void ensureFrame(int margin) { if (stale_ || margin != margin_) Step(margin); }

    rsv_space* h_ = nullptr;
    std::vector<Shape*> shapes_;
    std::vector<uint32_t> removed_;
    double *pos_, *laabb_, *radius_, *verts_;
    uint64_t* tags_;
    uint32_t *meta_, *voff_;
    uint32_t nVerts_ = 0, nextId_ = 0, maxShapes_, maxVerts_;
    bool stale_ = true, fullCommit_ = true, metaDirty_ = false;
    int margin_ = -1;
    std::vector<uint8_t> moved_;     // update() ran since the last device frame
    std::vector<uint32_t> still_;    // device frames since the shape last moved
    uint32_t posLo_ = 0xffffffffu, posHi_ = 0, frameNo_ = 0;
    uint64_t moveEpoch_ = 0, frameEpoch_ = 0;   // bumped by every shape move / edit; value at the last device frame
    std::vector<uint32_t> first_, count_;
    rsv_results res_{};
    rsv_counts counts_{};
};

// One IntersectionSet from a device record. The device sorts the contacts of a set as the reference does — except sets
// of more than 12 contacts of the two kinds the reference sorts (convexConvexTest by distance to the tester's Center,
// shape.go:467-469; convexCircleTest by distance to the circle, shape.go:385-387): sort.Slice is pdqsort_func there, so
// they arrive in generation order and are sorted here with the port of Go's sort.
inline void Space::fillSet(IntersectionSet& set, const rsv_hit& h, const rsv_contact* contacts, Shape* tester, Shape* other) {
    const uint32_t n = h.count_rank & 0x7f;
    if (n == 0) return;
    set.OtherShape = other;
    set.MTV = {h.mtv_x, h.mtv_y};
    for (uint32_t i = 0; i < n; i++) {
        const rsv_contact& q = contacts[h.first_contact + i];
        set.Intersections.push_back({{q.px, q.py}, {q.nx, q.ny}});
    }
    if (n > 12 && tester && !tester->IsCircle()) {
        Vector c = other->pos_;                                   // convexCircleTest: circle.position
        if (!other->IsCircle()) {                                 // convexConvexTest: convexA.Center() == Bounds().Center()
            const Bounds b = tester->GetBounds();
            c = {b.Min.X + (b.Max.X - b.Min.X) * 0.5, b.Min.Y + (b.Max.Y - b.Min.Y) * 0.5};
        }
        detail::GoSortSlice(set.Intersections, [&](const Intersection& k) { return k.Point.DistanceSquared(c); });
    }
}

inline void Shape::update() { if (space_) space_->markDirty(this); }
inline void Shape::updatePosition() { if (space_) space_->markMoved(this); }

inline ShapeFilter Shape::SelectTouchingCells(int margin) {
    ShapeFilter f;
    f.space = space_; f.owner = this; f.margin = margin;
    return f;
}

// shape.go:273-316, replayed from the compacted device buffer.
inline bool Shape::IntersectionTest(const IntersectionTestSettings& st) {
    if (!space_ || !st.TestAgainst || !st.TestAgainst->space) return false;
    Space& sp = *space_;
    const ShapeFilter& ta = *st.TestAgainst;
    // ---- possibleIntersections (shape.go:277-289): every shape the iterator yields except the tester, with its distance
    struct Cand { double d; Shape* other; bool known; IntersectionSet set; };
    std::vector<Cand> list;
    uint64_t validEpoch = ~0ull;   // move epoch the `set`s of the list were computed at (~0: none computed yet)
    if (ta.owner) {
        sp.ensureFrame(ta.margin);
        // the iterator is owner.SelectTouchingCells(margin).FilterShapes(): the owner's records, in generation order
        const uint32_t os = ta.owner->slot_;
        const bool own = ta.owner == this;
        for (uint32_t k = 0; k < sp.count_[os]; k++) {
            const rsv_hit& h = sp.res_.hits[sp.first_[os] + k];
            Shape* other = sp.shapes_[h.b];
            if (other == this || !ta.Accepts(*other)) continue;
            Cand c{other->pos_.DistanceSquared(pos_), other, own, {}};   // shape.go:179-181
            if (own) Space::fillSet(c.set, h, sp.res_.contacts, this, other);   // copied out: callbacks may invalidate the views
            list.push_back(std::move(c));
        }
        if (own) validEpoch = sp.frameEpoch_;
    } else {
        // space.FilterShapes(): every shape of the Space in Add order (space.go:100-110)
        for (Shape* other : sp.Shapes()) {
            if (other == this || !ta.Accepts(*other)) continue;
            list.push_back(Cand{other->pos_.DistanceSquared(pos_), other, false, {}});
        }
    }
    detail::GoSortSlice(list, [](const Cand& c) { return c.d; });   // shape.go:291-293
    bool collided = false;
    for (size_t idx = 0; idx < list.size(); idx++) {
        if (sp.moveEpoch_ != validEpoch || !list[idx].known) {
            // something moved since the sets were computed (the usual MoveVec(MTV) inside OnIntersect), or they never
            // were: the reference calls Intersection() on the CURRENT world for every remaining candidate (shape.go:298)
            std::vector<std::pair<Shape*, Shape*>> pairs;
            for (size_t k = idx; k < list.size(); k++) pairs.push_back({this, list[k].other});
            auto sets = sp.testPairs(pairs);
            for (size_t k = idx; k < list.size(); k++) { list[k].set = std::move(sets[k - idx]); list[k].known = true; }
            validEpoch = sp.moveEpoch_;
        }
        if (list[idx].set.IsEmpty()) continue;
        collided = true;
        if (!st.OnIntersect || !st.OnIntersect(list[idx].set)) break;
    }
    return collided;
}

// A single ordered pair (circle.go:69-82, convexPolygon.go:288-300) through rsv_test_pairs.
inline IntersectionSet Shape::Intersection(Shape& other) {
    if (!space_ || other.space_ != space_) return {};
    return space_->testPairs({{this, &other}})[0];
}

// utils.go:276-370 through rsv_linetest. TestAgainst built from SelectTouchingCells / FilterShapes: the device
// iterator is space.FilterCells(bounds of the cast line) (RECT), the host applies the filter chain and excludeSelf.
inline bool LineTest(const LineTestSettings& st) {
    if (!st.TestAgainst || !st.TestAgainst->space) return false;
    Space& sp = *st.TestAgainst->space;
    if (sp.stale_) { sp.prepare(); rsv_counts c{}; rsv_frame(sp.h_, 1, RSV_FRAME_GRID_ONLY | RSV_FRAME_INCREMENTAL, &c); }
    size_t n;
    double* seg = (double*)rsv_ray_staging(sp.h_, 0, &n);
    uint64_t* msk = (uint64_t*)rsv_ray_staging(sp.h_, 1, &n);
    seg[0] = st.Start.X; seg[1] = st.Start.Y; seg[2] = st.End.X; seg[3] = st.End.Y;
    msk[0] = msk[1] = msk[2] = 0;
    rsv_ray_counts rc{};
    if (rsv_linetest(sp.h_, 1, RSV_FRAME_DOWNLOAD, &rc) != RSV_OK) throw std::runtime_error(std::string("resolv: ") + rsv_last_error(sp.h_));
    rsv_ray_results rr{};
    rsv_ray_results_view(sp.h_, &rr);
    struct Cand { double key; uint32_t rank; const rsv_ray_hit* h; };
    std::vector<Cand> list;
    for (uint64_t k = 0; k < rr.n_hits; k++) {
        const rsv_ray_hit* h = rr.hits + k;
        Shape* other = sp.shapes_[h->b];
        if (other == st.TestAgainst->owner || !st.TestAgainst->Accepts(*other)) continue;
        list.push_back({h->key, h->count_rank >> 8, h});
    }
    // records come in candidate generation order; utils.go:355-357 sorts them with sort.Slice
    detail::GoSortSlice(list, [](const Cand& c) { return c.key; });
    int idx = 0;
    for (auto& c : list) {
        IntersectionSet set;
        set.OtherShape = sp.shapes_[c.h->b];
        set.MTV = {c.h->mtv_x, c.h->mtv_y};
        set.Center = {c.h->center_x, c.h->center_y};
        for (uint32_t i = 0; i < (c.h->count_rank & 0xff); i++) {
            const rsv_contact& k = rr.contacts[c.h->first_contact + i];
            set.Intersections.push_back({{k.px, k.py}, {k.nx, k.ny}});
        }
        if (st.OnIntersect && !st.OnIntersect(set, idx, (int)list.size())) break;
        idx++;
    }
    return !list.empty();
}

// convexPolygon.go:322-415: one LineTest per leading-edge vertex (insertion order; the reference iterates a Go map),
// merged per OtherShape keeping the smallest-|MTV|, replayed in order of |MTV|^2.
inline bool ConvexPolygon::ShapeLineTest(const ShapeLineTestSettings& st) {
    std::vector<Vector> verts;
    auto addv = [&](Vector v) {
        for (auto& o : verts) if (o.X == v.X && o.Y == v.Y) return;
        verts.push_back(v);
    };
    auto t = Transformed();
    if (!st.IncludeAllPoints) {
        size_t nl = (Closed && Points.size() > 2) ? t.size() : (t.empty() ? 0 : t.size() - 1);
        for (size_t i = 0; i < nl; i++) {
            // convexPolygon.go:329-340: `for lineIndex := range settings.Lines` ranges over the slice's INDICES, so line i
            // is cast from iff i < len(Lines) — whatever the slice holds (kept as in the reference)
            if (!st.Lines.empty() && i >= st.Lines.size()) continue;
            Vector s = t[i], e = t[(i + 1) % t.size()];
            Vector dir = e.Sub(s).Unit();
            Vector nrm = Vector{dir.Y, -dir.X}.Unit();
            if (nrm.Dot(st.Vec) < 0.01) continue;          // convexPolygon.go:347
            Vector v = dir.Scale(0.5).Invert();             // convexPolygon.go:352
            addv(s.Sub(v));
            addv(e.Sub(v.Invert()));
        }
    } else {
        for (auto p : t) addv(p);
    }
    Vector vu = st.Vec.Unit();
    std::vector<IntersectionSet> results;
    ShapeFilter self = st.TestAgainst ? *st.TestAgainst : ShapeFilter{};
    self.owner = this;  // callingShape (utils.go:290-292)
    for (auto p : verts) {
        LineTestSettings lt;
        lt.Start = p.Sub(vu.Add(st.StartOffset));           // convexPolygon.go:368
        lt.End = p.Add(st.Vec);
        lt.TestAgainst = &self;
        lt.OnIntersect = [&](const IntersectionSet& set, int, int) {
            for (auto& r : results)
                if (r.OtherShape == set.OtherShape) {
                    r.Intersections.insert(r.Intersections.end(), set.Intersections.begin(), set.Intersections.end());
                    if (set.MTV.MagnitudeSquared() < r.MTV.MagnitudeSquared()) r.MTV = set.MTV;
                    return true;
                }
            results.push_back(set);
            return true;
        };
        LineTest(lt);
    }
    detail::GoSortSlice(results, [](const IntersectionSet& r) { return r.MTV.MagnitudeSquared(); });   // convexPolygon.go:398-400
    for (size_t i = 0; i < results.size(); i++)
        if (st.OnIntersect && !st.OnIntersect(results[i], (int)i, (int)results.size())) break;
    return !results.empty();
}

inline std::unique_ptr<Space> NewSpace(int w, int h, int cw, int ch) { return std::make_unique<Space>(w, h, cw, ch); }

}  // namespace resolv
