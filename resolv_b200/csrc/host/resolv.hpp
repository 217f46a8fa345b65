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
// equivalent to the reference applying them before the pair test. Callbacks run nearest-first with ties in
// candidate generation order (shape.go:291-293; sort.Slice is an insertion sort for n <= 12).
#pragma once
#include <algorithm>
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
        const uint32_t flags = RSV_FRAME_DOWNLOAD | RSV_FRAME_TESTER_TABLE | RSV_FRAME_NO_TAG_FILTERS | RSV_FRAME_INCREMENTAL;
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
        margin_ = margin; stale_ = false; counts_ = c;
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
    void markDirty(Shape* s) { writeRow(s, false); moved_[s->slot_] = 1; stale_ = true; }
    void markMoved(Shape* s) {
        const uint32_t i = s->slot_;
        pos_[2 * i] = s->pos_.X; pos_[2 * i + 1] = s->pos_.Y;
        moved_[i] = 1;
        if (meta_[i] & RSV_META_STATIC) { meta_[i] &= ~RSV_META_STATIC; metaDirty_ = true; }   // demotion is immediate
        posLo_ = std::min(posLo_, i); posHi_ = std::max(posHi_, i + 1);
        stale_ = true;
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
    std::vector<uint32_t> first_, count_;
    rsv_results res_{};
    rsv_counts counts_{};
};

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
    sp.ensureFrame(st.TestAgainst->margin);
    // copy this tester's sets out of the borrowed result views first: a callback may move shapes, which re-runs the
    // frame and overwrites the views
    struct Cand { double d; uint32_t rank; Shape* other; IntersectionSet set; };
    std::vector<Cand> list;
    const rsv_hit* hits = sp.res_.hits;
    for (uint32_t k = 0; k < sp.count_[slot_]; k++) {
        const rsv_hit* h = hits + sp.first_[slot_] + k;
        const uint32_t n = h->count_rank & 0x7f;
        if (n == 0) continue;
        Shape* other = sp.shapes_[h->b];
        if (!st.TestAgainst->Accepts(*other)) continue;
        Cand c{other->pos_.DistanceSquared(pos_), h->count_rank >> 8, other, {}};   // shape.go:179-181
        c.set.OtherShape = other;
        c.set.MTV = {h->mtv_x, h->mtv_y};
        for (uint32_t i = 0; i < n; i++) {
            const rsv_contact& q = sp.res_.contacts[h->first_contact + i];
            c.set.Intersections.push_back({{q.px, q.py}, {q.nx, q.ny}});
        }
        list.push_back(std::move(c));
    }
    std::stable_sort(list.begin(), list.end(), [](const Cand& a, const Cand& b) { return a.d < b.d || (a.d == b.d && a.rank < b.rank); });
    bool collided = false;
    const Vector before = pos_;
    for (auto& c : list) {
        IntersectionSet set;
        if (pos_.X != before.X || pos_.Y != before.Y) {
            // a callback moved the tester: later sets must see the new position (shape.go:252-258) -> re-test that pair
            set = Intersection(*c.other);
            if (set.IsEmpty()) continue;
        } else {
            set = c.set;
        }
        collided = true;
        if (!st.OnIntersect || !st.OnIntersect(set)) break;
    }
    return collided;
}

// A single ordered pair through the device: the frame cache is refreshed and the record (this, other) looked up.
inline IntersectionSet Shape::Intersection(Shape& other) {
    IntersectionSet set;
    if (!space_ || other.space_ != space_) return set;
    Space& sp = *space_;
    // margin large enough that `other` is a candidate whenever the Bounds overlap (they then share a cell)
    sp.ensureFrame(sp.margin_ < 0 ? 1 : sp.margin_);
    for (uint32_t k = 0; k < sp.count_[slot_]; k++) {
        const rsv_hit* h = sp.res_.hits + sp.first_[slot_] + k;
        if (h->b != other.slot_ || (h->count_rank & 0x7f) == 0) continue;
        set.OtherShape = &other;
        set.MTV = {h->mtv_x, h->mtv_y};
        for (uint32_t i = 0; i < (h->count_rank & 0x7f); i++) {
            const rsv_contact& c = sp.res_.contacts[h->first_contact + i];
            set.Intersections.push_back({{c.px, c.py}, {c.nx, c.ny}});
        }
    }
    return set;
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
    std::stable_sort(list.begin(), list.end(), [](const Cand& a, const Cand& b) { return a.key < b.key || (a.key == b.key && a.rank < b.rank); });
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
    std::stable_sort(results.begin(), results.end(), [](const IntersectionSet& a, const IntersectionSet& b) { return a.MTV.MagnitudeSquared() < b.MTV.MagnitudeSquared(); });
    for (size_t i = 0; i < results.size(); i++)
        if (st.OnIntersect && !st.OnIntersect(results[i], (int)i, (int)results.size())) break;
    return !results.empty();
}

inline std::unique_ptr<Space> NewSpace(int w, int h, int cw, int ch) { return std::make_unique<Space>(w, h, cw, ch); }

}  // namespace resolv
