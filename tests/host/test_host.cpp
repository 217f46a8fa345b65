// Exercises the C++ host layer (resolv_b200/csrc/host/resolv.hpp) the way a resolv user would, and prints a
// transcript that tests/test_host_gpu.py compares with the oracle. Hex floats keep every bit.
#include <cstdio>

#include "../../resolv_b200/csrc/host/resolv.hpp"

using namespace resolv;

static void print_set(const char* tag, int tester, const IntersectionSet& s) {
    printf("%s %d %u %zu %a %a %a %a", tag, tester, s.OtherShape->ID(), s.Intersections.size(), s.MTV.X, s.MTV.Y, s.Center.X, s.Center.Y);
    for (auto& c : s.Intersections) printf(" %a %a %a %a", c.Point.X, c.Point.Y, c.Normal.X, c.Normal.Y);
    printf("\n");
}

int main() {
    const Tags SolidWall = 1, Platform = 2, Player = 4;
    auto space = NewSpace(640, 360, 16, 16);
    std::vector<std::unique_ptr<Shape>> own;
    auto add = [&](std::unique_ptr<Shape> s, Tags t) { s->GetTags() = t; space->Add(s.get()); own.push_back(std::move(s)); return own.back().get(); };
    // a slice of examples/worldPlatformer.go:34-92
    add(NewCircle(128, 128, 32), SolidWall | Platform);                      // 0
    add(NewRectangleFromTopLeft(0, 0, 640, 8), SolidWall | Platform);        // 1
    add(NewRectangleFromTopLeft(64, 200, 300, 8), SolidWall | Platform);     // 2
    add(NewRectangleFromTopLeft(512, 96, 32, 200), SolidWall | Platform);    // 3
    add(NewRectangleFromTopLeft(0, 352, 640, 8), SolidWall | Platform);      // 4
    add(NewRectangleFromTopLeft(400, 200, 32, 16), Platform);                // 5
    add(NewConvexPolygon(180, 175, {-24, 8, 8, -8, 48, -8, 80, 8}), Platform | SolidWall);  // 6 ramp
    Shape* player = add(NewRectangle(192, 128, 16, 12), Player);             // 7
    Shape* ball = add(NewCircle(150, 150, 8), Player);                       // 8

    // 1. IntersectionTest with a filter chain, callbacks nearest-first, early exit honoured
    const double px[] = {192, 100, 150, 520, 200, 405, 170};
    const double py[] = {128, 196, 140, 150, 346, 204, 172};
    for (int k = 0; k < 7; k++) {
        player->SetPosition(px[k], py[k]);
        auto near = player->SelectTouchingCells(4).ByTags(SolidWall);
        IntersectionTestSettings st;
        st.TestAgainst = &near;
        st.OnIntersect = [&](const IntersectionSet& set) { print_set("IT", k, set); return true; };
        bool hit = player->IntersectionTest(st);
        printf("ITR %d %d\n", k, (int)hit);
        int calls = 0;
        st.OnIntersect = [&](const IntersectionSet&) { calls++; return false; };
        player->IntersectionTest(st);
        printf("ITE %d %d\n", k, calls);
    }
    // 2. circle tester, no tag filter, Not(player)
    for (int k = 0; k < 3; k++) {
        ball->SetPosition(150 + 8 * k, 150 + 3 * k);
        auto near = ball->SelectTouchingCells(1).Not(player);
        IntersectionTestSettings st;
        st.TestAgainst = &near;
        st.OnIntersect = [&](const IntersectionSet& set) { print_set("IC", k, set); return true; };
        ball->IntersectionTest(st);
    }
    // (the walls never moved: the host layer flagged them static and the device kept their registrations)
    { auto si = space->StaticInfo(); printf("SI %u %u\n", si.first, si.second); }
    // 3. LineTest against the whole Space
    {
        auto all = space->FilterShapes().ByTags(Platform);
        LineTestSettings lt;
        lt.Start = {20, 30}; lt.End = {600, 330};
        lt.TestAgainst = &all;
        lt.OnIntersect = [&](const IntersectionSet& set, int i, int n) { print_set("LT", i * 100 + n, set); return true; };
        printf("LTR %d\n", (int)LineTest(lt));
    }
    // 4. ShapeLineTest downwards from the player (examples/worldPlatformer.go:221-246)
    player->SetPosition(200, 188);
    {
        auto near = player->SelectTouchingCells(4).ByTags(Platform);
        ShapeLineTestSettings sl;
        sl.Vec = {0, 12};
        sl.TestAgainst = &near;
        sl.OnIntersect = [&](const IntersectionSet& set, int i, int n) { print_set("SL", i * 100 + n, set); return true; };
        printf("SLR %d\n", (int)static_cast<ConvexPolygon*>(player)->ShapeLineTest(sl));
    }
    // 5. a callback that moves the tester (shape.go:252-258): later sets are re-tested at the new position
    player->SetPosition(100, 196);
    {
        auto near = player->SelectTouchingCells(4).ByTags(SolidWall);
        IntersectionTestSettings st;
        st.TestAgainst = &near;
        st.OnIntersect = [&](const IntersectionSet& set) { print_set("MV", 0, set); player->MoveVec(set.MTV); return true; };
        player->IntersectionTest(st);
        printf("MVP %a %a\n", player->Position().X, player->Position().Y);
    }
    // 6. Remove
    space->Remove(own[2].get());
    player->SetPosition(100, 196);
    {
        auto near = player->SelectTouchingCells(4);
        IntersectionTestSettings st;
        st.TestAgainst = &near;
        st.OnIntersect = [&](const IntersectionSet& set) { print_set("RM", 0, set); return true; };
        printf("RMR %d\n", (int)player->IntersectionTest(st));
    }
    { auto si = space->StaticInfo(); printf("SI %u %u\n", si.first, si.second); }
    // 7. the first callback's MoveVec(MTV) pushes the tester INTO a shape it did not touch before the move: the
    //    reference tests every remaining candidate against the current world, so the wall is reported in the same call
    Shape* block = add(NewRectangle(600, 66, 16, 8), SolidWall);               // 9
    Shape* wall = add(NewRectangle(606, 46, 8, 11), SolidWall);                // 10
    (void)block; (void)wall;
    player->SetPosition(600, 59);
    {
        auto near = player->SelectTouchingCells(2).ByTags(SolidWall);
        IntersectionTestSettings st;
        st.TestAgainst = &near;
        st.OnIntersect = [&](const IntersectionSet& set) { print_set("PW", 0, set); player->MoveVec(set.MTV); return true; };
        player->IntersectionTest(st);
        printf("PWP %a %a\n", player->Position().X, player->Position().Y);
    }
    // 8. TestAgainst = the whole Space (space.FilterShapes(), every shape in Add order), with ByDistance on top
    player->SetPosition(100, 196);
    {
        auto all = space->FilterShapes().ByTags(SolidWall);
        IntersectionTestSettings st;
        st.TestAgainst = &all;
        st.OnIntersect = [&](const IntersectionSet& set) { print_set("WS", 0, set); return true; };
        player->IntersectionTest(st);
        auto ring = space->FilterShapes().ByDistance(Vector{100, 196}, 50, 200);
        st.TestAgainst = &ring;
        st.OnIntersect = [&](const IntersectionSet& set) { print_set("BD", 0, set); return true; };
        player->IntersectionTest(st);
    }
    // 9. more than 12 candidates with exactly equal distances: the callback order is whatever Go's pdqsort leaves
    {
        Shape* hub = add(NewCircle(300, 100, 5.5), Player);                    // 11
        const int off[24][2] = {{3, 4}, {-3, 4}, {3, -4}, {-3, -4}, {4, 3}, {-4, 3}, {4, -3}, {-4, -3}, {5, 0}, {-5, 0}, {0, 5}, {0, -5},
                                {5, 5}, {-5, 5}, {5, -5}, {-5, -5}, {1, 7}, {-1, 7}, {1, -7}, {-1, -7}, {7, 1}, {-7, 1}, {7, -1}, {-7, -1}};
        for (auto& o : off) add(NewCircle(300 + o[0], 100 + o[1], 2), Platform);  // 12..35
        auto near = hub->SelectTouchingCells(1).ByTags(Platform);
        IntersectionTestSettings st;
        st.TestAgainst = &near;
        st.OnIntersect = [&](const IntersectionSet& set) { print_set("PQ", 0, set); return true; };
        hub->IntersectionTest(st);
    }
    // 10. ShapeLineTestSettings.Lines is matched against the slice's indices, not its values (convexPolygon.go:329-340)
    player->SetPosition(416, 190);
    {
        auto near = player->SelectTouchingCells(4).ByTags(Platform);
        for (int variant = 0; variant < 3; variant++) {
            ShapeLineTestSettings sl;
            sl.Vec = {0, 12};
            sl.TestAgainst = &near;
            if (variant == 0) sl.Lines = {3};            // one element: only line 0 is considered, whatever the value
            if (variant == 1) sl.Lines = {0};
            if (variant == 2) sl.Lines = {9, 9, 9, 9};   // four elements: all four lines
            sl.OnIntersect = [&](const IntersectionSet& set, int i, int n) { print_set("LQ", variant * 10000 + i * 100 + n, set); return true; };
            printf("LQR %d %d\n", variant, (int)static_cast<ConvexPolygon*>(player)->ShapeLineTest(sl));
        }
    }
    // 11. one IntersectionSet of more than 12 contacts: a circle cuts every edge of an octagon twice (16 points, all at
    // distance r from the circle: every sort key ties). sort.Slice is pdqsort beyond 12 items, so the device delivers the
    // set in generation order and Space::fillSet sorts it with the port of Go's sort.
    {
        Shape* oct = add(NewConvexPolygon(500, 40, {20, 0, 14, 14, 0, 20, -14, 14, -20, 0, -14, -14, 0, -20, 14, -14}), Player);
        add(NewCircle(500, 40, 19), Platform);
        auto near = oct->SelectTouchingCells(1).ByTags(Platform);
        IntersectionTestSettings st;
        st.TestAgainst = &near;
        st.OnIntersect = [&](const IntersectionSet& set) { print_set("MC", 0, set); return true; };
        oct->IntersectionTest(st);
    }
    return 0;
}
