// Host-side rectangle algebra of the compositor's tile scheduler: the product's own implementation of
// box2i (math/dplug/math/box.d), Mipmap.impactOnNextLevel (graphics/dplug/graphics/mipmap.d:623-645),
// tileAreas / removeOverlappingAreas (gui/dplug/gui/boxlist.d:54-167) and
// convertPBRLayerRectToRawLayerRect (gui/dplug/gui/graphics.d:981-1002).
#pragma once
#include <mutex>
#include <vector>
#include "../../include/b2pbr.h"

namespace b2 {

inline int boxW(const b2_box2i& b) { return b.max_x - b.min_x; }
inline int boxH(const b2_box2i& b) { return b.max_y - b.min_y; }
// box.d:188-192 — a box is empty as soon as one dimension is degenerate
inline bool boxEmpty(const b2_box2i& b) { return b.min_x == b.max_x || b.min_y == b.max_y; }
inline bool boxSorted(const b2_box2i& b) { return b.min_x <= b.max_x && b.min_y <= b.max_y; }
inline bool boxContains(const b2_box2i& a, const b2_box2i& o) {
    return !(o.min_x < a.min_x || o.max_x > a.max_x || o.min_y < a.min_y || o.max_y > a.max_y);
}
// box.d:298-319 — an empty operand is returned as is
inline b2_box2i boxIntersect(const b2_box2i& a, const b2_box2i& o) {
    if (boxEmpty(a)) return a;
    if (boxEmpty(o)) return o;
    b2_box2i r;
    r.min_x = a.min_x > o.min_x ? a.min_x : o.min_x;
    int hx = a.max_x < o.max_x ? a.max_x : o.max_x;
    r.max_x = hx >= r.min_x ? hx : r.min_x;
    r.min_y = a.min_y > o.min_y ? a.min_y : o.min_y;
    int hy = a.max_y < o.max_y ? a.max_y : o.max_y;
    r.max_y = hy >= r.min_y ? hy : r.min_y;
    return r;
}

inline b2_box2i impactOnNextLevel(int quality, b2_box2i a, int curW, int curH) {
    b2_box2i maxArea{0, 0, curW / 2, curH / 2};
    b2_box2i n;
    if (quality == B2_QUALITY_CUBIC) n = b2_box2i{(a.min_x - 1) / 2, (a.min_y - 1) / 2, (a.max_x + 2) / 2, (a.max_y + 2) / 2};
    else n = b2_box2i{a.min_x / 2, a.min_y / 2, (a.max_x + 1) / 2, (a.max_y + 1) / 2};
    return boxIntersect(n, maxArea);
}

inline void tileAreas(const b2_box2i* areas, int n, int maxW, int maxH, std::vector<b2_box2i>& out) {
    for (int k = 0; k < n; ++k) {
        const b2_box2i& a = areas[k];
        const int w = boxW(a), h = boxH(a);
        for (int y0 = 0; y0 < h; y0 += maxH)
            for (int x0 = 0; x0 < w; x0 += maxW) {
                int x1 = x0 + maxW < w ? x0 + maxW : w;
                int y1 = y0 + maxH < h ? y0 + maxH : h;
                out.push_back(b2_box2i{a.min_x + x0, a.min_y + y0, a.min_x + x1, a.min_y + y1});
            }
    }
}

inline void removeOverlappingAreas(std::vector<b2_box2i>& boxes, std::vector<b2_box2i>& filtered) {
    for (size_t i = 0; i < boxes.size(); ++i) {
        const b2_box2i A = boxes[i];
        if (boxW(A) * boxH(A) <= 0) continue;
        bool split = false;
        for (size_t j = i + 1; j < boxes.size(); ++j) {
            const b2_box2i B = boxes[j];
            const b2_box2i C = boxIntersect(A, B);
            if (!(boxSorted(C) && !boxEmpty(C))) continue;
            if (boxContains(A, B)) {  // B adds nothing: drop it, keep scanning the element swapped in
                boxes[j] = boxes.back();
                boxes.pop_back();
                --j;
                continue;
            }
            split = true;
            if (!boxContains(B, A)) {
                const b2_box2i parts[4] = {{A.min_x, A.min_y, A.max_x, C.min_y}, {A.min_x, C.min_y, C.min_x, C.max_y},
                                           {C.max_x, C.min_y, A.max_x, C.max_y}, {A.min_x, C.max_y, A.max_x, A.max_y}};
                for (const b2_box2i& p : parts)
                    if (!boxEmpty(p)) boxes.push_back(p);
            }
            break;
        }
        if (!split) filtered.push_back(A);
    }
}

// box.d:321-325 — two boxes intersect when their (sorted) intersection is not empty
inline bool boxIntersects(const b2_box2i& a, const b2_box2i& b) {
    const b2_box2i c = boxIntersect(a, b);
    return boxSorted(c) && !boxEmpty(c);
}
inline bool haveNoOverlap(const b2_box2i* areas, int n) {   // boxlist.d:171-186
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j)
            if (boxIntersects(areas[i], areas[j])) return false;
    return true;
}

// The widgets' dirty-rectangle accumulator (DirtyRectList, gui/dplug/gui/boxlist.d:200-303): rectangles are added
// from the UI thread and from host parameter-change callbacks (hence the lock), the draw thread pulls the whole
// list.  The list never holds overlapping rectangles: a new rectangle that lies inside an element is dropped,
// elements inside it are removed, and elements it cuts are replaced by their up-to-four remainders.
class DirtyRectList {
public:
    bool isEmpty() {
        std::lock_guard<std::mutex> g(lock_);
        return rects_.empty();
    }
    void pullAllRectangles(std::vector<b2_box2i>& out) {   // appends, then clears (boxlist.d:224-234)
        std::lock_guard<std::mutex> g(lock_);
        out.insert(out.end(), rects_.begin(), rects_.end());
        rects_.clear();
    }
    void addRect(const b2_box2i& rect) {
        if (boxEmpty(rect)) return;
        std::lock_guard<std::mutex> g(lock_);
        size_t i = 0;
        while (i < rects_.size()) {
            const b2_box2i other = rects_[i];
            if (boxContains(other, rect)) return;            // already covered
            const bool swallowed = boxContains(rect, other);
            const b2_box2i common = swallowed ? other : boxIntersect(other, rect);
            if (!swallowed && boxEmpty(common)) { ++i; continue; }   // disjoint: keep
            // `other` leaves the list (the last element takes its slot and is examined next) ...
            rects_[i] = rects_.back();
            rects_.pop_back();
            if (swallowed) continue;
            // ... and what is left of it outside the new rectangle goes to the end: above, left, right, below
            const b2_box2i rest[4] = {{other.min_x, other.min_y, other.max_x, common.min_y}, {other.min_x, common.min_y, common.min_x, common.max_y},
                                      {common.max_x, common.min_y, other.max_x, common.max_y}, {other.min_x, common.max_y, other.max_x, other.max_y}};
            for (const b2_box2i& r : rest)
                if (!boxEmpty(r)) rects_.push_back(r);
        }
        rects_.push_back(rect);
    }

private:
    std::vector<b2_box2i> rects_;
    std::mutex lock_;
};

// Coalesces a list of disjoint rectangles for bulk copies: horizontally adjacent boxes with the same rows are
// joined, then vertically adjacent spans with the same columns (a full-frame 64x64 tile list collapses to one
// rectangle).  Coverage is unchanged; used only to batch host<->device transfers.
inline void mergeRects(const b2_box2i* rects, int n, std::vector<b2_box2i>& out) {
    std::vector<b2_box2i> v;
    for (int i = 0; i < n; ++i)
        if (boxW(rects[i]) > 0 && boxH(rects[i]) > 0) v.push_back(rects[i]);
    auto pass = [&](bool horizontal) {
        std::vector<b2_box2i> w;
        for (const b2_box2i& r : v) {
            bool joined = false;
            for (auto it = w.rbegin(); it != w.rend() && (it - w.rbegin()) < 64; ++it) {
                if (horizontal && it->min_y == r.min_y && it->max_y == r.max_y && it->max_x == r.min_x) { it->max_x = r.max_x; joined = true; break; }
                if (!horizontal && it->min_x == r.min_x && it->max_x == r.max_x && it->max_y == r.min_y) { it->max_y = r.max_y; joined = true; break; }
            }
            if (!joined) w.push_back(r);
        }
        v.swap(w);
    };
    pass(true);
    pass(false);
    out.swap(v);
}

inline b2_box2i growRect(b2_box2i r, int margin, int width, int height) {
    int xmin = r.min_x - margin, ymin = r.min_y - margin, xmax = r.max_x + margin, ymax = r.max_y + margin;
    if (xmin < 0) xmin = 0;
    if (ymin < 0) ymin = 0;
    if (xmax > width) xmax = width;
    if (ymax > height) ymax = height;
    if (xmax < 0) xmax = 0;
    if (ymax < 0) ymax = 0;
    if (xmin > width) xmin = width;
    if (ymin > height) ymin = height;
    return b2_box2i{xmin, ymin, xmax, ymax};
}

}  // namespace b2
