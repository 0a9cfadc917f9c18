// Per-island-group work lists for the island schedule of the solver stage (host code, no CUDA).
// Constraints only couple the bodies of one connected component of dynamic bodies (solver/islands/mod.rs:3-4: static bodies do not merge
// islands), so a scene made of MANY SMALL islands — a field of ragdolls, many small piles — needs no grid-wide synchronisation inside the
// substep loop at all: one thread block can take a GROUP of islands through the whole loop, with __syncthreads() where the barrier schedule has
// a grid barrier, and the group's state stays in that SM's L1.  This builds, per group of `group` consecutive islands, the bodies, the
// manifold slots grouped by graph colour and the joint slots grouped by schedule level, all in the order the barrier schedule visits them
// (so the results are bit-identical, and the items of a level stay sorted by joint type across the block's warps).
#pragma once
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <vector>

#include "../../include/avian_b200.h"

namespace avn {

struct IslandLists {
    std::vector<int> root, island_of, body_off, bodies, m_off, mslots, j_off, jslots, cursor;
    int count = 0 /* groups */, islands = 0, max_bodies = 0 /* of one island */, levels = 0;
    bool ok = false;
};

// kind == NULL: every body dynamic.  body1/body2[M] = AVN_NO_BODY or a body index; manifold m of colour c sits in slot
// color_off[c] + (m - m_color_off[c]).  jb1/jb2[J] = bodies of the joint in schedule slot s; level_off[n_levels + 1].
inline void build_island_lists(int B, const uint8_t* kind, int M, const int32_t* body1, const int32_t* body2, const int* m_color_off, const int* color_off,
                               int J, const int* jb1, const int* jb2, const int* level_off, int n_levels, int target_units, IslandLists& out) {
    out.ok = false;
    out.count = 0;
    if (B <= 0) return;
    auto dyn = [&](int b) { return b >= 0 && b < B && (!kind || kind[b] == AVN_BODY_DYNAMIC); };
    if (kind)
        for (int b = 0; b < B; ++b)
            if (kind[b] == AVN_BODY_KINEMATIC) return;   // a moving body shared by islands couples them through its delta: barrier schedule
    if (m_color_off[AVN_COLOR_OVERFLOW + 1] > m_color_off[AVN_COLOR_OVERFLOW]) return;   // the serial overflow colour: barrier schedule
    std::vector<int>& root = out.root;
    root.resize(B);
    std::iota(root.begin(), root.end(), 0);
    auto find = [&](int x) { while (root[x] != x) { root[x] = root[root[x]]; x = root[x]; } return x; };
    auto unite = [&](int a, int b) { a = find(a); b = find(b); if (a != b) { if (a < b) root[b] = a; else root[a] = b; } };
    for (int m = 0; m < M; ++m) if (dyn(body1[m]) && dyn(body2[m])) unite(body1[m], body2[m]);
    for (int s = 0; s < J; ++s) if (dyn(jb1[s]) && dyn(jb2[s])) unite(jb1[s], jb2[s]);
    // islands numbered by their smallest body; bodies ascending inside an island
    std::vector<int>& isl = out.island_of;
    isl.assign(B, -1);
    int I = 0;
    for (int b = 0; b < B; ++b) if (dyn(b) && find(b) == b) isl[b] = I++;
    for (int b = 0; b < B; ++b) if (dyn(b)) isl[b] = isl[find(b)];
    // largest island (in bodies), then groups of consecutive islands as the unit of work
    {
        std::vector<int>& cnt = out.cursor;
        cnt.assign(I, 0);
        for (int b = 0; b < B; ++b) if (isl[b] >= 0) ++cnt[isl[b]];
        out.max_bodies = 0;
        for (int i = 0; i < I; ++i) out.max_bodies = std::max(out.max_bodies, cnt[i]);
    }
    out.islands = I;
    const int group = std::max(1, (I + std::max(target_units, 1) - 1) / std::max(target_units, 1));
    for (int b = 0; b < B; ++b) if (isl[b] >= 0) isl[b] /= group;
    I = (I + group - 1) / group;
    const int island_max_bodies = out.max_bodies;
    out.body_off.assign(I + 1, 0);
    for (int b = 0; b < B; ++b) if (isl[b] >= 0) ++out.body_off[isl[b] + 1];
    int max_bodies = 0;
    for (int i = 0; i < I; ++i) { max_bodies = std::max(max_bodies, out.body_off[i + 1]); out.body_off[i + 1] += out.body_off[i]; }
    out.bodies.resize(out.body_off[I]);
    out.cursor.assign(out.body_off.begin(), out.body_off.end() - 1);
    for (int b = 0; b < B; ++b) if (isl[b] >= 0) out.bodies[out.cursor[isl[b]]++] = b;
    auto owner = [&](int a, int b) { return dyn(a) ? isl[a] : (dyn(b) ? isl[b] : -1); };
    // manifold slots per (island, colour), ascending slot inside a colour
    const int NC = AVN_GRAPH_COLOR_COUNT;
    out.m_off.assign(size_t(I) * (NC + 1) + 1, 0);
    auto colour_of = [&](int m) { int c = 0; while (c < NC - 1 && m >= m_color_off[c + 1]) ++c; return c; };
    {
        int c = 0;
        for (int m = 0; m < M; ++m) {
            while (c < NC - 1 && m >= m_color_off[c + 1]) ++c;
            const int o = owner(body1[m], body2[m]);
            if (o >= 0) ++out.m_off[size_t(o) * (NC + 1) + c + 1];
        }
    }
    {   // exclusive scan in (island, colour) order; entry [i][NC] = end of island i = begin of island i + 1
        int run = 0;
        for (int i = 0; i < I; ++i) {
            int* row = &out.m_off[size_t(i) * (NC + 1)];
            int prev = run;
            for (int c = 0; c < NC; ++c) { const int n = row[c + 1]; row[c] = prev; prev += n; }
            row[NC] = prev;
            run = prev;
        }
        out.mslots.resize(run);
    }
    {
        std::vector<int> cur(size_t(I) * NC);
        for (int i = 0; i < I; ++i) for (int c = 0; c < NC; ++c) cur[size_t(i) * NC + c] = out.m_off[size_t(i) * (NC + 1) + c];
        int c = 0;
        for (int m = 0; m < M; ++m) {
            while (c < NC - 1 && m >= m_color_off[c + 1]) ++c;
            const int o = owner(body1[m], body2[m]);
            if (o >= 0) out.mslots[cur[size_t(o) * NC + c]++] = color_off[c] + (m - m_color_off[c]);
        }
        (void)colour_of;
    }
    // joint slots per (island, level), ascending slot inside a level
    const int L = n_levels;
    out.j_off.assign(size_t(I) * (L + 1) + 1, 0);
    if (J > 0) {
        int l = 0;
        for (int s = 0; s < J; ++s) {
            while (l < L - 1 && s >= level_off[l + 1]) ++l;
            const int o = owner(jb1[s], jb2[s]);
            if (o >= 0) ++out.j_off[size_t(o) * (L + 1) + l + 1];
        }
        int run = 0;
        for (int i = 0; i < I; ++i) {
            int* row = &out.j_off[size_t(i) * (L + 1)];
            int prev = run;
            for (int k = 0; k < L; ++k) { const int n = row[k + 1]; row[k] = prev; prev += n; }
            row[L] = prev;
            run = prev;
        }
        out.jslots.resize(run);
        std::vector<int> cur(size_t(I) * std::max(L, 1));
        for (int i = 0; i < I; ++i) for (int k = 0; k < L; ++k) cur[size_t(i) * L + k] = out.j_off[size_t(i) * (L + 1) + k];
        l = 0;
        for (int s = 0; s < J; ++s) {
            while (l < L - 1 && s >= level_off[l + 1]) ++l;
            const int o = owner(jb1[s], jb2[s]);
            if (o >= 0) out.jslots[cur[size_t(o) * L + l]++] = s;
        }
    } else {
        out.jslots.clear();
    }
    out.count = I;
    (void)max_bodies;
    out.max_bodies = island_max_bodies;
    out.levels = L;
    out.ok = true;
}

}  // namespace avn
