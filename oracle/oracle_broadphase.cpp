// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
// CPU restatement (oracle) of the sweep-and-prune broad phase (src/collision/broad_phase.rs:373-487) and of
// the persistent greedy constraint-graph colouring (src/dynamics/solver/constraint_graph.rs:163-296).
// PARITY UNPINNED: the reference has no unit test or golden vector for either (SURVEY.md §4 "Gaps").
// Both are integer/compare algorithms restated literally (insertion sort included), so the pair list is the
// exact sequence of ContactGraph::add_edge_and_key_with calls of the reference.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <unordered_set>
#include <vector>

#include "../include/avian_b200.h"

namespace {

inline uint64_t pair_key(uint32_t a, uint32_t b) {  // data_structures/pair_key.rs:15-21
    return a < b ? (uint64_t(a) << 32) | b : (uint64_t(b) << 32) | a;
}

template <class S>
int broadphase(AvnAabbColumns& ac, AvnPairList& out) {
    const uint32_t n_all = ac.count;
    const S* mn = static_cast<const S*>(ac.aabb_min);
    const S* mx = static_cast<const S*>(ac.aabb_max);
    // intervals in persistent order; `order` is the permutation being sorted.  update_aabb_intervals' retain drops the intervals whose AABB
    // is not finite before the sort ever sees them (broad_phase.rs:243-245)
    std::vector<uint32_t> order;
    order.reserve(n_all);
    for (uint32_t i = 0; i < n_all; ++i) {
        bool finite = true;
        for (int k = 0; k < 3; ++k) finite = finite && std::isfinite(mn[3 * i + k]) && std::isfinite(mx[3 * i + k]);
        if (finite) order.push_back(i);
    }
    const uint32_t n = uint32_t(order.size());
    ac.retained_count = n;
    const std::vector<uint32_t> retained = order;
    // insertion_sort(|a, b| a.min.x > b.min.x)  (broad_phase.rs:383,479-487)
    // The literal O(n^2) loop is kept while it is affordable; past a swap budget the oracle restarts with
    // std::stable_sort, which yields the identical permutation (a strict-'>' adjacent-swap insertion sort is a
    // stable sort under '<'; -0.0 and +0.0 compare equal in both; non-finite AABBs never reach the intervals,
    // broad_phase.rs:243-245).
    uint64_t swaps = 0, budget = 64ull * n + 1024;
    bool bailed = false;
    for (uint32_t i = 1; i < n && !bailed; ++i) {
        uint32_t j = i;
        while (j > 0 && mn[3 * order[j - 1]] > mn[3 * order[j]]) {
            std::swap(order[j - 1], order[j]);
            --j;
            if (++swaps > budget) { bailed = true; break; }
        }
    }
    if (bailed) {
        order = retained;
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return mn[3 * a] < mn[3 * b]; });
    }
    if (ac.order_out) std::memcpy(ac.order_out, order.data(), sizeof(uint32_t) * n);
    std::unordered_set<uint64_t> existing, jdis;
    existing.reserve(ac.existing_pair_count * 2 + 16);
    for (uint64_t k = 0; k < ac.existing_pair_count; ++k) existing.insert(ac.existing_pairs[k]);
    for (uint64_t k = 0; k < ac.joint_disabled_pair_count; ++k) jdis.insert(ac.joint_disabled_body_pairs[k]);
    uint64_t count = 0;
    for (uint32_t ii = 0; ii < n; ++ii) {
        const uint32_t a = order[ii];
        const uint8_t f1 = ac.flags ? ac.flags[a] : uint8_t(AVN_AABB_GENERATE_CONSTRAINTS);
        if (f1 & AVN_AABB_HALO) continue;  // x-slab partition: a halo copy never starts a sweep (include/avian_b200.h)
        const uint32_t m1 = ac.memberships ? ac.memberships[a] : 1u, fl1 = ac.filters ? ac.filters[a] : 0xFFFFFFFFu;
        for (uint32_t jj = ii + 1; jj < n; ++jj) {
            const uint32_t b = order[jj];
            if (mn[3 * b] > mx[3 * a]) break;
            if (mn[3 * a + 1] > mx[3 * b + 1] || mx[3 * a + 1] < mn[3 * b + 1]) continue;
            if (mn[3 * a + 2] > mx[3 * b + 2] || mx[3 * a + 2] < mn[3 * b + 2]) continue;
            const uint8_t f2 = ac.flags ? ac.flags[b] : uint8_t(AVN_AABB_GENERATE_CONSTRAINTS);
            const uint32_t m2 = ac.memberships ? ac.memberships[b] : 1u, fl2 = ac.filters ? ac.filters[b] : 0xFFFFFFFFu;
            const bool interacts = (m1 & fl2) != 0 && (m2 & fl1) != 0;  // layers.rs:423-426
            if ((f1 & f2 & AVN_AABB_IS_INACTIVE) || !interacts || ac.body[a] == ac.body[b]) continue;
            if ((f2 & AVN_AABB_NOT_J) || ((f1 & AVN_AABB_SPLIT_I) && (f2 & AVN_AABB_HALO))) continue;  // x-slab partition
            if (existing.count(pair_key(ac.collider[a], ac.collider[b]))) continue;
            if (!jdis.empty() && jdis.count(pair_key(ac.body[a], ac.body[b]))) continue;
            const uint8_t u = f1 | f2;
            uint8_t pf = 0;
            if (u & AVN_AABB_CONTACT_EVENTS) pf |= AVN_PAIR_CONTACT_EVENTS;
            if (u & AVN_AABB_MODIFY_CONTACTS) pf |= AVN_PAIR_MODIFY_CONTACTS;
            if (u & AVN_AABB_GENERATE_CONSTRAINTS) pf |= AVN_PAIR_GENERATE_CONSTRAINTS;
            if (u & AVN_AABB_CUSTOM_FILTER) pf |= AVN_PAIR_NEEDS_HOOK;
            if (count < out.capacity) {
                out.collider1[count] = ac.collider[a];
                out.collider2[count] = ac.collider[b];
                out.body1[count] = ac.body[a];
                out.body2[count] = ac.body[b];
                out.flags[count] = pf;
            }
            ++count;
            // the reference inserts the key into pair_set here (contact_graph.rs:521-565); a later (i,j) with the
            // same colliders cannot occur because every unordered pair is visited once.
        }
    }
    out.count = count;
    return count > out.capacity ? AVN_ERR_CAPACITY : AVN_OK;
}

}  // namespace

extern "C" {

int orc_broadphase(uint32_t scalar_bits, AvnAabbColumns* aabbs, AvnPairList* out) {
    if (!aabbs || !out) return AVN_ERR_INVALID_ARGUMENT;
    if (scalar_bits == 32) return broadphase<float>(*aabbs, *out);
    if (scalar_bits == 64) return broadphase<double>(*aabbs, *out);
    return AVN_ERR_INVALID_ARGUMENT;
}

// ---- ConstraintGraph (constraint_graph.rs) ---------------------------------------------------------------
// A minimal persistent graph: per-colour body bit sets + manifold handle vectors, push/pop as the narrow
// phase calls them.  Handles are opaque u64 ids chosen by the caller (contact id << 8 | manifold index).
struct OrcConstraintGraph {
    std::vector<std::vector<uint8_t>> body_set;       // [24][bits]
    std::vector<std::vector<uint64_t>> handles;       // [24]
    OrcConstraintGraph() : body_set(AVN_GRAPH_COLOR_COUNT), handles(AVN_GRAPH_COLOR_COUNT) {}
    bool get(int c, uint32_t i) const { return i < body_set[c].size() && body_set[c][i]; }
    void set(int c, uint32_t i) { if (i >= body_set[c].size()) body_set[c].resize(i + 1, 0); body_set[c][i] = 1; }
    void unset(int c, uint32_t i) { if (i < body_set[c].size()) body_set[c][i] = 0; }
};

OrcConstraintGraph* orc_graph_create() { return new OrcConstraintGraph(); }
void orc_graph_destroy(OrcConstraintGraph* g) { delete g; }

// push_manifold (constraint_graph.rs:163-236). Returns colour index; *local_index = position in the colour.
int orc_graph_push(OrcConstraintGraph* g, uint32_t body1, uint32_t body2, int is_static1, int is_static2, uint64_t handle,
                   uint32_t* local_index) {
    int color = AVN_COLOR_OVERFLOW;
    if (!is_static1 && !is_static2) {
        for (int i = 0; i < AVN_DYNAMIC_COLOR_COUNT; ++i) {
            if (g->get(i, body1) || g->get(i, body2)) continue;
            g->set(i, body1);
            g->set(i, body2);
            color = i;
            break;
        }
    } else if (!is_static1) {
        for (int i = AVN_COLOR_OVERFLOW - 1; i >= 1; --i) {
            if (g->get(i, body1)) continue;
            g->set(i, body1);
            color = i;
            break;
        }
    } else if (!is_static2) {
        for (int i = AVN_COLOR_OVERFLOW - 1; i >= 1; --i) {
            if (g->get(i, body2)) continue;
            g->set(i, body2);
            color = i;
            break;
        }
    }
    if (local_index) *local_index = uint32_t(g->handles[color].size());
    g->handles[color].push_back(handle);
    return color;
}

// pop_manifold (constraint_graph.rs:245-296): swap_remove; returns the handle that moved into local_index
// (or ~0 when none moved) so the caller can fix its back-reference.
uint64_t orc_graph_pop(OrcConstraintGraph* g, int color, uint32_t local_index, uint32_t body1, uint32_t body2) {
    if (color != AVN_COLOR_OVERFLOW) {
        g->unset(color, body1);
        g->unset(color, body2);
    }
    auto& v = g->handles[color];
    uint32_t moved = uint32_t(v.size()) - 1;
    v[local_index] = v[moved];
    v.pop_back();
    return moved != local_index ? v[local_index] : ~uint64_t(0);
}

uint32_t orc_graph_color_size(const OrcConstraintGraph* g, int color) { return uint32_t(g->handles[color].size()); }
void orc_graph_color_handles(const OrcConstraintGraph* g, int color, uint64_t* out) {
    std::memcpy(out, g->handles[color].data(), sizeof(uint64_t) * g->handles[color].size());
}
}
