// libavian_host.so — the host-side FIXTURE that stands in for the parts of an avian3d app that stay on the CPU
// around the GPU hot path, so the hot path can be exercised and benchmarked without Bevy:
//   * swept collider AABBs        (restates update_aabb for cuboids/spheres, collider/backend.rs:498-625)
//   * ContactGraph bookkeeping    (pair set, lowest-free ContactId, contact_graph.rs:521-631; id_pool.rs:43-52)
//   * a narrow phase for cuboid / sphere pairs (SAT + face clipping; the reference delegates this arithmetic
//     to parry3d 0.25, which is not vendored, so this is OUR manifold generator: a fixture, identical for the
//     oracle and the GPU path, not a parity claim), contact matching (contact_types/mod.rs:426-470),
//     the status-change loop and ConstraintGraph push/pop colouring (narrow_phase/system_param.rs:136-389,
//     constraint_graph.rs:163-296)
//   * export of the manifolds as per-colour columns in the layout of AvnManifoldColumns.
// It contains no solver or broad-phase code: those are the GPU library (product) or oracle/ (tests).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <queue>
#include <unordered_map>
#include <vector>

#include "../../include/avian_b200.h"
#include "../csrc/narrow_math.hpp"
#include "../csrc/contact_rows.hpp"

namespace {

using namespace nm;  // S = double, V3, Q, M3, Box, Contacts: the manifold geometry shared with the device (csrc/narrow_math.hpp)

inline Q qmul(Q a, Q b) {
    return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
            a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
inline Q q_from_scaled_axis(V3 v) {
    S l = len(v);
    if (l == 0) return {0, 0, 0, 1};
    S s = std::sin(l * 0.5) / l;
    return {v.x * s, v.y * s, v.z * s, std::cos(l * 0.5)};
}
template <class T> struct ColR {  // typed reader over a void* column of float or double
    const void* p; bool f64;
    S at(size_t i) const { return f64 ? static_cast<const double*>(p)[i] : S(static_cast<const float*>(p)[i]); }
    V3 v3(size_t i) const { return {at(3 * i), at(3 * i + 1), at(3 * i + 2)}; }
    Q q(size_t i) const { return {at(4 * i), at(4 * i + 1), at(4 * i + 2), at(4 * i + 3)}; }
};
using Col = ColR<void>;
struct ColW {
    void* p; bool f64;
    void set(size_t i, S v) const { if (f64) static_cast<double*>(p)[i] = v; else static_cast<float*>(p)[i] = float(v); }
    void set3(size_t i, V3 v) const { set(3 * i, v.x); set(3 * i + 1, v.y); set(3 * i + 2, v.z); }
};

struct Shape { int type; V3 he; S friction, restitution; };

struct Point {
    V3 anchor1, anchor2;  // world-space offsets from each body's centre of mass
    S penetration, normal_speed;
    S ws_normal = 0, ws_tx = 0, ws_ty = 0, normal_impulse = 0;
};
struct Manifold { V3 normal; std::vector<Point> pts; };
struct Handle { int color; uint32_t local; };

struct Pair {
    uint32_t collider1, collider2, body1, body2;
    uint8_t flags;                 // AVN_PAIR_*
    bool alive = false, touching = false, static1 = false, static2 = false;
    std::vector<Manifold> manifolds;
    std::vector<Handle> handles;   // ContactEdge::constraint_handles
};

// ConstraintGraph (constraint_graph.rs:39-296)
struct Color {
    std::vector<uint8_t> body_set;
    std::vector<std::pair<uint32_t, uint32_t>> handles;  // (contact id, manifold index)
    bool get(uint32_t i) const { return i < body_set.size() && body_set[i]; }
    void set(uint32_t i) { if (i >= body_set.size()) body_set.resize(size_t(i) + 1, 0); body_set[i] = 1; }
    void unset(uint32_t i) { if (i < body_set.size()) body_set[i] = 0; }
};

struct Pipeline {
    uint32_t n = 0;
    std::vector<Shape> shapes;
    std::vector<uint32_t> order;          // AabbIntervals persistent order (positions -> collider index)
    std::vector<Pair> pairs;              // indexed by ContactId (stable graph edges)
    std::priority_queue<uint32_t, std::vector<uint32_t>, std::greater<uint32_t>> free_ids;  // IdPool: lowest free id
    std::unordered_map<uint64_t, uint32_t> pair_set;  // PairKey -> ContactId
    std::vector<uint32_t> active;         // ContactGraph::active_pairs order
    Color colors[AVN_GRAPH_COLOR_COUNT];
    S contact_tolerance = 0.005, length_unit = 1.0;
    // export bookkeeping: the (contact id, manifold, point) of every exported point, in column order
    struct Ref { uint32_t id, mi, pi; };
    std::vector<Ref> export_refs;
};

inline uint64_t pair_key(uint32_t a, uint32_t b) { return a < b ? (uint64_t(a) << 32) | b : (uint64_t(b) << 32) | a; }

// ---- ConstraintGraph::push_manifold / pop_manifold -----------------------------------------------------------
void push_manifold(Pipeline& P, uint32_t id) {
    Pair& pr = P.pairs[id];
    int color = AVN_COLOR_OVERFLOW;
    if (!pr.static1 && !pr.static2) {
        for (int i = 0; i < AVN_DYNAMIC_COLOR_COUNT; ++i) {
            Color& c = P.colors[i];
            if (c.get(pr.body1) || c.get(pr.body2)) continue;
            c.set(pr.body1); c.set(pr.body2); color = i; break;
        }
    } else if (!pr.static1) {
        for (int i = AVN_COLOR_OVERFLOW - 1; i >= 1; --i) {
            Color& c = P.colors[i];
            if (c.get(pr.body1)) continue;
            c.set(pr.body1); color = i; break;
        }
    } else if (!pr.static2) {
        for (int i = AVN_COLOR_OVERFLOW - 1; i >= 1; --i) {
            Color& c = P.colors[i];
            if (c.get(pr.body2)) continue;
            c.set(pr.body2); color = i; break;
        }
    }
    Color& c = P.colors[color];
    uint32_t manifold_index = uint32_t(pr.handles.size());
    pr.handles.push_back({color, uint32_t(c.handles.size())});
    c.handles.push_back({id, manifold_index});
}
void pop_manifold(Pipeline& P, uint32_t id) {
    Pair& pr = P.pairs[id];
    if (pr.handles.empty()) return;
    Handle h = pr.handles.back();
    pr.handles.pop_back();
    Color& c = P.colors[h.color];
    if (h.color != AVN_COLOR_OVERFLOW) { c.unset(pr.body1); c.unset(pr.body2); }
    uint32_t moved = uint32_t(c.handles.size()) - 1;
    c.handles[h.local] = c.handles[moved];
    c.handles.pop_back();
    if (moved != h.local) {
        auto mh = c.handles[h.local];
        P.pairs[mh.first].handles[mh.second].local = h.local;
    }
}

// The row function of the device-resident contact store (csrc/contact_rows.hpp, what csrc/contacts.cu runs one thread per row) over host
// arrays: every live row gets its manifold, matched impulses and history exactly as on the device.  Column layouts as in contacts.cu.
template <class T>
static void rows_narrow(uint32_t E, uint32_t* c1, uint32_t* c2, uint32_t* b1, uint32_t* b2, uint8_t* live, uint8_t* count, uint8_t* disjoint, void* normal, void* a1,
                        void* a2, void* pen, void* ns, uint8_t* prev_count, double* prev_a1, double* prev_a2, void* ws_n_in, void* ws_t_in, void* ws_n_out,
                        void* ws_t_out, const uint8_t* shape, const void* dims, const void* pos, const void* rot, const void* lv, const void* av,
                        const void* amin, const void* amax, double dt, double tol, double length_unit, uint32_t match) {
    avn::NarrowEdgeArgs<T> a{};
    a.r.E = int(E);
    a.r.c1 = c1; a.r.c2 = c2; a.r.b1 = b1; a.r.b2 = b2; a.r.live = live; a.r.count = count; a.r.disjoint = disjoint;
    a.r.normal = static_cast<T*>(normal); a.r.a1 = static_cast<T*>(a1); a.r.a2 = static_cast<T*>(a2); a.r.pen = static_cast<T*>(pen); a.r.ns = static_cast<T*>(ns);
    a.r.prev_count = prev_count; a.r.prev_a1 = prev_a1; a.r.prev_a2 = prev_a2;
    a.r.ws_n_in = static_cast<T*>(ws_n_in); a.r.ws_t_in = static_cast<T*>(ws_t_in); a.r.ws_n_out = static_cast<T*>(ws_n_out); a.r.ws_t_out = static_cast<T*>(ws_t_out);
    a.r.nimp_in = nullptr; a.r.nimp_out = nullptr;
    a.shape = shape; a.dims = static_cast<const T*>(dims); a.pos = static_cast<const T*>(pos); a.rot = static_cast<const T*>(rot);
    a.lv = static_cast<const T*>(lv); a.av = static_cast<const T*>(av); a.amin = static_cast<const T*>(amin); a.amax = static_cast<const T*>(amax);
    a.dt = dt; a.tol = tol; a.thr2 = (0.1 * length_unit) * (0.1 * length_unit); a.match = match ? 1 : 0;
    for (uint32_t e = 0; e < E; ++e) avn::narrow_edge_row<T>(a, int(e));
}

}  // namespace

extern "C" {

struct AvhPipeline;  // opaque = Pipeline

AvhPipeline* avh_create(uint32_t n_bodies) {
    Pipeline* p = new Pipeline();
    p->n = n_bodies;
    p->shapes.assign(n_bodies, Shape{SHAPE_CUBOID, {0.5, 0.5, 0.5}, 0.5, 0.0});
    return reinterpret_cast<AvhPipeline*>(p);
}
void avh_destroy(AvhPipeline* h) { delete reinterpret_cast<Pipeline*>(h); }

// shape_type[n], dims[n][3] (half extents, or radius in [0]), friction[n], restitution[n] — doubles
void avh_set_shapes(AvhPipeline* h, const int32_t* shape_type, const double* dims, const double* friction, const double* restitution) {
    Pipeline& P = *reinterpret_cast<Pipeline*>(h);
    for (uint32_t i = 0; i < P.n; ++i)
        P.shapes[i] = Shape{shape_type[i], {dims[3 * i], dims[3 * i + 1], dims[3 * i + 2]}, friction[i], restitution[i]};
}

// update_aabb (collider/backend.rs:498-625) for the default configuration: speculative margin = MAX, no collision
// margin, collider at the body origin.  AABB = merge(aabb(start pose), aabb(end pose)) grown by contact_tolerance.
void avh_update_aabbs(AvhPipeline* h, uint32_t scalar_bits, const void* position, const void* rotation, const void* linvel, const void* angvel,
                      double dt, void* out_min, void* out_max) {
    Pipeline& P = *reinterpret_cast<Pipeline*>(h);
    const bool f64 = scalar_bits == 64;
    Col pos{position, f64}, rt{rotation, f64}, lv{linvel, f64}, av{angvel, f64};
    ColW omin{out_min, f64}, omax{out_max, f64};
    const S tol = P.contact_tolerance * P.length_unit;
    for (uint32_t i = 0; i < P.n; ++i) {
        const Shape& sh = P.shapes[i];
        V3 p0 = pos.v3(i), v = lv.v3(i), w = av.v3(i);
        Q q0 = rt.q(i);
        Q q1 = qmul(q_from_scaled_axis(w * dt), q0);
        S l2 = q1.x * q1.x + q1.y * q1.y + q1.z * q1.z + q1.w * q1.w, k = 0.5 * (3 - l2);
        q1 = {q1.x * k, q1.y * k, q1.z * k, q1.w * k};
        V3 p1 = p0 + v * dt;
        V3 lo{1e300, 1e300, 1e300}, hi{-1e300, -1e300, -1e300};
        for (int e = 0; e < 2; ++e) {
            V3 c = e ? p1 : p0;
            V3 ext;
            if (sh.type == SHAPE_SPHERE) {
                ext = {sh.he.x, sh.he.x, sh.he.x};
            } else {
                M3 m = to_mat(e ? q1 : q0);
                ext = {std::fabs(m.c[0].x) * sh.he.x + std::fabs(m.c[1].x) * sh.he.y + std::fabs(m.c[2].x) * sh.he.z,
                       std::fabs(m.c[0].y) * sh.he.x + std::fabs(m.c[1].y) * sh.he.y + std::fabs(m.c[2].y) * sh.he.z,
                       std::fabs(m.c[0].z) * sh.he.x + std::fabs(m.c[1].z) * sh.he.y + std::fabs(m.c[2].z) * sh.he.z};
            }
            lo = {std::min(lo.x, c.x - ext.x), std::min(lo.y, c.y - ext.y), std::min(lo.z, c.z - ext.z)};
            hi = {std::max(hi.x, c.x + ext.x), std::max(hi.y, c.y + ext.y), std::max(hi.z, c.z + ext.z)};
        }
        omin.set3(i, {lo.x - tol, lo.y - tol, lo.z - tol});
        omax.set3(i, {hi.x + tol, hi.y + tol, hi.z + tol});
    }
}

// AabbIntervals persistent order (broad_phase.rs:296-315).  order[] holds collider indices.
uint32_t avh_get_order(AvhPipeline* h, uint32_t* out) {
    Pipeline& P = *reinterpret_cast<Pipeline*>(h);
    if (P.order.size() != P.n) {  // first use: add_new_aabb_intervals appends in spawn order
        P.order.resize(P.n);
        for (uint32_t i = 0; i < P.n; ++i) P.order[i] = i;
    }
    if (out) std::memcpy(out, P.order.data(), sizeof(uint32_t) * P.n);
    return P.n;
}
void avh_set_order(AvhPipeline* h, const uint32_t* order) {
    Pipeline& P = *reinterpret_cast<Pipeline*>(h);
    P.order.assign(order, order + P.n);
}

uint64_t avh_existing_pairs(AvhPipeline* h, uint64_t* out, uint64_t capacity) {
    Pipeline& P = *reinterpret_cast<Pipeline*>(h);
    uint64_t k = 0;
    for (uint32_t id : P.active) {
        if (out && k < capacity) out[k] = pair_key(P.pairs[id].collider1, P.pairs[id].collider2);
        ++k;
    }
    return k;
}

// ContactGraph::add_edge_and_key_with for each emitted pair, in list order (broad_phase.rs:443-471)
void avh_add_pairs(AvhPipeline* h, const uint32_t* c1, const uint32_t* c2, const uint32_t* b1, const uint32_t* b2, const uint8_t* flags, uint64_t count) {
    Pipeline& P = *reinterpret_cast<Pipeline*>(h);
    for (uint64_t k = 0; k < count; ++k) {
        uint64_t key = pair_key(c1[k], c2[k]);
        if (P.pair_set.count(key)) continue;
        uint32_t id;
        if (!P.free_ids.empty()) { id = P.free_ids.top(); P.free_ids.pop(); }
        else { id = uint32_t(P.pairs.size()); P.pairs.emplace_back(); }
        Pair& pr = P.pairs[id];
        pr = Pair{};
        pr.collider1 = c1[k]; pr.collider2 = c2[k]; pr.body1 = b1[k]; pr.body2 = b2[k]; pr.flags = flags[k]; pr.alive = true;
        P.pair_set[key] = id;
        P.active.push_back(id);
    }
}

// The second half of NarrowPhase::update: status changes in ascending ContactId (system_param.rs:136-389) -> ContactGraph removals and
// ConstraintGraph push / pop.  Returns the number of manifolds in the constraint graph; *out_points = their points.
static uint32_t apply_status_changes(Pipeline& P, std::vector<uint32_t>& changed, const std::vector<uint8_t>& disjoint, const std::vector<uint8_t>& started,
                                     const std::vector<uint8_t>& stopped, const std::vector<int>& count_change, uint32_t* out_points) {
    std::sort(changed.begin(), changed.end());
    for (uint32_t id : changed) {
        Pair& pr = P.pairs[id];
        const bool gen = pr.flags & AVN_PAIR_GENERATE_CONSTRAINTS;
        if (disjoint[id]) {
            if (gen) while (!pr.handles.empty()) pop_manifold(P, id);
            P.pair_set.erase(pair_key(pr.collider1, pr.collider2));
            auto it = std::find(P.active.begin(), P.active.end(), id);  // remove_edge_by_id: swap_remove (contact_graph.rs:615-628)
            if (it != P.active.end()) { *it = P.active.back(); P.active.pop_back(); }
            pr = Pair{};
            P.free_ids.push(id);
        } else if (started[id]) {
            pr.touching = true;
            if (gen) for (size_t k = 0; k < pr.manifolds.size(); ++k) push_manifold(P, id);
        } else if (stopped[id]) {
            pr.touching = false;
            if (gen) while (!pr.handles.empty()) pop_manifold(P, id);
        } else if (pr.touching && gen && count_change[id] > 0) {
            for (int k = 0; k < count_change[id]; ++k) push_manifold(P, id);
        } else if (pr.touching && gen && count_change[id] < 0) {
            for (int k = 0; k < -count_change[id]; ++k) pop_manifold(P, id);
        }
    }
    uint32_t M = 0, Pn = 0;
    for (int c = 0; c < AVN_GRAPH_COLOR_COUNT; ++c)
        for (auto& hnd : P.colors[c].handles) { ++M; Pn += uint32_t(P.pairs[hnd.first].manifolds[hnd.second].pts.size()); }
    if (out_points) *out_points = Pn;
    return M;
}

// NarrowPhase::update (narrow_phase/system_param.rs:114-400) with the fixture manifold generator.
// kind[n] = AvnBodyKind.  Returns the number of exported manifolds; *out_points = number of points.
uint32_t avh_narrow_phase(AvhPipeline* h, uint32_t scalar_bits, const uint8_t* kind, const void* position, const void* rotation, const void* linvel,
                          const void* angvel, const void* aabb_min, const void* aabb_max, double dt, uint32_t match_contacts, uint32_t* out_points) {
    Pipeline& P = *reinterpret_cast<Pipeline*>(h);
    const bool f64 = scalar_bits == 64;
    Col pos{position, f64}, rt{rotation, f64}, lv{linvel, f64}, av{angvel, f64}, amin{aabb_min, f64}, amax{aabb_max, f64};
    const S tol = P.contact_tolerance * P.length_unit;
    std::vector<uint32_t> changed;  // contact ids whose status changed
    std::vector<uint8_t> disjoint(P.pairs.size(), 0), started(P.pairs.size(), 0), stopped(P.pairs.size(), 0);
    std::vector<int> count_change(P.pairs.size(), 0);
    Contacts pts;
    for (uint32_t id : P.active) {
        Pair& pr = P.pairs[id];
        const uint32_t a = pr.collider1, b = pr.collider2;
        V3 mina = amin.v3(a), maxa = amax.v3(a), minb = amin.v3(b), maxb = amax.v3(b);
        bool overlap = !(mina.x > maxb.x || maxa.x < minb.x || mina.y > maxb.y || maxa.y < minb.y || mina.z > maxb.z || maxa.z < minb.z);
        if (!overlap) { disjoint[id] = 1; changed.push_back(id); continue; }
        pr.static1 = kind[pr.body1] == AVN_BODY_STATIC;
        pr.static2 = kind[pr.body2] == AVN_BODY_STATIC;
        const Shape& sa = P.shapes[a];
        const Shape& sb = P.shapes[b];
        V3 v1 = lv.v3(pr.body1), v2 = lv.v3(pr.body2), w1 = av.v3(pr.body1), w2 = av.v3(pr.body2);
        V3 rel = v2 - v1;
        S eff_margin = dt * len(rel);  // effective speculative margin (system_param.rs:663-681) with margin = MAX
        S max_dist = std::max(eff_margin, tol);
        std::vector<Manifold> old = std::move(pr.manifolds);
        pr.manifolds.clear();
        V3 normal;
        V3 pa = pos.v3(a), pb = pos.v3(b);
        const bool hit = collide(sa.type, sa.he, pa, rt.q(a), sb.type, sb.he, pb, rt.q(b), max_dist, normal, pts);
        if (hit) {
            PointOut out[4];
            const int np = manifold_points(pts, normal, pa, pb, rel, w1, w2, dt, eff_margin, out);
            if (np > 0) {
                Manifold m;
                m.normal = normal;
                for (int k = 0; k < np; ++k) {
                    Point pt;
                    pt.anchor1 = out[k].anchor1;
                    pt.anchor2 = out[k].anchor2;
                    pt.penetration = out[k].penetration;
                    pt.normal_speed = out[k].normal_speed;
                    m.pts.push_back(pt);
                }
                pr.manifolds.push_back(std::move(m));
            }
        }
        bool touching = !pr.manifolds.empty();
        if (touching && match_contacts && pr.manifolds.size() <= 4) {
            // ContactManifold::match_contacts with unknown feature ids (contact_types/mod.rs:426-470)
            const S thr2 = (0.1 * P.length_unit) * (0.1 * P.length_unit);
            for (Manifold& m : pr.manifolds)
                for (const Manifold& om : old) {
                    V3 oa1[8], oa2[8];
                    const int n_old = int(std::min<size_t>(om.pts.size(), 8));
                    for (int k = 0; k < n_old; ++k) { oa1[k] = om.pts[k].anchor1; oa2[k] = om.pts[k].anchor2; }
                    for (Point& c : m.pts) {
                        const int k = match_point(c.anchor1, c.anchor2, oa1, oa2, n_old, thr2);
                        if (k >= 0) { c.ws_normal = om.pts[k].ws_normal; c.ws_tx = om.pts[k].ws_tx; c.ws_ty = om.pts[k].ws_ty; }
                    }
                }
        }
        count_change[id] = int(pr.manifolds.size()) - int(old.size());
        if (touching && !pr.touching) { started[id] = 1; changed.push_back(id); }
        else if (!touching && pr.touching) { stopped[id] = 1; changed.push_back(id); }
        else if (count_change[id] != 0) changed.push_back(id);
    }
    return apply_status_changes(P, changed, disjoint, started, stopped, count_change, out_points);
}

// Export the manifolds grouped by colour in manifold_handles order (what prepare_contact_constraints walks,
// solver/plugin.rs:389-434).  Arrays are sized from avh_narrow_phase's return values.
void avh_export_manifolds(AvhPipeline* h, uint32_t scalar_bits, uint32_t* color_offsets /*[25]*/, int32_t* body1, int32_t* body2, void* normal,
                          void* friction, void* restitution, uint32_t* point_offsets, void* anchor1, void* anchor2, void* penetration,
                          void* normal_speed, void* ws_normal, void* ws_tangent) {
    Pipeline& P = *reinterpret_cast<Pipeline*>(h);
    const bool f64 = scalar_bits == 64;
    ColW wn{normal, f64}, wf{friction, f64}, wr{restitution, f64}, wa1{anchor1, f64}, wa2{anchor2, f64}, wp{penetration, f64}, ws{normal_speed, f64},
        wwn{ws_normal, f64}, wwt{ws_tangent, f64};
    P.export_refs.clear();
    uint32_t m = 0, p = 0;
    for (int c = 0; c < AVN_GRAPH_COLOR_COUNT; ++c) {
        color_offsets[c] = m;
        for (auto& hnd : P.colors[c].handles) {
            const Pair& pr = P.pairs[hnd.first];
            const Manifold& mf = pr.manifolds[hnd.second];
            body1[m] = int32_t(pr.body1);
            body2[m] = int32_t(pr.body2);
            wn.set3(m, mf.normal);
            wf.set(m, (P.shapes[pr.collider1].friction + P.shapes[pr.collider2].friction) * 0.5);
            wr.set(m, (P.shapes[pr.collider1].restitution + P.shapes[pr.collider2].restitution) * 0.5);
            point_offsets[m] = p;
            for (uint32_t k = 0; k < mf.pts.size(); ++k) {
                const Point& pt = mf.pts[k];
                wa1.set3(p, pt.anchor1); wa2.set3(p, pt.anchor2);
                wp.set(p, pt.penetration); ws.set(p, pt.normal_speed);
                wwn.set(p, pt.ws_normal); wwt.set(2 * p, pt.ws_tx); wwt.set(2 * p + 1, pt.ws_ty);
                P.export_refs.push_back({hnd.first, hnd.second, k});
                ++p;
            }
            ++m;
        }
    }
    color_offsets[AVN_GRAPH_COLOR_COUNT] = m;
    point_offsets[m] = p;
}

// store_contact_impulses' destination: ContactPoint::{warm_start_*, normal_impulse} (solver/plugin.rs:741-750)
void avh_store_impulses(AvhPipeline* h, uint32_t scalar_bits, const void* ws_normal, const void* ws_tangent, const void* normal_impulse) {
    Pipeline& P = *reinterpret_cast<Pipeline*>(h);
    const bool f64 = scalar_bits == 64;
    Col n{ws_normal, f64}, t{ws_tangent, f64}, ni{normal_impulse, f64};
    for (size_t p = 0; p < P.export_refs.size(); ++p) {
        auto r = P.export_refs[p];
        Point& pt = P.pairs[r.id].manifolds[r.mi].pts[r.pi];
        pt.ws_normal = n.at(p); pt.ws_tx = t.at(2 * p); pt.ws_ty = t.at(2 * p + 1); pt.normal_impulse = ni.at(p);
    }
}

// The geometry stage alone for an explicit list of pairs, in the layout of AvnRawManifolds (4 point slots per pair): the CPU side of the
// device narrow-phase test.  scalar_bits selects the column type; evaluation is in double either way (csrc/narrow_math.hpp).
void avh_raw_manifolds(uint32_t scalar_bits, uint32_t pair_count, const uint32_t* c1, const uint32_t* c2, const uint32_t* b1, const uint32_t* b2,
                       const uint8_t* shape, const void* dims, const void* position, const void* rotation, const void* linvel, const void* angvel,
                       const void* aabb_min, const void* aabb_max, double dt, double tol, uint8_t* point_count, uint8_t* disjoint, void* normal,
                       void* anchor1, void* anchor2, void* penetration, void* normal_speed, double* anchor1_f64, double* anchor2_f64) {
    const bool f64 = scalar_bits == 64;
    Col dm{dims, f64}, pos{position, f64}, rt{rotation, f64}, lv{linvel, f64}, av{angvel, f64}, amin{aabb_min, f64}, amax{aabb_max, f64};
    ColW on{normal, f64}, oa1{anchor1, f64}, oa2{anchor2, f64}, op{penetration, f64}, os{normal_speed, f64};
    for (uint32_t k = 0; k < pair_count; ++k) {
        const uint32_t a = c1[k], b = c2[k];
        point_count[k] = 0;
        on.set3(k, V3{0, 0, 0});
        for (int p = 0; p < 4; ++p) { oa1.set3(4 * size_t(k) + p, V3{0, 0, 0}); oa2.set3(4 * size_t(k) + p, V3{0, 0, 0}); op.set(4 * size_t(k) + p, 0); os.set(4 * size_t(k) + p, 0); }
        if (aabb_min) {
            V3 mina = amin.v3(a), maxa = amax.v3(a), minb = amin.v3(b), maxb = amax.v3(b);
            bool overlap = !(mina.x > maxb.x || maxa.x < minb.x || mina.y > maxb.y || maxa.y < minb.y || mina.z > maxb.z || maxa.z < minb.z);
            if (disjoint) disjoint[k] = overlap ? 0 : 1;
            if (!overlap) continue;
        } else if (disjoint) {
            disjoint[k] = 0;
        }
        V3 pa = pos.v3(a), pb = pos.v3(b);
        V3 v1 = lv.v3(b1[k]), v2 = lv.v3(b2[k]), w1 = av.v3(b1[k]), w2 = av.v3(b2[k]);
        V3 rel = v2 - v1;
        S eff_margin = dt * len(rel);
        S max_dist = smax(eff_margin, tol);
        V3 nrm;
        Contacts pts;
        int ta = shape ? shape[a] : SHAPE_CUBOID, tb = shape ? shape[b] : SHAPE_CUBOID;
        if (!collide(ta, dm.v3(a), pa, rt.q(a), tb, dm.v3(b), pb, rt.q(b), max_dist, nrm, pts)) continue;
        PointOut out[4];
        int np = manifold_points(pts, nrm, pa, pb, rel, w1, w2, dt, eff_margin, out);
        point_count[k] = uint8_t(np);
        on.set3(k, nrm);
        for (int p = 0; p < np; ++p) {
            oa1.set3(4 * size_t(k) + p, out[p].anchor1);
            oa2.set3(4 * size_t(k) + p, out[p].anchor2);
            if (anchor1_f64)   // unrounded anchors: what the next step's match_contacts compares (the fixture matches in double)
                for (int c = 0; c < 3; ++c) {
                    anchor1_f64[(4 * size_t(k) + p) * 3 + c] = comp(out[p].anchor1, c);
                    anchor2_f64[(4 * size_t(k) + p) * 3 + c] = comp(out[p].anchor2, c);
                }
            op.set(4 * size_t(k) + p, out[p].penetration);
            os.set(4 * size_t(k) + p, out[p].normal_speed);
        }
    }
}

// ---- resident mode: the geometry runs elsewhere (avn_narrow_phase), the host keeps only the graphs -------------
// The active contact edges in ContactGraph order: edge id (ContactId), colliders, bodies.  Arrays sized avh_pair_count().
uint32_t avh_active_edges(AvhPipeline* h, uint32_t* ids, uint32_t* c1, uint32_t* c2, uint32_t* b1, uint32_t* b2) {
    Pipeline& P = *reinterpret_cast<Pipeline*>(h);
    uint32_t n = 0;
    for (uint32_t id : P.active) {
        const Pair& pr = P.pairs[id];
        ids[n] = id; c1[n] = pr.collider1; c2[n] = pr.collider2; b1[n] = pr.body1; b2[n] = pr.body2;
        ++n;
    }
    return n;
}

// The host half of the narrow phase from the per-edge results computed elsewhere: point_count[i] (0 = not touching) and disjoint[i] for
// edge ids[i].  Same status machine, same graph updates as avh_narrow_phase; the manifolds only remember how many points they have.
uint32_t avh_apply_counts(AvhPipeline* h, const uint8_t* kind, const uint32_t* ids, const uint8_t* point_count, const uint8_t* disjoint_in, uint32_t n,
                          uint32_t* out_points) {
    Pipeline& P = *reinterpret_cast<Pipeline*>(h);
    std::vector<uint32_t> changed;
    std::vector<uint8_t> disjoint(P.pairs.size(), 0), started(P.pairs.size(), 0), stopped(P.pairs.size(), 0);
    std::vector<int> count_change(P.pairs.size(), 0);
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t id = ids[i];
        Pair& pr = P.pairs[id];
        if (disjoint_in[i]) { disjoint[id] = 1; changed.push_back(id); continue; }
        pr.static1 = kind[pr.body1] == AVN_BODY_STATIC;
        pr.static2 = kind[pr.body2] == AVN_BODY_STATIC;
        const size_t old_count = pr.manifolds.size();
        pr.manifolds.clear();
        if (point_count[i] > 0) {
            Manifold m;
            m.normal = V3{0, 0, 0};
            m.pts.resize(point_count[i]);
            pr.manifolds.push_back(std::move(m));
        }
        const bool touching = !pr.manifolds.empty();
        count_change[id] = int(pr.manifolds.size()) - int(old_count);
        if (touching && !pr.touching) { started[id] = 1; changed.push_back(id); }
        else if (!touching && pr.touching) { stopped[id] = 1; changed.push_back(id); }
        else if (count_change[id] != 0) changed.push_back(id);
    }
    return apply_status_changes(P, changed, disjoint, started, stopped, count_change, out_points);
}

// The constraint graph as a colour-major list of edge ids (manifold_handles order) with the per-edge material (what the solver's prepare
// needs besides the geometry).  Arrays sized from avh_apply_counts' return value.
void avh_export_edges(AvhPipeline* h, uint32_t* color_offsets /*[25]*/, uint32_t* edge, int32_t* body1, int32_t* body2, double* friction, double* restitution) {
    Pipeline& P = *reinterpret_cast<Pipeline*>(h);
    uint32_t m = 0;
    for (int c = 0; c < AVN_GRAPH_COLOR_COUNT; ++c) {
        color_offsets[c] = m;
        for (auto& hnd : P.colors[c].handles) {
            const Pair& pr = P.pairs[hnd.first];
            edge[m] = hnd.first;
            body1[m] = int32_t(pr.body1);
            body2[m] = int32_t(pr.body2);
            friction[m] = (P.shapes[pr.collider1].friction + P.shapes[pr.collider2].friction) * 0.5;
            restitution[m] = (P.shapes[pr.collider1].restitution + P.shapes[pr.collider2].restitution) * 0.5;
            ++m;
        }
    }
    color_offsets[AVN_GRAPH_COLOR_COUNT] = m;
}

// match_contacts on edge-indexed resident state (what the device keeps between steps): for edge ids[i] with new_count[i] points whose anchors
// (double, unrounded: the fixture matches in double) are new_a1/new_a2[i][4][3], carry the warm-start impulses over from the matching old
// points and make the new points the resident ones.  ws_n[E][4], ws_t[E][4][2] in the column scalar type.
void avh_match_raw(uint32_t scalar_bits, uint32_t n, const uint32_t* ids, const uint8_t* new_count, const double* new_a1, const double* new_a2,
                   double length_unit, uint32_t match_contacts, uint8_t* prev_count, double* prev_a1, double* prev_a2, void* ws_n, void* ws_t) {
    const bool f64 = scalar_bits == 64;
    Col rn{ws_n, f64}, rtg{ws_t, f64};
    ColW wn{ws_n, f64}, wt{ws_t, f64};
    const S thr2 = (0.1 * length_unit) * (0.1 * length_unit);
    for (uint32_t i = 0; i < n; ++i) {
        const size_t e = ids[i];
        const int nc = new_count[i], oc = prev_count[e];
        V3 oa1[4], oa2[4];
        S on[4], otx[4], oty[4];
        for (int k = 0; k < oc; ++k) {
            oa1[k] = V3{prev_a1[(e * 4 + k) * 3], prev_a1[(e * 4 + k) * 3 + 1], prev_a1[(e * 4 + k) * 3 + 2]};
            oa2[k] = V3{prev_a2[(e * 4 + k) * 3], prev_a2[(e * 4 + k) * 3 + 1], prev_a2[(e * 4 + k) * 3 + 2]};
            on[k] = rn.at(e * 4 + k); otx[k] = rtg.at((e * 4 + k) * 2); oty[k] = rtg.at((e * 4 + k) * 2 + 1);
        }
        for (int k = 0; k < 4; ++k) {
            S vn = 0, vx = 0, vy = 0;
            if (k < nc) {
                const V3 a1{new_a1[(size_t(i) * 4 + k) * 3], new_a1[(size_t(i) * 4 + k) * 3 + 1], new_a1[(size_t(i) * 4 + k) * 3 + 2]};
                const V3 a2{new_a2[(size_t(i) * 4 + k) * 3], new_a2[(size_t(i) * 4 + k) * 3 + 1], new_a2[(size_t(i) * 4 + k) * 3 + 2]};
                const int j = match_contacts ? match_point(a1, a2, oa1, oa2, oc, thr2) : -1;
                if (j >= 0) { vn = on[j]; vx = otx[j]; vy = oty[j]; }
                for (int c = 0; c < 3; ++c) { prev_a1[(e * 4 + k) * 3 + c] = comp(a1, c); prev_a2[(e * 4 + k) * 3 + c] = comp(a2, c); }
            }
            wn.set(e * 4 + k, vn);
            wt.set((e * 4 + k) * 2, vx);
            wt.set((e * 4 + k) * 2 + 1, vy);
        }
        prev_count[e] = uint8_t(nc);
    }
}

// csrc/contact_rows.hpp over host arrays (see rows_narrow above)
void avh_rows_narrow(uint32_t scalar_bits, uint32_t E, uint32_t* c1, uint32_t* c2, uint32_t* b1, uint32_t* b2, uint8_t* live, uint8_t* count, uint8_t* disjoint,
                     void* normal, void* a1, void* a2, void* pen, void* ns, uint8_t* prev_count, double* prev_a1, double* prev_a2, void* ws_n_in, void* ws_t_in,
                     void* ws_n_out, void* ws_t_out, const uint8_t* shape, const void* dims, const void* pos, const void* rot, const void* lv, const void* av,
                     const void* amin, const void* amax, double dt, double tol, double length_unit, uint32_t match) {
    if (scalar_bits == 64)
        rows_narrow<double>(E, c1, c2, b1, b2, live, count, disjoint, normal, a1, a2, pen, ns, prev_count, prev_a1, prev_a2, ws_n_in, ws_t_in, ws_n_out, ws_t_out,
                            shape, dims, pos, rot, lv, av, amin, amax, dt, tol, length_unit, match);
    else
        rows_narrow<float>(E, c1, c2, b1, b2, live, count, disjoint, normal, a1, a2, pen, ns, prev_count, prev_a1, prev_a2, ws_n_in, ws_t_in, ws_n_out, ws_t_out,
                           shape, dims, pos, rot, lv, av, amin, amax, dt, tol, length_unit, match);
}

uint32_t avh_pair_count(AvhPipeline* h) { return uint32_t(reinterpret_cast<Pipeline*>(h)->active.size()); }

}  // extern "C"
