// Contact-manifold geometry for cuboid / sphere pairs, written once for the host fixture (g++, -ffp-contract=off) and for the device
// (nvcc, -fmad=false): the same expressions in the same order, IEEE double throughout, so both evaluate to the same bits.
//
// What this is: OUR manifold generator (SAT + face clipping for boxes, closed forms for spheres).  The reference delegates this
// arithmetic to parry3d 0.25, which is not vendored, so there is no parity claim against parry — the claim is
// fixture-level: whatever consumes these manifolds (oracle or CUDA solver) gets identical inputs.
// Reference call sites: collider/parry/contact_query.rs:156-261 (contact_manifolds), contact_types/mod.rs:478-566 (prune_points),
// narrow_phase/system_param.rs:663-681,736-756 (speculative margin, point keep rule).
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define NM_HD __host__ __device__
#else
#define NM_HD
#endif

namespace nm {

using S = double;  // geometry is evaluated in double and rounded to the column scalar type on export

struct V3 { S x, y, z; };
NM_HD inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
NM_HD inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
NM_HD inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
NM_HD inline V3 operator*(V3 a, S s) { return {a.x * s, a.y * s, a.z * s}; }
NM_HD inline S dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
NM_HD inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
NM_HD inline S len(V3 a) { return sqrt(dot(a, a)); }
NM_HD inline S comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
// std::max / std::min semantics (the first argument wins ties), usable on the device
NM_HD inline S smax(S a, S b) { return a < b ? b : a; }
NM_HD inline S smin(S a, S b) { return b < a ? b : a; }

struct Q { S x, y, z, w; };
NM_HD inline V3 rot(Q q, V3 v) {
    V3 b{q.x, q.y, q.z};
    S b2 = dot(b, b);
    return v * (q.w * q.w - b2) + b * (dot(v, b) * 2) + cross(b, v) * (q.w * 2);
}
struct M3 { V3 c[3]; };  // columns = world directions of the local axes
NM_HD inline M3 to_mat(Q q) { return {{rot(q, {1, 0, 0}), rot(q, {0, 1, 0}), rot(q, {0, 0, 1})}}; }

enum ShapeType { SHAPE_CUBOID = 0, SHAPE_SPHERE = 1 };
struct Box { V3 c; M3 r; V3 he; };

// witness points of a contact: on shape A, on shape B (world space).  A quad clipped by four planes has at most 8 vertices.
constexpr int MAX_RAW_POINTS = 8;
struct Witness { V3 a, b; };
struct Contacts {
    int n;
    Witness p[MAX_RAW_POINTS];
    NM_HD void clear() { n = 0; }
    NM_HD void push(V3 on_a, V3 on_b) { p[n].a = on_a; p[n].b = on_b; ++n; }
};

// the face of `b` on the `sign` side of local axis `axis`: 4 corners (world)
NM_HD inline void box_face(const Box& b, int axis, S sign, V3 out[4]) {
    int u = (axis + 1) % 3, v = (axis + 2) % 3;
    V3 n = b.r.c[axis] * (sign * comp(b.he, axis));
    V3 eu = b.r.c[u] * comp(b.he, u), ev = b.r.c[v] * comp(b.he, v);
    V3 fc = b.c + n;
    out[0] = fc + eu + ev; out[1] = fc - eu + ev; out[2] = fc - eu - ev; out[3] = fc + eu - ev;
}

NM_HD inline int clip_poly(const V3* in, int n, V3 plane_n, S plane_d, V3* out) {  // keep dot(n,p) <= d
    int m = 0;
    for (int i = 0; i < n; ++i) {
        V3 a = in[i], b = in[(i + 1) % n];
        S da = dot(plane_n, a) - plane_d, db = dot(plane_n, b) - plane_d;
        if (da <= 0) out[m++] = a;
        if ((da < 0 && db > 0) || (da > 0 && db < 0)) out[m++] = a + (b - a) * (da / (da - db));
    }
    return m;
}

NM_HD inline S box_radius(const Box& b, V3 n) {
    return fabs(dot(b.r.c[0], n)) * b.he.x + fabs(dot(b.r.c[1], n)) * b.he.y + fabs(dot(b.r.c[2], n)) * b.he.z;
}

// SAT over the 15 axes, then either the closest points of the two supporting edges or the incident face clipped against the
// reference face's side planes.  Returns false when the boxes are farther apart than max_dist.  normal points from A to B.
NM_HD inline bool box_box(const Box& A, const Box& B, S max_dist, V3& normal, Contacts& pts) {
    pts.clear();
    V3 d = B.c - A.c;
    S best_sep = -1e300;
    int best_kind = -1, best_i = 0, best_j = 0;
    V3 best_n{0, 1, 0};
    auto consider = [&](V3 n, int kind, int i, int j, S bias) {
        S l = len(n);
        if (l < 1e-9) return;
        n = n * (1 / l);
        if (dot(n, d) < 0) n = -n;
        S sep = dot(n, d) - box_radius(A, n) - box_radius(B, n);
        // face axes are preferred over edge axes by a small bias, earlier axes win ties: stable feature choice
        if (sep - bias > best_sep + 1e-9) { best_sep = sep - bias; best_kind = kind; best_i = i; best_j = j; best_n = n; }
    };
    for (int i = 0; i < 3; ++i) consider(A.r.c[i], 0, i, 0, 0);
    for (int i = 0; i < 3; ++i) consider(B.r.c[i], 1, i, 0, 0);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) consider(cross(A.r.c[i], B.r.c[j]), 2, i, j, 1e-4);
    S sep = best_kind == 2 ? best_sep + 1e-4 : best_sep;
    if (sep > max_dist) return false;
    normal = best_n;
    if (best_kind == 2) {
        // edge-edge: closest points of the two supporting edges
        V3 ea = A.r.c[best_i], eb = B.r.c[best_j];
        V3 pa = A.c, pb = B.c;
        for (int k = 0; k < 3; ++k) {
            if (k != best_i) pa = pa + A.r.c[k] * (comp(A.he, k) * (dot(A.r.c[k], normal) > 0 ? 1 : -1));
            if (k != best_j) pb = pb + B.r.c[k] * (comp(B.he, k) * (dot(B.r.c[k], normal) < 0 ? 1 : -1));
        }
        V3 r = pa - pb;
        S a = dot(ea, ea), e = dot(eb, eb), f = dot(eb, r), c = dot(ea, r), b = dot(ea, eb);
        S den = a * e - b * b;
        S s = den > 1e-12 ? (b * f - c * e) / den : 0;
        S t = (b * s + f) / e;
        S ha = comp(A.he, best_i), hb = comp(B.he, best_j);
        s = smax(-ha, smin(ha, s));
        t = smax(-hb, smin(hb, t));
        pts.push(pa + ea * s, pb + eb * t);
        return true;
    }
    // face contact: reference box R (face axis), incident box I
    const bool ref_is_a = best_kind == 0;
    const Box& R = ref_is_a ? A : B;
    const Box& I = ref_is_a ? B : A;
    V3 rn = ref_is_a ? normal : -normal;  // outward normal of the reference face
    // incident face: the face of I most anti-parallel to rn
    int inc_axis = 0;
    S inc_best = -1;
    for (int k = 0; k < 3; ++k) {
        S v = fabs(dot(I.r.c[k], rn));
        if (v > inc_best) { inc_best = v; inc_axis = k; }
    }
    S isign = dot(I.r.c[inc_axis], rn) > 0 ? -1 : 1;
    V3 poly[16], tmp[16];
    box_face(I, inc_axis, isign, poly);
    int np = 4;
    int u = (best_i + 1) % 3, v = (best_i + 2) % 3;
    const int axes[2] = {u, v};
    for (int a = 0; a < 2 && np > 0; ++a) {
        V3 sn = R.r.c[axes[a]];
        S he = comp(R.he, axes[a]);
        np = clip_poly(poly, np, sn, dot(sn, R.c) + he, tmp);
        np = clip_poly(tmp, np, -sn, -dot(sn, R.c) + he, poly);
    }
    S face_d = dot(rn, R.c) + comp(R.he, best_i) * 1.0;
    for (int k = 0; k < np; ++k) {
        S dist = dot(rn, poly[k]) - face_d;  // signed distance of the incident point above the reference face
        if (dist > max_dist) continue;
        V3 on_ref = poly[k] - rn * dist;
        // drop near-duplicates
        bool dup = false;
        for (int q = 0; q < pts.n; ++q) {
            V3 e = (ref_is_a ? pts.p[q].b : pts.p[q].a) - poly[k];
            if (dot(e, e) < 1e-12) { dup = true; break; }
        }
        if (dup) continue;
        if (pts.n == MAX_RAW_POINTS) break;
        if (ref_is_a) pts.push(on_ref, poly[k]); else pts.push(poly[k], on_ref);
    }
    return pts.n != 0;
}

// keep at most 4 points: deepest, farthest from it, farthest from that segment on each side (cf. prune_points,
// contact_types/mod.rs:478-566, itself after Jolt's PruneContactPoints); survivors stay in their original order
NM_HD inline void prune4(Contacts& pts, V3 n) {
    if (pts.n <= 4) return;
    auto depth = [&](int i) { return dot(pts.p[i].a - pts.p[i].b, n); };
    int p0 = 0;
    for (int i = 1; i < pts.n; ++i) if (depth(i) > depth(p0) + 1e-12) p0 = i;
    int p1 = p0;
    S best = -1;
    for (int i = 0; i < pts.n; ++i) { V3 e = pts.p[i].a - pts.p[p0].a; S v = dot(e, e); if (v > best) { best = v; p1 = i; } }
    V3 dir = cross(pts.p[p1].a - pts.p[p0].a, n);
    int p2 = p0, p3 = p0;
    S mx = 0, mn = 0;
    for (int i = 0; i < pts.n; ++i) {
        S v = dot(pts.p[i].a - pts.p[p0].a, dir);
        if (v > mx) { mx = v; p2 = i; }
        if (v < mn) { mn = v; p3 = i; }
    }
    bool keep[MAX_RAW_POINTS] = {false, false, false, false, false, false, false, false};
    keep[p0] = keep[p1] = keep[p2] = keep[p3] = true;
    int m = 0;
    for (int i = 0; i < pts.n; ++i)
        if (keep[i]) pts.p[m++] = pts.p[i];
    pts.n = m;
}

NM_HD inline bool sphere_sphere(V3 ca, S ra, V3 cb, S rb, S max_dist, V3& normal, Contacts& pts) {
    pts.clear();
    V3 d = cb - ca;
    S l = len(d);
    if (l - ra - rb > max_dist) return false;
    normal = l > 1e-12 ? d * (1 / l) : V3{0, 1, 0};
    pts.push(ca + normal * ra, cb - normal * rb);
    return true;
}

NM_HD inline bool box_sphere(const Box& A, V3 cs, S rs, S max_dist, V3& normal, Contacts& pts) {  // normal from box to sphere
    pts.clear();
    V3 d = cs - A.c;
    V3 local{dot(d, A.r.c[0]), dot(d, A.r.c[1]), dot(d, A.r.c[2])};
    V3 cl{smax(-A.he.x, smin(A.he.x, local.x)), smax(-A.he.y, smin(A.he.y, local.y)), smax(-A.he.z, smin(A.he.z, local.z))};
    V3 on_box = A.c + A.r.c[0] * cl.x + A.r.c[1] * cl.y + A.r.c[2] * cl.z;
    V3 e = cs - on_box;
    S l = len(e);
    if (l > 1e-9) {
        if (l - rs > max_dist) return false;
        normal = e * (1 / l);
    } else {  // centre inside the box: push out through the nearest face
        int ax = 0; S best = 1e300;
        for (int k = 0; k < 3; ++k) { S v = comp(A.he, k) - fabs(comp(local, k)); if (v < best) { best = v; ax = k; } }
        S sgn = comp(local, ax) >= 0 ? 1 : -1;
        normal = A.r.c[ax] * sgn;
        on_box = cs + normal * best;
    }
    pts.push(on_box, cs - normal * rs);
    return true;
}

// One collider pair -> normal (from A to B) and at most 4 witness pairs.  he = half extents of a cuboid, radius in he.x of a sphere.
NM_HD inline bool collide(int type_a, V3 he_a, V3 pa, Q qa, int type_b, V3 he_b, V3 pb, Q qb, S max_dist, V3& normal, Contacts& pts) {
    bool hit;
    if (type_a == SHAPE_CUBOID && type_b == SHAPE_CUBOID) {
        Box A{pa, to_mat(qa), he_a}, B{pb, to_mat(qb), he_b};
        hit = box_box(A, B, max_dist, normal, pts);
        if (hit) prune4(pts, normal);
    } else if (type_a == SHAPE_SPHERE && type_b == SHAPE_SPHERE) {
        hit = sphere_sphere(pa, he_a.x, pb, he_b.x, max_dist, normal, pts);
    } else if (type_a == SHAPE_CUBOID) {
        Box A{pa, to_mat(qa), he_a};
        hit = box_sphere(A, pb, he_b.x, max_dist, normal, pts);
    } else {
        Box B{pb, to_mat(qb), he_b};
        hit = box_sphere(B, pa, he_a.x, max_dist, normal, pts);
        if (hit) {
            normal = -normal;
            for (int k = 0; k < pts.n; ++k) { V3 t = pts.p[k].a; pts.p[k].a = pts.p[k].b; pts.p[k].b = t; }
        }
    }
    return hit;
}

// A manifold point as the solver's input columns want it (ContactPoint, contact_types/mod.rs:603-660): anchors relative to the body
// origins (collider at the body origin, centre of mass at the origin), penetration, normal speed.
struct PointOut { V3 anchor1, anchor2; S penetration, normal_speed; };

// From the witness pairs to the manifold's points: the speculative keep rule of narrow_phase/system_param.rs:748-756.
// rel = v2 - v1, eff_margin = dt * |rel| (margin = MAX).  Returns the number of points written to out[0..4).
NM_HD inline int manifold_points(const Contacts& pts, V3 normal, V3 pa, V3 pb, V3 rel, V3 w1, V3 w2, S dt, S eff_margin, PointOut out[4]) {
    int m = 0;
    for (int k = 0; k < pts.n && m < 4; ++k) {
        PointOut pt;
        pt.anchor1 = pts.p[k].a - pa;
        pt.anchor2 = pts.p[k].b - pb;
        pt.penetration = dot(pts.p[k].a - pts.p[k].b, normal);
        V3 rv = rel + cross(w2, pt.anchor2) - cross(w1, pt.anchor1);
        pt.normal_speed = dot(rv, normal);
        bool keep = -pt.penetration < eff_margin || (pt.normal_speed * dt - pt.penetration < eff_margin);
        if (!keep) continue;
        out[m++] = pt;
    }
    return m;
}

// ContactManifold::match_contacts with unknown feature ids (contact_types/mod.rs:426-470): a new point inherits the warm-start impulses
// of the first old point whose two anchors both lie within the distance threshold (squared: thr2) of its own, in either body order.
// Returns the index of that old point or -1.
NM_HD inline int match_point(V3 anchor1, V3 anchor2, const V3* old_anchor1, const V3* old_anchor2, int n_old, S thr2) {
    for (int k = 0; k < n_old; ++k) {
        V3 e11 = anchor1 - old_anchor1[k], e22 = anchor2 - old_anchor2[k], e12 = anchor1 - old_anchor2[k], e21 = anchor2 - old_anchor1[k];
        if ((dot(e11, e11) < thr2 && dot(e22, e22) < thr2) || (dot(e12, e12) < thr2 && dot(e21, e21) < thr2)) return k;
    }
    return -1;
}

}  // namespace nm
