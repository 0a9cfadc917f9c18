// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
// CPU restatement (oracle) of the avian3d solver stage: solver bodies, integrator, TGS-soft contacts,
// XPBD joints.  Serial, reference operation order, built with -ffp-contract=off.
//
// PARITY UNPINNED: the reference cannot be compiled here (no Rust toolchain) and holds no
// 3D golden vectors for contacts/joints; this file follows the reference line by line and is validated by
// the restated reference tests (integrator/mod.rs:561-629) and analytic invariants in tests/.
#pragma once
#include <algorithm>
#include <cstring>
#include <vector>

#include "../include/avian_b200.h"
#include "oracle_math.hpp"

namespace orc {

// ---- solver_body/mod.rs:59-104 ------------------------------------------------------------------------
template <class S>
struct SolverBody {
    V3<S> linear_velocity{0, 0, 0};
    V3<S> angular_velocity{0, 0, 0};
    V3<S> delta_position{0, 0, 0};
    Quat<S> delta_rotation{0, 0, 0, 1};
    uint32_t flags = 0;
    bool is_kinematic() const { return flags & (1u << 6); }
    bool is_gyroscopic() const { return flags & (1u << 7); }
    V3<S> velocity_at_point(V3<S> p) const { return linear_velocity + cross(angular_velocity, p); }
};
// solver_body/mod.rs:218-451
template <class S>
struct SolverBodyInertia {
    S inv_mass = 0;
    Sym3<S> inv_inertia = sym3_zero<S>();
    int16_t dominance = 128;  // DUMMY: i8::MAX + 1
    uint16_t flags = (1u << 6) | (1u << 7);
    V3<S> effective_inv_mass() const {
        V3<S> m{inv_mass, inv_mass, inv_mass};
        if (flags & AVN_LOCK_TRANSLATION_X) m.x = 0;
        if (flags & AVN_LOCK_TRANSLATION_Y) m.y = 0;
        if (flags & AVN_LOCK_TRANSLATION_Z) m.z = 0;
        return m;
    }
};
// solver_body/mod.rs:378-423 SolverBodyInertia::new (3D)
template <class S>
inline SolverBodyInertia<S> make_inertia(S inv_mass, Sym3<S> inv_inertia, uint8_t locked, int8_t dominance, bool is_dynamic) {
    SolverBodyInertia<S> r;
    uint16_t flags = locked;
    if (inv_mass == S(0)) flags |= (1u << 6);
    if (is_zero(inv_inertia)) flags |= (1u << 7);
    if (locked & AVN_LOCK_ROTATION_X) { inv_inertia.m00 = 0; inv_inertia.m01 = 0; inv_inertia.m02 = 0; }
    if (locked & AVN_LOCK_ROTATION_Y) { inv_inertia.m01 = 0; inv_inertia.m11 = 0; inv_inertia.m12 = 0; }
    if (locked & AVN_LOCK_ROTATION_Z) { inv_inertia.m02 = 0; inv_inertia.m12 = 0; inv_inertia.m22 = 0; }
    r.inv_mass = inv_mass;
    r.inv_inertia = inv_inertia;
    r.dominance = is_dynamic ? int16_t(dominance) : int16_t(128);
    r.flags = flags;
    return r;
}

// integrator/mod.rs:216-233
template <class S>
struct VelocityIntegrationData {
    V3<S> linear_increment{0, 0, 0}, angular_increment{0, 0, 0};
    S linear_damping_rhs = 1, angular_damping_rhs = 1;
};

// softness_parameters/mod.rs:64-79
template <class S>
struct Softness {
    S bias, mass_scale, impulse_scale;
};
template <class S>
inline Softness<S> softness_coefficients(S damping_ratio, S frequency_hz, S delta_secs) {
    // SoftnessParameters::new(damping_ratio, frequency_hz): double_damping_ratio = 2*ratio, angular_frequency = TAU*hz
    S double_damping_ratio = S(2) * damping_ratio;
    S angular_frequency = S(6.283185307179586476925286766559) * frequency_hz;
    S a1 = double_damping_ratio + angular_frequency * delta_secs;
    S a2 = angular_frequency * delta_secs * a1;
    S a3 = S(1) / (S(1) + a2);
    return {angular_frequency / a1, a2 * a3, a3};
}

// contact/mod.rs:32-106, normal_part.rs:14-27, tangent_part.rs:14-27
template <class S>
struct ContactPointC {
    // normal part
    S impulse = 0, total_impulse = 0, effective_mass = 0;
    Softness<S> softness;
    // tangent part
    bool has_tangent = false;
    V2<S> t_impulse{0, 0};
    S k[3] = {0, 0, 0};
    V3<S> anchor1, anchor2;
    S normal_speed = 0, initial_separation = 0;
};
template <class S>
struct ContactConstraint {
    int32_t body1, body2;
    int16_t relative_dominance;
    S friction, restitution;
    V3<S> tangent_velocity, normal, tangent1;
    int npoints = 0;
    ContactPointC<S> points[AVN_MAX_MANIFOLD_POINTS];
    uint32_t first_point = 0;  // index into the P arrays
};

template <class S>
struct Columns {
    // typed accessors over the ABI's void* columns
    static V3<S> vec3(const void* p, size_t i) {
        const S* f = static_cast<const S*>(p) + 3 * i;
        return {f[0], f[1], f[2]};
    }
    static V3<S> vec3_or(const void* p, size_t i, V3<S> d) { return p ? vec3(p, i) : d; }
    static Quat<S> quat(const void* p, size_t i) {
        const S* f = static_cast<const S*>(p) + 4 * i;
        return {f[0], f[1], f[2], f[3]};
    }
    static Quat<S> quat_or_identity(const void* p, size_t i) { return p ? quat(p, i) : quat_identity<S>(); }
    static S scalar(const void* p, size_t i) { return static_cast<const S*>(p)[i]; }
    static S scalar_or(const void* p, size_t i, S d) { return p ? scalar(p, i) : d; }
    static Sym3<S> sym3(const void* p, size_t i) {
        const S* f = static_cast<const S*>(p) + 6 * i;
        return {f[0], f[1], f[2], f[3], f[4], f[5]};
    }
    static void set_vec3(void* p, size_t i, V3<S> v) {
        S* f = static_cast<S*>(p) + 3 * i;
        f[0] = v.x; f[1] = v.y; f[2] = v.z;
    }
    static void set_quat(void* p, size_t i, Quat<S> q) {
        S* f = static_cast<S*>(p) + 4 * i;
        f[0] = q.x; f[1] = q.y; f[2] = q.z; f[3] = q.w;
    }
};

// ---- XPBD joint solver data (xpbd/joints/*.rs) -----------------------------------------------------------
template <class S>
struct JointData {
    int type;
    int32_t body1, body2;
    // point constraint (shared/point_constraint.rs:16-21) / distance / prismatic
    V3<S> world_r1, world_r2, center_difference;
    V3<S> total_position_lagrange{0, 0, 0};
    // fixed-angle (shared/fixed_angle_constraint.rs:15-21)
    Quat<S> rotation_difference{0, 0, 0, 1};
    // revolute a1,a2,b1,b2 / spherical swing1,swing2,twist1,twist2 / prismatic free_axis1 in a1
    V3<S> a1, a2, b1, b2;
    V3<S> total_rot_lagrange_a{0, 0, 0}, total_rot_lagrange_b{0, 0, 0};
    // parameters
    uint8_t limit_enabled = 0;
    S limit_min = 0, limit_max = 0, limit2_min = 0, limit2_max = 0;
    S c0 = 0, c1 = 0, c2 = 0;
    bool damping = false;
    S damping_linear = 0, damping_angular = 0;
    uint32_t index_in_type = 0;
};

template <class S>
struct SolverWorld {
    std::vector<SolverBody<S>> bodies;          // one per column entry (static entries stay DUMMY)
    std::vector<SolverBodyInertia<S>> inertias;
    std::vector<VelocityIntegrationData<S>> integ;
    std::vector<ContactConstraint<S>> constraints[AVN_GRAPH_COLOR_COUNT];
    std::vector<JointData<S>> joints;  // reference solve order
    Softness<S> soft_dynamic, soft_non_dynamic;
};

// ---------------------------------------------------------------------------------------------------------
// contacts
// ---------------------------------------------------------------------------------------------------------

// contact/mod.rs:427-449
template <class S>
inline void compute_tangent_directions(V3<S> normal, V3<S> v1, V3<S> v2, V3<S>& t1, V3<S>& t2) {
    V3<S> force_direction = -normal;
    V3<S> relative_velocity = v1 - v2;
    V3<S> tangent_velocity = relative_velocity - force_direction * dot(force_direction, relative_velocity);
    V3<S> tangent;
    if (!try_normalize(tangent_velocity, tangent)) tangent = any_orthonormal_vector(force_direction);
    t1 = tangent;
    t2 = cross(force_direction, tangent);
}

// contact/mod.rs:110-220 ContactConstraint::generate (+ normal_part.rs:39-112, tangent_part.rs:35-151)
template <class S>
inline ContactConstraint<S> generate_constraint(int32_t b1, int32_t b2, const SolverBodyInertia<S>& inertia1,
                                                const SolverBodyInertia<S>& inertia2, V3<S> lin_vel1, V3<S> lin_vel2,
                                                const AvnManifoldColumns& m, uint32_t mi, bool warm_start,
                                                const Softness<S>& soft_dyn, const Softness<S>& soft_non_dyn) {
    using C = Columns<S>;
    ContactConstraint<S> c;
    c.body1 = b1;
    c.body2 = b2;
    int16_t rel = int16_t(inertia1.dominance - inertia2.dominance);
    c.relative_dominance = rel;
    V3<S> inv_mass1, inv_mass2;
    Sym3<S> i1, i2;
    if (rel == 0) {
        inv_mass1 = inertia1.effective_inv_mass(); i1 = inertia1.inv_inertia;
        inv_mass2 = inertia2.effective_inv_mass(); i2 = inertia2.inv_inertia;
    } else if (rel > 0) {
        inv_mass1 = {0, 0, 0}; i1 = sym3_zero<S>();
        inv_mass2 = inertia2.effective_inv_mass(); i2 = inertia2.inv_inertia;
    } else {
        inv_mass1 = inertia1.effective_inv_mass(); i1 = inertia1.inv_inertia;
        inv_mass2 = {0, 0, 0}; i2 = sym3_zero<S>();
    }
    Softness<S> softness = (rel != 0) ? soft_non_dyn : soft_dyn;
    V3<S> mass_sum = inv_mass1 + inv_mass2;
    V3<S> normal = C::vec3(m.normal, mi);
    V3<S> t1, t2;
    compute_tangent_directions(normal, lin_vel1, lin_vel2, t1, t2);
    c.friction = C::scalar(m.friction, mi);
    c.restitution = C::scalar(m.restitution, mi);
    c.tangent_velocity = C::vec3_or(m.tangent_velocity, mi, V3<S>{0, 0, 0});
    c.normal = normal;
    c.tangent1 = t1;
    uint32_t p0 = m.point_offsets[mi], p1 = m.point_offsets[mi + 1];
    c.first_point = p0;
    c.npoints = int(p1 - p0);
    for (uint32_t p = p0; p < p1; ++p) {
        ContactPointC<S>& pt = c.points[p - p0];
        V3<S> r1 = C::vec3(m.anchor1, p), r2 = C::vec3(m.anchor2, p);
        // normal part
        V3<S> r1xn = cross(r1, normal), r2xn = cross(r2, normal);
        S k_linear = dot(normal, mass_sum * normal);
        S k = k_linear + dot(r1xn, mul(i1, r1xn)) + dot(r2xn, mul(i2, r2xn));
        pt.impulse = warm_start ? C::scalar(m.warm_start_normal_impulse, p) : S(0);
        pt.total_impulse = 0;
        pt.effective_mass = recip_or_zero(k);
        pt.softness = softness;
        // tangent part
        pt.has_tangent = c.friction > S(0);
        if (pt.has_tangent) {
            if (warm_start) {
                const S* wt = static_cast<const S*>(m.warm_start_tangent_impulse) + 2 * p;
                pt.t_impulse = {wt[0], wt[1]};
            } else {
                pt.t_impulse = {0, 0};
            }
            V3<S> rt11 = cross(r1, t1), rt12 = cross(r2, t1), rt21 = cross(r1, t2), rt22 = cross(r2, t2);
            V3<S> i1_rt11 = mul(i1, rt11), i2_rt12 = mul(i2, rt12), i1_rt21 = mul(i1, rt21), i2_rt22 = mul(i2, rt22);
            S k_linear1 = dot(t1, mass_sum * t1);
            S k_linear2 = dot(t2, mass_sum * t2);
            pt.k[0] = k_linear1 + dot(rt11, i1_rt11) + dot(rt12, i2_rt12);
            pt.k[1] = k_linear2 + dot(rt21, i1_rt21) + dot(rt22, i2_rt22);
            pt.k[2] = S(2) * (dot(rt11, i1_rt21) + dot(rt12, i2_rt22));
        }
        pt.anchor1 = r1;
        pt.anchor2 = r2;
        pt.normal_speed = C::scalar(m.normal_speed, p);
        pt.initial_separation = -C::scalar(m.penetration, p) - dot(r2 - r1, normal);
    }
    return c;
}

template <class S>
struct BodyRef {
    SolverBody<S>* body;
    const SolverBodyInertia<S>* inertia;
};

// The get_unchecked + dominance dance of warm_start_internal / solve_contacts_internal (solver/plugin.rs:484-515,583-619)
template <class S>
inline void resolve_bodies(SolverWorld<S>& w, int32_t b1, int32_t b2, int16_t rel, SolverBody<S>& dummy1, SolverBody<S>& dummy2,
                           const SolverBodyInertia<S>& dummy_inertia, BodyRef<S>& r1, BodyRef<S>& r2,
                           const std::vector<uint8_t>& has_solver_body) {
    r1 = {&dummy1, &dummy_inertia};
    r2 = {&dummy2, &dummy_inertia};
    if (b1 >= 0 && has_solver_body[b1]) r1 = {&w.bodies[b1], &w.inertias[b1]};
    if (b2 >= 0 && has_solver_body[b2]) r2 = {&w.bodies[b2], &w.inertias[b2]};
    if (rel > 0) r1.inertia = &dummy_inertia;
    else if (rel < 0) r2.inertia = &dummy_inertia;
}

// contact/mod.rs:223-264
template <class S>
inline void warm_start_constraint(const ContactConstraint<S>& c, SolverBody<S>& body1, SolverBody<S>& body2,
                                  const SolverBodyInertia<S>& in1, const SolverBodyInertia<S>& in2, S coeff) {
    V3<S> inv_mass1 = in1.effective_inv_mass(), inv_mass2 = in2.effective_inv_mass();
    V3<S> t1 = c.tangent1, t2 = cross(c.tangent1, c.normal);
    for (int i = 0; i < c.npoints; ++i) {
        const ContactPointC<S>& pt = c.points[i];
        V3<S> r1 = pt.anchor1, r2 = pt.anchor2;
        V2<S> ti = pt.has_tangent ? pt.t_impulse : V2<S>{0, 0};
        V3<S> p = coeff * (pt.impulse * c.normal + ti.x * t1 + ti.y * t2);
        body1.linear_velocity -= p * inv_mass1;
        body1.angular_velocity -= mul(in1.inv_inertia, cross(r1, p));
        body2.linear_velocity += p * inv_mass2;
        body2.angular_velocity += mul(in2.inv_inertia, cross(r2, p));
    }
}

// contact/mod.rs:267-354 (+ normal_part.rs:116-166, tangent_part.rs:155-244)
template <class S>
inline void solve_constraint(ContactConstraint<S>& c, SolverBody<S>& body1, SolverBody<S>& body2, const SolverBodyInertia<S>& in1,
                             const SolverBodyInertia<S>& in2, S delta_secs, bool use_bias, S max_overlap_solve_speed) {
    V3<S> inv_mass1 = in1.effective_inv_mass(), inv_mass2 = in2.effective_inv_mass();
    const Sym3<S>& ii1 = in1.inv_inertia;
    const Sym3<S>& ii2 = in2.inv_inertia;
    V3<S> delta_translation = body2.delta_position - body1.delta_position;
    for (int i = 0; i < c.npoints; ++i) {
        ContactPointC<S>& pt = c.points[i];
        V3<S> r1 = rotate(body1.delta_rotation, pt.anchor1);
        V3<S> r2 = rotate(body2.delta_rotation, pt.anchor2);
        V3<S> delta_separation = delta_translation + (r2 - r1);
        S separation = dot(delta_separation, c.normal) + pt.initial_separation;
        r1 = pt.anchor1;
        r2 = pt.anchor2;
        V3<S> relative_velocity = body2.velocity_at_point(r2) - body1.velocity_at_point(r1);
        // ContactNormalPart::solve_impulse
        S normal_speed = dot(relative_velocity, c.normal);
        S impulse;
        if (separation > S(0)) {
            impulse = -pt.effective_mass * (normal_speed + separation / delta_secs);
        } else if (use_bias) {
            S bias = std::fmax(pt.softness.bias * separation, -max_overlap_solve_speed);
            S scaled_mass = pt.softness.mass_scale * pt.effective_mass;
            S scaled_impulse = pt.softness.impulse_scale * pt.impulse;
            impulse = -scaled_mass * (normal_speed + bias) - scaled_impulse;
        } else {
            impulse = -pt.effective_mass * normal_speed;
        }
        S new_impulse = std::fmax(pt.impulse + impulse, S(0));
        impulse = new_impulse - pt.impulse;
        pt.impulse = new_impulse;
        pt.total_impulse += new_impulse;
        V3<S> imp = impulse * c.normal;
        body1.linear_velocity -= imp * inv_mass1;
        body1.angular_velocity -= mul(ii1, cross(r1, imp));
        body2.linear_velocity += imp * inv_mass2;
        body2.angular_velocity += mul(ii2, cross(r2, imp));
    }
    V3<S> t1 = c.tangent1, t2 = cross(c.tangent1, c.normal);
    for (int i = 0; i < c.npoints; ++i) {
        ContactPointC<S>& pt = c.points[i];
        if (!pt.has_tangent) continue;
        V3<S> r1 = pt.anchor1, r2 = pt.anchor2;
        V3<S> relative_velocity = body2.velocity_at_point(r2) - body1.velocity_at_point(r1);
        // ContactTangentPart::solve_impulse
        S impulse_limit = c.friction * pt.impulse;
        relative_velocity = relative_velocity + c.tangent_velocity;
        S ts1 = dot(relative_velocity, t1), ts2 = dot(relative_velocity, t2);
        S t11 = ts1 * ts1, t22 = ts2 * ts2, t12 = ts1 * ts2;
        S inv = t11 * pt.k[0] + t22 * pt.k[1] + t12 * pt.k[2];
        S effective_mass = (t11 + t22) * (S(1) / inv);
        V3<S> imp{0, 0, 0};
        if (std::isfinite(effective_mass)) {
            V2<S> delta_impulse = effective_mass * V2<S>{ts1, ts2};
            V2<S> new_impulse = clamp_length_max(pt.t_impulse - delta_impulse, impulse_limit);
            V2<S> d = new_impulse - pt.t_impulse;
            pt.t_impulse = new_impulse;
            imp = d.x * t1 + d.y * t2;
        }
        body1.linear_velocity -= imp * inv_mass1;
        body1.angular_velocity -= mul(ii1, cross(r1, imp));
        body2.linear_velocity += imp * inv_mass2;
        body2.angular_velocity += mul(ii2, cross(r2, imp));
    }
}

// contact/mod.rs:358-407
template <class S>
inline void apply_restitution(ContactConstraint<S>& c, SolverBody<S>& body1, SolverBody<S>& body2, const SolverBodyInertia<S>& in1,
                              const SolverBodyInertia<S>& in2, S threshold) {
    V3<S> inv_mass1 = in1.effective_inv_mass(), inv_mass2 = in2.effective_inv_mass();
    for (int i = 0; i < c.npoints; ++i) {
        ContactPointC<S>& pt = c.points[i];
        if (pt.normal_speed > -threshold || pt.total_impulse == S(0)) continue;
        V3<S> r1 = pt.anchor1, r2 = pt.anchor2;
        V3<S> relative_velocity = body2.velocity_at_point(r2) - body1.velocity_at_point(r1);
        S normal_speed = dot(relative_velocity, c.normal);
        S impulse = -pt.effective_mass * (normal_speed + c.restitution * pt.normal_speed);
        S new_impulse = std::fmax(pt.impulse + impulse, S(0));
        impulse = new_impulse - pt.impulse;
        pt.impulse = new_impulse;
        pt.total_impulse += impulse;
        V3<S> imp = impulse * c.normal;
        body1.linear_velocity -= imp * inv_mass1;
        body1.angular_velocity -= mul(in1.inv_inertia, cross(r1, imp));
        body2.linear_velocity += imp * inv_mass2;
        body2.angular_velocity += mul(in2.inv_inertia, cross(r2, imp));
    }
}

// ---------------------------------------------------------------------------------------------------------
// integrator
// ---------------------------------------------------------------------------------------------------------

// integrator/mod.rs:403-460
template <class S>
inline void solve_gyroscopic_torque(V3<S>& ang_vel, Quat<S> rotation, const Sym3<S>& local_inverse_inertia, S delta_secs) {
    V3<S> local_ang_vel = rotate(inverse(rotation), ang_vel);
    Sym3<S> tensor = inverse_or_zero(local_inverse_inertia);
    V3<S> local_momentum = mul(tensor, local_ang_vel);
    V3<S> new_local_momentum = local_momentum - delta_secs * cross(local_ang_vel, local_momentum);
    S l2 = length_squared(new_local_momentum);
    if (l2 == S(0)) { ang_vel = {0, 0, 0}; return; }
    new_local_momentum *= std::sqrt(length_squared(local_momentum) / l2);
    ang_vel = rotate(rotation, mul(local_inverse_inertia, new_local_momentum));
}

// ---------------------------------------------------------------------------------------------------------
// XPBD joints
// ---------------------------------------------------------------------------------------------------------

// xpbd/mod.rs:393-413 compute_lagrange_update (lagrange is always 0 at every call site)
template <class S>
inline S compute_lagrange_update(S lagrange, S c, S w1, S w2, S compliance, S dt) {
    S w_sum = S(0) + w1 + w2;  // iter().copied().sum()
    if (w_sum <= std::numeric_limits<S>::epsilon()) return S(0);
    S tilde = compliance / (dt * dt);
    return (-c - tilde * lagrange) / (w_sum + tilde);
}

// positional_constraint.rs:9-50
template <class S>
inline void apply_positional_impulse(SolverBody<S>& b1, SolverBody<S>& b2, const SolverBodyInertia<S>& in1,
                                     const SolverBodyInertia<S>& in2, V3<S> impulse, V3<S> r1, V3<S> r2) {
    V3<S> inv_mass1 = in1.effective_inv_mass(), inv_mass2 = in2.effective_inv_mass();
    b1.delta_position += impulse * inv_mass1;
    b1.delta_rotation = mul(quat_from_scaled_axis(mul(in1.inv_inertia, cross(r1, impulse))), b1.delta_rotation);
    b2.delta_position -= impulse * inv_mass2;
    b2.delta_rotation = mul(quat_from_scaled_axis(mul(in2.inv_inertia, cross(r2, -impulse))), b2.delta_rotation);
}
// positional_constraint.rs:63-76
template <class S>
inline S positional_w(S inv_mass, const Sym3<S>& ii, V3<S> r, V3<S> n) {
    V3<S> rxn = cross(r, n);
    return inv_mass + dot(rxn, mul(ii, rxn));
}
// angular_constraint.rs:149-194 (3D align_orientation) + :52-97
template <class S>
inline V3<S> align_orientation(SolverBody<S>& b1, SolverBody<S>& b2, const Sym3<S>& ii1, const Sym3<S>& ii2, V3<S> rotation_difference,
                               S lagrange, S compliance, S dt) {
    S angle = length(rotation_difference);
    if (angle <= std::numeric_limits<S>::epsilon()) return {0, 0, 0};
    V3<S> axis = rotation_difference / angle;
    S w1 = dot(axis, mul(ii1, axis));
    S w2 = dot(axis, mul(ii2, axis));
    S delta_lagrange = compute_lagrange_update(lagrange, angle, w1, w2, compliance, dt);
    if (!(std::fabs(delta_lagrange) <= std::numeric_limits<S>::epsilon())) {
        V3<S> impulse = -delta_lagrange * axis;
        b1.delta_rotation = mul(quat_from_scaled_axis(mul(ii1, impulse)), b1.delta_rotation);
        b2.delta_rotation = mul(quat_from_scaled_axis(mul(ii2, -impulse)), b2.delta_rotation);
    }
    return delta_lagrange * axis;
}
// joints/mod.rs:427-473 AngleLimit::compute_correction (3D)
template <class S>
inline bool angle_limit_correction(S lim_min, S lim_max, V3<S> limit_axis, V3<S> axis1, V3<S> axis2, S max_correction, V3<S>& out) {
    const S PI = S(3.14159265358979323846264338327950288);
    const S TAU = S(6.28318530717958647692528676655900577);
    S phi = cr_asin(dot(cross(axis1, axis2), limit_axis));
    if (dot(axis1, axis2) < S(0)) phi = PI - phi;
    if (phi > PI) phi -= TAU;
    if (phi < lim_min || phi > lim_max) {
        // f32::clamp
        phi = phi < lim_min ? lim_min : (phi > lim_max ? lim_max : phi);
        Quat<S> rot = quat_from_axis_angle(limit_axis, phi);
        out = clamp_length_max(cross(rotate(rot, axis1), axis2), max_correction);
        return true;
    }
    return false;
}
// joints/mod.rs:321-357 DistanceLimit
template <class S>
inline void distance_limit_correction(S lim_min, S lim_max, V3<S> separation, V3<S>& dir, S& dist) {
    S d2 = length_squared(separation);
    if (d2 <= std::numeric_limits<S>::epsilon()) { dir = {0, 0, 0}; dist = 0; return; }
    S distance = std::sqrt(d2);
    if (distance < lim_min) { dir = separation / distance; dist = lim_min - distance; }
    else if (distance > lim_max) { dir = -separation / distance; dist = distance - lim_max; }
    else { dir = {0, 0, 0}; dist = 0; }
}
template <class S>
inline V3<S> distance_limit_along_axis(S lim_min, S lim_max, V3<S> separation, V3<S> axis) {
    S a = dot(separation, axis);
    if (a < lim_min) return axis * (lim_min - a);
    if (a > lim_max) return -axis * (a - lim_max);
    return {0, 0, 0};
}

// shared/point_constraint.rs:53-108
template <class S>
inline void point_constraint_solve(JointData<S>& j, SolverBody<S>& b1, SolverBody<S>& b2, const SolverBodyInertia<S>& in1,
                                   const SolverBodyInertia<S>& in2, S compliance, S dt) {
    V3<S> world_r1 = rotate(b1.delta_rotation, j.world_r1);
    V3<S> world_r2 = rotate(b2.delta_rotation, j.world_r2);
    V3<S> separation = (b2.delta_position - b1.delta_position) + (world_r2 - world_r1) + j.center_difference;
    S m2 = length_squared(separation);
    if (m2 == S(0)) return;
    S magnitude = std::sqrt(m2);
    V3<S> dir = -separation / magnitude;
    S w1 = positional_w(max_element(in1.effective_inv_mass()), in1.inv_inertia, world_r1, dir);
    S w2 = positional_w(max_element(in2.effective_inv_mass()), in2.inv_inertia, world_r2, dir);
    S dl = compute_lagrange_update(S(0), magnitude, w1, w2, compliance, dt);
    V3<S> impulse = dl * dir;
    j.total_position_lagrange += impulse;
    apply_positional_impulse(b1, b2, in1, in2, impulse, world_r1, world_r2);
}
// shared/fixed_angle_constraint.rs:59-97
template <class S>
inline void fixed_angle_solve(JointData<S>& j, SolverBody<S>& b1, SolverBody<S>& b2, const SolverBodyInertia<S>& in1,
                              const SolverBodyInertia<S>& in2, S compliance, S dt) {
    Quat<S> q = mul(mul(j.rotation_difference, b1.delta_rotation), inverse(b2.delta_rotation));
    V3<S> difference = S(-2) * xyz(q);
    j.total_rot_lagrange_a += align_orientation(b1, b2, in1.inv_inertia, in2.inv_inertia, difference, S(0), compliance, dt);
}

template <class S>
inline void solve_joint(JointData<S>& j, SolverBody<S>& b1, SolverBody<S>& b2, const SolverBodyInertia<S>& in1,
                        const SolverBodyInertia<S>& in2, S dt) {
    const S PI = S(3.14159265358979323846264338327950288);
    switch (j.type) {
        case AVN_JOINT_FIXED:  // xpbd/joints/fixed.rs:73-89
            fixed_angle_solve(j, b1, b2, in1, in2, j.c1, dt);
            point_constraint_solve(j, b1, b2, in1, in2, j.c0, dt);
            break;
        case AVN_JOINT_REVOLUTE: {  // xpbd/joints/revolute.rs:92-187
            V3<S> a1 = rotate(b1.delta_rotation, j.a1);
            V3<S> a2 = rotate(b2.delta_rotation, j.a2);
            V3<S> difference = cross(a1, a2);
            j.total_rot_lagrange_a += align_orientation(b1, b2, in1.inv_inertia, in2.inv_inertia, difference, S(0), j.c1, dt);
            if (j.limit_enabled & 1) {
                V3<S> la1 = rotate(b1.delta_rotation, j.a1);
                V3<S> lb1 = rotate(b1.delta_rotation, j.b1);
                V3<S> lb2 = rotate(b2.delta_rotation, j.b2);
                V3<S> corr;
                if (angle_limit_correction(j.limit_min, j.limit_max, la1, lb1, lb2, PI, corr))
                    j.total_rot_lagrange_b += align_orientation(b1, b2, in1.inv_inertia, in2.inv_inertia, corr, S(0), j.c2, dt);
            }
            point_constraint_solve(j, b1, b2, in1, in2, j.c0, dt);
            break;
        }
        case AVN_JOINT_SPHERICAL: {  // xpbd/joints/spherical.rs:84-207
            point_constraint_solve(j, b1, b2, in1, in2, j.c0, dt);
            if (j.limit_enabled & 1) {  // swing: a1/a2 = swing_axis1/2
                V3<S> a1 = rotate(b1.delta_rotation, j.a1);
                V3<S> a2 = rotate(b2.delta_rotation, j.a2);
                V3<S> n = cross(a1, a2);
                S nm = length(n);
                if (!(nm <= std::numeric_limits<S>::epsilon())) {
                    n = n / nm;
                    V3<S> corr;
                    if (angle_limit_correction(j.limit_min, j.limit_max, n, a1, a2, PI, corr))
                        j.total_rot_lagrange_a += align_orientation(b1, b2, in1.inv_inertia, in2.inv_inertia, corr, S(0), j.c1, dt);
                }
            }
            if (j.limit_enabled & 2) {  // twist: b1/b2 = twist_axis1/2
                V3<S> a1 = rotate(b1.delta_rotation, j.a1);
                V3<S> a2 = rotate(b2.delta_rotation, j.a2);
                V3<S> n = a1 + a2;
                S nm = length(n);
                if (nm <= std::numeric_limits<S>::epsilon()) break;
                V3<S> tb1 = rotate(b1.delta_rotation, j.b1);
                V3<S> tb2 = rotate(b2.delta_rotation, j.b2);
                n = n / nm;
                V3<S> n1 = tb1 - dot(n, tb1) * n;
                V3<S> n2 = tb2 - dot(n, tb2) * n;
                S n1m = length(n1), n2m = length(n2);
                if (n1m <= std::numeric_limits<S>::epsilon() || n2m <= std::numeric_limits<S>::epsilon()) break;
                n1 = n1 / n1m;
                n2 = n2 / n2m;
                S max_correction = dot(a1, a2) > S(-0.5) ? S(2) * PI : dt;
                V3<S> corr;
                if (angle_limit_correction(j.limit2_min, j.limit2_max, n, n1, n2, max_correction, corr))
                    j.total_rot_lagrange_b += align_orientation(b1, b2, in1.inv_inertia, in2.inv_inertia, corr, S(0), j.c2, dt);
            }
            break;
        }
        case AVN_JOINT_PRISMATIC: {  // xpbd/joints/prismatic.rs:79-193
            fixed_angle_solve(j, b1, b2, in1, in2, j.c1, dt);
            V3<S> world_r1 = rotate(b1.delta_rotation, j.world_r1);
            V3<S> world_r2 = rotate(b2.delta_rotation, j.world_r2);
            V3<S> delta_x{0, 0, 0};
            V3<S> axis1 = rotate(b1.delta_rotation, j.a1);
            V3<S> separation = (b2.delta_position - b1.delta_position) + (world_r2 - world_r1) + j.center_difference;
            if (j.limit_enabled & 1) delta_x += distance_limit_along_axis(j.limit_min, j.limit_max, separation, axis1);
            V3<S> axis2 = any_orthogonal_vector(axis1);
            V3<S> axis3 = cross(axis1, axis2);
            delta_x += distance_limit_along_axis(S(0), S(0), separation, axis2);
            delta_x += distance_limit_along_axis(S(0), S(0), separation, axis3);
            S magnitude = length(delta_x);
            if (magnitude <= std::numeric_limits<S>::epsilon()) break;
            V3<S> dir = delta_x / magnitude;
            S w1 = positional_w(max_element(in1.effective_inv_mass()), in1.inv_inertia, world_r1, dir);
            S w2 = positional_w(max_element(in2.effective_inv_mass()), in2.inv_inertia, world_r2, dir);
            S dl = compute_lagrange_update(S(0), magnitude, w1, w2, j.c0, dt);
            V3<S> impulse = dl * dir;
            j.total_position_lagrange += impulse;
            apply_positional_impulse(b1, b2, in1, in2, impulse, world_r1, world_r2);
            break;
        }
        case AVN_JOINT_DISTANCE: {  // xpbd/joints/distance.rs:56-117
            V3<S> world_r1 = rotate(b1.delta_rotation, j.world_r1);
            V3<S> world_r2 = rotate(b2.delta_rotation, j.world_r2);
            V3<S> separation = (b2.delta_position - b1.delta_position) + (world_r2 - world_r1) + j.center_difference;
            V3<S> dir;
            S distance;
            distance_limit_correction(j.limit_min, j.limit_max, separation, dir, distance);
            if (distance <= std::numeric_limits<S>::epsilon()) break;
            S w1 = positional_w(max_element(in1.effective_inv_mass()), in1.inv_inertia, world_r1, dir);
            S w2 = positional_w(max_element(in2.effective_inv_mass()), in2.inv_inertia, world_r2, dir);
            S dl = compute_lagrange_update(S(0), distance, w1, w2, j.c0, dt);
            V3<S> impulse = dl * dir;
            j.total_position_lagrange += impulse;
            apply_positional_impulse(b1, b2, in1, in2, impulse, world_r1, world_r2);
            break;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// the whole solver stage
// ---------------------------------------------------------------------------------------------------------
template <class S>
int solver_step(const AvnStepParams& prm, AvnBodyColumns& bc, AvnManifoldColumns* mc, AvnJointSet* js, int threads);

}  // namespace orc
