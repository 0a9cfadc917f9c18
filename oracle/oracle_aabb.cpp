// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
// CPU restatement (oracle) of update_aabb for cuboid / sphere colliders (src/collision/collider/backend.rs:498-625).
// PARITY UNPINNED: the shape AABB arithmetic lives in parry3d 0.25.0 / nalgebra 0.34.1 (Cargo.lock:3564,3030), which are not vendored:
// Cuboid::aabb = center +- |R| * half_extents with R = UnitQuaternion::to_rotation_matrix (nalgebra's ww+ii-jj-kk form), Ball::aabb =
// center +- radius, Aabb::merged / ColliderAabb::grow componentwise.  Restated from the published algorithms.
#include <cmath>
#include <limits>

#include "../include/avian_b200.h"
#include "oracle_math.hpp"

namespace {
using namespace orc;

template <class S>
void shape_aabb(int shape, V3<S> dims, V3<S> pos, Quat<S> q, V3<S>& mn, V3<S>& mx) {
    V3<S> he;
    if (shape == AVN_SHAPE_SPHERE) {
        he = {dims.x, dims.x, dims.x};
    } else {
        // nalgebra UnitQuaternion::to_rotation_matrix
        S i = q.x, j = q.y, k = q.z, w = q.w;
        S ww = w * w, ii = i * i, jj = j * j, kk = k * k;
        S ij = i * j * S(2), wk = w * k * S(2), wj = w * j * S(2), ik = i * k * S(2), jk = j * k * S(2), wi = w * i * S(2);
        S m11 = ww + ii - jj - kk, m12 = ij - wk, m13 = wj + ik;
        S m21 = wk + ij, m22 = ww - ii + jj - kk, m23 = jk - wi;
        S m31 = ik - wj, m32 = wi + jk, m33 = ww - ii - jj + kk;
        // |M| * half_extents (matrix-vector product, row sums left to right)
        he = {std::fabs(m11) * dims.x + std::fabs(m12) * dims.y + std::fabs(m13) * dims.z,
              std::fabs(m21) * dims.x + std::fabs(m22) * dims.y + std::fabs(m23) * dims.z,
              std::fabs(m31) * dims.x + std::fabs(m32) * dims.y + std::fabs(m33) * dims.z};
    }
    mn = pos - he;
    mx = pos + he;
}

template <class S>
int update_aabbs(const AvnAabbParams& prm, AvnColliderColumns& c) {
    const S* dims = static_cast<const S*>(c.dims);
    const S* pos = static_cast<const S*>(c.position);
    const S* rot = static_cast<const S*>(c.rotation);
    const S* lv = static_cast<const S*>(c.linear_velocity);
    const S* av = static_cast<const S*>(c.angular_velocity);
    const S* cm = static_cast<const S*>(c.collision_margin);
    const S* sm = static_cast<const S*>(c.speculative_margin);
    S* omn = static_cast<S*>(c.aabb_min);
    S* omx = static_cast<S*>(c.aabb_max);
    const S dt = S(prm.dt), tol = S(prm.contact_tolerance);
    const S def_spec = std::isinf(prm.default_speculative_margin) ? std::numeric_limits<S>::max() : S(prm.default_speculative_margin);
    for (uint32_t n = 0; n < c.count; ++n) {
        V3<S> d{dims[3 * n], dims[3 * n + 1], dims[3 * n + 2]}, p{pos[3 * n], pos[3 * n + 1], pos[3 * n + 2]};
        Quat<S> q{rot[4 * n], rot[4 * n + 1], rot[4 * n + 2], rot[4 * n + 3]};
        S margin = cm ? cm[n] : S(0);
        S spec = sm ? (std::isinf(sm[n]) ? std::numeric_limits<S>::max() : sm[n]) : def_spec;
        V3<S> mn, mx;
        if (spec <= S(0)) {
            shape_aabb<S>(c.shape[n], d, p, q, mn, mx);
        } else {
            V3<S> v = lv ? V3<S>{lv[3 * n], lv[3 * n + 1], lv[3 * n + 2]} : V3<S>{0, 0, 0};
            V3<S> w = av ? V3<S>{av[3 * n], av[3 * n + 1], av[3 * n + 2]} : V3<S>{0, 0, 0};
            Quat<S> end_rot = fast_renormalize(mul(quat_from_scaled_axis(w * dt), q));
            V3<S> end_pos = p + clamp_length_max(v * dt, std::fmax(spec, tol));
            V3<S> mn0, mx0, mn1, mx1;
            shape_aabb<S>(c.shape[n], d, p, q, mn0, mx0);
            shape_aabb<S>(c.shape[n], d, end_pos, end_rot, mn1, mx1);
            mn = {std::fmin(mn0.x, mn1.x), std::fmin(mn0.y, mn1.y), std::fmin(mn0.z, mn1.z)};
            mx = {std::fmax(mx0.x, mx1.x), std::fmax(mx0.y, mx1.y), std::fmax(mx0.z, mx1.z)};
        }
        S g = tol + margin;
        omn[3 * n] = mn.x - g; omn[3 * n + 1] = mn.y - g; omn[3 * n + 2] = mn.z - g;
        omx[3 * n] = mx.x + g; omx[3 * n + 1] = mx.y + g; omx[3 * n + 2] = mx.z + g;
    }
    return AVN_OK;
}
}  // namespace

extern "C" int orc_update_aabbs(uint32_t scalar_bits, const AvnAabbParams* prm, AvnColliderColumns* c) {
    if (!prm || !c) return AVN_ERR_INVALID_ARGUMENT;
    if (scalar_bits == 32) return update_aabbs<float>(*prm, *c);
    if (scalar_bits == 64) return update_aabbs<double>(*prm, *c);
    return AVN_ERR_INVALID_ARGUMENT;
}
