// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
// CPU restatement (oracle) of the arithmetic primitives the avian3d substep hot path calls.
//
// PARITY UNPINNED: these primitives live in third-party crates that are NOT vendored under /root/reference
// (glam 0.30.8, glam_matrix_extras 0.1.0, bevy_math 0.17.2 — Cargo.lock:2472,2485,917) and no Rust toolchain
// exists in this environment, so the restatement follows the crates' published algorithms and the reference's
// call sites (integrator/mod.rs:426,459,529-530; contact/mod.rs:443-446; xpbd/positional_constraint.rs:92;
// xpbd/angular_constraint.rs:279; joints/mod.rs:437,466; mass_properties/components/computed.rs:663-668).
// Operation order follows glam's x86-64 build: `Quat` is SSE2 (lane-wise IEEE ops, association as in
// glam/src/sse2/quat.rs), `Vec3`/`Mat3`/all f64 types are scalar code.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>

namespace orc {

template <class S>
struct V3 {
    S x, y, z;
};
template <class S>
struct V2 {
    S x, y;
};
template <class S>
struct Quat {
    S x, y, z, w;
};
// glam_matrix_extras::SymmetricMat3 field order (visible at solver_body/mod.rs:396-410).
template <class S>
struct Sym3 {
    S m00, m01, m02, m11, m12, m22;
};
// column-major like glam::Mat3
template <class S>
struct Mat3 {
    V3<S> c0, c1, c2;
};

template <class S> inline V3<S> v3(S x, S y, S z) { return V3<S>{x, y, z}; }
template <class S> inline V3<S> operator+(V3<S> a, V3<S> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class S> inline V3<S> operator-(V3<S> a, V3<S> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class S> inline V3<S> operator-(V3<S> a) { return {-a.x, -a.y, -a.z}; }
template <class S> inline V3<S> operator*(V3<S> a, S s) { return {a.x * s, a.y * s, a.z * s}; }
template <class S> inline V3<S> operator*(S s, V3<S> a) { return {s * a.x, s * a.y, s * a.z}; }
template <class S> inline V3<S> operator*(V3<S> a, V3<S> b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
template <class S> inline V3<S> operator/(V3<S> a, S s) { return {a.x / s, a.y / s, a.z / s}; }
template <class S> inline V3<S>& operator+=(V3<S>& a, V3<S> b) { a = a + b; return a; }
template <class S> inline V3<S>& operator-=(V3<S>& a, V3<S> b) { a = a - b; return a; }
template <class S> inline V3<S>& operator*=(V3<S>& a, S s) { a = a * s; return a; }
// glam Vec3::dot: (x*x + y*y) + z*z
template <class S> inline S dot(V3<S> a, V3<S> b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }
// glam Vec3::cross
template <class S> inline V3<S> cross(V3<S> a, V3<S> b) {
    return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
template <class S> inline S length_squared(V3<S> a) { return dot(a, a); }
template <class S> inline S length(V3<S> a) { return std::sqrt(dot(a, a)); }
template <class S> inline bool is_finite(S s) { return std::isfinite(s); }
template <class S> inline S max_element(V3<S> a) { return std::fmax(a.x, std::fmax(a.y, a.z)); }

template <class S> inline V2<S> operator+(V2<S> a, V2<S> b) { return {a.x + b.x, a.y + b.y}; }
template <class S> inline V2<S> operator-(V2<S> a, V2<S> b) { return {a.x - b.x, a.y - b.y}; }
template <class S> inline V2<S> operator*(S s, V2<S> a) { return {s * a.x, s * a.y}; }
template <class S> inline V2<S> operator/(V2<S> a, S s) { return {a.x / s, a.y / s}; }

// src/math/mod.rs:244-262 RecipOrZero
template <class S> inline S recip_or_zero(S s) { return (s != S(0) && std::isfinite(s)) ? S(1) / s : S(0); }
template <class S> inline V3<S> recip_or_zero(V3<S> a) { return {recip_or_zero(a.x), recip_or_zero(a.y), recip_or_zero(a.z)}; }

// glam Vec3::try_normalize: rcp = 1/length; Some(self * rcp) iff rcp finite and > 0
template <class S> inline bool try_normalize(V3<S> a, V3<S>& out) {
    S rcp = S(1) / length(a);
    if (std::isfinite(rcp) && rcp > S(0)) { out = a * rcp; return true; }
    return false;
}
// glam Vec3::any_orthonormal_vector (Pixar "Building an Orthonormal Basis, Revisited")
template <class S> inline V3<S> any_orthonormal_vector(V3<S> a) {
    S sign = std::signbit(a.z) ? S(-1) : S(1);  // f32::signum: +0 -> 1, -0 -> -1
    S aa = S(-1) / (sign + a.z);
    S b = a.x * a.y * aa;
    return {b, sign + a.y * a.y * aa, -a.y};
}
// glam Vec3::any_orthogonal_vector
template <class S> inline V3<S> any_orthogonal_vector(V3<S> a) {
    if (std::fabs(a.x) > std::fabs(a.y)) return {-a.z, S(0), a.x};
    return {S(0), a.z, -a.y};
}
// glam Vec2::clamp_length_max
template <class S> inline V2<S> clamp_length_max(V2<S> a, S max) {
    S l2 = a.x * a.x + a.y * a.y;
    if (l2 > max * max) return max * (a / std::sqrt(l2));
    return a;
}
template <class S> inline V3<S> clamp_length_max(V3<S> a, S max) {
    S l2 = length_squared(a);
    if (l2 > max * max) return max * (a / std::sqrt(l2));
    return a;
}

template <class S> inline Quat<S> quat_identity() { return {S(0), S(0), S(0), S(1)}; }
template <class S> inline V3<S> xyz(Quat<S> q) { return {q.x, q.y, q.z}; }
// glam Quat::inverse == conjugate (unit quaternions)
template <class S> inline Quat<S> inverse(Quat<S> q) { return {-q.x, -q.y, -q.z, q.w}; }

// glam Quat::mul_quat. f32: sse2/quat.rs association ((w*r + x*r') + (y*r'' + z*r''')); f64: scalar left-assoc.
inline Quat<float> mul(Quat<float> a, Quat<float> b) {
    Quat<float> r;
    r.x = (a.w * b.x + (a.x * b.w) * 1.0f) + ((a.y * b.z) * 1.0f + (a.z * b.y) * -1.0f);
    r.y = (a.w * b.y + (a.x * b.z) * -1.0f) + ((a.y * b.w) * 1.0f + (a.z * b.x) * 1.0f);
    r.z = (a.w * b.z + (a.x * b.y) * 1.0f) + ((a.y * b.x) * -1.0f + (a.z * b.w) * 1.0f);
    r.w = (a.w * b.w + (a.x * b.x) * -1.0f) + ((a.y * b.y) * -1.0f + (a.z * b.z) * -1.0f);
    return r;
}
inline Quat<double> mul(Quat<double> a, Quat<double> b) {
    return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
            a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
// glam Quat * Vec3: v*(w*w - b.b) + b*((v.b)*2) + (b x v)*(w*2)
template <class S> inline V3<S> rotate(Quat<S> q, V3<S> v) {
    S w = q.w;
    V3<S> b = {q.x, q.y, q.z};
    S b2 = dot(b, b);
    return (v * (w * w - b2)) + (b * (dot(v, b) * S(2))) + (cross(b, v) * (w * S(2)));
}
// f32 transcendentals: CORRECTLY ROUNDED (evaluated in double, rounded once) — the definition the CUDA path uses (csrc/avn_math.cuh) and the
// one glibc >= 2.41 implements (CORE-MATH sinf/cosf/asinf).  The reference calls the platform libm through Rust's f32::sin/cos/asin, so its
// last bit is platform-dependent: glibc 2.39's asinf differs from the correctly rounded value in 7 % of arguments, sinf/cosf in 1.3 %
// (measured on this image).  Such one-ulp differences in joint angles are amplified by the XPBD velocity projection (x 2/h) to ~1e-4 rad/s,
// so "1e-5 of Avian" is only meaningful for jointed scenes once the libm is fixed; the oracle fixes it to the correctly rounded one.
inline float cr_sin(float x) { return float(std::sin(double(x))); }
inline float cr_cos(float x) { return float(std::cos(double(x))); }
inline float cr_asin(float x) { return float(std::asin(double(x))); }
inline double cr_sin(double x) { return std::sin(x); }
inline double cr_cos(double x) { return std::cos(x); }
inline double cr_asin(double x) { return std::asin(x); }
// glam Quat::from_axis_angle / from_scaled_axis
template <class S> inline Quat<S> quat_from_axis_angle(V3<S> axis, S angle) {
    S s = cr_sin(angle * S(0.5)), c = cr_cos(angle * S(0.5));
    V3<S> v = axis * s;
    return {v.x, v.y, v.z, c};
}
template <class S> inline Quat<S> quat_from_scaled_axis(V3<S> v) {
    S len = length(v);
    if (len == S(0)) return quat_identity<S>();
    return quat_from_axis_angle(v / len, len);
}
template <class S> inline S length_squared(Quat<S> q) { return ((q.x * q.x + q.y * q.y) + q.z * q.z) + q.w * q.w; }
// src/physics_transform/transform.rs:811-817 Rotation::fast_renormalize
template <class S> inline Quat<S> fast_renormalize(Quat<S> q) {
    S l2 = length_squared(q);
    S k = S(0.5) * (S(3) - l2);
    return {q.x * k, q.y * k, q.z * k, q.w * k};
}

// glam Mat3::from_quat
template <class S> inline Mat3<S> mat3_from_quat(Quat<S> r) {
    S x2 = r.x + r.x, y2 = r.y + r.y, z2 = r.z + r.z;
    S xx = r.x * x2, xy = r.x * y2, xz = r.x * z2;
    S yy = r.y * y2, yz = r.y * z2, zz = r.z * z2;
    S wx = r.w * x2, wy = r.w * y2, wz = r.w * z2;
    return {{S(1) - (yy + zz), xy + wz, xz - wy}, {xy - wz, S(1) - (xx + zz), yz + wx}, {xz + wy, yz - wx, S(1) - (xx + yy)}};
}
// glam Mat3::mul_vec3: (c0*x + c1*y) + c2*z
template <class S> inline V3<S> mul(const Mat3<S>& m, V3<S> v) { return (m.c0 * v.x + m.c1 * v.y) + m.c2 * v.z; }
template <class S> inline Mat3<S> mul(const Mat3<S>& a, const Mat3<S>& b) { return {mul(a, b.c0), mul(a, b.c1), mul(a, b.c2)}; }
template <class S> inline Mat3<S> transpose(const Mat3<S>& m) {
    return {{m.c0.x, m.c1.x, m.c2.x}, {m.c0.y, m.c1.y, m.c2.y}, {m.c0.z, m.c1.z, m.c2.z}};
}
template <class S> inline Mat3<S> to_mat3(const Sym3<S>& s) {
    return {{s.m00, s.m01, s.m02}, {s.m01, s.m11, s.m12}, {s.m02, s.m12, s.m22}};
}
template <class S> inline Sym3<S> sym3_from_mat3_unchecked(const Mat3<S>& m) {
    return {m.c0.x, m.c0.y, m.c0.z, m.c1.y, m.c1.z, m.c2.z};
}
template <class S> inline Sym3<S> sym3_zero() { return {S(0), S(0), S(0), S(0), S(0), S(0)}; }
template <class S> inline bool is_zero(const Sym3<S>& s) {
    return s.m00 == 0 && s.m01 == 0 && s.m02 == 0 && s.m11 == 0 && s.m12 == 0 && s.m22 == 0;
}
// SymmetricMat3 * Vec3: (col0*x + col1*y) + col2*z
template <class S> inline V3<S> mul(const Sym3<S>& s, V3<S> v) {
    return {(s.m00 * v.x + s.m01 * v.y) + s.m02 * v.z, (s.m01 * v.x + s.m11 * v.y) + s.m12 * v.z,
            (s.m02 * v.x + s.m12 * v.y) + s.m22 * v.z};
}
// SymmetricMat3::inverse_or_zero (src/math/mod.rs:515-524) — cofactor inverse.
template <class S> inline Sym3<S> inverse_or_zero(const Sym3<S>& s) {
    S c00 = s.m11 * s.m22 - s.m12 * s.m12;
    S c01 = s.m02 * s.m12 - s.m01 * s.m22;
    S c02 = s.m01 * s.m12 - s.m02 * s.m11;
    S det = (s.m00 * c00 + s.m01 * c01) + s.m02 * c02;
    if (det == S(0)) return sym3_zero<S>();
    S inv = S(1) / det;
    S c11 = s.m00 * s.m22 - s.m02 * s.m02;
    S c12 = s.m01 * s.m02 - s.m00 * s.m12;
    S c22 = s.m00 * s.m11 - s.m01 * s.m01;
    return {c00 * inv, c01 * inv, c02 * inv, c11 * inv, c12 * inv, c22 * inv};
}
// src/math/mod.rs:526-543
template <class S> inline bool is_isotropic(const Sym3<S>& s, S eps) {
    if (std::fabs(s.m00 - s.m11) > eps || std::fabs(s.m11 - s.m22) > eps) return false;
    return std::fabs(s.m01) < eps && std::fabs(s.m02) < eps && std::fabs(s.m12) < eps;
}
// ComputedAngularInertia::rotated(q).inverse(): (R * I^-1) * R^T  (mass_properties/components/computed.rs:663-668)
template <class S> inline Sym3<S> rotate_inverse_inertia(const Sym3<S>& inv_local, Quat<S> q) {
    Mat3<S> r = mat3_from_quat(q);
    Mat3<S> ri = mul(r, to_mat3(inv_local));
    return sym3_from_mat3_unchecked(mul(ri, transpose(r)));
}

}  // namespace orc
