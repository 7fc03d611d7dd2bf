// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into or called from the product path.
// PARITY UNPINNED: Box2D (box2d_catkin, 2.3.x, version un-pinned in /root/reference/package.xml:47,51)
// is not available in this container and the reference's own tests hold no numeric vectors
// (tests/test_Box2DWorld.cpp:101-117).  This file restates the published Box2D 2.3.1 math
// conventions (Common/b2Math.h, Common/b2Settings.h) that the reference reaches through
// `#include "Box2D/Box2D.h"` (include/sim_env/Box2DWorld.h:7).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

namespace b2o {

// ---- b2Settings.h (2.3.1) -------------------------------------------------------------------
constexpr float kMaxFloat = FLT_MAX;
constexpr float kEpsilon = FLT_EPSILON;
constexpr float kPi = 3.14159265359f;
constexpr int kMaxManifoldPoints = 2;
constexpr int kMaxPolygonVertices = 8;
constexpr float kAabbExtension = 0.1f;
constexpr float kAabbMultiplier = 2.0f;
constexpr float kLinearSlop = 0.005f;
constexpr float kAngularSlop = (2.0f / 180.0f * kPi);
constexpr float kPolygonRadius = (2.0f * kLinearSlop);
constexpr int kMaxSubSteps = 8;
constexpr int kMaxTOIContacts = 32;
constexpr float kVelocityThreshold = 1.0f;
constexpr float kMaxLinearCorrection = 0.2f;
constexpr float kMaxAngularCorrection = (8.0f / 180.0f * kPi);
constexpr float kMaxTranslation = 2.0f;
constexpr float kMaxTranslationSquared = (kMaxTranslation * kMaxTranslation);
constexpr float kMaxRotation = (0.5f * kPi);
constexpr float kMaxRotationSquared = (kMaxRotation * kMaxRotation);
constexpr float kBaumgarte = 0.2f;
constexpr float kToiBaumgarte = 0.75f;
constexpr float kTimeToSleep = 0.5f;
constexpr float kLinearSleepTolerance = 0.01f;
constexpr float kAngularSleepTolerance = (2.0f / 180.0f * kPi);

// ---- deterministic sin/cos --------------------------------------------------------------------
// Box2D calls sinf/cosf (b2Rot::Set).  libm and CUDA differ in the last ulp, so the oracle and the
// CUDA path both use this fixed sequence of IEEE single operations (Cody-Waite reduction by pi/2
// followed by the Cephes single-precision minimax cores).  Build with -ffp-contract=off.
// Defining B2O_LIBM_TRIG switches the oracle to libm for sensitivity studies.
inline void SinCos(float x, float& s, float& c) {
#ifdef B2O_LIBM_TRIG
    s = sinf(x);
    c = cosf(x);
#else
    float kf = rintf(x * 0.636619772367581343f);
    float r = x - kf * 1.5703125f;
    r = r - kf * 4.837512969970703125e-4f;
    r = r - kf * 7.54978995489188216e-8f;
    float z = r * r;
    float ps = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * r + r;
    float pc = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z;
    pc = pc - 0.5f * z;
    pc = pc + 1.0f;
    int q = ((int)kf) & 3;
    if (q == 0) { s = ps; c = pc; }
    else if (q == 1) { s = pc; c = -ps; }
    else if (q == 2) { s = -ps; c = -pc; }
    else { s = -pc; c = ps; }
#endif
}

template <typename T> inline T Min(T a, T b) { return a < b ? a : b; }
template <typename T> inline T Max(T a, T b) { return a > b ? a : b; }
template <typename T> inline T Clamp(T a, T low, T high) { return Max(low, Min(a, high)); }
inline float Abs(float a) { return a > 0.0f ? a : -a; }

struct Vec2 {
    float x, y;
    Vec2() : x(0.0f), y(0.0f) {}
    Vec2(float x_, float y_) : x(x_), y(y_) {}
    void SetZero() { x = 0.0f; y = 0.0f; }
    void Set(float x_, float y_) { x = x_; y = y_; }
    Vec2 operator-() const { return Vec2(-x, -y); }
    void operator+=(const Vec2& v) { x += v.x; y += v.y; }
    void operator-=(const Vec2& v) { x -= v.x; y -= v.y; }
    void operator*=(float a) { x *= a; y *= a; }
    float Length() const { return sqrtf(x * x + y * y); }
    float LengthSquared() const { return x * x + y * y; }
    float Normalize() {
        float length = Length();
        if (length < kEpsilon) return 0.0f;
        float invLength = 1.0f / length;
        x *= invLength;
        y *= invLength;
        return length;
    }
};

struct Vec3 {
    float x, y, z;
    Vec3() : x(0.0f), y(0.0f), z(0.0f) {}
    Vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    void SetZero() { x = y = z = 0.0f; }
    Vec3 operator-() const { return Vec3(-x, -y, -z); }
    void operator+=(const Vec3& v) { x += v.x; y += v.y; z += v.z; }
    void operator*=(float s) { x *= s; y *= s; z *= s; }
};

inline Vec2 operator+(const Vec2& a, const Vec2& b) { return Vec2(a.x + b.x, a.y + b.y); }
inline Vec2 operator-(const Vec2& a, const Vec2& b) { return Vec2(a.x - b.x, a.y - b.y); }
inline Vec2 operator*(float s, const Vec2& a) { return Vec2(s * a.x, s * a.y); }
inline float Dot(const Vec2& a, const Vec2& b) { return a.x * b.x + a.y * b.y; }
inline float Cross(const Vec2& a, const Vec2& b) { return a.x * b.y - a.y * b.x; }
inline Vec2 Cross(const Vec2& a, float s) { return Vec2(s * a.y, -s * a.x); }
inline Vec2 Cross(float s, const Vec2& a) { return Vec2(-s * a.y, s * a.x); }
inline float Distance(const Vec2& a, const Vec2& b) { Vec2 c = a - b; return c.Length(); }
inline float DistanceSquared(const Vec2& a, const Vec2& b) { Vec2 c = a - b; return Dot(c, c); }
inline Vec2 Min(const Vec2& a, const Vec2& b) { return Vec2(Min(a.x, b.x), Min(a.y, b.y)); }
inline Vec2 Max(const Vec2& a, const Vec2& b) { return Vec2(Max(a.x, b.x), Max(a.y, b.y)); }

inline float Dot(const Vec3& a, const Vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 Cross(const Vec3& a, const Vec3& b) {
    return Vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

struct Mat22 {
    Vec2 ex, ey;
    void SetZero() { ex.SetZero(); ey.SetZero(); }
    Mat22 GetInverse() const {
        float a = ex.x, b = ey.x, c = ex.y, d = ey.y;
        Mat22 B;
        float det = a * d - b * c;
        if (det != 0.0f) det = 1.0f / det;
        B.ex.x = det * d;  B.ey.x = -det * b;
        B.ex.y = -det * c; B.ey.y = det * a;
        return B;
    }
    Vec2 Solve(const Vec2& b) const {
        float a11 = ex.x, a12 = ey.x, a21 = ex.y, a22 = ey.y;
        float det = a11 * a22 - a12 * a21;
        if (det != 0.0f) det = 1.0f / det;
        Vec2 x;
        x.x = det * (a22 * b.x - a12 * b.y);
        x.y = det * (a11 * b.y - a21 * b.x);
        return x;
    }
};
inline Vec2 Mul(const Mat22& A, const Vec2& v) {
    return Vec2(A.ex.x * v.x + A.ey.x * v.y, A.ex.y * v.x + A.ey.y * v.y);
}

struct Mat33 {
    Vec3 ex, ey, ez;
    Vec3 Solve33(const Vec3& b) const {
        float det = Dot(ex, Cross(ey, ez));
        if (det != 0.0f) det = 1.0f / det;
        Vec3 x;
        x.x = det * Dot(b, Cross(ey, ez));
        x.y = det * Dot(ex, Cross(b, ez));
        x.z = det * Dot(ex, Cross(ey, b));
        return x;
    }
    Vec2 Solve22(const Vec2& b) const {
        float a11 = ex.x, a12 = ey.x, a21 = ex.y, a22 = ey.y;
        float det = a11 * a22 - a12 * a21;
        if (det != 0.0f) det = 1.0f / det;
        Vec2 x;
        x.x = det * (a22 * b.x - a12 * b.y);
        x.y = det * (a11 * b.y - a21 * b.x);
        return x;
    }
};

struct Rot {
    float s, c;
    Rot() : s(0.0f), c(1.0f) {}
    explicit Rot(float angle) { SinCos(angle, s, c); }
    void Set(float angle) { SinCos(angle, s, c); }
    void SetIdentity() { s = 0.0f; c = 1.0f; }
};

struct Transform {
    Vec2 p;
    Rot q;
    void SetIdentity() { p.SetZero(); q.SetIdentity(); }
    void Set(const Vec2& position, float angle) { p = position; q.Set(angle); }
};

inline Rot MulT(const Rot& q, const Rot& r) {
    Rot qr;
    qr.s = q.c * r.s - q.s * r.c;
    qr.c = q.c * r.c + q.s * r.s;
    return qr;
}
inline Vec2 Mul(const Rot& q, const Vec2& v) { return Vec2(q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y); }
inline Vec2 MulT(const Rot& q, const Vec2& v) { return Vec2(q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y); }
inline Vec2 Mul(const Transform& T, const Vec2& v) {
    float x = (T.q.c * v.x - T.q.s * v.y) + T.p.x;
    float y = (T.q.s * v.x + T.q.c * v.y) + T.p.y;
    return Vec2(x, y);
}
inline Vec2 MulT(const Transform& T, const Vec2& v) {
    float px = v.x - T.p.x;
    float py = v.y - T.p.y;
    float x = (T.q.c * px + T.q.s * py);
    float y = (-T.q.s * px + T.q.c * py);
    return Vec2(x, y);
}
inline Transform MulT(const Transform& A, const Transform& B) {
    Transform C;
    C.q = MulT(A.q, B.q);
    C.p = MulT(A.q, B.p - A.p);
    return C;
}

struct Sweep {
    Vec2 localCenter;
    Vec2 c0, c;
    float a0 = 0.0f, a = 0.0f;
    float alpha0 = 0.0f;
    void GetTransform(Transform* xf, float beta) const {
        xf->p = (1.0f - beta) * c0 + beta * c;
        float angle = (1.0f - beta) * a0 + beta * a;
        xf->q.Set(angle);
        xf->p -= Mul(xf->q, localCenter);
    }
    void Advance(float alpha) {
        float beta = (alpha - alpha0) / (1.0f - alpha0);
        c0 += beta * (c - c0);
        a0 += beta * (a - a0);
        alpha0 = alpha;
    }
    void Normalize() {
        float twoPi = 2.0f * kPi;
        float d = twoPi * floorf(a0 / twoPi);
        a0 -= d;
        a -= d;
    }
};

struct AABB {
    Vec2 lowerBound, upperBound;
    Vec2 GetCenter() const { return 0.5f * (lowerBound + upperBound); }
    Vec2 GetExtents() const { return 0.5f * (upperBound - lowerBound); }
    void Combine(const AABB& a1, const AABB& a2) {
        lowerBound = Min(a1.lowerBound, a2.lowerBound);
        upperBound = Max(a1.upperBound, a2.upperBound);
    }
    bool Contains(const AABB& aabb) const {
        bool result = true;
        result = result && lowerBound.x <= aabb.lowerBound.x;
        result = result && lowerBound.y <= aabb.lowerBound.y;
        result = result && aabb.upperBound.x <= upperBound.x;
        result = result && aabb.upperBound.y <= upperBound.y;
        return result;
    }
};
inline bool TestOverlap(const AABB& a, const AABB& b) {
    Vec2 d1 = b.lowerBound - a.upperBound;
    Vec2 d2 = a.lowerBound - b.upperBound;
    if (d1.x > 0.0f || d1.y > 0.0f) return false;
    if (d2.x > 0.0f || d2.y > 0.0f) return false;
    return true;
}

}  // namespace b2o
