// Sequential-impulse solver, joints, islands, sleeping and the per-step driver (one world per thread).
// Replaces b2World::Solve / b2Island::Solve / b2ContactSolver / b2RevoluteJoint / b2FrictionJoint of Box2D
// 2.3.1, reached by the reference through b2World::Step
// (src/sim_env/Box2DWorld.cpp:3279).  Island body/joint/contact order follows upstream's DFS so that the
// Gauss-Seidel sweeps visit constraints in the same order.
#pragma once
#include "b2g_device.cuh"

namespace b2g {

// ---- body velocity / position access (solver works in place on the world state) ---------------------------
struct Vel {
    V2 v;
    float w;
};
DEV Vel ldVel(const Ctx& c, int b) {
    Vel r;
    r.v = ldV2(c, bslot(c, b, SB_VX));
    r.w = ldF(c, bslot(c, b, SB_W));
    return r;
}
DEV void stVel(const Ctx& c, int b, const Vel& r) {
    stV2(c, bslot(c, b, SB_VX), r.v);
    stF(c, bslot(c, b, SB_W), r.w);
}
struct Pos {
    V2 c;
    float a;
};
DEV Pos ldPos(const Ctx& c, int b) {
    Pos r;
    r.c = ldV2(c, bslot(c, b, SB_CX));
    r.a = ldF(c, bslot(c, b, SB_A));
    return r;
}
DEV void stPos(const Ctx& c, int b, const Pos& r) {
    stV2(c, bslot(c, b, SB_CX), r.c);
    stF(c, bslot(c, b, SB_A), r.a);
}
DEV float invMass(const Ctx& c, int b) { return mF(c, mbody(c, b, MB_INVMASS)); }
DEV float invI(const Ctx& c, int b) { return mF(c, mbody(c, b, MB_INVI)); }

// =================================================================================================== joints
// b2FrictionJoint / b2RevoluteJoint::InitVelocityConstraints
DEVN void jointInit(const Ctx& c, int j, float dtRatio) {
    int type = mI(c, mjoint(c, j, MJ_TYPE));
    int bA = mI(c, mjoint(c, j, MJ_BODYA));
    int bB = mI(c, mjoint(c, j, MJ_BODYB));
    float aA = ldF(c, bslot(c, bA, SB_A));
    float aB = ldF(c, bslot(c, bB, SB_A));
    Vel A = ldVel(c, bA);
    Vel B = ldVel(c, bB);
    Rot qA = rotOf(aA), qB = rotOf(aB);
    V2 la = mk(mF(c, mjoint(c, j, MJ_LAX)), mF(c, mjoint(c, j, MJ_LAY)));
    V2 lb = mk(mF(c, mjoint(c, j, MJ_LBX)), mF(c, mjoint(c, j, MJ_LBY)));
    V2 rA = mul(qA, la - localCenter(c, bA));
    V2 rB = mul(qB, lb - localCenter(c, bB));
    float mA = invMass(c, bA), mB = invMass(c, bB);
    float iA = invI(c, bA), iB = invI(c, bB);
    stV2(c, jwslot(c, j, WJ_RAX), rA);
    stV2(c, jwslot(c, j, WJ_RBX), rB);

    if (type == 0) {
        float kxx = mA + mB + iA * rA.y * rA.y + iB * rB.y * rB.y;
        float kxy = -iA * rA.x * rA.y - iB * rB.x * rB.y;
        float kyx = kxy;
        float kyy = mA + mB + iA * rA.x * rA.x + iB * rB.x * rB.x;
        // b2Mat22::GetInverse
        float a = kxx, b = kyx, cc = kxy, d = kyy;
        float det = a * d - b * cc;
        if (det != 0.0f) det = 1.0f / det;
        stF(c, jwslot(c, j, WJ_M0 + 0), det * d);    // ex.x
        stF(c, jwslot(c, j, WJ_M0 + 1), -det * cc);  // ex.y
        stF(c, jwslot(c, j, WJ_M0 + 2), -det * b);   // ey.x
        stF(c, jwslot(c, j, WJ_M0 + 3), det * a);    // ey.y
        float angularMass = iA + iB;
        if (angularMass > 0.0f) angularMass = 1.0f / angularMass;
        stF(c, jwslot(c, j, WJ_M0 + 4), angularMass);
        // warm start
        V2 P = mk(ldF(c, jslot(c, j, SJ_IX)) * dtRatio, ldF(c, jslot(c, j, SJ_IY)) * dtRatio);
        float ang = ldF(c, jslot(c, j, SJ_IZ)) * dtRatio;
        stV2(c, jslot(c, j, SJ_IX), P);
        stF(c, jslot(c, j, SJ_IZ), ang);
        A.v = A.v - mA * P;
        A.w -= iA * (cross(rA, P) + ang);
        B.v = B.v + mB * P;
        B.w += iB * (cross(rB, P) + ang);
    } else {
        bool fixedRotation = (iA + iB == 0.0f);
        float exx = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
        float eyx = -rA.y * rA.x * iA - rB.y * rB.x * iB;
        float ezx = -rA.y * iA - rB.y * iB;
        float exy = eyx;
        float eyy = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
        float ezy = rA.x * iA + rB.x * iB;
        float exz = ezx;
        float eyz = ezy;
        float ezz = iA + iB;
        stF(c, jwslot(c, j, WJ_M0 + 0), exx); stF(c, jwslot(c, j, WJ_M0 + 1), exy); stF(c, jwslot(c, j, WJ_M0 + 2), exz);
        stF(c, jwslot(c, j, WJ_M0 + 3), eyx); stF(c, jwslot(c, j, WJ_M0 + 4), eyy); stF(c, jwslot(c, j, WJ_M0 + 5), eyz);
        stF(c, jwslot(c, j, WJ_M0 + 6), ezx); stF(c, jwslot(c, j, WJ_M0 + 7), ezy); stF(c, jwslot(c, j, WJ_M0 + 8), ezz);
        float motorMass = iA + iB;
        if (motorMass > 0.0f) motorMass = 1.0f / motorMass;
        stF(c, jwslot(c, j, WJ_M0 + 9), motorMass);

        int flags = ldI(c, jslot(c, j, SJ_FLAGS));
        bool enableMotor = (flags & 4) != 0;
        int limitState = flags & 3;
        bool enableLimit = (mI(c, mjoint(c, j, MJ_FLAGS)) & 1) != 0;
        float motorImpulse = ldF(c, jslot(c, j, SJ_MOTOR_IMPULSE));
        V3 imp;
        imp.x = ldF(c, jslot(c, j, SJ_IX)); imp.y = ldF(c, jslot(c, j, SJ_IY)); imp.z = ldF(c, jslot(c, j, SJ_IZ));
        if (enableMotor == false || fixedRotation) motorImpulse = 0.0f;
        if (enableLimit && fixedRotation == false) {
            float lower = mF(c, mjoint(c, j, MJ_LOWER)), upper = mF(c, mjoint(c, j, MJ_UPPER));
            float jointAngle = aB - aA - mF(c, mjoint(c, j, MJ_REF));
            if (absT(upper - lower) < 2.0f * B2G_ANGULAR_SLOP) {
                limitState = 3;
            } else if (jointAngle <= lower) {
                if (limitState != 1) imp.z = 0.0f;
                limitState = 1;
            } else if (jointAngle >= upper) {
                if (limitState != 2) imp.z = 0.0f;
                limitState = 2;
            } else {
                limitState = 0;
                imp.z = 0.0f;
            }
        } else {
            limitState = 0;
        }
        stI(c, jslot(c, j, SJ_FLAGS), (flags & ~3) | limitState);
        imp.x *= dtRatio; imp.y *= dtRatio; imp.z *= dtRatio;
        motorImpulse *= dtRatio;
        stF(c, jslot(c, j, SJ_IX), imp.x); stF(c, jslot(c, j, SJ_IY), imp.y); stF(c, jslot(c, j, SJ_IZ), imp.z);
        stF(c, jslot(c, j, SJ_MOTOR_IMPULSE), motorImpulse);
        V2 P = mk(imp.x, imp.y);
        A.v = A.v - mA * P;
        A.w -= iA * (cross(rA, P) + motorImpulse + imp.z);
        B.v = B.v + mB * P;
        B.w += iB * (cross(rB, P) + motorImpulse + imp.z);
    }
    stVel(c, bA, A);
    stVel(c, bB, B);
}

// b2FrictionJoint::SolveVelocityConstraints
DEV void frictionSolveVel(const Ctx& c, int j) {
    int bA = mI(c, mjoint(c, j, MJ_BODYA));
    int bB = mI(c, mjoint(c, j, MJ_BODYB));
    Vel A = ldVel(c, bA);
    Vel B = ldVel(c, bB);
    float mA = invMass(c, bA), mB = invMass(c, bB);
    float iA = invI(c, bA), iB = invI(c, bB);
    V2 rA = ldV2(c, jwslot(c, j, WJ_RAX));
    V2 rB = ldV2(c, jwslot(c, j, WJ_RBX));
    float h = c.dt;
    {
        float Cdot = B.w - A.w;
        float impulse = -ldF(c, jwslot(c, j, WJ_M0 + 4)) * Cdot;
        float oldImpulse = ldF(c, jslot(c, j, SJ_IZ));
        float maxImpulse = h * mF(c, mjoint(c, j, MJ_MAXTORQUE));
        float acc = clampT(oldImpulse + impulse, -maxImpulse, maxImpulse);
        stF(c, jslot(c, j, SJ_IZ), acc);
        impulse = acc - oldImpulse;
        A.w -= iA * impulse;
        B.w += iB * impulse;
    }
    {
        V2 Cdot = B.v + cross(B.w, rB) - A.v - cross(A.w, rA);
        float exx = ldF(c, jwslot(c, j, WJ_M0 + 0)), exy = ldF(c, jwslot(c, j, WJ_M0 + 1));
        float eyx = ldF(c, jwslot(c, j, WJ_M0 + 2)), eyy = ldF(c, jwslot(c, j, WJ_M0 + 3));
        V2 impulse = -mk(exx * Cdot.x + eyx * Cdot.y, exy * Cdot.x + eyy * Cdot.y);
        V2 oldImpulse = ldV2(c, jslot(c, j, SJ_IX));
        V2 acc = oldImpulse + impulse;
        float maxImpulse = h * mF(c, mjoint(c, j, MJ_MAXFORCE));
        if (dot(acc, acc) > maxImpulse * maxImpulse) {
            normalize(acc);
            acc = mk(acc.x * maxImpulse, acc.y * maxImpulse);
        }
        stV2(c, jslot(c, j, SJ_IX), acc);
        impulse = acc - oldImpulse;
        A.v = A.v - mA * impulse;
        A.w -= iA * cross(rA, impulse);
        B.v = B.v + mB * impulse;
        B.w += iB * cross(rB, impulse);
    }
    stVel(c, bA, A);
    stVel(c, bB, B);
}

DEV V3 solve33(const float* m, V3 b) {  // m = ex.xyz ey.xyz ez.xyz ; b2Mat33::Solve33
    V3 ex = {m[0], m[1], m[2]}, ey = {m[3], m[4], m[5]}, ez = {m[6], m[7], m[8]};
    V3 cyz = {ey.y * ez.z - ey.z * ez.y, ey.z * ez.x - ey.x * ez.z, ey.x * ez.y - ey.y * ez.x};
    float det = ex.x * cyz.x + ex.y * cyz.y + ex.z * cyz.z;
    if (det != 0.0f) det = 1.0f / det;
    V3 x;
    x.x = det * (b.x * cyz.x + b.y * cyz.y + b.z * cyz.z);
    V3 cbz = {b.y * ez.z - b.z * ez.y, b.z * ez.x - b.x * ez.z, b.x * ez.y - b.y * ez.x};
    x.y = det * (ex.x * cbz.x + ex.y * cbz.y + ex.z * cbz.z);
    V3 cyb = {ey.y * b.z - ey.z * b.y, ey.z * b.x - ey.x * b.z, ey.x * b.y - ey.y * b.x};
    x.z = det * (ex.x * cyb.x + ex.y * cyb.y + ex.z * cyb.z);
    return x;
}

// b2RevoluteJoint::SolveVelocityConstraints
DEV void revoluteSolveVel(const Ctx& c, int j) {
    int bA = mI(c, mjoint(c, j, MJ_BODYA));
    int bB = mI(c, mjoint(c, j, MJ_BODYB));
    Vel A = ldVel(c, bA);
    Vel B = ldVel(c, bB);
    float mA = invMass(c, bA), mB = invMass(c, bB);
    float iA = invI(c, bA), iB = invI(c, bB);
    V2 rA = ldV2(c, jwslot(c, j, WJ_RAX));
    V2 rB = ldV2(c, jwslot(c, j, WJ_RBX));
    int flags = ldI(c, jslot(c, j, SJ_FLAGS));
    bool enableMotor = (flags & 4) != 0;
    int limitState = flags & 3;
    bool enableLimit = (mI(c, mjoint(c, j, MJ_FLAGS)) & 1) != 0;
    bool fixedRotation = (iA + iB == 0.0f);

    if (enableMotor && limitState != 3 && fixedRotation == false) {
        float Cdot = B.w - A.w - ldF(c, jslot(c, j, SJ_MOTOR_SPEED));
        float impulse = -ldF(c, jwslot(c, j, WJ_M0 + 9)) * Cdot;
        float oldImpulse = ldF(c, jslot(c, j, SJ_MOTOR_IMPULSE));
        float maxImpulse = c.dt * ldF(c, jslot(c, j, SJ_MAX_MOTOR_TORQUE));
        float acc = clampT(oldImpulse + impulse, -maxImpulse, maxImpulse);
        stF(c, jslot(c, j, SJ_MOTOR_IMPULSE), acc);
        impulse = acc - oldImpulse;
        A.w -= iA * impulse;
        B.w += iB * impulse;
    }
    float m[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) m[i] = ldF(c, jwslot(c, j, WJ_M0 + i));
    if (enableLimit && limitState != 0 && fixedRotation == false) {
        V2 Cdot1 = B.v + cross(B.w, rB) - A.v - cross(A.w, rA);
        float Cdot2 = B.w - A.w;
        V3 Cdot = {Cdot1.x, Cdot1.y, Cdot2};
        V3 s = solve33(m, Cdot);
        V3 impulse = {-s.x, -s.y, -s.z};
        V3 acc;
        acc.x = ldF(c, jslot(c, j, SJ_IX)); acc.y = ldF(c, jslot(c, j, SJ_IY)); acc.z = ldF(c, jslot(c, j, SJ_IZ));
        if (limitState == 3) {
            acc.x += impulse.x; acc.y += impulse.y; acc.z += impulse.z;
        } else {
            float newImpulse = acc.z + impulse.z;
            bool reduce = (limitState == 1) ? (newImpulse < 0.0f) : (newImpulse > 0.0f);
            if (reduce) {
                V2 rhs = -Cdot1 + acc.z * mk(m[6], m[7]);
                V2 reduced = solve22(m[0], m[3], m[1], m[4], rhs);
                impulse.x = reduced.x;
                impulse.y = reduced.y;
                impulse.z = -acc.z;
                acc.x += reduced.x;
                acc.y += reduced.y;
                acc.z = 0.0f;
            } else {
                acc.x += impulse.x; acc.y += impulse.y; acc.z += impulse.z;
            }
        }
        stF(c, jslot(c, j, SJ_IX), acc.x); stF(c, jslot(c, j, SJ_IY), acc.y); stF(c, jslot(c, j, SJ_IZ), acc.z);
        V2 P = mk(impulse.x, impulse.y);
        A.v = A.v - mA * P;
        A.w -= iA * (cross(rA, P) + impulse.z);
        B.v = B.v + mB * P;
        B.w += iB * (cross(rB, P) + impulse.z);
    } else {
        V2 Cdot = B.v + cross(B.w, rB) - A.v - cross(A.w, rA);
        V2 impulse = solve22(m[0], m[3], m[1], m[4], -Cdot);
        stF(c, jslot(c, j, SJ_IX), ldF(c, jslot(c, j, SJ_IX)) + impulse.x);
        stF(c, jslot(c, j, SJ_IY), ldF(c, jslot(c, j, SJ_IY)) + impulse.y);
        A.v = A.v - mA * impulse;
        A.w -= iA * cross(rA, impulse);
        B.v = B.v + mB * impulse;
        B.w += iB * cross(rB, impulse);
    }
    stVel(c, bA, A);
    stVel(c, bB, B);
}

// b2RevoluteJoint::SolvePositionConstraints (friction joints have none and report true)
DEV bool revoluteSolvePos(const Ctx& c, int j) {
    int bA = mI(c, mjoint(c, j, MJ_BODYA));
    int bB = mI(c, mjoint(c, j, MJ_BODYB));
    Pos A = ldPos(c, bA);
    Pos B = ldPos(c, bB);
    float mA = invMass(c, bA), mB = invMass(c, bB);
    float iA = invI(c, bA), iB = invI(c, bB);
    int limitState = ldI(c, jslot(c, j, SJ_FLAGS)) & 3;
    bool enableLimit = (mI(c, mjoint(c, j, MJ_FLAGS)) & 1) != 0;
    float angularError = 0.0f;
    float positionError = 0.0f;
    bool fixedRotation = (iA + iB == 0.0f);
    if (enableLimit && limitState != 0 && fixedRotation == false) {
        float motorMass = ldF(c, jwslot(c, j, WJ_M0 + 9));
        float lower = mF(c, mjoint(c, j, MJ_LOWER)), upper = mF(c, mjoint(c, j, MJ_UPPER));
        float angle = B.a - A.a - mF(c, mjoint(c, j, MJ_REF));
        float limitImpulse = 0.0f;
        if (limitState == 3) {
            float C = clampT(angle - lower, -B2G_MAX_ANGULAR_CORRECTION, B2G_MAX_ANGULAR_CORRECTION);
            limitImpulse = -motorMass * C;
            angularError = absT(C);
        } else if (limitState == 1) {
            float C = angle - lower;
            angularError = -C;
            C = clampT(C + B2G_ANGULAR_SLOP, -B2G_MAX_ANGULAR_CORRECTION, 0.0f);
            limitImpulse = -motorMass * C;
        } else if (limitState == 2) {
            float C = angle - upper;
            angularError = C;
            C = clampT(C - B2G_ANGULAR_SLOP, 0.0f, B2G_MAX_ANGULAR_CORRECTION);
            limitImpulse = -motorMass * C;
        }
        A.a -= iA * limitImpulse;
        B.a += iB * limitImpulse;
    }
    {
        Rot qA = rotOf(A.a), qB = rotOf(B.a);
        V2 la = mk(mF(c, mjoint(c, j, MJ_LAX)), mF(c, mjoint(c, j, MJ_LAY)));
        V2 lb = mk(mF(c, mjoint(c, j, MJ_LBX)), mF(c, mjoint(c, j, MJ_LBY)));
        V2 rA = mul(qA, la - localCenter(c, bA));
        V2 rB = mul(qB, lb - localCenter(c, bB));
        V2 C = B.c + rB - A.c - rA;
        positionError = length(C);
        float kxx = mA + mB + iA * rA.y * rA.y + iB * rB.y * rB.y;
        float kxy = -iA * rA.x * rA.y - iB * rB.x * rB.y;
        float kyx = kxy;
        float kyy = mA + mB + iA * rA.x * rA.x + iB * rB.x * rB.x;
        V2 impulse = -solve22(kxx, kyx, kxy, kyy, C);
        A.c = A.c - mA * impulse;
        A.a -= iA * cross(rA, impulse);
        B.c = B.c + mB * impulse;
        B.a += iB * cross(rB, impulse);
    }
    stPos(c, bA, A);
    stPos(c, bB, B);
    return positionError <= B2G_LINEAR_SLOP && angularError <= B2G_ANGULAR_SLOP;
}

// ================================================================================================= contacts
DEV void contactBodies(const Ctx& c, int k, int& bA, int& bB, int& fA, int& fB) {
    contactFixtures(c, k, fA, fB);
    bA = mI(c, mfix(c, fA, MF_BODY));
    bB = mI(c, mfix(c, fB, MF_BODY));
}

// b2ContactSolver constructor (impulse hand-over) + InitializeVelocityConstraints for slot k
DEVN void contactInit(const Ctx& c, int k, float dtRatio, bool warmStarting) {
    int bA, bB, fA, fB;
    contactBodies(c, k, bA, bB, fA, fB);
    int man = ldI(c, cslot(c, k, SC_MANIFOLD));
    int type = man & 0xff;
    int pointCount = (man >> 8) & 0xff;
    float mA = invMass(c, bA), mB = invMass(c, bB);
    float iA = invI(c, bA), iB = invI(c, bB);
    Pos PA = ldPos(c, bA), PB = ldPos(c, bB);
    Vel A = ldVel(c, bA), B = ldVel(c, bB);
    XF xfA, xfB;
    xfA.q = rotOf(PA.a);
    xfB.q = rotOf(PB.a);
    xfA.p = PA.c - mul(xfA.q, localCenter(c, bA));
    xfB.p = PB.c - mul(xfB.q, localCenter(c, bB));
    float restitution;
    {
        float rA_ = mF(c, mfix(c, fA, MF_RESTITUTION)), rB_ = mF(c, mfix(c, fB, MF_RESTITUTION));
        restitution = rA_ > rB_ ? rA_ : rB_;
    }
    // b2WorldManifold::Initialize
    V2 localNormal = ldV2(c, cslot(c, k, SC_LNX));
    V2 localPoint = ldV2(c, cslot(c, k, SC_LPX));
    V2 normal;
    V2 wp[2];
    const float radius = B2G_POLYGON_RADIUS;
    if (type == 1) {
        normal = mul(xfA.q, localNormal);
        V2 planePoint = mul(xfA, localPoint);
        for (int i = 0; i < pointCount; ++i) {
            V2 clipPoint = mul(xfB, ldV2(c, cslot(c, k, i == 0 ? SC_P0X : SC_P1X)));
            V2 cA = clipPoint + (radius - dot(clipPoint - planePoint, normal)) * normal;
            V2 cB = clipPoint - radius * normal;
            wp[i] = 0.5f * (cA + cB);
        }
    } else {
        normal = mul(xfB.q, localNormal);
        V2 planePoint = mul(xfB, localPoint);
        for (int i = 0; i < pointCount; ++i) {
            V2 clipPoint = mul(xfA, ldV2(c, cslot(c, k, i == 0 ? SC_P0X : SC_P1X)));
            V2 cB = clipPoint + (radius - dot(clipPoint - planePoint, normal)) * normal;
            V2 cA = clipPoint - radius * normal;
            wp[i] = 0.5f * (cA + cB);
        }
        normal = -normal;
    }
    stV2(c, cwslot(c, k, WC_NX), normal);
    V2 rAs[2], rBs[2];
    for (int jn = 0; jn < pointCount; ++jn) {
        int base = WC_PT0 + jn * WC_PT_STRIDE;
        V2 rA = wp[jn] - PA.c;
        V2 rB = wp[jn] - PB.c;
        rAs[jn] = rA;
        rBs[jn] = rB;
        float rnA = cross(rA, normal);
        float rnB = cross(rB, normal);
        float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
        float normalMass = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;
        V2 tangent = cross(normal, 1.0f);
        float rtA = cross(rA, tangent);
        float rtB = cross(rB, tangent);
        float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
        float tangentMass = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;
        float velocityBias = 0.0f;
        float vRel = dot(normal, B.v + cross(B.w, rB) - A.v - cross(A.w, rA));
        if (vRel < -B2G_VELOCITY_THRESHOLD) velocityBias = -restitution * vRel;
        float ni = 0.0f, ti = 0.0f;
        if (warmStarting) {
            ni = dtRatio * ldF(c, cslot(c, k, jn == 0 ? SC_N0 : SC_N1));
            ti = dtRatio * ldF(c, cslot(c, k, jn == 0 ? SC_T0 : SC_T1));
        }
        stV2(c, cwslot(c, k, base + 0), rA);
        stV2(c, cwslot(c, k, base + 2), rB);
        stF(c, cwslot(c, k, base + 4), normalMass);
        stF(c, cwslot(c, k, base + 5), tangentMass);
        stF(c, cwslot(c, k, base + 6), velocityBias);
        stF(c, cwslot(c, k, base + 7), ni);
        stF(c, cwslot(c, k, base + 8), ti);
    }
    int vcCount = pointCount;
    if (pointCount == 2) {
        float rn1A = cross(rAs[0], normal);
        float rn1B = cross(rBs[0], normal);
        float rn2A = cross(rAs[1], normal);
        float rn2B = cross(rBs[1], normal);
        float k11 = mA + mB + iA * rn1A * rn1A + iB * rn1B * rn1B;
        float k22 = mA + mB + iA * rn2A * rn2A + iB * rn2B * rn2B;
        float k12 = mA + mB + iA * rn1A * rn2A + iB * rn1B * rn2B;
        const float k_maxConditionNumber = 1000.0f;
        if (k11 * k11 < k_maxConditionNumber * (k11 * k22 - k12 * k12)) {
            stF(c, cwslot(c, k, WC_K + 0), k11);
            stF(c, cwslot(c, k, WC_K + 1), k12);
            stF(c, cwslot(c, k, WC_K + 2), k22);
            // K = [ex=(k11,k12), ey=(k12,k22)]; GetInverse
            float a = k11, b = k12, cc = k12, d = k22;
            float det = a * d - b * cc;
            if (det != 0.0f) det = 1.0f / det;
            stF(c, cwslot(c, k, WC_NM + 0), det * d);    // ex.x
            stF(c, cwslot(c, k, WC_NM + 1), -det * cc);  // ex.y ( == ey.x )
            stF(c, cwslot(c, k, WC_NM + 2), det * a);    // ey.y
        } else {
            vcCount = 1;
        }
    }
    stI(c, cwslot(c, k, WC_COUNT), vcCount);
}

// b2ContactSolver::WarmStart for slot k
DEV void contactWarmStart(const Ctx& c, int k) {
    int bA, bB, fA, fB;
    contactBodies(c, k, bA, bB, fA, fB);
    float mA = invMass(c, bA), mB = invMass(c, bB);
    float iA = invI(c, bA), iB = invI(c, bB);
    Vel A = ldVel(c, bA), B = ldVel(c, bB);
    V2 normal = ldV2(c, cwslot(c, k, WC_NX));
    V2 tangent = cross(normal, 1.0f);
    int pointCount = ldI(c, cwslot(c, k, WC_COUNT));
    for (int jn = 0; jn < pointCount; ++jn) {
        int base = WC_PT0 + jn * WC_PT_STRIDE;
        V2 rA = ldV2(c, cwslot(c, k, base + 0));
        V2 rB = ldV2(c, cwslot(c, k, base + 2));
        V2 P = ldF(c, cwslot(c, k, base + 7)) * normal + ldF(c, cwslot(c, k, base + 8)) * tangent;
        A.w -= iA * cross(rA, P);
        A.v = A.v - mA * P;
        B.w += iB * cross(rB, P);
        B.v = B.v + mB * P;
    }
    stVel(c, bA, A);
    stVel(c, bB, B);
}

// b2ContactSolver::SolveVelocityConstraints for slot k
DEVN void contactSolveVel(const Ctx& c, int k) {
    int bA, bB, fA, fB;
    contactBodies(c, k, bA, bB, fA, fB);
    float mA = invMass(c, bA), mB = invMass(c, bB);
    float iA = invI(c, bA), iB = invI(c, bB);
    Vel A = ldVel(c, bA), B = ldVel(c, bB);
    V2 normal = ldV2(c, cwslot(c, k, WC_NX));
    V2 tangent = cross(normal, 1.0f);
    float friction = sqrtf(mF(c, mfix(c, fA, MF_FRICTION)) * mF(c, mfix(c, fB, MF_FRICTION)));
    int pointCount = ldI(c, cwslot(c, k, WC_COUNT));

    for (int jn = 0; jn < pointCount; ++jn) {
        int base = WC_PT0 + jn * WC_PT_STRIDE;
        V2 rA = ldV2(c, cwslot(c, k, base + 0));
        V2 rB = ldV2(c, cwslot(c, k, base + 2));
        V2 dv = B.v + cross(B.w, rB) - A.v - cross(A.w, rA);
        float vt = dot(dv, tangent) - 0.0f;
        float lambda = ldF(c, cwslot(c, k, base + 5)) * (-vt);
        float maxFriction = friction * ldF(c, cwslot(c, k, base + 7));
        float oldT = ldF(c, cwslot(c, k, base + 8));
        float newImpulse = clampT(oldT + lambda, -maxFriction, maxFriction);
        lambda = newImpulse - oldT;
        stF(c, cwslot(c, k, base + 8), newImpulse);
        V2 P = lambda * tangent;
        A.v = A.v - mA * P;
        A.w -= iA * cross(rA, P);
        B.v = B.v + mB * P;
        B.w += iB * cross(rB, P);
    }
    if (pointCount == 1) {
        int base = WC_PT0;
        V2 rA = ldV2(c, cwslot(c, k, base + 0));
        V2 rB = ldV2(c, cwslot(c, k, base + 2));
        V2 dv = B.v + cross(B.w, rB) - A.v - cross(A.w, rA);
        float vn = dot(dv, normal);
        float lambda = -ldF(c, cwslot(c, k, base + 4)) * (vn - ldF(c, cwslot(c, k, base + 6)));
        float oldN = ldF(c, cwslot(c, k, base + 7));
        float newImpulse = fmaxT(oldN + lambda, 0.0f);
        lambda = newImpulse - oldN;
        stF(c, cwslot(c, k, base + 7), newImpulse);
        V2 P = lambda * normal;
        A.v = A.v - mA * P;
        A.w -= iA * cross(rA, P);
        B.v = B.v + mB * P;
        B.w += iB * cross(rB, P);
    } else if (pointCount == 2) {
        int b1 = WC_PT0, b2 = WC_PT0 + WC_PT_STRIDE;
        V2 rA1 = ldV2(c, cwslot(c, k, b1 + 0)), rB1 = ldV2(c, cwslot(c, k, b1 + 2));
        V2 rA2 = ldV2(c, cwslot(c, k, b2 + 0)), rB2 = ldV2(c, cwslot(c, k, b2 + 2));
        V2 a = mk(ldF(c, cwslot(c, k, b1 + 7)), ldF(c, cwslot(c, k, b2 + 7)));
        V2 dv1 = B.v + cross(B.w, rB1) - A.v - cross(A.w, rA1);
        V2 dv2 = B.v + cross(B.w, rB2) - A.v - cross(A.w, rA2);
        float vn1 = dot(dv1, normal);
        float vn2 = dot(dv2, normal);
        V2 b;
        b.x = vn1 - ldF(c, cwslot(c, k, b1 + 6));
        b.y = vn2 - ldF(c, cwslot(c, k, b2 + 6));
        float k11 = ldF(c, cwslot(c, k, WC_K + 0)), k12 = ldF(c, cwslot(c, k, WC_K + 1)), k22 = ldF(c, cwslot(c, k, WC_K + 2));
        float n11 = ldF(c, cwslot(c, k, WC_NM + 0)), n12 = ldF(c, cwslot(c, k, WC_NM + 1)), n22 = ldF(c, cwslot(c, k, WC_NM + 2));
        b = b - mk(k11 * a.x + k12 * a.y, k12 * a.x + k22 * a.y);
        V2 x;
        bool solved = false;
        // case 1
        x = -mk(n11 * b.x + n12 * b.y, n12 * b.x + n22 * b.y);
        if (x.x >= 0.0f && x.y >= 0.0f) solved = true;
        if (!solved) {  // case 2
            x.x = -ldF(c, cwslot(c, k, b1 + 4)) * b.x;
            x.y = 0.0f;
            vn2 = k12 * x.x + b.y;
            if (x.x >= 0.0f && vn2 >= 0.0f) solved = true;
        }
        if (!solved) {  // case 3
            x.x = 0.0f;
            x.y = -ldF(c, cwslot(c, k, b2 + 4)) * b.y;
            vn1 = k12 * x.y + b.x;
            if (x.y >= 0.0f && vn1 >= 0.0f) solved = true;
        }
        if (!solved) {  // case 4
            x.x = 0.0f;
            x.y = 0.0f;
            vn1 = b.x;
            vn2 = b.y;
            if (vn1 >= 0.0f && vn2 >= 0.0f) solved = true;
        }
        if (solved) {
            V2 d = x - a;
            V2 P1 = d.x * normal;
            V2 P2 = d.y * normal;
            A.v = A.v - mA * (P1 + P2);
            A.w -= iA * (cross(rA1, P1) + cross(rA2, P2));
            B.v = B.v + mB * (P1 + P2);
            B.w += iB * (cross(rB1, P1) + cross(rB2, P2));
            stF(c, cwslot(c, k, b1 + 7), x.x);
            stF(c, cwslot(c, k, b2 + 7), x.y);
        }
    }
    stVel(c, bA, A);
    stVel(c, bB, B);
}

// b2ContactSolver::StoreImpulses for slot k
DEV void contactStoreImpulses(const Ctx& c, int k) {
    int n = ldI(c, cwslot(c, k, WC_COUNT));
    for (int jn = 0; jn < n; ++jn) {
        int base = WC_PT0 + jn * WC_PT_STRIDE;
        stF(c, cslot(c, k, jn == 0 ? SC_N0 : SC_N1), ldF(c, cwslot(c, k, base + 7)));
        stF(c, cslot(c, k, jn == 0 ? SC_T0 : SC_T1), ldF(c, cwslot(c, k, base + 8)));
    }
}

// b2ContactSolver::SolvePositionConstraints / SolveTOIPositionConstraints for slot k.
// toiA/toiB < 0 selects the regular solver.  Returns the minimum separation seen.
DEVN float contactSolvePos(const Ctx& c, int k, int toiA, int toiB, float minSeparation) {
    int bA, bB, fA, fB;
    contactBodies(c, k, bA, bB, fA, fB);
    int man = ldI(c, cslot(c, k, SC_MANIFOLD));
    int type = man & 0xff;
    int pointCount = (man >> 8) & 0xff;
    bool toi = toiA >= 0;
    float mA = invMass(c, bA), iA = invI(c, bA);
    float mB = invMass(c, bB), iB = invI(c, bB);
    if (toi) {
        if (!(bA == toiA || bA == toiB)) { mA = 0.0f; iA = 0.0f; }
        if (!(bB == toiA || bB == toiB)) { mB = 0.0f; iB = 0.0f; }
    }
    V2 lcA = localCenter(c, bA), lcB = localCenter(c, bB);
    Pos A = ldPos(c, bA), B = ldPos(c, bB);
    V2 localNormal = ldV2(c, cslot(c, k, SC_LNX));
    V2 localPoint = ldV2(c, cslot(c, k, SC_LPX));
    const float radius = B2G_POLYGON_RADIUS;
    for (int jn = 0; jn < pointCount; ++jn) {
        XF xfA, xfB;
        xfA.q = rotOf(A.a);
        xfB.q = rotOf(B.a);
        xfA.p = A.c - mul(xfA.q, lcA);
        xfB.p = B.c - mul(xfB.q, lcB);
        V2 lp = ldV2(c, cslot(c, k, jn == 0 ? SC_P0X : SC_P1X));
        V2 normal, point;
        float separation;
        if (type == 1) {
            normal = mul(xfA.q, localNormal);
            V2 planePoint = mul(xfA, localPoint);
            V2 clipPoint = mul(xfB, lp);
            separation = dot(clipPoint - planePoint, normal) - radius - radius;
            point = clipPoint;
        } else {
            normal = mul(xfB.q, localNormal);
            V2 planePoint = mul(xfB, localPoint);
            V2 clipPoint = mul(xfA, lp);
            separation = dot(clipPoint - planePoint, normal) - radius - radius;
            point = clipPoint;
            normal = -normal;
        }
        V2 rA = point - A.c;
        V2 rB = point - B.c;
        minSeparation = fminT(minSeparation, separation);
        float C = clampT((toi ? B2G_TOI_BAUMGARTE : B2G_BAUMGARTE) * (separation + B2G_LINEAR_SLOP), -B2G_MAX_LINEAR_CORRECTION, 0.0f);
        float rnA = cross(rA, normal);
        float rnB = cross(rB, normal);
        float K = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
        float impulse = K > 0.0f ? -C / K : 0.0f;
        V2 P = impulse * normal;
        A.c = A.c - mA * P;
        A.a -= iA * cross(rA, P);
        B.c = B.c + mB * P;
        B.a += iB * cross(rB, P);
    }
    stPos(c, bA, A);
    stPos(c, bB, B);
    return minSeparation;
}

// ================================================================================================== islands
struct Island {
    unsigned char bodies[kMaxBodies];
    unsigned char joints[kMaxJoints];
    unsigned char contacts[kMaxContacts];
    int nb, nj, nc;
};

// integrate positions with the translation / rotation caps (b2Island::Solve step 6, also SolveTOI)
DEV void integratePosition(const Ctx& c, int b, float h) {
    Pos P = ldPos(c, b);
    Vel V = ldVel(c, b);
    V2 translation = h * V.v;
    if (dot(translation, translation) > B2G_MAX_TRANSLATION * B2G_MAX_TRANSLATION) {
        float ratio = B2G_MAX_TRANSLATION / length(translation);
        V.v = mk(V.v.x * ratio, V.v.y * ratio);
    }
    float rotation = h * V.w;
    if (rotation * rotation > B2G_MAX_ROTATION * B2G_MAX_ROTATION) {
        float ratio = B2G_MAX_ROTATION / absT(rotation);
        V.w *= ratio;
    }
    P.c = P.c + h * V.v;
    P.a += h * V.w;
    stPos(c, b, P);
    stVel(c, b, V);
}

// b2Island::Solve
DEVN void islandSolve(const Ctx& c, const Island& isl, float dtRatio, bool allowSleep) {
    float h = c.dt;
    for (int i = 0; i < isl.nb; ++i) {
        int b = isl.bodies[i];
        Pos P = ldPos(c, b);
        stV2(c, bslot(c, b, SB_C0X), P.c);
        stF(c, bslot(c, b, SB_A0), P.a);
        if (bodyType(c, b) == 2) {
            Vel V = ldVel(c, b);
            V2 f = ldV2(c, bslot(c, b, SB_FX));
            float tq = ldF(c, bslot(c, b, SB_TORQUE));
            float im = invMass(c, b), ii = invI(c, b);
            // gravity is (0,0), gravityScale 1, damping 0 (Box2DWorld.cpp:3668; b2BodyDef defaults)
            V2 acc = 1.0f * mk(0.0f, 0.0f) + im * f;
            V.v = V.v + h * acc;
            V.w += h * ii * tq;
            float ld = 1.0f / (1.0f + h * 0.0f);
            V.v = mk(V.v.x * ld, V.v.y * ld);
            V.w *= ld;
            stVel(c, b, V);
        }
    }
    for (int i = 0; i < isl.nc; ++i) contactInit(c, isl.contacts[i], dtRatio, true);
    for (int i = 0; i < isl.nc; ++i) contactWarmStart(c, isl.contacts[i]);
    for (int i = 0; i < isl.nj; ++i) jointInit(c, isl.joints[i], dtRatio);

    for (int it = 0; it < c.velIters; ++it) {
        for (int i = 0; i < isl.nj; ++i) {
            int j = isl.joints[i];
            if (mI(c, mjoint(c, j, MJ_TYPE)) == 0) frictionSolveVel(c, j);
            else revoluteSolveVel(c, j);
        }
        for (int i = 0; i < isl.nc; ++i) contactSolveVel(c, isl.contacts[i]);
    }
    for (int i = 0; i < isl.nc; ++i) contactStoreImpulses(c, isl.contacts[i]);

    for (int i = 0; i < isl.nb; ++i) integratePosition(c, isl.bodies[i], h);

    bool positionSolved = false;
    for (int it = 0; it < c.posIters; ++it) {
        float minSeparation = 0.0f;
        for (int i = 0; i < isl.nc; ++i) minSeparation = contactSolvePos(c, isl.contacts[i], -1, -1, minSeparation);
        bool contactsOkay = minSeparation >= -3.0f * B2G_LINEAR_SLOP;
        bool jointsOkay = true;
        for (int i = 0; i < isl.nj; ++i) {
            int j = isl.joints[i];
            bool ok = true;
            if (mI(c, mjoint(c, j, MJ_TYPE)) == 1) ok = revoluteSolvePos(c, j);
            jointsOkay = jointsOkay && ok;
        }
        if (contactsOkay && jointsOkay) {
            positionSolved = true;
            break;
        }
    }
    for (int i = 0; i < isl.nb; ++i) syncTransform(c, isl.bodies[i]);

    if (allowSleep) {
        float minSleepTime = FLT_MAX;
        const float linTolSqr = B2G_LINEAR_SLEEP_TOL * B2G_LINEAR_SLEEP_TOL;
        const float angTolSqr = B2G_ANGULAR_SLEEP_TOL * B2G_ANGULAR_SLEEP_TOL;
        for (int i = 0; i < isl.nb; ++i) {
            int b = isl.bodies[i];
            if (bodyType(c, b) == 0) continue;
            Vel V = ldVel(c, b);
            if (V.w * V.w > angTolSqr || dot(V.v, V.v) > linTolSqr) {
                stF(c, bslot(c, b, SB_SLEEP), 0.0f);
                minSleepTime = 0.0f;
            } else {
                float st = ldF(c, bslot(c, b, SB_SLEEP)) + h;
                stF(c, bslot(c, b, SB_SLEEP), st);
                minSleepTime = fminT(minSleepTime, st);
            }
        }
        if (minSleepTime >= B2G_TIME_TO_SLEEP && positionSolved) {
            for (int i = 0; i < isl.nb; ++i) sleepBody(c, isl.bodies[i]);
        }
    }
}

// b2World::Solve: islands by DFS in upstream order, then SynchronizeFixtures + FindNewContacts
DEVN void solveIslands(const Ctx& c, float dtRatio, bool allowSleep) {
    const int nb = c.h->nb;
    uint64_t bodyFlag = 0ull;
    uint64_t jointFlag[2] = {0ull, 0ull};
    uint64_t contactFlag[2] = {0ull, 0ull};
    unsigned char stack[kMaxBodies];
    Island isl;
    for (int seed = nb - 1; seed >= 0; --seed) {  // body list: newest first
        if ((bodyFlag >> seed) & 1ull) continue;
        if (!isAwake(c, seed)) continue;
        if (bodyType(c, seed) == 0) continue;
        isl.nb = 0; isl.nj = 0; isl.nc = 0;
        int sp = 0;
        stack[sp++] = (unsigned char)seed;
        bodyFlag |= 1ull << seed;
        while (sp > 0) {
            int b = stack[--sp];
            isl.bodies[isl.nb++] = (unsigned char)b;
            wakeBody(c, b);
            if (bodyType(c, b) == 0) continue;
            // contact edges of b, newest first
            for (int k = contactCount(c) - 1; k >= 0; --k) {
                int fA, fB;
                contactFixtures(c, k, fA, fB);
                int cA = mI(c, mfix(c, fA, MF_BODY));
                int cB = mI(c, mfix(c, fB, MF_BODY));
                if (cA != b && cB != b) continue;
                if ((contactFlag[k >> 6] >> (k & 63)) & 1ull) continue;
                int fl = ldI(c, cslot(c, k, SC_FLAGS));
                if ((fl & CF_ENABLED) == 0 || (fl & CF_TOUCHING) == 0) continue;
                isl.contacts[isl.nc++] = (unsigned char)k;
                contactFlag[k >> 6] |= 1ull << (k & 63);
                int other = cA == b ? cB : cA;
                if ((bodyFlag >> other) & 1ull) continue;
                stack[sp++] = (unsigned char)other;
                bodyFlag |= 1ull << other;
            }
            // joint edges of b, newest first
            int js = c.h->oJEdge + mI(c, mbody(c, b, MB_JE_START));
            int jc = mI(c, mbody(c, b, MB_JE_COUNT));
            for (int e = 0; e < jc; ++e) {
                int packed = mI(c, js + e);
                int j = packed & 0xffff;
                int other = packed >> 16;
                if ((jointFlag[j >> 6] >> (j & 63)) & 1ull) continue;
                isl.joints[isl.nj++] = (unsigned char)j;
                jointFlag[j >> 6] |= 1ull << (j & 63);
                if ((bodyFlag >> other) & 1ull) continue;
                stack[sp++] = (unsigned char)other;
                bodyFlag |= 1ull << other;
            }
        }
        islandSolve(c, isl, dtRatio, allowSleep);
        for (int i = 0; i < isl.nb; ++i) {
            int b = isl.bodies[i];
            if (bodyType(c, b) == 0) bodyFlag &= ~(1ull << b);
        }
    }
    uint64_t moved = 0ull;
    for (int b = nb - 1; b >= 0; --b) {
        if (((bodyFlag >> b) & 1ull) == 0ull) continue;
        if (bodyType(c, b) == 0) continue;
        XF xf1;
        xf1.q = rotOf(ldF(c, bslot(c, b, SB_A0)));
        xf1.p = ldV2(c, bslot(c, b, SB_C0X)) - mul(xf1.q, localCenter(c, b));
        XF xf2 = ldXF(c, b);
        moved |= syncBodyFixtures(c, b, xf1, xf2);
    }
    if (moved) stMoved(c, ldMoved(c) | moved);
    findNewContacts(c);
}

}  // namespace b2g
