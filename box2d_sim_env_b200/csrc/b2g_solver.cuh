// Sequential-impulse solver, joints, islands, sleeping and the per-step driver (one world per thread).
// Replaces b2World::Solve / b2Island::Solve / b2ContactSolver / b2RevoluteJoint / b2FrictionJoint of Box2D
// 2.3.1, reached by the reference through b2World::Step
// (src/sim_env/Box2DWorld.cpp:3279).  Island body/joint/contact order follows upstream's DFS so that the
// Gauss-Seidel sweeps visit constraints in the same order.
#pragma once
#include "b2g_device.cuh"

namespace b2g {

// ---- body velocity / position access (solver works in place on the world state) ---------------------------
// One LDG.128 / STG.128 per body: the 4th word of the chunk (awake flags / alpha0) rides along untouched.
struct Vel {
    V2 v;
    float w;
    float keep;
};
DEV Vel ldVel(Ctx c, int b) {
    float4 q = ld4(c, bslot(c, b, SB_VEL4));
    Vel r;
    r.v = mk(q.x, q.y);
    r.w = q.z;
    r.keep = q.w;
    return r;
}
DEV void stVel(Ctx c, int b, const Vel& r) { st4(c, bslot(c, b, SB_VEL4), mk4(r.v.x, r.v.y, r.w, r.keep)); }
struct Pos {
    V2 c;
    float a;
    float keep;
};
DEV Pos ldPos(Ctx c, int b) {
    float4 q = ld4(c, bslot(c, b, SB_POS4));
    Pos r;
    r.c = mk(q.x, q.y);
    r.a = q.z;
    r.keep = q.w;
    return r;
}
DEV void stPos(Ctx c, int b, const Pos& r) { st4(c, bslot(c, b, SB_POS4), mk4(r.c.x, r.c.y, r.a, r.keep)); }
DEV float invMass(Ctx c, int b) { return mF(c, mbody(c, b, MB_INVMASS)); }
DEV float invI(Ctx c, int b) { return mF(c, mbody(c, b, MB_INVI)); }

// =================================================================================================== joints
// Every joint has a 128-byte solver record in the model blob (model.h: JointField) with precomputed state byte
// offsets, anchor arms and inverse masses, so a solve is LDS.128 broadcasts + LDG.128/STG.128 on the state.
struct JointRec {
    int base;  // word offset of the record in the blob
};
DEV JointRec jrec(Ctx c, int j) { JointRec r; r.base = mjoint(c, j, 0); return r; }
DEV int4 mi4(Ctx c, int off) { B2G_MODEL_OK(off, 4); return *reinterpret_cast<const int4*>(g_sm + kHdrWords + B2G_MO(c) + off); }
DEV Vel ldVelAt(Ctx c, int boff) {
    float4 q = ld4(c, boff + SB_VEL4);
    Vel r;
    r.v = mk(q.x, q.y);
    r.w = q.z;
    r.keep = q.w;
    return r;
}
DEV void stVelAt(Ctx c, int boff, const Vel& r) { st4(c, boff + SB_VEL4, mk4(r.v.x, r.v.y, r.w, r.keep)); }
DEV Pos ldPosAt(Ctx c, int boff) {
    float4 q = ld4(c, boff + SB_POS4);
    Pos r;
    r.c = mk(q.x, q.y);
    r.a = q.z;
    r.keep = q.w;
    return r;
}
DEV void stPosAt(Ctx c, int boff, const Pos& r) { st4(c, boff + SB_POS4, mk4(r.c.x, r.c.y, r.a, r.keep)); }
// Rotation of a body at the start of an island solve.  b2Island reads it as b2Rot(position.a) with position.a = m_sweep.a;
// every path that writes m_sweep.a (SynchronizeTransform, SetTransform, b2Body::Advance) also sets m_xf.q = b2Rot(a) with
// the same deterministic sin / cos, so the stored m_xf.q IS b2Rot(a), bit for bit, and the sin / cos is not recomputed.
DEV Rot ldRotAt(Ctx c, int boff) {
    Rot r;
    if (boff == 0) {  // body 0 is the ground: never moved, m_xf.q = b2Rot(0) = (0, 1) -- nothing to load
        r.s = 0.0f;
        r.c = 1.0f;
        return r;
    }
    float2 q = ld2(c, boff + SB_QS);
    r.s = q.x;
    r.c = q.y;
    return r;
}

// A static body never moves: b2Island copies v = w = +0 for it and every solver update is x -= 0 * impulse, which
// leaves +0 for any finite impulse.  Its velocity is therefore neither loaded nor stored.
DEV Vel zeroVel() {
    Vel r;
    r.v = mk(0.0f, 0.0f);
    r.w = 0.0f;
    r.keep = 0.0f;
    return r;
}

// b2FrictionJoint / b2RevoluteJoint::InitVelocityConstraints
// (friction and revolute joints; a prismatic joint goes to prismaticInit -- the caller dispatches, so that models without
// prismatic joints run the loops of solveIslands without the third kind: carrying its branches and calls through the hot
// loops cost 6 % on test_env.yaml)
DEVN void prismaticInit(Ctx c, int j, float dtRatio);
DEVN void jointInit(Ctx c, int j, float dtRatio) {
    const int jr = mjoint(c, j, 0);
    const int4 jh = mi4(c, jr + MJ_TYPE);     // type bodyA bodyB flags
    const int4 off = mi4(c, jr + MJ_SOFF);    // js jw bodyA bodyB (state byte offsets)
    const float4 arm = m4(c, jr + MJ_RA0X);   // localAnchorA - localCenterA, localAnchorB - localCenterB
    const float4 im = m4(c, jr + MJ_MA);      // mA iA mB iB
    const int type = jh.x;
    const bool staticA = (jh.w & 4) != 0, staticB = (jh.w & 8) != 0;
    const int js = off.x, jw = off.y;
    Vel A = staticA ? zeroVel() : ldVelAt(c, off.z);
    Vel B = staticB ? zeroVel() : ldVelAt(c, off.w);
    Rot qA = ldRotAt(c, off.z), qB = ldRotAt(c, off.w);
    V2 rA = mul(qA, mk(arm.x, arm.y));
    V2 rB = mul(qB, mk(arm.z, arm.w));
    float mA = im.x, mB = im.z;
    float iA = im.y, iB = im.w;

    if (type == 0) {
        float kxx = mA + mB + iA * rA.y * rA.y + iB * rB.y * rB.y;
        float kxy = -iA * rA.x * rA.y - iB * rB.x * rB.y;
        float kyx = kxy;
        float kyy = mA + mB + iA * rA.x * rA.x + iB * rB.x * rB.x;
        // b2Mat22::GetInverse; the inverse of a symmetric matrix is symmetric bit for bit (ex.y = -det * c, ey.x = -det * b, b == c)
        float a = kxx, b = kyx, cc = kxy, d = kyy;
        float det = a * d - b * cc;
        if (det != 0.0f) det = 1.0f / det;
        st4(c, jw + WJ_R4, mk4(rB.y, det * d, -det * cc, det * a));  // rB.y  M.xx  M.xy  M.yy
        // warm start
        float4 imp = ld4(c, js + SJ_IMPULSE4);
        V2 P = mk(imp.x * dtRatio, imp.y * dtRatio);
        float ang = imp.z * dtRatio;
        st4(c, js + SJ_IMPULSE4, mk4(P.x, P.y, ang, rB.x));
        A.v = A.v - mA * P;
        A.w -= iA * (cross(rA, P) + ang);
        B.v = B.v + mB * P;
        B.w += iB * (cross(rB, P) + ang);
    } else {
        bool fixedRotation = (iA + iB == 0.0f);
        float exx = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
        float eyx = -rA.y * rA.x * iA - rB.y * rB.x * iB;
        float ezx = -rA.y * iA - rB.y * iB;
        float eyy = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
        float ezy = rA.x * iA + rB.x * iB;
        // symmetric: ex.y = ey.x, ex.z = ez.x, ey.z = ez.y; ez.z = iA + iB is recomputed by the solver
        st4(c, jw + WJ_R4, mk4(rA.x, rA.y, rB.x, rB.y));
        st4(c, jw + WJ_M0, mk4(exx, eyx, eyy, ezx));

        float4 mot = ld4(c, js + SJ_MOTOR4);  // flags maxMotorTorque motorSpeed | K.zy
        int flags = __float_as_int(mot.x);
        bool enableMotor = (flags & 4) != 0;
        int limitState = flags & 3;
        bool enableLimit = (jh.w & 1) != 0;
        float4 si = ld4(c, js + SJ_IMPULSE4);
        float motorImpulse = si.w;
        V3 imp;
        imp.x = si.x; imp.y = si.y; imp.z = si.z;
        if (enableMotor == false || fixedRotation) motorImpulse = 0.0f;
        if (enableLimit && fixedRotation == false) {
            const float4 lim = m4(c, jr + MJ_REF);  // ref lower upper maxForce
            float lower = lim.y, upper = lim.z;
            float jointAngle = ldF(c, off.w + SB_A) - ldF(c, off.z + SB_A) - lim.x;
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
        mot.x = __int_as_float((flags & ~3) | limitState);
        mot.w = ezy;
        st4(c, js + SJ_MOTOR4, mot);
        if (enableLimit && limitState != 0 && fixedRotation == false) {
            // b2Mat33::Solve33's right-hand-side independent part for the velocity iterations: cross(ey, ez), 1 / det
            const float ezz = iA + iB;
            const float cx = eyy * ezz - ezy * ezy, cy = ezy * ezx - eyx * ezz, cz = eyx * ezy - eyy * ezx;
            float det = exx * cx + eyx * cy + ezx * cz;
            if (det != 0.0f) det = 1.0f / det;
            st4(c, jw + WJ_M1, mk4(cx, cy, cz, det));
        }
        imp.x *= dtRatio; imp.y *= dtRatio; imp.z *= dtRatio;
        motorImpulse *= dtRatio;
        st4(c, js + SJ_IMPULSE4, mk4(imp.x, imp.y, imp.z, motorImpulse));
        V2 P = mk(imp.x, imp.y);
        A.v = A.v - mA * P;
        A.w -= iA * (cross(rA, P) + motorImpulse + imp.z);
        B.v = B.v + mB * P;
        B.w += iB * (cross(rB, P) + motorImpulse + imp.z);
    }
    if (!staticA) stVelAt(c, off.z, A);
    if (!staticB) stVelAt(c, off.w, B);
}

// b2FrictionJoint::SolveVelocityConstraints.  bodyA is the static ground (every link's friction joint,
// Box2DWorld.cpp:134-147; enforced by the model builder): its velocity is the constant +0 and all of its updates are
// dead code, so neither its velocity nor its anchor arm is loaded.  Two joint chunks + one body chunk per solve.
struct FrictionSolve {
    int js, boff;
    float mB, iB, hmForce, hmTorque, angularMass;
    float4 wk, acc4;
    Vel B;
};
DEV FrictionSolve frictionLoad(Ctx c, int jr) {
    FrictionSolve f;
    const int4 off = mi4(c, jr + MJ_SOFF);         // js jw bodyA bodyB
    const float4 im = m4(c, jr + MJ_MA);           // mA iA mB iB
    const float4 hm = m4(c, jr + MJ_HMAXFORCE);    // dt * maxForce, dt * maxTorque, angularMass, -
    f.js = off.x;
    f.boff = off.w;
    f.mB = im.z;
    f.iB = im.w;
    f.hmForce = hm.x;
    f.hmTorque = hm.y;
    f.angularMass = hm.z;
    f.B = ldVelAt(c, off.w);
    f.wk = ld4(c, off.y + WJ_R4);                  // rB.y  M.xx  M.xy  M.yy
    f.acc4 = ld4(c, off.x + SJ_IMPULSE4);          // linearImpulse.xy angularImpulse rB.x
    return f;
}
// Straight-line: the saturation test is a select -- the normalised candidate is computed unconditionally and only selected
// where the reference takes the branch (same operations on the same values; measured 1.5 % faster than the branch).
DEV void frictionCompute(FrictionSolve& f) {
    const V2 rB = mk(f.acc4.w, f.wk.x);
    {
        float Cdot = f.B.w - 0.0f;
        float impulse = -f.angularMass * Cdot;
        float oldImpulse = f.acc4.z;
        float maxImpulse = f.hmTorque;
        float acc = clampT(oldImpulse + impulse, -maxImpulse, maxImpulse);
        f.acc4.z = acc;
        impulse = acc - oldImpulse;
        f.B.w += f.iB * impulse;
    }
    {
        V2 Cdot = f.B.v + cross(f.B.w, rB) - mk(0.0f, 0.0f) - mk(-0.0f, 0.0f);  // - vA - cross(wA, rA) with vA = wA = +0
        float exx = f.wk.y, exy = f.wk.z;
        float eyx = f.wk.z, eyy = f.wk.w;
        V2 impulse = -mk(exx * Cdot.x + eyx * Cdot.y, exy * Cdot.x + eyy * Cdot.y);
        V2 oldImpulse = mk(f.acc4.x, f.acc4.y);
        V2 acc = oldImpulse + impulse;
        float maxImpulse = f.hmForce;
        const bool saturated = dot(acc, acc) > maxImpulse * maxImpulse;
        // b2Vec2::Normalize, then * maxImpulse
        const float len = length(acc);
        const float inv = 1.0f / len;
        const V2 unit = len < FLT_EPSILON ? acc : mk(acc.x * inv, acc.y * inv);
        const V2 clamped = mk(unit.x * maxImpulse, unit.y * maxImpulse);
        acc = saturated ? clamped : acc;
        f.acc4.x = acc.x;
        f.acc4.y = acc.y;
        impulse = acc - oldImpulse;
        f.B.v = f.B.v + f.mB * impulse;
        f.B.w += f.iB * cross(rB, impulse);
    }
}
DEV void frictionStore(Ctx c, const FrictionSolve& f) {
    st4(c, f.js + SJ_IMPULSE4, f.acc4);
    stVelAt(c, f.boff, f.B);
}
DEV void frictionSolveVel(Ctx c, int jr) {
    FrictionSolve f = frictionLoad(c, jr);
    frictionCompute(f);
    frictionStore(c, f);
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

// b2Mat33::Solve33 with cross(ey, ez) and 1 / det taken from jointInit (same operations on the same values)
DEV V3 solve33pre(const float* m, float4 pre, V3 b) {
    V3 ex = {m[0], m[1], m[2]}, ey = {m[3], m[4], m[5]}, ez = {m[6], m[7], m[8]};
    const float det = pre.w;
    V3 x;
    x.x = det * (b.x * pre.x + b.y * pre.y + b.z * pre.z);
    V3 cbz = {b.y * ez.z - b.z * ez.y, b.z * ez.x - b.x * ez.z, b.x * ez.y - b.y * ez.x};
    x.y = det * (ex.x * cbz.x + ex.y * cbz.y + ex.z * cbz.z);
    V3 cyb = {ey.y * b.z - ey.z * b.y, ey.z * b.x - ey.x * b.z, ey.x * b.y - ey.y * b.x};
    x.z = det * (ex.x * cyb.x + ex.y * cyb.y + ex.z * cyb.z);
    return x;
}

// b2RevoluteJoint::SolveVelocityConstraints
DEV void revoluteSolveVel(Ctx c, int jr, int modelFlags) {
    const int4 off = mi4(c, jr + MJ_SOFF);  // js jw bodyA bodyB
    const float4 im = m4(c, jr + MJ_MA);    // mA iA mB iB
    const int js = off.x, jw = off.y;
    const bool staticA = (modelFlags & 4) != 0, staticB = (modelFlags & 8) != 0;
    Vel A = staticA ? zeroVel() : ldVelAt(c, off.z);
    Vel B = staticB ? zeroVel() : ldVelAt(c, off.w);
    float mA = im.x, mB = im.z;
    float iA = im.y, iB = im.w;
    const float4 r = ld4(c, jw + WJ_R4);
    const float4 k0 = ld4(c, jw + WJ_M0);       // K.xx K.xy(=K.yx) K.yy K.zx(=K.xz)
    const float4 mot = ld4(c, js + SJ_MOTOR4);  // flags maxMotorTorque motorSpeed K.zy(=K.yz)
    float4 imp4 = ld4(c, js + SJ_IMPULSE4);     // impulse.xyz motorImpulse
    const float motorMass = mF(c, jr + MJ_ROTMASS);
    V2 rA = mk(r.x, r.y);
    V2 rB = mk(r.z, r.w);
    int flags = __float_as_int(mot.x);
    bool enableMotor = (flags & 4) != 0;
    int limitState = flags & 3;
    bool enableLimit = (modelFlags & 1) != 0;
    bool fixedRotation = (iA + iB == 0.0f);

    if (enableMotor && limitState != 3 && fixedRotation == false) {
        float Cdot = B.w - A.w - mot.z;
        float impulse = -motorMass * Cdot;
        float oldImpulse = imp4.w;
        float maxImpulse = PRM().dt * mot.y;
        float acc = clampT(oldImpulse + impulse, -maxImpulse, maxImpulse);
        imp4.w = acc;
        impulse = acc - oldImpulse;
        A.w -= iA * impulse;
        B.w += iB * impulse;
    }
    if (enableLimit && limitState != 0 && fixedRotation == false) {
        float m[9] = {k0.x, k0.y, k0.w, k0.y, k0.z, mot.w, k0.w, mot.w, iA + iB};
        const float4 pre = ld4(c, jw + WJ_M1);  // cross(ey, ez), 1 / det (jointInit)
        V2 Cdot1 = B.v + cross(B.w, rB) - A.v - cross(A.w, rA);
        float Cdot2 = B.w - A.w;
        V3 Cdot = {Cdot1.x, Cdot1.y, Cdot2};
        V3 s = solve33pre(m, pre, Cdot);
        V3 impulse = {-s.x, -s.y, -s.z};
        V3 acc;
        acc.x = imp4.x; acc.y = imp4.y; acc.z = imp4.z;
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
        imp4.x = acc.x; imp4.y = acc.y; imp4.z = acc.z;
        V2 P = mk(impulse.x, impulse.y);
        A.v = A.v - mA * P;
        A.w -= iA * (cross(rA, P) + impulse.z);
        B.v = B.v + mB * P;
        B.w += iB * (cross(rB, P) + impulse.z);
    } else {
        V2 Cdot = B.v + cross(B.w, rB) - A.v - cross(A.w, rA);
        V2 impulse = solve22(k0.x, k0.y, k0.y, k0.z, -Cdot);
        imp4.x = imp4.x + impulse.x;
        imp4.y = imp4.y + impulse.y;
        A.v = A.v - mA * impulse;
        A.w -= iA * cross(rA, impulse);
        B.v = B.v + mB * impulse;
        B.w += iB * cross(rB, impulse);
    }
    st4(c, js + SJ_IMPULSE4, imp4);
    if (!staticA) stVelAt(c, off.z, A);
    if (!staticB) stVelAt(c, off.w, B);
}

// non-inlined copies for the velocity loop of models WITH prismatic joints (that loop is the rare one: it stays small)
DEVN void frictionSolveVelOutlined(Ctx c, int jr) { frictionSolveVel(c, jr); }
DEVN void revoluteSolveVelOutlined(Ctx c, int jr, int modelFlags) { revoluteSolveVel(c, jr, modelFlags); }

// ---- b2PrismaticJoint (2.3.1; created by the reference at Box2DWorld.cpp:920-940) ---------------------------------------------
// b2Mat33::Solve22 of the symmetric K = [k11 k12; k12 k22]
DEV V2 solveSym22(float k11, float k12, float k22, V2 b) {
    float a11 = k11, a12 = k12, a21 = k12, a22 = k22;
    float det = a11 * a22 - a12 * a21;
    if (det != 0.0f) det = 1.0f / det;
    V2 x;
    x.x = det * (a22 * b.x - a12 * b.y);
    x.y = det * (a11 * b.y - a21 * b.x);
    return x;
}

// b2PrismaticJoint::InitVelocityConstraints (warm start included)
DEVN void prismaticInit(Ctx c, int j, float dtRatio) {
    const int jr = mjoint(c, j, 0);
    const int4 jh = mi4(c, jr + MJ_TYPE);     // type bodyA bodyB flags
    const int4 off = mi4(c, jr + MJ_SOFF);    // js jw bodyA bodyB (state byte offsets)
    const float4 arm = m4(c, jr + MJ_RA0X);   // localAnchorA - localCenterA, localAnchorB - localCenterB
    const float4 im = m4(c, jr + MJ_MA);      // mA iA mB iB
    const bool staticA = (jh.w & 4) != 0, staticB = (jh.w & 8) != 0;
    const int js = off.x, jw = off.y;
    Vel A = staticA ? zeroVel() : ldVelAt(c, off.z);
    Vel B = staticB ? zeroVel() : ldVelAt(c, off.w);
    const Pos PA = ldPosAt(c, off.z), PB = ldPosAt(c, off.w);
    Rot qA = ldRotAt(c, off.z), qB = ldRotAt(c, off.w);
    V2 rA = mul(qA, mk(arm.x, arm.y));
    V2 rB = mul(qB, mk(arm.z, arm.w));
    V2 d = (PB.c - PA.c) + rB - rA;
    float mA = im.x, mB = im.z;
    float iA = im.y, iB = im.w;
    const V2 lx = mk(mF(c, jr + MJ_AXISX), mF(c, jr + MJ_AXISY));
    const V2 ly = cross(1.0f, lx);
    V2 axis = mul(qA, lx);
    float a1 = cross(d + rA, axis);
    float a2 = cross(rB, axis);
    V2 perp = mul(qA, ly);
    float s1 = cross(d + rA, perp);
    float s2 = cross(rB, perp);
    float k11 = mA + mB + iA * s1 * s1 + iB * s2 * s2;
    float k12 = iA * s1 + iB * s2;
    float k13 = iA * s1 * a1 + iB * s2 * a2;
    float k23 = iA * a1 + iB * a2;
    float k33 = mA + mB + iA * a1 * a1 + iB * a2 * a2;
    st4(c, jw + WJ_R4, mk4(axis.x, axis.y, perp.x, perp.y));
    st4(c, jw + WJ_M0, mk4(s1, s2, a1, a2));
    st4(c, jw + WJ_M1, mk4(k11, k12, k13, k23));

    float4 mot = ld4(c, js + SJ_MOTOR4);  // flags maxMotorForce motorSpeed | k33
    int flags = __float_as_int(mot.x);
    const bool enableMotor = (flags & 4) != 0;
    int limitState = flags & 3;
    const bool enableLimit = (jh.w & 1) != 0;
    float4 imp = ld4(c, js + SJ_IMPULSE4);  // impulse.xyz motorImpulse
    if (enableLimit) {
        const float4 lim = m4(c, jr + MJ_REF);  // ref lower upper -
        float jointTranslation = dot(axis, d);
        if (absT(lim.z - lim.y) < 2.0f * B2G_LINEAR_SLOP) {
            limitState = 3;
        } else if (jointTranslation <= lim.y) {
            if (limitState != 1) { limitState = 1; imp.z = 0.0f; }
        } else if (jointTranslation >= lim.z) {
            if (limitState != 2) { limitState = 2; imp.z = 0.0f; }
        } else {
            limitState = 0;
            imp.z = 0.0f;
        }
    } else {
        limitState = 0;
        imp.z = 0.0f;
    }
    if (enableMotor == false) imp.w = 0.0f;
    mot.x = __int_as_float((flags & ~3) | limitState);
    mot.w = k33;
    st4(c, js + SJ_MOTOR4, mot);
    imp.x *= dtRatio; imp.y *= dtRatio; imp.z *= dtRatio;
    imp.w *= dtRatio;
    st4(c, js + SJ_IMPULSE4, imp);
    V2 P = imp.x * perp + (imp.w + imp.z) * axis;
    float LA = imp.x * s1 + imp.y + (imp.w + imp.z) * a1;
    float LB = imp.x * s2 + imp.y + (imp.w + imp.z) * a2;
    A.v = A.v - mA * P;
    A.w -= iA * LA;
    B.v = B.v + mB * P;
    B.w += iB * LB;
    if (!staticA) stVelAt(c, off.z, A);
    if (!staticB) stVelAt(c, off.w, B);
}

// b2PrismaticJoint::SolveVelocityConstraints
DEVN void prismaticSolveVel(Ctx c, int jr, int modelFlags) {
    const int4 off = mi4(c, jr + MJ_SOFF);  // js jw bodyA bodyB
    const float4 im = m4(c, jr + MJ_MA);    // mA iA mB iB
    const int js = off.x, jw = off.y;
    const bool staticA = (modelFlags & 4) != 0, staticB = (modelFlags & 8) != 0;
    Vel A = staticA ? zeroVel() : ldVelAt(c, off.z);
    Vel B = staticB ? zeroVel() : ldVelAt(c, off.w);
    float mA = im.x, mB = im.z;
    float iA = im.y, iB = im.w;
    const float4 ap = ld4(c, jw + WJ_R4);    // axis.xy perp.xy
    const float4 sa = ld4(c, jw + WJ_M0);    // s1 s2 a1 a2
    const float4 kk = ld4(c, jw + WJ_M1);    // k11 k12 k13 k23
    const float4 mot = ld4(c, js + SJ_MOTOR4);  // flags maxMotorForce motorSpeed k33
    float4 imp = ld4(c, js + SJ_IMPULSE4);      // impulse.xyz motorImpulse
    const V2 axis = mk(ap.x, ap.y), perp = mk(ap.z, ap.w);
    const float s1 = sa.x, s2 = sa.y, a1 = sa.z, a2 = sa.w;
    const int flags = __float_as_int(mot.x);
    const bool enableMotor = (flags & 4) != 0;
    const int limitState = flags & 3;
    const bool enableLimit = (modelFlags & 1) != 0;
    const float k33 = mot.w;
    float k22 = iA + iB;
    if (k22 == 0.0f) k22 = 1.0f;

    if (enableMotor && limitState != 3) {
        float motorMass = k33;  // m_motorMass is the same expression as k33, inverted when positive
        if (motorMass > 0.0f) motorMass = 1.0f / motorMass;
        float Cdot = dot(axis, B.v - A.v) + a2 * B.w - a1 * A.w;
        float impulse = motorMass * (mot.z - Cdot);
        float oldImpulse = imp.w;
        float maxImpulse = PRM().dt * mot.y;
        imp.w = clampT(imp.w + impulse, -maxImpulse, maxImpulse);
        impulse = imp.w - oldImpulse;
        V2 P = impulse * axis;
        float LA = impulse * a1;
        float LB = impulse * a2;
        A.v = A.v - mA * P;
        A.w -= iA * LA;
        B.v = B.v + mB * P;
        B.w += iB * LB;
    }
    V2 Cdot1;
    Cdot1.x = dot(perp, B.v - A.v) + s2 * B.w - s1 * A.w;
    Cdot1.y = B.w - A.w;
    if (enableLimit && limitState != 0) {
        float Cdot2 = dot(axis, B.v - A.v) + a2 * B.w - a1 * A.w;
        V3 f1 = {imp.x, imp.y, imp.z};
        const float m[9] = {kk.x, kk.y, kk.z, kk.y, k22, kk.w, kk.z, kk.w, k33};
        V3 rhs = {-Cdot1.x, -Cdot1.y, -Cdot2};
        V3 df = solve33(m, rhs);
        imp.x += df.x; imp.y += df.y; imp.z += df.z;
        if (limitState == 1) imp.z = fmaxT(imp.z, 0.0f);
        else if (limitState == 2) imp.z = fminT(imp.z, 0.0f);
        V2 b = -Cdot1 - (imp.z - f1.z) * mk(kk.z, kk.w);
        V2 f2r = solveSym22(kk.x, kk.y, k22, b) + mk(f1.x, f1.y);
        imp.x = f2r.x;
        imp.y = f2r.y;
        df.x = imp.x - f1.x; df.y = imp.y - f1.y; df.z = imp.z - f1.z;
        V2 P = df.x * perp + df.z * axis;
        float LA = df.x * s1 + df.y + df.z * a1;
        float LB = df.x * s2 + df.y + df.z * a2;
        A.v = A.v - mA * P;
        A.w -= iA * LA;
        B.v = B.v + mB * P;
        B.w += iB * LB;
    } else {
        V2 df = solveSym22(kk.x, kk.y, k22, -Cdot1);
        imp.x += df.x;
        imp.y += df.y;
        V2 P = df.x * perp;
        float LA = df.x * s1 + df.y;
        float LB = df.x * s2 + df.y;
        A.v = A.v - mA * P;
        A.w -= iA * LA;
        B.v = B.v + mB * P;
        B.w += iB * LB;
    }
    st4(c, js + SJ_IMPULSE4, imp);
    if (!staticA) stVelAt(c, off.z, A);
    if (!staticB) stVelAt(c, off.w, B);
}

// b2PrismaticJoint::SolvePositionConstraints
DEVN bool prismaticSolvePos(Ctx c, int jr, int modelFlags) {
    const int4 off = mi4(c, jr + MJ_SOFF);
    const float4 im = m4(c, jr + MJ_MA);
    const float4 arm = m4(c, jr + MJ_RA0X);
    Pos A = ldPosAt(c, off.z);
    Pos B = ldPosAt(c, off.w);
    float mA = im.x, mB = im.z;
    float iA = im.y, iB = im.w;
    Rot qA = rotOf(A.a), qB = rotOf(B.a);
    V2 rA = mul(qA, mk(arm.x, arm.y));
    V2 rB = mul(qB, mk(arm.z, arm.w));
    V2 d = B.c + rB - A.c - rA;
    const V2 lx = mk(mF(c, jr + MJ_AXISX), mF(c, jr + MJ_AXISY));
    V2 axis = mul(qA, lx);
    float a1 = cross(d + rA, axis);
    float a2 = cross(rB, axis);
    V2 perp = mul(qA, cross(1.0f, lx));
    float s1 = cross(d + rA, perp);
    float s2 = cross(rB, perp);
    const float4 lim = m4(c, jr + MJ_REF);  // ref lower upper -
    V3 impulse;
    V2 C1;
    C1.x = dot(perp, d);
    C1.y = B.a - A.a - lim.x;
    float linearError = absT(C1.x);
    float angularError = absT(C1.y);
    bool active = false;
    float C2 = 0.0f;
    if ((modelFlags & 1) != 0) {
        float translation = dot(axis, d);
        if (absT(lim.z - lim.y) < 2.0f * B2G_LINEAR_SLOP) {
            C2 = clampT(translation, -B2G_MAX_LINEAR_CORRECTION, B2G_MAX_LINEAR_CORRECTION);
            linearError = fmaxT(linearError, absT(translation));
            active = true;
        } else if (translation <= lim.y) {
            C2 = clampT(translation - lim.y + B2G_LINEAR_SLOP, -B2G_MAX_LINEAR_CORRECTION, 0.0f);
            linearError = fmaxT(linearError, lim.y - translation);
            active = true;
        } else if (translation >= lim.z) {
            C2 = clampT(translation - lim.z - B2G_LINEAR_SLOP, 0.0f, B2G_MAX_LINEAR_CORRECTION);
            linearError = fmaxT(linearError, translation - lim.z);
            active = true;
        }
    }
    float k11 = mA + mB + iA * s1 * s1 + iB * s2 * s2;
    float k12 = iA * s1 + iB * s2;
    float k22 = iA + iB;
    if (k22 == 0.0f) k22 = 1.0f;
    if (active) {
        float k13 = iA * s1 * a1 + iB * s2 * a2;
        float k23 = iA * a1 + iB * a2;
        float k33 = mA + mB + iA * a1 * a1 + iB * a2 * a2;
        const float m[9] = {k11, k12, k13, k12, k22, k23, k13, k23, k33};
        V3 rhs = {-C1.x, -C1.y, -C2};
        impulse = solve33(m, rhs);
    } else {
        // b2Mat22::Solve: a11 = ex.x, a12 = ey.x, a21 = ex.y, a22 = ey.y
        V2 i1 = solve22(k11, k12, k12, k22, -C1);
        impulse.x = i1.x;
        impulse.y = i1.y;
        impulse.z = 0.0f;
    }
    V2 P = impulse.x * perp + impulse.z * axis;
    float LA = impulse.x * s1 + impulse.y + impulse.z * a1;
    float LB = impulse.x * s2 + impulse.y + impulse.z * a2;
    A.c = A.c - mA * P;
    A.a -= iA * LA;
    B.c = B.c + mB * P;
    B.a += iB * LB;
    stPosAt(c, off.z, A);
    stPosAt(c, off.w, B);
    return linearError <= B2G_LINEAR_SLOP && angularError <= B2G_ANGULAR_SLOP;
}

// b2Island::Solve's joint work for an island that consists of one dynamic body and its ground-friction joint:
// InitVelocityConstraints (warm start included) and all velocity iterations of b2FrictionJoint with the body's
// velocity, the joint's effective masses and its accumulated impulses held in registers -- the same operations in
// the same order as jointInit + velIters x frictionSolveVel<true>, minus the round trips through memory.
DEVN void frictionIslandSolve(Ctx c, int j, float dtRatio, int velIters) {
    const int jr = mjoint(c, j, 0);
    const int4 off = mi4(c, jr + MJ_SOFF);    // js jw bodyA bodyB (state byte offsets)
    const float4 arm = m4(c, jr + MJ_RA0X);   // localAnchorA - localCenterA, localAnchorB - localCenterB
    const float4 im = m4(c, jr + MJ_MA);      // mA iA mB iB
    const float2 hmax = m2(c, jr + MJ_HMAXFORCE);
    const int js = off.x;
    Vel A = zeroVel();
    Vel B = ldVelAt(c, off.w);
    float4 acc4 = ld4(c, js + SJ_IMPULSE4);
    Rot qA = ldRotAt(c, off.z), qB = ldRotAt(c, off.w);
    V2 rA = mul(qA, mk(arm.x, arm.y));
    V2 rB = mul(qB, mk(arm.z, arm.w));
    float mA = im.x, mB = im.z;
    float iA = im.y, iB = im.w;
    // ---- InitVelocityConstraints
    float kxx = mA + mB + iA * rA.y * rA.y + iB * rB.y * rB.y;
    float kxy = -iA * rA.x * rA.y - iB * rB.x * rB.y;
    float kyx = kxy;
    float kyy = mA + mB + iA * rA.x * rA.x + iB * rB.x * rB.x;
    float a = kxx, b = kyx, cc = kxy, d = kyy;
    float det = a * d - b * cc;
    if (det != 0.0f) det = 1.0f / det;
    const float exx = det * d, exy = -det * cc, eyx = -det * b, eyy = det * a;
    float angularMass = iA + iB;
    if (angularMass > 0.0f) angularMass = 1.0f / angularMass;
    {
        V2 P = mk(acc4.x * dtRatio, acc4.y * dtRatio);
        float ang = acc4.z * dtRatio;
        acc4 = mk4(P.x, P.y, ang, angularMass);
        A.v = A.v - mA * P;
        A.w -= iA * (cross(rA, P) + ang);
        B.v = B.v + mB * P;
        B.w += iB * (cross(rB, P) + ang);
    }
    // a static bodyA is written back unchanged by the reference (x -= 0 * impulse); here it stays the constant +0
    A = zeroVel();
    // ---- velocity iterations.  One iteration maps (B.v, B.w, accumulated impulses) to new values and reads nothing else
    // that changes, so once an iteration returns its input bit for bit every later one does too: stop there.  A free body
    // on its friction joint gets there in a few iterations (impulses saturate or the velocity reaches zero).
    for (int it = 0; it < velIters; ++it) {
        const float pvx = B.v.x, pvy = B.v.y, pw = B.w, pax = acc4.x, pay = acc4.y, paz = acc4.z;
        {
            float Cdot = B.w - A.w;
            float impulse = -acc4.w * Cdot;
            float oldImpulse = acc4.z;
            float maxImpulse = hmax.y;
            float acc = clampT(oldImpulse + impulse, -maxImpulse, maxImpulse);
            acc4.z = acc;
            impulse = acc - oldImpulse;
            B.w += iB * impulse;
        }
        {
            V2 Cdot = B.v + cross(B.w, rB) - A.v - cross(A.w, rA);
            V2 impulse = -mk(exx * Cdot.x + eyx * Cdot.y, exy * Cdot.x + eyy * Cdot.y);
            V2 oldImpulse = mk(acc4.x, acc4.y);
            V2 acc = oldImpulse + impulse;
            float maxImpulse = hmax.x;
            if (dot(acc, acc) > maxImpulse * maxImpulse) {
                normalize(acc);
                acc = mk(acc.x * maxImpulse, acc.y * maxImpulse);
            }
            acc4.x = acc.x;
            acc4.y = acc.y;
            impulse = acc - oldImpulse;
            B.v = B.v + mB * impulse;
            B.w += iB * cross(rB, impulse);
        }
        const int diff = (__float_as_int(B.v.x) ^ __float_as_int(pvx)) | (__float_as_int(B.v.y) ^ __float_as_int(pvy)) |
                         (__float_as_int(B.w) ^ __float_as_int(pw)) | (__float_as_int(acc4.x) ^ __float_as_int(pax)) |
                         (__float_as_int(acc4.y) ^ __float_as_int(pay)) | (__float_as_int(acc4.z) ^ __float_as_int(paz));
        if (diff == 0) break;
    }
    st4(c, js + SJ_IMPULSE4, acc4);
    stVelAt(c, off.w, B);
}

// b2RevoluteJoint::SolvePositionConstraints (friction joints have none and report true)
DEV bool revoluteSolvePos(Ctx c, int jr, int modelFlags) {
    const int4 off = mi4(c, jr + MJ_SOFF);  // js jw bodyA bodyB
    const float4 im = m4(c, jr + MJ_MA);    // mA iA mB iB
    const float4 arm = m4(c, jr + MJ_RA0X);
    Pos A = ldPosAt(c, off.z);
    Pos B = ldPosAt(c, off.w);
    float mA = im.x, mB = im.z;
    float iA = im.y, iB = im.w;
    const float4 mot = ld4(c, off.x + SJ_MOTOR4);
    int limitState = __float_as_int(mot.x) & 3;
    bool enableLimit = (modelFlags & 1) != 0;
    float angularError = 0.0f;
    float positionError = 0.0f;
    bool fixedRotation = (iA + iB == 0.0f);
    if (enableLimit && limitState != 0 && fixedRotation == false) {
        float motorMass = mF(c, jr + MJ_ROTMASS);
        const float4 lim = m4(c, jr + MJ_REF);
        float lower = lim.y, upper = lim.z;
        float angle = B.a - A.a - lim.x;
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
        V2 rA = mul(qA, mk(arm.x, arm.y));
        V2 rB = mul(qB, mk(arm.z, arm.w));
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
    stPosAt(c, off.z, A);
    stPosAt(c, off.w, B);
    return positionError <= B2G_LINEAR_SLOP && angularError <= B2G_ANGULAR_SLOP;
}

// position solve of a revolute or prismatic joint, out of line: the position loop of models WITH prismatic joints
DEVN bool positionSolveOutlined(Ctx c, int jr, int type, int modelFlags) {
    return type == 1 ? revoluteSolvePos(c, jr, modelFlags) : prismaticSolvePos(c, jr, modelFlags);
}

// ================================================================================================= contacts
DEV void contactBodies(Ctx c, int k, int& bA, int& bB, int& fA, int& fB) {
    contactFixtures(c, k, fA, fB);
    bA = mI(c, mfix(c, fA, MF_BODY));
    bB = mI(c, mfix(c, fB, MF_BODY));
}

// b2ContactSolver constructor (impulse hand-over) + InitializeVelocityConstraints for slot k
DEVN void contactInit(Ctx c, int k, float dtRatio, bool warmStarting) {
    const int ck = cslot(c, k, 0), cw = cwslot(c, k, 0);
    const float4 head = ld4(c, ck + SC_HEAD4);
    const float4 idn = ld4(c, ck + SC_ID4);
    const int fA = __float_as_int(head.x) & 0xff, fB = (__float_as_int(head.x) >> 8) & 0xff;
    const float4 fmA = m4(c, mfix(c, fA, MF_BODY));  // body count - -
    const float4 fmB = m4(c, mfix(c, fB, MF_BODY));
    const int bA = __float_as_int(fmA.x), bB = __float_as_int(fmB.x);
    int man = __float_as_int(head.z);
    int type = man & 0xff;
    int pointCount = (man >> 8) & 0xff;
    const float4 bmA = m4(c, mbody(c, bA, MB_INVMASS));
    const float4 bmB = m4(c, mbody(c, bB, MB_INVMASS));
    float mA = bmA.x, mB = bmB.x;
    float iA = bmA.y, iB = bmB.y;
    Pos PA = ldPos(c, bA), PB = ldPos(c, bB);
    Vel A = ldVel(c, bA), B = ldVel(c, bB);
    // b2Island::Solve (warmStarting): positions are the bodies' sweeps, so xf.q = b2Rot(a) is the stored m_xf.q (ldRotAt).
    // b2Island::SolveTOI: the TOI position solver has moved the positions away from the bodies' transforms -> recompute.
    // xf.p = c - q * localCenter is always recomputed: after a SetTransform the stored m_xf.p is the caller's position,
    // which differs from c - q * localCenter in the last bit.
    XF xfA, xfB;
    xfA.q = warmStarting ? ldRotAt(c, bslot(c, bA, 0)) : rotOf(PA.a);
    xfB.q = warmStarting ? ldRotAt(c, bslot(c, bB, 0)) : rotOf(PB.a);
    xfA.p = PA.c - mul(xfA.q, mk(bmA.z, bmA.w));
    xfB.p = PB.c - mul(xfB.q, mk(bmB.z, bmB.w));
    const float4 lpf = ld4(c, ck + SC_LP4);  // localPoint, b2Contact::m_friction, m_restitution (mixed at creation, addPair)
    float restitution = lpf.w;
    float friction = lpf.z;
    // b2WorldManifold::Initialize
    V2 localNormal = mk(idn.z, idn.w);
    V2 localPoint = mk(lpf.x, lpf.y);
    float4 pts[2];
    pts[0] = ld4(c, ck + SC_P04);
    pts[1] = ld4(c, ck + SC_P14);
    V2 normal;
    V2 wp[2];
    const float radius = B2G_POLYGON_RADIUS;
    if (type == 1) {
        normal = mul(xfA.q, localNormal);
        V2 planePoint = mul(xfA, localPoint);
        for (int i = 0; i < pointCount; ++i) {
            V2 clipPoint = mul(xfB, mk(pts[i].x, pts[i].y));
            V2 cA = clipPoint + (radius - dot(clipPoint - planePoint, normal)) * normal;
            V2 cB = clipPoint - radius * normal;
            wp[i] = 0.5f * (cA + cB);
        }
    } else {
        normal = mul(xfB.q, localNormal);
        V2 planePoint = mul(xfB, localPoint);
        for (int i = 0; i < pointCount; ++i) {
            V2 clipPoint = mul(xfA, mk(pts[i].x, pts[i].y));
            V2 cB = clipPoint + (radius - dot(clipPoint - planePoint, normal)) * normal;
            V2 cA = clipPoint - radius * normal;
            wp[i] = 0.5f * (cA + cB);
        }
        normal = -normal;
    }
    V2 rAs[2], rBs[2];
    float nMass[2] = {0.0f, 0.0f}, tMass[2] = {0.0f, 0.0f}, bias[2] = {0.0f, 0.0f};
    float nImp[2] = {0.0f, 0.0f}, tImp[2] = {0.0f, 0.0f};
    rAs[0] = rAs[1] = rBs[0] = rBs[1] = mk(0.0f, 0.0f);
    for (int jn = 0; jn < pointCount; ++jn) {
        V2 rA = wp[jn] - PA.c;
        V2 rB = wp[jn] - PB.c;
        rAs[jn] = rA;
        rBs[jn] = rB;
        float rnA = cross(rA, normal);
        float rnB = cross(rB, normal);
        float kNormal = mA + mB + iA * rnA * rnA + iB * rnB * rnB;
        nMass[jn] = kNormal > 0.0f ? 1.0f / kNormal : 0.0f;
        V2 tangent = cross(normal, 1.0f);
        float rtA = cross(rA, tangent);
        float rtB = cross(rB, tangent);
        float kTangent = mA + mB + iA * rtA * rtA + iB * rtB * rtB;
        tMass[jn] = kTangent > 0.0f ? 1.0f / kTangent : 0.0f;
        float vRel = dot(normal, B.v + cross(B.w, rB) - A.v - cross(A.w, rA));
        if (vRel < -B2G_VELOCITY_THRESHOLD) bias[jn] = -restitution * vRel;
        if (warmStarting) {
            nImp[jn] = dtRatio * pts[jn].z;
            tImp[jn] = dtRatio * pts[jn].w;
        }
    }
    int vcCount = pointCount;
    float k11 = 0.0f, k12 = 0.0f, k22 = 0.0f, n11 = 0.0f, n12 = 0.0f, n22 = 0.0f;
    if (pointCount == 2) {
        float rn1A = cross(rAs[0], normal);
        float rn1B = cross(rBs[0], normal);
        float rn2A = cross(rAs[1], normal);
        float rn2B = cross(rBs[1], normal);
        k11 = mA + mB + iA * rn1A * rn1A + iB * rn1B * rn1B;
        k22 = mA + mB + iA * rn2A * rn2A + iB * rn2B * rn2B;
        k12 = mA + mB + iA * rn1A * rn2A + iB * rn1B * rn2B;
        const float k_maxConditionNumber = 1000.0f;
        if (k11 * k11 < k_maxConditionNumber * (k11 * k22 - k12 * k12)) {
            // K = [ex=(k11,k12), ey=(k12,k22)]; GetInverse
            float a = k11, b = k12, cc = k12, d = k22;
            float det = a * d - b * cc;
            if (det != 0.0f) det = 1.0f / det;
            n11 = det * d;     // ex.x
            n12 = -det * cc;   // ex.y ( == ey.x )
            n22 = det * a;     // ey.y
        } else {
            vcCount = 1;
        }
    }
    st4(c, cw + WC_N4, mk4(normal.x, normal.y, __int_as_float(vcCount), friction));
    st4(c, cw + WC_R0, mk4(rAs[0].x, rAs[0].y, rBs[0].x, rBs[0].y));
    st4(c, cw + WC_R1, mk4(rAs[1].x, rAs[1].y, rBs[1].x, rBs[1].y));
    st4(c, cw + WC_MASS, mk4(nMass[0], tMass[0], nMass[1], tMass[1]));
    st4(c, cw + WC_BK4, mk4(bias[0], bias[1], k11, k12));
    st4(c, cw + WC_KN4, mk4(k22, n11, n12, n22));
    st4(c, cw + WC_IMP, mk4(nImp[0], tImp[0], nImp[1], tImp[1]));
}

// b2ContactSolver::WarmStart for slot k
DEV void contactWarmStart(Ctx c, int k) {
    int bA, bB, fA, fB;
    contactBodies(c, k, bA, bB, fA, fB);
    float mA = invMass(c, bA), mB = invMass(c, bB);
    float iA = invI(c, bA), iB = invI(c, bB);
    Vel A = ldVel(c, bA), B = ldVel(c, bB);
    const int cw = cwslot(c, k, 0);
    const float4 n4 = ld4(c, cw + WC_N4);
    const float4 imp = ld4(c, cw + WC_IMP);
    V2 normal = mk(n4.x, n4.y);
    V2 tangent = cross(normal, 1.0f);
    int pointCount = __float_as_int(n4.z);
    for (int jn = 0; jn < pointCount; ++jn) {
        const float4 r = ld4(c, cw + (jn == 0 ? WC_R0 : WC_R1));
        V2 rA = mk(r.x, r.y);
        V2 rB = mk(r.z, r.w);
        V2 P = (jn == 0 ? imp.x : imp.z) * normal + (jn == 0 ? imp.y : imp.w) * tangent;
        A.w -= iA * cross(rA, P);
        A.v = A.v - mA * P;
        B.w += iB * cross(rB, P);
        B.v = B.v + mB * P;
    }
    stVel(c, bA, A);
    stVel(c, bB, B);
}

// b2ContactSolver::SolveVelocityConstraints for slot k
DEVN void contactSolveVel(Ctx c, int k) {
    int bA, bB, fA, fB;
    contactBodies(c, k, bA, bB, fA, fB);
    float mA = invMass(c, bA), mB = invMass(c, bB);
    float iA = invI(c, bA), iB = invI(c, bB);
    Vel A = ldVel(c, bA), B = ldVel(c, bB);
    const int cw = cwslot(c, k, 0);
    const float4 n4 = ld4(c, cw + WC_N4);
    const float4 r0 = ld4(c, cw + WC_R0);
    const float4 mass = ld4(c, cw + WC_MASS);
    const float4 bk = ld4(c, cw + WC_BK4);
    float4 imp = ld4(c, cw + WC_IMP);  // n0 t0 n1 t1
    V2 normal = mk(n4.x, n4.y);
    V2 tangent = cross(normal, 1.0f);
    float friction = n4.w;
    int pointCount = __float_as_int(n4.z);
    float4 r1 = mk4(0.0f, 0.0f, 0.0f, 0.0f);
    if (pointCount == 2) r1 = ld4(c, cw + WC_R1);

    for (int jn = 0; jn < pointCount; ++jn) {
        V2 rA = jn == 0 ? mk(r0.x, r0.y) : mk(r1.x, r1.y);
        V2 rB = jn == 0 ? mk(r0.z, r0.w) : mk(r1.z, r1.w);
        V2 dv = B.v + cross(B.w, rB) - A.v - cross(A.w, rA);
        float vt = dot(dv, tangent) - 0.0f;
        float lambda = (jn == 0 ? mass.y : mass.w) * (-vt);
        float maxFriction = friction * (jn == 0 ? imp.x : imp.z);
        float oldT = jn == 0 ? imp.y : imp.w;
        float newImpulse = clampT(oldT + lambda, -maxFriction, maxFriction);
        lambda = newImpulse - oldT;
        if (jn == 0) imp.y = newImpulse; else imp.w = newImpulse;
        V2 P = lambda * tangent;
        A.v = A.v - mA * P;
        A.w -= iA * cross(rA, P);
        B.v = B.v + mB * P;
        B.w += iB * cross(rB, P);
    }
    if (pointCount == 1) {
        V2 rA = mk(r0.x, r0.y);
        V2 rB = mk(r0.z, r0.w);
        V2 dv = B.v + cross(B.w, rB) - A.v - cross(A.w, rA);
        float vn = dot(dv, normal);
        float lambda = -mass.x * (vn - bk.x);
        float oldN = imp.x;
        float newImpulse = fmaxT(oldN + lambda, 0.0f);
        lambda = newImpulse - oldN;
        imp.x = newImpulse;
        V2 P = lambda * normal;
        A.v = A.v - mA * P;
        A.w -= iA * cross(rA, P);
        B.v = B.v + mB * P;
        B.w += iB * cross(rB, P);
    } else if (pointCount == 2) {
        const float4 kn = ld4(c, cw + WC_KN4);
        V2 rA1 = mk(r0.x, r0.y), rB1 = mk(r0.z, r0.w);
        V2 rA2 = mk(r1.x, r1.y), rB2 = mk(r1.z, r1.w);
        V2 a = mk(imp.x, imp.z);
        V2 dv1 = B.v + cross(B.w, rB1) - A.v - cross(A.w, rA1);
        V2 dv2 = B.v + cross(B.w, rB2) - A.v - cross(A.w, rA2);
        float vn1 = dot(dv1, normal);
        float vn2 = dot(dv2, normal);
        V2 b;
        b.x = vn1 - bk.x;
        b.y = vn2 - bk.y;
        float k11 = bk.z, k12 = bk.w, k22 = kn.x;
        float n11 = kn.y, n12 = kn.z, n22 = kn.w;
        b = b - mk(k11 * a.x + k12 * a.y, k12 * a.x + k22 * a.y);
        V2 x;
        bool solved = false;
        // case 1
        x = -mk(n11 * b.x + n12 * b.y, n12 * b.x + n22 * b.y);
        if (x.x >= 0.0f && x.y >= 0.0f) solved = true;
        if (!solved) {  // case 2
            x.x = -mass.x * b.x;
            x.y = 0.0f;
            vn2 = k12 * x.x + b.y;
            if (x.x >= 0.0f && vn2 >= 0.0f) solved = true;
        }
        if (!solved) {  // case 3
            x.x = 0.0f;
            x.y = -mass.z * b.y;
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
            imp.x = x.x;
            imp.z = x.y;
        }
    }
    st4(c, cw + WC_IMP, imp);
    stVel(c, bA, A);
    stVel(c, bB, B);
}

// b2ContactSolver::StoreImpulses for slot k
DEV void contactStoreImpulses(Ctx c, int k) {
    const int cw = cwslot(c, k, 0);
    int n = ldI(c, cw + WC_COUNT);
    const float4 imp = ld4(c, cw + WC_IMP);
    if (n > 0) stV2(c, cslot(c, k, SC_N0), mk(imp.x, imp.y));
    if (n > 1) stV2(c, cslot(c, k, SC_N1), mk(imp.z, imp.w));
}

// b2ContactSolver::SolvePositionConstraints / SolveTOIPositionConstraints for slot k.
// toiA/toiB < 0 selects the regular solver.  Returns the minimum separation seen.
DEVN float contactSolvePos(Ctx c, int k, int toiA, int toiB, float minSeparation) {
    const int ck = cslot(c, k, 0);
    const float4 head = ld4(c, ck + SC_HEAD4);
    const float4 idn = ld4(c, ck + SC_ID4);
    const int fA = __float_as_int(head.x) & 0xff, fB = (__float_as_int(head.x) >> 8) & 0xff;
    const int bA = mI(c, mfix(c, fA, MF_BODY)), bB = mI(c, mfix(c, fB, MF_BODY));
    int man = __float_as_int(head.z);
    int type = man & 0xff;
    int pointCount = (man >> 8) & 0xff;
    bool toi = toiA >= 0;
    const float4 bmA = m4(c, mbody(c, bA, MB_INVMASS));
    const float4 bmB = m4(c, mbody(c, bB, MB_INVMASS));
    float mA = bmA.x, iA = bmA.y;
    float mB = bmB.x, iB = bmB.y;
    if (toi) {
        if (!(bA == toiA || bA == toiB)) { mA = 0.0f; iA = 0.0f; }
        if (!(bB == toiA || bB == toiB)) { mB = 0.0f; iB = 0.0f; }
    }
    V2 lcA = mk(bmA.z, bmA.w), lcB = mk(bmB.z, bmB.w);
    Pos A = ldPos(c, bA), B = ldPos(c, bB);
    V2 localNormal = mk(idn.z, idn.w);
    V2 localPoint = ldV2(c, ck + SC_LPX);
    const float radius = B2G_POLYGON_RADIUS;
    for (int jn = 0; jn < pointCount; ++jn) {
        XF xfA, xfB;
        xfA.q = rotOf(A.a);
        xfB.q = rotOf(B.a);
        xfA.p = A.c - mul(xfA.q, lcA);
        xfB.p = B.c - mul(xfB.q, lcB);
        V2 lp = ldV2(c, ck + (jn == 0 ? SC_P0X : SC_P1X));
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
// b2World::Solve builds one island at a time and solves it before looking for the next seed.  Islands never share
// a dynamic body, static bodies are never modified by the solver, and solving an island cannot change which bodies
// a later DFS reaches (connectivity = joints + touching contacts, both fixed during Solve), so here ALL islands of
// the world are built first (same seeds, same DFS, same per-island order) and then solved together phase by phase
// over the concatenated constraint lists.  Per-island semantics that remain are carried by an island id per body:
// the early exit of the position iterations and the sleep decision.  Why: with one world per thread, 32 worlds
// whose islands differ (sleeping objects, merged islands) would otherwise serialise island by island; over flat
// lists every lane just walks its own list and the warp stays converged.
struct WorldLists {
    unsigned char dyn[kMaxBodies];       // dynamic bodies in island order (each belongs to exactly one island)
    unsigned char bodyIsl[kMaxBodies];   // island id per body (static bodies: the LAST island that contained them)
    unsigned char joints[kMaxJoints];    // joints, island by island, each island in DFS order
    unsigned char contacts[kMaxContacts];
    unsigned char cA[kMaxContacts], cB[kMaxContacts];  // bodies of every live contact slot
    unsigned char jStart[kMaxBodies + 1], cStart[kMaxBodies + 1];  // first joint / contact of every island
    int ndyn, nj, nc, nIslands;
};

// integrate positions with the translation / rotation caps (b2Island::Solve step 6, also SolveTOI)
// updateSleep: b2Island::Solve's sleep-time bookkeeping (m_sleepTime = 0 if the body moves faster than the sleep tolerances,
// += h otherwise) is folded in here: it reads exactly the velocity this function leaves (the position iterations that
// follow do not touch velocities) and the sleep time lives in the 4th word of the position chunk that is stored anyway.
DEV void integratePosition(Ctx c, int b, float h, bool updateSleep) {
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
    if (updateSleep) {
        const float linTolSqr = B2G_LINEAR_SLEEP_TOL * B2G_LINEAR_SLEEP_TOL;
        const float angTolSqr = B2G_ANGULAR_SLEEP_TOL * B2G_ANGULAR_SLEEP_TOL;
        P.keep = (V.w * V.w > angTolSqr || dot(V.v, V.v) > linTolSqr) ? 0.0f : P.keep + h;
    }
    stPos(c, b, P);
    stVel(c, b, V);
}


// Phase barrier of the CTA-synchronised stepping kernel.  The step is ~17.6 k distinct instructions (k_step is 300 KB of
// SASS); resident warps that drift through different phases thrash the SM's instruction caches ("no_instruction" was the top
// stall).  With SYNC every warp of the CTA walks the phases together, so the caches hold one phase.
// All threads of the CTA must reach every phaseSync: they sit outside data-dependent control flow only.
// Which of the thirteen candidate barriers are taken is a launch parameter (StepParams::syncMask, bit ID; warp-uniform
// shared-memory read).  Measured on B200, 65536 worlds, ms per 100 steps (tools/sweep_sync.sh, profiles/r2_sync_sweep.txt):
//   round-1 inputs (contact-free), cfg 2: none 32.7, the two around the velocity iterations 24.6, every second of the first
//   ten (0x2aa) 18.2, all ten 18.5;
//   round-2 inputs (contacts, TOI), cfg 2 / cfg 3 / cfg 4: none 102 / 28.0 / 271, 0x30 75.6 / 15.9 / 188, 0x2aa 44.5 / 10.95 /
//   108.2, all ten (0x3ff) 42.3 / 10.4 / 110.4, ten + end of step (0x7ff) 41.4 / 10.35 / 110.5, all thirteen 41.8 / 10.3 / 110.4.
// Default (b2g_capi.cpp): 0x7ff for small models, 0x2aa for scenes with 16 bodies or more (cfg 4).
template <bool SYNC, int ID>
DEV void phaseSync() {
    if (SYNC && ((PRM().syncMask >> ID) & 1)) __syncthreads();
}


// ---- island lists ------------------------------------------------------------------------------------------------------
// Shared by the one-thread-per-world solver (lists in local memory, WorldLists) and the sub-warp-per-world solver (lists in
// shared memory, b2g_multi.cuh): `Lists` only needs WorldLists' members.
//
// Islands by DFS in upstream order (b2World::Solve: body list newest first, contact edges before joint edges, both newest
// first); returns the mask of static bodies that joined at least one island.
template <class Lists>
DEV uint64_t buildIslands(Ctx c, Lists& L) {
    const int nb = H().nb;
    L.ndyn = 0; L.nj = 0; L.nc = 0; L.nIslands = 0;
    uint64_t bodyFlag = 0ull;      // dynamic bodies that are in some island
    uint64_t staticSeen = 0ull;    // static bodies that joined at least one island
    uint64_t jointFlag[2] = {0ull, 0ull};
    uint64_t contactFlag[2] = {0ull, 0ull};
    const int ncAll = contactCount(c);
    // contacts that can be island edges (enabled && touching), oldest first: the DFS walks this short list newest first for
    // every body it pops instead of all ncAll slots (same visiting order: the slots it skips were skipped by the flag test)
    unsigned char liveList[kMaxContacts];
    int nLive = 0;
    for (int k = 0; k < ncAll; ++k) {
        const float4 head = ld4(c, cslot(c, k, SC_HEAD4));
        int fx = __float_as_int(head.x), fl = __float_as_int(head.y);
        L.cA[k] = (unsigned char)mI(c, mfix(c, fx & 0xff, MF_BODY));
        L.cB[k] = (unsigned char)mI(c, mfix(c, (fx >> 8) & 0xff, MF_BODY));
        if ((fl & CF_ENABLED) != 0 && (fl & CF_TOUCHING) != 0) liveList[nLive++] = (unsigned char)k;
    }

    unsigned char stack[kMaxBodies];
    // seeds: awake dynamic bodies that no earlier island holds.  A body the DFS wakes is flagged in the same breath, so the mask
    // read once is as good as the live flags.
    const uint64_t awakeAtStart = ldAwake(c);
    for (int seed = nb - 1; seed >= 0; --seed) {
        if ((bodyFlag >> seed) & 1ull) continue;
        if (((awakeAtStart >> seed) & 1ull) == 0ull) continue;
        if (bodyType(c, seed) == 0) continue;
        const int isl = L.nIslands++;
        L.jStart[isl] = (unsigned char)L.nj;
        L.cStart[isl] = (unsigned char)L.nc;
        uint64_t staticFlag = 0ull;  // static bodies join every island that reaches them, once per island
        int sp = 0;
        stack[sp++] = (unsigned char)seed;
        bodyFlag |= 1ull << seed;
        while (sp > 0) {
            int b = stack[--sp];
            L.bodyIsl[b] = (unsigned char)isl;
            if (bodyType(c, b) == 0) {
                staticSeen |= 1ull << b;  // its awake flag is settled after the sleep decisions (see below)
                continue;
            }
            B2G_ASSERT(L.ndyn < kMaxBodies && sp <= kMaxBodies, "island body list / DFS stack overflow", L.ndyn, sp);
            L.dyn[L.ndyn++] = (unsigned char)b;
            wakeBody(c, b);
            // contact edges of b, newest first
            for (int t = nLive - 1; t >= 0; --t) {
                const int k = liveList[t];
                int cA = L.cA[k], cB = L.cB[k];
                if (cA != b && cB != b) continue;
                if ((contactFlag[k >> 6] >> (k & 63)) & 1ull) continue;
                B2G_ASSERT(L.nc < kMaxContacts, "island contact list overflow", L.nc, k);
                L.contacts[L.nc++] = (unsigned char)k;
                contactFlag[k >> 6] |= 1ull << (k & 63);
                int other = cA == b ? cB : cA;
                if (((bodyFlag | staticFlag) >> other) & 1ull) continue;
                stack[sp++] = (unsigned char)other;
                if (bodyType(c, other) == 0) staticFlag |= 1ull << other; else bodyFlag |= 1ull << other;
            }
            // joint edges of b, newest first
            int js = H().oJEdge + mI(c, mbody(c, b, MB_JE_START));
            int jc = mI(c, mbody(c, b, MB_JE_COUNT));
            for (int e = 0; e < jc; ++e) {
                int packed = mI(c, js + e);
                int j = packed & 0xffff;
                int other = packed >> 16;
                if ((jointFlag[j >> 6] >> (j & 63)) & 1ull) continue;
                B2G_ASSERT(L.nj < kMaxJoints, "island joint list overflow", L.nj, j);
                L.joints[L.nj++] = (unsigned char)j;
                jointFlag[j >> 6] |= 1ull << (j & 63);
                if (((bodyFlag | staticFlag) >> other) & 1ull) continue;
                stack[sp++] = (unsigned char)other;
                if (bodyType(c, other) == 0) staticFlag |= 1ull << other; else bodyFlag |= 1ull << other;
            }
        }
    }
    return staticSeen;
}

// The order of islands inside the flat lists is free (they are independent).  Upstream finds them newest body
// first, which puts the free objects before the robot; emitting them in the opposite order keeps the lanes of
// a warp aligned on the robot's joints (same model, same DFS) and leaves the part that differs from world to
// world (which objects are awake) at the tail, where every entry has the same type.
//
// "Simple" islands -- one dynamic body held by its ground-friction joint, no contacts: a free object, or a rigid
// robot that touches nothing -- leave the lists altogether: their 15 velocity iterations only ever touch one body
// and one joint, so they run in registers (frictionIslandSolve) instead of streaming the same chunks through
// L1/L2 fifteen times.  Returns the mask of those islands; their joints go to the tail of L.joints (nSimple entries
// behind L.nj).
template <class Lists>
DEV uint64_t reorderIslands(Ctx c, Lists& L, int& nSimple) {
    uint64_t simple = 0ull;
    nSimple = 0;
    unsigned char tmp[kMaxJoints];
    unsigned char tail[kMaxBodies];
    L.jStart[L.nIslands] = (unsigned char)L.nj;
    L.cStart[L.nIslands] = (unsigned char)L.nc;
    int n = 0;
    for (int isl = L.nIslands - 1; isl >= 0; --isl) {
        const int j0 = L.jStart[isl], nj = L.jStart[isl + 1] - j0, nc = L.cStart[isl + 1] - L.cStart[isl];
        if (nj == 1 && nc == 0) {
            const int4 jh = mi4(c, mjoint(c, L.joints[j0], MJ_TYPE));
            if (jh.x == 0 && (jh.w & 12) == 4) {  // friction joint, bodyA static, bodyB dynamic
                simple |= 1ull << isl;
                tail[nSimple++] = L.joints[j0];
                continue;
            }
        }
        for (int i = j0; i < j0 + nj; ++i) tmp[n++] = L.joints[i];
    }
    for (int i = 0; i < n; ++i) L.joints[i] = tmp[i];
    for (int i = 0; i < nSimple; ++i) L.joints[n + i] = tail[i];
    L.nj = n;
    n = 0;
    for (int isl = L.nIslands - 1; isl >= 0; --isl)
        for (int i = L.cStart[isl]; i < L.cStart[isl + 1]; ++i) tmp[n++] = L.contacts[i];
    for (int i = 0; i < n; ++i) L.contacts[i] = tmp[i];
    return simple;
}

// b2Island::Solve "integrate velocities" for one island body (c0 = c, a0 = a, alpha0 = 0 ride along)
DEV void integrateVelocity(Ctx c, int b, float h) {
    const int bo = bslot(c, b, 0);
    const float4 bm = m4(c, mbody(c, b, MB_INVMASS));
    Pos P = ldPosAt(c, bo);
    st4(c, bo + SB_POS04, mk4(P.c.x, P.c.y, P.a, 0.0f));  // c0 = c, a0 = a, alpha0 = 0
    Vel V = ldVelAt(c, bo);
    // only a robot's base link ever receives forces; for every other body the chunk is the constant +0
    const float4 ft = mI(c, mbody(c, b, MB_FORCED)) ? ld4(c, bo + SB_FORCE4) : mk4(0.0f, 0.0f, 0.0f, 0.0f);
    V2 f = mk(ft.x, ft.y);
    float tq = ft.z;
    float im = bm.x, ii = bm.y;
    // gravity is (0,0), gravityScale 1, damping 0 (Box2DWorld.cpp:3668; b2BodyDef defaults)
    V2 acc = 1.0f * mk(0.0f, 0.0f) + im * f;
    V.v = V.v + h * acc;
    V.w += h * ii * tq;
    float ld = 1.0f / (1.0f + h * 0.0f);
    V.v = mk(V.v.x * ld, V.v.y * ld);
    V.w *= ld;
    stVelAt(c, bo, V);
}

// SynchronizeTransform of an island body fused with its SynchronizeFixtures; returns the bits for the move buffer and the
// body's sleep time.  xf1 = the transform at the start of the step: its rotation b2Rot(a0) is the m_xf.q still stored at this
// point, xf1.p = c0 - q * localCenter is recomputed (see contactInit).
DEV uint64_t syncIslandBody(Ctx c, int b, float& sleepTime) {
    const int bo = bslot(c, b, 0);
    const V2 lc = localCenter(c, b);
    XF xf1, xf2;
    xf1.q = ldRotAt(c, bo);
    xf1.p = ldV2(c, bo + SB_C0X) - mul(xf1.q, lc);
    const float4 p = ld4(c, bo + SB_POS4);  // c, a, sleep time
    xf2.q = rotOf(p.z);                     // b2Body::SynchronizeTransform
    xf2.p = mk(p.x, p.y) - mul(xf2.q, lc);
    stXF(c, b, xf2);
    sleepTime = p.w;
    return syncBodyFixtures(c, b, xf1, xf2);
}

// b2World::Solve + b2Island::Solve for every island of the world, then SynchronizeFixtures + FindNewContacts
template <bool SYNC>
DEVN void solveIslands(Ctx c, float dtRatio, bool allowSleep) {
    const float h = PRM().dt;
    WorldLists L;
    const uint64_t staticSeen = buildIslands(c, L);
    phaseSync<SYNC, 2>();
    int nSimple = 0;
    const uint64_t simple = reorderIslands(c, L, nSimple);
    tick(3);

    // ---- b2Island::Solve, all islands together
    for (int i = 0; i < L.ndyn; ++i) integrateVelocity(c, L.dyn[i], h);  // c0 = c, integrate velocities
    phaseSync<SYNC, 3>();
    tick(4);
    for (int i = 0; i < L.nc; ++i) contactInit(c, L.contacts[i], dtRatio, true);
    for (int i = 0; i < L.nc; ++i) contactWarmStart(c, L.contacts[i]);
    phaseSync<SYNC, 11>();
    tick(5);
    const bool prismatic = H().nPrismatic != 0;   // warp-uniform: the model is shared
    if (!prismatic) {
        for (int i = 0; i < L.nj; ++i) jointInit(c, L.joints[i], dtRatio);
    } else {
        for (int i = 0; i < L.nj; ++i) {
            const int j = L.joints[i];
            if (mI(c, mjoint(c, j, MJ_TYPE)) == 2) prismaticInit(c, j, dtRatio); else jointInit(c, j, dtRatio);
        }
    }
    tick(6);

    const int velIters = PRM().velIters, posIters = PRM().posIters;
    phaseSync<SYNC, 4>();
    if (!prismatic) {
        for (int it = 0; it < velIters; ++it) {
            for (int i = 0; i < L.nj; ++i) {
                const int jr = mjoint(c, L.joints[i], 0);
                const int4 jh = mi4(c, jr + MJ_TYPE);  // type bodyA bodyB flags
                if (jh.x == 0) {
                    frictionSolveVel(c, jr);
                } else {
                    revoluteSolveVel(c, jr, jh.w);
                }
            }
            for (int i = 0; i < L.nc; ++i) contactSolveVel(c, L.contacts[i]);
        }
    } else {
        for (int it = 0; it < velIters; ++it) {
            for (int i = 0; i < L.nj; ++i) {
                const int jr = mjoint(c, L.joints[i], 0);
                const int4 jh = mi4(c, jr + MJ_TYPE);  // type bodyA bodyB flags
                if (jh.x == 2) prismaticSolveVel(c, jr, jh.w);
                else if (jh.x == 0) frictionSolveVelOutlined(c, jr);
                else revoluteSolveVelOutlined(c, jr, jh.w);
            }
            for (int i = 0; i < L.nc; ++i) contactSolveVel(c, L.contacts[i]);
        }
    }
    phaseSync<SYNC, 5>();
    tick(7);
    for (int i = 0; i < L.nc; ++i) contactStoreImpulses(c, L.contacts[i]);
    for (int i = 0; i < nSimple; ++i) frictionIslandSolve(c, L.joints[L.nj + i], dtRatio, velIters);

    phaseSync<SYNC, 6>();
    tick(8);
    for (int i = 0; i < L.ndyn; ++i) integratePosition(c, L.dyn[i], h, allowSleep);
    phaseSync<SYNC, 12>();
    tick(9);

    // position iterations: every island leaves the loop on its own (b2Island::Solve breaks when contactsOkay &&
    // jointsOkay); `open` = islands still iterating, `solved` = islands that reported positionSolved
    const uint64_t allIslands = L.nIslands >= 64 ? ~0ull : ((1ull << L.nIslands) - 1ull);
    // simple islands have neither contacts nor position constraints: solved in their first (empty) iteration
    uint64_t open = posIters > 0 ? allIslands & ~simple : 0ull, solved = posIters > 0 ? simple : 0ull;
    for (int it = 0; it < posIters && open != 0ull; ++it) {
        uint64_t bad = 0ull;  // islands whose contacts or joints are not okay in this iteration
        for (int i = 0; i < L.nc; ++i) {
            const int k = L.contacts[i];
            // a contact's island is the island of its dynamic body (both, if both are dynamic)
            const int a = L.cA[k];
            const int isl = L.bodyIsl[bodyType(c, a) != 0 ? a : L.cB[k]];
            if (((open >> isl) & 1ull) == 0ull) continue;
            // minSeparation of an island is the minimum over its contacts: okay iff every contact's own minimum is
            float sep = contactSolvePos(c, k, -1, -1, 0.0f);
            if (!(sep >= -3.0f * B2G_LINEAR_SLOP)) bad |= 1ull << isl;
        }
        for (int i = 0; i < L.nj; ++i) {
            const int jr = mjoint(c, L.joints[i], 0);
            const int4 jh = mi4(c, jr + MJ_TYPE);
            if (jh.x == 0) continue;
            const int isl = L.bodyIsl[(jh.w & 4) ? jh.z : jh.y];
            if (((open >> isl) & 1ull) == 0ull) continue;
            bool ok;
            if (!prismatic) ok = revoluteSolvePos(c, jr, jh.w);
            else ok = positionSolveOutlined(c, jr, jh.x, jh.w);
            if (!ok) bad |= 1ull << isl;
        }
        solved |= open & ~bad;
        open &= bad;
    }
    phaseSync<SYNC, 7>();
    tick(10);
    // SynchronizeTransform of every island body, fused with its SynchronizeFixtures (upstream runs that in a second pass
    // over the body list; the order is free here: the move buffer is a bit mask and nothing in between touches transforms
    // or fat AABBs).  xf1 = the transform at the start of the step: its rotation b2Rot(a0) is the m_xf.q still stored at
    // this point, xf1.p = c0 - q * localCenter is recomputed (see contactInit).
    uint64_t moved = 0ull;
    float minSleep[kMaxBodies];  // per island: the smallest sleep time of its bodies (updated in integratePosition)
    for (int i = 0; i < L.nIslands; ++i) minSleep[i] = FLT_MAX;
    for (int i = 0; i < L.ndyn; ++i) {
        const int b = L.dyn[i];
        float sleepTime;
        moved |= syncIslandBody(c, b, sleepTime);
        const int isl = L.bodyIsl[b];
        minSleep[isl] = fminT(minSleep[isl], sleepTime);
    }
    for (uint64_t m = staticSeen; m; m &= m - 1ull) syncTransform(c, __ffsll((long long)m) - 1);

    // sleep decision per island; static bodies take the state their last island leaves them in (the DFS wakes
    // every body it reaches, a sleeping island then puts all of its bodies to sleep, static ones included)
    uint64_t asleep = 0ull;
    if (allowSleep) {
        for (int i = 0; i < L.nIslands; ++i)
            if (minSleep[i] >= B2G_TIME_TO_SLEEP && ((solved >> i) & 1ull)) asleep |= 1ull << i;
        if (asleep)
            for (int i = 0; i < L.ndyn; ++i)
                if ((asleep >> L.bodyIsl[L.dyn[i]]) & 1ull) sleepBody(c, L.dyn[i]);
    }
    for (uint64_t m = staticSeen; m; m &= m - 1ull) {
        const int b = __ffsll((long long)m) - 1;
        if ((asleep >> L.bodyIsl[b]) & 1ull) sleepBody(c, b); else wakeBody(c, b);
    }

    phaseSync<SYNC, 8>();
    tick(11);
    if (moved) stMoved(c, ldMoved(c) | moved);
    findNewContacts(c);
}

}  // namespace b2g
