// Per-world wrapper semantics on the device: DOF <-> body kinematics, the robot velocity controller,
// effort application and the stepPhysics driver.  Each routine cites the reference lines it replaces
// (/root/reference/src/sim_env/Box2DWorld.cpp, Box2DController.cpp).
#pragma once
#include "b2g_solver.cuh"
#include "b2g_toi.cuh"

namespace b2g {

DEV int olink(Ctx c, int o, int pos, int f) { return H().oLink + (mI(c, mobj(c, o, MO_LINK_START)) + pos) * ML_STRIDE + f; }
DEV float dofLimit(Ctx c, int dof, int k) { return mF(c, H().oDof + dof * 6 + k); }
DEV float clampStd(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }  // sim_env::utils::math::clamp

// Box2DLink::getPose (:215-224) for the base link of object o
DEV void basePose(Ctx c, int o, float pose[3]) {
    int b = mI(c, mobj(c, o, MO_BASE_BODY));
    XF xf = ldXF(c, b);
    V2 pos = mul(xf, mk(mF(c, mobj(c, o, MO_ORIGIN_X)), mF(c, mobj(c, o, MO_ORIGIN_Y))));
    float inv = H().invScale;
    pose[0] = inv * pos.x;
    pose[1] = inv * pos.y;
    pose[2] = ldF(c, bslot(c, b, SB_A));
}

// b2RevoluteJoint::GetJointAngle / GetJointSpeed (:972, :1037)
DEV float jointAngle(Ctx c, int j) {
    int bA = mI(c, mjoint(c, j, MJ_BODYA)), bB = mI(c, mjoint(c, j, MJ_BODYB));
    return ldF(c, bslot(c, bB, SB_A)) - ldF(c, bslot(c, bA, SB_A)) - mF(c, mjoint(c, j, MJ_REF));
}
DEV float jointSpeed(Ctx c, int j) {
    int bA = mI(c, mjoint(c, j, MJ_BODYA)), bB = mI(c, mjoint(c, j, MJ_BODYB));
    return ldF(c, bslot(c, bB, SB_W)) - ldF(c, bslot(c, bA, SB_W));
}
// b2PrismaticJoint::GetJointTranslation / GetJointSpeed (:974-977, :1039-1042): reported in Box2D units, like the reference
DEVN float prismaticTranslation(Ctx c, int j) {  // out of line: rare, keeps the DOF getters of the controller path small
    const int jr = mjoint(c, j, 0);
    int bA = mI(c, jr + MJ_BODYA), bB = mI(c, jr + MJ_BODYB);
    V2 pA = mul(ldXF(c, bA), mk(mF(c, jr + MJ_LAX), mF(c, jr + MJ_LAY)));
    V2 pB = mul(ldXF(c, bB), mk(mF(c, jr + MJ_LBX), mF(c, jr + MJ_LBY)));
    V2 d = pB - pA;
    V2 axis = mul(ldXF(c, bA).q, mk(mF(c, jr + MJ_AXISX), mF(c, jr + MJ_AXISY)));
    return dot(d, axis);
}
DEVN float prismaticSpeed(Ctx c, int j) {
    const int jr = mjoint(c, j, 0);
    int bA = mI(c, jr + MJ_BODYA), bB = mI(c, jr + MJ_BODYB);
    const float4 arm = m4(c, jr + MJ_RA0X);
    const Rot qA = ldXF(c, bA).q, qB = ldXF(c, bB).q;
    V2 rA = mul(qA, mk(arm.x, arm.y));
    V2 rB = mul(qB, mk(arm.z, arm.w));
    V2 p1 = ldV2(c, bslot(c, bA, SB_CX)) + rA;
    V2 p2 = ldV2(c, bslot(c, bB, SB_CX)) + rB;
    V2 d = p2 - p1;
    V2 axis = mul(qA, mk(mF(c, jr + MJ_AXISX), mF(c, jr + MJ_AXISY)));
    V2 vA = ldV2(c, bslot(c, bA, SB_VX)), vB = ldV2(c, bslot(c, bB, SB_VX));
    float wA = ldF(c, bslot(c, bA, SB_W)), wB = ldF(c, bslot(c, bB, SB_W));
    return dot(d, cross(wA, axis)) + dot(axis, vB + cross(wB, rB) - vA - cross(wA, rA));
}
// Box2DJoint::getPosition / getVelocity (:965-981, :1030-1046)
DEV float jointPosition(Ctx c, int j) { return H().nPrismatic != 0 && mI(c, mjoint(c, j, MJ_TYPE)) == 2 ? prismaticTranslation(c, j) : jointAngle(c, j); }
DEV float jointVelocity(Ctx c, int j) { return H().nPrismatic != 0 && mI(c, mjoint(c, j, MJ_TYPE)) == 2 ? prismaticSpeed(c, j) : jointSpeed(c, j); }
DEV int objJoint(Ctx c, int o, int i) { return mI(c, H().oObjJoint + mI(c, mobj(c, o, MO_JOINT_START)) + i) & 0xffff; }
DEV int objJointChildPos(Ctx c, int o, int i) { return mI(c, H().oObjJoint + mI(c, mobj(c, o, MO_JOINT_START)) + i) >> 16; }
DEV V2 jointAnchorA(Ctx c, int j) {  // b2RevoluteJoint::GetAnchorA
    int bA = mI(c, mjoint(c, j, MJ_BODYA));
    return mul(ldXF(c, bA), mk(mF(c, mjoint(c, j, MJ_LAX)), mF(c, mjoint(c, j, MJ_LAY))));
}

// Box2DObject::getDOFPositions / getDOFVelocities over all DOFs (:1507-1528, :1628-1649)
DEV void getDofPositions(Ctx c, int o, float* q) {
    int isStatic = mI(c, mobj(c, o, MO_STATIC));
    int n = 0;
    if (!isStatic) {
        float pose[3];
        basePose(c, o, pose);
        q[0] = pose[0]; q[1] = pose[1]; q[2] = pose[2];
        n = 3;
    }
    int nj = mI(c, mobj(c, o, MO_JOINT_COUNT));
    for (int i = 0; i < nj; ++i) q[n + i] = jointPosition(c, objJoint(c, o, i));
}
DEV void getDofVelocities(Ctx c, int o, float* qd) {
    int isStatic = mI(c, mobj(c, o, MO_STATIC));
    int n = 0;
    if (!isStatic) {
        int b = mI(c, mobj(c, o, MO_BASE_BODY));
        float inv = H().invScale;
        qd[0] = inv * ldF(c, bslot(c, b, SB_VX));
        qd[1] = inv * ldF(c, bslot(c, b, SB_VY));
        qd[2] = ldF(c, bslot(c, b, SB_W));
        n = 3;
    }
    int nj = mI(c, mobj(c, o, MO_JOINT_COUNT));
    for (int i = 0; i < nj; ++i) qd[n + i] = jointVelocity(c, objJoint(c, o, i));
}

// Box2DLink::setPose(pose, update_children = true, joint_override) (:405-438) together with the
// Box2DJoint::resetPosition recursion (:992-1028), flattened: every joint value of the subtree is read
// before anything moves (that is what the recursion does), then links are placed parent first.
// `pos` is the position of the subtree root in the object's pre-order link list.
DEVN void placeSubtree(Ctx c, int o, int pos, const float pose[3], bool jointOverride) {
    int count = mI(c, olink(c, o, pos, ML_SUBTREE));
    float jv[kMaxLinksPerObject];
    for (int i = 1; i < count; ++i) {
        int j = mI(c, olink(c, o, pos + i, ML_PARENT_JOINT));
        jv[i] = jointOverride ? 0.0f : jointPosition(c, j);
    }
    float s = H().scale, inv = H().invScale;
    int skipUntil = 0;  // links behind a prismatic joint stay where they are: Box2DJoint::resetPosition is a no-op there (:1021-1026)
    for (int i = 0; i < count; ++i) {
        int b = mI(c, olink(c, o, pos + i, ML_BODY));
        float px, py, th;
        if (i < skipUntil) continue;
        if (i == 0) {
            px = pose[0]; py = pose[1]; th = pose[2];
        } else {
            int j = mI(c, olink(c, o, pos + i, ML_PARENT_JOINT));
            if (mI(c, mjoint(c, j, MJ_TYPE)) == 2) {
                skipUntil = i + mI(c, olink(c, o, pos + i, ML_SUBTREE));
                continue;
            }
            int bA = mI(c, mjoint(c, j, MJ_BODYA));
            int dof = mI(c, mjoint(c, j, MJ_DOF));
            float clamped = clampStd(jv[i], dofLimit(c, dof, 0), dofLimit(c, dof, 1));
            th = ldF(c, bslot(c, bA, SB_A)) + clamped + mF(c, mjoint(c, j, MJ_REF));
            V2 axisWorld = jointAnchorA(c, j);
            px = inv * axisWorld.x;
            py = inv * axisWorld.y;
        }
        // only the base link of an object carries a non-zero origin offset (:653-671)
        V2 off = mk(0.0f, 0.0f);
        if (mI(c, olink(c, o, pos + i, ML_PARENT_POS)) < 0) off = mk(mF(c, mobj(c, o, MO_ORIGIN_X)), mF(c, mobj(c, o, MO_ORIGIN_Y)));
        V2 baseTranslation = mk(s * px, s * py);
        Rot r = rotOf(th);
        V2 offset = mul(r, off);
        setTransform(c, b, baseTranslation - offset, th);
    }
}

// Box2DObject::updateBodyVelocities + Box2DLink::updateBodyVelocities (:2018-2041, :759-794).
// qd holds all DOF velocities of the object (object-local indexing).
DEVN void updateBodyVelocities(Ctx c, int o, const float* qd) {
    int isStatic = mI(c, mobj(c, o, MO_STATIC));
    int nl = mI(c, mobj(c, o, MO_LINK_COUNT));
    float s = H().scale;
    float pose[3];
    basePose(c, o, pose);
    V2 baseAxis = mk(s * pose[0], s * pose[1]);
    float bvx = isStatic ? 0.0f : qd[0];
    float bvy = isStatic ? 0.0f : qd[1];
    float bw = isStatic ? 0.0f : qd[2];
    int dofOffset = mI(c, mobj(c, o, MO_DOF_OFFSET));
    for (int p = 0; p < nl; ++p) {
        int b = mI(c, olink(c, o, p, ML_BODY));
        V2 lin = mk(s * bvx, s * bvy);
        float ang = 0.0f;
        V2 com = ldV2(c, bslot(c, b, SB_CX));
        {
            V2 r = com - baseAxis;
            lin.x -= bw * r.y;
            lin.y += bw * r.x;
            ang += bw;
        }
        int cs = H().oChain + mI(c, olink(c, o, p, ML_CHAIN_START));
        int cl = mI(c, olink(c, o, p, ML_CHAIN_LEN));
        for (int e = 0; e < cl; ++e) {
            int j = mI(c, cs + e);
            float w = qd[mI(c, mjoint(c, j, MJ_DOF)) - dofOffset];
            V2 r = com - jointAnchorA(c, j);
            lin.x -= w * r.y;
            lin.y += w * r.x;
            ang += w;
        }
        if (bodyType(c, b) != 0) {  // b2Body::SetLinearVelocity / SetAngularVelocity
            if (dot(lin, lin) > 0.0f) wakeBody(c, b);
            stV2(c, bslot(c, b, SB_VX), lin);
            if (ang * ang > 0.0f) wakeBody(c, b);
            stF(c, bslot(c, b, SB_W), ang);
        }
    }
}

// Box2DObject::setPose(x, y, theta) / setTransform (:1940-1947, :1419-1433)
DEVN void objectSetPose(Ctx c, int o, float x, float y, float theta) {
    float qd[kMaxDofsPerObject];
    getDofVelocities(c, o, qd);
    float pose[3] = {x, y, theta};
    placeSubtree(c, o, 0, pose, false);
    updateBodyVelocities(c, o, qd);
}

// Box2DObject::setDOFPositions over all DOFs (:1585-1626)
DEVN void objectSetDofPositions(Ctx c, int o, const float* q) {
    float qd[kMaxDofsPerObject];
    getDofVelocities(c, o, qd);
    int isStatic = mI(c, mobj(c, o, MO_STATIC));
    int n = 0;
    if (!isStatic) {
        float pose[3] = {q[0], q[1], q[2]};
        placeSubtree(c, o, 0, pose, false);
        n = 3;
    }
    int nj = mI(c, mobj(c, o, MO_JOINT_COUNT));
    for (int i = 0; i < nj; ++i) {
        // Box2DJoint::resetPosition(values[i], false)
        int j = objJoint(c, o, i);
        int bA = mI(c, mjoint(c, j, MJ_BODYA));
        int dof = mI(c, mjoint(c, j, MJ_DOF));
        float clamped = clampStd(q[n + i], dofLimit(c, dof, 0), dofLimit(c, dof, 1));
        float pose[3];
        V2 axisWorld = jointAnchorA(c, j);
        pose[0] = H().invScale * axisWorld.x;
        pose[1] = H().invScale * axisWorld.y;
        pose[2] = ldF(c, bslot(c, bA, SB_A)) + clamped + mF(c, mjoint(c, j, MJ_REF));
        placeSubtree(c, o, objJointChildPos(c, o, i), pose, false);
    }
    updateBodyVelocities(c, o, qd);
}

// Box2DObject::setDOFVelocities over all DOFs (:1691-1715)
DEVN void objectSetDofVelocities(Ctx c, int o, const float* v) {
    float qd[kMaxDofsPerObject];
    int n = mI(c, mobj(c, o, MO_NDOFS));
    int dofOffset = mI(c, mobj(c, o, MO_DOF_OFFSET));
    for (int i = 0; i < n; ++i) qd[i] = clampStd(v[i], dofLimit(c, dofOffset + i, 2), dofLimit(c, dofOffset + i, 3));
    updateBodyVelocities(c, o, qd);
}

// Box2DObject::setDOFPositions(values, indices) (:1585-1626) for a subset of the object's DOFs: idx holds n object-local DOF
// indices in ascending order (the reference assumes sorted indices, :1611)
DEVN void objectSetDofPositionsIdx(Ctx c, int o, const float* values, const int* idx, int n) {
    float qd[kMaxDofsPerObject];
    getDofVelocities(c, o, qd);
    const int nBase = mI(c, mobj(c, o, MO_STATIC)) ? 0 : 3;
    int dofOffset = 0;
    float pose[3];
    basePose(c, o, pose);
    for (int i = 0; i < (nBase < n ? nBase : n); ++i) {
        const int d = idx[i];
        if (d < nBase) {
            ++dofOffset;
            pose[d] = values[i];
        } else {
            break;
        }
    }
    if (dofOffset > 0) placeSubtree(c, o, 0, pose, false);
    for (int i = dofOffset; i < n; ++i) {
        // Box2DJoint::resetPosition(values[i], false)
        const int ji = idx[i] - nBase;
        int j = objJoint(c, o, ji);
        int bA = mI(c, mjoint(c, j, MJ_BODYA));
        int dof = mI(c, mjoint(c, j, MJ_DOF));
        float clamped = clampStd(values[i], dofLimit(c, dof, 0), dofLimit(c, dof, 1));
        float jp[3];
        V2 axisWorld = jointAnchorA(c, j);
        jp[0] = H().invScale * axisWorld.x;
        jp[1] = H().invScale * axisWorld.y;
        jp[2] = ldF(c, bslot(c, bA, SB_A)) + clamped + mF(c, mjoint(c, j, MJ_REF));
        placeSubtree(c, o, objJointChildPos(c, o, ji), jp, false);
    }
    updateBodyVelocities(c, o, qd);
}

// Box2DObject::setDOFVelocities(values, indices) (:1691-1715) for a subset of the object's DOFs
DEVN void objectSetDofVelocitiesIdx(Ctx c, int o, const float* values, const int* idx, int n) {
    float qd[kMaxDofsPerObject];
    getDofVelocities(c, o, qd);
    int dofOffset = mI(c, mobj(c, o, MO_DOF_OFFSET));
    for (int i = 0; i < n; ++i)
        qd[idx[i]] = clampStd(values[i], dofLimit(c, dofOffset + idx[i], 2), dofLimit(c, dofOffset + idx[i], 3));
    updateBodyVelocities(c, o, qd);
}

// Box2DObject::setState (:1980-1989) for object o; q/qd are the object's slices of the world state vector
DEV void objectSetState(Ctx c, int o, const float* q, const float* qd) {
    if (!mI(c, mobj(c, o, MO_STATIC))) objectSetPose(c, o, q[0], q[1], q[2]);
    objectSetDofPositions(c, o, q);
    objectSetDofVelocities(c, o, qd);
}

// ---- freshly loaded world: what createWorld leaves behind (:3664-3721) ------------------------------------------
DEVN void initWorld(Ctx c) {
    const ModelHeader& h = H();
    for (int b = 0; b < h.nb; ++b) {
        XF xf;
        xf.p = mk(mF(c, mbody(c, b, MB_PX)), mF(c, mbody(c, b, MB_PY)));
        xf.q.s = 0.0f;
        xf.q.c = 1.0f;
        stXF(c, b, xf);
        V2 cc = mul(xf, localCenter(c, b));  // b2Body::ResetMassData: c = Mul(xf, localCenter)
        if (bodyType(c, b) == 0) cc = xf.p;
        stV2(c, bslot(c, b, SB_CX), cc);
        stV2(c, bslot(c, b, SB_C0X), cc);
        stF(c, bslot(c, b, SB_A), 0.0f);
        stF(c, bslot(c, b, SB_A0), 0.0f);
        stF(c, bslot(c, b, SB_VX), 0.0f); stF(c, bslot(c, b, SB_VY), 0.0f); stF(c, bslot(c, b, SB_W), 0.0f);
        stF(c, bslot(c, b, SB_SLEEP), 0.0f);
        stI(c, bslot(c, b, SB_FLAGS), 1);
        stF(c, bslot(c, b, SB_ALPHA0), 0.0f);
        stF(c, bslot(c, b, SB_FX), 0.0f); stF(c, bslot(c, b, SB_FY), 0.0f); stF(c, bslot(c, b, SB_TORQUE), 0.0f);
    }
    for (int f = 0; f < h.nf; ++f) st4(c, fslot(c, f), m4(c, mfix(c, f, MF_FAT0)));
    for (int j = 0; j < h.nj; ++j) {
        const float4 zero = mk4(0.0f, 0.0f, 0.0f, 0.0f);
        for (int k = 0; k < SJ_CHUNKS; ++k) st4(c, jslot(c, j, k * kChunkBytes), zero);
        int enableMotor = (mI(c, mjoint(c, j, MJ_FLAGS)) >> 1) & 1;
        stI(c, jslot(c, j, SJ_FLAGS), enableMotor << 2);
        // b2PrismaticJointDef::maxMotorForce (:936-937); motorSpeed stays 0: an actuated prismatic joint brakes
        if (mI(c, mjoint(c, j, MJ_TYPE)) == 2) stF(c, jslot(c, j, SJ_MAX_MOTOR_TORQUE), mF(c, mjoint(c, j, MJ_MAXFORCE)));
        for (int k = 0; k < WJ_CHUNKS; ++k) st4(c, jwslot(c, j, k * kChunkBytes), zero);
    }
    for (int k = 0; k < h.maxc; ++k) {
        const float4 zero = mk4(0.0f, 0.0f, 0.0f, 0.0f);
        for (int f = 0; f < SC_CHUNKS; ++f) st4(c, cslot(c, k, f * kChunkBytes), zero);
        for (int f = 0; f < WC_CHUNKS; ++f) st4(c, cwslot(c, k, f * kChunkBytes), zero);
    }
    stF(c, sslot(c, SS_INV_DT0), 0.0f);
    stI(c, sslot(c, SS_FLAGS), 1);  // e_newFixture
    stI(c, sslot(c, SS_NCONTACTS), 0);
    stI(c, sslot(c, SS_STATUS), 0);
    stMoved(c, h.nf >= 64 ? ~0ull : ((1ull << h.nf) - 1ull));  // every proxy is buffered at creation
    stI(c, sslot(c, SS_TARGET_AVAIL), 0);
    stI(c, sslot(c, SS_ABORT), 0);
    st4(c, sslot(c, SS_DISABLED_LO), mk4(0.0f, 0.0f, 0.0f, 0.0f));
    stAwake(c, h.nb >= 64 ? ~0ull : ((1ull << h.nb) - 1ull));  // every body is created awake (SB_FLAGS = 1 above)
    for (int t = 0; t < ((h.ntarget + 3) & ~3); ++t) stF(c, tslot(c, t), 0.0f);

    // Box2DObject ctor: _base_link->setPose((0,0,0), true, true) (:1335)
    for (int o = 0; o < h.no; ++o) {
        float zero[3] = {0.0f, 0.0f, 0.0f};
        placeSubtree(c, o, 0, zero, true);
    }
    // YAML `states` (:3684-3717)
    for (int o = 0; o < h.no; ++o) {
        int is = h.oInit + mI(c, mobj(c, o, MO_INIT_START));
        if (!mI(c, is)) continue;
        int n = mI(c, is + 1);
        float q[kMaxDofsPerObject + 3], qd[kMaxDofsPerObject + 3];
        for (int i = 0; i < n; ++i) { q[i] = mF(c, is + 2 + i); qd[i] = mF(c, is + 2 + n + i); }
        if (!mI(c, mobj(c, o, MO_STATIC))) {
            objectSetDofPositions(c, o, q);
            objectSetDofVelocities(c, o, qd);
        } else {
            objectSetPose(c, o, q[0], q[1], q[2]);
            objectSetDofPositions(c, o, q + 3);
            objectSetDofVelocities(c, o, qd + 3);
        }
    }
}

// Box2DRobotVelocityController::setTargetVelocity's scaleToLimits (Box2DController.cpp:42-63).  scaleToLimits belongs to the
// missing sim_env package; assumed: one uniform factor that brings every component inside its velocity limits.
DEV float scaleToLimits(Ctx c, int object, const float* v) {
    int n = mI(c, mobj(c, object, MO_NDOFS));
    int off = mI(c, mobj(c, object, MO_DOF_OFFSET));
    float factor = 1.0f;
    for (int i = 0; i < n; ++i) {
        float lo = dofLimit(c, off + i, 2), hi = dofLimit(c, off + i, 3);
        if (v[i] > hi && v[i] > 0.0f) factor = fminf(factor, hi / v[i]);
        if (v[i] < lo && v[i] < 0.0f) factor = fminf(factor, lo / v[i]);
    }
    return factor;
}

// ---- controller ------------------------------------------------------------------------------------------------
// Box2DRobot::control (:2284-2310) = Box2DRobotVelocityController::control (Box2DController.cpp:65-117)
// followed by Box2DRobot::commandEfforts (:2319-2379) and Box2DJoint::setControlTorque (:1198-1222).
// Position controller (b2g_batch_set_target_position): the velocity target of this control step.  ASSUMED law, see robotControl.
DEVN void positionControlTarget(Ctx c, int o, int n, int isStatic, int tOff, float* target) {
    float q[kMaxDofsPerObject];
    getDofPositions(c, o, q);
    const float kp = ldF(c, tslot(c, tOff + n));
    for (int i = 0; i < n; ++i) {
        float err = ldF(c, tslot(c, tOff + i)) - q[i];
        if (!isStatic && i == 2) err = err - 6.2831855f * rintf(err * 0.15915494f);
        target[i] = kp * err;
    }
    const float factor = scaleToLimits(c, o, target);
    if (factor < 1.0f) {
        stI(c, sslot(c, SS_STATUS), ldI(c, sslot(c, SS_STATUS)) | 4);  // result depends on the assumed scaling rule
        for (int i = 0; i < n; ++i) target[i] = factor * target[i];
    }
}

DEVN void robotControl(Ctx c, int o, int robotSlot) {
    const int avail = ldI(c, sslot(c, SS_TARGET_AVAIL));
    if (((avail >> robotSlot) & 1) == 0) return;  // "No target available."
    // bit 16 + slot: the robot's controller is a ControlCallback that returns fixed efforts (b2g_batch_set_control_efforts)
    const bool effortMode = ((avail >> (16 + robotSlot)) & 1) != 0;
    int n = mI(c, mobj(c, o, MO_NDOFS));
    int isStatic = mI(c, mobj(c, o, MO_STATIC));
    int baseDofs = isStatic ? 0 : 3;
    int dofOffset = mI(c, mobj(c, o, MO_DOF_OFFSET));
    int tOff = mI(c, mobj(c, o, MO_TARGET_OFFSET));
    float vel[kMaxDofsPerObject], effort[kMaxDofsPerObject], target[kMaxDofsPerObject];
    // bit 8 + slot: position controller (b2g_batch_set_target_position).  ASSUMED law (the reference's position controllers are
    // sim_env classes that are not in the tree): every control step v_target = kp * (q_target - q), the heading error of a
    // moving base wrapped to (-pi, pi], handed to setTargetVelocity (scaleToLimits included), then the velocity controller.
    if (((avail >> (8 + robotSlot)) & 1) != 0) {
        positionControlTarget(c, o, n, isStatic, tOff, target);   // out of line: the velocity-controlled path stays small
    } else {
        for (int i = 0; i < n; ++i) target[i] = ldF(c, tslot(c, tOff + i));
    }
    getDofVelocities(c, o, vel);
    float inv = H().invScale, s = H().scale;
    int baseBody = mI(c, mobj(c, o, MO_BASE_BODY));
    for (int i = 0; i < n; ++i) {
        if (effortMode) {
            effort[i] = ldF(c, tslot(c, tOff + i));
            continue;
        }
        float massInertia;
        if (i < baseDofs) {
            if (i < 2) {
                massInertia = mF(c, mobj(c, o, MO_MASS));
            } else {
                // Box2DObject::getInertia (:1909-1921): links in name order, parallel axis about the base COM
                V2 bc = ldV2(c, bslot(c, baseBody, SB_CX));
                float comx = inv * bc.x, comy = inv * bc.y;
                float inertia = 0.0f;
                int ns = H().oName + mI(c, mobj(c, o, MO_NAME_START));
                int nl = mI(c, mobj(c, o, MO_LINK_COUNT));
                for (int l = 0; l < nl; ++l) {
                    int b = mI(c, ns + l);
                    V2 lc = ldV2(c, bslot(c, b, SB_CX));
                    float dx = inv * lc.x - comx, dy = inv * lc.y - comy;
                    float distance = sqrtf(dx * dx + dy * dy);
                    // Box2DLink::getCOMInertia (:557-566)
                    float m = mF(c, mbody(c, b, MB_MASS));
                    V2 loc = localCenter(c, b);
                    float bodyInertia = mF(c, mbody(c, b, MB_I)) + m * dot(loc, loc);  // b2Body::GetInertia
                    float comDist = length(loc);
                    float comInertia = bodyInertia - m * comDist * comDist;
                    float linkComInertia = inv * inv * comInertia;
                    inertia += linkComInertia + distance * distance * m;
                }
                massInertia = inertia;
            }
        } else {
            // computeKinematicChainInertia (Box2DController.cpp:191-215): only the joint's own child link
            int j = objJoint(c, o, i - baseDofs);
            int child = mI(c, mjoint(c, j, MJ_BODYB));
            XF xfc = ldXF(c, child);
            V2 origin = mul(xfc, mk(0.0f, 0.0f));  // child getPose: GetWorldPoint(local origin offset = 0)
            float ax = inv * origin.x, ay = inv * origin.y;
            V2 cc = ldV2(c, bslot(c, child, SB_CX));
            float dx = inv * cc.x - ax, dy = inv * cc.y - ay;
            float distance = sqrtf(dx * dx + dy * dy);
            float m = mF(c, mbody(c, child, MB_MASS));
            V2 loc = localCenter(c, child);
            float bodyInertia = mF(c, mbody(c, child, MB_I)) + m * dot(loc, loc);
            float linkInertia = inv * inv * bodyInertia;  // Box2DLink::getInertia (:549-555)
            massInertia = 0.0f + (linkInertia + m * distance * distance);
        }
        float accel = (target[i] - vel[i]) / PRM().dt;
        accel = fminf(dofLimit(c, dofOffset + i, 5), fmaxf(accel, dofLimit(c, dofOffset + i, 4)));
        effort[i] = massInertia * accel;
    }
    // commandEfforts (:2319-2379).  The base link's force / torque accumulate in registers in the reference's order (x force,
    // y force, torque; each ApplyForceToCenter adds a full (fx, fy) vector) and go back with one store; a joint's motor
    // settings are one chunk.
    if (baseDofs > 0) {
        wakeBody(c, baseBody);  // ApplyForceToCenter / ApplyTorque(…, wake = true)
        float4 ft = ld4(c, bslot(c, baseBody, SB_FORCE4));
        ft.x = ft.x + s * effort[0];
        ft.y = ft.y + 0.0f;
        ft.x = ft.x + 0.0f;
        ft.y = ft.y + s * effort[1];
        ft.z = ft.z + s * s * effort[2];
        st4(c, bslot(c, baseBody, SB_FORCE4), ft);
    }
    for (int i = baseDofs; i < n; ++i) {
        int j = objJoint(c, o, i - baseDofs);
        wakeBody(c, mI(c, mjoint(c, j, MJ_BODYA)));
        wakeBody(c, mI(c, mjoint(c, j, MJ_BODYB)));
        // Box2DJoint::setControlTorque (:1198-1222): EnableMotor(true), maxMotorTorque = scale^2 |tau|, speed = sign(tau) FLT_MAX
        float4 mot = ld4(c, jslot(c, j, SJ_MOTOR4));
        mot.x = __int_as_float(__float_as_int(mot.x) | 4);
        mot.y = s * s * fabsf(effort[i]);
        mot.z = (effort[i] > 0.0f ? 1.0f : -1.0f) * FLT_MAX;
        st4(c, jslot(c, j, SJ_MOTOR4), mot);
    }
}

// Box2DWorld::stepPhysics body for one step (:3270-3282) = controllers + b2World::Step + ClearForces
template <bool SYNC>
DEVN void stepWorld(Ctx c, bool executeControllers, bool allowSleeping) {
    const ModelHeader& h = H();

    tick(-1);
    if (executeControllers) {
        for (int r = 0; r < h.nrobots; ++r) robotControl(c, mI(c, h.oRobotOrder + r), r);
    }
    if (!allowSleeping) {  // b2World::SetAllowSleeping(false) wakes every body
        for (int b = h.nb - 1; b >= 0; --b) wakeBody(c, b);
    }
    tick(0);
    // b2World::Step
    int flags = ldI(c, sslot(c, SS_FLAGS));
    if (flags & 1) {
        findNewContacts(c);
        stI(c, sslot(c, SS_FLAGS), flags & ~1);
    }
    float dtRatio = ldF(c, sslot(c, SS_INV_DT0)) * PRM().dt;
    phaseSync<SYNC, 0>();
    tick(1);
    collide(c);
    phaseSync<SYNC, 1>();
    tick(2);
    solveIslands<SYNC>(c, dtRatio, allowSleeping);
    phaseSync<SYNC, 9>();
    tick(12);
#ifndef B2G_EXPERIMENT_NO_TOI  /* tuning experiment only: wrong physics, shows the register cost of the TOI call chain */
    solveTOI(c);
#endif
    phaseSync<SYNC, 10>();
    tick(13);
    stF(c, sslot(c, SS_INV_DT0), PRM().inv_dt);
    for (int b = 0; b < h.nb; ++b) {  // ClearForces
        if (mI(c, mbody(c, b, MB_FORCED))) st4(c, bslot(c, b, SB_FORCE4), mk4(0.0f, 0.0f, 0.0f, 0.0f));
    }
    tick(14);
}

// One step of stepPhysics(callback, ...) (:3312-3333): "for (i < steps and not abort_prematurely)".  Kept out of stepWorld
// so that the plain path carries no extra live state across the step.
DEVN void stepWorldGuarded(Ctx c, bool executeControllers, bool allowSleeping) {
    if ((ldI(c, sslot(c, SS_ABORT)) & 1) != 0) return;
    stepWorld<false>(c, executeControllers, allowSleeping);
    stI(c, sslot(c, SS_ABORT), ldI(c, sslot(c, SS_ABORT)) + 256);  // executed steps in bits 8..
}

}  // namespace b2g
