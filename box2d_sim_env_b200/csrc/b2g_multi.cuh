// Sub-warp-per-world stepping: LPW lanes of a warp cooperate on ONE world (small and medium batches).
//
// The one-thread-per-world kernel (k_step) is the right mapping when the batch fills the GPU.  A batch of a few thousand
// worlds leaves most of the 592 warp schedulers with a single warp whose lanes each walk a strictly sequential chain of
// ~3 * 10^4 dependent instructions per step (profiles/r1f_*, r2a_*: 4-6 of 32 lanes busy, one issue every 4.6 cycles).
// Here a world is owned by a group of LPW lanes (north star: "each world is owned by a warp or CTA"; a sub-warp, because
// the parallelism that exists inside a test_env-sized world is 2-5 wide, not 32):
//
//   * per-body / per-joint / per-contact loops (velocity and position integration, transform + fixture synchronisation,
//     joint and contact initialisation, narrow phase) are strided over the lanes of the group;
//   * the Gauss-Seidel velocity iterations run as a DATAFLOW schedule that is exact: two constraint solves commute bit for
//     bit when they share no dynamic body, so the sequential sweep "for it: for joint: solve; for contact: solve" is a
//     partial order -- per body, the constraints that touch it, in list order, iteration after iteration.  Every dynamic
//     body carries a token that walks its constraint sequence; a constraint fires when it holds the tokens of all its
//     dynamic bodies, and up to LPW ready constraints of one kind fire per slot, one per lane.  Iterations overlap (the
//     finger tips start iteration t+1 while the base still finishes iteration t).  For the robot of test_env.yaml the 9
//     joint solves of an iteration take 4 slots (2 revolute + 2 friction) instead of 9;
//   * what is inherently sequential (the island DFS, contact list surgery, sleeping, TOI sub-stepping) runs on lane 0 of the
//     group while the others wait at the group barrier.
//
// Results are bit-identical to the one-thread-per-world path and to the oracle: every body sees the same operations in the
// same order.  Lists live in shared memory (WorldShared), state stays in the warp-tiled global array; lanes of a group
// hand data to each other through that state, ordered by __syncwarp(group mask).
//
// Reference for the order that is preserved: b2Island::Solve (Box2D 2.3.1) as reached from Box2DWorld::stepPhysics
// (/root/reference/src/sim_env/Box2DWorld.cpp:3279).
#pragma once
#include "b2g_world.cuh"

namespace b2g {

constexpr int kMaxNodes = 32;      // constraints (joints + contacts outside simple islands) one dataflow schedule handles
constexpr int kMaxLaps = 255;      // velocity iterations a lap counter can count
constexpr int kMultiManifolds = 48;  // contacts of a world whose manifolds the lanes evaluate in parallel (more: the rest sequentially)
constexpr int kSlotBytes = 1024;   // schedule table per world: kSlotBytes / LPW slots of LPW node indices

// Lists and schedule of one world, shared by the lanes of its group (shared memory).
struct WorldShared {
    int ndyn, nj, nc, nIslands;     // WorldLists
    int nSimple;
    int serial;                     // 1: the dataflow schedule's limits are exceeded, lane 0 sweeps sequentially
    int nSlots;                     // slots of the cached schedule
    int keyNj, keyNc, keyIters;     // what the cached schedule was built for (keyNj < 0: nothing cached)
    unsigned movedLo, movedHi;      // move-buffer bits gathered by the lanes (atomicOr)
    unsigned simpleLo, simpleHi;
    unsigned staticLo, staticHi;
    int minSleep[kMaxBodies];       // per island: min sleep time of its bodies, float bits (>= 0, so integer order = float order)
    unsigned char dyn[kMaxBodies], bodyIsl[kMaxBodies];
    unsigned char joints[kMaxJoints];
    unsigned char contacts[kMaxContacts], cA[kMaxContacts], cB[kMaxContacts];
    unsigned char jStart[kMaxBodies + 1], cStart[kMaxBodies + 1];
    unsigned char pad0[2];
    // Dataflow schedule: node i < nj is joint joints[i], node nj + k is contact contacts[k].  It only depends on which bodies
    // every node touches, so it is cached across steps (shared memory lives as long as the launch) and rebuilt when the
    // lists change: keyJoints / keyBodies hold the lists it was built from.
    unsigned char keyJoints[kMaxNodes];
    unsigned char keyBodies[2][kMaxNodes];   // contacts only: bodies of contact node i (at index i - nj)
    unsigned char slots[kSlotBytes];         // slot t, lane s: node index or 0xff
    // build-time scratch
    unsigned nodeW[kMaxNodes];  // bits 0-6 successor on the side-A body (node | side << 5 | wrap << 6), bit 7 side A is dynamic,
                                // bits 8-15 the same for side B, bits 16-21 / 22-27 the side-A / side-B body
    // narrow / broad phase shared by the lanes of the group: manifolds evaluated in parallel (one contact per lane) before the
    // sequential b2ContactManager::Collide pass commits them, fat-box hit masks per fixture before UpdatePairs adds the pairs
    Manifold man[kMultiManifolds];
    unsigned manValidLo, manValidHi;
    uint64_t hit[kMaxFixtures];
    unsigned char lap[kMaxBodies];   // laps (velocity iterations) the token of a dynamic body has completed
    unsigned char last[kMaxBodies];  // last node | side << 5 seen on a body, 0xff none
    unsigned char first[kMaxBodies];
};

template <int LPW>
struct Grp {
    int s;          // lane of the group, 0 .. LPW-1
    unsigned mask;  // the group's lanes inside the warp
    unsigned warp;  // all lanes of the warp that own a world (the kinds of a slot are voted warp-wide)
};

template <int LPW>
DEV void gsync(const Grp<LPW>& g) {
    __syncwarp(g.mask);
}

// n-th (0-based) set bit of m, or -1
DEV int nthBit(unsigned m, int n) {
    for (int k = 0; k < n; ++k) m &= m - 1u;
    return m ? __ffs((int)m) - 1 : -1;
}

// ---- joint initialisation split in two ------------------------------------------------------------------------------------
// jointInit = effective masses + impulse scaling (independent per joint: strided over the lanes) followed by the warm start,
// which adds to the body velocities and therefore has to keep the list order per body (jointWarmApply, sequential).
DEVN void jointInitPre(Ctx c, int j, float dtRatio) {
    const int jr = mjoint(c, j, 0);
    const int4 jh = mi4(c, jr + MJ_TYPE);     // type bodyA bodyB flags
    const int4 off = mi4(c, jr + MJ_SOFF);    // js jw bodyA bodyB (state byte offsets)
    const float4 arm = m4(c, jr + MJ_RA0X);   // localAnchorA - localCenterA, localAnchorB - localCenterB
    const float4 im = m4(c, jr + MJ_MA);      // mA iA mB iB
    const int type = jh.x;
    if (type == 2) return;  // prismatic joints are initialised as a whole by jointWarmApply (sequential phase)
    const int js = off.x, jw = off.y;
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
        float a = kxx, b = kyx, cc = kxy, d = kyy;
        float det = a * d - b * cc;
        if (det != 0.0f) det = 1.0f / det;
        st4(c, jw + WJ_R4, mk4(rB.y, det * d, -det * cc, det * a));  // rB.y  M.xx  M.xy  M.yy
        float4 imp = ld4(c, js + SJ_IMPULSE4);
        st4(c, js + SJ_IMPULSE4, mk4(imp.x * dtRatio, imp.y * dtRatio, imp.z * dtRatio, rB.x));
    } else {
        bool fixedRotation = (iA + iB == 0.0f);
        float exx = mA + mB + rA.y * rA.y * iA + rB.y * rB.y * iB;
        float eyx = -rA.y * rA.x * iA - rB.y * rB.x * iB;
        float ezx = -rA.y * iA - rB.y * iB;
        float eyy = mA + mB + rA.x * rA.x * iA + rB.x * rB.x * iB;
        float ezy = rA.x * iA + rB.x * iB;
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
            const float ezz = iA + iB;
            const float cx = eyy * ezz - ezy * ezy, cy = ezy * ezx - eyx * ezz, cz = eyx * ezy - eyy * ezx;
            float det = exx * cx + eyx * cy + ezx * cz;
            if (det != 0.0f) det = 1.0f / det;
            st4(c, jw + WJ_M1, mk4(cx, cy, cz, det));
        }
        imp.x *= dtRatio; imp.y *= dtRatio; imp.z *= dtRatio;
        motorImpulse *= dtRatio;
        st4(c, js + SJ_IMPULSE4, mk4(imp.x, imp.y, imp.z, motorImpulse));
    }
}

// warm start of joint j: the (already scaled) accumulated impulses act on the body velocities
DEV void jointWarmApply(Ctx c, int j, float dtRatio) {
    const int jr = mjoint(c, j, 0);
    const int4 jh = mi4(c, jr + MJ_TYPE);
    if (jh.x == 2) {
        prismaticInit(c, j, dtRatio);
        return;
    }
    const int4 off = mi4(c, jr + MJ_SOFF);
    const float4 im = m4(c, jr + MJ_MA);
    const bool staticA = (jh.w & 4) != 0, staticB = (jh.w & 8) != 0;
    const float4 imp = ld4(c, off.x + SJ_IMPULSE4);
    float mA = im.x, mB = im.z;
    float iA = im.y, iB = im.w;
    if (jh.x == 0) {
        // friction joint: bodyA is the static ground (model builder), only bodyB moves
        const V2 rB = mk(imp.w, ldF(c, off.y + WJ_R4));
        const V2 P = mk(imp.x, imp.y);
        Vel B = ldVelAt(c, off.w);
        B.v = B.v + mB * P;
        B.w += iB * (cross(rB, P) + imp.z);
        stVelAt(c, off.w, B);
    } else {
        const float4 r = ld4(c, off.y + WJ_R4);
        const V2 rA = mk(r.x, r.y), rB = mk(r.z, r.w);
        const V2 P = mk(imp.x, imp.y);
        if (!staticA) {
            Vel A = ldVelAt(c, off.z);
            A.v = A.v - mA * P;
            A.w -= iA * (cross(rA, P) + imp.w + imp.z);
            stVelAt(c, off.z, A);
        }
        if (!staticB) {
            Vel B = ldVelAt(c, off.w);
            B.v = B.v + mB * P;
            B.w += iB * (cross(rB, P) + imp.w + imp.z);
            stVelAt(c, off.w, B);
        }
    }
}

// ---- dataflow schedule of the velocity iterations ----------------------------------------------------------------------------
// Lane 0, once per step: is the cached schedule still the one for these lists?  If not, rebuild it by running the token rule
// symbolically: every dynamic body's token walks the body's constraint sequence (list order, iteration after iteration); a
// constraint fires in the first slot in which it holds the tokens of all its dynamic bodies; up to LPW constraints fire per
// slot.  Constraints of different kinds share a slot: the lanes of a group that take different branches are interleaved by
// the scheduler, and the slot count is what the dependent chain of a step is made of.
template <int LPW>
DEV void updateSchedule(Ctx c, WorldShared& W, int velIters) {
    const int nj = W.nj, nc = W.nc, n = nj + nc;
#ifdef B2G_MULTI_FORCE_SERIAL  /* tuning experiment: lane-strided phases only, the velocity iterations stay on lane 0 */
    if (true) {
#else
    if (n > kMaxNodes || velIters > kMaxLaps || velIters < 1) {
#endif
        W.serial = 1;
        W.keyNj = -1;
        return;
    }
    bool same = W.keyNj == nj && W.keyNc == nc && W.keyIters == velIters;
    if (same) {
        for (int i = 0; i < nj; ++i) same = same && W.keyJoints[i] == W.joints[i];
        for (int i = 0; i < nc; ++i) {
            const int k = W.contacts[i];
            same = same && W.keyBodies[0][i] == W.cA[k] && W.keyBodies[1][i] == W.cB[k];
        }
    }
    if (same) return;  // W.serial / W.nSlots / W.slots are still valid

    W.keyNj = nj; W.keyNc = nc; W.keyIters = velIters;
    for (int i = 0; i < nj; ++i) W.keyJoints[i] = W.joints[i];
    for (int i = 0; i < nc; ++i) {
        const int k = W.contacts[i];
        W.keyBodies[0][i] = W.cA[k];
        W.keyBodies[1][i] = W.cB[k];
    }
    W.serial = 0;
    for (int i = 0; i < W.ndyn; ++i) {
        W.last[W.dyn[i]] = 0xff;
        W.lap[W.dyn[i]] = 0;
    }
    unsigned tA = 0u, tB = 0u;
    for (int i = 0; i < n; ++i) {
        int body[2];
        bool dynamic[2];
        if (i < nj) {
            const int4 jh = mi4(c, mjoint(c, W.joints[i], MJ_TYPE));  // type bodyA bodyB flags
            body[0] = jh.y; body[1] = jh.z;
            dynamic[0] = (jh.w & 4) == 0; dynamic[1] = (jh.w & 8) == 0;
        } else {
            const int k = W.contacts[i - nj];
            body[0] = W.cA[k]; body[1] = W.cB[k];
            dynamic[0] = bodyType(c, body[0]) != 0; dynamic[1] = bodyType(c, body[1]) != 0;
        }
        unsigned word = 0u;
        for (int side = 0; side < 2; ++side) {
            const int b = body[side];
            if (!dynamic[side]) {  // a static body is never written: no token, the side is always ready
                if (side == 0) tA |= 1u << i; else tB |= 1u << i;
                continue;
            }
            const int here = i | (side << 5);
            const int prev = W.last[b];
            if (prev == 0xff) {
                W.first[b] = (unsigned char)here;
                if (side == 0) tA |= 1u << i; else tB |= 1u << i;  // the first constraint of the body's sequence holds its token
            } else {
                W.nodeW[prev & 31] |= (unsigned)here << (8 * ((prev >> 5) & 1));
            }
            W.last[b] = (unsigned char)here;
            word |= (0x80u << (8 * side)) | ((unsigned)b << (16 + 6 * side));
        }
        W.nodeW[i] = word;  // successors are filled in by the next constraint on each body, or by the wrap-around below
    }
    for (int i = 0; i < W.ndyn; ++i) {
        const int b = W.dyn[i];
        const int prev = W.last[b];
        if (prev == 0xff) continue;  // body of a simple island: not in the lists
        W.nodeW[prev & 31] |= (unsigned)(W.first[b] | 0x40) << (8 * ((prev >> 5) & 1));
    }
    // symbolic run.  Constraints of different kinds share a slot.  Measured alternatives (B200, test_env.yaml x 4096, ms per
    // 100 steps): kind-homogeneous slots chosen per world 20.9, kinds voted warp-wide 40, mixed slots 17.3 -- homogeneous slots
    // need 64 instead of 47 slots for the robot alone and several times more once a contact ties the fingers to an object.
    int t = 0;
    for (;;) {
        unsigned m = tA & tB;
        if (m == 0u) break;
        if ((t + 1) * LPW > kSlotBytes) {
            W.serial = 1;
            break;
        }
        for (int k = 0; k < LPW; ++k) {
            if (m == 0u) {
                W.slots[t * LPW + k] = 0xff;
                continue;
            }
            const int i = __ffs((int)m) - 1;
            m &= m - 1u;
            W.slots[t * LPW + k] = (unsigned char)i;
            const unsigned w = W.nodeW[i];
            const unsigned bit = 1u << i;
            for (int side = 0; side < 2; ++side) {
                const unsigned ws = w >> (8 * side);
                if ((ws & 0x80u) == 0u) continue;  // static side: no token
                if (side == 0) tA &= ~bit; else tB &= ~bit;
                if (ws & 0x40u) {  // end of the body's sequence: one more lap done; after the last lap the token retires
                    const int b = (w >> (16 + 6 * side)) & 63;
                    const int l = W.lap[b] + 1;
                    W.lap[b] = (unsigned char)l;
                    if (l >= velIters) continue;
                }
                if (ws & 0x20u) tB |= 1u << (ws & 31u); else tA |= 1u << (ws & 31u);
            }
        }
        ++t;
    }
    W.nSlots = t;
}

// All velocity iterations of the world's joint / contact lists: slot after slot of the cached schedule, one constraint per
// lane; the group barrier at the end of a slot orders its velocity stores before the loads of the next.
template <int LPW>
DEV void velocityIterations(Ctx c, const Grp<LPW>& g, const WorldShared& W, int velIters) {
    const int nj = W.nj;
    if (W.serial) {
        if (g.s == 0) {
            for (int it = 0; it < velIters; ++it) {
                for (int i = 0; i < nj; ++i) {
                    const int jr = mjoint(c, W.joints[i], 0);
                    const int4 jh = mi4(c, jr + MJ_TYPE);
                    if (jh.x == 0) frictionSolveVel(c, jr);
                    else if (jh.x == 1) revoluteSolveVel(c, jr, jh.w);
                    else prismaticSolveVel(c, jr, jh.w);
                }
                for (int i = 0; i < W.nc; ++i) contactSolveVel(c, W.contacts[i]);
            }
        }
        gsync(g);
        return;
    }
    const int nSlots = W.nSlots;
    for (int t = 0; t < nSlots; ++t) {
        const int i = W.slots[t * LPW + g.s];
        if (i != 0xff) {
            if (i < nj) {
                const int jr = mjoint(c, W.joints[i], 0);
                const int4 jh = mi4(c, jr + MJ_TYPE);
                if (jh.x == 0) frictionSolveVel(c, jr);
                else if (jh.x == 1) revoluteSolveVel(c, jr, jh.w);
                else prismaticSolveVel(c, jr, jh.w);
            } else {
                contactSolveVel(c, W.contacts[i - nj]);
            }
        }
        gsync(g);
    }
}

// ---- b2World::Solve for a world owned by LPW lanes ------------------------------------------------------------------------
template <int LPW>
DEV void solveIslandsMulti(Ctx c, const Grp<LPW>& g, WorldShared& W, float dtRatio, bool allowSleep) {
    const float h = PRM().dt;
    const int velIters = PRM().velIters, posIters = PRM().posIters;
    if (g.s == 0) {
        const uint64_t staticSeen = buildIslands(c, W);
        int nSimple = 0;
        const uint64_t simple = reorderIslands(c, W, nSimple);
        W.nSimple = nSimple;
        W.simpleLo = (unsigned)simple; W.simpleHi = (unsigned)(simple >> 32);
        W.staticLo = (unsigned)staticSeen; W.staticHi = (unsigned)(staticSeen >> 32);
        W.movedLo = W.movedHi = 0u;
        for (int i = 0; i < W.nIslands; ++i) W.minSleep[i] = __float_as_int(FLT_MAX);
        updateSchedule<LPW>(c, W, velIters);
    }
    gsync(g);
    const int ndyn = W.ndyn, nj = W.nj, nc = W.nc;

    // ---- b2Island::Solve, all islands together
    for (int i = g.s; i < ndyn; i += LPW) integrateVelocity(c, W.dyn[i], h);
    gsync(g);
    for (int i = g.s; i < nc; i += LPW) contactInit(c, W.contacts[i], dtRatio, true);
    for (int i = g.s; i < nj; i += LPW) jointInitPre(c, W.joints[i], dtRatio);
    gsync(g);
    if (g.s == 0) {  // warm starts add to body velocities: list order (contacts, then joints)
        for (int i = 0; i < nc; ++i) contactWarmStart(c, W.contacts[i]);
        for (int i = 0; i < nj; ++i) jointWarmApply(c, W.joints[i], dtRatio);
    }
    gsync(g);

    velocityIterations(c, g, W, velIters);

    for (int i = g.s; i < nc; i += LPW) contactStoreImpulses(c, W.contacts[i]);
    for (int i = g.s; i < W.nSimple; i += LPW) frictionIslandSolve(c, W.joints[nj + i], dtRatio, velIters);
    gsync(g);
    for (int i = g.s; i < ndyn; i += LPW) integratePosition(c, W.dyn[i], h, allowSleep);
    gsync(g);

    // position iterations (sequential Gauss-Seidel with a per-island early exit): lane 0
    const uint64_t simple = (uint64_t)W.simpleLo | ((uint64_t)W.simpleHi << 32);
    uint64_t solved = 0ull;
    if (g.s == 0) {
        const uint64_t allIslands = W.nIslands >= 64 ? ~0ull : ((1ull << W.nIslands) - 1ull);
        uint64_t open = posIters > 0 ? allIslands & ~simple : 0ull;
        solved = posIters > 0 ? simple : 0ull;
        for (int it = 0; it < posIters && open != 0ull; ++it) {
            uint64_t bad = 0ull;
            for (int i = 0; i < nc; ++i) {
                const int k = W.contacts[i];
                const int a = W.cA[k];
                const int isl = W.bodyIsl[bodyType(c, a) != 0 ? a : W.cB[k]];
                if (((open >> isl) & 1ull) == 0ull) continue;
                float sep = contactSolvePos(c, k, -1, -1, 0.0f);
                if (!(sep >= -3.0f * B2G_LINEAR_SLOP)) bad |= 1ull << isl;
            }
            for (int i = 0; i < nj; ++i) {
                const int jr = mjoint(c, W.joints[i], 0);
                const int4 jh = mi4(c, jr + MJ_TYPE);
                if (jh.x == 0) continue;
                const int isl = W.bodyIsl[(jh.w & 4) ? jh.z : jh.y];
                if (((open >> isl) & 1ull) == 0ull) continue;
                if (!(jh.x == 1 ? revoluteSolvePos(c, jr, jh.w) : prismaticSolvePos(c, jr, jh.w))) bad |= 1ull << isl;
            }
            solved |= open & ~bad;
            open &= bad;
        }
    }
    gsync(g);

    // SynchronizeTransform + SynchronizeFixtures of every island body, strided over the lanes
    {
        uint64_t moved = 0ull;
        for (int i = g.s; i < ndyn; i += LPW) {
            const int b = W.dyn[i];
            float sleepTime;
            moved |= syncIslandBody(c, b, sleepTime);
            atomicMin(&W.minSleep[W.bodyIsl[b]], __float_as_int(sleepTime));  // sleep times are >= 0: integer order = float order
        }
        if (moved) {
            if ((unsigned)moved) atomicOr(&W.movedLo, (unsigned)moved);
            if ((unsigned)(moved >> 32)) atomicOr(&W.movedHi, (unsigned)(moved >> 32));
        }
    }
    gsync(g);

    if (g.s == 0) {
        const uint64_t staticSeen = (uint64_t)W.staticLo | ((uint64_t)W.staticHi << 32);
        for (uint64_t m = staticSeen; m; m &= m - 1ull) syncTransform(c, __ffsll((long long)m) - 1);
        uint64_t asleep = 0ull;
        if (allowSleep) {
            for (int i = 0; i < W.nIslands; ++i)
                if (__int_as_float(W.minSleep[i]) >= B2G_TIME_TO_SLEEP && ((solved >> i) & 1ull)) asleep |= 1ull << i;
            if (asleep)
                for (int i = 0; i < ndyn; ++i)
                    if ((asleep >> W.bodyIsl[W.dyn[i]]) & 1ull) sleepBody(c, W.dyn[i]);
        }
        for (uint64_t m = staticSeen; m; m &= m - 1ull) {
            const int b = __ffsll((long long)m) - 1;
            if ((asleep >> W.bodyIsl[b]) & 1ull) sleepBody(c, b); else wakeBody(c, b);
        }
        const uint64_t moved = (uint64_t)W.movedLo | ((uint64_t)W.movedHi << 32);
        if (moved) stMoved(c, ldMoved(c) | moved);
    }
    gsync(g);
    // broad phase: the fat-box tests of every fixture against its candidates, fixtures strided over the lanes; lane 0 then adds
    // the pairs in ascending (a, b) order (b2BroadPhase::UpdatePairs)
    {
        const uint64_t moved = ldMoved(c);
        if (moved != 0ull) {
            const int nf = H().nf;
            for (int a = g.s; a < nf; a += LPW) W.hit[a] = newPairHits(c, a, moved);
        }
    }
    gsync(g);
    if (g.s == 0) findNewContacts(c, W.hit);
    gsync(g);
}

// Box2DWorld::stepPhysics body for one step, world owned by LPW lanes
template <int LPW>
DEV void stepWorldMulti(Ctx c, const Grp<LPW>& g, WorldShared& W, bool executeControllers, bool allowSleeping) {
    const ModelHeader& h = H();
    float dtRatio = 0.0f;
    if (g.s == 0) {
        if (executeControllers) {
            for (int r = 0; r < h.nrobots; ++r) robotControl(c, mI(c, h.oRobotOrder + r), r);
        }
        if (!allowSleeping) {
            for (int b = h.nb - 1; b >= 0; --b) wakeBody(c, b);
        }
        int flags = ldI(c, sslot(c, SS_FLAGS));
        if (flags & 1) {
            findNewContacts(c);
            stI(c, sslot(c, SS_FLAGS), flags & ~1);
        }
        W.manValidLo = W.manValidHi = 0u;
    }
    gsync(g);
    // narrow phase, one contact per lane: b2CollidePolygons of every contact the Collide pass is going to update (an awake
    // dynamic body, fat boxes still overlapping -- judged on the state before the pass; a contact the pass wants although it
    // was skipped here, because an earlier contact of the pass woke its body, is evaluated by the pass itself)
    {
        const int n = contactCount(c);
        const uint64_t awake = ldAwake(c);
        for (int k = g.s; k < n && k < kMultiManifolds; k += LPW) {
            int fA, fB;
            contactFixtures(c, k, fA, fB);
            const int bA = mI(c, mfix(c, fA, MF_BODY)), bB = mI(c, mfix(c, fB, MF_BODY));
            const bool activeA = ((awake >> bA) & 1ull) && bodyType(c, bA) != 0;
            const bool activeB = ((awake >> bB) & 1ull) && bodyType(c, bB) != 0;
            if (!activeA && !activeB) continue;
            if (!boxOverlap(ldFat(c, fA), ldFat(c, fB))) continue;
            evalContact(c, k, W.man[k]);
            if (k < 32) atomicOr(&W.manValidLo, 1u << k); else atomicOr(&W.manValidHi, 1u << (k - 32));
        }
    }
    gsync(g);
    if (g.s == 0) collide(c, W.man, (uint64_t)W.manValidLo | ((uint64_t)W.manValidHi << 32));
    dtRatio = ldF(c, sslot(c, SS_INV_DT0)) * PRM().dt;  // written at the end of the previous step, before a group barrier
    gsync(g);
    solveIslandsMulti<LPW>(c, g, W, dtRatio, allowSleeping);
    if (g.s == 0) {
        solveTOI(c);
        stF(c, sslot(c, SS_INV_DT0), PRM().inv_dt);
        for (int b = 0; b < h.nb; ++b) {  // ClearForces
            if (mI(c, mbody(c, b, MB_FORCED))) st4(c, bslot(c, b, SB_FORCE4), mk4(0.0f, 0.0f, 0.0f, 0.0f));
        }
    }
    gsync(g);
}

}  // namespace b2g
