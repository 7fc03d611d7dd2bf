// Flattened immutable model ("ModelBlob") and per-world state layout shared by the host builder and
// the CUDA kernels.  One model is shared by every world of a batch; it is staged into shared memory by
// each CTA.  All entries are 32-bit words (floats are bit-cast).
//
// What the model encodes (reference lines): body mass data and fixtures created by the Box2DLink ctor
// (src/sim_env/Box2DWorld.cpp:44-148), friction joints to the ground (:134-147), revolute joints (:899-919),
// object/robot DOF layout (:1314-1331, 2056-2066), creation order ground -> robots -> objects (:3677-3683).
#pragma once
#include <cstdint>

namespace b2g {

constexpr int kMaxBodies = 64;    // island / collide masks are 64-bit
constexpr int kMaxFixtures = 64;  // move buffer and pair masks are 64-bit
constexpr int kMaxJoints = 128;
constexpr int kMaxContacts = 128;
constexpr int kMaxLinksPerObject = 16;
constexpr int kMaxDofsPerObject = 3 + kMaxLinksPerObject;

// ---- model blob strides / fields --------------------------------------------------------------------
enum BodyField {  // per body; 16-byte groups are read with one LDS.128
    MB_INVMASS = 0, MB_INVI, MB_LCX, MB_LCY,
    MB_MASS, MB_I,
    MB_TYPE,      // 0 static, 2 dynamic (b2BodyType)
    MB_OBJECT,
    MB_JE_START,  // first joint edge (upstream list order: newest joint first)
    MB_JE_COUNT,
    MB_FIX_START, // fixtures of a body are consecutive proxies
    MB_FIX_COUNT,
    MB_COLLIDE_LO, MB_COLLIDE_HI,  // bit b: b2Body::ShouldCollide(this, b)
    MB_PX, MB_PY,  // creation position (non-zero for the ground body only)
    MB_SOFF,       // byte offset of this body's first state chunk (see the per-world state below)
    MB_FORCED,     // 1: forces / torques can reach this body (the base link of a robot, Box2DRobot::commandEfforts :2319-2379);
                   // every other body's force chunk stays +0 for ever and is neither read nor cleared
    MB_SPARE2, MB_SPARE3,
    MB_STRIDE
};
enum FixtureField {  // per fixture / broad-phase proxy (creation order == proxy id order)
    MF_BODY = 0, MF_COUNT, MF_FRICTION, MF_RESTITUTION,
    MF_PAIR_LO, MF_PAIR_HI,  // bit g: fixture g can ever form a contact with this one
    MF_SPARE0, MF_SPARE1,
    MF_FAT0,                  // fat AABB at creation (4 floats)
    MF_VERTS = MF_FAT0 + 4,   // 8 x (x, y)
    MF_NORMALS = MF_VERTS + 16,
    MF_STRIDE = MF_NORMALS + 16
};
enum JointField {
    MJ_TYPE = 0,  // 0 friction, 1 revolute, 2 prismatic
    MJ_BODYA, MJ_BODYB,
    MJ_FLAGS,     // bit0 enableLimit, bit1 enableMotor at creation, bit2 bodyA is static, bit3 bodyB is static
    MJ_LAX, MJ_LAY, MJ_LBX, MJ_LBY,
    MJ_REF, MJ_LOWER, MJ_UPPER, MJ_MAXFORCE,
    MJ_MAXTORQUE,
    MJ_DOF,       // index in the world state vector (revolute joints of objects), -1 otherwise
    MJ_OBJECT,
    MJ_AXISX,     // prismatic: b2PrismaticJoint::m_localXAxisA (normalised); .y is MJ_AXISY below
    // solver record: everything the velocity / position sweeps need, 16-byte groups, no dependent look-ups
    MJ_SOFF, MJ_WOFF, MJ_BAOFF, MJ_BBOFF,  // byte offsets into the per-world state: joint state, joint scratch, bodyA, bodyB
    MJ_RA0X, MJ_RA0Y, MJ_RB0X, MJ_RB0Y,    // localAnchorA - localCenterA, localAnchorB - localCenterB
    MJ_HMAXFORCE, MJ_HMAXTORQUE,           // dt * maxForce, dt * maxTorque: written by the kernel prologue (depends on dt)
    MJ_ROTMASS,                            // m = iA + iB; if (m > 0) m = 1 / m: b2FrictionJoint::m_angularMass / b2RevoluteJoint::m_motorMass
    MJ_AXISY,
    MJ_MA, MJ_IA, MJ_MB, MJ_IB,            // invMass / invI of bodyA and bodyB
    // 36 words, not 32: with a 128-byte stride the same field of every joint sits in the same shared-memory banks, and lanes
    // that read DIFFERENT joints (the one-body islands of a clutter scene, worlds whose lists differ) serialise 8 ways per
    // LDS.128; 36 = 4 (mod 32) spreads eight consecutive records over all 32 banks (body records: 20 words, fixtures: 44)
    MJ_PAD0, MJ_PAD1, MJ_PAD2, MJ_PAD3,
    MJ_STRIDE
};
static_assert(MJ_STRIDE == 36, "joint record stride");
enum ObjectField {
    MO_STATIC = 0, MO_ROBOT, MO_BASE_BODY, MO_NDOFS, MO_DOF_OFFSET,
    MO_LINK_START,   // first entry of this object's pre-order link list
    MO_LINK_COUNT,
    MO_JOINT_START,  // first entry of this object's joint list (YAML order): joint | childLinkPos << 16
    MO_JOINT_COUNT,
    MO_ORIGIN_X, MO_ORIGIN_Y,  // base link _local_origin_offset (Box2DWorld.cpp:653-671)
    MO_MASS,          // Box2DObject::_mass (:1303)
    MO_NAME_START,    // links in std::map (name) order, as body indices (Box2DObject::getInertia :1909-1921)
    MO_TARGET_OFFSET, // robots: offset of the controller target in the target block; -1 otherwise
    MO_INIT_START,    // YAML `states` entry: [has, n, configuration..., velocity...]
    MO_SPARE,
    MO_STRIDE
};
enum LinkField {  // pre-order (parent first) list per object
    ML_BODY = 0,
    ML_PARENT_JOINT,  // model joint index or -1 for the base link
    ML_PARENT_POS,    // position of the parent link in this object's list, -1 for the base link
    ML_SUBTREE,       // number of links in the subtree rooted here (contiguous in pre-order)
    ML_CHAIN_START,   // ancestor joints root -> this link (model joint indices)
    ML_CHAIN_LEN,
    ML_STRIDE
};

struct StepParams {
    float dt, inv_dt;  // inv_dt = dt > 0 ? 1 / dt : 0 (b2TimeStep)
    int velIters, posIters;
    // b2ContactListener::BeginContact hooks of the recording / guarded stepPhysics overloads
    // (src/sim_env/Box2DWorld.cpp:3290-3333, 2792-2818); all null / 0 for plain stepping
    int recMode;              // bit0: append every BeginContact to the world's record; bit1: guarded stepping
    int recMax;               // record capacity per world
    int* recCounts;           // [N]
    int* recI;                // [N][recMax][4] bodyA bodyB fixtureA fixtureB
    float* recF;              // [N][recMax][6] contact_point(3) contact_normal(3), both times 1/scale
    const unsigned char* allowed;  // [nb][nb]: 0 = a BeginContact between these bodies aborts the world's stepping
    long long worldOffset;    // per-world calls on a sub-range (b2g_batch_select_worlds): thread w owns world w + worldOffset
    const char* stateBase;    // start of the state array (world index of a thread = f(its state pointer), see worldIndex)
    int syncMask;             // k_step: which of its candidate phase barriers are taken (bit ID, b2g_solver.cuh: phaseSync)
    // Parameter sets (b2g_batch_set_link_property on a selection of worlds): the worlds of one 32-world tile share one variant
    // of the model blob; blob variants are blobStride words apart, tileVariant[tile] picks one.  nullptr: one model for all.
    int blobStride;
    const int* tileVariant;
    int lead;                 // per-world kernels on a selection whose first world is not tile-aligned: thread g owns world
                              // g - lead + worldOffset (lead = worldOffset % 32), so that a warp never straddles two tiles
    int pad0;
};

struct ModelHeader {
    int32_t nb, nf, nj, no, nrobots, nq, ntarget;
    int32_t words;  // blob size in 32-bit words
    float scale, invScale;
    // blob offsets (words; body / fixture / joint records start on 16-byte boundaries)
    int32_t oBody, oFix, oJoint, oJEdge, oObj, oLink, oObjJoint, oChain, oDof, oName, oRobotOrder, oInit;
    // per-world state layout, in 16-byte chunks (see below)
    int32_t maxc;       // contact slots per world
    int32_t cBody, cFix, cJoint, cJointW, cCont, cContW, cScalar, cTarget;
    int32_t stateChunks;
    int32_t nPrismatic;  // prismatic joints in the model: 0 lets the solver loops run without the third joint kind
    int32_t pad0;
};

// ---- per-world state ------------------------------------------------------------------------------------
// The state of a batch is one array of 16-byte chunks tiled by warp: state[tile][chunk][lane] with
// tile = world / 32, lane = world % 32.  A warp touching chunk k of its 32 worlds issues one coalesced 512-byte
// LDG.128 / STG.128, and all chunks of a tile are contiguous, so a thread reaches any word of its world at
//     base(tile, lane) + chunk * 512 + word * 4
// Field enumerators below are such byte offsets relative to the first chunk of their record.
// B2G_STATE_TILE (build-time experiment, default 32): worlds per tile of the state array.  32 = a warp's access to one chunk is
// one 512-byte row.  Smaller tiles (4, 8) make a row 64 / 128 bytes: dense access then touches 8 / 4 rows per chunk, sparse
// access (few worlds of a warp have a body awake) fetches less around what it needs.
#ifndef B2G_STATE_TILE
#define B2G_STATE_TILE 32
#endif
constexpr int kTileWorlds = B2G_STATE_TILE;
constexpr int kTileShift = kTileWorlds == 32 ? 5 : kTileWorlds == 16 ? 4 : kTileWorlds == 8 ? 3 : kTileWorlds == 4 ? 2 : kTileWorlds == 2 ? 1 : 0;
static_assert((1 << kTileShift) == kTileWorlds, "B2G_STATE_TILE must be a power of two <= 32");
constexpr int kWorldPad = 1024;  // the largest CTA of k_step
// Tiles of the state array of an n-world batch: n rounded up to a multiple of kWorldPad plus one more kWorldPad worlds.
// The padding worlds are valid idle copies of the loaded world: every thread of every k_step CTA (any CTA size that is a
// multiple of 32 up to kWorldPad) owns a world that exists in memory and can take part in the phase barriers.
inline long long paddedTiles(long long n) { return ((n + kWorldPad - 1) / kWorldPad + 1) * (kWorldPad / kTileWorlds); }
constexpr int kChunkBytes = 16 * kTileWorlds;  // 512
#define B2G_F(chunk, word) ((chunk) * (16 * B2G_STATE_TILE) + (word) * 4)

enum BodyState {
    SB_PX = B2G_F(0, 0), SB_PY = B2G_F(0, 1), SB_QS = B2G_F(0, 2), SB_QC = B2G_F(0, 3),    // b2Body::m_xf
    SB_CX = B2G_F(1, 0), SB_CY = B2G_F(1, 1), SB_A = B2G_F(1, 2),                           // m_sweep.c, a
    SB_SLEEP = B2G_F(1, 3),                                                                // m_sleepTime
    // m_sweep.c0, a0, alpha0: all three only live inside one step, one STG.128 writes them when the step begins
    SB_C0X = B2G_F(2, 0), SB_C0Y = B2G_F(2, 1), SB_A0 = B2G_F(2, 2), SB_ALPHA0 = B2G_F(2, 3),
    SB_VX = B2G_F(3, 0), SB_VY = B2G_F(3, 1), SB_W = B2G_F(3, 2),
    SB_FLAGS = B2G_F(3, 3),                                                                // bit0 awake
    SB_FX = B2G_F(4, 0), SB_FY = B2G_F(4, 1), SB_TORQUE = B2G_F(4, 2),
    SB_XF4 = B2G_F(0, 0), SB_POS4 = B2G_F(1, 0), SB_POS04 = B2G_F(2, 0), SB_VEL4 = B2G_F(3, 0), SB_FORCE4 = B2G_F(4, 0),
    SB_CHUNKS = 5
};
// Joint chunks are packed so that the velocity iterations -- the part of the step that sweeps the same chunks 15 times and
// whose footprint decides whether a resident warp set fits in L1 -- touch as few chunks as possible (friction joint 2,
// revolute joint 4, plus the bodies); constants of a joint (angular / motor mass, dt * limits) live in the model record.
//   friction (bodyA is the static ground, model_build.cpp):
//     SJ chunk 0  linearImpulse.xy angularImpulse | rB.x (scratch)        WJ chunk 0  rB.y  M.xx  M.xy(=M.yx)  M.yy   (M = K^-1)
//   revolute:
//     SJ chunk 0  impulse.xyz motorImpulse                                WJ chunk 0  rA.xy rB.xy
//     SJ chunk 1  flags maxMotorTorque motorSpeed | K.zy (scratch)        WJ chunk 1  K.xx K.xy(=K.yx) K.yy K.zx   (K.zz = iA + iB)
//     WJ chunk 2 (only written / read while the joint limit is active): the part of b2Mat33::Solve33 that does not depend
//     on the right-hand side -- cross(K.ey, K.ez) and 1 / det(K) -- computed once per step instead of once per iteration
//   prismatic (b2PrismaticJoint):
//     SJ chunk 0  impulse.xyz motorImpulse                                WJ chunk 0  axis.xy perp.xy
//     SJ chunk 1  flags maxMotorForce motorSpeed | k33 (scratch)          WJ chunk 1  s1 s2 a1 a2
//                                                                         WJ chunk 2  k11 k12 k13 k23   (k22 = iA + iB, or 1)
enum JointState {
    SJ_IX = B2G_F(0, 0), SJ_IY = B2G_F(0, 1), SJ_IZ = B2G_F(0, 2),  // friction: linearImpulse.xy, angularImpulse; revolute: m_impulse
    SJ_MOTOR_IMPULSE = B2G_F(0, 3),                                 // revolute only (friction: scratch)
    SJ_FLAGS = B2G_F(1, 0),                                         // bits0-1 limitState, bit2 enableMotor
    SJ_MAX_MOTOR_TORQUE = B2G_F(1, 1), SJ_MOTOR_SPEED = B2G_F(1, 2),
    SJ_IMPULSE4 = B2G_F(0, 0), SJ_MOTOR4 = B2G_F(1, 0),
    SJ_CHUNKS = 2
};
enum JointWork {  // solver scratch, valid inside one island solve
    WJ_R4 = B2G_F(0, 0),
    WJ_M0 = B2G_F(1, 0),
    WJ_M1 = B2G_F(2, 0),
    WJ_CHUNKS = 3
};
enum ContactState {
    SC_FIX = B2G_F(0, 0),       // fixtureA | fixtureB << 8
    SC_FLAGS = B2G_F(0, 1),     // bit0 touching, bit1 enabled, bit2 filter, bit4 toi valid; bits 8.. toiCount
    SC_MANIFOLD = B2G_F(0, 2),  // type | pointCount << 8
    SC_TOI = B2G_F(0, 3),
    SC_ID0 = B2G_F(1, 0), SC_ID1 = B2G_F(1, 1), SC_LNX = B2G_F(1, 2), SC_LNY = B2G_F(1, 3),
    SC_LPX = B2G_F(2, 0), SC_LPY = B2G_F(2, 1),
    SC_FRICTION = B2G_F(2, 2), SC_RESTITUTION = B2G_F(2, 3),  // b2Contact::m_friction / m_restitution, mixed at creation
    SC_P0X = B2G_F(3, 0), SC_P0Y = B2G_F(3, 1), SC_N0 = B2G_F(3, 2), SC_T0 = B2G_F(3, 3),
    SC_P1X = B2G_F(4, 0), SC_P1Y = B2G_F(4, 1), SC_N1 = B2G_F(4, 2), SC_T1 = B2G_F(4, 3),
    SC_HEAD4 = B2G_F(0, 0), SC_ID4 = B2G_F(1, 0), SC_LP4 = B2G_F(2, 0), SC_P04 = B2G_F(3, 0), SC_P14 = B2G_F(4, 0),
    SC_PT_STRIDE = B2G_F(1, 0),  // SC_P1* - SC_P0*
    SC_CHUNKS = 5
};
enum ContactWork {  // b2ContactVelocityConstraint, valid inside one island / TOI solve
    WC_NX = B2G_F(0, 0), WC_NY = B2G_F(0, 1),
    WC_COUNT = B2G_F(0, 2),     // velocity-constraint point count (may drop to 1)
    WC_FRICTION = B2G_F(0, 3),
    WC_R0 = B2G_F(1, 0),        // point 0: rA.xy rB.xy
    WC_R1 = B2G_F(2, 0),        // point 1
    WC_MASS = B2G_F(3, 0),      // normalMass0 tangentMass0 normalMass1 tangentMass1
    WC_BIAS0 = B2G_F(4, 0), WC_BIAS1 = B2G_F(4, 1), WC_K11 = B2G_F(4, 2), WC_K12 = B2G_F(4, 3),
    WC_K22 = B2G_F(5, 0), WC_NM11 = B2G_F(5, 1), WC_NM12 = B2G_F(5, 2), WC_NM22 = B2G_F(5, 3),  // normalMass = K^-1 (symmetric)
    WC_IMP = B2G_F(6, 0),       // normalImpulse0 tangentImpulse0 normalImpulse1 tangentImpulse1
    WC_N4 = B2G_F(0, 0), WC_BK4 = B2G_F(4, 0), WC_KN4 = B2G_F(5, 0),
    WC_CHUNKS = 7
};
enum ScalarState {
    SS_INV_DT0 = B2G_F(0, 0),
    SS_FLAGS = B2G_F(0, 1),      // bit0 newFixture, bit1 the TOI loop ran since the sweeps' alpha0 were last reset
    SS_NCONTACTS = B2G_F(0, 2),
    SS_STATUS = B2G_F(0, 3),     // bit0 contact capacity exceeded, bit1 non-finite (computed on request), bit2 a controller target
                                 // lay outside the velocity limits and was scaled by the ASSUMED scaleToLimits rule
    SS_MOVED_LO = B2G_F(1, 0), SS_MOVED_HI = B2G_F(1, 1),
    SS_TARGET_AVAIL = B2G_F(1, 2),  // bit r: robot r has a controller target
    SS_ABORT = B2G_F(1, 3),      // guarded stepping: bit0 a forbidden BeginContact fired, bits 8.. steps executed
    SS_DISABLED_LO = B2G_F(2, 0), SS_DISABLED_HI = B2G_F(2, 1),  // bit f: fixture f has categoryBits = 0 (Box2DLink::setEnabled(false))
    SS_AWAKE_LO = B2G_F(2, 2), SS_AWAKE_HI = B2G_F(2, 3),        // bit b: body b is awake -- mirrors bit0 of every body's SB_FLAGS (wakeBody / sleepBody keep
                                                                 // both): finding the awake bodies costs one 8-byte load instead of one 512-byte row per body
    SS_CHUNKS = 3
};

}  // namespace b2g
