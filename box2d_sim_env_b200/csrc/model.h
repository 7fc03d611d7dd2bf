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
enum BodyField {  // per body
    MB_INVMASS = 0, MB_INVI, MB_MASS, MB_I, MB_LCX, MB_LCY,
    MB_TYPE,      // 0 static, 2 dynamic (b2BodyType)
    MB_JE_START,  // first joint edge (upstream list order: newest joint first)
    MB_JE_COUNT,
    MB_FIX_START, // fixtures of a body are consecutive proxies
    MB_FIX_COUNT,
    MB_OBJECT,
    MB_COLLIDE_LO, MB_COLLIDE_HI,  // bit b: b2Body::ShouldCollide(this, b)
    MB_PX, MB_PY,  // creation position (non-zero for the ground body only)
    MB_STRIDE
};
enum FixtureField {  // per fixture / broad-phase proxy (creation order == proxy id order)
    MF_BODY = 0, MF_COUNT, MF_FRICTION, MF_RESTITUTION,
    MF_PAIR_LO, MF_PAIR_HI,  // bit g: fixture g can ever form a contact with this one
    MF_FAT0,                  // fat AABB at creation (4 floats)
    MF_VERTS = MF_FAT0 + 4,   // 8 x (x, y)
    MF_NORMALS = MF_VERTS + 16,
    MF_STRIDE = MF_NORMALS + 16
};
enum JointField {
    MJ_TYPE = 0,  // 0 friction, 1 revolute
    MJ_BODYA, MJ_BODYB,
    MJ_FLAGS,     // bit0 enableLimit, bit1 enableMotor at creation
    MJ_LAX, MJ_LAY, MJ_LBX, MJ_LBY, MJ_REF, MJ_LOWER, MJ_UPPER, MJ_MAXFORCE, MJ_MAXTORQUE,
    MJ_DOF,       // index in the world state vector (revolute joints of objects), -1 otherwise
    MJ_OBJECT,
    MJ_SPARE,
    MJ_STRIDE
};
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

struct ModelHeader {
    int32_t nb, nf, nj, no, nrobots, nq, ntarget;
    int32_t words;  // blob size in 32-bit words
    float scale, invScale;
    // blob offsets
    int32_t oBody, oFix, oJoint, oJEdge, oObj, oLink, oObjJoint, oChain, oDof, oName, oRobotOrder, oInit;
    // per-world state layout (word slots; state is stored [slot][world])
    int32_t maxc;       // contact slots per world
    int32_t sBody, sFix, sJoint, sJointW, sCont, sContW, sScalar, sTarget;
    int32_t stateWords;
};

// ---- per-world state strides / fields -----------------------------------------------------------------
enum BodyState {
    SB_PX = 0, SB_PY, SB_QS, SB_QC,  // b2Body::m_xf
    SB_CX, SB_CY, SB_A,              // m_sweep.c, a
    SB_C0X, SB_C0Y, SB_A0,           // m_sweep.c0, a0
    SB_VX, SB_VY, SB_W,
    SB_SLEEP,                         // m_sleepTime
    SB_FLAGS,                         // bit0 awake
    SB_ALPHA0,                        // m_sweep.alpha0
    SB_FX, SB_FY, SB_TORQUE,
    SB_STRIDE
};
enum JointState {
    SJ_IX = 0, SJ_IY, SJ_IZ,  // friction: linearImpulse.xy, angularImpulse; revolute: m_impulse
    SJ_MOTOR_IMPULSE,
    SJ_FLAGS,                  // bits0-1 limitState, bit2 enableMotor
    SJ_MAX_MOTOR_TORQUE, SJ_MOTOR_SPEED,
    SJ_SPARE,
    SJ_STRIDE
};
enum JointWork {  // solver scratch, valid inside one island solve
    WJ_RAX = 0, WJ_RAY, WJ_RBX, WJ_RBY,
    WJ_M0,  // friction: linearMass ex.x ex.y ey.x ey.y, angularMass; revolute: mass ex.xyz ey.xyz ez.xyz, motorMass
    WJ_STRIDE = WJ_M0 + 12
};
enum ContactState {
    SC_FIX = 0,    // fixtureA | fixtureB << 8
    SC_FLAGS,      // bit0 touching, bit1 enabled, bit2 filter, bit4 toi valid; bits 8.. toiCount
    SC_MANIFOLD,   // type | pointCount << 8
    SC_ID0, SC_ID1,
    SC_LNX, SC_LNY, SC_LPX, SC_LPY,
    SC_P0X, SC_P0Y, SC_N0, SC_T0,
    SC_P1X, SC_P1Y, SC_N1, SC_T1,
    SC_TOI,
    SC_STRIDE
};
enum ContactWork {
    WC_NX = 0, WC_NY,
    WC_PT0,                       // per point: rA(2) rB(2) normalMass tangentMass velocityBias normalImpulse tangentImpulse
    WC_PT_STRIDE = 9,
    WC_K = WC_PT0 + 2 * WC_PT_STRIDE,  // k11 k12 k22
    WC_NM = WC_K + 3,                  // normalMass: ex.x, ex.y(=ey.x), ey.y
    WC_COUNT = WC_NM + 3,              // velocity-constraint point count (may drop to 1)
    WC_STRIDE
};
enum ScalarState {
    SS_INV_DT0 = 0,
    SS_FLAGS,     // bit0 newFixture
    SS_NCONTACTS,
    SS_STATUS,    // bit0 contact capacity exceeded, bit1 non-finite
    SS_MOVED_LO, SS_MOVED_HI,
    SS_TARGET_AVAIL,  // bit r: robot r has a controller target
    SS_SPARE,
    SS_STRIDE
};

}  // namespace b2g
