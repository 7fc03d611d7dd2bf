// Device-side building blocks of the batched Box2D backend (sm_100a, one world per thread).
//
// Data layout: the mutable state of the batch is one array of 16-byte chunks tiled by warp,
// state[tile][chunk][lane] (model.h), so a warp touching chunk k of its 32 worlds issues one coalesced 512-byte
// LDG.128 / STG.128.  The immutable model blob (model.h) is staged in shared memory by every CTA.
//
// Arithmetic: IEEE single precision, compiled with --fmad=false, IEEE division and square root, and a
// fixed-sequence sin/cos, so that results are reproducible bit for bit across launches, GPU counts and
// against the CPU oracle.  Formulas follow Box2D 2.3.1, which is what the
// reference calls through b2World::Step (src/sim_env/Box2DWorld.cpp:3279).
#pragma once
#include <cfloat>
#include <cstdio>
#include <cstdint>

#include "model.h"

namespace b2g {

// ---- Box2D constants (b2Settings.h 2.3.1) ------------------------------------------------------------
#define B2G_PI 3.14159265359f
#define B2G_LINEAR_SLOP 0.005f
#define B2G_ANGULAR_SLOP (2.0f / 180.0f * B2G_PI)
#define B2G_POLYGON_RADIUS (2.0f * B2G_LINEAR_SLOP)
#define B2G_AABB_EXTENSION 0.1f
#define B2G_AABB_MULTIPLIER 2.0f
#define B2G_MAX_SUB_STEPS 8
#define B2G_MAX_TOI_CONTACTS 32
#define B2G_VELOCITY_THRESHOLD 1.0f
#define B2G_MAX_LINEAR_CORRECTION 0.2f
#define B2G_MAX_ANGULAR_CORRECTION (8.0f / 180.0f * B2G_PI)
#define B2G_MAX_TRANSLATION 2.0f
#define B2G_MAX_ROTATION (0.5f * B2G_PI)
#define B2G_BAUMGARTE 0.2f
#define B2G_TOI_BAUMGARTE 0.75f
#define B2G_TIME_TO_SLEEP 0.5f
#define B2G_LINEAR_SLEEP_TOL 0.01f
#define B2G_ANGULAR_SLEEP_TOL (2.0f / 180.0f * B2G_PI)

#define DEV __device__ __forceinline__
#define DEVN static __device__ __noinline__  /* static: the kernels are compiled twice (B2G_SETS), host stubs must not clash */

struct V2 {
    float x, y;
};
struct V3 {
    float x, y, z;
};
struct Rot {
    float s, c;
};
struct XF {
    V2 p;
    Rot q;
};

DEV V2 mk(float x, float y) { V2 v; v.x = x; v.y = y; return v; }
DEV V2 operator+(V2 a, V2 b) { return mk(a.x + b.x, a.y + b.y); }
DEV V2 operator-(V2 a, V2 b) { return mk(a.x - b.x, a.y - b.y); }
DEV V2 operator-(V2 a) { return mk(-a.x, -a.y); }
DEV V2 operator*(float s, V2 a) { return mk(s * a.x, s * a.y); }
DEV float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
DEV float cross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
DEV V2 cross(V2 a, float s) { return mk(s * a.y, -s * a.x); }
DEV V2 cross(float s, V2 a) { return mk(-s * a.y, s * a.x); }
DEV float fminT(float a, float b) { return a < b ? a : b; }  // b2Min / b2Max are ternaries, not fminf
DEV float fmaxT(float a, float b) { return a > b ? a : b; }
DEV float clampT(float a, float lo, float hi) { return fmaxT(lo, fminT(a, hi)); }
DEV float absT(float a) { return a > 0.0f ? a : -a; }
DEV float length(V2 v) { return sqrtf(v.x * v.x + v.y * v.y); }
DEV float normalize(V2& v) {
    float len = length(v);
    if (len < FLT_EPSILON) return 0.0f;
    float inv = 1.0f / len;
    v.x *= inv;
    v.y *= inv;
    return len;
}

// Shared deterministic sin/cos: Cody-Waite reduction by pi/2 + Cephes single-precision cores.
DEV Rot rotOf(float x) {
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
    Rot o;
    if (q == 0) { o.s = ps; o.c = pc; }
    else if (q == 1) { o.s = pc; o.c = -ps; }
    else if (q == 2) { o.s = -ps; o.c = -pc; }
    else { o.s = -pc; o.c = ps; }
    return o;
}

DEV V2 mul(Rot q, V2 v) { return mk(q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y); }
DEV V2 mulT(Rot q, V2 v) { return mk(q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y); }
DEV V2 mul(const XF& T, V2 v) {
    float x = (T.q.c * v.x - T.q.s * v.y) + T.p.x;
    float y = (T.q.s * v.x + T.q.c * v.y) + T.p.y;
    return mk(x, y);
}
DEV V2 mulT(const XF& T, V2 v) {
    float px = v.x - T.p.x;
    float py = v.y - T.p.y;
    return mk((T.q.c * px + T.q.s * py), (-T.q.s * px + T.q.c * py));
}
DEV XF mulT(const XF& A, const XF& B) {
    XF C;
    C.q.s = A.q.c * B.q.s - A.q.s * B.q.c;
    C.q.c = A.q.c * B.q.c + A.q.s * B.q.s;
    C.p = mulT(A.q, B.p - A.p);
    return C;
}
DEV V2 solve22(float a11, float a12, float a21, float a22, V2 b) {
    float det = a11 * a22 - a12 * a21;
    if (det != 0.0f) det = 1.0f / det;
    V2 x;
    x.x = det * (a22 * b.x - a12 * b.y);
    x.y = det * (a11 * b.y - a21 * b.x);
    return x;
}

// ---- world context ----------------------------------------------------------------------------------------
// Shared memory of every kernel: [ModelHeader][StepParams][pad to 16 B][model blob].  It is addressed through this
// one extern array so that the compiler emits LDS (never generic loads) and knows model reads cannot alias the
// global-memory state.
extern __shared__ __align__(16) uint32_t g_sm[];
constexpr int kHdrWords = (int)((sizeof(ModelHeader) + sizeof(StepParams) + 15) / 16 * 4);

DEV const ModelHeader& H() { return *reinterpret_cast<const ModelHeader*>(g_sm); }
DEV const StepParams& PRM() { return *reinterpret_cast<const StepParams*>(g_sm + sizeof(ModelHeader) / 4); }

// B2G_PHASE_TIMING (tools/phase_timing.py; never in the product build): cycles per phase of the step, summed over warps.
// tick(id) charges the cycles since the warp's previous tick to phase `id`; warps are converged at the phase boundaries.
#ifdef B2G_PHASE_TIMING
__device__ unsigned long long g_phaseCycles[32];
DEV void tick(int id) {
    __shared__ long long s_last[32];
    const long long t = clock64();
    const unsigned m = __activemask();
    if ((int)(threadIdx.x & 31) == __ffs((int)m) - 1) {
        const int wp = threadIdx.x >> 5;
        if (id >= 0) atomicAdd(&g_phaseCycles[id], (unsigned long long)(t - s_last[wp]));
        s_last[wp] = clock64();
    }
    __syncwarp(m);
}
#else
DEV void tick(int) {}
#endif

// B2G_SETS: the kernels are compiled twice (build.py, b2g_launch.h).  0: all worlds of a batch share one model, staged once
// per CTA.  1: the batch holds several parameter sets; every warp stages the blob variant of its tile behind the previous
// warp's copy (b2g_kernels.cu: stageModel) and model reads carry the warp's offset.
#ifndef B2G_SETS
#define B2G_SETS 0
#endif
struct Ctx {
    char* P;  // &state[tile][0][lane]: chunk k of this thread's world is the float4 at P + k * 512
#if B2G_SETS
    int mo;   // word offset of this warp's copy of the model blob behind the first one
#endif
};
#if B2G_SETS
#define B2G_MO(c) ((c).mo)
#define B2G_SET_MO(c, v) ((c).mo = (v))
#else
#define B2G_MO(c) 0
#define B2G_SET_MO(c, v) ((void)(v))
#endif

// per-world state access: `off` is a byte offset (chunk * 512 + word * 4, see model.h); signed so that constant
// field offsets fold into the instruction's immediate
// B2G_CHECK (tools/checked_build.sh): the stand-in for compute-sanitizer's memcheck on pools where the tool is closed.  Every
// state access is checked against the world's chunk range (and its alignment), every model read against the blob size, every
// list append against its capacity; a violation prints what and where and traps the kernel.
#ifdef B2G_CHECK
#define B2G_ASSERT(cond, what, a, b)                                                                                    \
    do {                                                                                                                 \
        if (!(cond)) {                                                                                                   \
            printf("B2G_CHECK failed: %s (%d, %d) block %d thread %d\n", what, (int)(a), (int)(b), blockIdx.x, threadIdx.x); \
            __trap();                                                                                                    \
        }                                                                                                                \
    } while (0)
#else
#define B2G_ASSERT(cond, what, a, b) ((void)0)
#endif

template <int BYTES>
DEV char* spn(Ctx c, int off) {
    __builtin_assume(__isGlobal(c.P));
    B2G_ASSERT(off >= 0 && (off % kChunkBytes) + BYTES <= 16 && (off / kChunkBytes) < H().stateChunks && (off & (BYTES - 1)) == 0,
               "state access outside the world's chunks / misaligned", off, BYTES);
    return c.P + off;
}
DEV char* sp(Ctx c, int off) { return spn<4>(c, off); }
DEV float ldF(Ctx c, int off) { return *reinterpret_cast<const float*>(spn<4>(c, off)); }
DEV void stF(Ctx c, int off, float v) { *reinterpret_cast<float*>(spn<4>(c, off)) = v; }
DEV int ldI(Ctx c, int off) { return *reinterpret_cast<const int*>(spn<4>(c, off)); }
DEV void stI(Ctx c, int off, int v) { *reinterpret_cast<int*>(spn<4>(c, off)) = v; }
DEV float4 ld4(Ctx c, int off) { return *reinterpret_cast<const float4*>(spn<16>(c, off)); }
DEV void st4(Ctx c, int off, float4 v) { *reinterpret_cast<float4*>(spn<16>(c, off)) = v; }
DEV float2 ld2(Ctx c, int off) { return *reinterpret_cast<const float2*>(spn<8>(c, off)); }
DEV void st2(Ctx c, int off, float2 v) { *reinterpret_cast<float2*>(spn<8>(c, off)) = v; }
DEV float4 mk4(float x, float y, float z, float w) { return make_float4(x, y, z, w); }

// model blob access (shared memory, warp-uniform addresses: broadcast reads)
// (the region query reads one scratch fixture record behind the blob, b2g_kernels.cu: launchObjectsInAABB)
#define B2G_MODEL_OK(off, n) B2G_ASSERT((off) >= 0 && (off) + (n) <= H().words + 2 * MF_STRIDE && ((off) & ((n) - 1)) == 0, "model read outside the blob / misaligned", off, n)
DEV float mF(Ctx c, int off) { B2G_MODEL_OK(off, 1); return __uint_as_float(g_sm[kHdrWords + B2G_MO(c) + off]); }
DEV int mI(Ctx c, int off) { B2G_MODEL_OK(off, 1); return (int)g_sm[kHdrWords + B2G_MO(c) + off]; }
DEV float4 m4(Ctx c, int off) { B2G_MODEL_OK(off, 4); return *reinterpret_cast<const float4*>(g_sm + kHdrWords + B2G_MO(c) + off); }  // off % 4 == 0
DEV float2 m2(Ctx c, int off) { B2G_MODEL_OK(off, 2); return *reinterpret_cast<const float2*>(g_sm + kHdrWords + B2G_MO(c) + off); }  // off % 2 == 0

DEV int bslot(Ctx c, int b, int f) { return (H().cBody + b * SB_CHUNKS) * kChunkBytes + f; }
DEV int fslot(Ctx c, int f) { return (H().cFix + f) * kChunkBytes; }
DEV int jslot(Ctx c, int j, int f) { return (H().cJoint + j * SJ_CHUNKS) * kChunkBytes + f; }
DEV int jwslot(Ctx c, int j, int f) { return (H().cJointW + j * WJ_CHUNKS) * kChunkBytes + f; }
DEV int cslot(Ctx c, int k, int f) { return (H().cCont + k * SC_CHUNKS) * kChunkBytes + f; }
DEV int cwslot(Ctx c, int k, int f) { return (H().cContW + k * WC_CHUNKS) * kChunkBytes + f; }
DEV int sslot(Ctx c, int f) { return H().cScalar * kChunkBytes + f; }
DEV int tslot(Ctx c, int t) { return (H().cTarget + (t >> 2)) * kChunkBytes + (t & 3) * 4; }
DEV int mbody(Ctx c, int b, int f) { return H().oBody + b * MB_STRIDE + f; }
DEV int mfix(Ctx c, int f, int k) { return H().oFix + f * MF_STRIDE + k; }
DEV int mjoint(Ctx c, int j, int f) { return H().oJoint + j * MJ_STRIDE + f; }
DEV int mobj(Ctx c, int o, int f) { return H().oObj + o * MO_STRIDE + f; }

DEV V2 ldV2(Ctx c, int off) { float2 v = ld2(c, off); return mk(v.x, v.y); }  // x, y share a chunk, 8-byte aligned
DEV void stV2(Ctx c, int off, V2 v) { st2(c, off, make_float2(v.x, v.y)); }
DEV XF ldXF(Ctx c, int b) {
    float4 v = ld4(c, bslot(c, b, SB_XF4));
    XF x;
    x.p = mk(v.x, v.y);
    x.q.s = v.z;
    x.q.c = v.w;
    return x;
}
DEV void stXF(Ctx c, int b, const XF& x) { st4(c, bslot(c, b, SB_XF4), mk4(x.p.x, x.p.y, x.q.s, x.q.c)); }
DEV V2 localCenter(Ctx c, int b) { float2 v = m2(c, mbody(c, b, MB_LCX)); return mk(v.x, v.y); }
DEV int bodyType(Ctx c, int b) { return mI(c, mbody(c, b, MB_TYPE)); }
// awake bodies of the world as one mask (model.h: SS_AWAKE_LO): the same information as bit0 of every body's SB_FLAGS
DEV uint64_t ldAwake(Ctx c) {
    const float2 v = ld2(c, H().cScalar * kChunkBytes + SS_AWAKE_LO);
    return (uint64_t)(uint32_t)__float_as_int(v.x) | ((uint64_t)(uint32_t)__float_as_int(v.y) << 32);
}
DEV void stAwake(Ctx c, uint64_t m) {
    st2(c, H().cScalar * kChunkBytes + SS_AWAKE_LO, make_float2(__int_as_float((int)(uint32_t)m), __int_as_float((int)(uint32_t)(m >> 32))));
}
DEV bool isAwake(Ctx c, int b) { return ((ldAwake(c) >> b) & 1ull) != 0ull; }

// b2Body::SetAwake
DEV void wakeBody(Ctx c, int b) {
    const uint64_t m = ldAwake(c);
    if (((m >> b) & 1ull) == 0ull) {
        stAwake(c, m | (1ull << b));
        stI(c, bslot(c, b, SB_FLAGS), ldI(c, bslot(c, b, SB_FLAGS)) | 1);
        stF(c, bslot(c, b, SB_SLEEP), 0.0f);
    }
}
DEV void sleepBody(Ctx c, int b) {
    int f = ldI(c, bslot(c, b, SB_FLAGS));
    stAwake(c, ldAwake(c) & ~(1ull << b));
    st4(c, bslot(c, b, SB_VEL4), mk4(0.0f, 0.0f, 0.0f, __int_as_float(f & ~1)));
    stF(c, bslot(c, b, SB_SLEEP), 0.0f);
    if (mI(c, mbody(c, b, MB_FORCED))) st4(c, bslot(c, b, SB_FORCE4), mk4(0.0f, 0.0f, 0.0f, 0.0f));
}

// ---- broad phase: fat AABBs, move buffer as a 64-bit mask ---------------------------------------------------
struct Box {
    float lx, ly, ux, uy;
};
DEV Box ldFat(Ctx c, int f) {
    float4 v = ld4(c, fslot(c, f));
    Box b;
    b.lx = v.x; b.ly = v.y; b.ux = v.z; b.uy = v.w;
    return b;
}
DEV bool boxOverlap(const Box& a, const Box& b) {
    float d1x = b.lx - a.ux, d1y = b.ly - a.uy;
    float d2x = a.lx - b.ux, d2y = a.ly - b.uy;
    if (d1x > 0.0f || d1y > 0.0f) return false;
    if (d2x > 0.0f || d2y > 0.0f) return false;
    return true;
}
DEV uint64_t ldMoved(Ctx c) {
    return (uint64_t)(uint32_t)ldI(c, sslot(c, SS_MOVED_LO)) | ((uint64_t)(uint32_t)ldI(c, sslot(c, SS_MOVED_HI)) << 32);
}
DEV void stMoved(Ctx c, uint64_t m) {
    stI(c, sslot(c, SS_MOVED_LO), (int)(uint32_t)(m & 0xffffffffull));
    stI(c, sslot(c, SS_MOVED_HI), (int)(uint32_t)(m >> 32));
}

DEV V2 fixVertAt(Ctx c, int base, int i) { float2 v = m2(c, base + 2 * i); return mk(v.x, v.y); }

// b2PolygonShape::ComputeAABB
DEV Box fixtureAABB(Ctx c, int f, const XF& xf) {
    int n = mI(c, mfix(c, f, MF_COUNT));
    int vo = mfix(c, f, MF_VERTS);
    V2 lower = mul(xf, fixVertAt(c, vo, 0));
    V2 upper = lower;
    for (int i = 1; i < n; ++i) {
        V2 v = mul(xf, fixVertAt(c, vo, i));
        lower = mk(fminT(lower.x, v.x), fminT(lower.y, v.y));
        upper = mk(fmaxT(upper.x, v.x), fmaxT(upper.y, v.y));
    }
    Box b;
    b.lx = lower.x - B2G_POLYGON_RADIUS; b.ly = lower.y - B2G_POLYGON_RADIUS;
    b.ux = upper.x + B2G_POLYGON_RADIUS; b.uy = upper.y + B2G_POLYGON_RADIUS;
    return b;
}

// b2Fixture::Synchronize + b2DynamicTree::MoveProxy; returns the bit to add to the move buffer
DEV uint64_t syncFixture(Ctx c, int f, const XF& xf1, const XF& xf2) {
    Box a1 = fixtureAABB(c, f, xf1);
    Box a2 = fixtureAABB(c, f, xf2);
    Box a;
    a.lx = fminT(a1.lx, a2.lx); a.ly = fminT(a1.ly, a2.ly);
    a.ux = fmaxT(a1.ux, a2.ux); a.uy = fmaxT(a1.uy, a2.uy);
    V2 disp = xf2.p - xf1.p;
    Box fat = ldFat(c, f);
    bool contains = fat.lx <= a.lx && fat.ly <= a.ly && a.ux <= fat.ux && a.uy <= fat.uy;
    if (contains) return 0ull;
    Box b;
    b.lx = a.lx - B2G_AABB_EXTENSION; b.ly = a.ly - B2G_AABB_EXTENSION;
    b.ux = a.ux + B2G_AABB_EXTENSION; b.uy = a.uy + B2G_AABB_EXTENSION;
    V2 d = B2G_AABB_MULTIPLIER * disp;
    if (d.x < 0.0f) b.lx += d.x; else b.ux += d.x;
    if (d.y < 0.0f) b.ly += d.y; else b.uy += d.y;
    st4(c, fslot(c, f), mk4(b.lx, b.ly, b.ux, b.uy));
    return 1ull << f;
}

// b2Body::SynchronizeFixtures / SetTransform tail: fixtures in upstream list order (newest first)
DEV uint64_t syncBodyFixtures(Ctx c, int b, const XF& xf1, const XF& xf2) {
    int start = mI(c, mbody(c, b, MB_FIX_START));
    int cnt = mI(c, mbody(c, b, MB_FIX_COUNT));
    uint64_t moved = 0ull;
    for (int f = start + cnt - 1; f >= start; --f) moved |= syncFixture(c, f, xf1, xf2);
    return moved;
}

// b2Body::SetTransform (2.3.1: no FindNewContacts here)
DEV void setTransform(Ctx c, int b, V2 position, float angle) {
    XF xf;
    xf.q = rotOf(angle);
    xf.p = position;
    stXF(c, b, xf);
    V2 cc = mul(xf, localCenter(c, b));
    stV2(c, bslot(c, b, SB_CX), cc);
    stF(c, bslot(c, b, SB_A), angle);
    stV2(c, bslot(c, b, SB_C0X), cc);
    stF(c, bslot(c, b, SB_A0), angle);
    uint64_t moved = syncBodyFixtures(c, b, xf, xf);
    if (moved) stMoved(c, ldMoved(c) | moved);
}

// b2Body::SynchronizeTransform
DEV void syncTransform(Ctx c, int b) {
    XF xf;
    float a = ldF(c, bslot(c, b, SB_A));
    xf.q = rotOf(a);
    V2 cc = ldV2(c, bslot(c, b, SB_CX));
    xf.p = cc - mul(xf.q, localCenter(c, b));
    stXF(c, b, xf);
}

// ---- narrow phase: b2CollidePolygons (2.3.1) ---------------------------------------------------------------
struct Manifold {
    int type;        // 1 faceA, 2 faceB
    int pointCount;
    V2 localNormal, localPoint;
    V2 pt[2];
    uint32_t id[2];
};

DEV uint32_t makeId(int indexA, int indexB, int typeA, int typeB) {
    return (uint32_t)(indexA & 0xff) | ((uint32_t)(indexB & 0xff) << 8) | ((uint32_t)(typeA & 0xff) << 16) |
           ((uint32_t)(typeB & 0xff) << 24);
}

struct ClipV {
    V2 v;
    uint32_t id;
};

DEV int clipSegmentToLine(ClipV out[2], const ClipV in[2], V2 normal, float offset, int vertexIndexA) {
    int numOut = 0;
    float distance0 = dot(normal, in[0].v) - offset;
    float distance1 = dot(normal, in[1].v) - offset;
    if (distance0 <= 0.0f) out[numOut++] = in[0];
    if (distance1 <= 0.0f) out[numOut++] = in[1];
    if (distance0 * distance1 < 0.0f) {
        float interp = distance0 / (distance0 - distance1);
        out[numOut].v = in[0].v + interp * (in[1].v - in[0].v);
        // vertex (type 0) of A hitting face (type 1) of B
        out[numOut].id = makeId(vertexIndexA, (in[0].id >> 8) & 0xff, 0, 1);
        ++numOut;
    }
    return numOut;
}

DEV V2 fixVert(Ctx c, int f, int i) { return fixVertAt(c, mfix(c, f, MF_VERTS), i); }
DEV V2 fixNormal(Ctx c, int f, int i) { return fixVertAt(c, mfix(c, f, MF_NORMALS), i); }

DEV float findMaxSeparation(Ctx c, int* edgeIndex, int f1, const XF& xf1, int f2, const XF& xf2) {
    int count1 = mI(c, mfix(c, f1, MF_COUNT));
    int count2 = mI(c, mfix(c, f2, MF_COUNT));
    XF xf = mulT(xf2, xf1);
    int bestIndex = 0;
    float maxSeparation = -FLT_MAX;
    for (int i = 0; i < count1; ++i) {
        V2 n = mul(xf.q, fixNormal(c, f1, i));
        V2 v1 = mul(xf, fixVert(c, f1, i));
        float si = FLT_MAX;
        for (int j = 0; j < count2; ++j) {
            float sij = dot(n, fixVert(c, f2, j) - v1);
            if (sij < si) si = sij;
        }
        if (si > maxSeparation) { maxSeparation = si; bestIndex = i; }
    }
    *edgeIndex = bestIndex;
    return maxSeparation;
}

DEVN void collidePolygons(Ctx c, Manifold& m, int fA, const XF& xfA, int fB, const XF& xfB) {
    m.pointCount = 0;
    const float totalRadius = B2G_POLYGON_RADIUS + B2G_POLYGON_RADIUS;
    int edgeA = 0;
    float separationA = findMaxSeparation(c, &edgeA, fA, xfA, fB, xfB);
    if (separationA > totalRadius) return;
    int edgeB = 0;
    float separationB = findMaxSeparation(c, &edgeB, fB, xfB, fA, xfA);
    if (separationB > totalRadius) return;

    int f1, f2, edge1, flip;
    XF xf1, xf2;
    const float k_tol = 0.1f * B2G_LINEAR_SLOP;
    if (separationB > separationA + k_tol) {
        f1 = fB; f2 = fA; xf1 = xfB; xf2 = xfA; edge1 = edgeB; m.type = 2; flip = 1;
    } else {
        f1 = fA; f2 = fB; xf1 = xfA; xf2 = xfB; edge1 = edgeA; m.type = 1; flip = 0;
    }
    // b2FindIncidentEdge
    ClipV incident[2];
    {
        int count2 = mI(c, mfix(c, f2, MF_COUNT));
        V2 normal1 = mulT(xf2.q, mul(xf1.q, fixNormal(c, f1, edge1)));
        int index = 0;
        float minDot = FLT_MAX;
        for (int i = 0; i < count2; ++i) {
            float d = dot(normal1, fixNormal(c, f2, i));
            if (d < minDot) { minDot = d; index = i; }
        }
        int i1 = index;
        int i2 = i1 + 1 < count2 ? i1 + 1 : 0;
        incident[0].v = mul(xf2, fixVert(c, f2, i1));
        incident[0].id = makeId(edge1, i1, 1, 0);
        incident[1].v = mul(xf2, fixVert(c, f2, i2));
        incident[1].id = makeId(edge1, i2, 1, 0);
    }
    int count1 = mI(c, mfix(c, f1, MF_COUNT));
    int iv1 = edge1;
    int iv2 = edge1 + 1 < count1 ? edge1 + 1 : 0;
    V2 v11 = fixVert(c, f1, iv1);
    V2 v12 = fixVert(c, f1, iv2);
    V2 localTangent = v12 - v11;
    normalize(localTangent);
    V2 localNormal = cross(localTangent, 1.0f);
    V2 planePoint = 0.5f * (v11 + v12);
    V2 tangent = mul(xf1.q, localTangent);
    V2 normal = cross(tangent, 1.0f);
    v11 = mul(xf1, v11);
    v12 = mul(xf1, v12);
    float frontOffset = dot(normal, v11);
    float sideOffset1 = -dot(tangent, v11) + totalRadius;
    float sideOffset2 = dot(tangent, v12) + totalRadius;

    ClipV clip1[2], clip2[2];
    int np = clipSegmentToLine(clip1, incident, -tangent, sideOffset1, iv1);
    if (np < 2) return;
    np = clipSegmentToLine(clip2, clip1, tangent, sideOffset2, iv2);
    if (np < 2) return;

    m.localNormal = localNormal;
    m.localPoint = planePoint;
    int pointCount = 0;
    for (int i = 0; i < 2; ++i) {
        float separation = dot(normal, clip2[i].v) - frontOffset;
        if (separation <= totalRadius) {
            m.pt[pointCount] = mulT(xf2, clip2[i].v);
            uint32_t id = clip2[i].id;
            if (flip) id = ((id & 0xffu) << 8) | ((id >> 8) & 0xffu) | (((id >> 16) & 0xffu) << 24) | (((id >> 24) & 0xffu) << 16);
            m.id[pointCount] = id;
            ++pointCount;
        }
    }
    m.pointCount = pointCount;
}

// ---- contact slots ---------------------------------------------------------------------------------------
#define CF_TOUCHING 1
#define CF_ENABLED 2
#define CF_FILTER 4
#define CF_TOI 16

DEV int contactCount(Ctx c) { return ldI(c, sslot(c, SS_NCONTACTS)); }
DEV void contactFixtures(Ctx c, int k, int& fA, int& fB) {
    int p = ldI(c, cslot(c, k, SC_FIX));
    fA = p & 0xff;
    fB = (p >> 8) & 0xff;
}

// b2Contact::GetWorldManifold of slot k with the given transforms, reported the way the reference's collision checker
// does: points[0] and the normal, both multiplied by 1/scale (Box2DWorld.cpp:2797-2807, 2853-2863)
DEV void reportContact(Ctx c, int type, V2 localNormal, V2 localPoint, V2 lp, const XF& xfA, const XF& xfB, int bA, int bB,
                       int fA, int fB, int* io, float* fo) {
    V2 normal, point;
    const float radius = B2G_POLYGON_RADIUS;
    if (type == 1) {
        normal = mul(xfA.q, localNormal);
        V2 planePoint = mul(xfA, localPoint);
        V2 clipPoint = mul(xfB, lp);
        V2 cA = clipPoint + (radius - dot(clipPoint - planePoint, normal)) * normal;
        V2 cB = clipPoint - radius * normal;
        point = 0.5f * (cA + cB);
    } else {
        normal = mul(xfB.q, localNormal);
        V2 planePoint = mul(xfB, localPoint);
        V2 clipPoint = mul(xfA, lp);
        V2 cB = clipPoint + (radius - dot(clipPoint - planePoint, normal)) * normal;
        V2 cA = clipPoint - radius * normal;
        point = 0.5f * (cA + cB);
        normal = -normal;
    }
    const float inv = H().invScale;
    io[0] = bA; io[1] = bB; io[2] = fA; io[3] = fB;
    fo[0] = inv * point.x; fo[1] = inv * point.y; fo[2] = 0.0f;
    fo[3] = inv * normal.x; fo[4] = inv * normal.y; fo[5] = 0.0f;
}

// Box2DCollisionChecker::BeginContact (:2792-2818) for contact slot k, whose manifold has just been updated with the
// bodies' current transforms: recording and / or the guard of stepPhysics(callback, ...)
// world index of the calling thread, recovered from its state pointer (threads are mapped to worlds differently by
// the different stepping kernels)
DEV long long worldIndex(Ctx c) {
    const size_t off = (size_t)(c.P - PRM().stateBase);
    const size_t tileBytes = (size_t)H().stateChunks * kChunkBytes;
    return (long long)(off / tileBytes) * kTileWorlds + (long long)((off % tileBytes) >> 4);
}

DEVN void onBeginContact(Ctx c, int k) {
    const StepParams& p = PRM();
    const long long w = worldIndex(c);
    const int ck = cslot(c, k, 0);
    const int fx = ldI(c, ck + SC_FIX);
    const int fA = fx & 0xff, fB = (fx >> 8) & 0xff;
    const int bA = mI(c, mfix(c, fA, MF_BODY)), bB = mI(c, mfix(c, fB, MF_BODY));
    if (p.recMode & 1) {
        int n = p.recCounts[w];
        if (n < p.recMax)
            reportContact(c, ldI(c, ck + SC_MANIFOLD) & 0xff, ldV2(c, ck + SC_LNX), ldV2(c, ck + SC_LPX), ldV2(c, ck + SC_P0X),
                          ldXF(c, bA), ldXF(c, bB), bA, bB, fA, fB, p.recI + ((size_t)w * p.recMax + n) * 4,
                          p.recF + ((size_t)w * p.recMax + n) * 6);
        p.recCounts[w] = n + 1;
    }
    if ((p.recMode & 2) && p.allowed[bA * H().nb + bB] == 0) stI(c, sslot(c, SS_ABORT), ldI(c, sslot(c, SS_ABORT)) | 1);
}

// The side-effect-free part of b2Contact::Update for slot k: b2CollidePolygons on the fixtures' current transforms (the
// sub-warp kernel evaluates the contacts of a world in parallel, one per lane, and commits them in list order afterwards)
DEV void evalContact(Ctx c, int k, Manifold& m) {
    const float4 head = ld4(c, cslot(c, k, SC_HEAD4));
    const int fA = __float_as_int(head.x) & 0xff, fB = (__float_as_int(head.x) >> 8) & 0xff;
    const XF xfA = ldXF(c, mI(c, mfix(c, fA, MF_BODY)));
    const XF xfB = ldXF(c, mI(c, mfix(c, fB, MF_BODY)));
    m.type = __float_as_int(head.z) & 0xff;
    collidePolygons(c, m, fA, xfA, fB, xfB);
}

// b2Contact::Update for slot k.  Returns bit0 = now touching, bit1 = was touching.  `pre`: the manifold evalContact computed
// for this slot from the same transforms, or nullptr to evaluate here.
DEVN int updateContact(Ctx c, int k, const Manifold* pre = nullptr) {
    const int ck = cslot(c, k, 0);
    float4 head = ld4(c, ck + SC_HEAD4);
    float4 idn = ld4(c, ck + SC_ID4);
    float4 p0 = ld4(c, ck + SC_P04);
    float4 p1 = ld4(c, ck + SC_P14);
    int fA = __float_as_int(head.x) & 0xff, fB = (__float_as_int(head.x) >> 8) & 0xff;
    int bA = mI(c, mfix(c, fA, MF_BODY));
    int bB = mI(c, mfix(c, fB, MF_BODY));
    int flags = __float_as_int(head.y);
    flags |= CF_ENABLED;
    bool wasTouching = (flags & CF_TOUCHING) != 0;
    int oldMan = __float_as_int(head.z);
    int oldCount = (oldMan >> 8) & 0xff;
    uint32_t oldId0 = (uint32_t)__float_as_int(idn.x);
    uint32_t oldId1 = (uint32_t)__float_as_int(idn.y);
    float oldN0 = p0.z, oldT0 = p0.w;
    float oldN1 = p1.z, oldT1 = p1.w;

    Manifold m;
    if (pre != nullptr) {
        m = *pre;
    } else {
        XF xfA = ldXF(c, bA);
        XF xfB = ldXF(c, bB);
        m.type = oldMan & 0xff;
        collidePolygons(c, m, fA, xfA, fB, xfB);
    }
    bool touching = m.pointCount > 0;
    if (touching) {
        // b2CollidePolygons returns early (pointCount = 0) without touching the rest of the manifold, so the
        // other fields are only rewritten when points were produced
        idn.z = m.localNormal.x;
        idn.w = m.localNormal.y;
        stV2(c, ck + SC_LPX, m.localPoint);
        for (int i = 0; i < m.pointCount; ++i) {
            float ni = 0.0f, ti = 0.0f;
            uint32_t id2 = m.id[i];
            if (oldCount > 0 && oldId0 == id2) { ni = oldN0; ti = oldT0; }
            else if (oldCount > 1 && oldId1 == id2) { ni = oldN1; ti = oldT1; }
            st4(c, ck + (i == 0 ? SC_P04 : SC_P14), mk4(m.pt[i].x, m.pt[i].y, ni, ti));
            if (i == 0) idn.x = __int_as_float((int)id2); else idn.y = __int_as_float((int)id2);
        }
        st4(c, ck + SC_ID4, idn);
    }
    if (touching != wasTouching) {
        wakeBody(c, bA);
        wakeBody(c, bB);
    }
    if (touching) flags |= CF_TOUCHING; else flags &= ~CF_TOUCHING;
    head.y = __int_as_float(flags);
    head.z = __int_as_float((m.type & 0xff) | (m.pointCount << 8));
    st4(c, ck + SC_HEAD4, head);
    return (touching ? 1 : 0) | (wasTouching ? 2 : 0);
}

// b2ContactManager::Destroy + b2Contact::Destroy for slot k; later slots move down (creation order kept)
DEVN void destroyContact(Ctx c, int k) {
    int fA, fB;
    contactFixtures(c, k, fA, fB);
    int pc = (ldI(c, cslot(c, k, SC_MANIFOLD)) >> 8) & 0xff;
    if (pc > 0) {
        wakeBody(c, mI(c, mfix(c, fA, MF_BODY)));
        wakeBody(c, mI(c, mfix(c, fB, MF_BODY)));
    }
    int n = contactCount(c);
    for (int i = k; i + 1 < n; ++i)
        for (int f = 0; f < SC_CHUNKS; ++f) st4(c, cslot(c, i, f * kChunkBytes), ld4(c, cslot(c, i + 1, f * kChunkBytes)));
    stI(c, sslot(c, SS_NCONTACTS), n - 1);
}

// b2ContactManager::AddPair (fixture a < fixture b by proxy id; filters are folded into the pair mask)
DEV uint64_t ldDisabled(Ctx c) {
    return (uint64_t)(uint32_t)ldI(c, sslot(c, SS_DISABLED_LO)) | ((uint64_t)(uint32_t)ldI(c, sslot(c, SS_DISABLED_HI)) << 32);
}
// b2ContactFilter::ShouldCollide with the only filter data the reference ever sets: categoryBits = enabled ? 1 : 0,
// maskBits = 0xFFFF, groupIndex = 0 (Box2DWorld.cpp:815-821)
DEV bool filterShouldCollide(Ctx c, int a, int b) {
    uint64_t d = ldDisabled(c);
    return ((d >> a) & 1ull) == 0ull && ((d >> b) & 1ull) == 0ull;
}

DEV void addPair(Ctx c, int a, int b) {
    if (!filterShouldCollide(c, a, b)) return;
    int n = contactCount(c);
    int key = a | (b << 8);
    for (int k = 0; k < n; ++k)
        if (ldI(c, cslot(c, k, SC_FIX)) == key) return;
    if (n >= H().maxc) {
        stI(c, sslot(c, SS_STATUS), ldI(c, sslot(c, SS_STATUS)) | 1);
        return;
    }
    const int cn = cslot(c, n, 0);
    st4(c, cn + SC_HEAD4, mk4(__int_as_float(key), __int_as_float(CF_ENABLED), 0.0f, 1.0f));
    const float4 zero = mk4(0.0f, 0.0f, 0.0f, 0.0f);
    st4(c, cn + SC_ID4, zero);
    {   // b2Contact::b2Contact: the mixed coefficients are fixed when the contact is created (b2Fixture::SetFriction does
        // not reach live contacts; Box2DLink::setContactFriction, Box2DWorld.cpp:538-547)
        const float2 frA = m2(c, mfix(c, a, MF_FRICTION)), frB = m2(c, mfix(c, b, MF_FRICTION));  // friction, restitution
        const float friction = sqrtf(frA.x * frB.x);                  // b2MixFriction
        const float restitution = frA.y > frB.y ? frA.y : frB.y;      // b2MixRestitution
        st4(c, cn + SC_LP4, mk4(0.0f, 0.0f, friction, restitution));
    }
    st4(c, cn + SC_P04, zero);
    st4(c, cn + SC_P14, zero);
    stI(c, sslot(c, SS_NCONTACTS), n + 1);
    wakeBody(c, mI(c, mfix(c, a, MF_BODY)));
    wakeBody(c, mI(c, mfix(c, b, MF_BODY)));
}

// b2ContactManager::FindNewContacts -> b2BroadPhase::UpdatePairs.  Upstream queries the tree for every
// buffered proxy, sorts the (min,max) pairs and drops duplicates; iterating a < b in order over pairs with
// at least one moved member and overlapping fat boxes visits exactly that sorted, de-duplicated sequence.
// fixtures b > a whose fat box overlaps fixture a's and that may pair with it in this UpdatePairs round (`moved` = move buffer)
DEV uint64_t newPairHits(Ctx c, int a, uint64_t moved) {
    uint64_t cand = (uint64_t)(uint32_t)mI(c, mfix(c, a, MF_PAIR_LO)) | ((uint64_t)(uint32_t)mI(c, mfix(c, a, MF_PAIR_HI)) << 32);
    cand &= ~((2ull << a) - 1ull);
    if (((moved >> a) & 1ull) == 0ull) cand &= moved;
    if (cand == 0ull) return 0ull;
    Box fa = ldFat(c, a);
    // the overlap tests first (independent loads, nothing stored in between: they pipeline), the pairs are added afterwards
    uint64_t hit = 0ull;
    while (cand) {
        const int b0 = __ffsll((long long)cand) - 1;
        cand &= cand - 1ull;
        const int b1 = cand ? __ffsll((long long)cand) - 1 : b0;
        cand &= cand - 1ull;   // 0 & (0 - 1) = 0
        const int b2 = cand ? __ffsll((long long)cand) - 1 : b0;
        cand &= cand - 1ull;
        const int b3 = cand ? __ffsll((long long)cand) - 1 : b0;
        cand &= cand - 1ull;
        const Box f0 = ldFat(c, b0), f1 = ldFat(c, b1), f2 = ldFat(c, b2), f3 = ldFat(c, b3);
        if (boxOverlap(f0, fa)) hit |= 1ull << b0;
        if (boxOverlap(f1, fa)) hit |= 1ull << b1;
        if (boxOverlap(f2, fa)) hit |= 1ull << b2;
        if (boxOverlap(f3, fa)) hit |= 1ull << b3;
    }
    return hit;
}

// b2ContactManager::FindNewContacts -> b2BroadPhase::UpdatePairs.  Upstream queries the tree for every
// buffered proxy, sorts the (min,max) pairs and drops duplicates; iterating a < b in order over pairs with
// at least one moved member and overlapping fat boxes visits exactly that sorted, de-duplicated sequence.
// `hits`: per-fixture hit masks computed beforehand by the lanes of the world's group (sub-warp kernel), or nullptr.
DEVN void findNewContacts(Ctx c, const uint64_t* hits = nullptr) {
    uint64_t moved = ldMoved(c);
    if (moved == 0ull) return;
    int nf = H().nf;
    for (int a = 0; a < nf; ++a) {
        uint64_t hit = hits ? hits[a] : newPairHits(c, a, moved);
        while (hit) {
            const int b = __ffsll((long long)hit) - 1;
            hit &= hit - 1ull;
            addPair(c, a, b);
        }
    }
    stMoved(c, 0ull);
}

// b2ContactManager::Collide: contacts in upstream list order (newest first)
// `man` / `valid`: manifolds evalContact computed for slots k < 64 with bit k of `valid` set (sub-warp kernel), else none
DEVN void collide(Ctx c, const Manifold* man = nullptr, uint64_t valid = 0ull) {
    for (int k = contactCount(c) - 1; k >= 0; --k) {
        int fA, fB;
        contactFixtures(c, k, fA, fB);
        int bA = mI(c, mfix(c, fA, MF_BODY));
        int bB = mI(c, mfix(c, fB, MF_BODY));
        {   // contact flagged for filtering (b2Fixture::Refilter): body-level ShouldCollide is static here (pair masks)
            int fl = ldI(c, cslot(c, k, SC_FLAGS));
            if (fl & CF_FILTER) {
                if (!filterShouldCollide(c, fA, fB)) { destroyContact(c, k); continue; }
                stI(c, cslot(c, k, SC_FLAGS), fl & ~CF_FILTER);
            }
        }
        bool activeA = isAwake(c, bA) && bodyType(c, bA) != 0;
        bool activeB = isAwake(c, bB) && bodyType(c, bB) != 0;
        if (!activeA && !activeB) continue;
        Box a = ldFat(c, fA);
        Box b = ldFat(c, fB);
        if (!boxOverlap(a, b)) { destroyContact(c, k); continue; }
        const Manifold* pre = (k < 64 && ((valid >> k) & 1ull)) ? man + k : nullptr;
#ifndef B2G_NO_HOOK_COLLIDE
        if (updateContact(c, k, pre) == 1 && PRM().recMode != 0) onBeginContact(c, k);
#else
        updateContact(c, k, pre);
#endif
    }
}

}  // namespace b2g
