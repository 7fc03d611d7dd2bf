// CUDA kernels of the batched Box2D backend (sm_100a).  One thread owns one world; the model blob is staged
// in shared memory once per CTA; all per-world state lives in HBM as S[slot][world].
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a --fmad=false -lineinfo (see build.py).  --fmad=false and
// IEEE div/sqrt keep the arithmetic identical to the scalar SSE code the reference is compiled to
// (/root/reference/CMakeLists.txt:54-59: -O3, no -march, no fast-math).
#include <cuda_runtime.h>

#include <atomic>
#include <cstdlib>

#include "b2g_launch.h"
#include "b2g_query.cuh"
#include "b2g_world.cuh"
#include "b2g_multi.cuh"

#if B2G_SETS
#define B2G_VNS sets
#else
#define B2G_VNS one
#endif

namespace b2g {
namespace B2G_VNS {

// Stage header + step parameters + model blob into shared memory.  One model for the whole batch: one copy per CTA, every
// thread reads it at offset 0.  Several parameter sets (StepParams::tileVariant): every warp stages the blob variant of ITS
// tile behind the previous warp's copy and its threads read that one (Ctx::mo); `warpWorld` = the first world of the calling
// thread's warp.  Returns the word offset of the copy this thread reads.
DEV int stageModel(const ModelHeader& hdr, const StepParams& prm, const uint32_t* __restrict__ blob, long long warpWorld) {
    const uint32_t* hsrc = reinterpret_cast<const uint32_t*>(&hdr);
    const uint32_t* psrc = reinterpret_cast<const uint32_t*>(&prm);
    const int hw = (int)(sizeof(ModelHeader) / 4), pw = (int)(sizeof(StepParams) / 4);
    for (int i = threadIdx.x; i < hw; i += blockDim.x) g_sm[i] = hsrc[i];
    for (int i = threadIdx.x; i < pw; i += blockDim.x) g_sm[hw + i] = psrc[i];
    int mo = 0;
    if (!B2G_SETS || prm.tileVariant == nullptr) {
        for (int i = threadIdx.x; i < hdr.words; i += blockDim.x) g_sm[kHdrWords + i] = blob[i];
        __syncthreads();
        for (int j = threadIdx.x; j < hdr.nj; j += blockDim.x) {  // dt-dependent joint constants
            uint32_t* jr = g_sm + kHdrWords + hdr.oJoint + j * MJ_STRIDE;
            jr[MJ_HMAXFORCE] = __float_as_uint(prm.dt * __uint_as_float(jr[MJ_MAXFORCE]));
            jr[MJ_HMAXTORQUE] = __float_as_uint(prm.dt * __uint_as_float(jr[MJ_MAXTORQUE]));
        }
    } else {
        const int lane = threadIdx.x & 31;
        mo = (int)(threadIdx.x >> 5) * prm.blobStride;
        const uint32_t* src = blob + (size_t)prm.tileVariant[warpWorld >> 5] * prm.blobStride;
        for (int i = lane; i < hdr.words; i += 32) g_sm[kHdrWords + mo + i] = src[i];
        __syncwarp();
        for (int j = lane; j < hdr.nj; j += 32) {
            uint32_t* jr = g_sm + kHdrWords + mo + hdr.oJoint + j * MJ_STRIDE;
            jr[MJ_HMAXFORCE] = __float_as_uint(prm.dt * __uint_as_float(jr[MJ_MAXFORCE]));
            jr[MJ_HMAXTORQUE] = __float_as_uint(prm.dt * __uint_as_float(jr[MJ_MAXTORQUE]));
        }
    }
    __syncthreads();
    return mo;
}

// per-world kernels: thread g of the grid owns world g - lead + worldOffset of the selection [worldOffset, worldOffset + N)
#if B2G_SETS
#define B2G_PROLOGUE()                                                                                          \
    const long long b2g_g = (long long)blockIdx.x * blockDim.x + threadIdx.x;                                   \
    const int b2g_mo = stageModel(hdr, prm, blob, (b2g_g & ~31ll) - prm.lead + prm.worldOffset);                \
    const long long w = b2g_g - prm.lead;                                                                       \
    if (w < 0 || w >= N) return;                                                                                \
    Ctx c;                                                                                                      \
    c.P = worldBase(S, hdr.stateChunks, w + prm.worldOffset);                                                   \
    c.mo = b2g_mo;
#else
#define B2G_PROLOGUE()                                                                                          \
    stageModel(hdr, prm, blob, 0);                                                                              \
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;                                       \
    if (w >= N) return;                                                                                         \
    Ctx c;                                                                                                      \
    c.P = worldBase(S, hdr.stateChunks, w + prm.worldOffset);
#endif

// &state[tile][0][lane] of world w
DEV char* worldBase(float* S, int stateChunks, long long w) {
    return reinterpret_cast<char*>(S) + (size_t)(w >> kTileShift) * ((size_t)stateChunks * kChunkBytes) + (size_t)(w & (kTileWorlds - 1)) * 16;
}

__global__ void k_init(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S,
                       long long N, StepParams prm) {
    B2G_PROLOGUE();
    initWorld(c);
}

// broadcast the state of a template world (stored [chunk]) to worlds [first, first + count):
// state[tile][chunk][lane] = tmpl[chunk]; walks whole tiles so that every store instruction is a coalesced 512-byte row
__global__ void k_broadcast(float4* S, const float4* __restrict__ tmpl, long long first, long long count, int chunks) {
    const long long tile0 = first >> kTileShift, tile1 = (first + count + kTileWorlds - 1) >> kTileShift;
    const long long total = (tile1 - tile0) * chunks * kTileWorlds;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long world = ((tile0 + i / ((long long)chunks * kTileWorlds)) << kTileShift) + (i & (kTileWorlds - 1));
        if (world >= first && world < first + count) S[tile0 * chunks * kTileWorlds + i] = tmpl[(i >> kTileShift) % chunks];
    }
}

// Padding worlds (behind the batch, model.h: paddedTiles) exist so that every thread of a k_step CTA owns real memory.  They
// are parked: every body asleep, nothing buffered for the broad phase, no new-fixture flag -- a step of a parked world finds
// no island, no pair and no contact, whatever the loaded scene looks like.
__global__ void k_park(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S, long long N,
                       StepParams prm) {
    B2G_PROLOGUE();
    for (int b = 0; b < H().nb; ++b) sleepBody(c, b);
    stI(c, sslot(c, SS_FLAGS), ldI(c, sslot(c, SS_FLAGS)) & ~1);
    stMoved(c, 0ull);
}

__global__ void k_extract_world(const float4* S, float4* tmpl, int chunks, long long world) {
    const float4* base = S + (size_t)(world >> kTileShift) * ((size_t)chunks * kTileWorlds) + (world & (kTileWorlds - 1));
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < chunks; i += gridDim.x * blockDim.x) tmpl[i] = base[(size_t)i * kTileWorlds];
}

__global__ void k_set_state(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S,
                            long long N, StepParams prm, const float* __restrict__ q, const float* __restrict__ qd) {
    B2G_PROLOGUE();
    const int nq = H().nq;
    for (int o = 0; o < H().no; ++o) {
        int n = mI(c, mobj(c, o, MO_NDOFS));
        int off = mI(c, mobj(c, o, MO_DOF_OFFSET));
        float lq[kMaxDofsPerObject], lqd[kMaxDofsPerObject];
        for (int i = 0; i < n; ++i) {
            lq[i] = q[w * nq + off + i];
            lqd[i] = qd[w * nq + off + i];
        }
        objectSetState(c, o, lq, lqd);
    }
}

__global__ void k_get_state(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S,
                            long long N, StepParams prm, float* __restrict__ q, float* __restrict__ qd) {
    B2G_PROLOGUE();
    const int nq = H().nq;
    for (int o = 0; o < H().no; ++o) {
        int n = mI(c, mobj(c, o, MO_NDOFS));
        int off = mI(c, mobj(c, o, MO_DOF_OFFSET));
        float lq[kMaxDofsPerObject], lqd[kMaxDofsPerObject];
        getDofPositions(c, o, lq);
        getDofVelocities(c, o, lqd);
        for (int i = 0; i < n; ++i) {
            q[w * nq + off + i] = lq[i];
            qd[w * nq + off + i] = lqd[i];
        }
    }
}

// Box2DObject::setDOFPositions / setDOFVelocities(values, indices) (:1585-1626, :1691-1715) of one object in every world;
// values: [N][sel.n]
__global__ void k_set_object_dofs(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S,
                                  long long N, StepParams prm, int object, const __grid_constant__ DofSel sel,
                                  const float* __restrict__ values, int velocities) {
    B2G_PROLOGUE();
    float v[kMaxDofsPerObject];
    int idx[kMaxDofsPerObject];
    for (int i = 0; i < sel.n; ++i) {
        v[i] = values[w * sel.n + i];
        idx[i] = sel.idx[i];
    }
    if (velocities) objectSetDofVelocitiesIdx(c, object, v, idx, sel.n);
    else objectSetDofPositionsIdx(c, object, v, idx, sel.n);
}

// Box2DObject::getDOFPositions / getDOFVelocities(indices) (:1507-1528, :1628-1649); out: [N][sel.n]
__global__ void k_get_object_dofs(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S,
                                  long long N, StepParams prm, int object, const __grid_constant__ DofSel sel,
                                  float* __restrict__ out, int velocities) {
    B2G_PROLOGUE();
    float all[kMaxDofsPerObject];
    if (velocities) getDofVelocities(c, object, all);
    else getDofPositions(c, object, all);
    for (int i = 0; i < sel.n; ++i) out[w * sel.n + i] = all[sel.idx[i]];
}

__global__ void k_set_pose(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S,
                           long long N, StepParams prm, int object, const float* __restrict__ pose) {
    B2G_PROLOGUE();
    objectSetPose(c, object, pose[w * 3], pose[w * 3 + 1], pose[w * 3 + 2]);
}

__global__ void k_get_pose(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S,
                           long long N, StepParams prm, int object, float* __restrict__ pose) {
    B2G_PROLOGUE();
    float p[3];
    basePose(c, object, p);
    pose[w * 3] = p[0];
    pose[w * 3 + 1] = p[1];
    pose[w * 3 + 2] = p[2];
}

// Box2DRobotVelocityController::setTargetVelocity (Box2DController.cpp:42-63).  scaleToLimits belongs to the
// missing sim_env package; assumed: one uniform factor that brings every component inside its velocity limits.
DEV void setTarget(const Ctx& c, int object, const float* v) {
    int n = mI(c, mobj(c, object, MO_NDOFS));
    int tOff = mI(c, mobj(c, object, MO_TARGET_OFFSET));
    const float factor = scaleToLimits(c, object, v);
    // the scaling rule is an assumption (sim_env is absent): worlds whose results depend on it say so in their status word
    if (factor < 1.0f) stI(c, sslot(c, SS_STATUS), ldI(c, sslot(c, SS_STATUS)) | 4);
    for (int i = 0; i < n; ++i) stF(c, tslot(c, tOff + i), factor < 1.0f ? factor * v[i] : v[i]);
    int slot = 0;
    for (int r = 0; r < H().nrobots; ++r)
        if (mI(c, H().oRobotOrder + r) == object) slot = r;
    stI(c, sslot(c, SS_TARGET_AVAIL), (ldI(c, sslot(c, SS_TARGET_AVAIL)) | (1 << slot)) & ~((1 << (16 + slot)) | (1 << (8 + slot))));
}

// Position controller on top of the velocity controller (sim_env::RobotPositionController / SE2RobotPositionController,
// instantiated at Box2DWorldViewer.cpp:1208-1216; the classes belong to the absent sim_env package -- assumed law, see
// robotControl): stores the target positions and the gain and switches the robot's controller to position mode (bit 8 + slot).
DEV void setTargetPosition(const Ctx& c, int object, const float* q, float kp) {
    int n = mI(c, mobj(c, object, MO_NDOFS));
    int tOff = mI(c, mobj(c, object, MO_TARGET_OFFSET));
    for (int i = 0; i < n; ++i) stF(c, tslot(c, tOff + i), q[i]);
    stF(c, tslot(c, tOff + n), kp);
    int slot = 0;
    for (int r = 0; r < H().nrobots; ++r)
        if (mI(c, H().oRobotOrder + r) == object) slot = r;
    stI(c, sslot(c, SS_TARGET_AVAIL), (ldI(c, sslot(c, SS_TARGET_AVAIL)) | (1 << slot) | (1 << (8 + slot))) & ~(1 << (16 + slot)));
}

// Box2DRobot::setController (:2312-2317) with a callback that returns the given efforts on every control(dt) call
// (:2284-2310): they go straight to commandEfforts (:2319-2379).  Shares the target slots with the velocity controller.
DEV void setEfforts(const Ctx& c, int object, const float* e) {
    int n = mI(c, mobj(c, object, MO_NDOFS));
    int tOff = mI(c, mobj(c, object, MO_TARGET_OFFSET));
    for (int i = 0; i < n; ++i) stF(c, tslot(c, tOff + i), e[i]);
    int slot = 0;
    for (int r = 0; r < H().nrobots; ++r)
        if (mI(c, H().oRobotOrder + r) == object) slot = r;
    stI(c, sslot(c, SS_TARGET_AVAIL), (ldI(c, sslot(c, SS_TARGET_AVAIL)) | (1 << slot) | (1 << (16 + slot))) & ~(1 << (8 + slot)));
}

__global__ void k_set_target(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S,
                             long long N, StepParams prm, int object, const float* __restrict__ v) {
    B2G_PROLOGUE();
    int n = mI(c, mobj(c, object, MO_NDOFS));
    float lv[kMaxDofsPerObject];
    for (int i = 0; i < n; ++i) lv[i] = v[w * n + i];
    setTarget(c, object, lv);
}

__global__ void k_set_target_position(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S, long long N,
                                      StepParams prm, int object, const float* __restrict__ q, float kp) {
    B2G_PROLOGUE();
    int n = mI(c, mobj(c, object, MO_NDOFS));
    float lq[kMaxDofsPerObject];
    for (int i = 0; i < n; ++i) lq[i] = q[w * n + i];
    setTargetPosition(c, object, lq, kp);
}

__global__ void k_set_efforts(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S, long long N,
                              StepParams prm, int object, const float* __restrict__ v) {
    B2G_PROLOGUE();
    int n = mI(c, mobj(c, object, MO_NDOFS));
    float lv[kMaxDofsPerObject];
    for (int i = 0; i < n; ++i) lv[i] = v[w * n + i];
    setEfforts(c, object, lv);
}

// stepPhysics / rollout: nTargets segments of stepsPerTarget steps; targets == nullptr keeps the stored target
// Two register budgets (the kernel is latency-bound: more resident warps hide more, up to a point where spills cost more):
//   k_step<640>: 96 registers (about 100 bytes of spills in cold code), up to 20 warps per SM -- batches up to 148 x 640 worlds
//   k_step<896>: 72 registers, up to 28 warps per SM -- larger batches
// Measured on 148 CTAs (one per SM) of cfg 2, world-steps/s: 16 warps 1.85e8 (113 or 96 registers), 20 warps at 96 registers
// 2.09e8, 24 warps at 80 2.2e8, 28 warps at 72 2.36e8, 32 warps at 64 2.34e8; at 14 warps per SM (65536 worlds) 96 registers
// take 38.4 ms, 72 registers 41.6 ms (profiles/r2_block_sweep.txt).  131072 worlds: 76.9 ms in two waves of 14 warps -> 56.4 ms
// in one wave of 28; 1 Mi worlds 505 -> 431 ms.
template <int BOUNDS>
__global__ void __launch_bounds__(BOUNDS) k_step(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S, long long N,
                       StepParams prm, int object, const float* __restrict__ targets, int nTargets, int stepsPerTarget,
                       int executeControllers, int allowSleeping) {
    // CTA-synchronised stepping: every thread of the CTA owns a world that exists in memory (the state array is padded to
    // a multiple of kWorldPad worlds, the padding worlds are valid idle copies of the loaded world), so all threads walk
    // the phase barriers of stepWorld<true> together.
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    Ctx c;
    B2G_SET_MO(c, stageModel(hdr, prm, blob, w & ~31ll));
    c.P = worldBase(S, hdr.stateChunks, w);
    for (int t = 0; t < nTargets; ++t) {
        if (targets != nullptr && w < N) {
            int n = mI(c, mobj(c, object, MO_NDOFS));
            float lv[kMaxDofsPerObject];
            for (int i = 0; i < n; ++i) lv[i] = targets[((size_t)t * N + w) * n + i];
            setTarget(c, object, lv);
        }
        for (int s = 0; s < stepsPerTarget; ++s) stepWorld<true>(c, executeControllers != 0, allowSleeping != 0);
    }
}

// Small batches: stepping with only `lanes` (1..31) worlds per warp.  A batch that cannot fill the GPU's 592 warp schedulers
// gains nothing from full warps, but every rare event (a contact, a TOI sub-step, a sleeping body) in one lane makes the
// whole warp walk that path: spreading the worlds over more, narrower warps -- at most one per scheduler -- cuts that
// serialisation.  Warp g owns worlds [g * lanes, (g + 1) * lanes): the state layout is unchanged, a narrow warp reads
// `lanes` x 16 contiguous bytes of each chunk (two pieces when its worlds straddle a tile boundary).
__global__ void k_step_narrow(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S, long long N,
                              StepParams prm, int object, const float* __restrict__ targets, int nTargets, int stepsPerTarget,
                              int executeControllers, int allowSleeping, int lanes) {
    const long long g = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long w = g * lanes + lane;
    // with parameter sets `lanes` divides 32 (b2g_capi.cpp), so the worlds of a warp share a tile; a warp behind the batch reads tile 0
    const int mo = stageModel(hdr, prm, blob, g * lanes < N ? g * lanes : 0);
    if (lane >= lanes || w >= N) return;
    Ctx c;
    B2G_SET_MO(c, mo);
    c.P = worldBase(S, hdr.stateChunks, w);
    for (int t = 0; t < nTargets; ++t) {
        if (targets != nullptr) {
            int n = mI(c, mobj(c, object, MO_NDOFS));
            float lv[kMaxDofsPerObject];
            for (int i = 0; i < n; ++i) lv[i] = targets[((size_t)t * N + w) * n + i];
            setTarget(c, object, lv);
        }
        for (int s = 0; s < stepsPerTarget; ++s) stepWorld<false>(c, executeControllers != 0, allowSleeping != 0);
    }
}

// Small and medium batches: a world is owned by LPW lanes of a warp (b2g_multi.cuh); `wpw` <= 32 / LPW worlds per warp, one
// warp per CTA.  Warp g owns worlds [g * wpw, (g + 1) * wpw); the lists of each world live in shared memory behind the model.
DEV WorldShared* worldShared(const ModelHeader& hdr, int slot) {
    return reinterpret_cast<WorldShared*>(g_sm + kHdrWords + ((hdr.words + 3) & ~3)) + slot;
}
template <int LPW>
__global__ void __launch_bounds__(32) k_step_multi(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S,
                                                   long long N, StepParams prm, int object, const float* __restrict__ targets,
                                                   int nTargets, int stepsPerTarget, int executeControllers, int allowSleeping, int wpw) {
    StepParams prm1 = prm;   // the sub-warp kernel serves one-model batches only (b2g_capi.cpp)
    prm1.tileVariant = nullptr;
    stageModel(hdr, prm1, blob, 0);
    const int lane = threadIdx.x & 31;
    const int slot = lane / LPW;
    const long long w = (long long)blockIdx.x * wpw + slot;
    const bool live = slot < wpw && w < N;
    const unsigned warpMask = __ballot_sync(0xffffffffu, live);
    if (!live) return;  // whole groups leave: the group barriers only name lanes of one group
    Grp<LPW> g;
    g.warp = warpMask;
    g.s = lane % LPW;
    g.mask = (LPW == 32 ? 0xffffffffu : ((1u << LPW) - 1u)) << (lane - g.s);
    WorldShared& W = *worldShared(hdr, slot);
    if (g.s == 0) W.keyNj = -1;  // no schedule cached yet
    Ctx c;
    B2G_SET_MO(c, 0);
    c.P = worldBase(S, hdr.stateChunks, w);
    for (int t = 0; t < nTargets; ++t) {
        if (targets != nullptr) {
            if (g.s == 0) {
                int n = mI(c, mobj(c, object, MO_NDOFS));
                float lv[kMaxDofsPerObject];
                for (int i = 0; i < n; ++i) lv[i] = targets[((size_t)t * N + w) * n + i];
                setTarget(c, object, lv);
            }
            gsync(g);
        }
        for (int s = 0; s < stepsPerTarget; ++s) stepWorldMulti<LPW>(c, g, W, executeControllers != 0, allowSleeping != 0);
    }
}

// stepPhysics(callback, steps, ...) (:3312-3333): a kernel of its own so that the plain stepping kernel stays lean
__global__ void k_step_guarded(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S, long long N,
                               StepParams prm, int steps, int executeControllers, int allowSleeping) {
    B2G_PROLOGUE();
    for (int s = 0; s < steps; ++s) stepWorldGuarded(c, executeControllers != 0, allowSleeping != 0);
}

// per-world status: the sticky bits kept in the state (bit0: contact capacity exceeded) plus bit1 = a body position or
// velocity is not finite right now
__global__ void k_world_status(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S, long long N,
                               StepParams prm, int* __restrict__ out) {
    B2G_PROLOGUE();
    int st = ldI(c, sslot(c, SS_STATUS)) & 5;
    bool finite = true;
    for (int b = 0; b < H().nb; ++b) {
        const float4 p = ld4(c, bslot(c, b, SB_POS4)), v = ld4(c, bslot(c, b, SB_VEL4));
        finite = finite && isfinite(p.x) && isfinite(p.y) && isfinite(p.z) && isfinite(v.x) && isfinite(v.y) && isfinite(v.z);
    }
    out[w] = st | (finite ? 0 : 2);
}

// Box2DLink::setEnabled (:802-822) for the link with body index `body`: categoryBits of its fixtures (newest first),
// b2Fixture::Refilter = flag the fixture's contacts for filtering and touch its proxy.  enabled == nullptr: `all`.
__global__ void k_set_link_enabled(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S, long long N,
                                   StepParams prm, int body, const unsigned char* __restrict__ enabled, int all) {
    B2G_PROLOGUE();
    const bool on = enabled ? enabled[w] != 0 : all != 0;
    const int start = mI(c, mbody(c, body, MB_FIX_START)), cnt = mI(c, mbody(c, body, MB_FIX_COUNT));
    uint64_t bits = 0ull;
    for (int f = start; f < start + cnt; ++f) bits |= 1ull << f;
    uint64_t d = ldDisabled(c);
    const bool wasOn = (d & bits) == 0ull;   // Box2DLink keeps one flag per link: "if (b_enable == _enabled) return"
    if (on == wasOn) return;
    d = on ? (d & ~bits) : (d | bits);
    stI(c, sslot(c, SS_DISABLED_LO), (int)(uint32_t)(d & 0xffffffffull));
    stI(c, sslot(c, SS_DISABLED_HI), (int)(uint32_t)(d >> 32));
    for (int k = contactCount(c) - 1; k >= 0; --k) {
        int fA, fB;
        contactFixtures(c, k, fA, fB);
        if (((bits >> fA) & 1ull) || ((bits >> fB) & 1ull)) stI(c, cslot(c, k, SC_FLAGS), ldI(c, cslot(c, k, SC_FLAGS)) | CF_FILTER);
    }
    stMoved(c, ldMoved(c) | bits);
}

// Tail of b2Body::SetMassData for dynamic body `body` (Box2DLink::setMass :478-489; the mass data themselves live in the
// model): the sweep centre is re-derived from the body transform, c0 = c, and the linear velocity follows the (round-off
// sized) shift of the centre.
__global__ void k_mass_center(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S, long long N,
                              StepParams prm, int body) {
    B2G_PROLOGUE();
    const V2 oldCenter = ldV2(c, bslot(c, body, SB_CX));
    const V2 cc = mul(ldXF(c, body), localCenter(c, body));
    stV2(c, bslot(c, body, SB_CX), cc);
    stV2(c, bslot(c, body, SB_C0X), cc);
    const float wz = ldF(c, bslot(c, body, SB_W));
    const V2 v = ldV2(c, bslot(c, body, SB_VX));
    stV2(c, bslot(c, body, SB_VX), v + cross(wz, cc - oldCenter));
}

// Box2DWorld::atRest(threshold) (:3435-3452): every object's DOF velocities are within the threshold (infinity norm)
__global__ void k_at_rest(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S, long long N,
                          StepParams prm, float threshold, int* __restrict__ out) {
    B2G_PROLOGUE();
    bool rest = true;
    for (int o = 0; o < H().no; ++o) {
        float qd[kMaxDofsPerObject];
        getDofVelocities(c, o, qd);
        int n = mI(c, mobj(c, o, MO_NDOFS));
        for (int i = 0; i < n; ++i) rest = rest && (fabsf(qd[i]) <= threshold);
    }
    out[w] = rest ? 1 : 0;
}

// Box2DWorld::getObjects(aabb, ...) (:3207-3254) for every world: hits [N][no].  The query box lives in a scratch fixture
// record behind the model blob (index fQuery) so that GJK reads it like any other proxy.
__global__ void k_objects_in_aabb(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S, long long N,
                                  StepParams prm, float lx, float ly, float ux, float uy, int fQuery,
                                  unsigned char* __restrict__ hits) {
    {   // geometry only: every parameter set shares it, so the first blob serves all worlds (one copy per CTA, the scratch
        // record sits behind it)
        StepParams one = prm;
        one.tileVariant = nullptr;
        stageModel(hdr, one, blob, 0);
    }
    if (threadIdx.x == 0) {
        uint32_t* r = g_sm + kHdrWords + hdr.oFix + fQuery * MF_STRIDE;
        const float hx = 0.5f * (ux - lx), hy = 0.5f * (uy - ly);  // b2AABB::GetExtents -> SetAsBox(hx, hy)
        const float vx[4] = {-hx, hx, hx, -hx}, vy[4] = {-hy, -hy, hy, hy};
        r[MF_BODY] = 0;
        r[MF_COUNT] = 4;
        for (int i = 0; i < 4; ++i) {
            r[MF_VERTS + 2 * i] = __float_as_uint(vx[i]);
            r[MF_VERTS + 2 * i + 1] = __float_as_uint(vy[i]);
        }
    }
    __syncthreads();
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x - prm.lead;
    if (w < 0 || w >= N) return;
    Ctx c;
    B2G_SET_MO(c, 0);
    c.P = worldBase(S, hdr.stateChunks, w + prm.worldOffset);
    unsigned char h8[kMaxBodies];
    objectsInAABB(c, lx, ly, ux, uy, fQuery, h8);
    for (int o = 0; o < hdr.no; ++o) hits[(size_t)w * hdr.no + o] = h8[o];
}

// Box2DWorld::isPhysicallyFeasible (:3370-3400) for every world
__global__ void k_physically_feasible(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S,
                                      long long N, StepParams prm, int* __restrict__ out) {
    B2G_PROLOGUE();
    out[w] = physicallyFeasible(c) ? 1 : 0;
}

// guarded stepping: clear the abort flag / executed-step counter of every world
__global__ void k_set_word(float* S, long long N, int stateChunks, int off, int value) {
    long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w < N) *reinterpret_cast<int*>(worldBase(S, stateChunks, w) + off) = value;
}

// Box2DWorld::setToRest (:3550-3558): setDOFVelocities(0) on every object
__global__ void k_set_to_rest(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S, long long N,
                              StepParams prm) {
    B2G_PROLOGUE();
    float zero[kMaxDofsPerObject];
    for (int i = 0; i < kMaxDofsPerObject; ++i) zero[i] = 0.0f;
    for (int o = 0; o < H().no; ++o) objectSetDofVelocities(c, o, zero);
}

// Box2DCollisionChecker::startRecordings (:2758-2765): every contact that is touching and enabled right now goes
// through BeginContact, in contact-list order (newest first)
__global__ void k_record_touching(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S,
                                  long long N, StepParams prm) {
    B2G_PROLOGUE();
    for (int k = contactCount(c) - 1; k >= 0; --k) {
        int flags = ldI(c, cslot(c, k, SC_FLAGS));
        if ((flags & CF_TOUCHING) == 0 || (flags & CF_ENABLED) == 0) continue;
        onBeginContact(c, k);
    }
}

__global__ void k_get_bodies(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S,
                             long long N, StepParams prm, float* __restrict__ out) {
    B2G_PROLOGUE();
    const int nb = H().nb;
    for (int b = 0; b < nb; ++b) {
        float* o = out + ((size_t)w * nb + b) * 16;
        o[0] = ldF(c, bslot(c, b, SB_PX)); o[1] = ldF(c, bslot(c, b, SB_PY));
        o[2] = ldF(c, bslot(c, b, SB_QS)); o[3] = ldF(c, bslot(c, b, SB_QC));
        o[4] = ldF(c, bslot(c, b, SB_CX)); o[5] = ldF(c, bslot(c, b, SB_CY)); o[6] = ldF(c, bslot(c, b, SB_A));
        o[7] = ldF(c, bslot(c, b, SB_C0X)); o[8] = ldF(c, bslot(c, b, SB_C0Y)); o[9] = ldF(c, bslot(c, b, SB_A0));
        o[10] = ldF(c, bslot(c, b, SB_VX)); o[11] = ldF(c, bslot(c, b, SB_VY)); o[12] = ldF(c, bslot(c, b, SB_W));
        o[13] = ldF(c, bslot(c, b, SB_SLEEP));
        o[14] = (ldI(c, bslot(c, b, SB_FLAGS)) & 1) ? 1.0f : 0.0f;
        o[15] = (float)bodyType(c, b);
    }
}

__global__ void k_get_fat(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S,
                          long long N, StepParams prm, float* __restrict__ out) {
    B2G_PROLOGUE();
    const int nf = H().nf;
    for (int f = 0; f < nf; ++f)
        *reinterpret_cast<float4*>(out + ((size_t)w * nf + f) * 4) = ld4(c, fslot(c, f));
}

// Box2DObject::getBallApproximation / Box2DLink::updateBallApproximation (:709-747, 1754-1763): centre = getTransform() * ball
// centre with getTransform() (:193-213) = rotation by the body angle about z plus invScale * GetWorldPoint(_local_origin_offset);
// the radius passes through.  balls: rows of 8 = body, origin offset x y, centre x y z, radius (HostModel::balls).
__global__ void k_get_balls(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S, long long N,
                            StepParams prm, const float* __restrict__ balls, int first, int count, float* __restrict__ out) {
    B2G_PROLOGUE();
    const float inv = H().invScale;
    for (int i = 0; i < count; ++i) {
        const float* r = balls + (size_t)(first + i) * 8;
        const XF xf = ldXF(c, (int)r[0]);
        const V2 pos = mul(xf, mk(r[1], r[2]));
        const float m00 = xf.q.c, m01 = -xf.q.s, m10 = -m01, m11 = m00;
        const float tx = inv * pos.x, ty = inv * pos.y;
        float4 o;
        o.x = ((m00 * r[3] + m01 * r[4]) + 0.0f * r[5]) + tx;  // Eigen Affine3f * Vector3f: linear part, then translation
        o.y = ((m10 * r[3] + m11 * r[4]) + 0.0f * r[5]) + ty;
        o.z = ((0.0f * r[3] + 0.0f * r[4]) + 1.0f * r[5]) + 0.0f;
        o.w = r[6];
        *reinterpret_cast<float4*>(out + ((size_t)w * count + i) * 4) = o;
    }
}

// Box2DLink::getPose (:215-224), getVelocityVector (:233-241) and getCenterOfMass (:595-603) of every link:
// out[w][body][8] = pose x y theta, velocity x y omega, centre of mass x y (user units); the ground's row is left alone
__global__ void k_get_links(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S, long long N,
                            StepParams prm, float* __restrict__ out) {
    B2G_PROLOGUE();
    const int nb = H().nb;
    const float inv = H().invScale;
    for (int b = 1; b < nb; ++b) {
        const int o = mI(c, mbody(c, b, MB_OBJECT));
        const bool base = mI(c, mobj(c, o, MO_BASE_BODY)) == b;  // only base links carry a _local_origin_offset (:653-671)
        const V2 off = base ? mk(mF(c, mobj(c, o, MO_ORIGIN_X)), mF(c, mobj(c, o, MO_ORIGIN_Y))) : mk(0.0f, 0.0f);
        const V2 pos = mul(ldXF(c, b), off);
        const V2 v = ldV2(c, bslot(c, b, SB_VX));
        const V2 cc = ldV2(c, bslot(c, b, SB_CX));
        float4* r = reinterpret_cast<float4*>(out + ((size_t)w * nb + b) * 8);
        r[0] = make_float4(inv * pos.x, inv * pos.y, ldF(c, bslot(c, b, SB_A)), inv * v.x);
        r[1] = make_float4(inv * v.y, ldF(c, bslot(c, b, SB_W)), inv * cc.x, inv * cc.y);
    }
}

__global__ void k_get_joints(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S,
                             long long N, StepParams prm, int* __restrict__ iout, float* __restrict__ fout) {
    B2G_PROLOGUE();
    const int nj = H().nj;
    for (int j = 0; j < nj; ++j) {
        int* io = iout + ((size_t)w * nj + j) * 6;
        float* fo = fout + ((size_t)w * nj + j) * 15;
        int type = mI(c, mjoint(c, j, MJ_TYPE));
        int flags = ldI(c, jslot(c, j, SJ_FLAGS));
        io[0] = type; io[1] = mI(c, mjoint(c, j, MJ_BODYA)); io[2] = mI(c, mjoint(c, j, MJ_BODYB));
        io[3] = flags & 3; io[4] = (flags >> 2) & 1; io[5] = mI(c, mjoint(c, j, MJ_FLAGS)) & 1;
        fo[0] = ldF(c, jslot(c, j, SJ_IX)); fo[1] = ldF(c, jslot(c, j, SJ_IY)); fo[2] = ldF(c, jslot(c, j, SJ_IZ));
        fo[3] = type == 0 ? 0.0f : ldF(c, jslot(c, j, SJ_MOTOR_IMPULSE));
        fo[4] = ldF(c, jslot(c, j, SJ_MAX_MOTOR_TORQUE)); fo[5] = ldF(c, jslot(c, j, SJ_MOTOR_SPEED));
        fo[6] = type == 2 ? 0.0f : mF(c, mjoint(c, j, MJ_MAXFORCE)); fo[7] = mF(c, mjoint(c, j, MJ_MAXTORQUE));
        fo[8] = mF(c, mjoint(c, j, MJ_LAX)); fo[9] = mF(c, mjoint(c, j, MJ_LAY));
        fo[10] = mF(c, mjoint(c, j, MJ_LBX)); fo[11] = mF(c, mjoint(c, j, MJ_LBY));
        fo[12] = mF(c, mjoint(c, j, MJ_REF)); fo[13] = mF(c, mjoint(c, j, MJ_LOWER)); fo[14] = mF(c, mjoint(c, j, MJ_UPPER));
    }
}

__global__ void k_get_contacts(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S,
                               long long N, StepParams prm, int* __restrict__ counts, int* __restrict__ iout,
                               float* __restrict__ fout, int maxPer) {
    B2G_PROLOGUE();
    int n = contactCount(c);
    counts[w] = n;
    for (int k = 0; k < n && k < maxPer; ++k) {
        int* io = iout + ((size_t)w * maxPer + k) * 8;
        float* fo = fout + ((size_t)w * maxPer + k) * 12;
        int fA, fB;
        contactFixtures(c, k, fA, fB);
        int flags = ldI(c, cslot(c, k, SC_FLAGS));
        int man = ldI(c, cslot(c, k, SC_MANIFOLD));
        io[0] = fA; io[1] = fB; io[2] = flags & CF_TOUCHING ? 1 : 0; io[3] = flags & CF_ENABLED ? 1 : 0;
        io[4] = (man >> 8) & 0xff; io[5] = man & 0xff;
        io[6] = ldI(c, cslot(c, k, SC_ID0)); io[7] = ldI(c, cslot(c, k, SC_ID1));
        const int ck = cslot(c, k, 0);
        fo[0] = ldF(c, ck + SC_LNX); fo[1] = ldF(c, ck + SC_LNY); fo[2] = ldF(c, ck + SC_LPX); fo[3] = ldF(c, ck + SC_LPY);
        fo[4] = ldF(c, ck + SC_P0X); fo[5] = ldF(c, ck + SC_P0Y); fo[6] = ldF(c, ck + SC_N0); fo[7] = ldF(c, ck + SC_T0);
        fo[8] = ldF(c, ck + SC_P1X); fo[9] = ldF(c, ck + SC_P1Y); fo[10] = ldF(c, ck + SC_N1); fo[11] = ldF(c, ck + SC_T1);
    }
}

// Box2DCollisionChecker::updateContactCache (:2880-2911): UpdateContacts() = FindNewContacts + Collide (assumed
// meaning of the fork-only call at :2898), then the touching+enabled contacts in contact-list order
__global__ void k_check_collision(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S,
                                  long long N, StepParams prm, int* __restrict__ counts, int* __restrict__ iout,
                                  float* __restrict__ fout, int maxPer) {
    B2G_PROLOGUE();
    findNewContacts(c);
    collide(c);
    int n = 0;
    for (int k = contactCount(c) - 1; k >= 0; --k) {
        int flags = ldI(c, cslot(c, k, SC_FLAGS));
        if ((flags & CF_TOUCHING) == 0 || (flags & CF_ENABLED) == 0) continue;
        if (n < maxPer) {
            int bA, bB, fA, fB;
            contactBodies(c, k, bA, bB, fA, fB);
            // b2Contact::GetWorldManifold with the bodies' current transforms; only points[0] is reported (:2857-2862)
            reportContact(c, ldI(c, cslot(c, k, SC_MANIFOLD)) & 0xff, ldV2(c, cslot(c, k, SC_LNX)), ldV2(c, cslot(c, k, SC_LPX)),
                          ldV2(c, cslot(c, k, SC_P0X)), ldXF(c, bA), ldXF(c, bB), bA, bB, fA, fB,
                          iout + ((size_t)w * maxPer + n) * 4, fout + ((size_t)w * maxPer + n) * 6);
        }
        ++n;
    }
    counts[w] = n;
}

__global__ void k_status(float* S, long long N, int stateChunks, int statusOff, int* __restrict__ out) {
    long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w < N) out[w] = *reinterpret_cast<const int*>(worldBase(S, stateChunks, w) + statusOff);
}

// gather / scatter of the world state vector and static poses for the save/restore stack
// ---------------------------------------------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------------------------------------------
// batches may be driven from different host threads (one stream per batch): the launch counter is shared
#if !B2G_SETS
// ---- regrouping worlds by island shape (large batches, b2g_capi.cpp: regroupedRollout) ----------------------------------------
// The worlds of a warp execute the union of their paths: a world whose box sits in the robot's island walks a different joint
// list, runs the contact solver and more position iterations, and every other lane of its warp waits.  Between the target
// segments of a rollout the worlds are therefore SORTED by a key that predicts their path over the next steps -- live contacts
// or not, which robot link a touching contact enters the robot through, whether a static body is involved (TOI candidates) --
// and their state moves to the slot the sort assigns (a second state array; slot != world until the rollout ends, the
// controller targets of a segment are gathered into slot order).  Worlds are independent: results are bit-identical.
constexpr int kRegroupClasses = 32;

__global__ void k_regroup_key(const __grid_constant__ ModelHeader hdr, const uint32_t* __restrict__ blob, float* S, long long N,
                              StepParams prm, unsigned char* __restrict__ key, int* __restrict__ count, int groupSize) {
    B2G_PROLOGUE();
    const int n = contactCount(c);
    int entry = 0, staticLive = 0, live = 0;
    for (int k = 0; k < n; ++k) {
        const int fl = ldI(c, cslot(c, k, SC_FLAGS));
        if ((fl & CF_ENABLED) == 0) continue;
        int fA, fB;
        contactFixtures(c, k, fA, fB);
        const int bA = mI(c, mfix(c, fA, MF_BODY)), bB = mI(c, mfix(c, fB, MF_BODY));
        live = 1;
        if (bodyType(c, bA) == 0 || bodyType(c, bB) == 0) staticLive = 1;
        if (fl & CF_TOUCHING) {
            const int lowBody = bA < bB ? bA : bB;   // the older body: for robot-vs-object contacts the robot link the DFS enters through
            const int e = 1 + (lowBody % 13);
            entry = entry == 0 || e < entry ? e : entry;
        }
    }
    // 0: no contact at all; 1: live contacts, none touching; 2..14: touching, by entry link; + 16 with a static body involved
    const int cls = (live ? (entry ? 1 + entry : 1) : 0) + (staticLive ? 16 : 0);
    key[w] = (unsigned char)cls;
    // worlds are sorted inside groups of `groupSize` consecutive slots (one CTA of the stepping kernel): every SM keeps its mix
    // of cheap and expensive worlds, only the warps inside it become homogeneous
    const int slot = (int)(w / groupSize) * kRegroupClasses + cls;
    const unsigned m = __match_any_sync(__activemask(), slot);
    if ((int)(threadIdx.x & 31) == __ffs((int)m) - 1) atomicAdd(&count[slot], __popc(m));
}

// exclusive scan of the class counts: cursor[c] = first slot of class c (classes in ascending order)
__global__ void k_regroup_scan(const int* __restrict__ count, int* __restrict__ cursor, int groups, int groupSize) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) return;
    int s = g * groupSize;
    for (int i = 0; i < kRegroupClasses; ++i) {
        cursor[g * kRegroupClasses + i] = s;
        s += count[g * kRegroupClasses + i];
    }
}

__global__ void k_regroup_assign(const unsigned char* __restrict__ key, int* __restrict__ cursor, int* __restrict__ newSlot, long long N,
                                 int groupSize) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= N) return;
    const int cls = (int)(w / groupSize) * kRegroupClasses + key[w];
    const unsigned m = __match_any_sync(__activemask(), cls);
    const int lane = threadIdx.x & 31, leader = __ffs((int)m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&cursor[cls], __popc(m));
    base = __shfl_sync(m, base, leader);
    newSlot[w] = base + __popc(m & ((1u << lane) - 1u));
}

__global__ void k_regroup_iota(int* __restrict__ a, long long N) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w < N) a[w] = (int)w;
}

// state of slot s of `src` -> slot newSlot[s] of `dst`: everything but the unused contact slots and the contact scratch (dead
// between steps); worldOfDst[newSlot[s]] = worldOfSrc[s]
__global__ void k_regroup_move(const __grid_constant__ ModelHeader hdr, const float4* __restrict__ src, float4* __restrict__ dst,
                               const int* __restrict__ newSlot, const int* __restrict__ worldOfSrc, int* __restrict__ worldOfDst,
                               long long N) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    const long long d = newSlot[s];
    const size_t tileChunks = (size_t)hdr.stateChunks * kTileWorlds;
    const float4* ps = src + (size_t)(s >> kTileShift) * tileChunks + (size_t)(s & (kTileWorlds - 1));
    float4* pd = dst + (size_t)(d >> kTileShift) * tileChunks + (size_t)(d & (kTileWorlds - 1));
    for (int k = 0; k < hdr.cCont; ++k) pd[(size_t)k * kTileWorlds] = ps[(size_t)k * kTileWorlds];
    const float4 sc0 = ps[(size_t)hdr.cScalar * kTileWorlds];
    const int nc = __float_as_int(sc0.z);   // SS_NCONTACTS
    for (int k = hdr.cCont; k < hdr.cCont + nc * SC_CHUNKS; ++k) pd[(size_t)k * kTileWorlds] = ps[(size_t)k * kTileWorlds];
    for (int k = hdr.cScalar; k < hdr.stateChunks; ++k) pd[(size_t)k * kTileWorlds] = ps[(size_t)k * kTileWorlds];
    worldOfDst[d] = worldOfSrc[s];
}

// controller targets of one segment, world order -> slot order
__global__ void k_regroup_targets(const float* __restrict__ targets, float* __restrict__ out, const int* __restrict__ worldOf, long long N, int nd) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * nd) return;
    const long long s = i / nd;
    out[i] = targets[(size_t)worldOf[s] * nd + (i - s * nd)];
}
#endif

}  // namespace B2G_VNS
#if B2G_SETS
extern std::atomic<unsigned long long> g_launches;
#else
std::atomic<unsigned long long> g_launches{0};
#endif
namespace B2G_VNS {
#if defined(B2G_PHASE_TIMING) && !B2G_SETS
extern "C" int b2g_debug_phase_cycles(unsigned long long* out32, int reset) {
    if (out32 && cudaMemcpyFromSymbol(out32, g_phaseCycles, sizeof(unsigned long long) * 32) != cudaSuccess) return 1;
    if (reset) {
        unsigned long long zero[32] = {0};
        if (cudaMemcpyToSymbol(g_phaseCycles, zero, sizeof(zero)) != cudaSuccess) return 1;
    }
    return 0;
}
#endif
}  // namespace B2G_VNS
#if !B2G_SETS
unsigned long long launchCount() { return g_launches.load(std::memory_order_relaxed); }
#endif
namespace B2G_VNS {

static size_t smemPad() {  // tuning knob: extra dynamic shared memory per CTA throttles occupancy (B2G_SMEM_PAD bytes)
    static long pad = -1;
    if (pad < 0) {
        const char* e = getenv("B2G_SMEM_PAD");
        pad = e ? atol(e) : 0;
    }
    return (size_t)pad;
}
// dynamic shared memory of a CTA of `threads` threads: header + parameters + one blob, or one blob per warp when the batch holds
// several parameter sets (stageModel)
static inline size_t smemBytes(const Launch& L, int threads) {
    const size_t copies = L.prm.tileVariant ? (size_t)((threads + 31) / 32) : 1;
    const size_t blob = L.prm.tileVariant ? (size_t)L.prm.blobStride : (size_t)L.h.words;
    return (size_t)kHdrWords * 4 + copies * blob * 4 + smemPad();
}
static inline size_t smemBytes(const ModelHeader& h) { return (size_t)kHdrWords * 4 + (size_t)h.words * 4 + smemPad(); }

static cudaError_t prepare(const void* fn, size_t bytes) {
    if (bytes > 227 * 1024) return cudaErrorInvalidConfiguration;
    if (bytes > 48 * 1024) return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    return cudaSuccess;
}
static cudaError_t prepare(const void* fn, const ModelHeader& h) { return prepare(fn, smemBytes(h)); }

#define LAUNCH(kernel, ...)                                                                         \
    do {                                                                                            \
        int threads = L.block;                                                                      \
        const size_t bytes = smemBytes(L, threads);                                                 \
        cudaError_t e = prepare((const void*)kernel, bytes);                                        \
        if (e != cudaSuccess) return e;                                                             \
        unsigned blocks = (unsigned)((L.N + L.prm.lead + threads - 1) / threads);                   \
        kernel<<<blocks, threads, bytes, L.stream>>>(L.h, L.blob, L.S, L.N, L.prm, ##__VA_ARGS__);  \
        ++g_launches;                                                                               \
        return cudaGetLastError();                                                                  \
    } while (0)

cudaError_t launchInit(const Launch& L) { LAUNCH(k_init); }
cudaError_t launchPark(const Launch& L) { LAUNCH(k_park); }
cudaError_t launchSetState(const Launch& L, const float* q, const float* qd) { LAUNCH(k_set_state, q, qd); }
cudaError_t launchGetState(const Launch& L, float* q, float* qd) { LAUNCH(k_get_state, q, qd); }
cudaError_t launchSetPose(const Launch& L, int object, const float* pose) { LAUNCH(k_set_pose, object, pose); }
cudaError_t launchGetPose(const Launch& L, int object, float* pose) { LAUNCH(k_get_pose, object, pose); }
cudaError_t launchSetObjectDofs(const Launch& L, int object, const DofSel& sel, const float* values, int velocities) {
    LAUNCH(k_set_object_dofs, object, sel, values, velocities);
}
cudaError_t launchGetObjectDofs(const Launch& L, int object, const DofSel& sel, float* out, int velocities) {
    LAUNCH(k_get_object_dofs, object, sel, out, velocities);
}
cudaError_t launchSetEfforts(const Launch& L, int object, const float* v) { LAUNCH(k_set_efforts, object, v); }
cudaError_t launchSetTarget(const Launch& L, int object, const float* v) { LAUNCH(k_set_target, object, v); }
cudaError_t launchSetTargetPosition(const Launch& L, int object, const float* q, float kp) { LAUNCH(k_set_target_position, object, q, kp); }
cudaError_t launchStep(const Launch& L, int object, const float* targets, int nTargets, int stepsPerTarget, int execCtrl,
                       int allowSleep) {
    if (L.stepLpw > 1) {  // small / medium batch: LPW lanes per world, stepLanes worlds per warp, one warp per CTA
        const size_t bytes = (((size_t)kHdrWords + ((L.h.words + 3) & ~3)) * 4) + (size_t)L.stepLanes * sizeof(WorldShared) + smemPad();
        const unsigned blocks = (unsigned)((L.N + L.stepLanes - 1) / L.stepLanes);
#define B2G_MULTI(LPW)                                                                                                          \
    {                                                                                                                            \
        if (bytes > 48 * 1024) {                                                                                                 \
            cudaError_t e = cudaFuncSetAttribute((const void*)k_step_multi<LPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); \
            if (e != cudaSuccess) return e;                                                                                      \
        }                                                                                                                        \
        k_step_multi<LPW><<<blocks, 32, bytes, L.stream>>>(L.h, L.blob, L.S, L.N, L.prm, object, targets, nTargets, stepsPerTarget, \
                                                           execCtrl, allowSleep, L.stepLanes);                                  \
    }
        if (L.stepLpw == 2) B2G_MULTI(2)
        else if (L.stepLpw == 4) B2G_MULTI(4)
        else if (L.stepLpw == 8) B2G_MULTI(8)
        else if (L.stepLpw == 16) B2G_MULTI(16)
        else B2G_MULTI(32)
#undef B2G_MULTI
        ++g_launches;
        return cudaGetLastError();
    }
    if (L.stepLanes < 32) {  // small batch: narrow warps, one warp per CTA
        const unsigned warps = (unsigned)((L.N + L.stepLanes - 1) / L.stepLanes);
        static const int wpc = [] { const char* e = getenv("B2G_NARROW_WARPS"); int v = e ? atoi(e) : 1; return v >= 1 && v <= 32 ? v : 1; }();
        const unsigned blocks = (warps + wpc - 1) / wpc;  // warps per CTA: tuning experiments only (default one warp per CTA)
        const size_t bytes = smemBytes(L, 32 * wpc);
        cudaError_t e = prepare((const void*)k_step_narrow, bytes);
        if (e != cudaSuccess) return e;
        k_step_narrow<<<blocks, 32 * wpc, bytes, L.stream>>>(L.h, L.blob, L.S, L.N, L.prm, object, targets, nTargets,
                                                           stepsPerTarget, execCtrl, allowSleep, L.stepLanes);
        ++g_launches;
        return cudaGetLastError();
    }
    int threads = L.stepBlock;
    // several parameter sets: one blob copy per warp has to fit the SM's shared memory (large scenes get smaller CTAs)
    while (threads > 32 && smemBytes(L, threads) > 200 * 1024) threads -= 32;
    const size_t bytes = smemBytes(L, threads);
#if B2G_SETS
    if (threads > 640) threads = 640;   // the parameter-set build carries the 96-register kernel only
#else
    if (threads > 640) {
        cudaError_t e = prepare((const void*)k_step<896>, bytes);
        if (e != cudaSuccess) return e;
        unsigned blocks = (unsigned)((L.N + threads - 1) / threads);  // every thread has a (possibly padding) world
        k_step<896><<<blocks, threads, bytes, L.stream>>>(L.h, L.blob, L.S, L.N, L.prm, object, targets, nTargets, stepsPerTarget,
                                                       execCtrl, allowSleep);
        ++g_launches;
        return cudaGetLastError();
    }
#endif
    const size_t bytes640 = smemBytes(L, threads);
    cudaError_t e = prepare((const void*)k_step<640>, bytes640);
    if (e != cudaSuccess) return e;
    unsigned blocks = (unsigned)((L.N + threads - 1) / threads);  // every thread has a (possibly padding) world
    k_step<640><<<blocks, threads, bytes640, L.stream>>>(L.h, L.blob, L.S, L.N, L.prm, object, targets, nTargets, stepsPerTarget,
                                                      execCtrl, allowSleep);
    ++g_launches;
    return cudaGetLastError();
}
cudaError_t launchSetToRest(const Launch& L) { LAUNCH(k_set_to_rest); }
cudaError_t launchSetLinkEnabled(const Launch& L, int body, const unsigned char* enabled, int all) {
    LAUNCH(k_set_link_enabled, body, enabled, all);
}
cudaError_t launchMassCenter(const Launch& L, int body) { LAUNCH(k_mass_center, body); }
cudaError_t launchAtRest(const Launch& L, float threshold, int* out) { LAUNCH(k_at_rest, threshold, out); }
cudaError_t launchPhysicallyFeasible(const Launch& L, int* out) { LAUNCH(k_physically_feasible, out); }
cudaError_t launchObjectsInAABB(const Launch& L, float lx, float ly, float ux, float uy, unsigned char* hits) {
    // scratch fixture record: first MF_STRIDE-aligned slot (relative to oFix) behind the blob
    const int fQuery = (L.h.words - L.h.oFix + MF_STRIDE - 1) / MF_STRIDE;
    const size_t bytes = ((size_t)kHdrWords + L.h.oFix + (size_t)(fQuery + 1) * MF_STRIDE) * 4 + smemPad();
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute((const void*)k_objects_in_aabb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return e;
    }
    int threads = L.block;
    unsigned blocks = (unsigned)((L.N + L.prm.lead + threads - 1) / threads);
    k_objects_in_aabb<<<blocks, threads, bytes, L.stream>>>(L.h, L.blob, L.S, L.N, L.prm, lx, ly, ux, uy, fQuery, hits);
    ++g_launches;
    return cudaGetLastError();
}
cudaError_t launchRecordTouching(const Launch& L) { LAUNCH(k_record_touching); }
cudaError_t launchStepGuarded(const Launch& L, int steps, int execCtrl, int allowSleep) {
    LAUNCH(k_step_guarded, steps, execCtrl, allowSleep);
}
cudaError_t launchGetBodies(const Launch& L, float* out) { LAUNCH(k_get_bodies, out); }
cudaError_t launchGetFat(const Launch& L, float* out) { LAUNCH(k_get_fat, out); }
cudaError_t launchGetJoints(const Launch& L, int* iout, float* fout) { LAUNCH(k_get_joints, iout, fout); }
cudaError_t launchGetBalls(const Launch& L, const float* balls, int first, int count, float* out) {
    LAUNCH(k_get_balls, balls, first, count, out);
}
cudaError_t launchGetLinks(const Launch& L, float* out) { LAUNCH(k_get_links, out); }
cudaError_t launchGetContacts(const Launch& L, int* counts, int* iout, float* fout, int maxPer) {
    LAUNCH(k_get_contacts, counts, iout, fout, maxPer);
}
cudaError_t launchCheckCollision(const Launch& L, int* counts, int* iout, float* fout, int maxPer) {
    LAUNCH(k_check_collision, counts, iout, fout, maxPer);
}
cudaError_t launchBroadcast(const Launch& L, const float* tmpl) {  // worlds [prm.worldOffset, prm.worldOffset + N)
    const long long first = L.prm.worldOffset, count = L.N;
    long long total = (((first + count + kTileWorlds - 1) >> kTileShift) - (first >> kTileShift)) * L.h.stateChunks * kTileWorlds;
    int threads = 256;
    long long want = (total + threads - 1) / threads;
    unsigned blocks = (unsigned)(want < 4096 ? want : 4096);  // grid-stride copy: a few CTAs per SM saturate HBM
    k_broadcast<<<blocks, threads, 0, L.stream>>>(reinterpret_cast<float4*>(L.S), reinterpret_cast<const float4*>(tmpl), first, count,
                                                 L.h.stateChunks);
    ++g_launches;
    return cudaGetLastError();
}
cudaError_t launchExtractWorld(const Launch& L, float* tmpl, long long world) {
    k_extract_world<<<(L.h.stateChunks + 255) / 256, 256, 0, L.stream>>>(reinterpret_cast<const float4*>(L.S),
                                                                        reinterpret_cast<float4*>(tmpl), L.h.stateChunks, world);
    ++g_launches;
    return cudaGetLastError();
}
cudaError_t launchClearAbort(const Launch& L) {
    int threads = 256;
    k_set_word<<<(unsigned)((L.N + threads - 1) / threads), threads, 0, L.stream>>>(L.S, L.N, L.h.stateChunks,
                                                                                    L.h.cScalar * kChunkBytes + SS_ABORT, 0);
    ++g_launches;
    return cudaGetLastError();
}
cudaError_t launchAbortFlags(const Launch& L, int* out) {
    int threads = 256;
    k_status<<<(unsigned)((L.N + threads - 1) / threads), threads, 0, L.stream>>>(L.S, L.N, L.h.stateChunks,
                                                                                  L.h.cScalar * kChunkBytes + SS_ABORT, out);
    ++g_launches;
    return cudaGetLastError();
}
cudaError_t launchStatus(const Launch& L, int* out) { LAUNCH(k_world_status, out); }

#if !B2G_SETS
cudaError_t launchRegroupKey(const Launch& L, unsigned char* key, int* count, int groupSize) {
    const long long groups = (L.N + groupSize - 1) / groupSize;
    cudaError_t e = cudaMemsetAsync(count, 0, sizeof(int) * kRegroupClasses * groups, L.stream);
    if (e != cudaSuccess) return e;
    LAUNCH(k_regroup_key, key, count, groupSize);
}
cudaError_t launchRegroupAssign(const Launch& L, const unsigned char* key, const int* count, int* cursor, int* newSlot, int groupSize) {
    const int groups = (int)((L.N + groupSize - 1) / groupSize);
    k_regroup_scan<<<(unsigned)((groups + 127) / 128), 128, 0, L.stream>>>(count, cursor, groups, groupSize);
    const int threads = 256;
    k_regroup_assign<<<(unsigned)((L.N + threads - 1) / threads), threads, 0, L.stream>>>(key, cursor, newSlot, L.N, groupSize);
    g_launches += 2;
    return cudaGetLastError();
}
cudaError_t launchRegroupIota(const Launch& L, int* a) {
    const int threads = 256;
    k_regroup_iota<<<(unsigned)((L.N + threads - 1) / threads), threads, 0, L.stream>>>(a, L.N);
    ++g_launches;
    return cudaGetLastError();
}
cudaError_t launchRegroupMove(const Launch& L, const float* src, float* dst, const int* newSlot, const int* worldOfSrc, int* worldOfDst) {
    const int threads = 128;
    k_regroup_move<<<(unsigned)((L.N + threads - 1) / threads), threads, 0, L.stream>>>(L.h, (const float4*)src, (float4*)dst, newSlot,
                                                                                       worldOfSrc, worldOfDst, L.N);
    ++g_launches;
    return cudaGetLastError();
}
cudaError_t launchRegroupTargets(const Launch& L, const float* targets, float* out, const int* worldOf, int nd) {
    const int threads = 256;
    const long long total = L.N * nd;
    k_regroup_targets<<<(unsigned)((total + threads - 1) / threads), threads, 0, L.stream>>>(targets, out, worldOf, L.N, nd);
    ++g_launches;
    return cudaGetLastError();
}
#endif
}  // namespace B2G_VNS
}  // namespace b2g
