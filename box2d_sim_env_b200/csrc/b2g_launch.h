// Host-visible launch interface between the C ABI (b2g_capi.cu) and the kernels (b2g_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include "model.h"

namespace b2g {

struct Launch {
    ModelHeader h;
    const uint32_t* blob;  // device copy of the model blob
    float* S;              // device state, [stateWords][N]
    long long N;
    StepParams prm;
    cudaStream_t stream;
    int block;             // threads per CTA
    int stepBlock;         // threads per CTA of k_step (power of two <= kWorldPad)
    int stepLanes;         // worlds per warp of the stepping kernel: 32, or fewer for small batches
    int stepLpw;           // lanes per world: 1 = one thread per world (k_step / k_step_narrow); 2, 4, 8 = k_step_multi
};

// a selection of object-local DOF indices (ascending) passed to the per-object DOF kernels by value
struct DofSel {
    int n;
    int idx[kMaxDofsPerObject];
};

unsigned long long launchCount();

// The kernels are compiled twice (build.py): namespace `one` for batches whose worlds share one model (the model blob is
// staged once per CTA, model reads carry no per-warp offset: the hot path is exactly the single-model code), namespace `sets`
// for batches that hold several parameter sets (every warp stages the blob variant of its tile, Ctx::mo).  The C ABI calls
// the dispatching wrappers below: a launch goes to `sets` when the batch has a tile table.
//   launchPark parks worlds [prm.worldOffset, prm.worldOffset + N): the padding worlds
//   launchGetLinks: [N][nb][8] link pose / velocity / centre of mass
//   launchMassCenter: b2Body::SetMassData's sweep-centre update (setMass)
//   launchAbortFlags: out[w] = abort flag (bit 0) | executed steps << 8
#define B2G_LAUNCH_DECLS \
    cudaError_t launchInit(const Launch& L); \
    cudaError_t launchPark(const Launch& L); \
    cudaError_t launchBroadcast(const Launch& L, const float* tmpl); \
    cudaError_t launchExtractWorld(const Launch& L, float* tmpl, long long world); \
    cudaError_t launchSetState(const Launch& L, const float* q, const float* qd); \
    cudaError_t launchGetState(const Launch& L, float* q, float* qd); \
    cudaError_t launchSetPose(const Launch& L, int object, const float* pose); \
    cudaError_t launchGetPose(const Launch& L, int object, float* pose); \
    cudaError_t launchSetObjectDofs(const Launch& L, int object, const DofSel& sel, const float* values, int velocities); \
    cudaError_t launchGetObjectDofs(const Launch& L, int object, const DofSel& sel, float* out, int velocities); \
    cudaError_t launchSetEfforts(const Launch& L, int object, const float* v); \
    cudaError_t launchSetTarget(const Launch& L, int object, const float* v); \
    cudaError_t launchSetTargetPosition(const Launch& L, int object, const float* q, float kp); \
    cudaError_t launchStep(const Launch& L, int object, const float* targets, int nTargets, int stepsPerTarget, int execCtrl, \
    int allowSleep); \
    cudaError_t launchSetToRest(const Launch& L); \
    cudaError_t launchGetLinks(const Launch& L, float* out); \
    cudaError_t launchMassCenter(const Launch& L, int body); \
    cudaError_t launchSetLinkEnabled(const Launch& L, int body, const unsigned char* enabled, int all); \
    cudaError_t launchAtRest(const Launch& L, float threshold, int* out); \
    cudaError_t launchPhysicallyFeasible(const Launch& L, int* out); \
    cudaError_t launchObjectsInAABB(const Launch& L, float lx, float ly, float ux, float uy, unsigned char* hits); \
    cudaError_t launchRecordTouching(const Launch& L); \
    cudaError_t launchStepGuarded(const Launch& L, int steps, int execCtrl, int allowSleep); \
    cudaError_t launchGetBodies(const Launch& L, float* out); \
    cudaError_t launchGetFat(const Launch& L, float* out); \
    cudaError_t launchGetJoints(const Launch& L, int* iout, float* fout); \
    cudaError_t launchGetBalls(const Launch& L, const float* balls, int first, int count, float* out); \
    cudaError_t launchGetContacts(const Launch& L, int* counts, int* iout, float* fout, int maxPer); \
    cudaError_t launchCheckCollision(const Launch& L, int* counts, int* iout, float* fout, int maxPer); \
    cudaError_t launchStatus(const Launch& L, int* out); \
    cudaError_t launchAbortFlags(const Launch& L, int* out); \
    cudaError_t launchClearAbort(const Launch& L);

namespace one {
B2G_LAUNCH_DECLS
}
namespace sets {
B2G_LAUNCH_DECLS
}
// regrouping worlds by island shape between the target segments of a large rollout (single-model batches only)
namespace one {
cudaError_t launchRegroupKey(const Launch& L, unsigned char* key, int* count, int groupSize);
cudaError_t launchRegroupAssign(const Launch& L, const unsigned char* key, const int* count, int* cursor, int* newSlot, int groupSize);
cudaError_t launchRegroupIota(const Launch& L, int* a);
cudaError_t launchRegroupMove(const Launch& L, const float* src, float* dst, const int* newSlot, const int* worldOfSrc, int* worldOfDst);
cudaError_t launchRegroupTargets(const Launch& L, const float* targets, float* out, const int* worldOf, int nd);
}
#define B2G_DISPATCH(name)                                                                     \
    template <class... A>                                                                      \
    inline cudaError_t name(const Launch& L, A... a) {                                         \
        return L.prm.tileVariant ? sets::name(L, a...) : one::name(L, a...);                   \
    }
B2G_DISPATCH(launchInit)
B2G_DISPATCH(launchPark)
B2G_DISPATCH(launchBroadcast)
B2G_DISPATCH(launchExtractWorld)
B2G_DISPATCH(launchSetState)
B2G_DISPATCH(launchGetState)
B2G_DISPATCH(launchSetPose)
B2G_DISPATCH(launchGetPose)
B2G_DISPATCH(launchSetObjectDofs)
B2G_DISPATCH(launchGetObjectDofs)
B2G_DISPATCH(launchSetEfforts)
B2G_DISPATCH(launchSetTarget)
B2G_DISPATCH(launchSetTargetPosition)
B2G_DISPATCH(launchStep)
B2G_DISPATCH(launchSetToRest)
B2G_DISPATCH(launchGetLinks)
B2G_DISPATCH(launchMassCenter)
B2G_DISPATCH(launchSetLinkEnabled)
B2G_DISPATCH(launchAtRest)
B2G_DISPATCH(launchPhysicallyFeasible)
B2G_DISPATCH(launchObjectsInAABB)
B2G_DISPATCH(launchRecordTouching)
B2G_DISPATCH(launchStepGuarded)
B2G_DISPATCH(launchGetBodies)
B2G_DISPATCH(launchGetFat)
B2G_DISPATCH(launchGetJoints)
B2G_DISPATCH(launchGetBalls)
B2G_DISPATCH(launchGetContacts)
B2G_DISPATCH(launchCheckCollision)
B2G_DISPATCH(launchStatus)
B2G_DISPATCH(launchAbortFlags)
B2G_DISPATCH(launchClearAbort)
#undef B2G_DISPATCH

}  // namespace b2g
