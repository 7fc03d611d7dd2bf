// C ABI of the batched backend (include/b2g.h).  Host-side plumbing only: handles, device buffers, copies and
// kernel launches.  There is deliberately no CPU implementation of any compute entry point: without a CUDA
// device every one of them fails with B2G_ERR_CUDA.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/b2g.h"
#include "b2g_launch.h"
#include "desc.h"
#include "model_build.h"

using namespace b2g;

struct b2g_env {
    EnvironmentDescription d;
};
struct b2g_model {
    EnvironmentDescription env;
    HostModel m;
};

namespace {
thread_local std::string g_error;

b2g_status fail(b2g_status code, const std::string& msg) {
    g_error = msg;
    return code;
}
b2g_status cudaFail(cudaError_t e, const char* what) {
    return fail(B2G_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(call)                                        \
    do {                                                \
        cudaError_t e__ = (call);                       \
        if (e__ != cudaSuccess) return cudaFail(e__, #call); \
    } while (0)

// Every entry point works on its batch's device and leaves the calling thread's current device as it found it (callers
// that share the process -- torch tensors handed to the *_dev entry points -- keep allocating where they were).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err;
    explicit DeviceGuard(int device) {
        err = cudaGetDevice(&prev);
        if (err != cudaSuccess) { prev = -1; return; }
        if (prev != device) err = cudaSetDevice(device); else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define ON_DEVICE(dev)                                                        \
    DeviceGuard deviceGuard__(dev);                                           \
    if (deviceGuard__.err != cudaSuccess) return cudaFail(deviceGuard__.err, "cudaSetDevice")

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct SavedState {
    float* q = nullptr;
    float* qd = nullptr;
    std::vector<std::pair<int, float*>> staticPoses;
};
}  // namespace

constexpr int kStepWarpsPerSM = 20;      // k_step<640>: 96 registers per thread
constexpr int kStepWarpsPerSMBig = 28;   // k_step<896>: 72 registers per thread, batches that need more than 20 warps per SM

struct b2g_batch {
    HostModel m;
    // the model as loaded: Box2DWorld::clone reloads the environment description (Box2DWorld.cpp:3030), so link parameters
    // changed afterwards (b2g_batch_set_link_property) are not inherited by a clone
    std::shared_ptr<const HostModel> loaded;
    int device = 0;
    long long N = 0;
    long long selFirst = 0, selCount = 0;  // b2g_batch_select_worlds: per-world calls act on [selFirst, selFirst + selCount)
    long long n() const { return selCount; }
    bool allSelected() const { return selFirst == 0 && selCount == N; }
    uint32_t* blob = nullptr;   // device: sets.size() blob variants, blobStride words apart
    // Parameter sets (b2g_batch_set_link_property on a tile-aligned selection of worlds): sets[0] is `m`; sets[v > 0] are copies of
    // the model whose link parameters differ.  tileSet[tile] names the set the 32 worlds of a tile use (all 0 while one set).
    std::vector<std::shared_ptr<HostModel>> extraSets;   // sets 1 ..
    std::vector<int> tileSet;                            // host copy, one entry per tile of the padded state array
    int* dTileSet = nullptr;
    int blobStride = 0;
    size_t blobCapacitySets = 0;
    size_t nSets() const { return 1 + extraSets.size(); }
    HostModel& set(int v) { return v == 0 ? m : *extraSets[(size_t)v - 1]; }
    const HostModel& set(int v) const { return v == 0 ? m : *extraSets[(size_t)v - 1]; }
    float* S = nullptr;
    // regrouped rollouts (regroupedRollout): the second state array the worlds are sorted into, the slot <-> world tables
    float* S2 = nullptr;
    size_t stateBytes = 0;
    DevBuf rgKey, rgNewSlot, rgWorldOfA, rgWorldOfB, rgCount, rgTargets;
    float* tmpl = nullptr;
    float* balls = nullptr;  // device copy of HostModel::balls (ball approximation queries)
    cudaStream_t stream = nullptr;
    StepParams prm{0.01f, 1.0f / 0.01f, 15, 15, 0, 0, nullptr, nullptr, nullptr, nullptr, 0, nullptr};
    int block = 32;
    int stepBlock = 32;
    int stepLanes = 32;
    int stepLpw = 1;
    int sms = 148;
    DevBuf stageA, stageB, stageC, stageD, stageE;
    std::vector<SavedState> stack;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timed;
    std::vector<cudaEvent_t> eventPool;

    Launch L() const {
        Launch l;
        l.h = m.h;
        l.blob = blob;
        l.S = S;
        l.N = selCount;
        l.prm = prm;
        l.prm.worldOffset = selFirst;
        l.prm.stateBase = (const char*)S;
        l.prm.blobStride = blobStride;
        l.prm.tileVariant = nSets() > 1 ? dTileSet : nullptr;
        l.prm.lead = nSets() > 1 ? (int)(selFirst & 31) : 0;
        l.stream = stream;
        l.block = block;
        l.stepBlock = stepBlock;
        l.stepLanes = stepLanes;
        if (nSets() > 1)  // parameter sets: the worlds of a warp share a 32-world tile, so the lane count divides 32
            while (32 % l.stepLanes) ++l.stepLanes;
        l.stepLpw = nSets() > 1 ? 1 : stepLpw;                        // the sub-warp kernel serves one-model batches only
        return l;
    }
};

namespace {
ObjectDescription* pickObject(b2g_env* env, int isRobot, int idx) {
    auto& v = isRobot ? env->d.robots : env->d.objects;
    if (idx < 0 || idx >= (int)v.size()) return nullptr;
    return &v[idx];
}
// Timing events of the stepping launches (b2g_batch_take_kernel_ms).  A planner that never asks for the timings must not
// accumulate events: at most kMaxTimed launches are remembered, the oldest pair is recycled beyond that.
constexpr size_t kMaxTimed = 64;
cudaError_t getEvent(b2g_batch* b, cudaEvent_t* out) {
    if (!b->eventPool.empty()) {
        *out = b->eventPool.back();
        b->eventPool.pop_back();
        return cudaSuccess;
    }
    return cudaEventCreate(out);
}
}  // namespace

// Objects with a prismatic joint can be loaded, stepped and read, but not placed: in the reference every DOF setter of such an
// object ends in Box2DLink::updateBodyVelocities -> Box2DJoint::getGlobalAxisPosition, which throws std::logic_error for a
// prismatic joint (Box2DWorld.cpp:789, 1282-1285), after Box2DJoint::resetPosition left the child link where it was (:1021-1026).
static b2g_status refusePrismatic(const b2g_batch* b, int object) {
    for (int o = 0; o < (int)b->m.objects.size(); ++o)
        if ((object < 0 || o == object) && b->m.objects[o].hasPrismatic)
            return fail(B2G_ERR_UNSUPPORTED, "object '" + b->m.objects[o].name + "' has a prismatic joint: its DOFs cannot be set (the "
                        "reference throws in getGlobalAxisPosition, Box2DWorld.cpp:1282-1285)");
    return B2G_OK;
}

extern "C" {

const char* b2g_last_error(void) { return g_error.c_str(); }

// ------------------------------------------------------------------------------------------------ environment
b2g_status b2g_env_create(float scale, const float world_bounds[4], b2g_env** out) {
    if (!out) return fail(B2G_ERR_INVALID, "out is null");
    b2g_env* e = new b2g_env();
    e->d.scale = scale;
    for (int i = 0; i < 4; ++i) e->d.world_bounds[i] = world_bounds ? world_bounds[i] : 0.0f;
    *out = e;
    return B2G_OK;
}

b2g_status b2g_env_load_yaml(const char* path, b2g_env** out) {
    if (!path || !out) return fail(B2G_ERR_INVALID, "null argument");
    b2g_env* e = new b2g_env();
    try {
        parseYAML(path, e->d);
    } catch (const std::exception& ex) {
        delete e;
        return fail(B2G_ERR_IO, ex.what());
    }
    *out = e;
    return B2G_OK;
}

void b2g_env_destroy(b2g_env* env) { delete env; }

b2g_status b2g_env_add_object(b2g_env* env, const char* name, int is_robot, int is_static, const float* base_limits8,
                              int use_center_of_mass, int* out_index) {
    if (!env || !name) return fail(B2G_ERR_INVALID, "null argument");
    ObjectDescription o;
    o.name = name;
    o.is_robot = is_robot != 0;
    o.is_static = is_static != 0;
    o.use_center_of_mass = use_center_of_mass != 0;
    if (base_limits8) {
        for (int k = 0; k < 2; ++k) {
            o.translational_velocity_limits[k] = base_limits8[k];
            o.translational_acceleration_limits[k] = base_limits8[2 + k];
            o.rotational_velocity_limits[k] = base_limits8[4 + k];
            o.rotational_acceleration_limits[k] = base_limits8[6 + k];
        }
    }
    auto& v = is_robot ? env->d.robots : env->d.objects;
    v.push_back(o);
    if (out_index) *out_index = (int)v.size() - 1;
    return B2G_OK;
}

b2g_status b2g_env_add_link(b2g_env* env, int is_robot, int object, const char* name, float mass, float ground_friction,
                            float ground_torque_integral, float contact_friction, float restitution, int* out_index) {
    ObjectDescription* o = env ? pickObject(env, is_robot, object) : nullptr;
    if (!o || !name) return fail(B2G_ERR_INVALID, "unknown object");
    LinkDescription l;
    l.name = name;
    l.mass = mass;
    l.ground_friction = ground_friction;
    l.ground_torque_integral = ground_torque_integral;
    l.contact_friction = contact_friction;
    l.restitution = restitution;
    o->links.push_back(l);
    if (out_index) *out_index = (int)o->links.size() - 1;
    return B2G_OK;
}

b2g_status b2g_env_add_polygon(b2g_env* env, int is_robot, int object, int link, const float* xy, int n_floats) {
    ObjectDescription* o = env ? pickObject(env, is_robot, object) : nullptr;
    if (!o || link < 0 || link >= (int)o->links.size() || !xy) return fail(B2G_ERR_INVALID, "unknown link");
    if (n_floats < 6 || n_floats % 2) return fail(B2G_ERR_INVALID, "a polygon needs an even number (>= 6) of floats");
    o->links[link].polygons.emplace_back(xy, xy + n_floats);
    return B2G_OK;
}

b2g_status b2g_env_add_ball(b2g_env* env, int is_robot, int object, int link, const float center[3], float radius) {
    ObjectDescription* o = env ? pickObject(env, is_robot, object) : nullptr;
    if (!o || link < 0 || link >= (int)o->links.size() || !center) return fail(B2G_ERR_INVALID, "unknown link");
    o->links[link].balls.push_back({center[0], center[1], center[2], radius});
    return B2G_OK;
}

b2g_status b2g_env_add_joint(b2g_env* env, int is_robot, int object, const char* name, const char* link_a, const char* link_b,
                             const float* p, int actuated, const char* joint_type) {
    ObjectDescription* o = env ? pickObject(env, is_robot, object) : nullptr;
    if (!o || !name || !link_a || !link_b || !p || !joint_type) return fail(B2G_ERR_INVALID, "bad joint arguments");
    JointDescription j;
    j.name = name;
    j.link_a = link_a;
    j.link_b = link_b;
    j.axis[0] = p[0]; j.axis[1] = p[1];
    j.position_limits[0] = p[2]; j.position_limits[1] = p[3];
    j.velocity_limits[0] = p[4]; j.velocity_limits[1] = p[5];
    j.acceleration_limits[0] = p[6]; j.acceleration_limits[1] = p[7];
    j.axis_orientation = p[8];
    j.actuated = actuated != 0;
    j.joint_type = joint_type;
    o->joints.push_back(j);
    return B2G_OK;
}

b2g_status b2g_env_add_state(b2g_env* env, const char* name, const float* configuration, const float* velocity, int n) {
    if (!env || !name || !configuration || !velocity || n < 0) return fail(B2G_ERR_INVALID, "bad state arguments");
    StateDescription s;
    s.configuration.assign(configuration, configuration + n);
    s.velocity.assign(velocity, velocity + n);
    env->d.states[name] = s;
    return B2G_OK;
}

// ------------------------------------------------------------------------------------------------------ model
b2g_status b2g_model_create(const b2g_env* env, b2g_model** out) {
    if (!env || !out) return fail(B2G_ERR_INVALID, "null argument");
    b2g_model* m = new b2g_model();
    m->env = env->d;
    try {
        for (auto& o : m->env.robots) {
            o.is_robot = true;
            std::string msg;
            if (!verifyKinematicStructure(o, msg)) throw std::runtime_error("robot '" + o.name + "': " + msg);
        }
        for (auto& o : m->env.objects) {
            o.is_robot = false;
            std::string msg;
            if (!verifyKinematicStructure(o, msg)) throw std::runtime_error("object '" + o.name + "': " + msg);
        }
        buildModel(m->env, 0, m->m);
    } catch (const std::exception& ex) {
        std::string what = ex.what();
        delete m;
        bool unsupported = what.find("not supported") != std::string::npos || what.find("unsupported") != std::string::npos;
        return fail(unsupported ? B2G_ERR_UNSUPPORTED : B2G_ERR_INVALID, what);
    }
    *out = m;
    return B2G_OK;
}

void b2g_model_destroy(b2g_model* model) { delete model; }

b2g_status b2g_model_get_info(const b2g_model* model, b2g_model_info* out) {
    if (!model || !out) return fail(B2G_ERR_INVALID, "null argument");
    const ModelHeader& h = model->m.h;
    out->n_bodies = h.nb; out->n_fixtures = h.nf; out->n_joints = h.nj; out->n_objects = h.no;
    out->n_robots = h.nrobots; out->nq = h.nq; out->scale = h.scale;
    return B2G_OK;
}

b2g_status b2g_model_get_warnings(const b2g_model* model, char* buf, int cap, int32_t* n_out) {
    if (!model || !n_out || (cap > 0 && !buf)) return fail(B2G_ERR_INVALID, "null argument");
    std::string all;
    for (const std::string& w : model->m.warnings) {
        all += w;
        all += '\n';
    }
    *n_out = (int32_t)model->m.warnings.size();
    if (cap > 0) {
        strncpy(buf, all.c_str(), (size_t)cap - 1);
        buf[cap - 1] = 0;
    }
    return B2G_OK;
}

b2g_status b2g_model_get_object(const b2g_model* model, int object, b2g_object_info* out) {
    if (!model || !out || object < 0 || object >= (int)model->m.objects.size()) return fail(B2G_ERR_INVALID, "unknown object");
    const HostObjectInfo& o = model->m.objects[object];
    memset(out, 0, sizeof(*out));
    strncpy(out->name, o.name.c_str(), sizeof(out->name) - 1);
    out->is_robot = o.isRobot; out->is_static = o.isStatic; out->n_dofs = o.nDofs; out->dof_offset = o.dofOffset;
    out->first_body = o.firstBody; out->n_links = o.nLinks; out->n_joints = o.nJoints;
    out->n_balls = o.nBalls;
    return B2G_OK;
}

b2g_status b2g_model_get_balls(const b2g_model* model, int object, int32_t* bodies, float* xyzr) {
    if (!model || object < 0 || object >= (int)model->m.objects.size()) return fail(B2G_ERR_INVALID, "unknown object");
    const HostObjectInfo& o = model->m.objects[object];
    if (o.nBalls > 0 && (!bodies || !xyzr)) return fail(B2G_ERR_INVALID, "null argument");
    for (int i = 0; i < o.nBalls; ++i) {
        const float* r = &model->m.balls[(size_t)(o.ballStart + i) * 8];
        bodies[i] = (int32_t)r[0];
        for (int k = 0; k < 4; ++k) xyzr[i * 4 + k] = r[3 + k];
    }
    return B2G_OK;
}

b2g_status b2g_model_get_dof_limits(const b2g_model* model, float* out) {
    if (!model || !out) return fail(B2G_ERR_INVALID, "null argument");
    memcpy(out, model->m.dofLimits.data(), model->m.dofLimits.size() * sizeof(float));
    return B2G_OK;
}
b2g_status b2g_model_get_bodies(const b2g_model* model, float* out) {
    if (!model || !out) return fail(B2G_ERR_INVALID, "null argument");
    memcpy(out, model->m.bodyModel.data(), model->m.bodyModel.size() * sizeof(float));
    return B2G_OK;
}
b2g_status b2g_model_get_fixtures(const b2g_model* model, int32_t* iout, float* fout) {
    if (!model || !iout || !fout) return fail(B2G_ERR_INVALID, "null argument");
    memcpy(iout, model->m.fixInts.data(), model->m.fixInts.size() * sizeof(int32_t));
    memcpy(fout, model->m.fixFloats.data(), model->m.fixFloats.size() * sizeof(float));
    return B2G_OK;
}
b2g_status b2g_model_get_joints(const b2g_model* model, int32_t* iout, float* fout) {
    if (!model || !iout || !fout) return fail(B2G_ERR_INVALID, "null argument");
    memcpy(iout, model->m.jointInts.data(), model->m.jointInts.size() * sizeof(int32_t));
    memcpy(fout, model->m.jointFloats.data(), model->m.jointFloats.size() * sizeof(float));
    return B2G_OK;
}

// ------------------------------------------------------------------------------------------------------ batch
static b2g_status createBatch(const HostModel& hostModel, int64_t n_worlds, int device, int max_contacts, b2g_batch** out);

b2g_status b2g_batch_create(const b2g_model* model, int64_t n_worlds, int device, int max_contacts, b2g_batch** out) {
    if (!model || !out || n_worlds <= 0) return fail(B2G_ERR_INVALID, "bad batch arguments");
    return createBatch(model->m, n_worlds, device, max_contacts, out);
}

// Box2DWorld::clone (Box2DWorld.cpp:3021-3038): a new world from the same environment description, same step parameters,
// then setWorldState(getWorldState()) -- DOF positions / velocities of every object plus the poses of static objects; the
// clone starts with fresh caches (no contacts, zero warm-start impulses), exactly like the reference's.
b2g_status b2g_batch_clone(b2g_batch* src, b2g_batch** out) {
    if (!src || !out) return fail(B2G_ERR_INVALID, "null argument");
    if (!src->allSelected()) return fail(B2G_ERR_INVALID, "stepping and the state stack act on the whole batch: call b2g_batch_select_worlds(batch, 0, 0) first");
    b2g_batch* dst = nullptr;
    b2g_status st = createBatch(src->loaded ? *src->loaded : src->m, src->N, src->device, src->m.h.maxc, &dst);
    if (st) return st;
    dst->prm.dt = src->prm.dt;
    dst->prm.inv_dt = src->prm.inv_dt;
    dst->prm.velIters = src->prm.velIters;
    dst->prm.posIters = src->prm.posIters;
    dst->prm.syncMask = src->prm.syncMask;
    auto bail = [&](cudaError_t e, const char* what) {
        b2g_batch_destroy(dst);
        return cudaFail(e, what);
    };
    cudaError_t e;
    const size_t bytes = (size_t)src->N * src->m.h.nq * 4;
    if ((e = src->stageA.reserve(bytes)) != cudaSuccess) return bail(e, "cudaMalloc(stage)");
    if ((e = src->stageB.reserve(bytes)) != cudaSuccess) return bail(e, "cudaMalloc(stage)");
    if ((e = src->stageC.reserve((size_t)src->N * 3 * 4)) != cudaSuccess) return bail(e, "cudaMalloc(stage)");
    // everything runs on the source's stream, in order; the clone's own stream starts after the final synchronisation
    Launch ld = dst->L();
    ld.stream = src->stream;
    for (int o = 0; o < src->m.h.no; ++o) {
        if (!src->m.objects[o].isStatic) continue;
        if ((e = launchGetPose(src->L(), o, (float*)src->stageC.p)) != cudaSuccess) return bail(e, "get pose kernel");
        if ((e = launchSetPose(ld, o, (const float*)src->stageC.p)) != cudaSuccess) return bail(e, "set pose kernel");
    }
    if ((e = launchGetState(src->L(), (float*)src->stageA.p, (float*)src->stageB.p)) != cudaSuccess) return bail(e, "get state kernel");
    if ((e = launchSetState(ld, (const float*)src->stageA.p, (const float*)src->stageB.p)) != cudaSuccess) return bail(e, "set state kernel");
    if ((e = cudaStreamSynchronize(src->stream)) != cudaSuccess) return bail(e, "clone sync");
    *out = dst;
    return B2G_OK;
}

static b2g_status createBatch(const HostModel& hostModel, int64_t n_worlds, int device, int max_contacts, b2g_batch** out) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(B2G_ERR_CUDA, std::string("no CUDA device available (the batched backend has no CPU fallback): ") +
                                      cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(B2G_ERR_INVALID, "unknown device");
    ON_DEVICE(device);
    b2g_batch* b = new b2g_batch();
    b->m = hostModel;
    if (max_contacts > 0) layoutState(b->m.h, max_contacts > kMaxContacts ? kMaxContacts : max_contacts);
    b->loaded = std::make_shared<const HostModel>(b->m);
    b->device = device;
    b->N = n_worlds;
    b->selFirst = 0;
    b->selCount = n_worlds;
    int sms = 148;  // SM count of the batch's device (B200: 148; 4 warp schedulers each)
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sms < 1) sms = 148;
    b->sms = sms;
    const char* envBlock = getenv("B2G_BLOCK");
    if (envBlock) b->block = atoi(envBlock);
    else b->block = n_worlds <= (long long)sms * 32 * 8 ? 32 : 64;
    if (b->block < 32 || b->block > 1024 || (b->block & (b->block - 1))) b->block = 32;
    // CTA size of the stepping kernel: its warps walk the solver phases together (phase barriers), which keeps the
    // instruction cache on one phase; one warp per CTA while the batch does not fill the GPU anyway
    const char* envStep = getenv("B2G_STEP_BLOCK");
    if (envStep) b->stepBlock = atoi(envStep);
    else if (envBlock) b->stepBlock = b->block;
    else {
        // ONE CTA per SM, sized so that the grid fills all 148 SMs evenly: with 96 registers per thread (__launch_bounds__(640))
        // an SM holds 20 warps of k_step, so the batch runs in `waves` rounds of at most 148 x 16 warps; the warps of one round are split
        // evenly over the SMs (CTA size = any multiple of 32) instead of leaving a partly filled last round.  All warps of an
        // SM then walk the phase barriers together and share one phase's code in the instruction caches; two CTAs per SM
        // (round 1) drift against each other.  Measured, 65536 worlds, ms per 100 steps, 2 x 224 -> 1 x 448 threads per SM:
        // cfg 2 41.5 -> 38.8, cfg 3 10.5 -> 9.6, cfg 4 109.6 -> 103.0; cfg 2 x 262144 162.5 -> 151.8 (1 x 512: 40.0 / 10.0 /
        // 106.2 / 155.7; profiles/r2_block_sweep.txt).
        const long long warps = (n_worlds + 31) / 32;
        // batches beyond one wave of 20 warps per SM run the 72-register kernel with up to 28 warps per SM (b2g_kernels.cu: k_step)
        const long long cap = warps <= (long long)sms * kStepWarpsPerSM ? kStepWarpsPerSM : kStepWarpsPerSMBig;
        const long long waves = (warps + sms * cap - 1) / (sms * cap);
        long long w = (warps + waves * sms - 1) / (waves * sms);
        if (w < 1) w = 1;
        if (w > cap) w = cap;
        b->stepBlock = (int)(32 * w);
    }
    if (b->stepBlock < 32 || b->stepBlock > 32 * kStepWarpsPerSMBig || (b->stepBlock % 32)) b->stepBlock = 32;
    // worlds per warp: a B200 has 148 x 4 warp schedulers; a batch with fewer than 592 x 16 worlds is spread over narrower
    // warps (k_step_narrow), one warp per scheduler at most, so that rare per-world events serialise fewer neighbours
    const char* envLanes = getenv("B2G_LANES");
    if (envLanes) b->stepLanes = atoi(envLanes);
    else {
        // Up to 4 worlds per warp while that keeps the batch on at most 518 warps (a B200 has 592 warp schedulers); beyond
        // that 16, then full warps.  The stepping kernel streams ~400 KB of once-per-step code per warp and step through the
        // GPC-level instruction caches, so FEWER, wider warps win as soon as the batch does not fit 4 lanes: measured on the
        // contact-rich workloads, 4096 worlds, ms per 100 steps (profiles/r2_mode_sweep.txt):
        //   cfg 2: 4 -> 27.2, 6 -> 20.4, 7 -> 17.2, 8 -> 16.8, 10 -> 15.6, 12 -> 15.8, 16 -> 15.9, 24 -> 16.7, 32 -> 17.3
        //   cfg 3: 7 -> 7.8, 8 -> 7.3, 12 -> 6.0, 16 -> 5.7, 24 -> 5.7, 32 -> 5.7
        // Re-measured after the later kernel changes (cfg 2 / cfg 3 / contact-free inputs): 10 -> 15.4 / 6.8 / 8.4, 12 -> 15.3 / 6.3 / 8.35,
        // 16 -> 15.5 / 5.8 / 8.5: 12 worlds per warp while that fits 518 warps (the headline shape), then 16.
        const long long warpsTarget = (long long)sms * 4 * 7 / 8;  // 518 on a 148-SM B200
        long long l = (n_worlds + warpsTarget - 1) / warpsTarget;
        b->stepLanes = l <= 4 ? (int)l : (l <= 12 ? 12 : (l <= 16 ? 16 : 32));
    }
    if (b->stepLanes < 1 || b->stepLanes > 32) b->stepLanes = 32;
    // phase barriers of k_step (b2g_solver.cuh: phaseSync); B2G_SYNC_MASK overrides for experiments
    {
        const char* envMask = getenv("B2G_SYNC_MASK");
        b->prm.syncMask = envMask ? (int)strtol(envMask, nullptr, 0) : (b->m.h.nb >= 16 ? 0x2aa : 0x7ff);
    }
    // lanes per world (b2g_multi.cuh: a world owned by a group of lanes).  One thread per world is the faster mapping for small
    // models at every batch size; a scene with many bodies at a small batch is where the lanes of a group have
    // enough per-body / per-joint / per-contact work to share (narrow phase and broad-phase box tests included).  Measured on
    // the 38-body clutter scene, ms per 100 steps, one thread per world against the best group size: 512 worlds 35.3 -> 18.9
    // (16 lanes), 1024: 39.1 -> 22.0 (16), 2048: 45.4 -> 31.6 (8), 4096: 62.6 -> 55.6 (8), 8192: 71.0 -> 90.0 (8: no).
    // B2G_LPW overrides.
    const char* envLpw = getenv("B2G_LPW");
    if (envLpw) b->stepLpw = atoi(envLpw);
    else if (b->m.h.nb >= 16 && !envLanes)
        b->stepLpw = n_worlds <= 192 ? 32 : (n_worlds <= 1036 ? 16 : (n_worlds <= 4144 ? 8 : 1));
    else b->stepLpw = 1;
    if (b->stepLpw != 1 && b->stepLpw != 2 && b->stepLpw != 4 && b->stepLpw != 8 && b->stepLpw != 16 && b->stepLpw != 32) b->stepLpw = 1;
    if (b->stepLpw > 1) {
        const int maxWorlds = 32 / b->stepLpw;
        if (!envLanes || b->stepLanes > maxWorlds) b->stepLanes = maxWorlds;
    }
    auto cleanup = [&](cudaError_t err, const char* what) {
        b2g_batch_destroy(b);
        return cudaFail(err, what);
    };
    if ((e = cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking)) != cudaSuccess) return cleanup(e, "cudaStreamCreate");
    b->blobStride = (int)((b->m.blob.size() + 3) & ~(size_t)3);
    b->blobCapacitySets = 1;
    if ((e = cudaMalloc((void**)&b->blob, (size_t)b->blobStride * 4)) != cudaSuccess) return cleanup(e, "cudaMalloc(model)");
    if ((e = cudaMemcpy(b->blob, b->m.blob.data(), b->m.blob.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess)
        return cleanup(e, "cudaMemcpy(model)");
    // state[tile][chunk][lane]: tiles of 32 worlds, 16-byte chunks (model.h)
    // padded to a multiple of kWorldPad worlds: the padding worlds are idle copies that let every thread of a k_step CTA
    // take part in the phase barriers
    size_t tiles = (size_t)paddedTiles(n_worlds);
    size_t stateBytes = tiles * (size_t)b->m.h.stateChunks * kChunkBytes;
    b->stateBytes = stateBytes;
    if ((e = cudaMalloc((void**)&b->S, stateBytes)) != cudaSuccess) return cleanup(e, "cudaMalloc(state)");
    if ((e = cudaMalloc((void**)&b->tmpl, (size_t)b->m.h.stateChunks * 16)) != cudaSuccess) return cleanup(e, "cudaMalloc(template)");
    // build the freshly loaded world once (world 0), keep its chunks as the template, broadcast to every world
    {
        Launch l = b->L();
        l.N = 1;
        if ((e = launchInit(l)) != cudaSuccess) return cleanup(e, "init kernel");
        if ((e = launchExtractWorld(b->L(), b->tmpl, 0)) != cudaSuccess) return cleanup(e, "extract kernel");
        {   // every world, padding included, starts as a copy of the loaded world; the padding worlds are then parked
            Launch p = b->L();
            p.N = paddedTiles(n_worlds) * kTileWorlds;
            if ((e = launchBroadcast(p, b->tmpl)) != cudaSuccess) return cleanup(e, "broadcast kernel");
            p.prm.worldOffset = n_worlds;
            p.N = paddedTiles(n_worlds) * kTileWorlds - n_worlds;
            if ((e = launchPark(p)) != cudaSuccess) return cleanup(e, "park kernel");
        }
        if ((e = cudaStreamSynchronize(b->stream)) != cudaSuccess) return cleanup(e, "init sync");
    }
    *out = b;
    return B2G_OK;
}

void b2g_batch_destroy(b2g_batch* b) {
    if (!b) return;
    DeviceGuard guard(b->device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    for (auto& s : b->stack) {
        cudaFree(s.q);
        cudaFree(s.qd);
        for (auto& p : s.staticPoses) cudaFree(p.second);
    }
    for (auto& p : b->timed) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    for (auto& ev : b->eventPool) cudaEventDestroy(ev);
    b->stageA.release(); b->stageB.release(); b->stageC.release(); b->stageD.release(); b->stageE.release();
    b->rgKey.release(); b->rgNewSlot.release(); b->rgWorldOfA.release(); b->rgWorldOfB.release(); b->rgCount.release(); b->rgTargets.release();
    if (b->blob) cudaFree(b->blob);
    if (b->dTileSet) cudaFree(b->dTileSet);
    if (b->S) cudaFree(b->S);
    if (b->S2) cudaFree(b->S2);
    if (b->tmpl) cudaFree(b->tmpl);
    if (b->balls) cudaFree(b->balls);
    if (b->stream) cudaStreamDestroy(b->stream);
    delete b;
}

b2g_status b2g_batch_select_worlds(b2g_batch* b, int64_t first, int64_t count) {
    if (!b) return fail(B2G_ERR_INVALID, "null batch");
    if (count <= 0) { first = 0; count = b->N; }
    if (first < 0 || first + count > b->N) return fail(B2G_ERR_INVALID, "world range outside the batch");
    b->selFirst = first;
    b->selCount = count;
    return B2G_OK;
}

b2g_status b2g_batch_set_params(b2g_batch* b, float dt, int velocity_steps, int position_steps) {
    if (!b || !(dt > 0.0f) || velocity_steps < 0 || position_steps < 0) return fail(B2G_ERR_INVALID, "bad step parameters");
    b->prm.dt = dt;
    b->prm.inv_dt = dt > 0.0f ? 1.0f / dt : 0.0f;
    b->prm.velIters = velocity_steps;
    b->prm.posIters = position_steps;
    return B2G_OK;
}

void* b2g_batch_stream(b2g_batch* b) { return b ? (void*)b->stream : nullptr; }

b2g_status b2g_batch_sync(b2g_batch* b) {
    if (!b) return fail(B2G_ERR_INVALID, "null batch");
    ON_DEVICE(b->device);
    CU(cudaStreamSynchronize(b->stream));
    return B2G_OK;
}

b2g_status b2g_batch_status(b2g_batch* b, int32_t* out) {
    if (!b || !out) return fail(B2G_ERR_INVALID, "null argument");
    ON_DEVICE(b->device);
    CU(b->stageA.reserve((size_t)b->n() * 4));
    CU(launchStatus(b->L(), (int*)b->stageA.p));
    CU(cudaMemcpyAsync(out, b->stageA.p, (size_t)b->n() * 4, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    for (long long i = 0; i < b->n(); ++i)
        if (out[i] & 1) return fail(B2G_ERR_CAPACITY, "world " + std::to_string(i) + " exceeded its contact-slot capacity");
    return B2G_OK;
}

// reset_caches: every selected world becomes what Box2DWorld::clone() gives (Box2DWorld.cpp:3026-3037: a freshly loaded world
// that then receives the world state) -- contact list, warm-start impulses, sleep timers, broad-phase move buffer, link
// enable flags and controller targets as loaded -- EXCEPT the poses of static objects, which are part of the world state the
// clone receives (and which b2g_batch_clone carries over the same way): they are read before the template is broadcast and set
// again afterwards, as setPose calls on the fresh world.
static b2g_status resetToFreshClones(b2g_batch* b) {
    int nStatic = 0;
    for (const auto& o : b->m.objects) nStatic += o.isStatic ? 1 : 0;
    const size_t per = (size_t)b->n() * 3 * 4;
    if (nStatic) CU(b->stageE.reserve(per * nStatic));
    int k = 0;
    for (int o = 0; o < b->m.h.no; ++o)
        if (b->m.objects[o].isStatic) CU(launchGetPose(b->L(), o, (float*)((char*)b->stageE.p + per * k++)));
    CU(launchBroadcast(b->L(), b->tmpl));
    k = 0;
    for (int o = 0; o < b->m.h.no; ++o)
        if (b->m.objects[o].isStatic) CU(launchSetPose(b->L(), o, (const float*)((char*)b->stageE.p + per * k++)));
    return B2G_OK;
}

b2g_status b2g_batch_set_state_dev(b2g_batch* b, const float* q, const float* qd, int reset_caches) {
    if (!b || !q || !qd) return fail(B2G_ERR_INVALID, "null argument");
    if (b2g_status st = refusePrismatic(b, -1)) return st;
    ON_DEVICE(b->device);
    if (reset_caches)
        if (b2g_status st = resetToFreshClones(b)) return st;
    CU(launchSetState(b->L(), q, qd));
    return B2G_OK;
}

b2g_status b2g_batch_set_state(b2g_batch* b, const float* q, const float* qd, int reset_caches) {
    if (!b || !q || !qd) return fail(B2G_ERR_INVALID, "null argument");
    ON_DEVICE(b->device);
    size_t bytes = (size_t)b->n() * b->m.h.nq * 4;
    CU(b->stageA.reserve(bytes));
    CU(b->stageB.reserve(bytes));
    CU(cudaMemcpyAsync(b->stageA.p, q, bytes, cudaMemcpyHostToDevice, b->stream));
    CU(cudaMemcpyAsync(b->stageB.p, qd, bytes, cudaMemcpyHostToDevice, b->stream));
    b2g_status s = b2g_batch_set_state_dev(b, (const float*)b->stageA.p, (const float*)b->stageB.p, reset_caches);
    if (s) return s;
    CU(cudaStreamSynchronize(b->stream));
    return B2G_OK;
}

b2g_status b2g_batch_get_state_dev(b2g_batch* b, float* q, float* qd) {
    if (!b || !q || !qd) return fail(B2G_ERR_INVALID, "null argument");
    ON_DEVICE(b->device);
    CU(launchGetState(b->L(), q, qd));
    return B2G_OK;
}

b2g_status b2g_batch_get_state(b2g_batch* b, float* q, float* qd) {
    if (!b || !q || !qd) return fail(B2G_ERR_INVALID, "null argument");
    ON_DEVICE(b->device);
    size_t bytes = (size_t)b->n() * b->m.h.nq * 4;
    CU(b->stageA.reserve(bytes));
    CU(b->stageB.reserve(bytes));
    b2g_status s = b2g_batch_get_state_dev(b, (float*)b->stageA.p, (float*)b->stageB.p);
    if (s) return s;
    CU(cudaMemcpyAsync(q, b->stageA.p, bytes, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaMemcpyAsync(qd, b->stageB.p, bytes, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    return B2G_OK;
}

namespace {
// object-local DOF selection of the per-object DOF calls: null / 0 = all DOFs; indices ascending (Box2DWorld.cpp:1611)
b2g_status makeSel(const b2g_batch* b, int object, const int32_t* indices, int n, DofSel& sel) {
    if (!b || object < 0 || object >= b->m.h.no) return fail(B2G_ERR_INVALID, "unknown object");
    const int nd = b->m.objects[object].nDofs;
    if (!indices || n <= 0) {
        sel.n = nd;
        for (int i = 0; i < nd; ++i) sel.idx[i] = i;
        return B2G_OK;
    }
    if (n > nd) return fail(B2G_ERR_INVALID, "more DOF indices than the object has DOFs");
    sel.n = n;
    for (int i = 0; i < n; ++i) {
        if (indices[i] < 0 || indices[i] >= nd) return fail(B2G_ERR_INVALID, "DOF index out of range");
        if (i > 0 && indices[i] <= indices[i - 1]) return fail(B2G_ERR_INVALID, "DOF indices must be ascending");
        sel.idx[i] = indices[i];
    }
    return B2G_OK;
}
}  // namespace

b2g_status b2g_batch_set_object_dofs(b2g_batch* b, int object, int velocities, const int32_t* indices, int n_indices,
                                     const float* values) {
    DofSel sel;
    b2g_status st = makeSel(b, object, indices, n_indices, sel);
    if (st) return st;
    if (sel.n == 0) return B2G_OK;
    if (!values) return fail(B2G_ERR_INVALID, "null argument");
    if (b2g_status st2 = refusePrismatic(b, object)) return st2;
    ON_DEVICE(b->device);
    size_t bytes = (size_t)b->n() * sel.n * 4;
    CU(b->stageA.reserve(bytes));
    CU(cudaMemcpyAsync(b->stageA.p, values, bytes, cudaMemcpyHostToDevice, b->stream));
    CU(launchSetObjectDofs(b->L(), object, sel, (const float*)b->stageA.p, velocities));
    CU(cudaStreamSynchronize(b->stream));
    return B2G_OK;
}

b2g_status b2g_batch_get_object_dofs(b2g_batch* b, int object, int velocities, const int32_t* indices, int n_indices, float* out) {
    DofSel sel;
    b2g_status st = makeSel(b, object, indices, n_indices, sel);
    if (st) return st;
    if (sel.n == 0) return B2G_OK;
    if (!out) return fail(B2G_ERR_INVALID, "null argument");
    ON_DEVICE(b->device);
    size_t bytes = (size_t)b->n() * sel.n * 4;
    CU(b->stageA.reserve(bytes));
    CU(launchGetObjectDofs(b->L(), object, sel, (float*)b->stageA.p, velocities));
    CU(cudaMemcpyAsync(out, b->stageA.p, bytes, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    return B2G_OK;
}

b2g_status b2g_batch_set_object_pose(b2g_batch* b, int object, const float* pose) {
    if (!b || !pose || object < 0 || object >= b->m.h.no) return fail(B2G_ERR_INVALID, "unknown object");
    if (b2g_status st = refusePrismatic(b, object)) return st;
    ON_DEVICE(b->device);
    size_t bytes = (size_t)b->n() * 3 * 4;
    CU(b->stageA.reserve(bytes));
    CU(cudaMemcpyAsync(b->stageA.p, pose, bytes, cudaMemcpyHostToDevice, b->stream));
    CU(launchSetPose(b->L(), object, (const float*)b->stageA.p));
    CU(cudaStreamSynchronize(b->stream));
    return B2G_OK;
}

b2g_status b2g_batch_get_object_pose(b2g_batch* b, int object, float* pose) {
    if (!b || !pose || object < 0 || object >= b->m.h.no) return fail(B2G_ERR_INVALID, "unknown object");
    ON_DEVICE(b->device);
    size_t bytes = (size_t)b->n() * 3 * 4;
    CU(b->stageA.reserve(bytes));
    CU(launchGetPose(b->L(), object, (float*)b->stageA.p));
    CU(cudaMemcpyAsync(pose, b->stageA.p, bytes, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    return B2G_OK;
}

static b2g_status checkRobot(b2g_batch* b, int object) {
    if (!b || object < 0 || object >= b->m.h.no) return fail(B2G_ERR_INVALID, "unknown object");
    if (!b->m.objects[object].isRobot)
        return fail(B2G_ERR_INVALID, "object '" + b->m.objects[object].name + "' is not a robot: it has no velocity controller");
    return B2G_OK;
}

b2g_status b2g_batch_set_target_velocity_dev(b2g_batch* b, int object, const float* v) {
    b2g_status s = checkRobot(b, object);
    if (s) return s;
    if (!v) return fail(B2G_ERR_INVALID, "null argument");
    ON_DEVICE(b->device);
    CU(launchSetTarget(b->L(), object, v));
    return B2G_OK;
}

b2g_status b2g_batch_set_target_velocity(b2g_batch* b, int object, const float* v) {
    b2g_status s = checkRobot(b, object);
    if (s) return s;
    if (!v) return fail(B2G_ERR_INVALID, "null argument");
    ON_DEVICE(b->device);
    size_t bytes = (size_t)b->n() * b->m.objects[object].nDofs * 4;
    CU(b->stageA.reserve(bytes));
    CU(cudaMemcpyAsync(b->stageA.p, v, bytes, cudaMemcpyHostToDevice, b->stream));
    CU(launchSetTarget(b->L(), object, (const float*)b->stageA.p));
    CU(cudaStreamSynchronize(b->stream));
    return B2G_OK;
}

b2g_status b2g_batch_set_target_position_dev(b2g_batch* b, int object, const float* q_target, float kp) {
    b2g_status s = checkRobot(b, object);
    if (s) return s;
    if (!q_target || !(kp >= 0.0f)) return fail(B2G_ERR_INVALID, "bad position target");
    ON_DEVICE(b->device);
    CU(launchSetTargetPosition(b->L(), object, q_target, kp));
    return B2G_OK;
}

b2g_status b2g_batch_set_target_position(b2g_batch* b, int object, const float* q_target, float kp) {
    b2g_status s = checkRobot(b, object);
    if (s) return s;
    if (!q_target || !(kp >= 0.0f)) return fail(B2G_ERR_INVALID, "bad position target");
    ON_DEVICE(b->device);
    size_t bytes = (size_t)b->n() * b->m.objects[object].nDofs * 4;
    CU(b->stageA.reserve(bytes));
    CU(cudaMemcpyAsync(b->stageA.p, q_target, bytes, cudaMemcpyHostToDevice, b->stream));
    CU(launchSetTargetPosition(b->L(), object, (const float*)b->stageA.p, kp));
    CU(cudaStreamSynchronize(b->stream));
    return B2G_OK;
}

b2g_status b2g_batch_set_control_efforts_dev(b2g_batch* b, int object, const float* efforts) {
    b2g_status s = checkRobot(b, object);
    if (s) return s;
    if (!efforts) return fail(B2G_ERR_INVALID, "null argument");
    ON_DEVICE(b->device);
    CU(launchSetEfforts(b->L(), object, efforts));
    return B2G_OK;
}

b2g_status b2g_batch_set_control_efforts(b2g_batch* b, int object, const float* efforts) {
    b2g_status s = checkRobot(b, object);
    if (s) return s;
    if (!efforts) return fail(B2G_ERR_INVALID, "null argument");
    ON_DEVICE(b->device);
    size_t bytes = (size_t)b->n() * b->m.objects[object].nDofs * 4;
    CU(b->stageA.reserve(bytes));
    CU(cudaMemcpyAsync(b->stageA.p, efforts, bytes, cudaMemcpyHostToDevice, b->stream));
    CU(launchSetEfforts(b->L(), object, (const float*)b->stageA.p));
    CU(cudaStreamSynchronize(b->stream));
    return B2G_OK;
}

// Large rollouts, OPT-IN (B2G_REGROUP=1; measured: it does not pay): before every target segment the worlds are
// sorted by island shape into the other state array (b2g_kernels.cu: k_regroup_*), so that the worlds of a warp walk the same
// joint lists; after the last segment they go back to slot == world.  Single-model batches in the full-warp regime only.
static bool regroupWanted(const b2g_batch* b, const float* targets, int nTargets, int stepsPerTarget) {
    const char* env = getenv("B2G_REGROUP");
    const int mode = env ? atoi(env) : 0;
    return mode != 0 && targets != nullptr && nTargets >= 2 && stepsPerTarget >= 5 && b->nSets() == 1 && b->stepLpw == 1 &&
           b->stepLanes == 32 && b->N >= 32768 && b->prm.recMode == 0;
}

static cudaError_t regroupedRollout(b2g_batch* b, int object, const float* targets, int nTargets, int stepsPerTarget, int execCtrl,
                                    int allowSleep) {
    cudaError_t e;
    const int nd = b->m.objects[object].nDofs;
    const size_t n = (size_t)b->N;
    if (!b->S2) {
        if ((e = cudaMalloc((void**)&b->S2, b->stateBytes)) != cudaSuccess) return e;
        // whole array once: the parked padding worlds behind the batch have to exist in both arrays
        if ((e = cudaMemcpyAsync(b->S2, b->S, b->stateBytes, cudaMemcpyDeviceToDevice, b->stream)) != cudaSuccess) return e;
    }
    if ((e = b->rgKey.reserve(n)) != cudaSuccess || (e = b->rgNewSlot.reserve(n * 4)) != cudaSuccess ||
        (e = b->rgWorldOfA.reserve(n * 4)) != cudaSuccess || (e = b->rgWorldOfB.reserve(n * 4)) != cudaSuccess ||
        (e = b->rgTargets.reserve(n * nd * 4)) != cudaSuccess)
        return e;
    // worlds are sorted inside the slot range of one CTA of the stepping kernel (B2G_REGROUP_GROUP overrides: 0 = the whole batch)
    const char* groupStr = getenv("B2G_REGROUP_GROUP");
    const long long groupEnv = groupStr ? atoll(groupStr) : -1ll;
    const int groupSize = groupEnv < 0 ? b->stepBlock : (groupEnv == 0 ? (int)std::min<long long>(b->N, 1ll << 30) : (int)groupEnv);
    const size_t groups = (n + groupSize - 1) / groupSize;
    if ((e = b->rgCount.reserve(groups * 64 * 4)) != cudaSuccess) return e;
    int* worldOf = (int*)b->rgWorldOfA.p;
    int* worldOfNext = (int*)b->rgWorldOfB.p;
    int* count = (int*)b->rgCount.p;
    int* cursor = count + groups * 32;
    if ((e = one::launchRegroupIota(b->L(), worldOf)) != cudaSuccess) return e;
    for (int t = 0; t < nTargets; ++t) {
        if ((e = one::launchRegroupKey(b->L(), (unsigned char*)b->rgKey.p, count, groupSize)) != cudaSuccess) return e;
        if ((e = one::launchRegroupAssign(b->L(), (const unsigned char*)b->rgKey.p, count, cursor, (int*)b->rgNewSlot.p, groupSize)) != cudaSuccess) return e;
        if ((e = one::launchRegroupMove(b->L(), b->S, b->S2, (const int*)b->rgNewSlot.p, worldOf, worldOfNext)) != cudaSuccess) return e;
        std::swap(b->S, b->S2);
        std::swap(worldOf, worldOfNext);
        if ((e = one::launchRegroupTargets(b->L(), targets + (size_t)t * n * nd, (float*)b->rgTargets.p, worldOf, nd)) != cudaSuccess) return e;
        if ((e = launchStep(b->L(), object, (const float*)b->rgTargets.p, 1, stepsPerTarget, execCtrl, allowSleep)) != cudaSuccess) return e;
    }
    // back to slot == world: the slot -> world table is the destination of every slot
    if ((e = one::launchRegroupMove(b->L(), b->S, b->S2, worldOf, worldOf, worldOfNext)) != cudaSuccess) return e;
    std::swap(b->S, b->S2);
    return cudaSuccess;
}

static b2g_status timedStep(b2g_batch* b, int object, const float* targets, int nTargets, int stepsPerTarget, int execCtrl,
                            int allowSleep) {
    if (b->timed.size() >= kMaxTimed) {  // nobody collected the timings: forget the oldest launch, reuse its events
        b->eventPool.push_back(b->timed.front().first);
        b->eventPool.push_back(b->timed.front().second);
        b->timed.erase(b->timed.begin());
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t e = getEvent(b, &e0);
    if (e == cudaSuccess) e = getEvent(b, &e1);
    if (e == cudaSuccess) e = cudaEventRecord(e0, b->stream);
    if (e == cudaSuccess)
        e = regroupWanted(b, targets, nTargets, stepsPerTarget) ? regroupedRollout(b, object, targets, nTargets, stepsPerTarget, execCtrl, allowSleep)
                                                                 : launchStep(b->L(), object, targets, nTargets, stepsPerTarget, execCtrl, allowSleep);
    if (e == cudaSuccess) e = cudaEventRecord(e1, b->stream);
    if (e != cudaSuccess) {  // the events go back to the pool, nothing is leaked on the error path
        if (e0) b->eventPool.push_back(e0);
        if (e1) b->eventPool.push_back(e1);
        return cudaFail(e, "stepping launch");
    }
    b->timed.emplace_back(e0, e1);
    return B2G_OK;
}

b2g_status b2g_batch_step(b2g_batch* b, int steps, int execute_controllers, int allow_sleeping) {
    if (!b || steps < 0) return fail(B2G_ERR_INVALID, "bad step arguments");
    if (!b->allSelected()) return fail(B2G_ERR_INVALID, "stepping and the state stack act on the whole batch: call b2g_batch_select_worlds(batch, 0, 0) first");
    ON_DEVICE(b->device);
    if (steps == 0) return B2G_OK;
    b2g_status s = timedStep(b, 0, nullptr, 1, steps, execute_controllers, allow_sleeping);
    if (s) return s;
    CU(cudaStreamSynchronize(b->stream));
    return B2G_OK;
}

b2g_status b2g_batch_rollout_dev(b2g_batch* b, int object, const float* targets, int n_targets, int steps_per_target) {
    b2g_status s = checkRobot(b, object);
    if (s) return s;
    if (!targets || n_targets < 0 || steps_per_target < 0) return fail(B2G_ERR_INVALID, "bad rollout arguments");
    if (!b->allSelected()) return fail(B2G_ERR_INVALID, "stepping and the state stack act on the whole batch: call b2g_batch_select_worlds(batch, 0, 0) first");
    ON_DEVICE(b->device);
    if (n_targets == 0) return B2G_OK;
    return timedStep(b, object, targets, n_targets, steps_per_target, 1, 1);
}

b2g_status b2g_batch_rollout(b2g_batch* b, int object, const float* q, const float* qd, int reset_caches, const float* targets,
                             int n_targets, int steps_per_target, float* q_out, float* qd_out) {
    b2g_status s = checkRobot(b, object);
    if (s) return s;
    if (!targets || n_targets < 0 || steps_per_target < 0 || !q_out || !qd_out || (q == nullptr) != (qd == nullptr))
        return fail(B2G_ERR_INVALID, "bad rollout arguments");
    if (!b->allSelected()) return fail(B2G_ERR_INVALID, "stepping and the state stack act on the whole batch: call b2g_batch_select_worlds(batch, 0, 0) first");
    if (q)
        if (b2g_status st = refusePrismatic(b, -1)) return st;
    ON_DEVICE(b->device);
    const size_t sb = (size_t)b->N * b->m.h.nq * 4;
    const size_t tb = (size_t)n_targets * b->N * b->m.objects[object].nDofs * 4;
    CU(b->stageA.reserve(sb));
    CU(b->stageB.reserve(sb));
    CU(b->stageC.reserve(tb > 0 ? tb : 4));
    if (q) {
        CU(cudaMemcpyAsync(b->stageA.p, q, sb, cudaMemcpyHostToDevice, b->stream));
        CU(cudaMemcpyAsync(b->stageB.p, qd, sb, cudaMemcpyHostToDevice, b->stream));
        if (reset_caches)
            if (b2g_status st = resetToFreshClones(b)) return st;
        CU(launchSetState(b->L(), (const float*)b->stageA.p, (const float*)b->stageB.p));
    }
    if (n_targets > 0) {
        CU(cudaMemcpyAsync(b->stageC.p, targets, tb, cudaMemcpyHostToDevice, b->stream));
        s = timedStep(b, object, (const float*)b->stageC.p, n_targets, steps_per_target, 1, 1);
        if (s) return s;
    }
    CU(launchGetState(b->L(), (float*)b->stageA.p, (float*)b->stageB.p));
    CU(cudaMemcpyAsync(q_out, b->stageA.p, sb, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaMemcpyAsync(qd_out, b->stageB.p, sb, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    return B2G_OK;
}

b2g_status b2g_batch_set_to_rest(b2g_batch* b) {
    if (!b) return fail(B2G_ERR_INVALID, "null batch");
    if (b2g_status st = refusePrismatic(b, -1)) return st;  // setToRest = setDOFVelocities(0) on every object
    ON_DEVICE(b->device);
    CU(launchSetToRest(b->L()));
    CU(cudaStreamSynchronize(b->stream));
    return B2G_OK;
}

b2g_status b2g_batch_at_rest(b2g_batch* b, float threshold, int32_t* out) {
    if (!b || !out) return fail(B2G_ERR_INVALID, "null argument");
    ON_DEVICE(b->device);
    CU(b->stageC.reserve((size_t)b->n() * 4));
    CU(launchAtRest(b->L(), threshold, (int*)b->stageC.p));
    CU(cudaMemcpyAsync(out, b->stageC.p, (size_t)b->n() * 4, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    return B2G_OK;
}

// getObjects(aabb, objects, exclude_robots) (:3207-3254): the box is scaled like the reference does (:3210-3214)
b2g_status b2g_batch_objects_in_aabb(b2g_batch* b, const float aabb[4], int exclude_robots, uint8_t* hits) {
    if (!b || !aabb || !hits) return fail(B2G_ERR_INVALID, "null argument");
    ON_DEVICE(b->device);
    const float s = b->m.h.scale;
    const int no = b->m.h.no;
    size_t bytes = (size_t)b->n() * no;
    CU(b->stageC.reserve(bytes));
    CU(launchObjectsInAABB(b->L(), s * aabb[0], s * aabb[1], s * aabb[2], s * aabb[3], (unsigned char*)b->stageC.p));
    CU(cudaMemcpyAsync(hits, b->stageC.p, bytes, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    if (exclude_robots)
        for (long long w = 0; w < b->n(); ++w)
            for (int o = 0; o < no; ++o)
                if (b->m.objects[o].isRobot) hits[(size_t)w * no + o] = 0;
    return B2G_OK;
}

b2g_status b2g_batch_is_physically_feasible(b2g_batch* b, int32_t* out) {
    if (!b || !out) return fail(B2G_ERR_INVALID, "null argument");
    ON_DEVICE(b->device);
    CU(b->stageC.reserve((size_t)b->n() * 4));
    CU(launchPhysicallyFeasible(b->L(), (int*)b->stageC.p));
    CU(cudaMemcpyAsync(out, b->stageC.p, (size_t)b->n() * 4, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    return B2G_OK;
}

b2g_status b2g_batch_set_link_enabled(b2g_batch* b, int body, const uint8_t* enabled, int enabled_all) {
    if (!b || body <= 0 || body >= b->m.h.nb) return fail(B2G_ERR_INVALID, "unknown link (body 0 is the ground)");
    ON_DEVICE(b->device);
    const unsigned char* dev = nullptr;
    if (enabled) {
        CU(b->stageC.reserve((size_t)b->n()));
        CU(cudaMemcpyAsync(b->stageC.p, enabled, (size_t)b->n(), cudaMemcpyHostToDevice, b->stream));
        dev = (const unsigned char*)b->stageC.p;
    }
    CU(launchSetLinkEnabled(b->L(), body, dev, enabled_all));
    CU(cudaStreamSynchronize(b->stream));
    return B2G_OK;
}

// upstream asserts on these (b2IsValid, b2FrictionJoint::SetMaxForce / SetMaxTorque: force >= 0, b2Fixture friction is used
// under a square root): refused here instead of poisoning every world of the batch with NaNs
static const char* badLinkValue(int property, float value) {
    if (!std::isfinite(value)) return "link parameter is not finite";
    if (property != B2G_LINK_MASS && value < 0.0f) return "friction parameters must not be negative";
    return nullptr;
}

// Box2DLink::setMass / setGroundFrictionCoefficient / setGroundFrictionTorqueIntegral / setContactFriction (:478-547) for
// link `body` of every world: the model is shared by the worlds of a batch, so the records of the batch's own model copy
// are re-derived on the host and uploaded; setMass also runs b2Body::SetMassData's sweep-centre update on every world.
b2g_status b2g_batch_set_link_property(b2g_batch* b, int body, int property, float value) {
    if (!b) return fail(B2G_ERR_INVALID, "null batch");
    if (!isLinkBody(b->m, body)) return fail(B2G_ERR_INVALID, "unknown link (body 0 is the ground)");
    if (property < 0 || property > B2G_LINK_CONTACT_FRICTION)
        return fail(B2G_ERR_INVALID, "this link property is derived (read-only) or unknown");
    if (const char* why = badLinkValue(property, value)) return fail(B2G_ERR_INVALID, why);
    ON_DEVICE(b->device);
    if (b->allSelected() && b->nSets() == 1) {   // one model for the whole batch: patch it in place
        bool massCenter = false;
        try {
            massCenter = setLinkProperty(b->m, body, property, value);
        } catch (const std::exception& ex) {
            return fail(B2G_ERR_INVALID, ex.what());
        }
        // stream order: behind whatever still reads the old blob, ahead of the sweep-centre kernel and everything after
        CU(cudaMemcpyAsync(b->blob, b->m.blob.data(), b->m.blob.size() * 4, cudaMemcpyHostToDevice, b->stream));
        if (massCenter) CU(launchMassCenter(b->L(), body));
        CU(cudaStreamSynchronize(b->stream));
        return B2G_OK;
    }
    // A selection of worlds gets its own parameter set.  The worlds of a 32-world tile share one set (every warp stages the
    // blob of its tile), so the selection has to cover whole tiles.
    if ((b->selFirst & 31) != 0 || ((b->selCount & 31) != 0 && b->selFirst + b->selCount != b->N))
        return fail(B2G_ERR_INVALID, "link parameters are kept per tile of 32 worlds: select a range that starts at a multiple of 32 and "
                    "ends at a multiple of 32 or at the last world (b2g_batch_select_worlds)");
    const size_t tiles = ((size_t)paddedTiles(b->N) * kTileWorlds + 31) / 32;   // parameter-set tiles are always 32 worlds
    if (b->tileSet.size() != tiles) b->tileSet.assign(tiles, 0);
    const size_t t0 = (size_t)(b->selFirst >> 5), t1 = (size_t)((b->selFirst + b->selCount + 31) >> 5);
    const size_t lastRealTile = (size_t)((b->N + 31) >> 5);
    std::vector<size_t> users(b->nSets(), 0), inside(b->nSets(), 0);
    for (size_t t = 0; t < lastRealTile; ++t) users[(size_t)b->tileSet[t]]++;
    for (size_t t = t0; t < t1; ++t) inside[(size_t)b->tileSet[t]]++;
    bool massCenter = false;
    std::vector<int> remap(b->nSets(), -1);
    const size_t setsBefore = b->nSets();
    try {
        for (size_t v = 0; v < setsBefore; ++v) {
            if (inside[v] == 0) continue;
            if (inside[v] == users[v] && v != 0) {   // every tile that uses this set is selected: patch the set itself
                massCenter = setLinkProperty(b->set((int)v), body, property, value) || massCenter;
                remap[v] = (int)v;
            } else {                                 // split: the selected tiles move to a patched copy (set 0 always stays the
                if (b->nSets() >= 65536) return fail(B2G_ERR_CAPACITY, "too many parameter sets");  // loaded-model default)
                auto copy = std::make_shared<HostModel>(b->set((int)v));
                massCenter = setLinkProperty(*copy, body, property, value) || massCenter;
                b->extraSets.push_back(copy);
                remap[v] = (int)b->nSets() - 1;
            }
        }
    } catch (const std::exception& ex) {
        b->extraSets.resize(setsBefore - 1);
        return fail(B2G_ERR_INVALID, ex.what());
    }
    for (size_t t = t0; t < t1; ++t) b->tileSet[t] = remap[(size_t)b->tileSet[t]];
    // device copies: all blobs (the buffer grows geometrically), the tile table
    CU(cudaStreamSynchronize(b->stream));
    if (b->nSets() > b->blobCapacitySets) {
        size_t cap = b->blobCapacitySets;
        while (cap < b->nSets()) cap *= 2;
        uint32_t* nb = nullptr;
        CU(cudaMalloc((void**)&nb, cap * (size_t)b->blobStride * 4));
        cudaFree(b->blob);
        b->blob = nb;
        b->blobCapacitySets = cap;
    }
    for (size_t v = 0; v < b->nSets(); ++v)
        CU(cudaMemcpyAsync(b->blob + v * (size_t)b->blobStride, b->set((int)v).blob.data(), b->set((int)v).blob.size() * 4,
                           cudaMemcpyHostToDevice, b->stream));
    if (!b->dTileSet) CU(cudaMalloc((void**)&b->dTileSet, tiles * sizeof(int)));
    CU(cudaMemcpyAsync(b->dTileSet, b->tileSet.data(), tiles * sizeof(int), cudaMemcpyHostToDevice, b->stream));
    if (massCenter) CU(launchMassCenter(b->L(), body));
    CU(cudaStreamSynchronize(b->stream));
    return B2G_OK;
}

// Names <-> indices: Box2DObject::getLink(name) / getJoint(name) / getLinks / getJoints (Box2DWorld.cpp:1411-1505) hand out
// entities by name; here links are body indices and joints are model joint indices, these calls translate.
b2g_status b2g_model_get_link(const b2g_model* m, int body, b2g_link_info* out) {
    if (!m || !out) return fail(B2G_ERR_INVALID, "null argument");
    if (!isLinkBody(m->m, body)) return fail(B2G_ERR_INVALID, "unknown link (body 0 is the ground)");
    const HostModel& hm = m->m;
    std::memset(out, 0, sizeof(*out));
    const std::string& full = hm.bodyNames[body];  // "<object>/<link>"
    const int object = (int)hm.blob[hm.h.oBody + body * MB_STRIDE + MB_OBJECT];
    const std::string link = full.substr(std::min(full.size(), hm.objects[object].name.size() + 1));
    std::snprintf(out->name, sizeof(out->name), "%s", link.c_str());
    out->object = object;
    out->is_base_link = (int)hm.blob[hm.h.oObj + object * MO_STRIDE + MO_BASE_BODY] == body;
    out->parent_joint = -1;
    for (int j = 0; j < hm.h.nj; ++j)
        if (hm.jointInts[j * 6] == 1 && hm.jointInts[j * 6 + 2] == body) out->parent_joint = j;
    out->n_fixtures = (int)hm.blob[hm.h.oBody + body * MB_STRIDE + MB_FIX_COUNT];
    out->is_static = hm.blob[hm.h.oBody + body * MB_STRIDE + MB_TYPE] != 2u;
    return B2G_OK;
}

b2g_status b2g_model_get_link_aabb(const b2g_model* m, int body, float out[4]) {
    if (!m || !out) return fail(B2G_ERR_INVALID, "null argument");
    if (!isLinkBody(m->m, body)) return fail(B2G_ERR_INVALID, "unknown link (body 0 is the ground)");
    for (int i = 0; i < 4; ++i) out[i] = m->m.linkAabb[4 * body + i];
    return B2G_OK;
}

b2g_status b2g_model_get_object_aabb(const b2g_model* m, int object, float out[4]) {
    if (!m || !out) return fail(B2G_ERR_INVALID, "null argument");
    if (object < 0 || object >= (int)m->m.objects.size()) return fail(B2G_ERR_INVALID, "unknown object");
    objectLocalAabb(m->m, object, out);
    return B2G_OK;
}

b2g_status b2g_model_find_link(const b2g_model* m, int object, const char* name, int32_t* body) {
    if (!m || !name || !body) return fail(B2G_ERR_INVALID, "null argument");
    const HostModel& hm = m->m;
    if (object < 0 || object >= (int)hm.objects.size()) return fail(B2G_ERR_INVALID, "unknown object");
    const std::string full = hm.objects[object].name + "/" + name;
    for (int b = 1; b < hm.h.nb; ++b)
        if (hm.bodyNames[b] == full && (int)hm.blob[hm.h.oBody + b * MB_STRIDE + MB_OBJECT] == object) {
            *body = b;
            return B2G_OK;
        }
    return fail(B2G_ERR_INVALID, "object '" + hm.objects[object].name + "' has no link '" + name + "'");
}

b2g_status b2g_model_get_joint(const b2g_model* m, int joint, b2g_joint_info* out) {
    if (!m || !out) return fail(B2G_ERR_INVALID, "null argument");
    const HostModel& hm = m->m;
    if (joint < 0 || joint >= hm.h.nj) return fail(B2G_ERR_INVALID, "unknown joint");
    std::memset(out, 0, sizeof(*out));
    std::snprintf(out->name, sizeof(out->name), "%s", hm.jointNames[joint].c_str());
    const uint32_t* jr = &hm.blob[hm.h.oJoint + joint * MJ_STRIDE];
    out->kind = (int32_t)jr[MJ_TYPE];
    out->object = (int32_t)jr[MJ_OBJECT];
    out->body_a = (int32_t)jr[MJ_BODYA];
    out->body_b = (int32_t)jr[MJ_BODYB];
    out->dof_index = (int32_t)jr[MJ_DOF];
    out->actuated = (hm.jointInts[joint * 6 + 4] != 0);
    out->limited = (hm.jointInts[joint * 6 + 5] != 0);
    return B2G_OK;
}

b2g_status b2g_model_find_joint(const b2g_model* m, int object, const char* name, int32_t* joint) {
    if (!m || !name || !joint) return fail(B2G_ERR_INVALID, "null argument");
    const HostModel& hm = m->m;
    if (object < 0 || object >= (int)hm.objects.size()) return fail(B2G_ERR_INVALID, "unknown object");
    for (int j = 0; j < hm.h.nj; ++j) {
        const uint32_t* jr = &hm.blob[hm.h.oJoint + j * MJ_STRIDE];
        if (jr[MJ_TYPE] == 1u && (int)jr[MJ_OBJECT] == object && hm.jointNames[j] == name) {
            *joint = j;
            return B2G_OK;
        }
    }
    return fail(B2G_ERR_INVALID, "object '" + hm.objects[object].name + "' has no joint '" + name + "'");
}

// the same on a model: batches created from it afterwards start with the changed parameters
b2g_status b2g_model_set_link_property(b2g_model* m, int body, int property, float value) {
    if (!m) return fail(B2G_ERR_INVALID, "null model");
    if (!isLinkBody(m->m, body)) return fail(B2G_ERR_INVALID, "unknown link (body 0 is the ground)");
    if (property < 0 || property > B2G_LINK_CONTACT_FRICTION)
        return fail(B2G_ERR_INVALID, "this link property is derived (read-only) or unknown");
    if (const char* why = badLinkValue(property, value)) return fail(B2G_ERR_INVALID, why);
    try {
        setLinkProperty(m->m, body, property, value);
    } catch (const std::exception& ex) {
        return fail(B2G_ERR_INVALID, ex.what());
    }
    return B2G_OK;
}

b2g_status b2g_model_get_link_property(const b2g_model* m, int body, int property, float* out) {
    if (!m || !out) return fail(B2G_ERR_INVALID, "null argument");
    if (!isLinkBody(m->m, body)) return fail(B2G_ERR_INVALID, "unknown link (body 0 is the ground)");
    if (property < 0 || property >= LP_COUNT) return fail(B2G_ERR_INVALID, "unknown link property");
    *out = getLinkProperty(m->m, body, property);
    return B2G_OK;
}

b2g_status b2g_batch_get_link_property(const b2g_batch* b, int body, int property, float* out) {
    if (!b || !out) return fail(B2G_ERR_INVALID, "null argument");
    if (!isLinkBody(b->m, body)) return fail(B2G_ERR_INVALID, "unknown link (body 0 is the ground)");
    if (property < 0 || property >= LP_COUNT) return fail(B2G_ERR_INVALID, "unknown link property");
    // several parameter sets: the value the first selected world uses
    const int v = b->nSets() > 1 && !b->tileSet.empty() ? b->tileSet[(size_t)(b->selFirst >> 5)] : 0;
    *out = getLinkProperty(b->set(v), body, property);
    return B2G_OK;
}

// stepPhysics(std::vector<Contact>&, steps, ctrl, sleep) (:3290-3310) incl. startRecordings (:2743-2768)
b2g_status b2g_batch_step_record(b2g_batch* b, int steps, int execute_controllers, int allow_sleeping, int32_t* counts,
                                 int32_t* iout, float* fout, int max_per_world) {
    if (!b || steps < 0 || !counts || !iout || !fout || max_per_world <= 0) return fail(B2G_ERR_INVALID, "bad arguments");
    if (!b->allSelected()) return fail(B2G_ERR_INVALID, "stepping and the state stack act on the whole batch: call b2g_batch_select_worlds(batch, 0, 0) first");
    ON_DEVICE(b->device);
    size_t cb = (size_t)b->n() * 4, ib = (size_t)b->n() * max_per_world * 4 * 4, fb = (size_t)b->n() * max_per_world * 6 * 4;
    size_t sb = (size_t)b->n() * b->m.h.nq * 4;
    CU(b->stageA.reserve(ib));
    CU(b->stageB.reserve(fb));
    CU(b->stageC.reserve(cb));
    CU(b->stageD.reserve(sb));
    CU(b->stageE.reserve(sb));
    CU(cudaMemsetAsync(b->stageA.p, 0, ib, b->stream));
    CU(cudaMemsetAsync(b->stageB.p, 0, fb, b->stream));
    CU(cudaMemsetAsync(b->stageC.p, 0, cb, b->stream));
    // startRecordings: saveState; setToRest; stepPhysics(1, false, false); BeginContact for every touching contact; restoreState
    CU(launchGetState(b->L(), (float*)b->stageD.p, (float*)b->stageE.p));
    CU(launchSetToRest(b->L()));
    CU(launchStep(b->L(), 0, nullptr, 1, 1, 0, 0));
    StepParams saved = b->prm;
    b->prm.recMode = 1;
    b->prm.recMax = max_per_world;
    b->prm.recCounts = (int*)b->stageC.p;
    b->prm.recI = (int*)b->stageA.p;
    b->prm.recF = (float*)b->stageB.p;
    cudaError_t e = launchRecordTouching(b->L());
    b->prm.recMode = 0;
    if (e == cudaSuccess) e = launchSetState(b->L(), (const float*)b->stageD.p, (const float*)b->stageE.p);
    b->prm.recMode = 1;
    if (e == cudaSuccess && steps > 0) e = launchStep(b->L(), 0, nullptr, 1, steps, execute_controllers, allow_sleeping);
    b->prm = saved;
    if (e != cudaSuccess) return cudaFail(e, "recorded stepping");
    CU(cudaMemcpyAsync(counts, b->stageC.p, cb, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaMemcpyAsync(iout, b->stageA.p, ib, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaMemcpyAsync(fout, b->stageB.p, fb, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    return B2G_OK;
}

// stepPhysics(callback, steps, ctrl, sleep) (:3312-3333) with the callback given as a body-pair table
b2g_status b2g_batch_step_guarded(b2g_batch* b, int steps, int execute_controllers, int allow_sleeping,
                                  const uint8_t* allowed_pairs, int32_t* completed, int32_t* steps_done) {
    if (!b || steps < 0 || !allowed_pairs || !completed) return fail(B2G_ERR_INVALID, "bad arguments");
    if (!b->allSelected()) return fail(B2G_ERR_INVALID, "stepping and the state stack act on the whole batch: call b2g_batch_select_worlds(batch, 0, 0) first");
    ON_DEVICE(b->device);
    const int nb = b->m.h.nb;
    size_t tb = (size_t)nb * nb, cb = (size_t)b->n() * 4;
    CU(b->stageA.reserve(tb));
    CU(cudaMemcpyAsync(b->stageA.p, allowed_pairs, tb, cudaMemcpyHostToDevice, b->stream));
    CU(launchClearAbort(b->L()));
    StepParams saved = b->prm;
    b->prm.recMode = 2;
    b->prm.allowed = (const unsigned char*)b->stageA.p;
    cudaError_t e = steps > 0 ? launchStepGuarded(b->L(), steps, execute_controllers, allow_sleeping) : cudaSuccess;
    b->prm = saved;
    if (e != cudaSuccess) return cudaFail(e, "guarded stepping");
    // the reference returns `not abort_prematurely`: a world that aborted on its very last step still reports false, so the
    // abort flag itself is read back (bit 0) next to the number of executed steps (bits 8..)
    std::vector<int32_t> status((size_t)b->n());
    CU(b->stageB.reserve(cb));
    CU(launchAbortFlags(b->L(), (int*)b->stageB.p));
    CU(cudaMemcpyAsync(status.data(), b->stageB.p, cb, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    for (long long i = 0; i < b->n(); ++i) {
        completed[i] = (status[i] & 1) ? 0 : 1;
        if (steps_done) steps_done[i] = status[i] >> 8;
    }
    return B2G_OK;
}

b2g_status b2g_batch_take_kernel_ms(b2g_batch* b, float* out_ms, int32_t* out_launches) {
    if (!b || !out_ms) return fail(B2G_ERR_INVALID, "null argument");
    ON_DEVICE(b->device);
    CU(cudaStreamSynchronize(b->stream));
    float total = 0.0f;
    for (auto& p : b->timed) {
        float ms = 0.0f;
        CU(cudaEventElapsedTime(&ms, p.first, p.second));
        total += ms;
        b->eventPool.push_back(p.first);
        b->eventPool.push_back(p.second);
    }
    if (out_launches) *out_launches = (int32_t)b->timed.size();
    b->timed.clear();
    *out_ms = total;
    return B2G_OK;
}

b2g_status b2g_batch_get_bodies(b2g_batch* b, float* out) {
    if (!b || !out) return fail(B2G_ERR_INVALID, "null argument");
    ON_DEVICE(b->device);
    size_t bytes = (size_t)b->n() * b->m.h.nb * 16 * 4;
    CU(b->stageA.reserve(bytes));
    CU(launchGetBodies(b->L(), (float*)b->stageA.p));
    CU(cudaMemcpyAsync(out, b->stageA.p, bytes, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    return B2G_OK;
}

// Box2DLink::getPose / getVelocityVector / getCenterOfMass of every link of every (selected) world
b2g_status b2g_batch_get_link_states(b2g_batch* b, float* out) {
    if (!b || !out) return fail(B2G_ERR_INVALID, "null argument");
    ON_DEVICE(b->device);
    size_t bytes = (size_t)b->n() * b->m.h.nb * 8 * 4;
    CU(b->stageA.reserve(bytes));
    CU(cudaMemsetAsync(b->stageA.p, 0, bytes, b->stream));  // the ground's rows
    CU(launchGetLinks(b->L(), (float*)b->stageA.p));
    CU(cudaMemcpyAsync(out, b->stageA.p, bytes, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    return B2G_OK;
}

b2g_status b2g_batch_get_fat_aabbs(b2g_batch* b, float* out) {
    if (!b || !out) return fail(B2G_ERR_INVALID, "null argument");
    ON_DEVICE(b->device);
    size_t bytes = (size_t)b->n() * b->m.h.nf * 4 * 4;
    CU(b->stageA.reserve(bytes));
    CU(launchGetFat(b->L(), (float*)b->stageA.p));
    CU(cudaMemcpyAsync(out, b->stageA.p, bytes, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    return B2G_OK;
}

b2g_status b2g_batch_get_ball_approximation(b2g_batch* b, int object, float* out) {
    if (!b || object < 0 || object >= b->m.h.no) return fail(B2G_ERR_INVALID, "unknown object");
    const HostObjectInfo& o = b->m.objects[object];
    if (o.nBalls == 0) return B2G_OK;
    if (!out) return fail(B2G_ERR_INVALID, "null argument");
    ON_DEVICE(b->device);
    if (!b->balls) {  // the ball table of the model, uploaded on first use
        CU(cudaMalloc((void**)&b->balls, b->m.balls.size() * sizeof(float)));
        CU(cudaMemcpyAsync(b->balls, b->m.balls.data(), b->m.balls.size() * sizeof(float), cudaMemcpyHostToDevice, b->stream));
    }
    size_t bytes = (size_t)b->n() * o.nBalls * 4 * 4;
    CU(b->stageA.reserve(bytes));
    CU(launchGetBalls(b->L(), b->balls, o.ballStart, o.nBalls, (float*)b->stageA.p));
    CU(cudaMemcpyAsync(out, b->stageA.p, bytes, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    return B2G_OK;
}

b2g_status b2g_batch_get_joints(b2g_batch* b, int32_t* iout, float* fout) {
    if (!b || !iout || !fout) return fail(B2G_ERR_INVALID, "null argument");
    ON_DEVICE(b->device);
    size_t ib = (size_t)b->n() * b->m.h.nj * 6 * 4, fb = (size_t)b->n() * b->m.h.nj * 15 * 4;
    CU(b->stageA.reserve(ib));
    CU(b->stageB.reserve(fb));
    CU(launchGetJoints(b->L(), (int*)b->stageA.p, (float*)b->stageB.p));
    CU(cudaMemcpyAsync(iout, b->stageA.p, ib, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaMemcpyAsync(fout, b->stageB.p, fb, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    return B2G_OK;
}

b2g_status b2g_batch_get_contacts(b2g_batch* b, int32_t* counts, int32_t* iout, float* fout, int max_per_world) {
    if (!b || !counts || !iout || !fout || max_per_world <= 0) return fail(B2G_ERR_INVALID, "bad arguments");
    ON_DEVICE(b->device);
    size_t cb = (size_t)b->n() * 4, ib = (size_t)b->n() * max_per_world * 8 * 4, fb = (size_t)b->n() * max_per_world * 12 * 4;
    CU(b->stageA.reserve(ib));
    CU(b->stageB.reserve(fb));
    CU(b->stageC.reserve(cb));
    CU(cudaMemsetAsync(b->stageA.p, 0, ib, b->stream));
    CU(cudaMemsetAsync(b->stageB.p, 0, fb, b->stream));
    CU(launchGetContacts(b->L(), (int*)b->stageC.p, (int*)b->stageA.p, (float*)b->stageB.p, max_per_world));
    CU(cudaMemcpyAsync(counts, b->stageC.p, cb, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaMemcpyAsync(iout, b->stageA.p, ib, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaMemcpyAsync(fout, b->stageB.p, fb, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    return B2G_OK;
}

b2g_status b2g_batch_check_collision_dev(b2g_batch* b, int32_t* counts, int32_t* iout, float* fout, int max_per_world) {
    if (!b || !counts || max_per_world < 0) return fail(B2G_ERR_INVALID, "bad arguments");
    if (!iout || !fout) max_per_world = 0;
    ON_DEVICE(b->device);
    CU(launchCheckCollision(b->L(), counts, iout, fout, max_per_world));
    return B2G_OK;
}

b2g_status b2g_batch_check_collision(b2g_batch* b, int32_t* counts, int32_t* iout, float* fout, int max_per_world) {
    if (!b || !counts || !iout || !fout || max_per_world <= 0) return fail(B2G_ERR_INVALID, "bad arguments");
    ON_DEVICE(b->device);
    size_t cb = (size_t)b->n() * 4, ib = (size_t)b->n() * max_per_world * 4 * 4, fb = (size_t)b->n() * max_per_world * 6 * 4;
    CU(b->stageA.reserve(ib));
    CU(b->stageB.reserve(fb));
    CU(b->stageC.reserve(cb));
    CU(cudaMemsetAsync(b->stageA.p, 0, ib, b->stream));
    CU(cudaMemsetAsync(b->stageB.p, 0, fb, b->stream));
    CU(launchCheckCollision(b->L(), (int*)b->stageC.p, (int*)b->stageA.p, (float*)b->stageB.p, max_per_world));
    CU(cudaMemcpyAsync(counts, b->stageC.p, cb, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaMemcpyAsync(iout, b->stageA.p, ib, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaMemcpyAsync(fout, b->stageB.p, fb, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    return B2G_OK;
}

// ---------------------------------------------------------------------------------------------- state stack
b2g_status b2g_batch_save_state(b2g_batch* b) {
    if (!b) return fail(B2G_ERR_INVALID, "null batch");
    if (!b->allSelected()) return fail(B2G_ERR_INVALID, "stepping and the state stack act on the whole batch: call b2g_batch_select_worlds(batch, 0, 0) first");
    ON_DEVICE(b->device);
    SavedState s;
    size_t bytes = (size_t)b->n() * b->m.h.nq * 4;
    CU(cudaMalloc((void**)&s.q, bytes));
    CU(cudaMalloc((void**)&s.qd, bytes));
    CU(launchGetState(b->L(), s.q, s.qd));
    for (int o = 0; o < b->m.h.no; ++o) {
        if (!b->m.objects[o].isStatic) continue;
        float* p = nullptr;
        CU(cudaMalloc((void**)&p, (size_t)b->n() * 3 * 4));
        CU(launchGetPose(b->L(), o, p));
        s.staticPoses.emplace_back(o, p);
    }
    b->stack.push_back(s);
    CU(cudaStreamSynchronize(b->stream));
    return B2G_OK;
}

b2g_status b2g_batch_drop_state(b2g_batch* b) {
    if (!b) return fail(B2G_ERR_INVALID, "null batch");
    if (b->stack.empty()) return B2G_OK;
    ON_DEVICE(b->device);
    SavedState s = b->stack.back();
    b->stack.pop_back();
    cudaFree(s.q);
    cudaFree(s.qd);
    for (auto& p : s.staticPoses) cudaFree(p.second);
    return B2G_OK;
}

b2g_status b2g_batch_restore_state(b2g_batch* b) {
    if (!b) return fail(B2G_ERR_INVALID, "null batch");
    if (!b->allSelected()) return fail(B2G_ERR_INVALID, "stepping and the state stack act on the whole batch: call b2g_batch_select_worlds(batch, 0, 0) first");
    if (b->stack.empty()) return fail(B2G_ERR_INVALID, "state stack is empty");  // reference returns false (:3526-3528)
    ON_DEVICE(b->device);
    SavedState& s = b->stack.back();
    for (auto& p : s.staticPoses) CU(launchSetPose(b->L(), p.first, p.second));
    CU(launchSetState(b->L(), s.q, s.qd));
    CU(cudaStreamSynchronize(b->stream));
    return b2g_batch_drop_state(b);
}

// ------------------------------------------------------------------------------------------------------ misc
b2g_status b2g_host_alloc(void** out, uint64_t bytes) {
    if (!out) return fail(B2G_ERR_INVALID, "null argument");
    CU(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return B2G_OK;
}
void b2g_host_free(void* p) {
    if (p) cudaFreeHost(p);
}
uint64_t b2g_launch_count(void) { return launchCount(); }

}  // extern "C"

// ------------------------------------------------------------------------------------------------ several GPUs
// One process, one host thread and one b2g_batch per device, contiguous world slices, ncclAllGather of the final states on
// the batches' streams.  NCCL is loaded with dlopen on first use so that the library itself has no link-time
// dependency on it (the single-GPU path and the CPU-side tests never touch it).
namespace {
struct Nccl {
    void* lib = nullptr;
    decltype(&ncclCommInitAll) commInitAll = nullptr;
    decltype(&ncclCommDestroy) commDestroy = nullptr;
    decltype(&ncclAllGather) allGather = nullptr;
    decltype(&ncclGroupStart) groupStart = nullptr;
    decltype(&ncclGroupEnd) groupEnd = nullptr;
    decltype(&ncclGetErrorString) errorString = nullptr;
    bool load(std::string& why) {
        if (lib) return true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) {
            why = std::string("cannot load libnccl.so.2: ") + dlerror();
            return false;
        }
#define B2G_SYM(field, symbol)                                       \
    field = reinterpret_cast<decltype(field)>(dlsym(lib, symbol));   \
    if (!field) {                                                    \
        why = std::string("libnccl has no symbol ") + symbol;        \
        lib = nullptr;                                               \
        return false;                                                \
    }
        B2G_SYM(commInitAll, "ncclCommInitAll")
        B2G_SYM(commDestroy, "ncclCommDestroy")
        B2G_SYM(allGather, "ncclAllGather")
        B2G_SYM(groupStart, "ncclGroupStart")
        B2G_SYM(groupEnd, "ncclGroupEnd")
        B2G_SYM(errorString, "ncclGetErrorString")
#undef B2G_SYM
        return true;
    }
};
Nccl g_nccl;

struct Shard {
    b2g_batch* batch = nullptr;
    int device = 0;
    long long first = 0, count = 0;
    float *q = nullptr, *qd = nullptr, *tg = nullptr;  // this slice's inputs on the device
    float *sendQ = nullptr, *sendQd = nullptr;        // [rowsPer][nq], padded
    float *allQ = nullptr, *allQd = nullptr;          // [nDev][rowsPer][nq]
    size_t tgCap = 0;
    cudaEvent_t g0 = nullptr, g1 = nullptr;
    float kernelMs = 0.0f, gatherMs = 0.0f;
    b2g_status status = B2G_OK;
    std::string error;
};
}  // namespace

struct b2g_multi {
    std::vector<Shard> shards;
    std::vector<ncclComm_t> comms;
    long long N = 0, rowsPer = 0;
    int nq = 0;
};

extern "C" {

void b2g_multi_destroy(b2g_multi* m) {
    if (!m) return;
    for (size_t i = 0; i < m->shards.size(); ++i) {
        Shard& s = m->shards[i];
        DeviceGuard guard(s.device);
        if (s.batch) b2g_batch_sync(s.batch);
        if (i < m->comms.size() && m->comms[i] && g_nccl.commDestroy) g_nccl.commDestroy(m->comms[i]);
        for (float* p : {s.q, s.qd, s.tg, s.sendQ, s.sendQd, s.allQ, s.allQd}) cudaFree(p);
        if (s.g0) cudaEventDestroy(s.g0);
        if (s.g1) cudaEventDestroy(s.g1);
        if (s.batch) b2g_batch_destroy(s.batch);
    }
    delete m;
}

b2g_status b2g_multi_create(const b2g_model* model, int64_t n_total, const int32_t* devices, int n_devices, int max_contacts,
                            b2g_multi** out) {
    if (!model || !out || !devices || n_devices < 1 || n_total < n_devices) return fail(B2G_ERR_INVALID, "bad multi-GPU batch arguments");
    for (int i = 0; i < n_devices; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) return fail(B2G_ERR_INVALID, "every device may appear once (NCCL: one rank per GPU)");
    std::string why;
    if (!g_nccl.load(why)) return fail(B2G_ERR_UNSUPPORTED, why);
    b2g_multi* m = new b2g_multi();
    m->N = n_total;
    m->shards.resize(n_devices);
    m->rowsPer = (n_total + n_devices - 1) / n_devices;
    // contiguous slices that differ by at most one world (the rule of box2d_sim_env_b200/sharding.py: world_slice)
    const long long base = n_total / n_devices, rem = n_total % n_devices;
    long long first = 0;
    for (int i = 0; i < n_devices; ++i) {
        Shard& s = m->shards[i];
        s.device = devices[i];
        s.first = first;
        s.count = base + (i < rem ? 1 : 0);
        first += s.count;
        b2g_status st = b2g_batch_create(model, s.count, s.device, max_contacts, &s.batch);
        if (st) {
            const std::string msg = g_error;
            b2g_multi_destroy(m);
            return fail(st, msg);
        }
        m->nq = s.batch->m.h.nq;
        DeviceGuard guard(s.device);
        const size_t slice = (size_t)s.count * m->nq * 4, padded = (size_t)m->rowsPer * m->nq * 4;
        cudaError_t e = cudaMalloc((void**)&s.q, slice);
        if (e == cudaSuccess) e = cudaMalloc((void**)&s.qd, slice);
        if (e == cudaSuccess) e = cudaMalloc((void**)&s.sendQ, padded);
        if (e == cudaSuccess) e = cudaMalloc((void**)&s.sendQd, padded);
        if (e == cudaSuccess) e = cudaMalloc((void**)&s.allQ, padded * n_devices);
        if (e == cudaSuccess) e = cudaMalloc((void**)&s.allQd, padded * n_devices);
        if (e == cudaSuccess) e = cudaMemset(s.sendQ, 0, padded);
        if (e == cudaSuccess) e = cudaMemset(s.sendQd, 0, padded);
        if (e == cudaSuccess) e = cudaEventCreate(&s.g0);
        if (e == cudaSuccess) e = cudaEventCreate(&s.g1);
        if (e != cudaSuccess) {
            b2g_multi_destroy(m);
            return cudaFail(e, "multi-GPU buffers");
        }
    }
    m->comms.assign(n_devices, nullptr);
    ncclResult_t r = g_nccl.commInitAll(m->comms.data(), n_devices, devices);
    if (r != ncclSuccess) {
        const std::string msg = std::string("ncclCommInitAll: ") + g_nccl.errorString(r);
        b2g_multi_destroy(m);
        return fail(B2G_ERR_CUDA, msg);
    }
    *out = m;
    return B2G_OK;
}

int32_t b2g_multi_device_count(const b2g_multi* m) { return m ? (int32_t)m->shards.size() : 0; }

b2g_status b2g_multi_get_shard(b2g_multi* m, int i, b2g_batch** batch, int64_t* first_world, int64_t* n_worlds) {
    if (!m || i < 0 || i >= (int)m->shards.size()) return fail(B2G_ERR_INVALID, "unknown shard");
    if (batch) *batch = m->shards[i].batch;
    if (first_world) *first_world = m->shards[i].first;
    if (n_worlds) *n_worlds = m->shards[i].count;
    return B2G_OK;
}

b2g_status b2g_multi_rollout(b2g_multi* m, int object, const float* q, const float* qd, int reset_caches, const float* targets,
                             int n_targets, int steps_per_target, float* q_out, float* qd_out) {
    if (!m || !targets || n_targets < 1 || steps_per_target < 0 || !q_out || !qd_out || (q == nullptr) != (qd == nullptr))
        return fail(B2G_ERR_INVALID, "bad rollout arguments");
    const int nDev = (int)m->shards.size();
    const int nq = m->nq;
    {
        b2g_status s = checkRobot(m->shards[0].batch, object);
        if (s) return s;
    }
    const int nd = m->shards[0].batch->m.objects[object].nDofs;
    auto work = [&](int i) {
        Shard& s = m->shards[i];
        b2g_batch* b = s.batch;
        s.status = B2G_OK;
        auto bad = [&](cudaError_t e, const char* what) {
            s.status = B2G_ERR_CUDA;
            s.error = std::string(what) + ": " + cudaGetErrorString(e);
        };
        cudaError_t e = cudaSetDevice(s.device);
        if (e != cudaSuccess) return bad(e, "cudaSetDevice");
        cudaStream_t st = b->stream;
        const size_t slice = (size_t)s.count * nq * 4;
        if (q) {
            e = cudaMemcpyAsync(s.q, q + (size_t)s.first * nq, slice, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(s.qd, qd + (size_t)s.first * nq, slice, cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) return bad(e, "state H2D");
            if ((s.status = b2g_batch_set_state_dev(b, s.q, s.qd, reset_caches)) != B2G_OK) {
                s.error = g_error;
                return;
            }
        }
        const size_t tgBytes = (size_t)n_targets * s.count * nd * 4;
        if (tgBytes > s.tgCap) {
            cudaFree(s.tg);
            s.tg = nullptr;
            s.tgCap = 0;
            if ((e = cudaMalloc((void**)&s.tg, tgBytes)) != cudaSuccess) return bad(e, "cudaMalloc(targets)");
            s.tgCap = tgBytes;
        }
        // rows first .. first + count of every targets[t]: one strided copy
        e = cudaMemcpy2DAsync(s.tg, (size_t)s.count * nd * 4, targets + (size_t)s.first * nd, (size_t)m->N * nd * 4,
                              (size_t)s.count * nd * 4, (size_t)n_targets, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return bad(e, "targets H2D");
        float ms = 0.0f;
        int32_t launches = 0;
        b2g_batch_take_kernel_ms(b, &ms, &launches);  // forget earlier launches
        if ((s.status = b2g_batch_rollout_dev(b, object, s.tg, n_targets, steps_per_target)) != B2G_OK ||
            (s.status = b2g_batch_get_state_dev(b, s.sendQ, s.sendQd)) != B2G_OK) {
            s.error = g_error;
            return;
        }
        // the path's only exchange: every device receives every slice
        e = cudaEventRecord(s.g0, st);
        if (e != cudaSuccess) return bad(e, "cudaEventRecord");
        const size_t rowsBytes = (size_t)m->rowsPer * nq;
        ncclResult_t r = g_nccl.groupStart();
        if (r == ncclSuccess) r = g_nccl.allGather(s.sendQ, s.allQ, rowsBytes, ncclFloat, m->comms[i], st);
        if (r == ncclSuccess) r = g_nccl.allGather(s.sendQd, s.allQd, rowsBytes, ncclFloat, m->comms[i], st);
        if (r == ncclSuccess) r = g_nccl.groupEnd();
        if (r != ncclSuccess) {
            s.status = B2G_ERR_CUDA;
            s.error = std::string("ncclAllGather: ") + g_nccl.errorString(r);
            return;
        }
        e = cudaEventRecord(s.g1, st);
        if (e == cudaSuccess && i == 0) {  // device 0 hands the gathered states to the caller, slice by slice (padding skipped)
            for (int j = 0; j < nDev && e == cudaSuccess; ++j) {
                const Shard& o = m->shards[j];
                const size_t bytes = (size_t)o.count * nq * 4;
                e = cudaMemcpyAsync(q_out + (size_t)o.first * nq, s.allQ + (size_t)j * rowsBytes, bytes, cudaMemcpyDeviceToHost, st);
                if (e == cudaSuccess)
                    e = cudaMemcpyAsync(qd_out + (size_t)o.first * nq, s.allQd + (size_t)j * rowsBytes, bytes, cudaMemcpyDeviceToHost, st);
            }
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return bad(e, "gather / D2H");
        cudaEventElapsedTime(&s.gatherMs, s.g0, s.g1);
        b2g_batch_take_kernel_ms(b, &s.kernelMs, &launches);
    };
    if (nDev == 1) {
        DeviceGuard guard(m->shards[0].device);
        work(0);
    } else {
        std::vector<std::thread> threads;
        for (int i = 0; i < nDev; ++i) threads.emplace_back(work, i);
        for (auto& t : threads) t.join();
    }
    for (const Shard& s : m->shards)
        if (s.status != B2G_OK) return fail(s.status, "device " + std::to_string(s.device) + ": " + s.error);
    return B2G_OK;
}

b2g_status b2g_multi_gathered_dev(b2g_multi* m, int i, const float** q_all, const float** qd_all, int64_t* rows_per_device) {
    if (!m || i < 0 || i >= (int)m->shards.size()) return fail(B2G_ERR_INVALID, "unknown shard");
    if (q_all) *q_all = m->shards[i].allQ;
    if (qd_all) *qd_all = m->shards[i].allQd;
    if (rows_per_device) *rows_per_device = m->rowsPer;
    return B2G_OK;
}

b2g_status b2g_multi_last_timing(b2g_multi* m, float* kernel_ms_min, float* kernel_ms_max, float* gather_ms) {
    if (!m) return fail(B2G_ERR_INVALID, "null argument");
    float lo = 1e30f, hi = 0.0f, g = 0.0f;
    for (const Shard& s : m->shards) {
        lo = std::min(lo, s.kernelMs);
        hi = std::max(hi, s.kernelMs);
        g = std::max(g, s.gatherMs);
    }
    if (kernel_ms_min) *kernel_ms_min = lo;
    if (kernel_ms_max) *kernel_ms_max = hi;
    if (gather_ms) *gather_ms = g;
    return B2G_OK;
}

}  // extern "C"
