/* b2g.h — C ABI of the B200-native batched Box2D backend (libb2g.so).
 *
 * This is the drop-in boundary for the hot path of JoshuaHaustein/box2d_sim_env: every entry point
 * replaces one (family of) sim_env::World / Box2DWorld method(s), applied to N independent worlds at
 * once.  Citations are file:line in /root/reference.  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions
 *   - All functions return b2g_status (0 = ok).  b2g_last_error() gives the message of the last failure
 *     on the calling thread (the reference throws std::logic_error / std::runtime_error instead,
 *     e.g. src/sim_env/Box2DWorld.cpp:1700-1704).
 *   - Arrays are row-major with the world index leading: q[n_worlds][nq].
 *   - Units are the reference's user units (metres, radians, seconds); the `scale` of the YAML file
 *     is applied internally exactly as Box2DWorld does (include/sim_env/Box2DWorld.h:829-841).
 *   - "_dev" variants take device pointers (on the batch's device) and are asynchronous on the batch's
 *     CUDA stream; the plain variants take host pointers, copy, and synchronise.
 *   - A batch handle is not thread-safe (the reference serialises every call on one recursive mutex,
 *     include/sim_env/Box2DWorld.h:852).
 *   - There is no CPU fallback: every compute entry point fails with B2G_ERR_CUDA when no usable
 *     sm_100 device is present.
 */
#ifndef B2G_H_
#define B2G_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t b2g_status;
enum {
    B2G_OK = 0,
    B2G_ERR_INVALID = 1,   /* bad argument / size mismatch (reference: std::runtime_error) */
    B2G_ERR_IO = 2,        /* file missing or YAML not decodable */
    B2G_ERR_CUDA = 3,      /* CUDA runtime failure or no device */
    B2G_ERR_UNSUPPORTED = 4, /* e.g. prismatic joints (reference throws too: Box2DWorld.cpp:1204-1208) */
    B2G_ERR_CAPACITY = 5   /* a world exceeded its contact-slot capacity (see b2g_batch_status) */
};

typedef struct b2g_env b2g_env;     /* mutable environment description (Box2DEnvironmentDescription) */
typedef struct b2g_model b2g_model; /* immutable flattened model shared by all worlds of a batch */
typedef struct b2g_batch b2g_batch; /* N worlds resident in HBM on one device */

const char* b2g_last_error(void);

/* ---------------------------------------------------------------------------------------------------
 * Environment description.  Replaces parseYAML (src/sim_env/Box2DIOUtils.cpp:63-92) and the description
 * structs of include/sim_env/Box2DIOUtils.h:18-70.  The YAML reader accepts both the decoder's key names
 * (ground_friction, ground_torque_integral; Box2DIOUtils.h:133-134) and the stale names of the shipped
 * data (trans_friction, rot_friction; test_data/robots/test_robot.yaml:14-15) and ignores unknown keys.
 * ------------------------------------------------------------------------------------------------- */
b2g_status b2g_env_create(float scale, const float world_bounds[4], b2g_env** out);
b2g_status b2g_env_load_yaml(const char* path, b2g_env** out);
void b2g_env_destroy(b2g_env* env);
/* base_limits8 = translational velocity(2), translational acceleration(2), rotational velocity(2),
 * rotational acceleration(2); NULL for plain objects.  Returns the object index through *out_index
 * (robots and objects are indexed separately, file order). */
b2g_status b2g_env_add_object(b2g_env* env, const char* name, int is_robot, int is_static,
                              const float* base_limits8, int use_center_of_mass, int* out_index);
b2g_status b2g_env_add_link(b2g_env* env, int is_robot, int object, const char* name, float mass,
                            float ground_friction, float ground_torque_integral, float contact_friction,
                            float restitution, int* out_index);
b2g_status b2g_env_add_polygon(b2g_env* env, int is_robot, int object, int link, const float* xy, int n_floats);
/* one ball of the link's ball approximation: centre in the link frame (user units) and radius
 * (Box2DLinkDescription::balls, Box2DIOUtils.h:21, 139-142; kept by the Box2DLink ctor, Box2DWorld.cpp:54-57) */
b2g_status b2g_env_add_ball(b2g_env* env, int is_robot, int object, int link, const float center[3], float radius);
/* params9 = axis(2), position_limits(2), velocity_limits(2), acceleration_limits(2), axis_orientation */
b2g_status b2g_env_add_joint(b2g_env* env, int is_robot, int object, const char* name, const char* link_a,
                             const char* link_b, const float* params9, int actuated, const char* joint_type);
b2g_status b2g_env_add_state(b2g_env* env, const char* name, const float* configuration, const float* velocity, int n);

/* ---------------------------------------------------------------------------------------------------
 * Model.  Replaces Box2DWorld::createWorld / createGround and the Box2DLink / Box2DJoint / Box2DObject /
 * Box2DRobot constructors (src/sim_env/Box2DWorld.cpp:3664-3721, 3644-3662, 44-148, 872-948, 1292-1363,
 * 2046-2067): polygon hulls, densities, mass data, friction joints to the ground, revolute joints,
 * creation order, DOF indexing.
 * ------------------------------------------------------------------------------------------------- */
b2g_status b2g_model_create(const b2g_env* env, b2g_model** out);
void b2g_model_destroy(b2g_model* model);

typedef struct b2g_model_info {
    int32_t n_bodies;    /* incl. the ground body (index 0) */
    int32_t n_fixtures;  /* broad-phase proxies incl. the ground sensor (index 0) */
    int32_t n_joints;    /* friction + revolute joints, creation order */
    int32_t n_objects;   /* robots first, then objects (creation order) */
    int32_t n_robots;
    int32_t nq;          /* DOFs of the world state vector (Box2DObject::getDOFPositions over all objects) */
    float scale;
} b2g_model_info;
b2g_status b2g_model_get_info(const b2g_model* model, b2g_model_info* out);
/* What the reference reports through its logger while loading (logWarn in parseYAML / createWorld, e.g. "No scale
 * specified", Box2DIOUtils.cpp:70-84) plus this loader's own notes (stale key names accepted as aliases): the messages,
 * one per line, copied into buf (NUL-terminated, truncated to cap); *n_out = number of messages. */
b2g_status b2g_model_get_warnings(const b2g_model* model, char* buf, int cap, int32_t* n_out);

typedef struct b2g_object_info {
    char name[64];
    int32_t is_robot, is_static;
    int32_t n_dofs;      /* 3 base DOFs unless static + one per joint (Box2DWorld.cpp:1314-1331) */
    int32_t dof_offset;  /* first index in the world state vector */
    int32_t first_body, n_links, n_joints;
    int32_t n_balls;     /* Box2DObject::_num_balls (Box2DWorld.cpp:1305) */
} b2g_object_info;
b2g_status b2g_model_get_object(const b2g_model* model, int object, b2g_object_info* out);
/* DOF limits in world-state order: out[nq][6] = position lo/hi, velocity lo/hi, acceleration lo/hi
 * (Box2DObject::getDOFInformation, Box2DWorld.cpp:1557-1583) */
b2g_status b2g_model_get_dof_limits(const b2g_model* model, float* out);
/* per body [8]: mass invMass I invI localCenter(2) type n_fixtures  (parity probe; b2Body mass data) */
b2g_status b2g_model_get_bodies(const b2g_model* model, float* out);
/* per fixture: iout[3] = body, vertex count, sensor; fout[38] = vertices(16) normals(16) centroid(2)
 * friction restitution density radius  (b2PolygonShape::Set results) */
b2g_status b2g_model_get_fixtures(const b2g_model* model, int32_t* iout, float* fout);
/* per joint: iout[6] = kind(0 friction,1 revolute) bodyA bodyB 0 enableMotor enableLimit;
 * fout[15] = 0 0 0 0 0 0 maxForce maxTorque anchorA(2) anchorB(2) referenceAngle lower upper */
b2g_status b2g_model_get_joints(const b2g_model* model, int32_t* iout, float* fout);
/* Box2DLink::getLocalBallApproximation (Box2DWorld.cpp:727-731) for every link of the object, links in name order:
 * bodies[n_balls] = body index of the ball's link, xyzr[n_balls][4] = centre in the link frame and radius */
b2g_status b2g_model_get_balls(const b2g_model* model, int object, int32_t* bodies, float* xyzr);

/* Names <-> indices.  The reference hands out links and joints by name (Box2DObject::getLink / getJoint / getLinks /
 * getJoints, Box2DWorld.cpp:1411-1505; Contact::link_a / link_b are LinkPtrs); here a link is its body index (what contacts
 * report) and a joint its model joint index (b2g_model_get_joints order). */
typedef struct b2g_link_info {
    char name[64];         /* the link's own name (YAML `name`) */
    int32_t object;        /* owning object (b2g_model_get_object index) */
    int32_t is_base_link;
    int32_t is_static;
    int32_t parent_joint;  /* model joint index of the revolute joint whose child link this is; -1 for the base link */
    int32_t n_fixtures;
} b2g_link_info;
b2g_status b2g_model_get_link(const b2g_model* model, int body, b2g_link_info* out);
b2g_status b2g_model_find_link(const b2g_model* model, int object, const char* name, int32_t* body);
/* Box2DLink::getLocalBoundingBox (Box2DWorld.cpp:258-261; the box of the link's polygons as given in the description) and
 * Box2DObject::getLocalAABB (:1928-1931; the link boxes merged, shifted by the base link's local centre of mass, :1304-1312):
 * out = min x, min y, max x, max y in user units */
b2g_status b2g_model_get_link_aabb(const b2g_model* model, int body, float out[4]);
b2g_status b2g_model_get_object_aabb(const b2g_model* model, int object, float out[4]);
typedef struct b2g_joint_info {
    char name[64];         /* YAML name; empty for the ground-friction joint of a link */
    int32_t kind;          /* 0 ground friction, 1 revolute, 2 prismatic */
    int32_t object;
    int32_t body_a, body_b;  /* parent / child link (ground friction: 0 / the link) */
    int32_t dof_index;     /* index in the world state vector, -1 for ground-friction joints */
    int32_t actuated, limited;
} b2g_joint_info;
b2g_status b2g_model_get_joint(const b2g_model* model, int joint, b2g_joint_info* out);
b2g_status b2g_model_find_joint(const b2g_model* model, int object, const char* name, int32_t* joint);

/* ---------------------------------------------------------------------------------------------------
 * Batch of worlds.  b2g_batch_create == N x Box2DWorld::loadWorld (Box2DWorld.cpp:3040-3054): every world
 * starts in the state createWorld leaves (bodies placed, YAML `states` applied, nothing stepped).
 * max_contacts <= 0 selects a default from the model.
 * ------------------------------------------------------------------------------------------------- */
b2g_status b2g_batch_create(const b2g_model* model, int64_t n_worlds, int device, int max_contacts, b2g_batch** out);
void b2g_batch_destroy(b2g_batch* batch);
/* Box2DWorld::clone (Box2DWorld.cpp:3021-3038) for the whole batch: same environment, same time step / iteration counts,
 * setWorldState(getWorldState()) incl. the poses of static objects; fresh caches like a freshly loaded world. */
b2g_status b2g_batch_clone(b2g_batch* batch, b2g_batch** out);
/* setPhysicsTimeStep / setVelocitySteps / setPositionSteps (Box2DWorld.cpp:3340-3353); defaults 0.01, 15, 15 */
/* Restrict the per-world calls below (state / pose / target get and set, collision and region queries, ball
 * approximation, probes, status) to worlds [first, first + count): their arrays then have `count` rows.  This is what a
 * single-world view (BatchBox2DWorld::world(i) in b2g_world.hpp, the reference's one-world Box2DWorld signatures) is built
 * on.  count <= 0 selects the whole batch again.  Stepping and the save/restore stack always act on the whole batch and
 * fail with B2G_ERR_INVALID while a sub-range is selected. */
b2g_status b2g_batch_select_worlds(b2g_batch* batch, int64_t first, int64_t count);
b2g_status b2g_batch_set_params(b2g_batch* batch, float dt, int velocity_steps, int position_steps);
/* per-world status after the last call: 0 ok, bit0 contact capacity exceeded (-> B2G_ERR_CAPACITY), bit1 non-finite state,
 * bit2 (a note, not an error) a controller target lay outside the robot's velocity limits and was scaled: the scaling rule of
 * sim_env's utils::eigen::scaleToLimits is an assumption here (one uniform factor), so results of such worlds depend on it */
b2g_status b2g_batch_status(b2g_batch* batch, int32_t* out);
void* b2g_batch_stream(b2g_batch* batch); /* cudaStream_t the batch launches on */
b2g_status b2g_batch_sync(b2g_batch* batch);

/* setWorldState (Box2DWorld.cpp:3492-3513 -> Box2DObject::setState :1980-1989): per object
 * setTransform(pose) [pose = first three DOFs; static objects keep their pose], setDOFPositions(all),
 * setDOFVelocities(all).  reset_caches != 0 first turns every world into what clone() (:3026-3038) gives a planner: a
 * freshly loaded world -- no contacts, zero warm-start impulses, inv_dt0 = 0, sleep timers, link enable flags and controller
 * targets as loaded -- that keeps the poses of its static objects (b2g_batch_set_object_pose), which belong to the world
 * state a clone receives.  q, qd: [n_worlds][nq]. */
b2g_status b2g_batch_set_state(b2g_batch* batch, const float* q, const float* qd, int reset_caches);
b2g_status b2g_batch_set_state_dev(b2g_batch* batch, const float* q, const float* qd, int reset_caches);
/* getWorldState (:3462-3474): dof_positions / dof_velocities of every object */
b2g_status b2g_batch_get_state(b2g_batch* batch, float* q, float* qd);
b2g_status b2g_batch_get_state_dev(b2g_batch* batch, float* q, float* qd);
/* Box2DObject::setPose(x, y, theta) (:1940-1947) for one object in every world; pose: [n_worlds][3] */
/* Box2DObject::setDOFPositions / setDOFVelocities(values, indices) (Box2DWorld.cpp:1585-1626, 1691-1715) and
 * getDOFPositions / getDOFVelocities(indices) (:1507-1528, 1628-1649) of one object in every selected world.
 * indices: n_indices object-local DOF indices in ascending order (base x, y, theta = 0, 1, 2 unless the object is
 * static; joints in YAML order after); null / 0 = all DOFs.  values / out: [n_worlds][n_indices].  velocities != 0
 * selects the velocity variant (values are clamped to the DOF velocity limits like the reference). */
b2g_status b2g_batch_set_object_dofs(b2g_batch* batch, int object, int velocities, const int32_t* indices, int n_indices,
                                     const float* values);
b2g_status b2g_batch_get_object_dofs(b2g_batch* batch, int object, int velocities, const int32_t* indices, int n_indices,
                                     float* out);
b2g_status b2g_batch_set_object_pose(b2g_batch* batch, int object, const float* pose);
b2g_status b2g_batch_get_object_pose(b2g_batch* batch, int object, float* pose);

/* Box2DWorld::setToRest (:3550-3558): setDOFVelocities(0) on every object of every world */
b2g_status b2g_batch_set_to_rest(b2g_batch* batch);
/* Box2DWorld::atRest(threshold) (:3435-3452): out[w] = 1 iff every DOF velocity of world w is within the threshold */
b2g_status b2g_batch_at_rest(b2g_batch* batch, float threshold, int32_t* out);
/* Box2DWorld::getObjects(aabb, objects, exclude_robots) (:3207-3254, Box2DAABBQuery :2979-2998): which objects have a
 * fixture overlapping the box aabb = (min_x, min_y, max_x, max_y) in user units — fat-AABB query, then the exact
 * b2TestOverlap (GJK) against the box.  One box for all worlds; hits [n_worlds][n_objects] (model object order). */
b2g_status b2g_batch_objects_in_aabb(b2g_batch* batch, const float aabb[4], int exclude_robots, uint8_t* hits);
/* Box2DWorld::isPhysicallyFeasible (:3370-3400): out[w] = 0 iff two touching links of world w overlap by more than
 * 1e-4 m^2 (convex-polygon clipping in place of boost::geometry::intersection). Runs checkCollision first. */
b2g_status b2g_batch_is_physically_feasible(b2g_batch* batch, int32_t* out);
/* Box2DLink::setEnabled (:802-822) for the link whose body index is `body` (links in creation order, body 0 = ground):
 * a disabled link keeps its body but collides with nothing (categoryBits = 0).  enabled: [n_worlds] or NULL to apply
 * enabled_all to every world. */
b2g_status b2g_batch_set_link_enabled(b2g_batch* batch, int body, const uint8_t* enabled, int enabled_all);

/* Box2DLink's dynamics parameters (src/sim_env/Box2DWorld.cpp:471-566).  A setter acts on link `body` (body index, 0 = ground
 * is not a link) of every SELECTED world, like calling the reference's setter on each of them:
 *   - the whole batch (the default selection): the one model all worlds share is patched;
 *   - a selection that covers whole tiles of 32 worlds (b2g_batch_select_worlds(batch, first, count) with first % 32 == 0 and
 *     count % 32 == 0 or first + count == n_worlds): those worlds get their own PARAMETER SET -- per-world dynamics for domain
 *     randomisation at tile granularity; the stepping kernels then stage the model variant of each warp's tile (a second
 *     build of the kernels, 1-3 % slower than the single-model path).  Any other selection is B2G_ERR_INVALID.
 * b2g_batch_get_link_property answers for the first selected world.
 *   MASS                    setMass :478-489 (b2Body mass and inertia scaled by mass / old mass, sweep centre re-derived,
 *                           ground-friction joint updated); no effect on the mass data of a static link, as upstream
 *   GROUND_FRICTION         setGroundFrictionCoefficient :496-500; both this and the next run updateFrictionJoint
 *   GROUND_TORQUE_INTEGRAL  setGroundFrictionTorqueIntegral :512-516 (:523-531: maxForce = mu m g scale, maxTorque = mu integral scale^2)
 *   CONTACT_FRICTION        setContactFriction :538-547 (b2Fixture::SetFriction on every fixture of the link).  As upstream,
 *                           contacts that exist at the time of the call keep the coefficient they were created with
 *                           (b2Contact mixes friction and restitution in its constructor); later contacts use the new one
 * Read-only: INERTIA getInertia :549-555 (about the link origin), COM_INERTIA getCOMInertia :557-566,
 * GROUND_FRICTION_LIMIT_FORCE :507-510 (mu m 9.81, unscaled), GROUND_FRICTION_LIMIT_TORQUE :502-505 (mu integral).
 * Like the reference's clone (it reloads the environment description, :3030), b2g_batch_clone does not inherit changes. */
typedef enum b2g_link_property {
    B2G_LINK_MASS = 0,
    B2G_LINK_GROUND_FRICTION = 1,
    B2G_LINK_GROUND_TORQUE_INTEGRAL = 2,
    B2G_LINK_CONTACT_FRICTION = 3,
    B2G_LINK_INERTIA = 4,
    B2G_LINK_COM_INERTIA = 5,
    B2G_LINK_GROUND_FRICTION_LIMIT_FORCE = 6,
    B2G_LINK_GROUND_FRICTION_LIMIT_TORQUE = 7
} b2g_link_property;
b2g_status b2g_batch_set_link_property(b2g_batch* batch, int body, int property, float value);
b2g_status b2g_batch_get_link_property(const b2g_batch* batch, int body, int property, float* out);
/* The same on a model (no per-world state involved): batches created from it afterwards start with the changed parameters,
 * as if the setter had been called on every freshly loaded world. */
b2g_status b2g_model_set_link_property(b2g_model* model, int body, int property, float value);
b2g_status b2g_model_get_link_property(const b2g_model* model, int body, int property, float* out);

/* Box2DRobotVelocityController::setTargetVelocity (src/sim_env/Box2DController.cpp:42-63) for robot
 * `object` in every world; v: [n_worlds][n_dofs of that robot].  The controller is then evaluated
 * in-kernel before every step (Box2DRobot::control, Box2DWorld.cpp:2284-2310). */
b2g_status b2g_batch_set_target_velocity(b2g_batch* batch, int object, const float* v);
b2g_status b2g_batch_set_target_velocity_dev(b2g_batch* batch, int object, const float* v);
/* Position control: sim_env::RobotPositionController / SE2RobotPositionController on top of the velocity controller
 * (instantiated at Box2DWorldViewer.cpp:1208-1216).  The classes belong to the sim_env package, which is not in the tree, so the
 * law is an ASSUMPTION stated here: at every control step v_target = kp * (q_target - q) over all DOFs of the robot (the heading
 * error of a moving base wrapped to (-pi, pi]), passed through setTargetVelocity (scaleToLimits included) to the velocity
 * controller.  q_target: [n_worlds][n_dofs of the robot]; kp in 1/s.  A later b2g_batch_set_target_velocity or
 * b2g_batch_set_control_efforts switches the robot back. */
b2g_status b2g_batch_set_target_position(b2g_batch* batch, int object, const float* q_target, float kp);
b2g_status b2g_batch_set_target_position_dev(b2g_batch* batch, int object, const float* q_target, float kp);
/* Box2DRobot::setController (Box2DWorld.cpp:2312-2317) with a ControlCallback that returns these efforts on every
 * control(dt) call (:2284-2310): each step they go to Box2DRobot::commandEfforts (:2319-2379) -- force on the base
 * centre (x scale), torque (x scale^2), joint motors through Box2DJoint::setControlTorque (:1198-1222).  This is the
 * hook for user control laws: compute efforts outside, install them, step.  efforts: [n_worlds][n_dofs of the robot].
 * Replaces the robot's velocity-controller target until b2g_batch_set_target_velocity is called again. */
b2g_status b2g_batch_set_control_efforts(b2g_batch* batch, int object, const float* efforts);
b2g_status b2g_batch_set_control_efforts_dev(b2g_batch* batch, int object, const float* efforts);

/* Box2DWorld::stepPhysics(steps, execute_controllers, allow_sleeping) (:3266-3283). */
b2g_status b2g_batch_step(b2g_batch* batch, int steps, int execute_controllers, int allow_sleeping);
/* Box2DWorld::stepPhysics(std::vector<Contact>& contacts, steps, execute_controllers, allow_sleeping) (:3290-3310):
 * steps like b2g_batch_step and returns, per world, one Contact for every b2ContactListener::BeginContact that fired
 * (Box2DCollisionChecker::BeginContact :2792-2818), preceded by the contacts that are touching at the start, found the
 * way startRecordings does (:2743-2768: saveState, setToRest, one step without controllers, restoreState).
 * counts [n_worlds] (may exceed max_per_world: the record is truncated, the count is not); iout [n_worlds][max][4] =
 * bodyA bodyB fixtureA fixtureB; fout [n_worlds][max][6] = contact_point(3) contact_normal(3), times 1/scale. */
b2g_status b2g_batch_step_record(b2g_batch* batch, int steps, int execute_controllers, int allow_sleeping, int32_t* counts,
                                 int32_t* iout, float* fout, int max_per_world);
/* Box2DWorld::stepPhysics(callback, steps, execute_controllers, allow_sleeping) (:3312-3333).  The reference calls
 * callback(link_a, link_b) on every BeginContact and stops stepping that world after the current step once it returns
 * false; here the callback is the table allowed_pairs[n_bodies][n_bodies] (0 = this pair aborts; bodies = links in
 * creation order, body 0 = ground).  completed [n_worlds]: 1 if no forbidden contact began (the reference's return
 * value); steps_done [n_worlds] (optional): steps actually executed. */
b2g_status b2g_batch_step_guarded(b2g_batch* batch, int steps, int execute_controllers, int allow_sleeping,
                                  const uint8_t* allowed_pairs, int32_t* completed, int32_t* steps_done);
/* "propagate" (planner-side composition): for t in [0, n_targets): setTargetVelocity(
 * targets[t]) on robot `object`, stepPhysics(steps_per_target).  targets: [n_targets][n_worlds][n_dofs]
 * on the device. One kernel launch. */
b2g_status b2g_batch_rollout_dev(b2g_batch* batch, int object, const float* targets, int n_targets, int steps_per_target);
/* The same propagate as ONE host-pointer call, the way a planner evaluates a batch of candidate controls: (optionally)
 * setWorldState(q, qd) on every world -- fresh clones when reset_caches != 0; q == qd == NULL keeps the current state --,
 * the n_targets x steps_per_target rollout, getWorldState into q_out / qd_out.  All buffers are host memory
 * ([n_worlds][nq], targets [n_targets][n_worlds][n_dofs]); copies run on the batch's stream (asynchronously when the
 * buffers are pinned, see b2g_host_alloc) and the call returns after one synchronisation. */
b2g_status b2g_batch_rollout(b2g_batch* batch, int object, const float* q, const float* qd, int reset_caches, const float* targets,
                             int n_targets, int steps_per_target, float* q_out, float* qd_out);

/* Box2DObject::getBallApproximation (Box2DWorld.cpp:1754-1763, Box2DLink::updateBallApproximation :733-747) of one
 * object in every world: out[n_worlds][n_balls][4] = centre (world frame, user units) and radius, links in name order */
b2g_status b2g_batch_get_ball_approximation(b2g_batch* batch, int object, float* out);

/* Box2DLink::getPose (Box2DWorld.cpp:215-224), getVelocityVector (:233-241) and getCenterOfMass (:595-603) of every link:
 * out[n_worlds][n_bodies][8] = pose x y theta, velocity x y omega, centre of mass x y (user units; row 0, the ground, is 0) */
b2g_status b2g_batch_get_link_states(b2g_batch* batch, float* out);

/* Parity probes (b2Body / b2Joint / b2Contact state).  bodies: [n_worlds][n_bodies][16] =
 * xf.p(2) xf.q.s xf.q.c c(2) a c0(2) a0 v(2) w sleepTime awake type */
b2g_status b2g_batch_get_bodies(b2g_batch* batch, float* out);
/* fat AABBs: [n_worlds][n_fixtures][4] */
b2g_status b2g_batch_get_fat_aabbs(b2g_batch* batch, float* out);
/* joints: iout [n_worlds][n_joints][6], fout [n_worlds][n_joints][15] (layout of b2g_model_get_joints with the
 * impulses, limit state, motor settings filled in) */
b2g_status b2g_batch_get_joints(b2g_batch* batch, int32_t* iout, float* fout);
/* contacts in creation order (b2World contact list reversed): counts [n_worlds]; iout [n_worlds][max][8] =
 * fixtureA fixtureB touching enabled pointCount type id0 id1; fout [n_worlds][max][12] = localNormal(2)
 * localPoint(2) p0(2) n0 t0 p1(2) n1 t1 */
b2g_status b2g_batch_get_contacts(b2g_batch* batch, int32_t* counts, int32_t* iout, float* fout, int max_per_world);

/* Box2DCollisionChecker::updateContactCache (:2880-2911): b2World::UpdateContacts() then every contact with
 * IsTouching && IsEnabled, in b2World contact-list order.  counts [n_worlds]; iout [n_worlds][max][4] = bodyA
 * bodyB fixtureA fixtureB; fout [n_worlds][max][6] = contact_point(3) contact_normal(3), both multiplied by
 * 1/scale exactly as the reference does (:2857-2862). */
b2g_status b2g_batch_check_collision(b2g_batch* batch, int32_t* counts, int32_t* iout, float* fout, int max_per_world);
/* same with device pointers, asynchronous on the batch's stream; iout / fout may be NULL to get the counts only */
b2g_status b2g_batch_check_collision_dev(b2g_batch* batch, int32_t* counts, int32_t* iout, float* fout, int max_per_world);

/* saveState / restoreState / dropState (:3515-3541): a stack of world states (q, qd and static poses) */
b2g_status b2g_batch_save_state(b2g_batch* batch);
b2g_status b2g_batch_restore_state(b2g_batch* batch);
b2g_status b2g_batch_drop_state(b2g_batch* batch);

/* pinned host memory for the end-to-end path */
b2g_status b2g_host_alloc(void** out, uint64_t bytes);
void b2g_host_free(void* p);

/* counters: kernels launched by this library since load (bench evidence) */
uint64_t b2g_launch_count(void);
/* device-side timing of the launches made since the last call (CUDA events on the batch stream):
 * returns the accumulated milliseconds of step/rollout kernels and resets the accumulator.  At most the 64 most recent
 * stepping launches are remembered (callers that never collect timings do not accumulate events). */
b2g_status b2g_batch_take_kernel_ms(b2g_batch* batch, float* out_ms, int32_t* out_launches);

/* ---- several GPUs of one node behind the same ABI ------------------------------------------------------
 * The reference's planner steps one Box2DWorld per call (include/sim_env/Box2DWorld.h:748-844); worlds are independent, so a
 * batch shards trivially: device i of `devices` owns the contiguous slice [first_i, first_i + count_i) of the n_worlds_total
 * worlds (slices differ by at most one world), with its own b2g_batch, stream and host thread inside THIS process.  The
 * only exchange is the final gather of the rollout states: ncclAllGather over NVLink on the batches' streams
 * (ncclCommInitAll, one communicator per device; libnccl.so.2 is loaded on first use).  No data-path collective otherwise. */
typedef struct b2g_multi b2g_multi;
b2g_status b2g_multi_create(const b2g_model* model, int64_t n_worlds_total, const int32_t* devices, int n_devices, int max_contacts,
                            b2g_multi** out);
void b2g_multi_destroy(b2g_multi* multi);
int32_t b2g_multi_device_count(const b2g_multi* multi);
/* shard i: its batch (every b2g_batch_* call works on it; worlds are numbered from 0 inside the shard), the global index of
 * its first world and its world count */
b2g_status b2g_multi_get_shard(b2g_multi* multi, int i, b2g_batch** batch, int64_t* first_world, int64_t* n_worlds);
/* b2g_batch_rollout over all devices at once.  Host buffers cover the whole job: q, qd, q_out, qd_out [n_worlds_total][nq],
 * targets [n_targets][n_worlds_total][n_dofs].  Every device thread copies its slice in, runs the rollout kernel and
 * getWorldState, then all devices all-gather the states (every device ends up with all n_worlds_total rows, see
 * b2g_multi_gathered_dev) and device 0 copies them to q_out / qd_out. */
b2g_status b2g_multi_rollout(b2g_multi* multi, int object, const float* q, const float* qd, int reset_caches, const float* targets,
                             int n_targets, int steps_per_target, float* q_out, float* qd_out);
/* the gathered states as device i holds them after b2g_multi_rollout: q_all / qd_all [n_devices][rows_per_device][nq]
 * (device j's slice starts at row j * rows_per_device; its last rows_per_device - count_j rows are padding) */
b2g_status b2g_multi_gathered_dev(b2g_multi* multi, int i, const float** q_all, const float** qd_all, int64_t* rows_per_device);
/* timings of the last b2g_multi_rollout, CUDA events on the batches' streams: the rollout kernel per device (min / max over
 * the devices) and the all-gather (max over the devices; it includes waiting for the slowest device's kernel) */
b2g_status b2g_multi_last_timing(b2g_multi* multi, float* kernel_ms_min, float* kernel_ms_max, float* gather_ms);

#ifdef __cplusplus
}
#endif
#endif /* B2G_H_ */
