// b2g_world.hpp — header-only C++ facade over the C ABI (include/b2g.h, libb2g.so).
//
// BatchBox2DWorld keeps the method names and argument meaning of sim_env::Box2DWorld
// (/root/reference/include/sim_env/Box2DWorld.h:748-844); every array gains a leading world index.  The C ABI's
// status codes are turned back into the exceptions the reference throws (std::runtime_error for bad sizes / missing
// files, std::logic_error for impossible states; e.g. src/sim_env/Box2DWorld.cpp:1700-1704).
//
//   sim_env::b2g::BatchBox2DWorld batch(4096, /*cuda device*/ 0);
//   batch.loadWorld("test_data/test_env.yaml");
//   batch.setWorldState(q, qd, /*reset_caches=*/true);
//   batch.setTargetVelocity("my_robot", u);
//   batch.stepPhysics(10);
//   batch.getWorldState(q, qd);
#ifndef B2G_WORLD_HPP_
#define B2G_WORLD_HPP_

#include <cstdint>
#include <cstdio>
#include <limits>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "b2g.h"

namespace sim_env {
namespace b2g {

// sim_env::Contact restated for batched use: links / objects are given by body index (links in creation order, body 0 is
// the ground) instead of weak pointers (Box2DWorld.cpp:2799-2807)
struct Contact {
    float contact_point[3];
    float contact_normal[3];
    int body_a, body_b;        // Box2DLink of fixture A / B
    int fixture_a, fixture_b;  // broad-phase proxy ids
};

struct ObjectInfo {
    std::string name;
    bool is_robot = false, is_static = false;
    int n_dofs = 0, dof_offset = 0, first_body = 0, n_links = 0, n_joints = 0, n_balls = 0;
};

// sim_env types the single-world views speak (restated; Eigen is replaced by std::vector / plain arrays):
// DOFInformation (Box2DWorld.cpp:1338-1361), ObjectState (:1962-1970), WorldState (:3462-3473), Ball (:709-720)
struct DOFInformation {
    int dof_index = 0;
    float position_limits[2], velocity_limits[2], acceleration_limits[2];
    bool cyclic = false;
};
struct ObjectState {
    std::vector<int> active_dofs;
    std::vector<float> dof_positions, dof_velocities;  // all DOFs of the object (getDOFIndices order)
    float pose[3] = {0.0f, 0.0f, 0.0f};                // base link pose x, y, theta (the reference stores an Affine3f)
};
typedef std::map<std::string, ObjectState> WorldState;
struct Ball {
    float center[3];
    float radius;
};
enum class EntityType { Link, Joint, Object, Robot };

class BatchBox2DWorld;
class World;
class Object;
class Joint;

// sim_env::Logger as the reference uses it (Box2DWorld::getLogger, Box2DWorld.cpp:3011 / 3565-3568: DefaultLogger with a level
// and a prefix per message).  Loader warnings (what the reference logs through logWarn while parsing, Box2DIOUtils.cpp:70-84) and
// the facade's own notes go through it; the default sink is stderr, setSink redirects (or silences) it.
class Logger {
public:
    enum class LogLevel { Debug = 0, Info = 1, Warn = 2, Error = 3 };
    typedef void (*Sink)(LogLevel level, const std::string& message, const std::string& prefix);
    void setLevel(LogLevel lvl) { _level = lvl; }
    LogLevel getLevel() const { return _level; }
    void setSink(Sink sink) { _sink = sink; }
    void logDebug(const std::string& msg, const std::string& prefix = "") const { log(LogLevel::Debug, msg, prefix); }
    void logInfo(const std::string& msg, const std::string& prefix = "") const { log(LogLevel::Info, msg, prefix); }
    void logWarn(const std::string& msg, const std::string& prefix = "") const { log(LogLevel::Warn, msg, prefix); }
    void logErr(const std::string& msg, const std::string& prefix = "") const { log(LogLevel::Error, msg, prefix); }
    void log(LogLevel lvl, const std::string& msg, const std::string& prefix) const {
        if ((int)lvl < (int)_level) return;
        if (_sink) {
            _sink(lvl, msg, prefix);
            return;
        }
        static const char* names[] = {"[Debug]", "[Info]", "[Warning]", "[Error]"};
        std::fprintf(stderr, "%s %s %s\n", names[(int)lvl], prefix.c_str(), msg.c_str());
    }

private:
    LogLevel _level = LogLevel::Info;
    Sink _sink = nullptr;
};

// One link of ONE world with the method names of Box2DLink (/root/reference/include/sim_env/Box2DWorld.h:100-220).  Obtained
// from Object::getLink / getBaseLink / getLinks.  A link is identified by its body index (what contacts report).
class Link {
public:
    std::string getName() const;
    EntityType getType() const { return EntityType::Link; }
    int body() const { return _body; }
    int objectIndex() const;
    // Box2DLink::getPose (:215-224), getVelocityVector (:233-241), getCenterOfMass (:595-603): x, y, theta / vx, vy, omega / x, y
    void getPose(float pose[3]) const;
    void getVelocityVector(float vel[3]) const;
    void getCenterOfMass(float com[2]) const;
    // link parameters (:463-603); the setters change the model the worlds of the batch share
    float getMass() const;
    float getInertia() const;
    float getGroundFrictionCoefficient() const;
    float getGroundFrictionTorqueIntegral() const;
    float getGroundFrictionLimitForce() const;
    float getGroundFrictionLimitTorque() const;
    float getContactFriction() const;
    // kinematic tree (:243-256): the joint this link is the child of (false for a base link), the joints it is the parent of
    bool getParentJoint(Joint& joint) const;
    std::vector<Joint> getChildJoints() const;
    // Box2DLink::checkCollision overloads (Box2DWorld.h:124-129) on this world's contact list
    bool checkCollision();
    bool checkCollision(std::vector<Contact>& contacts);
    bool checkCollision(const std::vector<Link>& other_links, std::vector<Contact>* contacts = nullptr);

private:
    friend class Object;
    friend class Joint;
    Link(BatchBox2DWorld* batch, int64_t world, int body) : _batch(batch), _world(world), _body(body) {}
    void row(float out[8]) const;
    BatchBox2DWorld* _batch;
    int64_t _world;
    int _body;
};

// One joint of ONE world with the method names of Box2DJoint (Box2DWorld.h:226-345).  Obtained from Object::getJoint / getJoints.
class Joint {
public:
    enum class JointType { Revolute, Prismatic };
    Joint() : _batch(nullptr), _world(0), _joint(-1) {}
    std::string getName() const;
    EntityType getType() const { return EntityType::Joint; }
    JointType getJointType() const;
    int index() const { return _joint; }          // model joint index
    unsigned int getJointIndex() const;           // position among the object's joints (Box2DJoint::getJointIndex)
    unsigned int getDOFIndex() const;             // object-local DOF index (Box2DJoint::getDOFIndex)
    // getPosition / getVelocity (:965-981, :1030-1046) and their setters (:983-990, :1048-1055): one DOF of the owning object
    float getPosition() const;
    float getVelocity() const;
    void setPosition(float v);
    void setVelocity(float v);
    void getPositionLimits(float limits[2]) const;
    void getVelocityLimits(float limits[2]) const;
    void getAccelerationLimits(float limits[2]) const;
    DOFInformation getDOFInformation() const;
    Link getChildLink() const;
    Link getParentLink() const;

private:
    friend class Object;
    friend class Link;
    Joint(BatchBox2DWorld* batch, int64_t world, int joint) : _batch(batch), _world(world), _joint(joint) {}
    BatchBox2DWorld* _batch;
    int64_t _world;
    int _joint;
};

// One object (or robot) of ONE world of the batch with the method names of Box2DObject / Box2DRobot
// (/root/reference/include/sim_env/Box2DWorld.h:350-590).  Obtained from World::getObject / getRobot.
class Object {
public:
    const std::string& getName() const;
    EntityType getType() const;
    bool isStatic() const;
    int index() const { return _object; }
    // DOF bookkeeping (:1435-1505): active DOFs are kept per object name in the batch (one set for all worlds)
    unsigned int getNumDOFs() const;
    unsigned int getNumBaseDOFs() const { return isStatic() ? 0u : 3u; }
    std::vector<int> getDOFIndices() const;
    std::vector<int> getActiveDOFs() const;
    unsigned int getNumActiveDOFs() const { return (unsigned int)getActiveDOFs().size(); }
    void setActiveDOFs(const std::vector<int>& indices);
    DOFInformation getDOFInformation(unsigned int dof_index) const;
    // getDOFPositions / getDOFVelocities / setDOFPositions / setDOFVelocities (:1507-1528, 1628-1649, 1585-1626, 1691-1715):
    // an empty index list means the active DOFs; indices must be ascending
    std::vector<float> getDOFPositions(const std::vector<int>& indices = std::vector<int>()) const;
    std::vector<float> getDOFVelocities(const std::vector<int>& indices = std::vector<int>()) const;
    void setDOFPositions(const std::vector<float>& values, const std::vector<int>& indices = std::vector<int>());
    void setDOFVelocities(const std::vector<float>& values, const std::vector<int>& indices = std::vector<int>());
    // base link pose (Box2DLink::getPose :215-224, Box2DObject::setPose :1940-1947)
    void getPose(float pose[3]) const;
    void setPose(const float pose[3]);
    // getState / setState (:1962-1989)
    void getState(ObjectState& state) const;
    ObjectState getState() const {
        ObjectState s;
        getState(s);
        return s;
    }
    void setState(const ObjectState& state);
    void setToRest() { setDOFVelocities(std::vector<float>(getNumDOFs(), 0.0f), getDOFIndices()); }  // (:1717-1722)
    // checkCollision overloads (Box2DWorld.h:372-377)
    bool checkCollision();
    bool checkCollision(std::vector<Contact>& contacts);
    bool checkCollision(const Object& other);
    bool checkCollision(const Object& other, std::vector<Contact>& contacts);
    bool checkCollision(const std::vector<Object>& others);
    bool checkCollision(const std::vector<Object>& others, std::vector<Contact>& contacts);
    // getBallApproximation (:1754-1763)
    void getBallApproximation(std::vector<Ball>& balls) const;
    // Box2DRobotVelocityController::setTargetVelocity (Box2DController.cpp:42-63); one value per DOF of the robot
    void setTargetVelocity(const std::vector<float>& target);
    // Box2DObject::getLink / getBaseLink / getLinks / getJoint / getJoints (:1411-1505)
    Link getLink(const std::string& link_name) const;
    Link getBaseLink() const;
    std::vector<Link> getLinks() const;
    Joint getJoint(const std::string& joint_name) const;
    std::vector<Joint> getJoints() const;

private:
    friend class World;
    friend class Link;
    friend class Joint;
    Object(BatchBox2DWorld* batch, int64_t world, int object) : _batch(batch), _world(world), _object(object) {}
    BatchBox2DWorld* _batch;
    int64_t _world;
    int _object;
};

// ONE world of the batch with the single-world signatures of sim_env::Box2DWorld (Box2DWorld.h:748-844): what a planner
// written against sim_env::World calls.  Every call touches only this world (b2g_batch_select_worlds); stepping is a
// batch operation and is only forwarded when the batch holds a single world.
class World {
public:
    int64_t index() const { return _world; }
    Logger& getLogger() const;   // Box2DWorld::getLogger (:3565-3568): the batch's logger
    Object getRobot(const std::string& name) const;
    Object getObject(const std::string& name, bool exclude_robots = true) const;
    void getObjects(std::vector<Object>& objects, bool exclude_robots = true) const;
    void getRobots(std::vector<Object>& robots) const;
    // stepPhysics overloads (:3256-3333): only for a batch of one world (use BatchBox2DWorld::stepPhysics otherwise)
    void stepPhysics(int steps = 1, bool execute_controllers = true, bool allow_sleeping = true);
    bool stepPhysics(std::vector<Contact>& contacts, int steps = 1, bool execute_controllers = true, bool allow_sleeping = true);
    float getPhysicsTimeStep() const;
    int getVelocitySteps() const;
    int getPositionSteps() const;
    // getWorldState / setWorldState (:3462-3513); setWorldState returns false for unknown object names like the reference
    void getWorldState(WorldState& state) const;
    WorldState getWorldState() const {
        WorldState s;
        getWorldState(s);
        return s;
    }
    bool setWorldState(const WorldState& state);
    // checkCollision overloads (Box2DWorld.h:797-815); links are given by body index (Contact::body_a / body_b)
    bool checkCollision(std::vector<Contact>& contacts);
    bool checkCollision(const Object& object_a, const Object& object_b);
    bool checkCollision(const Object& object_a, const Object& object_b, std::vector<Contact>& contacts);
    bool checkCollision(const Object& object);
    bool checkCollision(const Object& object, std::vector<Contact>& contacts);
    bool checkCollision(const Object& object, std::vector<Object>& collidors);
    bool checkCollision(const Object& object_a, const std::vector<Object>& other_objects);
    bool checkCollision(const Object& object_a, const std::vector<Object>& other_objects, std::vector<Contact>& contacts);
    bool checkCollisionLink(int link_body);
    bool checkCollisionLink(int link_body, std::vector<Contact>& contacts);
    bool checkCollisionLink(int link_body, const std::vector<int>& other_link_bodies, std::vector<Contact>* contacts = nullptr);
    bool checkCollisionLink(int link_body, const std::vector<Object>& other_objects, std::vector<Contact>* contacts = nullptr);
    bool checkCollision(const Object& object_a, int link_body_b, std::vector<Contact>* contacts = nullptr);
    // atRest (:3435-3452), isPhysicallyFeasible (:3370-3400), getObjects(aabb, ...) (:3207-3254), setToRest (:3550-3558)
    bool atRest(float threshold = 0.0001f) const;
    bool isPhysicallyFeasible() const;
    void getObjects(const float aabb[4], std::vector<Object>& objects, bool exclude_robots = true) const;

private:
    friend class BatchBox2DWorld;
    friend class Object;
    friend class Link;
    World(BatchBox2DWorld* batch, int64_t world) : _batch(batch), _world(world) {}
    // Box2DCollisionChecker (:2551-2741) on top of the world's touching-contact list: a contact is reported for a body set A
    // when one of its bodies is in A; with `others` the other body must be in it too.  link_a of a reported contact is the
    // body that belongs to A (the checker fills link_a from the queried collidable, :2596-2609).
    bool filterContacts(const std::vector<char>& in_a, const std::vector<char>* in_b, std::vector<Contact>* out,
                        std::vector<char>* hit_bodies = nullptr);
    std::vector<char> bodiesOf(const Object& o) const;
    BatchBox2DWorld* _batch;
    int64_t _world;
};

class BatchBox2DWorld {
public:
    explicit BatchBox2DWorld(int64_t n_worlds = 1, int device = 0, int max_contacts = 0)
        : _n(n_worlds), _device(device), _max_contacts(max_contacts) {}
    BatchBox2DWorld(const BatchBox2DWorld&) = delete;
    BatchBox2DWorld& operator=(const BatchBox2DWorld&) = delete;
    ~BatchBox2DWorld() { destroy(); }

    // Box2DWorld::loadWorld(path) (Box2DWorld.cpp:3040-3054)
    void loadWorld(const std::string& path) {
        destroy();
        _path = path;
        check(b2g_env_load_yaml(path.c_str(), &_env));
        finishLoad();
    }

    // Box2DWorld::clone (:3021-3038): same environment and step parameters, setWorldState(getWorldState()) incl. the poses of
    // static objects, fresh caches.  The clone re-reads the environment file the original was loaded from.
    BatchBox2DWorld* clone() {
        if (!_batch) throw std::logic_error("[sim_env::b2g::BatchBox2DWorld::clone] no world loaded");
        if (_path.empty()) throw std::logic_error("[sim_env::b2g::BatchBox2DWorld::clone] the world was not loaded from a file");
        BatchBox2DWorld* other = new BatchBox2DWorld(_n, _device, _max_contacts);
        try {
            check(b2g_env_load_yaml(_path.c_str(), &other->_env));
            check(b2g_model_create(other->_env, &other->_model));
            other->_info = _info;
            other->_objects = _objects;
            other->_active = _active;
            other->_limits = _limits;
            other->_path = _path;
            other->_dt = _dt; other->_vs = _vs; other->_ps = _ps;
            check(b2g_batch_clone(_batch, &other->_batch));
        } catch (...) {
            delete other;
            throw;
        }
        return other;
    }

    // Box2DWorld::loadWorld(const Box2DEnvironmentDescription&) (:3049-3054): a description built with b2g_env_create /
    // b2g_env_add_* (the batch takes ownership of `env`).  clone() is not available for such a world (no file to re-read).
    void loadWorld(b2g_env* env) {
        destroy();
        _path.clear();
        _env = env;
        finishLoad();
    }
    // loader diagnostics: what the reference sends to its logger while loading (getLogger()->logWarn), one per entry
    std::vector<std::string> getLoadWarnings() const {
        std::vector<std::string> out;
        if (!_model) return out;
        std::vector<char> buf(16384);
        int32_t n = 0;
        check(b2g_model_get_warnings(_model, buf.data(), (int)buf.size(), &n));
        std::string all(buf.data()), line;
        for (size_t i = 0; i < all.size(); ++i) {
            if (all[i] == '\n') { if (!line.empty()) out.push_back(line); line.clear(); }
            else line.push_back(all[i]);
        }
        return out;
    }
    Logger& getLogger() const { return _logger; }                      // Box2DWorld::getLogger (:3565-3568)
    bool supportsPhysics() const { return true; }                      // (:3335-3338)
    bool isRobot(const std::string& name) const {                      // (:3202-3205)
        for (size_t i = 0; i < _objects.size(); ++i)
            if (_objects[i].name == name) return _objects[i].is_robot;
        return false;
    }
    float getGravity() const { return 9.81f * _info.scale; }           // GRAVITY * getScale() (Box2DWorld.h:865-869)
    // printWorldState (:3475-3488) of one world, to a stdio stream instead of the sim_env logger
    void printWorldState(int64_t world_index, std::FILE* out = stdout);

    // model introspection (host only; valid after loadWorld's model stage even when batch creation failed)
    int64_t numWorlds() const { return _n; }
    int nq() const { return _info.nq; }
    int numBodies() const { return _info.n_bodies; }
    float getScale() const { return _info.scale; }
    float getInverseScale() const { return 1.0f / _info.scale; }
    const std::vector<ObjectInfo>& getObjects() const { return _objects; }
    int objectIndex(const std::string& name) const {
        for (size_t i = 0; i < _objects.size(); ++i)
            if (_objects[i].name == name) return (int)i;
        throw std::runtime_error("[sim_env::b2g::BatchBox2DWorld] unknown object '" + name + "'");
    }
    int robotDofs(const std::string& name) const { return _objects[objectIndex(name)].n_dofs; }

    // setPhysicsTimeStep / setVelocitySteps / setPositionSteps and getters (:3340-3368)
    void setPhysicsTimeStep(float dt) { _dt = dt; pushParams(); }
    void setVelocitySteps(int n) { _vs = n; pushParams(); }
    void setPositionSteps(int n) { _ps = n; pushParams(); }
    float getPhysicsTimeStep() const { return _dt; }
    int getVelocitySteps() const { return _vs; }
    int getPositionSteps() const { return _ps; }

    // stepPhysics(steps, execute_controllers, allow_sleeping) (:3256-3283)
    void stepPhysics(int steps = 1, bool execute_controllers = true, bool allow_sleeping = true) {
        check(b2g_batch_step(_batch, steps, execute_controllers, allow_sleeping));
    }
    // stepPhysics(std::vector<Contact>& contacts, steps, ...) (:3290-3310): contacts of world w are
    // contacts[offsets[w] .. offsets[w + 1])
    void stepPhysics(std::vector<Contact>& contacts, std::vector<int64_t>& offsets, int steps = 1, bool execute_controllers = true,
                     bool allow_sleeping = true, int max_per_world = 32) {
        std::vector<int32_t> counts(_n), io((size_t)_n * max_per_world * 4);
        std::vector<float> fo((size_t)_n * max_per_world * 6);
        check(b2g_batch_step_record(_batch, steps, execute_controllers, allow_sleeping, counts.data(), io.data(), fo.data(),
                                    max_per_world));
        pack(counts, io, fo, max_per_world, contacts, offsets);
    }
    // stepPhysics(callback, steps, ...) (:3312-3333) with the callback as an [n_bodies][n_bodies] table (0 = abort);
    // returns per world what the reference returns (true = no forbidden contact began)
    std::vector<bool> stepPhysics(const std::vector<uint8_t>& allowed_pairs, int steps, bool execute_controllers = true,
                                  bool allow_sleeping = true, std::vector<int32_t>* steps_done = nullptr) {
        if ((int64_t)allowed_pairs.size() != (int64_t)_info.n_bodies * _info.n_bodies)
            throw std::runtime_error("[sim_env::b2g::BatchBox2DWorld::stepPhysics] allowed_pairs must be n_bodies x n_bodies");
        std::vector<int32_t> completed(_n), done(_n);
        check(b2g_batch_step_guarded(_batch, steps, execute_controllers, allow_sleeping, allowed_pairs.data(), completed.data(),
                                     done.data()));
        if (steps_done) *steps_done = done;
        std::vector<bool> ok(_n);
        for (int64_t i = 0; i < _n; ++i) ok[i] = completed[i] != 0;
        return ok;
    }

    // planner "propagate" in one call: setWorldState(q, qd) (fresh clones), then for every targets[t] ([n_targets][n_worlds]
    // [n_dofs]) setTargetVelocity + stepPhysics(steps_per_target), then getWorldState into q_out / qd_out
    void propagate(const std::string& robot, const float* q, const float* qd, const float* targets, int n_targets,
                   int steps_per_target, float* q_out, float* qd_out, bool reset_caches = true) {
        check(b2g_batch_rollout(_batch, objectIndex(robot), q, qd, reset_caches, targets, n_targets, steps_per_target, q_out, qd_out));
    }

    // getWorldState / setWorldState (:3462-3513); q, qd: [n_worlds][nq]
    void getWorldState(float* q, float* qd) { check(b2g_batch_get_state(_batch, q, qd)); }
    void setWorldState(const float* q, const float* qd, bool reset_caches = false) {
        check(b2g_batch_set_state(_batch, q, qd, reset_caches));
    }
    void getWorldState(std::vector<float>& q, std::vector<float>& qd) {
        q.resize((size_t)_n * nq());
        qd.resize((size_t)_n * nq());
        getWorldState(q.data(), qd.data());
    }
    void setWorldState(const std::vector<float>& q, const std::vector<float>& qd, bool reset_caches = false) {
        if (q.size() != (size_t)_n * nq() || qd.size() != q.size())
            throw std::runtime_error("[sim_env::b2g::BatchBox2DWorld::setWorldState] state arrays must be [n_worlds][nq]");
        setWorldState(q.data(), qd.data(), reset_caches);
    }
    // saveState / restoreState / dropState (:3515-3541); restoreState returns false on an empty stack like the reference
    void saveState() { check(b2g_batch_save_state(_batch)); }
    bool restoreState() {
        b2g_status s = b2g_batch_restore_state(_batch);
        if (s == B2G_ERR_INVALID && std::string(b2g_last_error()).find("stack is empty") != std::string::npos) return false;
        check(s);
        return true;
    }
    void dropState() { check(b2g_batch_drop_state(_batch)); }
    // Box2DObject::setPose / getPose (:1940-1947, :215-224); pose: [n_worlds][3]
    void setObjectPose(const std::string& object, const float* pose) { check(b2g_batch_set_object_pose(_batch, objectIndex(object), pose)); }
    void getObjectPose(const std::string& object, float* pose) { check(b2g_batch_get_object_pose(_batch, objectIndex(object), pose)); }

    // Box2DRobotVelocityController::setTargetVelocity (Box2DController.cpp:42-63); v: [n_worlds][n_dofs of the robot]
    void setTargetVelocity(const std::string& robot, const float* v) {
        check(b2g_batch_set_target_velocity(_batch, objectIndex(robot), v));
    }

    // sim_env::RobotPositionController / SE2RobotPositionController on top of the velocity controller (instantiated by the
    // reference's viewer, Box2DWorldViewer.cpp:1208-1216; assumed law, include/b2g.h): q_target: [n_worlds][n_dofs of the robot]
    void setTargetPosition(const std::string& robot, const float* q_target, float kp = 5.0f) {
        check(b2g_batch_set_target_position(_batch, objectIndex(robot), q_target, kp));
    }

    // Box2DRobot::setController (:2312-2317) with a ControlCallback that returns these efforts on every step (they reach
    // Box2DRobot::commandEfforts :2319-2379): the hook for user control laws; efforts: [n_worlds][n_dofs of the robot]
    void setControlEfforts(const std::string& robot, const float* efforts) {
        check(b2g_batch_set_control_efforts(_batch, objectIndex(robot), efforts));
    }

    // checkCollision(std::vector<Contact>&) (:3846-3859 -> updateContactCache :2880-2911) for every world
    void checkCollision(std::vector<Contact>& contacts, std::vector<int64_t>& offsets, int max_per_world = 32) {
        std::vector<int32_t> counts(_n), io((size_t)_n * max_per_world * 4);
        std::vector<float> fo((size_t)_n * max_per_world * 6);
        check(b2g_batch_check_collision(_batch, counts.data(), io.data(), fo.data(), max_per_world));
        pack(counts, io, fo, max_per_world, contacts, offsets);
    }
    // checkCollision() -> bool per world
    std::vector<bool> checkCollision() {
        std::vector<Contact> c;
        std::vector<int64_t> off;
        checkCollision(c, off, 1);
        std::vector<bool> hit(_n);
        for (int64_t i = 0; i < _n; ++i) hit[i] = off[i + 1] > off[i];
        return hit;
    }

    // getObjects(aabb, objects, exclude_robots) (:3207-3254): hits[w * n_objects + o]
    std::vector<uint8_t> getObjects(const float aabb[4], bool exclude_robots = true) {
        std::vector<uint8_t> hits((size_t)_n * _objects.size());
        check(b2g_batch_objects_in_aabb(_batch, aabb, exclude_robots, hits.data()));
        return hits;
    }
    // isPhysicallyFeasible (:3370-3400), atRest (:3435-3452), setToRest (:3550-3558), Box2DLink::setEnabled (:802-822)
    std::vector<bool> isPhysicallyFeasible() {
        std::vector<int32_t> out(_n);
        check(b2g_batch_is_physically_feasible(_batch, out.data()));
        return std::vector<bool>(out.begin(), out.end());
    }
    std::vector<bool> atRest(float threshold = 0.0001f) {
        std::vector<int32_t> out(_n);
        check(b2g_batch_at_rest(_batch, threshold, out.data()));
        return std::vector<bool>(out.begin(), out.end());
    }
    void setToRest() { check(b2g_batch_set_to_rest(_batch)); }
    void setLinkEnabled(int body, bool enabled) { check(b2g_batch_set_link_enabled(_batch, body, nullptr, enabled)); }

    // Box2DObject::getLink(name) / getJoint(name) / getBaseLink (Box2DWorld.cpp:1411-1505): a link is its body index (what
    // Contact::body_a / body_b report), a joint its model joint index
    int getLink(const std::string& object, const std::string& link) const {
        int32_t body = -1;
        check(b2g_model_find_link(_model, objectIndex(object), link.c_str(), &body));
        return body;
    }
    int getJoint(const std::string& object, const std::string& joint) const {
        int32_t j = -1;
        check(b2g_model_find_joint(_model, objectIndex(object), joint.c_str(), &j));
        return j;
    }
    b2g_link_info getLinkInfo(int body) const {
        b2g_link_info info;
        check(b2g_model_get_link(_model, body, &info));
        return info;
    }
    b2g_joint_info getJointInfo(int joint) const {
        b2g_joint_info info;
        check(b2g_model_get_joint(_model, joint, &info));
        return info;
    }
    std::string getLinkName(int body) const { return getLinkInfo(body).name; }
    // Box2DObject::getLocalAABB (:1928-1931) / Box2DLink::getLocalBoundingBox (:258-261): min x, min y, max x, max y
    void getLocalAABB(const std::string& object, float aabb[4]) const {
        check(b2g_model_get_object_aabb(_model, objectIndex(object), aabb));
    }
    void getLocalBoundingBox(int body, float aabb[4]) const { check(b2g_model_get_link_aabb(_model, body, aabb)); }
    int getBaseLink(const std::string& object) const {
        const ObjectInfo& o = _objects[objectIndex(object)];
        for (int b = o.first_body; b < o.first_body + o.n_links; ++b)
            if (getLinkInfo(b).is_base_link) return b;
        throw std::logic_error("[sim_env::b2g::BatchBox2DWorld::getBaseLink] object without a base link");
    }
    std::vector<int> getLinks(const std::string& object) const {
        const ObjectInfo& o = _objects[objectIndex(object)];
        std::vector<int> links;
        for (int b = o.first_body; b < o.first_body + o.n_links; ++b) links.push_back(b);
        return links;
    }

    // Box2DLink::getPose (:215-224), getVelocityVector (:233-241), getCenterOfMass (:595-603) of every link of every world:
    // out[(w * numBodies() + body) * 8 + k], k = pose x y theta, velocity x y omega, centre of mass x y
    void getLinkStates(std::vector<float>& out) {
        out.resize((size_t)_n * _info.n_bodies * 8);
        check(b2g_batch_get_link_states(_batch, out.data()));
    }

    // Box2DLink's dynamics parameters (Box2DWorld.cpp:471-566) for the link with body index `body` in every selected world:
    // the whole batch, or -- per-world parameters at tile granularity -- a selectWorlds(first, count) range that covers whole
    // tiles of 32 worlds (include/b2g.h)
    void setMass(int body, float mass) { check(b2g_batch_set_link_property(_batch, body, B2G_LINK_MASS, mass)); }
    void setGroundFrictionCoefficient(int body, float mu) {
        check(b2g_batch_set_link_property(_batch, body, B2G_LINK_GROUND_FRICTION, mu));
    }
    void setGroundFrictionTorqueIntegral(int body, float val) {
        check(b2g_batch_set_link_property(_batch, body, B2G_LINK_GROUND_TORQUE_INTEGRAL, val));
    }
    void setContactFriction(int body, float mu) { check(b2g_batch_set_link_property(_batch, body, B2G_LINK_CONTACT_FRICTION, mu)); }
    // Per-world dynamics (domain randomisation): the same setters for worlds [first, first + count) only.  The range has to
    // cover whole tiles of 32 worlds (first % 32 == 0, count % 32 == 0 or first + count == numWorlds()): those worlds then
    // carry their own parameter set.  getLinkPropertyOfWorld answers for one world.
    void setLinkPropertyOfWorlds(int64_t first, int64_t count, int body, b2g_link_property property, float value) {
        check(b2g_batch_select_worlds(_batch, first, count));
        const b2g_status st = b2g_batch_set_link_property(_batch, body, property, value);
        b2g_batch_select_worlds(_batch, 0, 0);
        check(st);
    }
    float getLinkPropertyOfWorld(int64_t world, int body, b2g_link_property property) const {
        float v = 0.0f;
        check(b2g_batch_select_worlds(_batch, world, 1));
        const b2g_status st = b2g_batch_get_link_property(_batch, body, property, &v);
        b2g_batch_select_worlds(_batch, 0, 0);
        check(st);
        return v;
    }
    float getLinkProperty(int body, b2g_link_property property) const {
        float v = 0.0f;
        check(b2g_batch_get_link_property(_batch, body, property, &v));
        return v;
    }
    float getMass(int body) const { return getLinkProperty(body, B2G_LINK_MASS); }
    float getInertia(int body) const { return getLinkProperty(body, B2G_LINK_INERTIA); }
    float getGroundFrictionCoefficient(int body) const { return getLinkProperty(body, B2G_LINK_GROUND_FRICTION); }
    float getGroundFrictionTorqueIntegral(int body) const { return getLinkProperty(body, B2G_LINK_GROUND_TORQUE_INTEGRAL); }
    float getGroundFrictionLimitForce(int body) const { return getLinkProperty(body, B2G_LINK_GROUND_FRICTION_LIMIT_FORCE); }
    float getGroundFrictionLimitTorque(int body) const { return getLinkProperty(body, B2G_LINK_GROUND_FRICTION_LIMIT_TORQUE); }
    float getContactFriction(int body) const { return getLinkProperty(body, B2G_LINK_CONTACT_FRICTION); }

    // Box2DObject::getBallApproximation (:1754-1763) of one object in every world: out[w * n_balls + k]
    std::vector<Ball> getBallApproximation(const std::string& object) {
        const int o = objectIndex(object);
        std::vector<float> raw((size_t)_n * _objects[o].n_balls * 4);
        check(b2g_batch_get_ball_approximation(_batch, o, raw.data()));
        std::vector<Ball> balls(raw.size() / 4);
        for (size_t i = 0; i < balls.size(); ++i) {
            for (int d = 0; d < 3; ++d) balls[i].center[d] = raw[i * 4 + d];
            balls[i].radius = raw[i * 4 + 3];
        }
        return balls;
    }
    // per-object DOF access for every world: values / result [n_worlds][indices.size()], empty indices = all DOFs
    std::vector<float> getDOFPositions(const std::string& object, const std::vector<int>& indices = std::vector<int>(),
                                       bool velocities = false) {
        const int o = objectIndex(object);
        const size_t n = indices.empty() ? (size_t)_objects[o].n_dofs : indices.size();
        std::vector<float> out((size_t)_n * n);
        std::vector<int32_t> ix(indices.begin(), indices.end());
        check(b2g_batch_get_object_dofs(_batch, o, velocities, ix.data(), (int)ix.size(), out.data()));
        return out;
    }
    void setDOFPositions(const std::string& object, const std::vector<float>& values,
                         const std::vector<int>& indices = std::vector<int>(), bool velocities = false) {
        const int o = objectIndex(object);
        const size_t n = indices.empty() ? (size_t)_objects[o].n_dofs : indices.size();
        if (values.size() != (size_t)_n * n)
            throw std::runtime_error("[sim_env::b2g::BatchBox2DWorld::setDOFPositions] values must be [n_worlds][n_indices]");
        std::vector<int32_t> ix(indices.begin(), indices.end());
        check(b2g_batch_set_object_dofs(_batch, o, velocities, ix.data(), (int)ix.size(), values.data()));
    }

    // single-world view with the reference's one-world signatures
    World world(int64_t i) {
        if (i < 0 || i >= _n) throw std::runtime_error("[sim_env::b2g::BatchBox2DWorld::world] world index out of range");
        return World(this, i);
    }

    // per-world status after the last call (contact capacity exceeded -> std::runtime_error)
    std::vector<int32_t> status() {
        std::vector<int32_t> s(_n);
        check(b2g_batch_status(_batch, s.data()));
        return s;
    }
    b2g_batch* handle() { return _batch; }

private:
    friend class World;
    friend class Object;
    friend class Link;
    friend class Joint;
    void finishLoad() {
        check(b2g_model_create(_env, &_model));
        check(b2g_model_get_info(_model, &_info));
        _objects.clear();
        for (int i = 0; i < _info.n_objects; ++i) {
            b2g_object_info oi;
            check(b2g_model_get_object(_model, i, &oi));
            ObjectInfo o;
            o.name = oi.name;
            o.is_robot = oi.is_robot != 0;
            o.is_static = oi.is_static != 0;
            o.n_dofs = oi.n_dofs;
            o.dof_offset = oi.dof_offset;
            o.first_body = oi.first_body;
            o.n_links = oi.n_links;
            o.n_joints = oi.n_joints;
            o.n_balls = oi.n_balls;
            _objects.push_back(o);
        }
        _active.assign(_objects.size(), std::vector<int>());
        for (size_t i = 0; i < _objects.size(); ++i)  // all DOFs are active after construction (:1333)
            for (int d = 0; d < _objects[i].n_dofs; ++d) _active[i].push_back(d);
        _limits.assign((size_t)_info.nq * 6, 0.0f);
        if (_info.nq > 0) check(b2g_model_get_dof_limits(_model, _limits.data()));
        // what the reference logs while it loads a world goes to the logger here too
        const std::vector<std::string> warnings = getLoadWarnings();
        for (size_t i = 0; i < warnings.size(); ++i) _logger.logWarn(warnings[i], "[sim_env::b2g::BatchBox2DWorld::loadWorld]");
        check(b2g_batch_create(_model, _n, _device, _max_contacts, &_batch));
    }
    // scope guard: per-world calls act on world i only while it lives
    struct Select {
        Select(const BatchBox2DWorld* b, int64_t i) : _b(b) { check(b2g_batch_select_worlds(_b->_batch, i, 1)); }
        ~Select() { b2g_batch_select_worlds(_b->_batch, 0, 0); }
        const BatchBox2DWorld* _b;
    };
    static void check(b2g_status s) {
        if (s == B2G_OK) return;
        std::string msg = b2g_last_error();
        if (s == B2G_ERR_UNSUPPORTED) throw std::logic_error(msg);
        throw std::runtime_error(msg);
    }
    void pushParams() {
        if (_batch) check(b2g_batch_set_params(_batch, _dt, _vs, _ps));
    }
    void pack(const std::vector<int32_t>& counts, const std::vector<int32_t>& io, const std::vector<float>& fo, int max_per_world,
              std::vector<Contact>& contacts, std::vector<int64_t>& offsets) const {
        contacts.clear();
        offsets.assign((size_t)_n + 1, 0);
        for (int64_t w = 0; w < _n; ++w) {
            int n = counts[w] < max_per_world ? counts[w] : max_per_world;
            for (int k = 0; k < n; ++k) {
                const int32_t* i4 = &io[((size_t)w * max_per_world + k) * 4];
                const float* f6 = &fo[((size_t)w * max_per_world + k) * 6];
                Contact c;
                for (int d = 0; d < 3; ++d) { c.contact_point[d] = f6[d]; c.contact_normal[d] = f6[3 + d]; }
                c.body_a = i4[0]; c.body_b = i4[1]; c.fixture_a = i4[2]; c.fixture_b = i4[3];
                contacts.push_back(c);
            }
            offsets[w + 1] = (int64_t)contacts.size();
        }
    }
    void destroy() {
        if (_batch) b2g_batch_destroy(_batch);
        if (_model) b2g_model_destroy(_model);
        if (_env) b2g_env_destroy(_env);
        _batch = nullptr;
        _model = nullptr;
        _env = nullptr;
    }

    int64_t _n;
    int _device, _max_contacts;
    float _dt = 0.01f;
    int _vs = 15, _ps = 15;
    b2g_env* _env = nullptr;
    b2g_model* _model = nullptr;
    b2g_batch* _batch = nullptr;
    b2g_model_info _info{};
    std::string _path;
    std::vector<ObjectInfo> _objects;
    std::vector<std::vector<int> > _active;  // active DOFs per object (Box2DObject::_active_dof_indices)
    mutable Logger _logger;
    std::vector<float> _limits;              // nq x 6 (b2g_model_get_dof_limits)
};

// ------------------------------------------------------------------------------------------------ Object (one world)
inline const std::string& Object::getName() const { return _batch->_objects[_object].name; }
inline EntityType Object::getType() const { return _batch->_objects[_object].is_robot ? EntityType::Robot : EntityType::Object; }
inline bool Object::isStatic() const { return _batch->_objects[_object].is_static; }
inline unsigned int Object::getNumDOFs() const { return (unsigned int)_batch->_objects[_object].n_dofs; }
inline std::vector<int> Object::getDOFIndices() const {
    std::vector<int> v(getNumDOFs());
    for (size_t i = 0; i < v.size(); ++i) v[i] = (int)i;
    return v;
}
inline std::vector<int> Object::getActiveDOFs() const { return _batch->_active[_object]; }
inline void Object::setActiveDOFs(const std::vector<int>& indices) {
    for (size_t i = 0; i < indices.size(); ++i)
        if (indices[i] < 0 || indices[i] >= (int)getNumDOFs() || (i > 0 && indices[i] <= indices[i - 1]))
            throw std::runtime_error("[sim_env::b2g::Object::setActiveDOFs] indices must be ascending DOF indices of the object");
    _batch->_active[_object] = indices;
}
inline DOFInformation Object::getDOFInformation(unsigned int dof_index) const {
    if (dof_index >= getNumDOFs()) throw std::runtime_error("[sim_env::b2g::Object::getDOFInformation] invalid DOF index");
    DOFInformation d;
    const float* l = &_batch->_limits[(size_t)(_batch->_objects[_object].dof_offset + (int)dof_index) * 6];
    d.dof_index = (int)dof_index;
    for (int k = 0; k < 2; ++k) {
        d.position_limits[k] = l[k];
        d.velocity_limits[k] = l[2 + k];
        d.acceleration_limits[k] = l[4 + k];
    }
    d.cyclic = !isStatic() && dof_index == 2;  // base rotation (:1355)
    return d;
}
inline std::vector<float> Object::getDOFPositions(const std::vector<int>& indices) const {
    const std::vector<int>& ix = indices.empty() ? _batch->_active[_object] : indices;
    std::vector<float> out(ix.size());
    if (ix.empty()) return out;
    BatchBox2DWorld::Select sel(_batch, _world);
    std::vector<int32_t> i32(ix.begin(), ix.end());
    BatchBox2DWorld::check(b2g_batch_get_object_dofs(_batch->_batch, _object, 0, i32.data(), (int)i32.size(), out.data()));
    return out;
}
inline std::vector<float> Object::getDOFVelocities(const std::vector<int>& indices) const {
    const std::vector<int>& ix = indices.empty() ? _batch->_active[_object] : indices;
    std::vector<float> out(ix.size());
    if (ix.empty()) return out;
    BatchBox2DWorld::Select sel(_batch, _world);
    std::vector<int32_t> i32(ix.begin(), ix.end());
    BatchBox2DWorld::check(b2g_batch_get_object_dofs(_batch->_batch, _object, 1, i32.data(), (int)i32.size(), out.data()));
    return out;
}
inline void Object::setDOFPositions(const std::vector<float>& values, const std::vector<int>& indices) {
    const std::vector<int>& ix = indices.empty() ? _batch->_active[_object] : indices;
    if (ix.size() != values.size())
        throw std::runtime_error("[sim_env::Box2DObject::setDOFPositions] Could not set DOF positions."
                                 " Number of indices to set and number of provided values do not match.");
    if (ix.empty()) return;
    BatchBox2DWorld::Select sel(_batch, _world);
    std::vector<int32_t> i32(ix.begin(), ix.end());
    BatchBox2DWorld::check(b2g_batch_set_object_dofs(_batch->_batch, _object, 0, i32.data(), (int)i32.size(), values.data()));
}
inline void Object::setDOFVelocities(const std::vector<float>& values, const std::vector<int>& indices) {
    const std::vector<int>& ix = indices.empty() ? _batch->_active[_object] : indices;
    if (ix.size() != values.size())  // (:1700-1704)
        throw std::runtime_error("[sim_env::Box2DObject::setDOFVelocities] Could not set DOF velocities."
                                 " Number of indices to set and number of provided values do not match.");
    if (ix.empty()) return;
    BatchBox2DWorld::Select sel(_batch, _world);
    std::vector<int32_t> i32(ix.begin(), ix.end());
    BatchBox2DWorld::check(b2g_batch_set_object_dofs(_batch->_batch, _object, 1, i32.data(), (int)i32.size(), values.data()));
}
inline void Object::getPose(float pose[3]) const {
    BatchBox2DWorld::Select sel(_batch, _world);
    BatchBox2DWorld::check(b2g_batch_get_object_pose(_batch->_batch, _object, pose));
}
inline void Object::setPose(const float pose[3]) {
    BatchBox2DWorld::Select sel(_batch, _world);
    BatchBox2DWorld::check(b2g_batch_set_object_pose(_batch->_batch, _object, pose));
}
inline void Object::getState(ObjectState& state) const {
    state.active_dofs = _batch->_active[_object];
    state.dof_positions = getNumDOFs() ? getDOFPositions(getDOFIndices()) : std::vector<float>();
    state.dof_velocities = getNumDOFs() ? getDOFVelocities(getDOFIndices()) : std::vector<float>();
    getPose(state.pose);
}
inline void Object::setState(const ObjectState& state) {
    if (state.dof_positions.size() != getNumDOFs() || state.dof_velocities.size() != getNumDOFs())
        throw std::runtime_error("[sim_env::b2g::Object::setState] the state must hold every DOF of the object");
    setActiveDOFs(state.active_dofs);
    setPose(state.pose);  // setTransform(object_state.pose) (:1984)
    if (getNumDOFs()) {
        setDOFPositions(state.dof_positions, getDOFIndices());
        setDOFVelocities(state.dof_velocities, getDOFIndices());
    }
}
inline void Object::getBallApproximation(std::vector<Ball>& balls) const {
    const int n = _batch->_objects[_object].n_balls;
    balls.resize((size_t)n);
    if (n == 0) return;
    std::vector<float> raw((size_t)n * 4);
    BatchBox2DWorld::Select sel(_batch, _world);
    BatchBox2DWorld::check(b2g_batch_get_ball_approximation(_batch->_batch, _object, raw.data()));
    for (int i = 0; i < n; ++i) {
        for (int d = 0; d < 3; ++d) balls[(size_t)i].center[d] = raw[(size_t)i * 4 + d];
        balls[(size_t)i].radius = raw[(size_t)i * 4 + 3];
    }
}
inline void Object::setTargetVelocity(const std::vector<float>& target) {
    if (!_batch->_objects[_object].is_robot) throw std::logic_error("[sim_env::b2g::Object::setTargetVelocity] not a robot");
    if (target.size() != getNumDOFs())
        throw std::runtime_error("[sim_env::b2g::Object::setTargetVelocity] one target value per DOF of the robot is needed");
    BatchBox2DWorld::Select sel(_batch, _world);
    BatchBox2DWorld::check(b2g_batch_set_target_velocity(_batch->_batch, _object, target.data()));
}
inline bool Object::checkCollision() { return World(_batch, _world).checkCollision(*this); }
inline bool Object::checkCollision(std::vector<Contact>& contacts) { return World(_batch, _world).checkCollision(*this, contacts); }
inline bool Object::checkCollision(const Object& other) { return World(_batch, _world).checkCollision(*this, other); }
inline bool Object::checkCollision(const Object& other, std::vector<Contact>& contacts) {
    return World(_batch, _world).checkCollision(*this, other, contacts);
}
inline bool Object::checkCollision(const std::vector<Object>& others) { return World(_batch, _world).checkCollision(*this, others); }
inline bool Object::checkCollision(const std::vector<Object>& others, std::vector<Contact>& contacts) {
    return World(_batch, _world).checkCollision(*this, others, contacts);
}

inline void BatchBox2DWorld::printWorldState(int64_t world_index, std::FILE* out) {
    WorldState state = world(world_index).getWorldState();
    std::fprintf(out, "World state: \n");
    for (WorldState::const_iterator it = state.begin(); it != state.end(); ++it) {
        std::fprintf(out, "%s positions:", it->first.c_str());
        for (size_t i = 0; i < it->second.dof_positions.size(); ++i) std::fprintf(out, " %g", it->second.dof_positions[i]);
        std::fprintf(out, " velocities:");
        for (size_t i = 0; i < it->second.dof_velocities.size(); ++i) std::fprintf(out, " %g", it->second.dof_velocities[i]);
        std::fprintf(out, " active_dofs:");
        for (size_t i = 0; i < it->second.active_dofs.size(); ++i) std::fprintf(out, " %d", it->second.active_dofs[i]);
        std::fprintf(out, "\n");
    }
}

// ------------------------------------------------------------------------------------------------- World (one world)
// ---- Link / Joint views ------------------------------------------------------------------------------------------------------
inline Logger& World::getLogger() const { return _batch->getLogger(); }
inline std::string Link::getName() const { return _batch->getLinkInfo(_body).name; }
inline int Link::objectIndex() const { return _batch->getLinkInfo(_body).object; }
inline void Link::row(float out[8]) const {
    BatchBox2DWorld::Select sel(_batch, _world);
    std::vector<float> all((size_t)_batch->numBodies() * 8);
    BatchBox2DWorld::check(b2g_batch_get_link_states(_batch->_batch, all.data()));
    for (int k = 0; k < 8; ++k) out[k] = all[(size_t)_body * 8 + k];
}
inline void Link::getPose(float pose[3]) const {
    float r[8];
    row(r);
    pose[0] = r[0]; pose[1] = r[1]; pose[2] = r[2];
}
inline void Link::getVelocityVector(float vel[3]) const {
    float r[8];
    row(r);
    vel[0] = r[3]; vel[1] = r[4]; vel[2] = r[5];
}
inline void Link::getCenterOfMass(float com[2]) const {
    float r[8];
    row(r);
    com[0] = r[6]; com[1] = r[7];
}
inline float Link::getMass() const { return _batch->getMass(_body); }
inline float Link::getInertia() const { return _batch->getInertia(_body); }
inline float Link::getGroundFrictionCoefficient() const { return _batch->getGroundFrictionCoefficient(_body); }
inline float Link::getGroundFrictionTorqueIntegral() const { return _batch->getGroundFrictionTorqueIntegral(_body); }
inline float Link::getGroundFrictionLimitForce() const { return _batch->getGroundFrictionLimitForce(_body); }
inline float Link::getGroundFrictionLimitTorque() const { return _batch->getGroundFrictionLimitTorque(_body); }
inline float Link::getContactFriction() const { return _batch->getContactFriction(_body); }
inline bool Link::getParentJoint(Joint& joint) const {
    const int j = _batch->getLinkInfo(_body).parent_joint;
    if (j < 0) return false;
    joint = Joint(_batch, _world, j);
    return true;
}
inline std::vector<Joint> Link::getChildJoints() const {
    std::vector<Joint> out;
    const b2g_link_info li = _batch->getLinkInfo(_body);
    const ObjectInfo& o = _batch->_objects[(size_t)li.object];
    for (int b = o.first_body; b < o.first_body + o.n_links; ++b) {
        const int j = _batch->getLinkInfo(b).parent_joint;
        if (j >= 0 && _batch->getJointInfo(j).body_a == _body) out.push_back(Joint(_batch, _world, j));
    }
    return out;
}
inline bool Link::checkCollision() { return World(_batch, _world).checkCollisionLink(_body); }
inline bool Link::checkCollision(std::vector<Contact>& contacts) { return World(_batch, _world).checkCollisionLink(_body, contacts); }
inline bool Link::checkCollision(const std::vector<Link>& other_links, std::vector<Contact>* contacts) {
    std::vector<int> bodies;
    for (size_t i = 0; i < other_links.size(); ++i) bodies.push_back(other_links[i]._body);
    return World(_batch, _world).checkCollisionLink(_body, bodies, contacts);
}

inline std::string Joint::getName() const { return _batch->getJointInfo(_joint).name; }
inline Joint::JointType Joint::getJointType() const { return _batch->getJointInfo(_joint).kind == 2 ? JointType::Prismatic : JointType::Revolute; }
inline unsigned int Joint::getDOFIndex() const {
    const b2g_joint_info ji = _batch->getJointInfo(_joint);
    return (unsigned int)(ji.dof_index - _batch->_objects[(size_t)ji.object].dof_offset);
}
inline unsigned int Joint::getJointIndex() const {
    const b2g_joint_info ji = _batch->getJointInfo(_joint);
    return getDOFIndex() - (_batch->_objects[(size_t)ji.object].is_static ? 0u : 3u);
}
inline float Joint::getPosition() const {
    const std::vector<int> ix(1, (int)getDOFIndex());
    return Object(_batch, _world, _batch->getJointInfo(_joint).object).getDOFPositions(ix)[0];
}
inline float Joint::getVelocity() const {
    const std::vector<int> ix(1, (int)getDOFIndex());
    return Object(_batch, _world, _batch->getJointInfo(_joint).object).getDOFVelocities(ix)[0];
}
inline void Joint::setPosition(float v) {
    Object(_batch, _world, _batch->getJointInfo(_joint).object).setDOFPositions(std::vector<float>(1, v), std::vector<int>(1, (int)getDOFIndex()));
}
inline void Joint::setVelocity(float v) {
    Object(_batch, _world, _batch->getJointInfo(_joint).object).setDOFVelocities(std::vector<float>(1, v), std::vector<int>(1, (int)getDOFIndex()));
}
inline DOFInformation Joint::getDOFInformation() const {
    return Object(_batch, _world, _batch->getJointInfo(_joint).object).getDOFInformation(getDOFIndex());
}
inline void Joint::getPositionLimits(float limits[2]) const {
    const DOFInformation d = getDOFInformation();
    limits[0] = d.position_limits[0]; limits[1] = d.position_limits[1];
}
inline void Joint::getVelocityLimits(float limits[2]) const {
    const DOFInformation d = getDOFInformation();
    limits[0] = d.velocity_limits[0]; limits[1] = d.velocity_limits[1];
}
inline void Joint::getAccelerationLimits(float limits[2]) const {
    const DOFInformation d = getDOFInformation();
    limits[0] = d.acceleration_limits[0]; limits[1] = d.acceleration_limits[1];
}
inline Link Joint::getChildLink() const { return Link(_batch, _world, _batch->getJointInfo(_joint).body_b); }
inline Link Joint::getParentLink() const { return Link(_batch, _world, _batch->getJointInfo(_joint).body_a); }

inline Link Object::getLink(const std::string& link_name) const { return Link(_batch, _world, _batch->getLink(getName(), link_name)); }
inline Link Object::getBaseLink() const { return Link(_batch, _world, _batch->getBaseLink(getName())); }
inline std::vector<Link> Object::getLinks() const {
    std::vector<Link> out;
    const std::vector<int> bodies = _batch->getLinks(getName());
    for (size_t i = 0; i < bodies.size(); ++i) out.push_back(Link(_batch, _world, bodies[i]));
    return out;
}
inline Joint Object::getJoint(const std::string& joint_name) const { return Joint(_batch, _world, _batch->getJoint(getName(), joint_name)); }
inline std::vector<Joint> Object::getJoints() const {
    std::vector<Joint> out;
    const ObjectInfo& o = _batch->_objects[(size_t)_object];
    for (int b = o.first_body; b < o.first_body + o.n_links; ++b) {
        const int j = _batch->getLinkInfo(b).parent_joint;
        if (j >= 0) out.push_back(Joint(_batch, _world, j));
    }
    return out;
}

inline Object World::getRobot(const std::string& name) const {
    const int o = _batch->objectIndex(name);
    if (!_batch->_objects[(size_t)o].is_robot) throw std::runtime_error("[sim_env::b2g::World::getRobot] '" + name + "' is not a robot");
    return Object(_batch, _world, o);
}
inline Object World::getObject(const std::string& name, bool exclude_robots) const {
    const int o = _batch->objectIndex(name);
    if (exclude_robots && _batch->_objects[(size_t)o].is_robot)
        throw std::runtime_error("[sim_env::b2g::World::getObject] '" + name + "' is a robot");
    return Object(_batch, _world, o);
}
inline void World::getObjects(std::vector<Object>& objects, bool exclude_robots) const {
    for (size_t o = 0; o < _batch->_objects.size(); ++o)
        if (!exclude_robots || !_batch->_objects[o].is_robot) objects.push_back(Object(_batch, _world, (int)o));
}
inline void World::getRobots(std::vector<Object>& robots) const {
    for (size_t o = 0; o < _batch->_objects.size(); ++o)
        if (_batch->_objects[o].is_robot) robots.push_back(Object(_batch, _world, (int)o));
}
inline void World::stepPhysics(int steps, bool execute_controllers, bool allow_sleeping) {
    if (_batch->_n != 1)
        throw std::logic_error("[sim_env::b2g::World::stepPhysics] stepping advances every world of the batch: call "
                               "BatchBox2DWorld::stepPhysics (a view steps only when the batch holds one world)");
    _batch->stepPhysics(steps, execute_controllers, allow_sleeping);
}
inline bool World::stepPhysics(std::vector<Contact>& contacts, int steps, bool execute_controllers, bool allow_sleeping) {
    if (_batch->_n != 1)
        throw std::logic_error("[sim_env::b2g::World::stepPhysics] stepping advances every world of the batch: call "
                               "BatchBox2DWorld::stepPhysics (a view steps only when the batch holds one world)");
    std::vector<int64_t> offsets;
    _batch->stepPhysics(contacts, offsets, steps, execute_controllers, allow_sleeping);
    return true;
}
inline float World::getPhysicsTimeStep() const { return _batch->getPhysicsTimeStep(); }
inline int World::getVelocitySteps() const { return _batch->getVelocitySteps(); }
inline int World::getPositionSteps() const { return _batch->getPositionSteps(); }
inline void World::getWorldState(WorldState& state) const {
    for (size_t o = 0; o < _batch->_objects.size(); ++o) state[_batch->_objects[o].name] = Object(_batch, _world, (int)o).getState();
}
inline bool World::setWorldState(const WorldState& state) {
    bool ok = true;
    for (WorldState::const_iterator it = state.begin(); it != state.end(); ++it) {
        int o = -1;
        for (size_t i = 0; i < _batch->_objects.size(); ++i)
            if (_batch->_objects[i].name == it->first) o = (int)i;
        if (o < 0) {  // "Unknown object encountered!" (:3505-3508): warn and report failure
            ok = false;
            continue;
        }
        Object(_batch, _world, o).setState(it->second);
    }
    return ok;
}
inline std::vector<char> World::bodiesOf(const Object& o) const {
    std::vector<char> in((size_t)_batch->_info.n_bodies, 0);
    const ObjectInfo& oi = _batch->_objects[(size_t)o.index()];
    for (int b = oi.first_body; b < oi.first_body + oi.n_links; ++b) in[(size_t)b] = 1;
    return in;
}
inline bool World::checkCollision(std::vector<Contact>& contacts) {
    const int cap = 128;  // kMaxContacts
    std::vector<int32_t> io((size_t)cap * 4);
    std::vector<float> fo((size_t)cap * 6);
    int32_t count = 0;
    {
        BatchBox2DWorld::Select sel(_batch, _world);
        BatchBox2DWorld::check(b2g_batch_check_collision(_batch->_batch, &count, io.data(), fo.data(), cap));
    }
    const int n = count < cap ? count : cap;
    for (int k = 0; k < n; ++k) {
        Contact c;
        for (int d = 0; d < 3; ++d) { c.contact_point[d] = fo[(size_t)k * 6 + d]; c.contact_normal[d] = fo[(size_t)k * 6 + 3 + d]; }
        c.body_a = io[(size_t)k * 4]; c.body_b = io[(size_t)k * 4 + 1];
        c.fixture_a = io[(size_t)k * 4 + 2]; c.fixture_b = io[(size_t)k * 4 + 3];
        contacts.push_back(c);
    }
    return n > 0;
}
inline bool World::filterContacts(const std::vector<char>& in_a, const std::vector<char>* in_b, std::vector<Contact>* out,
                                  std::vector<char>* hit_bodies) {
    std::vector<Contact> all;
    checkCollision(all);
    bool hit = false;
    for (size_t k = 0; k < all.size(); ++k) {
        for (int flip = 0; flip < 2; ++flip) {  // the cache holds every contact under both of its bodies (:2904-2905)
            Contact c = all[k];
            if (flip) { int t = c.body_a; c.body_a = c.body_b; c.body_b = t; t = c.fixture_a; c.fixture_a = c.fixture_b; c.fixture_b = t; }
            if (!in_a[(size_t)c.body_a]) continue;
            if (in_b && !(*in_b)[(size_t)c.body_b]) continue;
            hit = true;
            if (out) out->push_back(c);
            if (hit_bodies) (*hit_bodies)[(size_t)c.body_b] = 1;
        }
    }
    return hit;
}
inline bool World::checkCollision(const Object& a, const Object& b) {
    std::vector<char> ia = bodiesOf(a), ib = bodiesOf(b);
    return filterContacts(ia, &ib, nullptr);
}
inline bool World::checkCollision(const Object& a, const Object& b, std::vector<Contact>& contacts) {
    std::vector<char> ia = bodiesOf(a), ib = bodiesOf(b);
    return filterContacts(ia, &ib, &contacts);
}
inline bool World::checkCollision(const Object& object) { return filterContacts(bodiesOf(object), nullptr, nullptr); }
inline bool World::checkCollision(const Object& object, std::vector<Contact>& contacts) {
    return filterContacts(bodiesOf(object), nullptr, &contacts);
}
inline bool World::checkCollision(const Object& object, std::vector<Object>& collidors) {
    std::vector<char> hitBodies((size_t)_batch->_info.n_bodies, 0);
    const bool hit = filterContacts(bodiesOf(object), nullptr, nullptr, &hitBodies);
    for (size_t o = 0; o < _batch->_objects.size(); ++o) {
        const ObjectInfo& oi = _batch->_objects[o];
        bool any = false;
        for (int b = oi.first_body; b < oi.first_body + oi.n_links; ++b) any = any || hitBodies[(size_t)b];
        if (any) collidors.push_back(Object(_batch, _world, (int)o));
    }
    return hit;
}
inline bool World::checkCollision(const Object& a, const std::vector<Object>& others) {
    std::vector<char> ib((size_t)_batch->_info.n_bodies, 0);
    for (size_t i = 0; i < others.size(); ++i) {
        std::vector<char> t = bodiesOf(others[i]);
        for (size_t b = 0; b < t.size(); ++b) ib[b] = ib[b] || t[b];
    }
    return filterContacts(bodiesOf(a), &ib, nullptr);
}
inline bool World::checkCollision(const Object& a, const std::vector<Object>& others, std::vector<Contact>& contacts) {
    std::vector<char> ib((size_t)_batch->_info.n_bodies, 0);
    for (size_t i = 0; i < others.size(); ++i) {
        std::vector<char> t = bodiesOf(others[i]);
        for (size_t b = 0; b < t.size(); ++b) ib[b] = ib[b] || t[b];
    }
    return filterContacts(bodiesOf(a), &ib, &contacts);
}
inline bool World::checkCollisionLink(int link_body) {
    std::vector<char> ia((size_t)_batch->_info.n_bodies, 0);
    ia.at((size_t)link_body) = 1;
    return filterContacts(ia, nullptr, nullptr);
}
inline bool World::checkCollisionLink(int link_body, std::vector<Contact>& contacts) {
    std::vector<char> ia((size_t)_batch->_info.n_bodies, 0);
    ia.at((size_t)link_body) = 1;
    return filterContacts(ia, nullptr, &contacts);
}
inline bool World::checkCollisionLink(int link_body, const std::vector<int>& other_link_bodies, std::vector<Contact>* contacts) {
    std::vector<char> ia((size_t)_batch->_info.n_bodies, 0), ib((size_t)_batch->_info.n_bodies, 0);
    ia.at((size_t)link_body) = 1;
    for (size_t i = 0; i < other_link_bodies.size(); ++i) ib.at((size_t)other_link_bodies[i]) = 1;
    return filterContacts(ia, &ib, contacts);
}
inline bool World::checkCollisionLink(int link_body, const std::vector<Object>& other_objects, std::vector<Contact>* contacts) {
    std::vector<char> ia((size_t)_batch->_info.n_bodies, 0), ib((size_t)_batch->_info.n_bodies, 0);
    ia.at((size_t)link_body) = 1;
    for (size_t i = 0; i < other_objects.size(); ++i) {
        std::vector<char> t = bodiesOf(other_objects[i]);
        for (size_t b = 0; b < t.size(); ++b) ib[b] = ib[b] || t[b];
    }
    return filterContacts(ia, &ib, contacts);
}
inline bool World::checkCollision(const Object& a, int link_body_b, std::vector<Contact>* contacts) {
    std::vector<char> ib((size_t)_batch->_info.n_bodies, 0);
    ib.at((size_t)link_body_b) = 1;
    return filterContacts(bodiesOf(a), &ib, contacts);
}
inline bool World::atRest(float threshold) const {
    int32_t out = 0;
    BatchBox2DWorld::Select sel(_batch, _world);
    BatchBox2DWorld::check(b2g_batch_at_rest(_batch->_batch, threshold, &out));
    return out != 0;
}
inline bool World::isPhysicallyFeasible() const {
    int32_t out = 0;
    BatchBox2DWorld::Select sel(_batch, _world);
    BatchBox2DWorld::check(b2g_batch_is_physically_feasible(_batch->_batch, &out));
    return out != 0;
}
inline void World::getObjects(const float aabb[4], std::vector<Object>& objects, bool exclude_robots) const {
    std::vector<uint8_t> hits(_batch->_objects.size());
    {
        BatchBox2DWorld::Select sel(_batch, _world);
        BatchBox2DWorld::check(b2g_batch_objects_in_aabb(_batch->_batch, aabb, exclude_robots, hits.data()));
    }
    for (size_t o = 0; o < hits.size(); ++o)
        if (hits[o]) objects.push_back(Object(_batch, _world, (int)o));
}

}  // namespace b2g
}  // namespace sim_env

#endif  // B2G_WORLD_HPP_
