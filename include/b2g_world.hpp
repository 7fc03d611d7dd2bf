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
    int n_dofs = 0, dof_offset = 0, first_body = 0, n_links = 0, n_joints = 0;
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
        check(b2g_env_load_yaml(path.c_str(), &_env));
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
            _objects.push_back(o);
        }
        check(b2g_batch_create(_model, _n, _device, _max_contacts, &_batch));
    }

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

    // per-world status after the last call (contact capacity exceeded -> std::runtime_error)
    std::vector<int32_t> status() {
        std::vector<int32_t> s(_n);
        check(b2g_batch_status(_batch, s.data()));
        return s;
    }
    b2g_batch* handle() { return _batch; }

private:
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
    std::vector<ObjectInfo> _objects;
};

}  // namespace b2g
}  // namespace sim_env

#endif  // B2G_WORLD_HPP_
