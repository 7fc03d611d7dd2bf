// Host-side environment description of the batched backend.
// Mirrors the reference's description structs (include/sim_env/Box2DIOUtils.h:18-70): link / joint /
// object / robot / state / environment.  Eigen is not available; plain arrays are used.
#pragma once
#include <array>
#include <map>
#include <string>
#include <vector>

namespace b2g {

struct LinkDescription {
    std::string name;
    std::vector<std::vector<float>> polygons;  // x0,y0,x1,y1,... (user units)
    float mass = 0.0f;
    float ground_friction = 0.0f;
    float ground_torque_integral = 0.0f;
    float contact_friction = 0.0f;
    float restitution = 0.0f;
    // ball approximation: centre (x, y, z) in the link frame and radius, user units (Box2DIOUtils.h:21, 139-142)
    std::vector<std::array<float, 4>> balls;
};

struct JointDescription {
    std::string name, link_a, link_b;
    float axis[2] = {0.0f, 0.0f};
    float position_limits[2] = {0.0f, 0.0f};
    float velocity_limits[2] = {0.0f, 0.0f};
    float acceleration_limits[2] = {0.0f, 0.0f};
    std::string joint_type = "revolute";
    bool actuated = false;
    float axis_orientation = 0.0f;
};

struct ObjectDescription {
    std::string name;
    std::string base_link;
    std::vector<LinkDescription> links;
    std::vector<JointDescription> joints;
    bool is_static = false;
    // robot part (Box2DRobotDescription)
    bool is_robot = false;
    float translational_velocity_limits[2] = {0.0f, 0.0f};
    float translational_acceleration_limits[2] = {0.0f, 0.0f};
    float rotational_velocity_limits[2] = {0.0f, 0.0f};
    float rotational_acceleration_limits[2] = {0.0f, 0.0f};
    bool use_center_of_mass = true;
};

struct StateDescription {
    std::vector<float> configuration, velocity;
};

struct EnvironmentDescription {
    float scale = 1.0f;
    float world_bounds[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    std::map<std::string, StateDescription> states;
    std::vector<ObjectDescription> robots;
    std::vector<ObjectDescription> objects;
    std::vector<std::string> warnings;  // loader diagnostics (alias keys, ignored keys)
};

// yaml_env.cpp — throws std::runtime_error on I/O or decode failure
void parseYAML(const std::string& filename, EnvironmentDescription& ed);
// kinematic-tree validation of Box2DIOUtils.h:212-257; fills base_link; returns false + message if invalid
bool verifyKinematicStructure(ObjectDescription& od, std::string& message);

}  // namespace b2g
