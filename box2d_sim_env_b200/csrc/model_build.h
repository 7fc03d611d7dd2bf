// Host model builder: EnvironmentDescription -> flattened model (setup time, CPU).
#pragma once
#include <string>
#include <vector>

#include "desc.h"
#include "model.h"

namespace b2g {

struct HostObjectInfo {
    std::string name;
    bool isRobot = false, isStatic = false;
    int nDofs = 0, dofOffset = 0, firstBody = 0, nLinks = 0, nJoints = 0;
    int ballStart = 0, nBalls = 0;  // rows of HostModel::balls
};

struct HostModel {
    ModelHeader h;
    std::vector<uint32_t> blob;
    std::vector<HostObjectInfo> objects;
    std::vector<std::string> bodyNames;
    // introspection copies (parity probes)
    std::vector<float> bodyModel;     // nb x 8
    std::vector<int32_t> fixInts;     // nf x 3
    std::vector<float> fixFloats;     // nf x 38
    std::vector<int32_t> jointInts;   // nj x 6
    std::vector<float> jointFloats;   // nj x 15
    std::vector<float> dofLimits;     // nq x 6
    // ball approximation (Box2DLink::_balls), object by object, links in name order (Box2DObject::getBallApproximation
    // walks its std::map of links, Box2DWorld.cpp:1754-1763): rows of 8 = body, origin offset x y (scaled), centre x y z, radius
    std::vector<float> balls;
    std::vector<std::string> warnings;
};

// throws std::runtime_error (unsupported joint type, capacity, malformed description)
void buildModel(const EnvironmentDescription& env, int maxContacts, HostModel& out);
// recompute the per-world state layout for a different contact capacity
void layoutState(ModelHeader& h, int maxContacts);

}  // namespace b2g
