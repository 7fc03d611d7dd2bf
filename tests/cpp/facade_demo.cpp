// Exercises include/b2g_world.hpp the way a planner built on sim_env::Box2DWorld would (C++11, no CUDA headers needed).
//   facade_demo <env.yaml> host   -> model introspection works, creating the batch fails loudly without a GPU
//   facade_demo <env.yaml> gpu    -> steps 64 worlds and prints world 0's state as hex floats for the Python-side check
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "b2g_world.hpp"

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    const std::string path = argv[1];
    const bool gpu = std::strcmp(argv[2], "gpu") == 0;
    sim_env::b2g::BatchBox2DWorld batch(64, 0);
    try {
        batch.loadWorld(path);
    } catch (const std::runtime_error& e) {
        if (gpu) {
            std::printf("FAIL %s\n", e.what());
            return 1;
        }
        // no CUDA device: the model stage must still have succeeded and the failure must name CUDA
        bool named = std::string(e.what()).find("cuda") != std::string::npos || std::string(e.what()).find("CUDA") != std::string::npos;
        std::printf("HOST nq=%d bodies=%d objects=%zu first=%s cuda_named=%d\n", batch.nq(), batch.numBodies(), batch.getObjects().size(),
                    batch.getObjects().empty() ? "-" : batch.getObjects()[0].name.c_str(), (int)named);
        return named ? 0 : 1;
    }
    if (!gpu) {
        std::printf("HOST nq=%d bodies=%d objects=%zu first=%s cuda_named=-1\n", batch.nq(), batch.numBodies(), batch.getObjects().size(),
                    batch.getObjects()[0].name.c_str());
        return 0;
    }
    const int n = (int)batch.numWorlds(), nq = batch.nq();
    std::vector<float> q, qd;
    batch.getWorldState(q, qd);
    for (int w = 0; w < n; ++w) q[(size_t)w * nq + 7] = 0.2f + 0.01f * w;  // obj1 x: slides it towards the robot's fingers
    batch.setWorldState(q, qd, true);
    const int nd = batch.robotDofs("my_robot");
    std::vector<float> u((size_t)n * nd, 0.0f);
    for (int w = 0; w < n; ++w) { u[(size_t)w * nd] = 0.5f; u[(size_t)w * nd + 2] = 0.3f; }
    batch.setTargetVelocity("my_robot", u.data());
    batch.stepPhysics(10);
    batch.saveState();
    std::vector<sim_env::b2g::Contact> contacts;
    std::vector<int64_t> offsets;
    batch.stepPhysics(contacts, offsets, 20);
    size_t recorded = contacts.size();
    bool restored = batch.restoreState();
    batch.stepPhysics(20);
    batch.getWorldState(q, qd);
    std::vector<bool> hit = batch.checkCollision();
    int hits = 0;
    for (int w = 0; w < n; ++w) hits += hit[w] ? 1 : 0;
    std::printf("GPU recorded=%zu restored=%d hits=%d q0=", recorded, (int)restored, hits);
    for (int i = 0; i < nq; ++i) std::printf("%a ", q[i]);
    std::printf("\n");
    bool threw = false;
    try {
        std::vector<float> bad(3);
        batch.setWorldState(bad, bad);
    } catch (const std::runtime_error&) {
        threw = true;
    }
    return threw ? 0 : 1;
}
