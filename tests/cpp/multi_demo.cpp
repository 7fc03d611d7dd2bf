// A C++ planner on several GPUs of one node through the plain C ABI (include/b2g.h, no CUDA or NCCL headers needed):
//   multi_demo <env.yaml> <n_devices> <n_worlds>
// shards the batch over devices 0 .. n_devices-1 (b2g_multi_*: one host thread and one batch per device, ncclAllGather of the
// final states), runs a 3 x 10-step rollout and checks it against the same rollout on a single batch, bit for bit.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "b2g.h"

#define CHECK(call)                                                        \
    do {                                                                   \
        if ((call) != B2G_OK) {                                            \
            std::printf("FAIL %s: %s\n", #call, b2g_last_error());         \
            return 1;                                                      \
        }                                                                  \
    } while (0)

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const int nDev = std::atoi(argv[2]);
    const long long n = std::atoll(argv[3]);
    b2g_env* env = nullptr;
    b2g_model* model = nullptr;
    CHECK(b2g_env_load_yaml(argv[1], &env));
    CHECK(b2g_model_create(env, &model));
    b2g_model_info info;
    CHECK(b2g_model_get_info(model, &info));
    b2g_object_info robot;
    CHECK(b2g_model_get_object(model, 0, &robot));
    const int nq = info.nq, nd = robot.n_dofs, nT = 3;

    std::vector<int32_t> devices(nDev);
    for (int i = 0; i < nDev; ++i) devices[i] = i;
    b2g_multi* multi = nullptr;
    CHECK(b2g_multi_create(model, n, devices.data(), nDev, 0, &multi));
    b2g_batch* single = nullptr;
    CHECK(b2g_batch_create(model, n, 0, 0, &single));

    // inputs: the loaded state with obj1 slid towards the fingers, per-world base / finger velocity targets
    std::vector<float> q((size_t)n * nq), qd((size_t)n * nq), tg((size_t)nT * n * nd, 0.0f);
    CHECK(b2g_batch_get_state(single, q.data(), qd.data()));
    for (long long w = 0; w < n; ++w) q[(size_t)w * nq + 7] = 0.1f + 0.002f * (float)(w % 200);
    for (int t = 0; t < nT; ++t)
        for (long long w = 0; w < n; ++w) {
            float* u = &tg[((size_t)t * n + w) * nd];
            u[0] = 0.3f * (float)(t + 1);
            u[1] = 0.5f;
            u[2] = 0.1f * (float)(w % 7);
            if (nd > 3) u[3] = (w & 1) ? 2.0f : -2.0f;
        }
    std::vector<float> mq((size_t)n * nq), mqd((size_t)n * nq), sq((size_t)n * nq), sqd((size_t)n * nq);
    CHECK(b2g_multi_rollout(multi, 0, q.data(), qd.data(), 1, tg.data(), nT, 10, mq.data(), mqd.data()));
    CHECK(b2g_batch_rollout(single, 0, q.data(), qd.data(), 1, tg.data(), nT, 10, sq.data(), sqd.data()));
    const bool same = std::memcmp(mq.data(), sq.data(), mq.size() * 4) == 0 && std::memcmp(mqd.data(), sqd.data(), mqd.size() * 4) == 0;
    float kmin = 0, kmax = 0, gather = 0;
    CHECK(b2g_multi_last_timing(multi, &kmin, &kmax, &gather));
    long long total = 0;
    for (int i = 0; i < b2g_multi_device_count(multi); ++i) {
        int64_t first = 0, count = 0;
        CHECK(b2g_multi_get_shard(multi, i, nullptr, &first, &count));
        if (first != total) return 1;
        total += count;
    }
    std::printf("MULTI devices=%d worlds=%lld covered=%lld identical=%d kernel_ms=[%.3f, %.3f] gather_ms=%.3f\n", nDev, n, total, (int)same,
                kmin, kmax, gather);
    b2g_multi_destroy(multi);
    b2g_batch_destroy(single);
    b2g_model_destroy(model);
    b2g_env_destroy(env);
    return same && total == n ? 0 : 1;
}
