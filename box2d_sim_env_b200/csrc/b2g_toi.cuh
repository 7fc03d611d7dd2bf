// Continuous collision (b2World::SolveTOI) — placeholder until the TOI kernels land; see b2g_toi.cuh history.
#pragma once
#include "b2g_solver.cuh"

namespace b2g {
DEV void solveTOI(const Ctx& c) { (void)c; }
}  // namespace b2g
