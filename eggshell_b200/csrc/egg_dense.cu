// Dense mixed-LCP path (placeholder until the Murty kernel lands).
#include "egg_internal.cuh"
size_t egg_dense_scratch_bytes(const EggDev& d) { return 8; }
void egg_launch_solve_dense(const EggDev& d, double dt, cudaStream_t s, void* scratch, size_t scratch_bytes) {}
