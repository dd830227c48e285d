// (all entry points of include/gp3d_b200.h are implemented; this translation unit is intentionally empty)
#include "common.cuh"
