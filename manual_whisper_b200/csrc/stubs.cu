// Every entry point of include/mw_b200.h is implemented; this file is intentionally empty of stubs.
#include "mw_common.cuh"
