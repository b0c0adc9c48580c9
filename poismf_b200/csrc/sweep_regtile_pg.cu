#define PMF_RT_METHOD 3
#include "sweep_regtile.cuh"
