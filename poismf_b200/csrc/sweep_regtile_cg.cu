#define PMF_RT_METHOD 2
#include "sweep_regtile.cuh"
