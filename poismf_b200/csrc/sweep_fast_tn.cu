#define PMF_INST_STRICT 0
#define PMF_INST_TN 1
#include "sweep_inst.cuh"
