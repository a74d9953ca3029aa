// placeholder until K7/K8 land
#include "svo_common.cuh"
