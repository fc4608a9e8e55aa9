// placeholder until the transport kernel lands (next commit)
#include "common.cuh"
