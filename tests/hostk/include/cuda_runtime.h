// TEST INFRASTRUCTURE ONLY -- stands in for <cuda_runtime.h> when the kernel SOURCES under
// mpconstellation_b200/csrc/*.cuh are compiled for the host by tests/hostk (see hostk_shim.h).
#pragma once
#include "hostk_shim.h"
