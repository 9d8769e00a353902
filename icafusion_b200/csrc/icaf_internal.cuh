// Shared internals of libicaf_b200: error reporting, launch checks, device info.
#pragma once
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/icaf_b200.h"
#include "ptx.cuh"

namespace icaf {

int set_error(int code, const char* msg);
int set_cuda_error(cudaError_t e, const char* where);
int check_launch(const char* where);   // cudaGetLastError after a launch; never synchronises
int sm_count_cached();

}  // namespace icaf
