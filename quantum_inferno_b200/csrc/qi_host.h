// qi_host.h -- host-side helpers shared by the translation units of libqi_b200.so
#pragma once
#include <stdio.h>
#include "qi_platform.cuh"

namespace qi {
extern thread_local char g_last_cuda_error[256];
int check_cuda(const char* where);
}  // namespace qi
