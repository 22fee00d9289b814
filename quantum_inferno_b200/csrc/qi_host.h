// qi_host.h -- host-side helpers shared by the translation units of libqi_b200.so
#pragma once
#include <stdio.h>
#include "qi_platform.cuh"

namespace qi {
extern thread_local char g_last_cuda_error[256];
int check_cuda(const char* where);
// Asynchronous copy of a small host-side table (band descriptors, index lists) to the device through kernel
// parameters: the caller's buffer may die immediately, the host never waits for the stream, and no copy engine is used.
// The size must be a multiple of 4 bytes.
void stage_to_device(void* dst, const void* src, size_t bytes, cudaStream_t st);
}  // namespace qi
