// Kernel-side view of one batch and the launch entry point (implemented in kzg_kernels.cu).
#pragma once
#include "kzg_common.cuh"
#include <cuda_runtime.h>

struct KzgLaunchInfo { int sm_count; int launches; };

// Enqueues the whole decode pipeline for one batch on `stream`. Returns number of kernel launches.
int kzg_launch_batch(const KzgBatch& b, int sm_count, cudaStream_t stream);
