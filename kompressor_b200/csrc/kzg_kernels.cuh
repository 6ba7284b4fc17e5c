// Kernel-side view of one batch and the launch entry point (implemented in kzg_kernels.cu).
#pragma once
#include "kzg_common.cuh"
#include <cuda_runtime.h>

struct KzgLaunchInfo { int sm_count; int launches; };

// Enqueues the whole decode pipeline for one batch on `stream`. Returns number of kernel launches.
#define KZG_N_STAGES 11
// If ev != nullptr it must hold KZG_N_STAGES + 1 events; event k is recorded before stage k.
int kzg_launch_batch(const KzgBatch& b, int sm_count, cudaStream_t stream, cudaEvent_t* ev);
extern const char* const kzg_stage_names[KZG_N_STAGES];
// Opts the kernels into their dynamic shared-memory sizes on `device` (the current device); idempotent, thread-safe.
int kzg_kernels_init_device(int device);
