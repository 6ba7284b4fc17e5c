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

// Running XXH64 of a frame that is decoded in segments (kzg_decompress_segment_host): folds `len` more content bytes
// (device memory) into the state and keeps its digest current.  One warp; the state lives in device memory.
struct KzgXxhState { u64 acc[4]; u64 total; u8 tail[32]; u32 want, digest; };
void kzg_launch_xxh_segment(KzgXxhState* d_state, const u8* d_data, u64 len, cudaStream_t stream);
