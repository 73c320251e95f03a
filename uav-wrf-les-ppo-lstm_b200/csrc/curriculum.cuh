// curriculum.cuh -- the interface between the curriculum kernels (learner_kernels.cu) and the peer-memory exchange
// (comm_kernels.cu): where the packed done/reached flags of every rank live.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace plume {

constexpr int kCommMaxWorld = 16;
constexpr int kCurMaxBlocks = 8192;        // curriculum windows one segment may complete

// flag codes (bit 0 done, bit 1 reached) of rank r: base[r][t * N + n].  One contiguous [world][T][N] array (what an
// all-gather produces) or one mapped peer buffer per rank (no gathered copy).
struct CodeSrc {
    const uint8_t* base[kCommMaxWorld];
};

// PPOTrainer.update (model.py:188-221) over the finished episodes of a [T][N] x world segment in canonical order
// (step-major, then global env id).  window_radius_out (may be NULL, double[kCurMaxBlocks + 2]): [0] = length of the
// carried partial window at entry, [1] = number of radii that follow, [2 + b] = the trainer's radius in force for the
// episodes whose ordinal (carried length + position in the segment) lies in window b -- what the reference logs as
// 'Current_Radius' (train_ppo2.0.py:247).  comm_error (may be NULL): device word; non-zero = the flags are not
// trustworthy, nothing is applied.
int launch_curriculum_packed(const CodeSrc& src, int horizon, int n_envs, int world, double* state, double* curriculum,
                             double initial_radius, double min_radius, double radius_decay, double success_threshold,
                             int window, double decay_factor, double* window_radius_out, const uint32_t* comm_error,
                             cudaStream_t stream);

// the published flag-code buffers of all ranks (after plume_comm_publish_codes + the wait of
// plume_curriculum_update_peer for the same segment of `bytes` bytes); 0 on success
int comm_code_sources(void* comm, int64_t bytes, CodeSrc* out, int* world, int* rank);

}  // namespace plume
