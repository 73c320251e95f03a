// comm_kernels.cu -- every exchange step of a data-parallel PPO iteration over NVLink peer memory, without a
// host-launched collective:
//
//   (1) gradient all-reduce + clip_grad_norm_ + Adam as ONE kernel (train_ppo2.0.py:85-87 on N ranks), 20 x per
//       iteration;
//   (2) the three doubles of the global advantage statistics (train_ppo2.0.py:34-38 over the whole batch), one tiny
//       kernel per iteration;
//   (3) the segment's packed done/reached flags for the curriculum (model.py:188-221 on the GLOBAL episode stream):
//       each rank copies its [T][N] bytes into its mapped block and signals; the curriculum kernels of every rank
//       then read all ranks' flags straight through the peer mappings -- there is no gathered copy at all.
//
// Every rank owns one cudaMalloc'd block that all peers have mapped (CUDA IPC, one process per GPU).  A rank
// publishes data in its OWN block, signals "published" with a system-scope release store into each PEER's flag
// array, waits (locally) for the flags of all ranks and then reads the peers' blocks.  For (1) CTA c sums slice c of
// all ranks' buffers in rank order (so every rank obtains the bitwise identical sum -- a one-shot all-reduce), the
// CTAs combine their slice norms through a grid barrier, and each thread finishes clip + Adam for its element from
// registers.  Buffers alternate by step parity, so one cross-GPU barrier per step is enough: a rank can only publish
// step k+2 into the buffer of step k after every peer has arrived at step k+1, i.e. finished reading step k.
//
// Failure handling: every spin is bounded.  A timeout sets the block's sticky error word; the kernel that saw it
// applies NOTHING (no parameter, moment or gradient is written from partial data), and every later exchange kernel
// of this communicator returns at once.  The host reads the word with plume_comm_error (PlumeTrainer does so once
// per iteration) and has to rebuild the exchange state collectively (plume_comm_reset on every rank between two
// host barriers) before it may continue.
//
// Algorithmic bytes per optimiser step and rank: world x 145 KB read (peer), 145 KB published, 16 B read + 12 B
// written per parameter.  Per iteration: T*N flag bytes copied locally, 2 x world x T*N read through the mappings
// (count pass + bin pass of the curriculum).
#include <cstring>

#include "common.cuh"
#include "curriculum.cuh"

namespace plume {

constexpr int kCommThreads = 1024;
constexpr int kCommSmallMax = 32;          // doubles per small all-reduce
constexpr uint32_t kCommSpinLimit = 1u << 26;

struct CommLayout {            // one cudaMalloc'd block per rank, mapped by every peer
    // [0]      gflags  uint32[16]   gflags[r] = last gradient step rank r has published (written by rank r)
    // [64]     sflags  uint32[16]   last small all-reduce rank r has published
    // [128]    cflags  uint32[16]   last flag-code segment rank r has published
    // [256]    grid    uint32[4]    [0] grid-barrier counter, [1] publish "blocks done" counter, [2] sticky error
    // [512]    partial double[64]   slice norms (local)
    // [1024]   small   double[2][kCommSmallMax]   published small vectors, by parity
    // [2048]   pub     float[2][n_pad]            published gradients, by step parity
    // [...]    codes   uint8[2][code_pad]         published flag codes, by segment parity
    static constexpr size_t gflags = 0, sflags = 64, cflags = 128, grid = 256, partial = 512, small = 1024, pub = 2048;
};
static_assert(kCommMaxWorld * sizeof(uint32_t) <= 64, "flag arrays are 64 bytes apart");

struct Comm {
    int world, rank, n, n_pad;
    int64_t code_pad;                    // bytes per flag-code buffer (multiple of 256)
    char* local;                         // this rank's block
    char* peer[kCommMaxWorld];           // every rank's block as mapped here (peer[rank] == local)
    bool opened[kCommMaxWorld];
    uint32_t step;                       // optimiser steps exchanged so far
    uint32_t grid_epoch;                 // grid barriers passed so far
    uint32_t small_epoch, code_epoch;
    int64_t code_bytes_published;
    size_t codes_off() const { return CommLayout::pub + 2 * (size_t)n_pad * sizeof(float); }
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct CommPtrs {
    char* peer[kCommMaxWorld];
};

__device__ __forceinline__ double comm_block_sum(double v, double* scratch) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += scratch[w];
    return t;
}

// all CTAs of the grid are co-resident (<= 36 CTAs of 1024 threads on 148 SMs); a timeout records error 2
__device__ __forceinline__ void grid_barrier(uint32_t* counter, uint32_t target, uint32_t* err) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        uint32_t spins = 0;
        while (ld_acquire_gpu(counter) < target) {
            if (++spins > kCommSpinLimit) {
                atomicExch(err, 2u);
                break;
            }
        }
    }
    __syncthreads();
}

// thread r < world: tell rank r that this rank has published `epoch`, then wait until rank r has (error 1 on timeout)
__device__ __forceinline__ void signal_and_wait(const CommPtrs& ptrs, int world, int rank, size_t flag_off,
                                                uint32_t epoch, bool signal, uint32_t* err) {
    if ((int)threadIdx.x < world) {
        if (signal) {
            __threadfence_system();
            st_release_sys(reinterpret_cast<uint32_t*>(ptrs.peer[threadIdx.x] + flag_off) + rank, epoch);
        }
        const uint32_t* mine = reinterpret_cast<const uint32_t*>(ptrs.peer[rank] + flag_off) + threadIdx.x;
        uint32_t spins = 0;
        while (ld_acquire_sys(mine) < epoch) {
            if (++spins > kCommSpinLimit) {
                atomicExch(err, 1u);
                break;
            }
        }
    }
}

__global__ void __launch_bounds__(kCommThreads) allreduce_clip_adam_kernel(
    CommPtrs ptrs, int world, int rank, int n, int n_pad, uint32_t step, uint32_t grid_target, float* __restrict__ p,
    float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, float max_norm, float lr, float b1, float b2,
    float eps, float bc1, float bc2_sqrt, float* grad_norm_out) {
    __shared__ double scratch[32];
    char* local = ptrs.peer[rank];
    uint32_t* grid = reinterpret_cast<uint32_t*>(local + CommLayout::grid);
    uint32_t* err = grid + 2;
    if (ld_acquire_gpu(err) != 0u) return;       // an earlier exchange failed: nothing may be applied any more
    double* partial = reinterpret_cast<double*>(local + CommLayout::partial);
    const int buf = (int)(step & 1u);
    float* pub = reinterpret_cast<float*>(local + CommLayout::pub) + (size_t)buf * n_pad;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;

    // 1. publish this rank's gradient, then tell every peer (after the whole grid has written)
    if (i < n) pub[i] = g[i];
    grid_barrier(grid, grid_target - gridDim.x, err);               // first of the two barriers of this launch
    // 2. block 0 signals; every block waits until every rank has published this step
    signal_and_wait(ptrs, world, rank, CommLayout::gflags, step + 1u, blockIdx.x == 0, err);
    __syncthreads();
    // 3. one-shot all-reduce of this thread's element, in rank order
    // (all peer loads are issued before the first one is used: a load over NVLink takes ~1.5 us, and a loop over a run-time
    // `world` would pay that once per rank)
    float gi = 0.0f;
    if (i < n) {
        float part[kCommMaxWorld];
#pragma unroll
        for (int r = 0; r < kCommMaxWorld; ++r) {
            part[r] = 0.0f;
            if (r < world) {
                const float* src = reinterpret_cast<const float*>(ptrs.peer[r] + CommLayout::pub) + (size_t)buf * n_pad;
                part[r] = __ldcv(src + i);
            }
        }
#pragma unroll
        for (int r = 0; r < kCommMaxWorld; ++r) gi += part[r];          // rank order: every rank gets the bitwise identical sum
    }
    const double ss = comm_block_sum((double)gi * (double)gi, scratch);
    if (threadIdx.x == 0) partial[blockIdx.x] = ss;
    grid_barrier(grid, grid_target, err);
    // a timeout recorded by ANY CTA before it arrived at the barrier above is visible here: the step is dropped as a
    // whole -- g keeps the local gradient, parameters and moments stay as they were
    if (ld_acquire_gpu(err) != 0u) return;
    double tot = 0.0;
    for (int c = 0; c < (int)gridDim.x; ++c) tot += __ldcv(partial + c);
    const float norm = (float)sqrt(tot);
    if (blockIdx.x == 0 && threadIdx.x == 0 && grad_norm_out) *grad_norm_out = norm;
    // 4. clip_grad_norm_ + Adam (same arithmetic as learner_kernels.cu::clip_adam_kernel)
    float coef = max_norm / (norm + 1e-6f);
    coef = coef > 1.0f ? 1.0f : coef;
    const float step_size = lr / bc1;
    if (i < n) {
        g[i] = gi;                       // the reduced gradient, for the caller (grad-norm records, tests)
        const float gc = gi * coef;
        const float mi = m[i] + (gc - m[i]) * (1.0f - b1);
        const float vi = v[i] * b2 + (1.0f - b2) * gc * gc;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = p[i] - step_size * (mi / denom);
    }
}

// values[count] <- sum over the ranks, in rank order (bitwise identical everywhere).  One CTA.
__global__ void __launch_bounds__(64) comm_small_allreduce_kernel(CommPtrs ptrs, int world, int rank, uint32_t epoch,
                                                                  double* __restrict__ values, int count) {
    char* local = ptrs.peer[rank];
    uint32_t* err = reinterpret_cast<uint32_t*>(local + CommLayout::grid) + 2;
    if (ld_acquire_gpu(err) != 0u) return;
    const int slot = (int)(epoch & 1u) * kCommSmallMax;
    double* mine = reinterpret_cast<double*>(local + CommLayout::small) + slot;
    const int tid = threadIdx.x;
    if (tid < count) mine[tid] = values[tid];
    __syncthreads();
    signal_and_wait(ptrs, world, rank, CommLayout::sflags, epoch, true, err);
    __syncthreads();
    if (ld_acquire_gpu(err) != 0u) return;
    if (tid < count) {
        double t = 0.0;
        for (int r = 0; r < world; ++r)
            t += __ldcv(reinterpret_cast<const double*>(ptrs.peer[r] + CommLayout::small) + slot + tid);
        values[tid] = t;
    }
}

// copies this rank's flag codes into its mapped block; the LAST block to finish tells the peers
__global__ void __launch_bounds__(256) comm_publish_codes_kernel(CommPtrs ptrs, int world, int rank, uint32_t epoch,
                                                                 const uint8_t* __restrict__ code, long long bytes,
                                                                 size_t codes_off, long long code_pad) {
    __shared__ int is_last;
    char* local = ptrs.peer[rank];
    uint32_t* grid = reinterpret_cast<uint32_t*>(local + CommLayout::grid);
    if (ld_acquire_gpu(grid + 2) != 0u) return;
    uint8_t* dst = reinterpret_cast<uint8_t*>(local + codes_off) + (size_t)(epoch & 1u) * code_pad;
    const long long vec = bytes / 16;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < vec; i += stride)
        reinterpret_cast<uint4*>(dst)[i] = __ldg(reinterpret_cast<const uint4*>(code) + i);
    for (long long i = vec * 16 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < bytes; i += stride)
        dst[i] = code[i];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t prev = atomicAdd(grid + 1, 1u);
        is_last = prev == gridDim.x - 1;
        if (is_last) {
            __threadfence();
            atomicExch(grid + 1, 0u);
        }
    }
    __syncthreads();
    if (is_last && (int)threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<uint32_t*>(ptrs.peer[threadIdx.x] + CommLayout::cflags) + rank, epoch);
    }
}

__global__ void __launch_bounds__(32) comm_wait_codes_kernel(CommPtrs ptrs, int world, int rank, uint32_t epoch) {
    uint32_t* err = reinterpret_cast<uint32_t*>(ptrs.peer[rank] + CommLayout::grid) + 2;
    if (ld_acquire_gpu(err) != 0u) return;
    signal_and_wait(ptrs, world, rank, CommLayout::cflags, epoch, false, err);
}

int comm_code_sources(void* comm, int64_t bytes, CodeSrc* out, int* world, int* rank) {
    Comm* c = static_cast<Comm*>(comm);
    if (c->code_epoch == 0 || c->code_bytes_published != bytes)
        return fail("the communicator has not published this segment's flag codes (%lld B wanted, %lld B published)",
                    (long long)bytes, (long long)c->code_bytes_published);
    for (int r = 0; r < kCommMaxWorld; ++r)
        out->base[r] = r < c->world ? reinterpret_cast<const uint8_t*>(c->peer[r] + c->codes_off()) +
                                          (size_t)(c->code_epoch & 1u) * c->code_pad
                                    : nullptr;
    *world = c->world;
    *rank = c->rank;
    return 0;
}

static CommPtrs comm_ptrs(const Comm* c) {
    CommPtrs ptrs;
    for (int r = 0; r < kCommMaxWorld; ++r) ptrs.peer[r] = c->peer[r];
    return ptrs;
}

}  // namespace plume

using namespace plume;

extern "C" int plume_comm_create(int32_t world, int32_t rank, int32_t n_params, int64_t code_bytes, void** comm_out,
                                 uint8_t* handle_out /* 64 bytes */) {
    PLUME_CHECK_ARG(comm_out && handle_out, "null pointer");
    PLUME_CHECK_ARG(world >= 1 && world <= kCommMaxWorld && rank >= 0 && rank < world, "bad world/rank");
    PLUME_CHECK_ARG(n_params > 0 && code_bytes >= 0, "n_params must be positive, code_bytes non-negative");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    Comm* c = new Comm();
    c->world = world;
    c->rank = rank;
    c->n = n_params;
    c->n_pad = (n_params + 255) / 256 * 256;
    c->code_pad = (code_bytes + 255) / 256 * 256;
    c->step = 0;
    c->grid_epoch = 0;
    c->small_epoch = 0;
    c->code_epoch = 0;
    c->code_bytes_published = 0;
    for (int r = 0; r < kCommMaxWorld; ++r) {
        c->peer[r] = nullptr;
        c->opened[r] = false;
    }
    const size_t bytes = c->codes_off() + 2 * (size_t)c->code_pad;
    if (cudaMalloc(&c->local, bytes) != cudaSuccess) {
        delete c;
        return fail("plume_comm_create: cudaMalloc of %zu B failed", bytes);
    }
    cudaMemset(c->local, 0, bytes);
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, c->local) != cudaSuccess) {
        cudaFree(c->local);
        delete c;
        return fail("plume_comm_create: cudaIpcGetMemHandle failed (%s)", cudaGetErrorString(cudaGetLastError()));
    }
    memcpy(handle_out, &h, 64);
    c->peer[rank] = c->local;
    *comm_out = c;
    return 0;
}

extern "C" int plume_comm_connect(void* comm, const uint8_t* all_handles /* [world][64] */) {
    PLUME_CHECK_ARG(comm && all_handles, "null pointer");
    Comm* c = static_cast<Comm*>(comm);
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, all_handles + 64 * r, 64);
        void* ptr = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail("plume_comm_connect: cannot map rank %d's buffer (%s)", r, cudaGetErrorString(e));
        c->peer[r] = static_cast<char*>(ptr);
        c->opened[r] = true;
    }
    return 0;
}

extern "C" int plume_comm_destroy(void* comm) {
    if (!comm) return 0;
    Comm* c = static_cast<Comm*>(comm);
    for (int r = 0; r < c->world; ++r)
        if (c->opened[r]) cudaIpcCloseMemHandle(c->peer[r]);
    cudaFree(c->local);
    delete c;
    return 0;
}

// HOST out: 0 = fine, 1 = a peer did not publish in time, 2 = local grid barrier timed out (synchronises the stream)
extern "C" int plume_comm_error(void* comm, int32_t* error_out, void* stream) {
    PLUME_CHECK_ARG(comm && error_out, "null pointer");
    Comm* c = static_cast<Comm*>(comm);
    uint32_t e = 0;
    PLUME_CUDA(cudaMemcpyAsync(&e, c->local + CommLayout::grid + 2 * sizeof(uint32_t), sizeof(uint32_t),
                               cudaMemcpyDeviceToHost, as_stream(stream)));
    PLUME_CUDA(cudaStreamSynchronize(as_stream(stream)));
    *error_out = (int32_t)e;
    return 0;
}

// queues a copy of the error word into pinned host memory behind the work already in `stream` (no synchronisation)
extern "C" int plume_comm_error_async(void* comm, int32_t* pinned_error_out, void* stream) {
    PLUME_CHECK_ARG(comm && pinned_error_out, "null pointer");
    Comm* c = static_cast<Comm*>(comm);
    PLUME_CUDA(cudaMemcpyAsync(pinned_error_out, c->local + CommLayout::grid + 2 * sizeof(uint32_t), sizeof(uint32_t),
                               cudaMemcpyDeviceToHost, as_stream(stream)));
    return 0;
}

// Collective: every rank calls it between two host barriers, with no exchange kernel in flight anywhere.
extern "C" int plume_comm_reset(void* comm) {
    PLUME_CHECK_ARG(comm, "null pointer");
    Comm* c = static_cast<Comm*>(comm);
    PLUME_CUDA(cudaDeviceSynchronize());
    PLUME_CUDA(cudaMemset(c->local, 0, CommLayout::partial));       // flag arrays, counters, error word
    PLUME_CUDA(cudaDeviceSynchronize());
    c->step = c->grid_epoch = c->small_epoch = c->code_epoch = 0;
    return 0;
}

extern "C" int plume_allreduce_clip_adam(void* comm, float* params, float* grads, float* exp_avg, float* exp_avg_sq,
                                         int32_t n, float max_norm, float lr, float beta1, float beta2, float eps,
                                         int32_t step, float* grad_norm_out, void* stream) {
    PLUME_CHECK_ARG(comm && params && grads && exp_avg && exp_avg_sq, "null pointer");
    PLUME_CHECK_ARG(step >= 1, "Adam step is 1-based");
    Comm* c = static_cast<Comm*>(comm);
    PLUME_CHECK_ARG(n == c->n, "parameter count differs from the communicator's");
    for (int r = 0; r < c->world; ++r) PLUME_CHECK_ARG(c->peer[r] != nullptr, "communicator is not connected");
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    const int blocks = (n + kCommThreads - 1) / kCommThreads;
    c->grid_epoch += 2;                                         // two grid barriers per launch
    allreduce_clip_adam_kernel<<<blocks, kCommThreads, 0, as_stream(stream)>>>(
        comm_ptrs(c), c->world, c->rank, n, c->n_pad, c->step, c->grid_epoch * (uint32_t)blocks, params, grads, exp_avg,
        exp_avg_sq, max_norm, lr, beta1, beta2, eps, (float)bc1, (float)sqrt(bc2), grad_norm_out);
    PLUME_LAUNCH_CHECK();
    c->step += 1;
    return 0;
}

extern "C" int plume_comm_allreduce_small(void* comm, double* values, int32_t count, void* stream) {
    PLUME_CHECK_ARG(comm && values, "null pointer");
    PLUME_CHECK_ARG(count >= 1 && count <= kCommSmallMax, "count must be in [1, 32]");
    Comm* c = static_cast<Comm*>(comm);
    for (int r = 0; r < c->world; ++r) PLUME_CHECK_ARG(c->peer[r] != nullptr, "communicator is not connected");
    c->small_epoch += 1;
    comm_small_allreduce_kernel<<<1, 64, 0, as_stream(stream)>>>(comm_ptrs(c), c->world, c->rank, c->small_epoch, values,
                                                                 count);
    PLUME_LAUNCH_CHECK();
    return 0;
}

extern "C" int plume_comm_publish_codes(void* comm, const uint8_t* flag_code, int64_t bytes, void* stream) {
    PLUME_CHECK_ARG(comm && flag_code, "null pointer");
    Comm* c = static_cast<Comm*>(comm);
    PLUME_CHECK_ARG(bytes > 0 && bytes <= c->code_pad, "segment larger than the communicator's code buffers");
    PLUME_CHECK_ARG((reinterpret_cast<uintptr_t>(flag_code) & 15) == 0, "flag_code must be 16-byte aligned");
    for (int r = 0; r < c->world; ++r) PLUME_CHECK_ARG(c->peer[r] != nullptr, "communicator is not connected");
    c->code_epoch += 1;
    c->code_bytes_published = bytes;
    long long blocks = (bytes / 16 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 128) blocks = 128;
    comm_publish_codes_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
        comm_ptrs(c), c->world, c->rank, c->code_epoch, flag_code, (long long)bytes, c->codes_off(),
        (long long)c->code_pad);
    PLUME_LAUNCH_CHECK();
    return 0;
}

extern "C" int plume_curriculum_update_peer(void* comm, int32_t horizon, int32_t n_envs, double* state,
                                            double* curriculum, double initial_radius, double min_radius,
                                            double radius_decay, double success_threshold, int32_t window,
                                            double decay_factor, double* window_radius_out, void* stream) {
    PLUME_CHECK_ARG(comm && state && curriculum, "null pointer");
    Comm* c = static_cast<Comm*>(comm);
    PLUME_CHECK_ARG(c->code_epoch > 0 && c->code_bytes_published == (int64_t)horizon * n_envs,
                    "publish this segment's flag codes first (plume_comm_publish_codes)");
    comm_wait_codes_kernel<<<1, 32, 0, as_stream(stream)>>>(comm_ptrs(c), c->world, c->rank, c->code_epoch);
    PLUME_LAUNCH_CHECK();
    CodeSrc src;
    int world = 0, rank = 0;
    if (comm_code_sources(comm, (int64_t)horizon * n_envs, &src, &world, &rank) != 0) return -1;
    return launch_curriculum_packed(src, horizon, n_envs, c->world, state, curriculum, initial_radius, min_radius,
                                    radius_decay, success_threshold, window, decay_factor, window_radius_out,
                                    reinterpret_cast<const uint32_t*>(c->local + CommLayout::grid) + 2,
                                    as_stream(stream));
}
