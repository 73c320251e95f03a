// comm_kernels.cu -- the one exchange step of the PPO update as ONE kernel over NVLink peer memory:
// gradient all-reduce + clip_grad_norm_ + Adam (train_ppo2.0.py:85-87 on N data-parallel ranks).
//
// Every rank publishes its 145 KB minibatch gradient in a buffer that all peers have mapped (CUDA IPC,
// one process per GPU), signals "published" with a system-scope release store into each peer's flag
// array, waits for the flags of all ranks, and then reduces its own copy: CTA c sums slice c of all
// ranks' buffers in rank order (so every rank obtains the bitwise identical sum -- a one-shot all-reduce,
// 7/8 of the 1.2 MB read over NVLink), the CTAs combine their slice norms through a grid barrier, and each
// thread finishes clip + Adam for its element from registers.  Buffers alternate by step parity, so one
// cross-GPU barrier per step is enough: a rank can only publish step k+2 into the buffer of step k after
// every peer has arrived at step k+1, i.e. finished reading step k.
//
// Replaces, per optimiser step, an NCCL all-reduce (latency bound: ~35/70/180 us at 2/4/8 GPUs when
// launched from torch.distributed) plus the clip+Adam kernel.  Algorithmic bytes per step and rank:
// world x 145 KB read (peer), 145 KB published, 16 B read + 12 B written per parameter.
#include <cstring>

#include "common.cuh"

namespace plume {

constexpr int kCommMaxWorld = 16;
constexpr int kCommThreads = 1024;

struct CommLayout {            // one cudaMalloc'd block per rank, mapped by every peer
    // [0]                     flags   uint32[kCommMaxWorld]   flags[r] = last step rank r has published
    // [256]                   grid    uint32[4]               local grid-barrier counter, error flag
    // [512]                   partial double[64]              slice norms (local)
    // [1024]                  pub     float[2][n_pad]         published gradients, by step parity
    static constexpr size_t flags = 0, grid = 256, partial = 512, pub = 1024;
};

struct Comm {
    int world, rank, n, n_pad;
    char* local;                         // this rank's block
    char* peer[kCommMaxWorld];           // every rank's block as mapped here (peer[rank] == local)
    bool opened[kCommMaxWorld];
    uint32_t step;                       // optimiser steps exchanged so far
    uint32_t grid_epoch;                 // grid barriers passed so far
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

// all CTAs of the grid are co-resident (<= 36 CTAs of 1024 threads on 148 SMs)
__device__ __forceinline__ bool grid_barrier(uint32_t* counter, uint32_t target, uint32_t* err) {
    __syncthreads();
    bool ok = true;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        uint32_t spins = 0;
        while (ld_acquire_gpu(counter) < target) {
            if (++spins > (1u << 26)) {
                atomicExch(err, 2u);
                ok = false;
                break;
            }
        }
    }
    __syncthreads();
    return ok;
}

__global__ void __launch_bounds__(kCommThreads) allreduce_clip_adam_kernel(
    CommPtrs ptrs, int world, int rank, int n, int n_pad, uint32_t step, uint32_t grid_target, float* __restrict__ p,
    float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, float max_norm, float lr, float b1, float b2,
    float eps, float bc1, float bc2_sqrt, float* grad_norm_out) {
    __shared__ double scratch[32];
    char* local = ptrs.peer[rank];
    uint32_t* flags = reinterpret_cast<uint32_t*>(local + CommLayout::flags);
    uint32_t* grid = reinterpret_cast<uint32_t*>(local + CommLayout::grid);
    double* partial = reinterpret_cast<double*>(local + CommLayout::partial);
    const int buf = (int)(step & 1u);
    float* pub = reinterpret_cast<float*>(local + CommLayout::pub) + (size_t)buf * n_pad;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;

    // 1. publish this rank's gradient, then tell every peer (after the whole grid has written)
    if (i < n) pub[i] = g[i];
    grid_barrier(grid, grid_target - gridDim.x, grid + 2);          // first of the two barriers of this launch
    if (blockIdx.x == 0 && threadIdx.x < world) {
        __threadfence_system();
        uint32_t* peer_flags = reinterpret_cast<uint32_t*>(ptrs.peer[threadIdx.x] + CommLayout::flags);
        st_release_sys(peer_flags + rank, step + 1u);
    }
    // 2. wait until every rank has published this step
    if (threadIdx.x < world) {
        uint32_t spins = 0;
        while (ld_acquire_sys(flags + threadIdx.x) < step + 1u) {
            if (++spins > (1u << 26)) {
                atomicExch(grid + 2, 1u);
                break;
            }
        }
    }
    __syncthreads();
    // 3. one-shot all-reduce of this thread's element, in rank order
    float gi = 0.0f;
    if (i < n) {
        for (int r = 0; r < world; ++r) {
            const float* src = reinterpret_cast<const float*>(ptrs.peer[r] + CommLayout::pub) + (size_t)buf * n_pad;
            gi += __ldcv(src + i);
        }
        g[i] = gi;                       // the reduced gradient, for the caller (grad-norm records, tests)
    }
    const double ss = comm_block_sum((double)gi * (double)gi, scratch);
    if (threadIdx.x == 0) partial[blockIdx.x] = ss;
    grid_barrier(grid, grid_target, grid + 2);
    double tot = 0.0;
    for (int c = 0; c < (int)gridDim.x; ++c) tot += __ldcv(partial + c);
    const float norm = (float)sqrt(tot);
    if (blockIdx.x == 0 && threadIdx.x == 0 && grad_norm_out) *grad_norm_out = norm;
    // 4. clip_grad_norm_ + Adam (same arithmetic as learner_kernels.cu::clip_adam_kernel)
    float coef = max_norm / (norm + 1e-6f);
    coef = coef > 1.0f ? 1.0f : coef;
    const float step_size = lr / bc1;
    if (i < n) {
        const float gc = gi * coef;
        const float mi = m[i] + (gc - m[i]) * (1.0f - b1);
        const float vi = v[i] * b2 + (1.0f - b2) * gc * gc;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = p[i] - step_size * (mi / denom);
    }
}

}  // namespace plume

using namespace plume;

extern "C" int plume_comm_create(int32_t world, int32_t rank, int32_t n_params, void** comm_out,
                                 uint8_t* handle_out /* 64 bytes */) {
    PLUME_CHECK_ARG(comm_out && handle_out, "null pointer");
    PLUME_CHECK_ARG(world >= 1 && world <= kCommMaxWorld && rank >= 0 && rank < world, "bad world/rank");
    PLUME_CHECK_ARG(n_params > 0, "n_params must be positive");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    Comm* c = new Comm();
    c->world = world;
    c->rank = rank;
    c->n = n_params;
    c->n_pad = (n_params + 255) / 256 * 256;
    c->step = 0;
    c->grid_epoch = 0;
    for (int r = 0; r < kCommMaxWorld; ++r) {
        c->peer[r] = nullptr;
        c->opened[r] = false;
    }
    const size_t bytes = CommLayout::pub + 2 * (size_t)c->n_pad * sizeof(float);
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
    CommPtrs ptrs;
    for (int r = 0; r < kCommMaxWorld; ++r) ptrs.peer[r] = c->peer[r];
    const int blocks = (n + kCommThreads - 1) / kCommThreads;
    c->grid_epoch += 2;                                         // two grid barriers per launch
    allreduce_clip_adam_kernel<<<blocks, kCommThreads, 0, as_stream(stream)>>>(
        ptrs, c->world, c->rank, n, c->n_pad, c->step, c->grid_epoch * (uint32_t)blocks, params, grads, exp_avg,
        exp_avg_sq, max_norm, lr, beta1, beta2, eps, (float)bc1, (float)sqrt(bc2), grad_norm_out);
    PLUME_LAUNCH_CHECK();
    c->step += 1;
    return 0;
}
