// trajectory_kernels.cu -- N2: per-episode trajectory / statistics logging in the layouts of the reference's
// training_data.nc (NetCDFWriter.write_episode_data, PPOV2.1/model.py:351-419) and training_results*.csv
// (train_ppo2.0.py:128-134,140-180,194-199,236-248), assembled on the device from the [T][N] rollout buffers.
//
// The reference appends (x, y, conc) per step to python lists and writes one row per finished episode; here a
// segment of T lockstep steps of N envs is scattered in five small launches:
//   A  rank      one CTA per step t: position of every `done` among the dones of its row (ballot scan), row
//                counts; with several ranks also the dones of the other ranks' rows (global canonical order)
//   B  scan      exclusive scans over the T rows -> episode slot base / global episode ordinal base per row
//   C1 walk      one thread per env, forward over t: step index inside the episode, float64 running sums (time
//                order, like the reference's `+=`), per-episode scalars at every done; backward over t: which episode
//                row (or the env's carry row) each transition belongs to.  All [T][N] accesses coalesced.
//   C3 migrate   one CTA per env whose open episode from earlier segments closes here: carry row -> episode row
//   C2 scatter   one thread per transition: x / y / concentration into its episode row or the carry row
// Canonical episode order = step-major, then (global) env id -- the order the curriculum replays.
//
// Per-episode scalars (train_ppo2.0.py): Steps = number of steps; Final_Conc = conc_field at the final cell if the
// source was reached, else 0.0 (:145,194-196,246); Current_Radius = the TRAINER's radius when the episode ended,
// before its own curriculum update (:247,251) -- taken from the window radii the curriculum kernel reports for this
// segment (plume_curriculum_update_packed: window_radius_out).
#include "common.cuh"
#include "curriculum.cuh"

namespace plume {

struct TrajWs {                  // workspace layout (int32 units unless noted)
    int* rank;                   // [T][N] position of a done among the dones of its row (valid where done)
    int* next_slot;              // [T][N] episode row of the transition; -1 = carry (still open), -2 = dropped
    uint16_t* kidx;              // [T][N] step index inside the episode
    int* row_cnt;                // [T] local dones per row
    int* row_before;             // [T] dones of lower ranks in the same row
    int* row_all;                // [T] dones of all ranks per row
    int* row_base;               // [T] first episode slot of the row (count0 + exclusive local scan)
    int* ord_base;               // [T] global ordinal of the row's first local episode
    int* first_slot;             // [N] slot of the env's first done in this segment (-1 none, -2 dropped)
    int* len_in;                 // [N] carried length at entry
};

__host__ __device__ inline size_t traj_align(size_t x) { return (x + 255) / 256 * 256; }

static size_t traj_ws_bytes(int T, int N) {
    const size_t tn = (size_t)T * N;
    return traj_align(tn * 4) * 2 + traj_align(tn * 2) + traj_align((size_t)T * 4) * 5 + traj_align((size_t)N * 4) * 2 + 256;
}

static TrajWs traj_ws_carve(void* ws, int T, int N) {
    char* p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
    const size_t tn = (size_t)T * N;
    TrajWs w;
    w.rank = reinterpret_cast<int*>(p);            p += traj_align(tn * 4);
    w.next_slot = reinterpret_cast<int*>(p);       p += traj_align(tn * 4);
    w.kidx = reinterpret_cast<uint16_t*>(p);       p += traj_align(tn * 2);
    w.row_cnt = reinterpret_cast<int*>(p);         p += traj_align((size_t)T * 4);
    w.row_before = reinterpret_cast<int*>(p);      p += traj_align((size_t)T * 4);
    w.row_all = reinterpret_cast<int*>(p);         p += traj_align((size_t)T * 4);
    w.row_base = reinterpret_cast<int*>(p);        p += traj_align((size_t)T * 4);
    w.ord_base = reinterpret_cast<int*>(p);        p += traj_align((size_t)T * 4);
    w.first_slot = reinterpret_cast<int*>(p);      p += traj_align((size_t)N * 4);
    w.len_in = reinterpret_cast<int*>(p);
    return w;
}

// ---- A: rank of every done inside its row -------------------------------------------------------------------
__global__ void __launch_bounds__(256) traj_rank_kernel(const float* __restrict__ dones, int N, CodeSrc peers, int world,
                                                        int my_rank, TrajWs w) {
    __shared__ int warp_cnt[8];
    __shared__ int red[8];
    const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* row = dones + (size_t)t * N;
    int running = 0;
    for (int n0 = 0; n0 < N; n0 += 256) {
        const int n = n0 + tid;
        const bool d = n < N && row[n] != 0.0f;
        const unsigned m = __ballot_sync(0xffffffffu, d);
        if (lane == 0) warp_cnt[warp] = __popc(m);
        __syncthreads();
        int before = running, total = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (q < warp) before += warp_cnt[q];
            total += warp_cnt[q];
        }
        if (d) w.rank[(size_t)t * N + n] = before + __popc(m & ((1u << lane) - 1u));
        running += total;
        __syncthreads();
    }
    // the other ranks' dones of this row (their published flag codes): only counts are needed
    int before_me = 0, others = 0;
    for (int r = 0; r < world; ++r) {
        if (r == my_rank) continue;
        const uint8_t* code = peers.base[r] + (size_t)t * N;
        int c = 0;
        for (int n = tid; n < N; n += 256) c += code[n] & 1;
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) red[warp] = c;
        __syncthreads();
        int tot = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) tot += red[q];
        __syncthreads();
        others += tot;
        if (r < my_rank) before_me += tot;
    }
    if (tid == 0) {
        w.row_cnt[t] = running;
        w.row_before[t] = before_me;
        w.row_all[t] = running + others;
    }
}

// ---- B: exclusive scans over the rows (T <= 1024) ------------------------------------------------------------
__global__ void __launch_bounds__(1024) traj_scan_kernel(int T, TrajWs w, int* __restrict__ count, int max_episodes) {
    __shared__ int s_loc[1024], s_all[1024];
    const int tid = threadIdx.x;
    s_loc[tid] = tid < T ? w.row_cnt[tid] : 0;
    s_all[tid] = tid < T ? w.row_all[tid] : 0;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {          // inclusive Hillis-Steele scans
        const int a = tid >= off ? s_loc[tid - off] : 0, b = tid >= off ? s_all[tid - off] : 0;
        __syncthreads();
        s_loc[tid] += a;
        s_all[tid] += b;
        __syncthreads();
    }
    const int count0 = *count;
    if (tid < T) {
        w.row_base[tid] = count0 + s_loc[tid] - w.row_cnt[tid];
        w.ord_base[tid] = s_all[tid] - w.row_all[tid] + w.row_before[tid];
    }
    __syncthreads();
    if (tid == 0) {
        const long long c = (long long)count0 + s_loc[1023];
        *count = c < max_episodes ? (int)c : max_episodes;
    }
}

// ---- C1: one thread per env ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) traj_walk_kernel(plume_traj_log lg, plume_rollout_buffers buf, int T, TrajWs w,
                                                        int window, const double* __restrict__ window_radius,
                                                        double radius_fallback) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int N = lg.n_envs, S = lg.max_steps, E = lg.max_episodes;
    if (n >= N) return;
    int len = lg.c_len[n];
    w.len_in[n] = len;
    double sums[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) sums[k] = lg.c_sums[(size_t)n * 6 + k];
    int first = -1;
    const long long hist_len = window_radius ? (long long)window_radius[0] : 0;
    const int n_radii = window_radius ? (int)window_radius[1] : 0;
    for (int t = 0; t < T; ++t) {
        const size_t i = (size_t)t * N + n;
        w.kidx[i] = (uint16_t)(len < S ? len : S - 1);
        sums[0] += (double)buf.rewards[i];
        if (buf.info) {
            const float* inf = buf.info + (size_t)t * 5 * N + n;
#pragma unroll
            for (int k = 0; k < 5; ++k) sums[1 + k] += (double)inf[(size_t)k * N];
        }
        if (buf.dones[i] != 0.0f) {
            const int rk = w.rank[i];
            const int slot = w.row_base[t] + rk;
            if (slot < E) {
                lg.steps[slot] = len + 1 < S ? len + 1 : S;
                const bool ok = buf.reached[i] != 0;
                lg.success[slot] = ok ? 1 : 0;
                if (buf.src_out) {
                    lg.source[(size_t)slot * 2] = buf.src_out[i * 2];
                    lg.source[(size_t)slot * 2 + 1] = buf.src_out[i * 2 + 1];
                }
                lg.final_conc[slot] = (ok && buf.conc_out) ? buf.conc_out[i] : 0.0f;      // train_ppo2.0.py:145,194-196
                double rad = radius_fallback;
                if (window_radius && n_radii > 0) {
                    long long b = (hist_len + (long long)w.ord_base[t] + rk) / window;
                    if (b >= n_radii) b = n_radii - 1;
                    rad = window_radius[2 + b];
                }
                lg.radius[slot] = rad;                                                    // :247
#pragma unroll
                for (int k = 0; k < 6; ++k) lg.sums[(size_t)slot * 6 + k] = sums[k];
            }
            if (first == -1) first = slot < E ? slot : -2;
            len = 0;
#pragma unroll
            for (int k = 0; k < 6; ++k) sums[k] = 0.0;
        } else {
            ++len;
        }
    }
    lg.c_len[n] = len;
#pragma unroll
    for (int k = 0; k < 6; ++k) lg.c_sums[(size_t)n * 6 + k] = sums[k];
    w.first_slot[n] = first;
    int cur = -1;
    for (int t = T - 1; t >= 0; --t) {
        const size_t i = (size_t)t * N + n;
        if (buf.dones[i] != 0.0f) {
            const int slot = w.row_base[t] + w.rank[i];
            cur = slot < E ? slot : -2;
        }
        w.next_slot[i] = cur;
    }
}

// ---- C3: carried prefix of an episode that closes in this segment -> its episode row -----------------------------
__global__ void __launch_bounds__(128) traj_migrate_kernel(plume_traj_log lg, TrajWs w) {
    const int n = blockIdx.x;
    const int slot = w.first_slot[n], len = w.len_in[n], S = lg.max_steps;
    if (slot < 0 || len <= 0) return;
    const size_t src = (size_t)n * S, dst = (size_t)slot * S;
    for (int k = threadIdx.x; k < len && k < S; k += blockDim.x) {
        lg.x[dst + k] = lg.c_x[src + k];
        lg.y[dst + k] = lg.c_y[src + k];
        lg.conc[dst + k] = lg.c_conc[src + k];
    }
}

// ---- C2: one thread per transition -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) traj_scatter_kernel(plume_traj_log lg, plume_rollout_buffers buf, int T, TrajWs w) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int N = lg.n_envs, S = lg.max_steps;
    if (i >= (long long)T * N) return;
    const int slot = w.next_slot[i];
    if (slot == -2) return;
    const int n = (int)(i % N), k = w.kidx[i];
    const float2 p = reinterpret_cast<const float2*>(buf.pos_out)[i];
    const float cc = buf.conc_out[i];
    if (slot >= 0) {
        const size_t d = (size_t)slot * S + k;
        lg.x[d] = p.x;
        lg.y[d] = p.y;
        lg.conc[d] = cc;
    } else {
        const size_t d = (size_t)n * S + k;
        lg.c_x[d] = p.x;
        lg.c_y[d] = p.y;
        lg.c_conc[d] = cc;
    }
}

}  // namespace plume

using namespace plume;

extern "C" int64_t plume_trajectory_workspace_bytes(int32_t horizon, int32_t n_envs) {
    if (horizon <= 0 || n_envs <= 0) return 256;
    return (int64_t)traj_ws_bytes(horizon, n_envs);
}

extern "C" int plume_trajectory_log(const plume_traj_log* log, const plume_rollout_buffers* buf, int32_t horizon,
                                    void* comm, int32_t window, const double* window_radius, double radius_fallback,
                                    void* workspace, int64_t workspace_bytes, void* stream) {
    PLUME_CHECK_ARG(log && buf && workspace, "null pointer");
    PLUME_CHECK_ARG(log->x && log->y && log->conc && log->steps && log->source && log->success && log->radius &&
                        log->sums && log->final_conc && log->c_x && log->c_y && log->c_conc && log->c_sums &&
                        log->c_len && log->count, "null log table");
    PLUME_CHECK_ARG(buf->dones && buf->reached && buf->rewards && buf->pos_out && buf->conc_out,
                    "the logger needs dones, reached, rewards, pos_out and conc_out (RolloutEngine(with_trajectory=True))");
    PLUME_CHECK_ARG(log->max_episodes > 0 && log->max_steps > 0 && log->max_steps <= 65535 && log->n_envs > 0,
                    "bad table sizes");
    PLUME_CHECK_ARG(horizon <= 1024, "at most 1024 steps per segment");
    PLUME_CHECK_ARG(window > 0, "window must be positive");
    if (horizon <= 0) return 0;
    const int T = horizon, N = log->n_envs;
    PLUME_CHECK_ARG(workspace_bytes >= (int64_t)traj_ws_bytes(T, N), "workspace too small");
    CodeSrc peers;
    for (int r = 0; r < kCommMaxWorld; ++r) peers.base[r] = nullptr;
    int world = 1, my_rank = 0;
    if (comm) {
        if (comm_code_sources(comm, (int64_t)T * N, &peers, &world, &my_rank) != 0) return -1;
    }
    const TrajWs w = traj_ws_carve(workspace, T, N);
    cudaStream_t s = as_stream(stream);
    traj_rank_kernel<<<T, 256, 0, s>>>(buf->dones, N, peers, world, my_rank, w);
    traj_scan_kernel<<<1, 1024, 0, s>>>(T, w, log->count, log->max_episodes);
    traj_walk_kernel<<<(N + 127) / 128, 128, 0, s>>>(*log, *buf, T, w, window, window_radius, radius_fallback);
    traj_migrate_kernel<<<N, 128, 0, s>>>(*log, w);
    const long long tn = (long long)T * N;
    traj_scatter_kernel<<<(unsigned)((tn + 255) / 256), 256, 0, s>>>(*log, *buf, T, w);
    PLUME_LAUNCH_CHECK();
    return 0;
}
