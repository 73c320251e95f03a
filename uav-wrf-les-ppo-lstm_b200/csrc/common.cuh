// common.cuh -- error plumbing and launch helpers of libplume_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <string>

#include "plume_core.h"

namespace plume {

std::string& last_error_ref();
int fail(const char* fmt, ...);

#define PLUME_CHECK_ARG(cond, msg)                                  \
    do {                                                            \
        if (!(cond)) return ::plume::fail("%s: %s", __func__, msg); \
    } while (0)

#define PLUME_CUDA(expr)                                                                          \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            return ::plume::fail("%s: CUDA error %s at %s:%d", __func__, cudaGetErrorString(_e), \
                                 __FILE__, __LINE__);                                             \
    } while (0)

#define PLUME_LAUNCH_CHECK() PLUME_CUDA(cudaGetLastError())

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();   // cached multiprocessor count of the current device

// loads / stores of the SoA env state ------------------------------------------------------
__device__ __forceinline__ EnvRegs load_env(const plume_env_state& st, int i) {
    EnvRegs e;
    e.px = st.pos_x[i];
    e.py = st.pos_y[i];
    e.sx = st.src_x[i];
    e.sy = st.src_y[i];
    e.step = st.step_count[i];
    e.episode = (uint32_t)st.episode_idx[i];
    e.radius = st.radius[i];
    e.ebonus = st.explore_bonus[i];
    e.last_move = st.last_move ? (int32_t)st.last_move[i] : 0;
    return e;
}

__device__ __forceinline__ void store_env(const plume_env_state& st, int i, const EnvRegs& e) {
    st.pos_x[i] = e.px;
    st.pos_y[i] = e.py;
    st.src_x[i] = e.sx;
    st.src_y[i] = e.sy;
    st.step_count[i] = e.step;
    st.episode_idx[i] = (int32_t)e.episode;
    st.radius[i] = e.radius;
    st.explore_bonus[i] = e.ebonus;
    if (st.last_move) st.last_move[i] = (int8_t)e.last_move;
}

}  // namespace plume
