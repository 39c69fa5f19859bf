// Shared host/device helpers for the buffalo_b200 CUDA backend (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/buffalo_b200.h"

namespace bfl {

// ---------------------------------------------------------------------------------------
// error plumbing (replaces CHECK_CUDA's throw, include/buffalo/cuda/utils.cuh:24-31)
// ---------------------------------------------------------------------------------------
void set_error(const std::string& msg);
extern std::atomic<long long> g_launches;

#define BFL_CUDA(expr)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            bfl::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +       \
                           __FILE__ + ":" + std::to_string(__LINE__) + ")");                 \
            return BFL_ERR_CUDA;                                                             \
        }                                                                                    \
    } while (0)

#define BFL_LAUNCHED()                                                                       \
    do {                                                                                     \
        bfl::g_launches.fetch_add(1, std::memory_order_relaxed);                             \
        BFL_CUDA(cudaGetLastError());                                                        \
    } while (0)

#define BFL_FAIL(code, msg)                                                                  \
    do {                                                                                     \
        bfl::set_error(msg);                                                                 \
        return (code);                                                                       \
    } while (0)

int require_device();  // BFL_OK iff a CUDA device of compute capability 10.x is current

// ---------------------------------------------------------------------------------------
// minimal JSON reader for the option file (the reference uses json11, lib/algo.cc:19-37).
// Flat access to top-level scalars; nested objects/arrays are skipped.
// ---------------------------------------------------------------------------------------
struct JsonOpt {
    std::map<std::string, double> num;
    std::map<std::string, bool> boolean;
    std::map<std::string, std::string> str;
    bool parse(const std::string& text, std::string* err);
    bool load(const char* path, std::string* err);
    double number(const char* k, double dflt) const {
        auto it = num.find(k);
        if (it != num.end()) return it->second;
        auto ib = boolean.find(k);
        if (ib != boolean.end()) return ib->second ? 1.0 : 0.0;
        return dflt;
    }
    int integer(const char* k, int dflt) const { return (int)number(k, (double)dflt); }
    bool flag(const char* k, bool dflt) const {
        auto ib = boolean.find(k);
        if (ib != boolean.end()) return ib->second;
        auto it = num.find(k);
        if (it != num.end()) return it->second != 0.0;
        return dflt;
    }
    std::string string(const char* k, const char* dflt) const {
        auto it = str.find(k);
        return it != str.end() ? it->second : std::string(dflt);
    }
};

// device buffer with explicit ownership
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    int reserve(size_t n) {
        if (n <= cap) return BFL_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        BFL_CUDA(cudaMalloc(&p, n * sizeof(T)));
        cap = n;
        return BFL_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    ~DevBuf() { release(); }
};

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
// Warp-uniform broadcasts: all lanes already hold the same value, but routing it through lane 0 tells the
// compiler so; branches on the result are then uniform and shuffles under them stay plain SHFL instead of
// WARPSYNC.COLLECTIVE calls.
__device__ __forceinline__ int uni(int v) { return __shfl_sync(FULL, v, 0); }
__device__ __forceinline__ float uni(float v) { return __shfl_sync(FULL, v, 0); }
__device__ __forceinline__ bool uni(bool v) { return __shfl_sync(FULL, (int)v, 0) != 0; }
__device__ __forceinline__ long long uni(long long v) {
    const int lo = __shfl_sync(FULL, (int)(v & 0xffffffffll), 0), hi = __shfl_sync(FULL, (int)(v >> 32), 0);
    return ((long long)hi << 32) | (unsigned int)lo;
}
__device__ __forceinline__ int warp_id_uniform() { return __shfl_sync(FULL, (int)(threadIdx.x >> 5), 0); }

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// Philox4x32-10, identical to oracle/buffalo_oracle.c::philox4x32_10
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ uint32_t draw_u32(uint32_t seed, uint32_t epoch, uint64_t idx, uint32_t t) {
    uint32_t o[4];
    philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), t >> 2, epoch, seed, 0x5EEDu, o);
    return o[t & 3];
}
__device__ __forceinline__ int32_t draw_range(uint32_t seed, uint32_t epoch, uint64_t idx, uint32_t t,
                                              uint32_t range) {
    return (int32_t)__umulhi(draw_u32(seed, epoch, idx, t), range);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
#endif  // __CUDACC__

}  // namespace bfl
