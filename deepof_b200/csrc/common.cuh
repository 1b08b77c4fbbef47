// Common helpers for the deepof_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <vector>

#define DOF_OK 0
#define DOF_ERR_ARG -1
#define DOF_ERR_CUDA -2
#define DOF_ERR_UNSUPPORTED -3
#define DOF_ERR_WORKSPACE -4

extern thread_local char g_dof_err[512];

#define DOF_FAIL(code, ...)                                   \
    do {                                                      \
        snprintf(g_dof_err, sizeof(g_dof_err), __VA_ARGS__);  \
        return (code);                                        \
    } while (0)

#define DOF_CUDA(expr)                                                               \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess)                                                       \
            DOF_FAIL(DOF_ERR_CUDA, "%s:%d CUDA error %s: %s", __FILE__, __LINE__,    \
                     cudaGetErrorName(_e), cudaGetErrorString(_e));                  \
    } while (0)

#define DOF_LAUNCH_CHECK() DOF_CUDA(cudaGetLastError())

#define DOF_TRY(expr)              \
    do {                           \
        int _r = (expr);           \
        if (_r != DOF_OK) return _r; \
    } while (0)

// ---- optional per-kernel-class CUDA-event timing + launch counting ---------------
struct DofProfRec { const char* name; cudaEvent_t a, b; double flops, bytes; };
struct DofProf {
    bool enabled = false;
    long long launches = 0;
    std::vector<DofProfRec> recs;
};
extern DofProf g_prof;
struct ProfScope {
    const char* name; cudaStream_t st; cudaEvent_t a, b; bool on; double flops, bytes;
    // flops / bytes: ALGORITHMIC work of this launch (unpadded), used for roofline reporting
    ProfScope(const char* n, cudaStream_t s, double fl = 0.0, double by = 0.0)
        : name(n), st(s), on(g_prof.enabled), flops(fl), bytes(by) {
        g_prof.launches++;
        if (on) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, st); }
    }
    ~ProfScope() {
        if (on) { cudaEventRecord(b, st); g_prof.recs.push_back({name, a, b, flops, bytes}); }
    }
};

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// ---- device math -----------------------------------------------------------
// Gate non-linearities.  fp32 throughout; fast MUFU-based exp is accurate to ~2 ulp
// on the bounded arguments seen here, far inside the 1e-4 rel-L2 parity budget.
__device__ __forceinline__ float sigmoid_f(float x) {
    return __fdividef(1.0f, 1.0f + __expf(-x));
}
__device__ __forceinline__ float tanh_f(float x) {
    // 1 - 2/(exp(2x)+1): absolute error ~1e-7, saturates correctly for |x| large.
    float e = __expf(2.0f * x);
    return 1.0f - __fdividef(2.0f, e + 1.0f);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- counter-based random numbers (Philox4x32-10): the same (seed, site, element) always gives the same number, so a
// forward and a backward kernel regenerate identical noise and nothing is stored in HBM ---------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int i = 0; i < 10; i++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}

// standard normal number `idx` of stream (seed, site): Box-Muller on one Philox block per 4 consecutive elements
__device__ __forceinline__ float philox_normal(unsigned long long seed, unsigned int site, unsigned long long idx) {
    const unsigned long long blk = idx >> 2;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), site, 2u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t a = (idx & 2) ? r.z : r.x, b = (idx & 2) ? r.w : r.y;
    const float u1 = ((float)a + 1.0f) * 2.3283064365386963e-10f;          // (0, 1]
    const float rad = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincospif((float)b * 4.656612873077393e-10f, &sn, &cs);                // angle = 2 pi b / 2^32
    return rad * ((idx & 1) ? sn : cs);
}
