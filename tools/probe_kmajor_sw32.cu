// Probe: ONE shared-memory tile [R rows][32*nblk columns] fp32 in the SWIZZLE_128B_BASE32B arrangement
//   off(r, c) = (c / 32) * R * 128 + r * 128 + (((c % 32) / 8) ^ (r % 4)) * 32 + (c % 8) * 4
// read by tcgen05.mma kind::tf32 in BOTH roles:
//   (1) K-major A operand (M = tile row, K = tile column):   D1[m][n] = sum_c T[m][c] * W[n][c]
//       (W: canonical no-swizzle K-major B operand, as the GRU kernels hold their weights)
//   (2) MN-major A operand (M = tile column, K = tile row):  D2[c][n] = sum_r T[r][c] * X[r][n]
//       (X: a second tile in the same arrangement, MN-major B operand)
//   (3) MN-major A with a K-major all-ones B tile [16][8] (the same tile at every K step):  D3[c][*] = sum_r T[r][c]
// The fused BPTT + weight-gradient kernel needs all three on the same dG tile (gru_bwd_tc.cuh).
// nvcc -gencode arch=compute_100a,code=sm_100a -o tools/probe_kmajor_sw32 tools/probe_kmajor_sw32.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
           (1ull << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ uint32_t idesc_tf32(int M, int N, int amn, int bmn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)amn << 15) | ((uint32_t)bmn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ inline uint32_t sw_off(int r, int c, int R) {
    return (uint32_t)(c / 32) * (uint32_t)(R * 128) + (uint32_t)r * 128 + (uint32_t)((((c % 32) / 8) ^ (r % 4)) * 32) + (uint32_t)(c % 8) * 4;
}

struct Args { int R, C, N, NX, mode, layout, sbo, lbo; };   // tile [R][C]; W [N][C]; X [R][NX]

__device__ __forceinline__ void mma(uint32_t tm, uint64_t da, uint64_t db, uint32_t id, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tm), "l"(da), "l"(db), "r"(id), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128) probe(const float* T, const float* W, const float* X, float* D, Args a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar[1];
    const int tid = threadIdx.x;
    unsigned char* st = smem;                                            // the dual-use tile
    unsigned char* sx = st + (size_t)(a.C / 32) * a.R * 128;             // X tile, same arrangement
    unsigned char* sw = sx + (size_t)(a.NX / 32) * a.R * 128;            // W canonical K-major: (n, k) at n*16 + (k/4)*lbo + (k%4)*4
    const uint32_t w_lbo = (uint32_t)a.N * 16 + 16;
    unsigned char* so = sw + (size_t)(a.C / 4) * w_lbo + 1024;           // ones [8][R] canonical K-major
    so = (unsigned char*)(((uintptr_t)so + 127) & ~(uintptr_t)127);
    const uint32_t o_lbo = 16 * 16;                                      // ones [N = 16][K = 8]: every K step reads the same tile
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // poison the tile area first (rows >= R are read by an M = 128 instruction when R < 128)
    for (int i = tid; i < (a.C / 32) * 128 * 32; i += 128) ((float*)st)[i] = 1e30f;
    __syncthreads();
    for (int i = tid; i < a.R * a.C; i += 128) *(float*)(st + sw_off(i / a.C, i % a.C, a.R)) = T[i];
    for (int i = tid; i < a.R * a.NX; i += 128) *(float*)(sx + sw_off(i / a.NX, i % a.NX, a.R)) = X[i];
    for (int i = tid; i < a.N * a.C; i += 128) {
        const int n = i / a.C, k = i % a.C;
        *(float*)(sw + (uint32_t)n * 16 + (uint32_t)(k >> 2) * w_lbo + (k & 3) * 4) = W[i];
    }
    for (int i = tid; i < 16 * 8; i += 128) ((float*)so)[i] = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_slot;
    int ncols = 0;
    if (tid == 0) {
        if (a.mode == 1) {
            // K-major A from the tile: K step of 8 columns = +32 B inside a 128-byte row, next 32-column block = + R*128
            const uint32_t id = idesc_tf32(128, a.N, 0, 0);
            for (int ks = 0; ks < a.C / 8; ks++) {
                const uint32_t ao = (uint32_t)(ks / 4) * (uint32_t)(a.R * 128) + (uint32_t)(ks % 4) * 32;
                const uint64_t da = umma_desc(smem_u32(st) + ao, (uint32_t)a.lbo, (uint32_t)a.sbo, (uint32_t)a.layout);
                const uint64_t db = umma_desc(smem_u32(sw) + (uint32_t)ks * 2 * w_lbo, w_lbo, 128, 0);
                mma(tm, da, db, id, ks > 0);
            }
        } else if (a.mode == 2) {
            const uint32_t id = idesc_tf32(128, a.NX, 1, 1);   // M = 128 always (C = 64: lanes 64-127 read past the tile and are ignored)
            const uint32_t blk = (uint32_t)a.R * 128;
            for (int ks = 0; ks < a.R / 8; ks++) {
                const uint64_t da = umma_desc(smem_u32(st) + ks * 1024, blk, 512, 1);
                const uint64_t db = umma_desc(smem_u32(sx) + ks * 1024, blk, 512, 1);
                mma(tm, da, db, id, ks > 0);
            }
        } else {
            const uint32_t id = idesc_tf32(128, 16, 1, 0);
            const uint32_t blk = (uint32_t)a.R * 128;
            for (int ks = 0; ks < a.R / 8; ks++) {
                const uint64_t da = umma_desc(smem_u32(st) + ks * 1024, blk, 512, 1);
                const uint64_t db = umma_desc(smem_u32(so), o_lbo, 128, 0);
                mma(tm, da, db, id, ks > 0);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[0])) : "memory");
    }
    ncols = a.mode == 1 ? a.N : (a.mode == 2 ? a.NX : 16);
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar[0])), "r"(0) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = tid >> 5, lane = tid & 31;
    for (int c0 = 0; c0 < ncols; c0 += 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(tm + ((uint32_t)(warp * 32) << 16) + c0) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; i++) D[(warp * 32 + lane) * ncols + c0 + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(128) : "memory");
}

static void run(Args a, const char* name) {
    std::vector<float> T(a.R * a.C), W(a.N * a.C), X(a.R * a.NX), D(128 * 128, 0.f);
    srand(11);
    for (auto& v : T) v = (float)(rand() % 17 - 8);
    for (auto& v : W) v = (float)(rand() % 13 - 6);
    for (auto& v : X) v = (float)(rand() % 11 - 5);
    const int ncols = a.mode == 1 ? a.N : (a.mode == 2 ? a.NX : 16);
    const int nrows = a.mode == 1 ? a.R : a.C;
    std::vector<double> Rf((size_t)nrows * ncols, 0.0);
    if (a.mode == 1) {
        for (int m = 0; m < a.R; m++) for (int n = 0; n < a.N; n++) for (int c = 0; c < a.C; c++) Rf[m * ncols + n] += (double)T[m * a.C + c] * W[n * a.C + c];
    } else if (a.mode == 2) {
        for (int c = 0; c < a.C; c++) for (int n = 0; n < a.NX; n++) for (int r = 0; r < a.R; r++) Rf[c * ncols + n] += (double)T[r * a.C + c] * X[r * a.NX + n];
    } else {
        for (int c = 0; c < a.C; c++) for (int n = 0; n < 16; n++) for (int r = 0; r < a.R; r++) Rf[c * ncols + n] += (double)T[r * a.C + c];
    }
    float *dT, *dW, *dX, *dD;
    cudaMalloc(&dT, T.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dT, T.data(), T.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, D.size() * 4);
    size_t smem = (size_t)(a.C / 32 + a.NX / 32) * 128 * 128 + (size_t)a.C / 4 * (a.N * 16 + 16) + 8 * 128 * 4 + 4096;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<<<1, 128, smem>>>(dT, dW, dX, dD, a);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0; int bad = 0;
    for (int i = 0; i < nrows * ncols; i++) {
        double d = fabs((double)D[i] - Rf[i]);
        if (!(d <= 1e-3)) bad++;
        if (!(d <= err)) err = d;
    }
    printf("%-64s max err=%.3g bad=%d/%d cuda=%s D=%g %g %g R=%g %g %g\n", name, err, bad, nrows * ncols, cudaGetErrorString(e),
           D[0], D[1], D[ncols], Rf[0], Rf[1], Rf[ncols]);
    cudaFree(dT); cudaFree(dW); cudaFree(dX); cudaFree(dD);
}

int main(int argc, char** argv) {
    // one configuration per process (a bad descriptor faults the context): probe_kmajor_sw32 <case 0..5> <R>
    const int c = argc > 1 ? atoi(argv[1]) : 0, R = argc > 2 ? atoi(argv[2]) : 128;
    char nm[128];
    if (c == 0) { snprintf(nm, sizeof nm, "K-major A  R=%d C=128 N=32 layout=1 sbo=512", R); run(Args{R, 128, 32, 64, 1, 1, 512, 0}, nm); }
    if (c == 1) { snprintf(nm, sizeof nm, "K-major A  R=%d C=64 N=64 layout=1 sbo=512", R); run(Args{R, 64, 64, 64, 1, 1, 512, 0}, nm); }
    if (c == 2) { snprintf(nm, sizeof nm, "MN-major A (tile^T . X)  R=%d C=128 NX=64", R); run(Args{R, 128, 32, 64, 2, 1, 0, 0}, nm); }
    if (c == 3) { snprintf(nm, sizeof nm, "MN-major A (tile^T . X)  R=%d C=64 NX=96", R); run(Args{R, 64, 32, 96, 2, 1, 0, 0}, nm); }
    if (c == 4) { snprintf(nm, sizeof nm, "MN-major A . K-major ones N=16 (column sums)  R=%d C=128", R); run(Args{R, 128, 32, 64, 3, 1, 0, 0}, nm); }
    if (c == 5) { snprintf(nm, sizeof nm, "MN-major A . K-major ones N=16 (column sums)  R=%d C=64", R); run(Args{R, 64, 32, 64, 3, 1, 0, 0}, nm); }
    return 0;
}
