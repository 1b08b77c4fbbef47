// Probe: which shared-memory floats does tcgen05.mma kind::tf32 read for a K-major A operand under a given
// (layout type, SBO, LBO, start offset)?  One MMA (K = 8) against an identity B tile: D[m][k] = A[m][k]; the tile holds
// its own float index (two passes: low 11 bits / high bits, tf32 keeps integers < 2048 exactly).
// usage: probe_kmajor_readout layout sbo lbo startoff      (one configuration per process: a bad descriptor faults)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
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
struct Args { int layout, sbo, lbo, start, pass; };

__global__ void __launch_bounds__(128) probe(float* D, Args a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar[1];
    const int tid = threadIdx.x;
    float* st = (float*)smem;                       // 32 KB tile area
    float* sw = st + 8192;                          // identity B [N = 16][K = 8], canonical K-major: (n,k) at n*16 + (k/4)*256 + (k%4)*4 bytes
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 8192; i += 128) st[i] = a.pass == 0 ? (float)(i & 2047) : (float)(i >> 11);
    for (int i = tid; i < 128; i += 128) sw[i] = 0.f;
    __syncthreads();
    if (tid < 8) *(float*)((unsigned char*)sw + tid * 16 + (tid >> 2) * 256 + (tid & 3) * 4) = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_slot;
    if (tid == 0) {
        const uint32_t id = idesc_tf32(128, 16, 0, 0);
        const uint64_t da = umma_desc(smem_u32(st) + (uint32_t)a.start, (uint32_t)a.lbo, (uint32_t)a.sbo, (uint32_t)a.layout);
        const uint64_t db = umma_desc(smem_u32(sw), 256, 128, 0);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tm), "l"(da), "l"(db), "r"(id), "r"(0) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[0])) : "memory");
    }
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar[0])), "r"(0) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = tid >> 5, lane = tid & 31;
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(tm + ((uint32_t)(warp * 32) << 16)) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; i++) D[(warp * 32 + lane) * 8 + i] = __uint_as_float(r[i]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(32) : "memory");
}

int main(int argc, char** argv) {
    if (argc < 5) { printf("usage: layout sbo lbo startoff\n"); return 1; }
    Args a{atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), 0};
    float* dD;
    cudaMalloc(&dD, 128 * 8 * 4);
    std::vector<float> lo(1024), hi(1024);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    for (int p = 0; p < 2; p++) {
        a.pass = p;
        probe<<<1, 128, 40000>>>(dD, a);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("layout=%d sbo=%d lbo=%d start=%d: cuda error %s\n", a.layout, a.sbo, a.lbo, a.start, cudaGetErrorString(e)); return 0; }
        cudaMemcpy((p ? hi : lo).data(), dD, 1024 * 4, cudaMemcpyDeviceToHost);
    }
    printf("layout=%d sbo=%d lbo=%d start=%d: per row, byte offsets of k=0..7 (row: off/128 : off%%128)\n", a.layout, a.sbo, a.lbo, a.start);
    for (int m : {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 15, 16, 31, 32, 64, 100, 127}) {
        printf("  m=%3d:", m);
        for (int k = 0; k < 8; k++) {
            int idx = (int)hi[m * 8 + k] * 2048 + (int)lo[m * 8 + k];
            printf(" %d:%d", idx * 4 / 128, idx * 4 % 128);
        }
        printf("\n");
    }
    return 0;
}
