// Probe: tcgen05.mma kind::tf32 with MN-major operands in the SWIZZLE_128B_BASE32B layout (the only MN-major layout
// tf32 accepts), filled (a) by hand and (b) by TMA with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; (c) how the tensor core
// narrows fp32 bit patterns to tf32.  D[m][n] = sum_k A[k][m] * B[k][n], A and B row-major [k][*] in global memory.
// nvcc -gencode arch=compute_100a,code=sm_100a -o tools/probe_mn_major tools/probe_mn_major.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
           (1ull << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ uint32_t idesc_tf32(int M, int N, int amn, int bmn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)amn << 15) | ((uint32_t)bmn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Args { int K, M, N, use_tma, swap_lbo_sbo; };

// byte offset of element (k, c) of a [K rows][32*nblk cols] tile in the SW128_32B MN-major arrangement
__host__ __device__ inline uint32_t sw_off(int k, int c, int K) {
    return (uint32_t)(c / 32) * (uint32_t)(K * 128) + (uint32_t)k * 128 + (uint32_t)((((c % 32) / 8) ^ (k % 4)) * 32) + (uint32_t)(c % 8) * 4;
}

__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* D, Args a,
                                             const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar[2];
    unsigned char* sa = smem;
    unsigned char* sb = smem + (size_t)(a.M / 32) * a.K * 128;
    const int tid = threadIdx.x;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (!a.use_tma) {
        for (int i = tid; i < a.K * a.M; i += 128) *(float*)(sa + sw_off(i / a.M, i % a.M, a.K)) = A[i];
        for (int i = tid; i < a.K * a.N; i += 128) *(float*)(sb + sw_off(i / a.N, i % a.N, a.K)) = B[i];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    } else if (tid == 0) {
        const uint32_t bytes = (uint32_t)(a.K * (a.M + a.N) * 4);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[1])), "r"(bytes) : "memory");
        for (int b = 0; b < a.M / 32; b++)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(sa + (size_t)b * a.K * 128)), "l"(&tmA), "r"(b * 32), "r"(0), "r"(smem_u32(&bar[1])) : "memory");
        for (int b = 0; b < a.N / 32; b++)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(sb + (size_t)b * a.K * 128)), "l"(&tmB), "r"(b * 32), "r"(0), "r"(smem_u32(&bar[1])) : "memory");
    }
    if (a.use_tma) {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar[1])), "r"(0) : "memory");
    }
    __syncthreads();
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_slot;
    if (tid == 0) {
        const uint32_t id = idesc_tf32(a.M, a.N, 1, 1);
        const uint32_t blk = (uint32_t)a.K * 128, grp = 512;
        const uint32_t lbo = a.swap_lbo_sbo ? grp : blk, sbo = a.swap_lbo_sbo ? blk : grp;
        for (int ks = 0; ks < a.K / 8; ks++) {
            const uint64_t da = umma_desc(smem_u32(sa) + ks * 1024, lbo, sbo, 1);
            const uint64_t db = umma_desc(smem_u32(sb) + ks * 1024, lbo, sbo, 1);
            const uint32_t acc = ks > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tm), "l"(da), "l"(db), "r"(id), "r"(acc) : "memory");
        }
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
    for (int c0 = 0; c0 < a.N; c0 += 16) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(tm + ((uint32_t)(warp * 32) << 16) + c0) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; i++) D[(warp * 32 + lane) * a.N + c0 + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(64) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    return (EncodeFn)fn;
}

static CUtensorMap make_map(EncodeFn enc, float* base, int rows, int cols, int box_rows) {
    CUtensorMap m;
    memset(&m, 0, sizeof(m));
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
    return m;
}

static float trunc_tf32(float v) { uint32_t u; memcpy(&u, &v, 4); u &= 0xFFFFE000u; memcpy(&v, &u, 4); return v; }
static float rne_tf32(float v) { uint32_t u; memcpy(&u, &v, 4); u += 0xFFFu + ((u >> 13) & 1u); u &= 0xFFFFE000u; memcpy(&v, &u, 4); return v; }
static float rna_tf32(float v) { uint32_t u; memcpy(&u, &v, 4); u += 0x1000u; u &= 0xFFFFE000u; memcpy(&v, &u, 4); return v; }

static void run(Args a, const char* name, bool ints, EncodeFn enc) {
    std::vector<float> A(a.K * a.M), B(a.K * a.N), D(a.M * a.N);
    std::vector<double> R(a.M * a.N, 0.0), Rt(a.M * a.N, 0.0), Rn(a.M * a.N, 0.0), Ra(a.M * a.N, 0.0);
    srand(7);
    for (auto& v : A) v = ints ? (float)(rand() % 17 - 8) : (float)rand() / RAND_MAX - 0.5f;
    for (auto& v : B) v = ints ? (float)(rand() % 13 - 6) : (float)rand() / RAND_MAX - 0.5f;
    for (int k = 0; k < a.K; k++) for (int m = 0; m < a.M; m++) for (int n = 0; n < a.N; n++) {
        R[m * a.N + n] += (double)A[k * a.M + m] * B[k * a.N + n];
        Rt[m * a.N + n] += (double)trunc_tf32(A[k * a.M + m]) * trunc_tf32(B[k * a.N + n]);
        Rn[m * a.N + n] += (double)rne_tf32(A[k * a.M + m]) * rne_tf32(B[k * a.N + n]);
        Ra[m * a.N + n] += (double)rna_tf32(A[k * a.M + m]) * rna_tf32(B[k * a.N + n]);
    }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, D.size() * 4);
    CUtensorMap tA = make_map(enc, dA, a.K, a.M, a.K), tB = make_map(enc, dB, a.K, a.N, a.K);
    size_t smem = (size_t)a.K * (a.M + a.N) * 4 + 2048;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<<<1, 128, smem>>>(dA, dB, dD, a, tA, tB);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double err = 0, et = 0, en = 0, ea = 0; int nz = 0;
    for (size_t i = 0; i < D.size(); i++) {
        double d = fabs((double)D[i] - R[i]); if (!(d <= err)) err = d;
        et = fmax(et, fabs((double)D[i] - Rt[i])); en = fmax(en, fabs((double)D[i] - Rn[i])); ea = fmax(ea, fabs((double)D[i] - Ra[i]));
        nz += D[i] != 0.f;
    }
    printf("%-40s err=%.3g (vs trunc %.3g, rne %.3g, rna %.3g) nonzero=%d/%zu cuda=%s D=%g %g R=%g %g\n", name, err, et, en, ea, nz, D.size(),
           cudaGetErrorString(e), D[0], D[1], R[0], R[1]);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

int main() {
    EncodeFn enc = get_encode();
    if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    for (int N : {32, 64, 48}) {
        printf("--- M=128 N=%d K=64\n", N);
        const int NB = (N + 31) / 32 * 32;   // the tile always holds whole 32-column blocks
        (void)NB;
        if (N % 32 == 0) {
            run(Args{64, 128, N, 0, 0}, "manual fill, lbo=block sbo=512", true, enc);
            run(Args{64, 128, N, 1, 0}, "TMA fill,    lbo=block sbo=512", true, enc);
            run(Args{64, 128, N, 1, 0}, "TMA fill, random floats (lbo=block)", false, enc);
        }
    }
    return 0;
}
