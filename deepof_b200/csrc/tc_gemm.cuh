// tcgen05 (5th-gen tensor core) versions of the two tall-skinny GEMMs, sm_100a only.
//
// Accuracy: operands are fp32 in HBM.  Each is split in-kernel into tf32 "hi" (top 19 bits)
// and "lo" (the exact remainder, again truncated to tf32) and the product is formed with
// three kind::tf32 MMAs  (hi.hi + lo.hi + hi.lo)  accumulated in fp32 in TMEM — the classic
// 3xTF32 scheme, ~2^-22 relative error per product, i.e. fp32-class results, which is what
// the 1e-4 rel-L2 parity gate of the embeddings needs (single-pass tf32/bf16 does not hold it
// through 2x25 recurrent steps + LayerNorms, SURVEY section 7 "hard parts").
//
// Data movement: these GEMMs are HBM-bound (K, N <= 96): operand tiles are staged by the CTA's
// threads with coalesced float4 loads straight into the canonical no-swizzle UMMA shared-memory
// layouts (core matrix = 8 x 16 B), so the split costs no extra pass; accumulators live in
// TMEM; one elected thread issues the MMAs and commits to an mbarrier; the epilogue reads TMEM
// with tcgen05.ld (32 lanes x 32 bit per warp) and applies bias / ReLU / accumulate / mask.
//
//   gemm_rows_tc  : C[M,N] = epi(A[M,K] . W^T + b)     A row tile [128 x K] is operand A
//                   (K-major), W [N x K] operand B (K-major), persistent CTAs over row tiles.
//   gemm_wgrad_tc : dW[N,K] += P[M,N]^T . Q[M,K]       reduction over rows = MMA K dimension;
//                   P^T / Q^T are MN-major operands (rows of P/Q are contiguous along n/k);
//                   a constant-one column appended to Q yields the bias gradient for free.
#pragma once
#include "common.cuh"
#include "gemm.cuh"

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    // bounded spin: a lost arrival must fault the kernel, never hang the GPU
    for (unsigned long long i = 0; i < (1ull << 31); i++)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor, SWIZZLE_NONE, sm_100 version field = 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor: D=f32, A=B=tf32, M=128
__host__ __device__ __forceinline__ uint32_t umma_idesc_tf32(int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    lo = __uint_as_float(__float_as_uint(v - hi) & 0xFFFFE000u);
}
__device__ __forceinline__ void split_tf32x4(const float4 v, float4& hi, float4& lo) {
    split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y);
    split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
}

static inline int tmem_cols_for(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

// ---------------------------------------------------------------------------
// gemm_rows on tcgen05
// ---------------------------------------------------------------------------
#define TC_A_LBO (128 * 16 + 16)    // K-chunk (4 floats) stride of the A tile; +16 B breaks STS bank conflicts

struct TcRowsGeom { int KP, NP, tmem_cols, w_lbo; uint32_t a_bytes, w_bytes; };

__global__ void __launch_bounds__(128) gemm_rows_tc_kernel(const GemmBatch gb, const TcRowsGeom geo) {
    const GemmArgs& g = gb.g[blockIdx.z];
    extern __shared__ __align__(128) unsigned char tsm[];
    unsigned char* A_hi = tsm;
    unsigned char* A_lo = A_hi + geo.a_bytes;
    unsigned char* W_hi = A_lo + geo.a_bytes;
    unsigned char* W_lo = W_hi + geo.w_bytes;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(W_lo + geo.w_bytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KP = geo.KP, NP = geo.NP, KQ = KP >> 2;
    const int ntiles = (g.M + 127) / 128;
    if ((int)blockIdx.x >= ntiles) return;

    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), geo.tmem_cols);
    if (tid == 0) {
        mbar_init(smem_u32(mbar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // stage W (hi/lo) once: canonical K-major, rows n at 16 B, K-chunks at w_lbo
    for (int i = tid; i < NP * KP; i += 128) {
        int n, k;
        if (g.wT == 0) { n = i / KP; k = i % KP; } else { k = i / NP; n = i % NP; }
        float v = 0.f;
        if (n < g.N && k < g.K) v = g.wT == 0 ? __ldg(g.W + (size_t)n * g.ldw + k) : __ldg(g.W + (size_t)k * g.ldw + n);
        float hi, lo;
        split_tf32(v, hi, lo);
        uint32_t off = (uint32_t)n * 16 + (uint32_t)(k >> 2) * geo.w_lbo + (k & 3) * 4;
        *reinterpret_cast<float*>(W_hi + off) = hi;
        *reinterpret_cast<float*>(W_lo + off) = lo;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t idesc = umma_idesc_tf32(NP, 0, 0);
    const uint32_t a_hi_s = smem_u32(A_hi), a_lo_s = smem_u32(A_lo), w_hi_s = smem_u32(W_hi), w_lo_s = smem_u32(W_lo);
    const uint32_t bar = smem_u32(mbar);
    uint32_t phase = 0;
    const bool split = g.A.mode == A_SPLIT;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int m0 = tile * 128;
        // ---- stage the A tile: coalesced float4 loads -> hi/lo -> canonical K-major smem
        for (int i = tid; i < 128 * KQ; i += 128) {
            int row = i / KQ, kq = i - row * KQ;
            int m = m0 + row, c = kq * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < g.M && c < g.K) {
                if (split && c >= g.A.split) c += g.A.skip;
                v = __ldg(reinterpret_cast<const float4*>(g.A.p + (size_t)m * g.A.ld + c));
            }
            float4 hi, lo;
            split_tf32x4(v, hi, lo);
            uint32_t off = (uint32_t)row * 16 + (uint32_t)kq * TC_A_LBO;
            *reinterpret_cast<float4*>(A_hi + off) = hi;
            *reinterpret_cast<float4*>(A_lo + off) = lo;
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            for (int ks = 0; ks < (KP >> 3); ks++) {
                uint32_t ao = (uint32_t)ks * 2 * TC_A_LBO, wo = (uint32_t)ks * 2 * geo.w_lbo;
                uint64_t dah = umma_desc(a_hi_s + ao, TC_A_LBO, 128), dal = umma_desc(a_lo_s + ao, TC_A_LBO, 128);
                uint64_t dbh = umma_desc(w_hi_s + wo, geo.w_lbo, 128), dbl = umma_desc(w_lo_s + wo, geo.w_lbo, 128);
                umma_tf32(tmem, dah, dbh, idesc, ks > 0 ? 1u : 0u);
                umma_tf32(tmem, dal, dbh, idesc, 1u);
                umma_tf32(tmem, dah, dbl, idesc, 1u);
            }
            umma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        // ---- epilogue: thread = row
        const int m = m0 + warp * 32 + lane;
        const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
        const bool vec = ((g.ldc & 3) == 0) && ((g.N & 3) == 0) && (g.mask == nullptr || (g.ldmask & 3) == 0);
        for (int c0 = 0; c0 < NP; c0 += 16) {
            float v[16];
            tmem_ld16(trow + c0, v);
            if (m < g.M) {
                float* cp = g.C + (size_t)m * g.ldc + c0;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    int n = c0 + q * 4;
                    if (n >= g.N) break;
                    float o[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) o[j] = v[q * 4 + j] + ((g.bias && n + j < g.N) ? __ldg(g.bias + n + j) : 0.f);
                    if (vec) {
                        if (g.accum) {
                            float4 c = *reinterpret_cast<const float4*>(cp + q * 4);
                            o[0] += c.x; o[1] += c.y; o[2] += c.z; o[3] += c.w;
                        }
                        if (g.relu) {
#pragma unroll
                            for (int j = 0; j < 4; j++) o[j] = fmaxf(o[j], 0.f);
                        }
                        if (g.mask) {
                            float4 mk = *reinterpret_cast<const float4*>(g.mask + (size_t)m * g.ldmask + n);
                            o[0] = mk.x > 0.f ? o[0] : 0.f; o[1] = mk.y > 0.f ? o[1] : 0.f;
                            o[2] = mk.z > 0.f ? o[2] : 0.f; o[3] = mk.w > 0.f ? o[3] : 0.f;
                        }
                        *reinterpret_cast<float4*>(cp + q * 4) = make_float4(o[0], o[1], o[2], o[3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            if (n + j >= g.N) continue;
                            float x = o[j];
                            if (g.accum) x += cp[q * 4 + j];
                            if (g.relu) x = fmaxf(x, 0.f);
                            if (g.mask) x = g.mask[(size_t)m * g.ldmask + n + j] > 0.f ? x : 0.f;
                            cp[q * 4 + j] = x;
                        }
                    }
                }
            }
        }
        tc_fence_before();   // TMEM reads done before the next tile's MMAs (ordered by the next __syncthreads)
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, geo.tmem_cols);
}

static bool g_tc_enabled = true;
static bool g_tc_env_checked = false;
static inline bool tc_enabled() {
    if (!g_tc_env_checked) {
        const char* e = getenv("DOF_DISABLE_TC");
        if (e && e[0] == '1') g_tc_enabled = false;
        g_tc_env_checked = true;
    }
    return g_tc_enabled;
}

static inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

static bool tc_rows_eligible(const GemmArgs& g) {
    if (g.M < 2048 || g.N > 256 || g.K > 128 || g.N < 8) return false;
    if (g.A.mode != A_PLAIN && g.A.mode != A_SPLIT) return false;
    if ((g.A.ld & 3) || (g.K & 3) || !aligned16(g.A.p)) return false;
    if (g.A.mode == A_SPLIT && ((g.A.split & 3) || (g.A.skip & 3))) return false;
    if (!aligned16(g.C) || (g.mask && !aligned16(g.mask))) return false;
    return true;
}

static int launch_gemm_rows_tc(const GemmArgs* gs, int nbatch, cudaStream_t st, int sm_count) {
    GemmBatch gb;
    memset(&gb, 0, sizeof(gb));
    int M = 0;
    for (int i = 0; i < nbatch; i++) {
        gb.g[i] = gs[i];
        if (gs[i].M > M) M = gs[i].M;
        if (gs[i].N != gs[0].N || gs[i].K != gs[0].K) DOF_FAIL(DOF_ERR_ARG, "batched TC GEMMs must share N and K");
    }
    TcRowsGeom geo;
    geo.KP = round_up(gs[0].K, 8);
    geo.NP = round_up(gs[0].N, 16);
    geo.tmem_cols = tmem_cols_for(geo.NP);
    geo.w_lbo = geo.NP * 16 + 16;
    geo.a_bytes = (uint32_t)(geo.KP / 4) * TC_A_LBO;
    geo.w_bytes = (uint32_t)(geo.KP / 4) * geo.w_lbo;
    size_t smem = 2 * (size_t)geo.a_bytes + 2 * (size_t)geo.w_bytes + 64;
    static size_t attr_max = 0;
    if (smem > attr_max) {
        DOF_CUDA(cudaFuncSetAttribute(gemm_rows_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_max = 200 * 1024;
    }
    if (smem > 200 * 1024) DOF_FAIL(DOF_ERR_UNSUPPORTED, "TC GEMM tile does not fit shared memory");
    int occ = (int)((220 * 1024) / (smem + 1024));
    int by_tmem = 512 / geo.tmem_cols;
    if (occ > by_tmem) occ = by_tmem;
    if (occ > 4) occ = 4;
    if (occ < 1) occ = 1;
    int ntiles = cdiv(M, 128);
    int ctas = sm_count * occ / nbatch;
    if (ctas < 1) ctas = 1;
    if (ctas > ntiles) ctas = ntiles;
    double fl = 0.0;
    for (int i = 0; i < nbatch; i++) fl += 2.0 * gs[i].M * gs[i].N * gs[i].K;
    ProfScope ps("gemm_rows_tc", st, fl);
    gemm_rows_tc_kernel<<<dim3(ctas, 1, nbatch), 128, smem, st>>>(gb, geo);
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

// ---------------------------------------------------------------------------
// gemm_wgrad on tcgen05:  D[n, k] (+ ones column) accumulated in TMEM over this CTA's rows
// ---------------------------------------------------------------------------
#define TCW_BM 64                       // rows (MMA-K) per stage
#define TCW_SBO (TCW_BM * 16 + 16)      // stride between 4-wide n (or k) blocks; +16 B vs bank conflicts

struct TcWgradGeom { int KWP, tmem_cols; uint32_t p_bytes, q_bytes; };

__global__ void __launch_bounds__(128) gemm_wgrad_tc_kernel(const WGradBatch wb, const TcWgradGeom geo) {
    const WGradArgs& g = wb.g[blockIdx.z];
    extern __shared__ __align__(128) unsigned char tsm[];
    unsigned char* P_hi = tsm;
    unsigned char* P_lo = P_hi + geo.p_bytes;
    unsigned char* Q_hi = P_lo + geo.p_bytes;
    unsigned char* Q_lo = Q_hi + geo.q_bytes;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(Q_lo + geo.q_bytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int rows_per = (g.M + gridDim.x - 1) / gridDim.x;
    rows_per = (rows_per + TCW_BM - 1) / TCW_BM * TCW_BM;
    const int mbeg = blockIdx.x * rows_per;
    const int mend = min(g.M, mbeg + rows_per);
    if (mbeg >= mend) return;

    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), geo.tmem_cols);
    if (tid == 0) {
        mbar_init(smem_u32(mbar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // zero everything once: padded n rows (>= N) and k columns (> K) stay zero forever
    for (uint32_t i = tid * 16; i < 2 * geo.p_bytes + 2 * geo.q_bytes; i += 128 * 16)
        *reinterpret_cast<float4*>(tsm + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t idesc = umma_idesc_tf32(geo.KWP, 1, 1);
    const uint32_t p_hi_s = smem_u32(P_hi), p_lo_s = smem_u32(P_lo), q_hi_s = smem_u32(Q_hi), q_lo_s = smem_u32(Q_lo);
    const uint32_t bar = smem_u32(mbar);
    uint32_t phase = 0;
    const int NQ = g.N >> 2, KQ = g.K >> 2;
    const bool psplit = g.P.mode == A_SPLIT;
    const bool qshift = g.Q.mode == A_TSHIFT;
    bool first = true;

    for (int mb = mbeg; mb < mend; mb += TCW_BM) {
        // ---- stage P (as A^T: n-blocks of 4 at TCW_SBO, rows at 16 B) and Q likewise
        for (int i = tid; i < TCW_BM * NQ; i += 128) {
            int r = i / NQ, nq = i - r * NQ;
            int m = mb + r, c = nq * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < mend) {
                if (psplit && c >= g.P.split) c += g.P.skip;
                v = __ldg(reinterpret_cast<const float4*>(g.P.p + (size_t)m * g.P.ld + c));
            }
            float4 hi, lo;
            split_tf32x4(v, hi, lo);
            uint32_t off = (uint32_t)nq * TCW_SBO + (uint32_t)r * 16;
            *reinterpret_cast<float4*>(P_hi + off) = hi;
            *reinterpret_cast<float4*>(P_lo + off) = lo;
        }
        for (int i = tid; i < TCW_BM * (KQ + 1); i += 128) {
            int r = i / (KQ + 1), kq = i - r * (KQ + 1);
            int m = mb + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < mend) {
                if (kq == KQ) {
                    v.x = 1.0f;                       // ones column -> bias gradient
                } else if (qshift) {
                    int t = m % g.Q.T, tt = t + g.Q.shift;
                    if (tt >= 0 && tt < g.Q.T)
                        v = __ldg(reinterpret_cast<const float4*>(g.Q.p + (size_t)(m + g.Q.shift) * g.Q.ld + kq * 4));
                } else {
                    v = __ldg(reinterpret_cast<const float4*>(g.Q.p + (size_t)m * g.Q.ld + kq * 4));
                }
            }
            float4 hi, lo;
            split_tf32x4(v, hi, lo);
            uint32_t off = (uint32_t)kq * TCW_SBO + (uint32_t)r * 16;
            *reinterpret_cast<float4*>(Q_hi + off) = hi;
            *reinterpret_cast<float4*>(Q_lo + off) = lo;
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            for (int kb = 0; kb < TCW_BM / 8; kb++) {
                uint32_t o = (uint32_t)kb * 128;
                uint64_t dah = umma_desc(p_hi_s + o, 128, TCW_SBO), dal = umma_desc(p_lo_s + o, 128, TCW_SBO);
                uint64_t dbh = umma_desc(q_hi_s + o, 128, TCW_SBO), dbl = umma_desc(q_lo_s + o, 128, TCW_SBO);
                umma_tf32(tmem, dah, dbh, idesc, (first && kb == 0) ? 0u : 1u);
                umma_tf32(tmem, dal, dbh, idesc, 1u);
                umma_tf32(tmem, dah, dbl, idesc, 1u);
            }
            umma_commit(bar);
        }
        first = false;
        mbar_wait(bar, phase);   // MMAs done reading smem -> safe to restage
        phase ^= 1;
    }
    tc_fence_after();
    // ---- epilogue: thread = output row n; atomics into dW / db
    const int n = warp * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < geo.KWP; c0 += 16) {
        float v[16];
        tmem_ld16(trow + c0, v);
        if (n < g.N) {
#pragma unroll
            for (int j = 0; j < 16; j++) {
                int k = c0 + j;
                if (k < g.K) {
                    float* o = g.oT ? g.dW + (size_t)k * g.ldo + n : g.dW + (size_t)n * g.ldo + k;
                    atomicAdd(o, v[j]);
                } else if (k == g.K && g.db) {
                    atomicAdd(g.db + n, v[j]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, geo.tmem_cols);
}

static bool tc_wgrad_eligible(const WGradArgs& g) {
    if (g.M < 4096 || g.N > 128 || g.K + 4 > 256 || g.N < 4 || g.K < 4) return false;
    if ((g.N & 3) || (g.K & 3)) return false;
    if (g.P.mode != A_PLAIN && g.P.mode != A_SPLIT) return false;
    if (g.Q.mode != A_PLAIN && g.Q.mode != A_TSHIFT) return false;
    if ((g.P.ld & 3) || (g.Q.ld & 3) || !aligned16(g.P.p) || !aligned16(g.Q.p)) return false;
    if (g.P.mode == A_SPLIT && ((g.P.split & 3) || (g.P.skip & 3))) return false;
    return true;
}

static int launch_gemm_wgrad_tc(const WGradArgs* gs, int nbatch, cudaStream_t st, int sm_count) {
    WGradBatch wb;
    memset(&wb, 0, sizeof(wb));
    int M = 0;
    for (int i = 0; i < nbatch; i++) {
        wb.g[i] = gs[i];
        if (gs[i].M > M) M = gs[i].M;
        if (gs[i].N != gs[0].N || gs[i].K != gs[0].K) DOF_FAIL(DOF_ERR_ARG, "batched TC wgrads must share N and K");
    }
    TcWgradGeom geo;
    geo.KWP = round_up(gs[0].K + 4, 16);
    geo.tmem_cols = tmem_cols_for(geo.KWP);
    geo.p_bytes = 32u * TCW_SBO;                       // 128 n rows
    geo.q_bytes = (uint32_t)(geo.KWP / 4) * TCW_SBO;
    size_t smem = 2 * (size_t)geo.p_bytes + 2 * (size_t)geo.q_bytes + 64;
    static bool attr = false;
    if (!attr) {
        DOF_CUDA(cudaFuncSetAttribute(gemm_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    if (smem > 200 * 1024) DOF_FAIL(DOF_ERR_UNSUPPORTED, "TC wgrad tile does not fit shared memory");
    int occ = (int)((220 * 1024) / (smem + 1024));
    int by_tmem = 512 / geo.tmem_cols;
    if (occ > by_tmem) occ = by_tmem;
    if (occ > 3) occ = 3;
    if (occ < 1) occ = 1;
    int ctas = sm_count * occ / nbatch;
    int maxsplit = cdiv(M, 4 * TCW_BM);
    if (ctas > maxsplit) ctas = maxsplit;
    if (ctas < 1) ctas = 1;
    double fl = 0.0;
    for (int i = 0; i < nbatch; i++) fl += 2.0 * gs[i].M * gs[i].N * gs[i].K;
    ProfScope ps("gemm_wgrad_tc", st, fl);
    gemm_wgrad_tc_kernel<<<dim3(ctas, 1, nbatch), 128, smem, st>>>(wb, geo);
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

// ---------------------------------------------------------------------------
// dispatch: tensor-core path when every problem of the batch is eligible, SIMT otherwise
// ---------------------------------------------------------------------------
static int g_sm_count = 148;

static int launch_gemm_rows(const GemmArgs* gs, int nbatch, cudaStream_t st) {
    bool tc = tc_enabled();
    for (int i = 0; i < nbatch && tc; i++)
        tc = tc_rows_eligible(gs[i]) && gs[i].N == gs[0].N && gs[i].K == gs[0].K;
    if (tc) return launch_gemm_rows_tc(gs, nbatch, st, g_sm_count);
    return launch_gemm_rows_simt(gs, nbatch, st);
}

static int launch_gemm_wgrad(const WGradArgs* gs, int nbatch, cudaStream_t st, int sm_count) {
    bool tc = tc_enabled();
    for (int i = 0; i < nbatch && tc; i++)
        tc = tc_wgrad_eligible(gs[i]) && gs[i].N == gs[0].N && gs[i].K == gs[0].K;
    if (tc) return launch_gemm_wgrad_tc(gs, nbatch, st, sm_count);
    return launch_gemm_wgrad_simt(gs, nbatch, st, sm_count);
}
