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
//                   P^T / Q^T are staged TRANSPOSED into K-major tiles (bank-conflict free);
//                   a constant-one row appended to Q^T yields the bias gradient for free.
// Both kernels are warp-specialised and persistent: 4 producer warps stream tiles from HBM
// (register double-buffered float4 loads -> hi/lo split -> shared-memory ring; the split needs
// a register pass, so TMA cannot stage these operands), one thread issues the MMAs, and (rows
// kernel) 4 epilogue warps drain a double-buffered TMEM accumulator; all hand-offs are
// mbarriers, there is no block-wide barrier in the steady state.
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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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

// One elected lane of a CONVERGED warp (elect.sync).  The MMA-issuing warps run their loops with all 32 lanes and guard only
// the tcgen05.mma / commit with this: the operands are then warp-uniform values in uniform registers.  Inside an
// `if (lane == 0)` region ptxas wraps every tcgen05.mma in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop.
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// TMEM base address as a provably warp-uniform value (REDUX result)
__device__ __forceinline__ uint32_t tmem_base_uniform(const uint32_t* slot) { return __reduce_or_sync(0xffffffffu, *slot); }

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
// Same split with the remainder ROUNDED to nearest tf32 (add half an ulp of the 10-bit mantissa, then truncate): what is lost is
// at most 2^-22 |v| and has no preferred sign.  The truncated remainder of split_tf32 always errs towards zero (mean 2^-22 |v| per
// operand), which showed up as a 6e-7 relative error of K = 128 products; the general-purpose GEMM kernels use this one.
__device__ __forceinline__ void split_tf32_rn(float v, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
#ifdef DOF_AB_OLD_SPLIT
    lo = __uint_as_float(__float_as_uint(v - hi) & 0xFFFFE000u);
#else
    lo = __uint_as_float((__float_as_uint(v - hi) + 0x1000u) & 0xFFFFE000u);
#endif
}
__device__ __forceinline__ void split_tf32x4_rn(const float4 v, float4& hi, float4& lo) {
    split_tf32_rn(v.x, hi.x, lo.x); split_tf32_rn(v.y, hi.y, lo.y);
    split_tf32_rn(v.z, hi.z, lo.z); split_tf32_rn(v.w, hi.w, lo.w);
}

__host__ __device__ static inline int tmem_cols_for(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

// ---------------------------------------------------------------------------
// gemm_rows on tcgen05
// ---------------------------------------------------------------------------
#define TC_A_LBO (128 * 16 + 16)    // K-chunk (4 floats) stride of the A tile; +16 B breaks STS bank conflicts
#define TCR_THREADS 288             // warps 0-3 producers, 4-7 epilogue (TMEM lane group = warp & 3), warp 8 MMA

struct TcRowsGeom { int KP, NP, tmem_cols, w_lbo, stages; uint32_t a_bytes, w_bytes; };

// KQM: float4 items held in registers per producer thread and per set; NSET register sets; PW producer warps (4, or 8 for the
// K > 64 shapes: one producer warp per scheduler cannot hide its own instruction latency — the K = 128 TCN convolution spent
// 8.4 us per tile in 4.4 k dependent instructions per producer warp).  Warps [0, PW) produce, [PW, PW + 4) drain TMEM (lane
// group = warp & 3), warp PW + 4 issues the MMAs.
template <int KQM, int NSET, int PW = 4>
__global__ void __launch_bounds__(PW * 32 + 160) gemm_rows_tc_kernel(const GemmBatch gb, const TcRowsGeom geo) {
    constexpr int PT = PW * 32;
    const GemmArgs& g = gb.g[blockIdx.z];
    extern __shared__ __align__(128) unsigned char tsm[];
    const int S = geo.stages;
    const int ksp = g.ksplit > 1 ? g.ksplit : 1;                      // column blocks per matrix
    const int nkb = (g.nkb == 2 ? 2 : 1) * ksp;                       // K blocks per tile: (A | A2) x column block
    const int Kb = g.K / ksp;                                         // K of one block
    unsigned char* W_hi = tsm;                                        // [nkb] x (W_hi | W_lo)
    unsigned char* W_lo = W_hi + geo.w_bytes;
    unsigned char* A_base = tsm + (size_t)nkb * 2 * geo.w_bytes;      // S x (A_hi | A_lo)
    uint64_t* mbar = reinterpret_cast<uint64_t*>(A_base + (size_t)S * 2 * geo.a_bytes);
    // mbar[0..S) full, [S..2S) empty, [2S..2S+2) tmem full, [2S+2..2S+4) tmem empty
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 2 * S + 4);
    float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);          // [NP]
    float* epi_s = bias_s + geo.NP;                                     // [4 warps][32 rows][36]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KP = geo.KP, NP = geo.NP, KQ = KP >> 2;
    const int ntiles = (g.M + 127) / 128;
    if ((int)blockIdx.x >= ntiles) return;

    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), geo.tmem_cols);
    if (tid == 32) {
        for (int i = 0; i < S; i++) { mbar_init(smem_u32(mbar + i), PT); mbar_init(smem_u32(mbar + S + i), 1); }
        for (int i = 0; i < 2; i++) { mbar_init(smem_u32(mbar + 2 * S + i), 1); mbar_init(smem_u32(mbar + 2 * S + 2 + i), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // stage W (hi/lo) once: canonical K-major, rows n at 16 B, K-chunks at w_lbo
    for (int kb = 0; kb < nkb; kb++) {
        const float* Wp = (kb / ksp) ? g.W2 : g.W;
        const int k0 = (kb % ksp) * Kb;
        for (int i = tid; i < NP * KP; i += PT + 160) {
            int n, k;
            if (g.wT == 0) { n = i / KP; k = i % KP; } else { k = i / NP; n = i % NP; }
            float v = 0.f;
            if (n < g.N && k < Kb) v = gemm_w_at(g, Wp, n, k0 + k);
            float hi, lo;
            split_tf32_rn(v, hi, lo);
            uint32_t off = (uint32_t)kb * 2 * geo.w_bytes + (uint32_t)n * 16 + (uint32_t)(k >> 2) * geo.w_lbo + (k & 3) * 4;
            *reinterpret_cast<float*>(W_hi + off) = hi;
            *reinterpret_cast<float*>(W_lo + off) = lo;
        }
    }
    for (int i = tid; i < NP; i += PT + 160) bias_s[i] = (g.bias && i < g.N) ? __ldg(g.bias + i) : 0.f;
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_uniform(tmem_slot);
    const uint32_t bar_full = smem_u32(mbar), bar_empty = smem_u32(mbar + S);
    const uint32_t bar_tfull = smem_u32(mbar + 2 * S), bar_tempty = smem_u32(mbar + 2 * S + 2);

    if (warp < PW) {
        // ===================== producers =====================
        const bool split = g.A.mode == A_SPLIT, taps = g.A.mode == A_TAPS;
        const int items = (128 * KQ + PT - 1) / PT;                       // float4 items per producer thread and block
        float4 pre[NSET][KQM];
        // work item w = (local tile index) * nkb + kb.  Thread tid owns the float4 items i = tid + 128 j of a [128 x KQ] block:
        // (row, kq) = (i / KQ, i % KQ) advance by (128 / KQ, 128 % KQ) per j — tracked incrementally, no division per item
        // (the per-item divisions by runtime KQ / T / channel count were 190 instructions per float4: the producers of the
        // K = 128 TCN convolution issued 24 k warp instructions per tile and bound the kernel at 1.03 ms per 1.4 M rows).
        const int rstep = PT / KQ, kstep = PT - rstep * KQ;
        const int row_t = tid / KQ, kq_t = tid - row_t * KQ;
        // the operand description in registers: `g` is an element of a by-value kernel argument selected by blockIdx.z, every
        // access to it is an indexed constant load (LDC) — 880 of them per tile in the K = 128 convolution before this
        const float* const ap0 = g.A.p; const float* const ap1 = g.nkb == 2 ? g.A2.p : g.A.p;
        const int a_ld = g.A.ld, a_T = g.A.T, a_cc = g.A.cc, a_dil = g.A.dil, a_taps = g.A.taps, a_off = g.A.off, a_split = g.A.split, a_skip = g.A.skip;
        const int gM = g.M, n_grid = gridDim.x, bx = blockIdx.x;
        auto load_regs = [&](float4 (&r)[KQM], int w) {
            const int tile = bx + (w / nkb) * n_grid;
            const int kb = w % nkb;
            const float* const ap = (kb / ksp) ? ap1 : ap0;
            const int k0 = (kb % ksp) * Kb;
            const int m0 = tile * 128;
            int row = row_t, kq = kq_t;
            int t = taps ? (m0 + row) % a_T : 0, ch = 0, sh = 0;
            if (taps && kstep == 0) {           // KQ divides 128: this thread's column (tap, channel) never changes
                const int c0 = kq * 4 + k0, tj = c0 / a_cc;
                ch = c0 - tj * a_cc; sh = a_dil * (a_taps - 1 - tj) + a_off;
            }
#pragma unroll
            for (int j = 0; j < KQM; j++) {
                r[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j < items) {
                    const int m = m0 + row;
                    int c = kq * 4;
                    bool ok = row < 128 && m < gM && c < Kb;
                    const float* ptr;
                    c += k0;
                    if (taps) {                 // dilated causal convolution: column block tj reads the row `sh` steps away
                        if (kstep != 0) {
                            const int tj = c / a_cc;
                            ch = c - tj * a_cc; sh = a_dil * (a_taps - 1 - tj) + a_off;
                        }
                        ok = ok && (unsigned)(t + sh) < (unsigned)a_T;
                        ptr = ap + (long long)(m + sh) * a_ld + ch;
                    } else {
                        if (split && c >= a_split) c += a_skip;
                        ptr = ap + (size_t)m * a_ld + c;
                    }
                    if (ok) r[j] = __ldg(reinterpret_cast<const float4*>(ptr));
                    int dr = rstep;
                    row += rstep; kq += kstep;
                    if (kq >= KQ) { kq -= KQ; row++; dr++; }
                    if (taps) { t += dr; while (t >= a_T) t -= a_T; }
                }
            }
        };
        auto store_smem = [&](const float4 (&r)[KQM], int stage) {
            unsigned char* A_hi = A_base + (size_t)stage * 2 * geo.a_bytes;
            unsigned char* A_lo = A_hi + geo.a_bytes;
            int row = row_t, kq = kq_t;
#pragma unroll
            for (int j = 0; j < KQM; j++) {
                if (j < items && row < 128) {
                    float4 hi, lo;
                    split_tf32x4_rn(r[j], hi, lo);
                    uint32_t off = (uint32_t)row * 16 + (uint32_t)kq * TC_A_LBO;
                    *reinterpret_cast<float4*>(A_hi + off) = hi;
                    *reinterpret_cast<float4*>(A_lo + off) = lo;
                    row += rstep; kq += kstep;
                    if (kq >= KQ) { kq -= KQ; row++; }
                }
            }
        };
        const int mytiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const int nwork = mytiles * nkb;
        if (NSET == 2) load_regs(pre[0], 0);
        for (int it = 0; it < nwork; it++) {
            const int s = it % S;
            const uint32_t ph = (uint32_t)((it / S) & 1);
            if (NSET == 2) {
                // alternate register sets: set (it&1) holds this work item, set (it&1)^1 receives the next
                if ((it & 1) == 0) {
                    if (it + 1 < nwork) load_regs(pre[1 % NSET], it + 1);
                    mbar_wait(bar_empty + 8u * s, ph ^ 1u);
                    store_smem(pre[0], s);
                } else {
                    if (it + 1 < nwork) load_regs(pre[0], it + 1);
                    mbar_wait(bar_empty + 8u * s, ph ^ 1u);
                    store_smem(pre[1 % NSET], s);
                }
            } else {
                load_regs(pre[0], it);
                mbar_wait(bar_empty + 8u * s, ph ^ 1u);
                store_smem(pre[0], s);
            }
            fence_async_smem();
            mbar_arrive(bar_full + 8u * s);
        }
    } else if (warp < PW + 4) {
        // ===================== epilogue =====================
        // TMEM lane = row.  Each warp moves its 32 rows in 32-column chunks through a private padded
        // smem buffer so that every global access is a full 128 B line per row (4 rows / instruction).
        const int ew = warp & 3;
        float* stg = epi_s + ew * (32 * 36);
        const bool vec = ((g.ldc & 3) == 0) && (((uintptr_t)g.C & 15) == 0) &&
                         (g.mask == nullptr || ((g.ldmask & 3) == 0 && ((uintptr_t)g.mask & 15) == 0));
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
            const int a = it & 1;
            mbar_wait(bar_tfull + 8u * a, (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            const int mw = tile * 128 + ew * 32;
            const uint32_t trow = tmem + ((uint32_t)(ew * 32) << 16) + (uint32_t)a * (uint32_t)NP;
            for (int c0 = 0; c0 < NP; c0 += 32) {
                float v[16];
                tmem_ld16(trow + c0, v);
#pragma unroll
                for (int q = 0; q < 4; q++)
                    *reinterpret_cast<float4*>(stg + lane * 36 + q * 4) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                if (c0 + 16 < NP) {
                    tmem_ld16(trow + c0 + 16, v);
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        *reinterpret_cast<float4*>(stg + lane * 36 + 16 + q * 4) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                }
                __syncwarp();
                const int cq = lane & 7, n = c0 + cq * 4;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int r = i * 4 + (lane >> 3);
                    const int m = mw + r;
                    if (m < g.M && n < g.N) {
                        float4 o4 = *reinterpret_cast<const float4*>(stg + r * 36 + cq * 4);
                        float4 b4 = *reinterpret_cast<const float4*>(bias_s + n);
                        float o[4] = {o4.x + b4.x, o4.y + b4.y, o4.z + b4.z, o4.w + b4.w};
                        float* cp = g.C + (size_t)m * g.ldc + n;
                        if (vec && n + 3 < g.N) {
                            if (g.accum) {
                                float4 c = *reinterpret_cast<const float4*>(cp);
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
                            *reinterpret_cast<float4*>(cp) = make_float4(o[0], o[1], o[2], o[3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                if (n + j >= g.N) continue;
                                float x = o[j];
                                if (g.accum) x += cp[j];
                                if (g.relu) x = fmaxf(x, 0.f);
                                if (g.mask) x = g.mask[(size_t)m * g.ldmask + n + j] > 0.f ? x : 0.f;
                                cp[j] = x;
                            }
                        }
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            mbar_arrive(bar_tempty + 8u * a);
        }
    } else {
        // ===================== MMA issuer: warp 8 runs the loop converged, one elected lane issues =====================
        const uint32_t idesc = umma_idesc_tf32(NP, 0, 0);
        const uint32_t w_hi_s = smem_u32(W_hi), w_lo_s = smem_u32(W_lo);
        int it = 0, wk = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
            const int a = it & 1;
            mbar_wait(bar_tempty + 8u * a, (uint32_t)(((it >> 1) & 1) ^ 1));
            const uint32_t acc = tmem + (uint32_t)a * (uint32_t)NP;
            for (int kb = 0; kb < nkb; kb++, wk++) {
                const int s = wk % S;
                mbar_wait(bar_full + 8u * s, (uint32_t)((wk / S) & 1));
                tc_fence_after();
                const uint32_t a_hi_s = smem_u32(A_base + (size_t)s * 2 * geo.a_bytes), a_lo_s = a_hi_s + geo.a_bytes;
                const uint32_t wkb = (uint32_t)kb * 2 * geo.w_bytes;
                if (elect_one_sync()) {
                    // The TMEM accumulator TRUNCATES on every accumulate (an error of up to one ulp of its current value): the
                    // two correction terms of the 3xTF32 split (2^-11 of the result) go in first, while the accumulator is
                    // still small, so only the K/8 hi.hi accumulates cost precision (measured: 1.3e-6 -> relative error of a
                    // K = 128 product with the three terms interleaved).
                    for (int ks = 0; ks < (KP >> 3); ks++) {
                        uint32_t ao = (uint32_t)ks * 2 * TC_A_LBO, wo = wkb + (uint32_t)ks * 2 * geo.w_lbo;
                        uint64_t dah = umma_desc(a_hi_s + ao, TC_A_LBO, 128), dal = umma_desc(a_lo_s + ao, TC_A_LBO, 128);
                        uint64_t dbh = umma_desc(w_hi_s + wo, geo.w_lbo, 128), dbl = umma_desc(w_lo_s + wo, geo.w_lbo, 128);
                        umma_tf32(acc, dal, dbh, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
                        umma_tf32(acc, dah, dbl, idesc, 1u);
                    }
                    for (int ks = 0; ks < (KP >> 3); ks++) {
                        uint32_t ao = (uint32_t)ks * 2 * TC_A_LBO, wo = wkb + (uint32_t)ks * 2 * geo.w_lbo;
                        uint64_t dah = umma_desc(a_hi_s + ao, TC_A_LBO, 128);
                        uint64_t dbh = umma_desc(w_hi_s + wo, geo.w_lbo, 128);
                        umma_tf32(acc, dah, dbh, idesc, 1u);
                    }
                    umma_commit(bar_empty + 8u * s);    // smem stage may be refilled once these MMAs retire
                    if (kb == nkb - 1) umma_commit(bar_tfull + 8u * a);        // accumulator ready for the epilogue warps
                }
                __syncwarp();
            }
        }
    }
    tc_fence_before();
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

#define TC_SMEM_MAX (200 * 1024)

static bool tc_rows_geom(int N, int K, TcRowsGeom& geo, size_t& smem, int nkb = 1, int max_stages = 2) {
    geo.KP = round_up(K, 8);
    geo.NP = round_up(N, 16);
    if (2 * geo.NP > 512) return false;
    geo.tmem_cols = tmem_cols_for(2 * geo.NP);
    geo.w_lbo = geo.NP * 16 + 16;
    geo.a_bytes = (uint32_t)(geo.KP / 4) * TC_A_LBO;
    geo.w_bytes = (uint32_t)(geo.KP / 4) * geo.w_lbo;
    for (int S = max_stages; S >= 1; S--) {
        smem = (size_t)nkb * 2 * geo.w_bytes + (size_t)S * 2 * geo.a_bytes + (2 * S + 4) * 8 + 16 + (size_t)geo.NP * 4 +
               4 * 32 * 36 * 4 + 128;
        if (smem <= TC_SMEM_MAX) { geo.stages = S; return true; }
    }
    return false;
}

static bool tc_rows_eligible(const GemmArgs& g) {
    if (g.M < 2048 || g.N > 256 || g.K > 128 * (g.ksplit > 1 ? g.ksplit : 1) || g.N < 8) return false;
    if (g.A.mode != A_PLAIN && g.A.mode != A_SPLIT && g.A.mode != A_TAPS) return false;
    if ((g.A.ld & 3) || (g.K & 3) || !aligned16(g.A.p)) return false;
    if (g.A.mode == A_SPLIT && ((g.A.split & 3) || (g.A.skip & 3))) return false;
    if (g.A.mode == A_TAPS && ((g.A.cc & 3) || g.nkb == 2)) return false;
    if (!aligned16(g.C) || (g.mask && !aligned16(g.mask))) return false;
    if (g.nkb == 2) {
        if (g.A2.mode != g.A.mode || g.A2.ld != g.A.ld || g.A2.split != g.A.split || g.A2.skip != g.A.skip) return false;
        if (!aligned16(g.A2.p) || !g.W2) return false;
    }
    const int ksp = g.ksplit > 1 ? g.ksplit : 1;
    if (g.K % ksp || ((g.K / ksp) & 3)) return false;
    TcRowsGeom geo; size_t smem;
    return tc_rows_geom(g.N, g.K / ksp, geo, smem, (g.nkb == 2 ? 2 : 1) * ksp);
}

template <int KQM, int NSET, int PW = 4>
static int launch_rows_tc_t(const GemmBatch& gb, const TcRowsGeom& geo, size_t smem, dim3 grid, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        DOF_CUDA(cudaFuncSetAttribute(gemm_rows_tc_kernel<KQM, NSET, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_MAX));
        attr = true;
    }
    gemm_rows_tc_kernel<KQM, NSET, PW><<<grid, PW * 32 + 160, smem, st>>>(gb, geo);
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

static int launch_gemm_rows_tc(const GemmArgs* gs, int nbatch, cudaStream_t st, int sm_count) {
    GemmBatch gb;
    memset(&gb, 0, sizeof(gb));
    int M = 0;
    for (int i = 0; i < nbatch; i++) {
        gb.g[i] = gs[i];
        if (gs[i].M > M) M = gs[i].M;
        if (gs[i].N != gs[0].N || gs[i].K != gs[0].K || (gs[i].nkb == 2) != (gs[0].nkb == 2))
            DOF_FAIL(DOF_ERR_ARG, "batched TC GEMMs must share N, K and the K-block count");
    }
    TcRowsGeom geo;
    size_t smem = 0;
    const int nmat = gs[0].nkb == 2 ? 2 : 1;
    // ksplit > 1 stages every matrix in column blocks (smaller stages, register prefetch).  Measured on B200 for the
    // K = 96 / N = 32 input-gradient GEMM: two 48-column blocks were 11 % SLOWER than one 96-column stage (the
    // per-stage barrier round trips outweigh the overlap), so it is opt-in only.
    int ksp = gs[0].ksplit > 1 ? gs[0].ksplit : 1;
    for (int i = 0; i < nbatch; i++) gb.g[i].ksplit = ksp;
    const int nkb = nmat * ksp;
    if (!tc_rows_geom(gs[0].N, gs[0].K / ksp, geo, smem, nkb)) DOF_FAIL(DOF_ERR_UNSUPPORTED, "TC GEMM tile does not fit");
    const int KQ = geo.KP / 4;
    // 8 < K/4 <= 16 (the transformer's K = 40 projections): two register sets of 16 float4 per producer thread cap the CTA at one
    // per SM.  One stage + one register set fits TWO CTAs per SM (all roles doubled), which hides the producers' load -> split
    // -> store latency better than the deeper pipeline of a single CTA (DOF_ROWS_OCC2=0 restores the old choice for A/B runs).
    static const bool occ2_on = !(getenv("DOF_ROWS_OCC2") && getenv("DOF_ROWS_OCC2")[0] == '0');
    bool occ2 = false;
    if (occ2_on && KQ > 8 && KQ <= 16) {
        TcRowsGeom g1; size_t sm1 = 0;
        if (tc_rows_geom(gs[0].N, gs[0].K / ksp, g1, sm1, nkb, 1) && 2 * (sm1 + 1024) <= 228 * 1024 && 2 * g1.tmem_cols <= 512) {
            geo = g1; smem = sm1; occ2 = true;
        }
    }
    int occ = (int)((228 * 1024) / (smem + 1024));
    int by_tmem = 512 / geo.tmem_cols;
    if (occ > by_tmem) occ = by_tmem;
    int by_regs = (KQ <= 8 || occ2) ? 2 : 1;           // register budget of the 288-thread CTA
    if (occ > by_regs) occ = by_regs;
    if (occ < 1) occ = 1;
    int ntiles = cdiv(M, 128);
    int ctas = sm_count * occ / nbatch;
    if (ctas < 1) ctas = 1;
    if (ctas > ntiles) ctas = ntiles;
    double fl = 0.0, by = 0.0;
    for (int i = 0; i < nbatch; i++) {
        fl += 2.0 * gs[i].M * gs[i].N * gs[i].K * nmat;
        by += 4.0 * gs[i].M * ((double)gs[i].K * nmat + (double)gs[i].N * (1 + (gs[i].accum ? 1 : 0) + (gs[i].mask ? 1 : 0)));
    }
    ProfScope ps("gemm_rows_tc", st, fl, by);
    dim3 grid(ctas, 1, nbatch);
    static const int pws = getenv("DOF_ROWS_PW_SMALL") ? atoi(getenv("DOF_ROWS_PW_SMALL")) : 8;
    if (pws == 8) {
        if (KQ <= 8) return launch_rows_tc_t<4, 2, 8>(gb, geo, smem, grid, st);
        if (KQ <= 16 && occ2) return launch_rows_tc_t<8, 1, 8>(gb, geo, smem, grid, st);
        if (KQ <= 16) return launch_rows_tc_t<8, 2, 8>(gb, geo, smem, grid, st);
    }
    if (KQ <= 8) return launch_rows_tc_t<8, 2>(gb, geo, smem, grid, st);
    if (KQ <= 16 && occ2) return launch_rows_tc_t<16, 1>(gb, geo, smem, grid, st);
    if (KQ <= 16) return launch_rows_tc_t<16, 2>(gb, geo, smem, grid, st);
    // K > 64: eight producer warps with 16 items each (DOF_ROWS_PW8=0 restores four warps with 32 items for A/B runs)
    static const int pw = getenv("DOF_ROWS_PW") ? atoi(getenv("DOF_ROWS_PW")) : 16;
    if (pw == 16) return launch_rows_tc_t<8, 1, 16>(gb, geo, smem, grid, st);
    if (pw == 12) return launch_rows_tc_t<11, 1, 12>(gb, geo, smem, grid, st);
    if (pw == 8) return launch_rows_tc_t<16, 1, 8>(gb, geo, smem, grid, st);
    return launch_rows_tc_t<32, 1>(gb, geo, smem, grid, st);
}

// ---------------------------------------------------------------------------
// gemm_wgrad on tcgen05:  D[n, k] = sum_m P[m,n] Q[m,k]  accumulated in TMEM over the CTA's rows.
// Both operands are K-major with the reduction index m as the MMA K dimension: the row-major
// [m, n] / [m, k] tiles are TRANSPOSED while they are staged (float4 global loads, 4 scalar
// shared stores per float4).  Bank-conflict-free staging: a warp covers 8 rows x 4 float4
// (64 B contiguous per row -> full sectors), the K-chunk stride LBO = 144 B (== 4 words mod 32)
// and the 8-row-group stride SBO == 8 words mod 32.  A constant-one row appended to Q^T yields
// the bias gradient.  Rows of A beyond N (and of B beyond K+1) hold don't-care data: they only
// reach accumulator rows / columns that are never read.
// ---------------------------------------------------------------------------
#define TCW_BM 32                      // rows (MMA-K) per stage
#define TCW_LBO 144
#define TCW_SBO ((TCW_BM / 4) * TCW_LBO + 32)   // 1184 B
#define TCW_MAXPRE 6                   // float4 prefetch registers per producer thread and set
#define TCW_PW 8                       // producer warps
#define TCW_THREADS 288                // warps 0-7 producers (0-3 also run the final epilogue), warp 8 MMA issuer
#define TCW_STAGES 2

struct TcWgradGeom { int KWP, tmem_cols, ptasks, qtasks; uint32_t p_bytes, q_bytes; };

__global__ void __launch_bounds__(TCW_THREADS, 2) gemm_wgrad_tc_kernel(const WGradBatch wb, const TcWgradGeom geo) {
    const WGradArgs& g = wb.g[blockIdx.z];
    extern __shared__ __align__(128) unsigned char tsm[];
    const uint32_t stage_bytes = 2 * geo.p_bytes + 2 * geo.q_bytes;     // P_hi | P_lo | Q_hi | Q_lo
    uint64_t* mbar = reinterpret_cast<uint64_t*>(tsm + (size_t)TCW_STAGES * stage_bytes);
    // mbar[0..S) full (128 arrivals), [S..2S) empty (1 commit), [2S] all MMAs done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 2 * TCW_STAGES + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int rows_per = (g.M + gridDim.x - 1) / gridDim.x;
    rows_per = (rows_per + TCW_BM - 1) / TCW_BM * TCW_BM;
    const int mbeg = blockIdx.x * rows_per;
    const int mend = min(g.M, mbeg + rows_per);
    if (mbeg >= mend) return;
    const int nstage = (mend - mbeg + TCW_BM - 1) / TCW_BM;

    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), geo.tmem_cols);
    if (tid == 32) {
        for (int i = 0; i < TCW_STAGES; i++) { mbar_init(smem_u32(mbar + i), TCW_PW * 32); mbar_init(smem_u32(mbar + TCW_STAGES + i), 1); }
        mbar_init(smem_u32(mbar + 2 * TCW_STAGES), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // zero the operand tiles once (padding rows stay finite), then the constant-one row k = K of Q^T
    for (uint32_t i = tid * 16; i < TCW_STAGES * stage_bytes; i += TCW_THREADS * 16)
        *reinterpret_cast<float4*>(tsm + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    if (tid < TCW_BM * TCW_STAGES) {
        int sidx = tid / TCW_BM, mm = tid % TCW_BM;
        uint32_t off = (uint32_t)(g.K >> 3) * TCW_SBO + (uint32_t)(g.K & 7) * 16 + (uint32_t)(mm >> 2) * TCW_LBO + (mm & 3) * 4;
        *reinterpret_cast<float*>(tsm + (size_t)sidx * stage_bytes + 2 * geo.p_bytes + off) = 1.0f;
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_uniform(tmem_slot);
    const uint32_t bar_full = smem_u32(mbar), bar_empty = smem_u32(mbar + TCW_STAGES), bar_done = smem_u32(mbar + 2 * TCW_STAGES);

    if (warp < TCW_PW) {
        // ===================== producers =====================
        const int NQ = g.N >> 2, KQ = g.K >> 2;
        const bool psplit = g.P.mode == A_SPLIT;
        const bool qshift = g.Q.mode == A_TSHIFT, qtaps = g.Q.mode == A_TAPS;
        const int r8 = lane & 7, c4 = lane >> 3;
        const int ntask = geo.ptasks + geo.qtasks;
        // task t: operand (P if t < ptasks), 8-row block mblk = t % (BM/8), 4-float4 column block cb
        float4 pre[2][TCW_MAXPRE];
        // scalars of the operand description in registers (g is an indexed element of the by-value argument: every access is
        // an LDC); step index and tap of a time-shifted Q item by multiply-high with precomputed reciprocals instead of the
        // ~20-instruction runtime division (exact for m < 2^32 / T and c < 2^32 / cc — rows are int, T and cc are small)
        // (only the two reciprocals live in registers: the kernel sits at its register budget for two CTAs per SM — keeping the
        // whole operand description in registers spilled and cost the plain weight gradients of the transformer 65 %)
        const uint32_t magic_T = g.Q.T > 0 ? (uint32_t)((0x100000000ull + (uint32_t)g.Q.T - 1) / (uint32_t)g.Q.T) : 0u;
        const uint32_t magic_c = g.Q.cc > 0 ? (uint32_t)((0x100000000ull + (uint32_t)g.Q.cc - 1) / (uint32_t)g.Q.cc) : 0u;
        auto load_regs = [&](float4 (&r)[TCW_MAXPRE], int mb0) {
#pragma unroll
            for (int j = 0; j < TCW_MAXPRE; j++) {
                int t = warp + TCW_PW * j;
                r[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t < ntask) {
                    const bool isP = t < geo.ptasks;
                    int tt = isP ? t : t - geo.ptasks;
                    int mblk = tt % (TCW_BM / 8), cb = tt / (TCW_BM / 8);
                    int m = mb0 + mblk * 8 + r8, cq = cb * 4 + c4;
                    if (m < mend) {
                        if (isP) {
                            if (cq < NQ) {
                                int c = cq * 4;
                                if (psplit && c >= g.P.split) c += g.P.skip;
                                r[j] = __ldg(reinterpret_cast<const float4*>(g.P.p + (size_t)m * g.P.ld + c));
                            }
                        } else if (cq < KQ) {
                            if (qtaps) {            // all taps of a dilated causal convolution in one operand (see mv_taps)
                                const int c = cq * 4, tj = (int)__umulhi((uint32_t)c, magic_c), ch = c - tj * g.Q.cc, sh = g.Q.dil * (g.Q.taps - 1 - tj) + g.Q.off;
                                const int t2 = m - (int)__umulhi((uint32_t)m, magic_T) * g.Q.T + sh;
                                if (t2 >= 0 && t2 < g.Q.T)
                                    r[j] = __ldg(reinterpret_cast<const float4*>(g.Q.p + (long long)(m + sh) * g.Q.ld + ch));
                            } else if (qshift) {
                                const int t2 = m - (int)__umulhi((uint32_t)m, magic_T) * g.Q.T + g.Q.shift;
                                if (t2 >= 0 && t2 < g.Q.T)
                                    r[j] = __ldg(reinterpret_cast<const float4*>(g.Q.p + (long long)(m + g.Q.shift) * g.Q.ld + cq * 4));
                            } else {
                                r[j] = __ldg(reinterpret_cast<const float4*>(g.Q.p + (size_t)m * g.Q.ld + cq * 4));
                            }
                        }
                    }
                }
            }
        };
        auto store_smem = [&](const float4 (&r)[TCW_MAXPRE], int stage) {
            unsigned char* P_hi = tsm + (size_t)stage * stage_bytes;
            unsigned char* P_lo = P_hi + geo.p_bytes;
            unsigned char* Q_hi = P_lo + geo.p_bytes;
            unsigned char* Q_lo = Q_hi + geo.q_bytes;
#pragma unroll
            for (int j = 0; j < TCW_MAXPRE; j++) {
                int t = warp + TCW_PW * j;
                if (t < ntask) {
                    const bool isP = t < geo.ptasks;
                    int tt = isP ? t : t - geo.ptasks;
                    int mblk = tt % (TCW_BM / 8), cb = tt / (TCW_BM / 8);
                    int ml = mblk * 8 + r8, cq = cb * 4 + c4;
                    if (cq < (isP ? NQ : KQ)) {
                        float4 hi, lo;
                        split_tf32x4_rn(r[j], hi, lo);
                        // rows 4cq..4cq+3 of the transposed tile: 8-row group cq>>1, row-in-group 4*(cq&1)+e
                        uint32_t off = (uint32_t)(cq >> 1) * TCW_SBO + (uint32_t)(4 * (cq & 1)) * 16 + (uint32_t)(ml >> 2) * TCW_LBO + (ml & 3) * 4;
                        unsigned char* bh = (isP ? P_hi : Q_hi) + off;
                        unsigned char* bl = (isP ? P_lo : Q_lo) + off;
                        *reinterpret_cast<float*>(bh) = hi.x; *reinterpret_cast<float*>(bh + 16) = hi.y;
                        *reinterpret_cast<float*>(bh + 32) = hi.z; *reinterpret_cast<float*>(bh + 48) = hi.w;
                        *reinterpret_cast<float*>(bl) = lo.x; *reinterpret_cast<float*>(bl + 16) = lo.y;
                        *reinterpret_cast<float*>(bl + 32) = lo.z; *reinterpret_cast<float*>(bl + 48) = lo.w;
                    }
                }
            }
        };
        load_regs(pre[0], mbeg);
        for (int it = 0; it < nstage; it++) {
            const int s = it % TCW_STAGES;
            const uint32_t ph = (uint32_t)((it / TCW_STAGES) & 1);
            const int mnext = mbeg + (it + 1) * TCW_BM;
            if ((it & 1) == 0) {
                if (it + 1 < nstage) load_regs(pre[1], mnext);
                mbar_wait(bar_empty + 8u * s, ph ^ 1u);
                store_smem(pre[0], s);
            } else {
                if (it + 1 < nstage) load_regs(pre[0], mnext);
                mbar_wait(bar_empty + 8u * s, ph ^ 1u);
                store_smem(pre[1], s);
            }
            fence_async_smem();
            mbar_arrive(bar_full + 8u * s);
        }
        // ---- epilogue (warps 0-3): thread = output row n; atomics into dW / db
        if (warp < 4) {
        mbar_wait(bar_done, 0u);
        tc_fence_after();
        const int n = warp * 32 + lane;
        const int nvl = g.nv > 0 ? g.nv : g.N, kvl = g.kv > 0 ? g.kv : g.K;
        const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < geo.KWP; c0 += 16) {
            float v[16];
            tmem_ld16(trow + c0, v);
            if (n < nvl) {
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    int k = c0 + j;
                    if (k < kvl) {
                        float* o = g.oT ? g.dW + (size_t)k * g.ldo + n : g.dW + (size_t)n * g.ldo + (size_t)k * (g.ks > 0 ? g.ks : 1);
                        if (g.otaps > 0) {              // Conv1d weight [N, ocv, otaps] from an A_TAPS operand (column = tap * cc + channel)
                            const int cc = g.K / g.otaps, tj = k / cc, ch = k - tj * cc;
                            if (ch >= g.ocv) continue;
                            o = g.dW + (size_t)n * g.ldo + ch * g.otaps + tj;
                        }
                        atomicAdd(o, v[j]);
                    } else if (k == g.K && g.db) {
                        atomicAdd(g.db + n, v[j]);
                    }
                }
            }
        }
        }
    } else {
        // ===================== MMA issuer: warp 8 runs the loop converged, one elected lane issues =====================
        const uint32_t idesc = umma_idesc_tf32(geo.KWP, 0, 0);
        for (int it = 0; it < nstage; it++) {
            const int s = it % TCW_STAGES;
            mbar_wait(bar_full + 8u * s, (uint32_t)((it / TCW_STAGES) & 1));
            tc_fence_after();
            const uint32_t p_hi_s = smem_u32(tsm + (size_t)s * stage_bytes), p_lo_s = p_hi_s + geo.p_bytes;
            const uint32_t q_hi_s = p_lo_s + geo.p_bytes, q_lo_s = q_hi_s + geo.q_bytes;
            if (elect_one_sync()) {
#pragma unroll
                for (int kb = 0; kb < TCW_BM / 8; kb++) {
                    uint32_t o = (uint32_t)kb * 2 * TCW_LBO;
                    uint64_t dah = umma_desc(p_hi_s + o, TCW_LBO, TCW_SBO), dal = umma_desc(p_lo_s + o, TCW_LBO, TCW_SBO);
                    uint64_t dbh = umma_desc(q_hi_s + o, TCW_LBO, TCW_SBO), dbl = umma_desc(q_lo_s + o, TCW_LBO, TCW_SBO);
                    umma_tf32(tmem, dah, dbh, idesc, (it == 0 && kb == 0) ? 0u : 1u);
                    umma_tf32(tmem, dal, dbh, idesc, 1u);
                    umma_tf32(tmem, dah, dbl, idesc, 1u);
                }
                umma_commit(bar_empty + 8u * s);
                if (it == nstage - 1) umma_commit(bar_done);
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, geo.tmem_cols);
}

static void tc_wgrad_geom(int N, int K, TcWgradGeom& geo) {
    geo.KWP = round_up(K + 1, 16);
    geo.tmem_cols = tmem_cols_for(geo.KWP);
    geo.ptasks = (TCW_BM / 8) * cdiv(N >> 2, 4);
    geo.qtasks = (TCW_BM / 8) * cdiv(K >> 2, 4);
    geo.p_bytes = 16u * TCW_SBO;                          // 128 rows of P^T (MMA M = 128)
    geo.q_bytes = (uint32_t)(geo.KWP / 8) * TCW_SBO;
}

static size_t tc_wgrad_smem(const TcWgradGeom& geo) {
    return (size_t)TCW_STAGES * (2 * (size_t)geo.p_bytes + 2 * (size_t)geo.q_bytes) + (2 * TCW_STAGES + 1) * 8 + 16 + 128;
}

static bool tc_wgrad_eligible(const WGradArgs& g) {
    if (g.M < 4096 || g.N > 128 || g.K + 1 > 256 || g.N < 4 || g.K < 4) return false;
    if ((g.N & 3) || (g.K & 3)) return false;
    if (g.P.mode != A_PLAIN && g.P.mode != A_SPLIT) return false;
    if (g.Q.mode != A_PLAIN && g.Q.mode != A_TSHIFT && g.Q.mode != A_TAPS) return false;
    if (g.Q.mode == A_TAPS && (g.Q.cc & 3)) return false;
    // the producers divide by T with a multiply-high reciprocal: exact while M * T < 2^32
    if ((g.Q.mode == A_TAPS || g.Q.mode == A_TSHIFT) && (g.Q.T < 1 || (long long)g.M * g.Q.T >= (1ll << 32))) return false;
    if ((g.P.ld & 3) || (g.Q.ld & 3) || !aligned16(g.P.p) || !aligned16(g.Q.p)) return false;
    if (g.P.mode == A_SPLIT && ((g.P.split & 3) || (g.P.skip & 3))) return false;
    TcWgradGeom geo;
    tc_wgrad_geom(g.N, g.K, geo);
    if (cdiv(geo.ptasks + geo.qtasks, TCW_PW) > TCW_MAXPRE) return false;
    if (tc_wgrad_smem(geo) > TC_SMEM_MAX) return false;
    return true;
}

static int launch_gemm_wgrad_tc(const WGradArgs* gs, int nbatch, cudaStream_t st, int sm_count) {
    WGradBatch wb;
    memset(&wb, 0, sizeof(wb));
    int M = 0;
    if (nbatch > WG_MAXBATCH) DOF_FAIL(DOF_ERR_ARG, "at most %d batched TC wgrads", WG_MAXBATCH);
    for (int i = 0; i < nbatch; i++) {
        wb.g[i] = gs[i];
        if (gs[i].M > M) M = gs[i].M;
        if (gs[i].N != gs[0].N || gs[i].K != gs[0].K) DOF_FAIL(DOF_ERR_ARG, "batched TC wgrads must share N and K");
    }
    TcWgradGeom geo;
    tc_wgrad_geom(gs[0].N, gs[0].K, geo);
    size_t smem = tc_wgrad_smem(geo);
    static bool attr = false;
    if (!attr) {
        DOF_CUDA(cudaFuncSetAttribute(gemm_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_MAX));
        attr = true;
    }
    if (smem > TC_SMEM_MAX) DOF_FAIL(DOF_ERR_UNSUPPORTED, "TC wgrad tile does not fit shared memory");
    int occ = (int)((228 * 1024) / (smem + 1024));
    int by_tmem = 512 / geo.tmem_cols;
    if (occ > by_tmem) occ = by_tmem;
    if (occ > 2) occ = 2;
    if (occ < 1) occ = 1;
    int ctas = sm_count * occ / nbatch;
    int maxsplit = cdiv(M, 8 * TCW_BM);
    if (ctas > maxsplit) ctas = maxsplit;
    if (ctas < 1) ctas = 1;
    double fl = 0.0, by = 0.0;
    for (int i = 0; i < nbatch; i++) {
        fl += 2.0 * gs[i].M * gs[i].N * gs[i].K;
        by += 4.0 * gs[i].M * ((double)gs[i].N + gs[i].K);
    }
    ProfScope ps("gemm_wgrad_tc", st, fl, by);
    gemm_wgrad_tc_kernel<<<dim3(ctas, 1, nbatch), TCW_THREADS, smem, st>>>(wb, geo);
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

// ---------------------------------------------------------------------------
// dispatch: tensor-core path when every problem of the batch is eligible, SIMT otherwise
// ---------------------------------------------------------------------------
static int g_sm_count = 148;
static bool tc_big_eligible(const GemmArgs& g);                                  // tc_gemm_big.cuh
static int launch_gemm_big_tc(const GemmArgs& g, cudaStream_t st, int sm_count);

static int launch_gemm_rows(const GemmArgs* gs, int nbatch, cudaStream_t st) {
    bool tc = tc_enabled();
    for (int i = 0; i < nbatch && tc; i++)
        tc = tc_rows_eligible(gs[i]) && gs[i].N == gs[0].N && gs[i].K == gs[0].K;
    if (tc) return launch_gemm_rows_tc(gs, nbatch, st, g_sm_count);
    // shapes whose weights do not fit the weight-resident kernel: both operands streamed (tc_gemm_big.cuh)
    bool big = tc_enabled();
    for (int i = 0; i < nbatch && big; i++) big = tc_big_eligible(gs[i]);
    if (big) {
        for (int i = 0; i < nbatch; i++) DOF_TRY(launch_gemm_big_tc(gs[i], st, g_sm_count));
        return DOF_OK;
    }
    bool any2 = false;
    for (int i = 0; i < nbatch; i++) any2 = any2 || gs[i].nkb == 2;
    if (!any2) return launch_gemm_rows_simt(gs, nbatch, st);
    for (int i = 0; i < nbatch; i++) {
        if (gs[i].nkb != 2) { DOF_TRY(launch_gemm_rows_simt(gs + i, 1, st)); continue; }
        // not eligible for the tensor-core kernel: two passes, the second accumulating onto the first
        GemmArgs a = gs[i], b = gs[i];
        a.nkb = 0; a.mask = nullptr; a.relu = 0;
        b.nkb = 0; b.A = gs[i].A2; b.W = gs[i].W2; b.bias = nullptr; b.accum = 1;
        DOF_TRY(launch_gemm_rows(&a, 1, st));
        DOF_TRY(launch_gemm_rows(&b, 1, st));
    }
    return DOF_OK;
}

static int launch_gemm_wgrad(const WGradArgs* gs, int nbatch, cudaStream_t st, int sm_count) {
    bool tc = tc_enabled();
    for (int i = 0; i < nbatch && tc; i++)
        tc = tc_wgrad_eligible(gs[i]) && gs[i].N == gs[0].N && gs[i].K == gs[0].K;
    if (tc) return launch_gemm_wgrad_tc(gs, nbatch, st, sm_count);
    return launch_gemm_wgrad_simt(gs, nbatch, st, sm_count);
}

#include "tc_gemm_big.cuh"
