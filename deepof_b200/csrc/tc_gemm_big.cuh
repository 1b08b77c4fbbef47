// tcgen05 GEMM for the shapes the weight-resident kernel of tc_gemm.cuh cannot hold in shared memory (the transformer
// decoder at model_dim 256, the encoder head with K = (N+E)*D): C[M, N] = epi(A[M, K] . W^T + b), any K (multiple of 4), any N.
// One CTA owns a 128 x 128 output tile at a time (persistent over tiles, the N tiles of one row tile adjacent in the
// schedule so A is re-read from L2); BOTH operands stream through a shared-memory ring in K blocks of 32 columns:
// 4 producer warps load fp32 with float4, split every value into tf32 hi / lo (3xTF32, fp32-class accuracy) and store the
// canonical no-swizzle K-major core-matrix layout; one thread issues the tcgen05.mma (hi.hi + lo.hi + hi.lo) into a
// double-buffered TMEM accumulator; 4 epilogue warps drain it (bias / ReLU / accumulate / mask) while the next tile's
// MMAs run.  Hand-offs are mbarriers only.
#pragma once

#define TCB_KB 32                      // K columns per stage
#define TCB_NT 128                     // output columns per tile
#define TCB_STAGES 2
#define TCB_THREADS 288                // warps 0-3 producers, 4-7 epilogue, 8 MMA issuer
#define TCB_W_LBO (TCB_NT * 16 + 16)
#define TCB_A_BYTES ((TCB_KB / 4) * TC_A_LBO)
#define TCB_W_BYTES ((TCB_KB / 4) * TCB_W_LBO)
#define TCB_STAGE_BYTES (2 * TCB_A_BYTES + 2 * TCB_W_BYTES)

static inline size_t tc_big_smem() {
    return (size_t)TCB_STAGES * TCB_STAGE_BYTES + (2 * TCB_STAGES + 4) * 8 + 16 + TCB_NT * 4 + 4 * 32 * 36 * 4 + 128;
}

__global__ void __launch_bounds__(TCB_THREADS, 1) gemm_big_tc_kernel(const GemmArgs g, const int vecW) {
    extern __shared__ __align__(128) unsigned char bsm[];
    unsigned char* ring = bsm;                                                  // stages x (A_hi | A_lo | W_hi | W_lo)
    uint64_t* mbar = reinterpret_cast<uint64_t*>(ring + (size_t)TCB_STAGES * TCB_STAGE_BYTES);
    // mbar[0..S) full, [S..2S) empty, [2S..2S+2) tmem full, [2S+2..2S+4) tmem empty
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 2 * TCB_STAGES + 4);
    float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);                    // [2][TCB_NT]? one tile at a time per accumulator
    float* epi_s = bias_s + TCB_NT;                                             // [4 warps][32 rows][36]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = TCB_STAGES;
    const int mtiles = (g.M + 127) / 128, ntiles = (g.N + TCB_NT - 1) / TCB_NT, tiles = mtiles * ntiles;
    const int nkb = (g.K + TCB_KB - 1) / TCB_KB;
    if ((int)blockIdx.x >= tiles) return;

    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 2 * TCB_NT);
    if (tid == 32) {
        for (int i = 0; i < S; i++) { mbar_init(smem_u32(mbar + i), 128); mbar_init(smem_u32(mbar + S + i), 1); }
        for (int i = 0; i < 2; i++) { mbar_init(smem_u32(mbar + 2 * S + i), 1); mbar_init(smem_u32(mbar + 2 * S + 2 + i), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_uniform(tmem_slot);
    const uint32_t bar_full = smem_u32(mbar), bar_empty = smem_u32(mbar + S);
    const uint32_t bar_tfull = smem_u32(mbar + 2 * S), bar_tempty = smem_u32(mbar + 2 * S + 2);

    if (warp < 4) {
        // ===================== producers =====================
        float4 ra[TCB_KB / 4], rw[TCB_KB / 4];
        // A block: 128 rows x 8 float4; thread t takes float4 index i = t + 128 j: row = i / 8, kq = i % 8
        auto load_block = [&](int tile, int kb) {
            const int m0 = (tile / ntiles) * 128, n0 = (tile % ntiles) * TCB_NT, k0 = kb * TCB_KB;
#pragma unroll
            for (int j = 0; j < TCB_KB / 4; j++) {
                const int i = tid + j * 128, row = i >> 3, kq = i & 7;
                const int m = m0 + row, k = k0 + kq * 4;
                ra[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m < g.M && k < g.K) ra[j] = __ldg(reinterpret_cast<const float4*>(g.A.p + (size_t)m * g.A.ld + k));
            }
            if (g.wT == 0) {                       // W[n * ldw + k]: float4 along k
#pragma unroll
                for (int j = 0; j < TCB_KB / 4; j++) {
                    const int i = tid + j * 128, nr = i >> 3, kq = i & 7;
                    const int n = n0 + nr, k = k0 + kq * 4;
                    rw[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (n < g.N && k < g.K) {
                        const float* p = g.W + (size_t)n * g.ldw + k;
                        if (vecW) rw[j] = __ldg(reinterpret_cast<const float4*>(p));
                        else rw[j] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
                    }
                }
            } else {                               // W[k * ldw + n]: float4 along n (4 consecutive output columns at one k)
#pragma unroll
                for (int j = 0; j < TCB_KB / 4; j++) {
                    const int i = tid + j * 128, kk = i >> 5, nq = i & 31;
                    const int n = n0 + nq * 4, k = k0 + kk;
                    rw[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (k < g.K && n < g.N) {
                        const float* p = g.W + (size_t)k * g.ldw + n;
                        if (vecW && n + 3 < g.N) rw[j] = __ldg(reinterpret_cast<const float4*>(p));
                        else {
                            rw[j].x = __ldg(p);
                            if (n + 1 < g.N) rw[j].y = __ldg(p + 1);
                            if (n + 2 < g.N) rw[j].z = __ldg(p + 2);
                            if (n + 3 < g.N) rw[j].w = __ldg(p + 3);
                        }
                    }
                }
            }
        };
        auto store_block = [&](int stage) {
            unsigned char* A_hi = ring + (size_t)stage * TCB_STAGE_BYTES;
            unsigned char* A_lo = A_hi + TCB_A_BYTES;
            unsigned char* W_hi = A_lo + TCB_A_BYTES;
            unsigned char* W_lo = W_hi + TCB_W_BYTES;
#pragma unroll
            for (int j = 0; j < TCB_KB / 4; j++) {
                const int i = tid + j * 128, row = i >> 3, kq = i & 7;
                float4 hi, lo;
                split_tf32x4(ra[j], hi, lo);
                const uint32_t off = (uint32_t)row * 16 + (uint32_t)kq * TC_A_LBO;
                *reinterpret_cast<float4*>(A_hi + off) = hi;
                *reinterpret_cast<float4*>(A_lo + off) = lo;
            }
            if (g.wT == 0) {
#pragma unroll
                for (int j = 0; j < TCB_KB / 4; j++) {
                    const int i = tid + j * 128, nr = i >> 3, kq = i & 7;
                    float4 hi, lo;
                    split_tf32x4(rw[j], hi, lo);
                    const uint32_t off = (uint32_t)nr * 16 + (uint32_t)kq * TCB_W_LBO;
                    *reinterpret_cast<float4*>(W_hi + off) = hi;
                    *reinterpret_cast<float4*>(W_lo + off) = lo;
                }
            } else {
#pragma unroll
                for (int j = 0; j < TCB_KB / 4; j++) {
                    const int i = tid + j * 128, kk = i >> 5, nq = i & 31;
                    float4 hi, lo;
                    split_tf32x4(rw[j], hi, lo);
                    const uint32_t base = (uint32_t)(kk >> 2) * TCB_W_LBO + (uint32_t)(kk & 3) * 4;
                    const float h4[4] = {hi.x, hi.y, hi.z, hi.w}, l4[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const uint32_t off = base + (uint32_t)(nq * 4 + q) * 16;
                        *reinterpret_cast<float*>(W_hi + off) = h4[q];
                        *reinterpret_cast<float*>(W_lo + off) = l4[q];
                    }
                }
            }
        };
        int it = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            for (int kb = 0; kb < nkb; kb++, it++) {
                const int s = it % S;
                load_block(tile, kb);
                mbar_wait(bar_empty + 8u * s, (uint32_t)(((it / S) & 1) ^ 1));
                store_block(s);
                fence_async_smem();
                mbar_arrive(bar_full + 8u * s);
            }
        }
    } else if (warp < 8) {
        // ===================== epilogue =====================
        const int ew = warp & 3;
        float* stg = epi_s + ew * (32 * 36);
        const bool vec = ((g.ldc & 3) == 0) && (((uintptr_t)g.C & 15) == 0) &&
                         (g.mask == nullptr || ((g.ldmask & 3) == 0 && ((uintptr_t)g.mask & 15) == 0));
        int it = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, it++) {
            const int a = it & 1;
            const int m0 = (tile / ntiles) * 128, n0 = (tile % ntiles) * TCB_NT;
            mbar_wait(bar_tfull + 8u * a, (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            const int mw = m0 + ew * 32;
            const uint32_t trow = tmem + ((uint32_t)(ew * 32) << 16) + (uint32_t)a * (uint32_t)TCB_NT;
            const int ncols = g.N - n0 < TCB_NT ? g.N - n0 : TCB_NT;
            for (int c0 = 0; c0 < ncols; c0 += 32) {
                float v[16];
                tmem_ld16(trow + c0, v);
#pragma unroll
                for (int q = 0; q < 4; q++)
                    *reinterpret_cast<float4*>(stg + lane * 36 + q * 4) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                tmem_ld16(trow + c0 + 16, v);
#pragma unroll
                for (int q = 0; q < 4; q++)
                    *reinterpret_cast<float4*>(stg + lane * 36 + 16 + q * 4) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                __syncwarp();
                const int cq = lane & 7, nl = c0 + cq * 4, n = n0 + nl;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int r = i * 4 + (lane >> 3);
                    const int m = mw + r;
                    if (m < g.M && n < g.N) {
                        const float4 o4 = *reinterpret_cast<const float4*>(stg + r * 36 + cq * 4);
                        float o[4] = {o4.x, o4.y, o4.z, o4.w};
                        float* cp = g.C + (size_t)m * g.ldc + n;
#pragma unroll
                        for (int j = 0; j < 4; j++) if (g.bias && n + j < g.N) o[j] += __ldg(g.bias + n + j);
                        if (vec && n + 3 < g.N) {
                            if (g.accum) {
                                const float4 c = *reinterpret_cast<const float4*>(cp);
                                o[0] += c.x; o[1] += c.y; o[2] += c.z; o[3] += c.w;
                            }
                            if (g.relu) {
#pragma unroll
                                for (int j = 0; j < 4; j++) o[j] = fmaxf(o[j], 0.f);
                            }
                            if (g.mask) {
                                const float4 mk = *reinterpret_cast<const float4*>(g.mask + (size_t)m * g.ldmask + n);
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
        const uint32_t idesc = umma_idesc_tf32(TCB_NT, 0, 0);
        int it = 0, wk = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, it++) {
            const int a = it & 1;
            mbar_wait(bar_tempty + 8u * a, (uint32_t)(((it >> 1) & 1) ^ 1));
            const uint32_t acc = tmem + (uint32_t)a * (uint32_t)TCB_NT;
            for (int kb = 0; kb < nkb; kb++, wk++) {
                const int s = wk % S;
                mbar_wait(bar_full + 8u * s, (uint32_t)((wk / S) & 1));
                tc_fence_after();
                const uint32_t a_hi_s = smem_u32(ring + (size_t)s * TCB_STAGE_BYTES), a_lo_s = a_hi_s + TCB_A_BYTES;
                const uint32_t w_hi_s = a_lo_s + TCB_A_BYTES, w_lo_s = w_hi_s + TCB_W_BYTES;
                const int kcols = g.K - kb * TCB_KB < TCB_KB ? g.K - kb * TCB_KB : TCB_KB;
                const int ksteps = (kcols + 7) >> 3;
                if (elect_one_sync()) {
                    for (int ks = 0; ks < ksteps; ks++) {
                        const uint32_t ao = (uint32_t)ks * 2 * TC_A_LBO, wo = (uint32_t)ks * 2 * TCB_W_LBO;
                        const uint64_t dah = umma_desc(a_hi_s + ao, TC_A_LBO, 128), dal = umma_desc(a_lo_s + ao, TC_A_LBO, 128);
                        const uint64_t dbh = umma_desc(w_hi_s + wo, TCB_W_LBO, 128), dbl = umma_desc(w_lo_s + wo, TCB_W_LBO, 128);
                        umma_tf32(acc, dah, dbh, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
                        umma_tf32(acc, dal, dbh, idesc, 1u);
                        umma_tf32(acc, dah, dbl, idesc, 1u);
                    }
                    umma_commit(bar_empty + 8u * s);    // the stage may be refilled once these MMAs retire
                    if (kb == nkb - 1) umma_commit(bar_tfull + 8u * a);        // accumulator ready for the epilogue warps
                }
                __syncwarp();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 2 * TCB_NT);
}

static bool tc_big_eligible(const GemmArgs& g) {
    // same row threshold as the weight-resident kernel (tc_rows_eligible): batches chunked below it stay on ONE arithmetic path
    if (!tc_enabled() || g.M < 2048 || g.N < 8 || g.K < 8 || (g.K & 3)) return false;
    if (g.A.mode != A_PLAIN || (g.A.ld & 3) || !aligned16(g.A.p) || g.nkb == 2 || g.wconv) return false;
    return true;
}

static int launch_gemm_big_tc(const GemmArgs& g, cudaStream_t st, int sm_count) {
    static bool attr = false;
    const size_t smem = tc_big_smem();
    if (!attr) {
        DOF_CUDA(cudaFuncSetAttribute(gemm_big_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    const int tiles = cdiv(g.M, 128) * cdiv(g.N, TCB_NT);
    const int ctas = tiles < sm_count ? tiles : sm_count;
    const int vecW = (g.wT == 0) ? (aligned16(g.W) && (g.ldw & 3) == 0) : (aligned16(g.W) && (g.ldw & 3) == 0);
    ProfScope ps("gemm_big_tc", st, 2.0 * g.M * g.N * g.K,
                 4.0 * g.M * ((double)g.K + (double)g.N * (1 + (g.accum ? 1 : 0) + (g.mask ? 1 : 0))));
    gemm_big_tc_kernel<<<ctas, TCB_THREADS, smem, st>>>(g, vecW);
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}
