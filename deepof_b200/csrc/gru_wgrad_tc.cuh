// Parameter gradients of one bidirectional GRU layer in ONE pass over the gate gradients (tcgen05 + TMA).
//
// Replaces, for the recurrent layers of the path, what autograd does for torch.nn.GRU's four parameter tensors
// (reference: deepof/clustering/models_new.py, every nn.GRU of the encoder / decoder):
//     dW_ih[c, i] = sum_m dGi[m, c] x[m, i]        db_ih[c] = sum_m dGi[m, c]
//     dW_hh[c, j] = sum_m dGh[m, c] h_prev[m, j]   db_hh[c] = sum_m dGh[m, c]
// The BPTT kernel leaves dG[m] = [dr, dz, dn*r, dn] (4H columns): dGh = columns [0, 3H), dGi = columns [0, 2H) + [3H, 4H).
//
// One accumulator D[4H lanes][I | H | 1 columns] = dG^T . [x | h_prev | 1] per CTA: the dG rows are read ONCE for all four
// gradients (the two-launch version read them twice).  Both operands have the reduction index m as the MMA K dimension
// and are stored row-major in HBM, i.e. they are MN-major operands: for tf32 the only MN-major shared-memory layout is
// SWIZZLE_128B_BASE32B, which is exactly what TMA's CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B writes (tools/probe_mn_major.cu
// verified layout, LBO = stride between 32-column blocks, SBO = 512 B, and that the tensor core TRUNCATES fp32 bit
// patterns to tf32).  So the raw TMA tile IS the high part of the 3xTF32 split; eight warps only compute the low part
// (v - trunc(v), an elementwise pass in the same layout) and zero the h_prev rows that sit on a window boundary.
//
// warps 0-7: low-part pass (+ epilogue on 0-3), warp 8: TMA producer, warp 9: MMA issuer.
#pragma once
#include <cuda.h>
#include "tc_gemm.cuh"
#include "tmap.cuh"

#define GW_BM 32                    // rows (MMA K) per stage
#define GW_BLK (GW_BM * 128)        // bytes of one [GW_BM rows][32 columns] block
#define GW_STAGES 3
#define GW_LO_WARPS 8
#define GW_THREADS (32 * (GW_LO_WARPS + 2))

struct GruWgradMaps { CUtensorMap P[2], X[2], Hs[2]; };      // per direction: dG [M, 4H], x [M, I], h [M, H] (column slice of Hout)
struct GruWgradArgs {
    float* dWih[2]; float* dWhh[2]; float* dbih[2]; float* dbhh[2];
    int M, T, I, H;
    int nbp, nbx;                   // 32-column blocks of dG (4H / 32; dual: both directions) and of x (ceil(I / 32))
    int dual;                       // H = 16: BOTH directions in one accumulator (lanes 0-63 forward gates, 64-127 reversed);
                                    //   x is read once, h_prev comes as two full-width [32-column] blocks of Hout: rows m-1 (its
                                    //   forward half is used) and rows m+1 (its reversed half); maps.Hs[0] spans all 2H columns
};

__device__ __forceinline__ uint64_t umma_desc_mn32(uint32_t saddr, uint32_t lbo_bytes) {
    // MN-major, SWIZZLE_128B_BASE32B (layout type 1): LBO = stride between 32-element MN blocks, SBO = 512 B (4 K rows)
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | ((uint64_t)(512u >> 4) << 32) |
           (1ull << 46) | (1ull << 61);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(GW_THREADS, 1) gru_wgrad_tc_kernel(const __grid_constant__ GruWgradMaps maps, const GruWgradArgs a) {
    extern __shared__ unsigned char gw_raw[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gw_raw) + 1023) & ~(uintptr_t)1023);
    const int d = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nhb = a.dual ? 2 : 1;                               // h_prev blocks
    const int nbq = a.nbx + nhb;                                  // x blocks + the h_prev block(s)
    const int nb = a.nbp + nbq;
    const uint32_t raw_bytes = (uint32_t)(nb + 1) * GW_BLK;       // + the constant-one block (bias gradients)
    const uint32_t lo_bytes = (uint32_t)nb * GW_BLK;
    const uint32_t stage_bytes = raw_bytes + lo_bytes;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(sm + (size_t)GW_STAGES * stage_bytes);
    // mbar[0..S) tma-full, [S..2S) lo-ready (256 arrivals), [2S..3S) empty (MMA commit), [3S] done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 3 * GW_STAGES + 1);

    int rows_per = (a.M + gridDim.x - 1) / gridDim.x;
    rows_per = (rows_per + GW_BM - 1) / GW_BM * GW_BM;
    const int mbeg = blockIdx.x * rows_per;
    const int mend = min(a.M, mbeg + rows_per);
    if (mbeg >= mend) return;
    const int nstage = (mend - mbeg + GW_BM - 1) / GW_BM;
    const int shift = d ? +1 : -1;                                // h_prev of the reversed direction is the NEXT step
    const int tcols = 32 * nbq + 16;                              // accumulator columns: x | h_prev | ones (+ pad)

    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), (uint32_t)tmem_cols_for(tcols));
    if (tid == 32) {
        for (int i = 0; i < GW_STAGES; i++) {
            mbar_init(smem_u32(mbar + i), 1);
            mbar_init(smem_u32(mbar + GW_STAGES + i), GW_LO_WARPS * 32);
            mbar_init(smem_u32(mbar + 2 * GW_STAGES + i), 1);
        }
        mbar_init(smem_u32(mbar + 3 * GW_STAGES), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the constant-one block of every stage: logical column 0 of row r sits in 32-byte unit (0 ^ (r & 3))
    for (int i = tid; i < GW_STAGES * GW_BM * 32; i += GW_THREADS) {
        const int s = i / (GW_BM * 32), w = i % (GW_BM * 32), r = w >> 5, q = w & 31;
        reinterpret_cast<float*>(sm + (size_t)s * stage_bytes + (size_t)nb * GW_BLK)[w] = (q == ((r & 3) << 3)) ? 1.0f : 0.0f;
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t bar_full = smem_u32(mbar), bar_ready = smem_u32(mbar + GW_STAGES), bar_empty = smem_u32(mbar + 2 * GW_STAGES),
                   bar_done = smem_u32(mbar + 3 * GW_STAGES);

    if (warp < GW_LO_WARPS) {
        // ===================== low-part pass =====================
        const int n4 = nb * (GW_BLK / 16);
        const int hblk0 = (a.nbp + a.nbx) * (GW_BLK / 16);
        for (int it = 0; it < nstage; it++) {
            const int s = it % GW_STAGES;
            const int m0 = mbeg + it * GW_BM;
            float4* raw4 = reinterpret_cast<float4*>(sm + (size_t)s * stage_bytes);
            float4* lo4 = reinterpret_cast<float4*>(sm + (size_t)s * stage_bytes + raw_bytes);
            mbar_wait(bar_full + 8u * s, (uint32_t)((it / GW_STAGES) & 1));
#pragma unroll 4
            for (int i = tid; i < n4; i += GW_LO_WARPS * 32) {
                float4 v = raw4[i];
                if (i >= hblk0) {
                    const int hb = (i - hblk0) / (GW_BLK / 16);           // dual: block 0 = rows m-1 (forward), 1 = rows m+1 (reversed)
                    const int t = (m0 + (((i - hblk0) % (GW_BLK / 16)) >> 3)) % a.T;        // 8 float4 per 128-byte row
                    const bool rev = a.dual ? hb == 1 : d == 1;
                    if (rev ? (t == a.T - 1) : (t == 0)) { v = make_float4(0.f, 0.f, 0.f, 0.f); raw4[i] = v; }
                }
                float4 lo;
                lo.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
                lo.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
                lo.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
                lo.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
                lo4[i] = lo;
            }
            fence_async_smem();
            mbar_arrive(bar_ready + 8u * s);
        }
        // ---- epilogue (warps 0-3): thread = gate column n of dG; vector reductions into the four gradients
        if (warp < 4) {
            mbar_wait(bar_done, 0u);
            tc_fence_after();
            const int H = a.H, I = a.I;
            const int n = warp * 32 + lane;
            const int dd = a.dual ? n / (4 * H) : d;           // direction this accumulator lane belongs to
            const int gcol = a.dual ? n % (4 * H) : n;         // gate column inside the direction's dG
            const bool valid = a.dual ? true : n < 4 * H;
            const int row_hh = gcol < 3 * H ? gcol : -1;
            const int row_ih = gcol < 2 * H ? gcol : (gcol >= 3 * H ? gcol - H : -1);
            const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
            const int hcol0 = 32 * a.nbx, onecol = 32 * nbq;
            // the gradient tensors live at arbitrary float offsets of the flat state: vector reductions only when aligned
            const bool vih = ((reinterpret_cast<uintptr_t>(a.dWih[dd]) & 15) == 0), vhh = ((reinterpret_cast<uintptr_t>(a.dWhh[dd]) & 15) == 0);
            for (int c0 = 0; c0 < tcols; c0 += 16) {
                float v[16];
                tmem_ld16(trow + c0, v);
                if (!valid) continue;
                if (c0 < hcol0) {
                    if (row_ih >= 0) {
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const int j = c0 + 4 * q;
                            if (j < I) {
                                float* o = a.dWih[dd] + (size_t)row_ih * I + j;
                                if (vih)
                                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(v[4 * q]), "f"(v[4 * q + 1]),
                                                 "f"(v[4 * q + 2]), "f"(v[4 * q + 3]) : "memory");
                                else { atomicAdd(o, v[4 * q]); atomicAdd(o + 1, v[4 * q + 1]); atomicAdd(o + 2, v[4 * q + 2]); atomicAdd(o + 3, v[4 * q + 3]); }
                            }
                        }
                    }
                } else if (c0 < onecol) {
                    // h_prev columns.  single: block columns [0, H).  dual: block 0 (rows m-1) columns [0, H) belong to the
                    // forward lanes, block 1 (rows m+1) columns [H, 2H) to the reversed lanes
                    const int hb = (c0 - hcol0) / 32, jb = (c0 - hcol0) % 32;
                    const int joff = a.dual ? (hb == 1 ? H : 0) : 0;
                    if (row_hh >= 0 && (!a.dual || hb == dd)) {
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const int j = jb + 4 * q - joff;
                            if (j >= 0 && j < H) {
                                float* o = a.dWhh[dd] + (size_t)row_hh * H + j;
                                if (vhh)
                                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(v[4 * q]), "f"(v[4 * q + 1]),
                                                 "f"(v[4 * q + 2]), "f"(v[4 * q + 3]) : "memory");
                                else { atomicAdd(o, v[4 * q]); atomicAdd(o + 1, v[4 * q + 1]); atomicAdd(o + 2, v[4 * q + 2]); atomicAdd(o + 3, v[4 * q + 3]); }
                            }
                        }
                    }
                } else {
                    if (row_hh >= 0) atomicAdd(a.dbhh[dd] + row_hh, v[0]);
                    if (row_ih >= 0) atomicAdd(a.dbih[dd] + row_ih, v[0]);
                }
            }
        }
    } else if (warp == GW_LO_WARPS) {
        // ===================== TMA producer (one thread) =====================
        if (lane == 0) {
            for (int it = 0; it < nstage; it++) {
                const int s = it % GW_STAGES;
                const int m0 = mbeg + it * GW_BM;
                mbar_wait(bar_empty + 8u * s, (uint32_t)(((it / GW_STAGES) & 1) ^ 1));
                const uint32_t raw = smem_u32(sm + (size_t)s * stage_bytes), bar = bar_full + 8u * s;
                mbar_expect_tx(bar, lo_bytes);
                const int pper = a.dual ? a.nbp / 2 : a.nbp;               // dG blocks per direction
                for (int b = 0; b < a.nbp; b++)
                    tma_load_2d(raw + (uint32_t)b * GW_BLK, &maps.P[a.dual ? b / pper : d], (b % pper) * 32, m0, bar);
                for (int b = 0; b < a.nbx; b++) tma_load_2d(raw + (uint32_t)(a.nbp + b) * GW_BLK, &maps.X[d], b * 32, m0, bar);
                if (a.dual) {
                    tma_load_2d(raw + (uint32_t)(a.nbp + a.nbx) * GW_BLK, &maps.Hs[0], 0, m0 - 1, bar);
                    tma_load_2d(raw + (uint32_t)(a.nbp + a.nbx + 1) * GW_BLK, &maps.Hs[0], 0, m0 + 1, bar);
                } else {
                    tma_load_2d(raw + (uint32_t)(a.nbp + a.nbx) * GW_BLK, &maps.Hs[d], 0, m0 + shift, bar);
                }
            }
        }
    } else if (lane == 0) {
        // ===================== MMA issuer (one thread) =====================
        const uint32_t idesc_full = umma_idesc_tf32(tcols, 1, 1), idesc_lo = umma_idesc_tf32(32 * nbq, 1, 1);
        for (int it = 0; it < nstage; it++) {
            const int s = it % GW_STAGES;
            mbar_wait(bar_ready + 8u * s, (uint32_t)((it / GW_STAGES) & 1));
            tc_fence_after();
            const uint32_t p_hi = smem_u32(sm + (size_t)s * stage_bytes), p_lo = p_hi + raw_bytes;
            const uint32_t q_hi = p_hi + (uint32_t)a.nbp * GW_BLK, q_lo = p_lo + (uint32_t)a.nbp * GW_BLK;
#pragma unroll
            for (int ks = 0; ks < GW_BM / 8; ks++) {
                const uint32_t o = (uint32_t)ks * 1024u;
                const uint64_t dah = umma_desc_mn32(p_hi + o, GW_BLK), dal = umma_desc_mn32(p_lo + o, GW_BLK);
                const uint64_t dbh = umma_desc_mn32(q_hi + o, GW_BLK), dbl = umma_desc_mn32(q_lo + o, GW_BLK);
                umma_tf32(tmem, dah, dbh, idesc_full, (it == 0 && ks == 0) ? 0u : 1u);
                umma_tf32(tmem, dal, dbh, idesc_full, 1u);
                umma_tf32(tmem, dah, dbl, idesc_lo, 1u);
            }
            umma_commit(bar_empty + 8u * s);
        }
        umma_commit(bar_done);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, (uint32_t)tmem_cols_for(tcols));
}

// ---------------------------------------------------------------------------
// host side: tensor maps (cached per buffer), eligibility, launch
// ---------------------------------------------------------------------------
struct TmapKey { const void* base; int rows, cols, ld; };
struct TmapEntry { TmapKey k; CUtensorMap m; };
static std::vector<TmapEntry> g_tmaps;

// [rows][cols] fp32 view with row pitch ld floats; box = [GW_BM rows][32 columns], 128B_ATOM_32B swizzle, zero OOB fill
static int tmap_rows32(const float* base, int rows, int cols, int ld, CUtensorMap* out) {
    for (const TmapEntry& e : g_tmaps)
        if (e.k.base == base && e.k.rows == rows && e.k.cols == cols && e.k.ld == ld) { *out = e.m; return DOF_OK; }
    dof_tmap_encode_fn enc = tmap_encoder();
    if (!enc) DOF_FAIL(DOF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    TmapEntry e;
    e.k = TmapKey{base, rows, cols, ld};
    memset(&e.m, 0, sizeof(e.m));
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32, GW_BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&e.m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) DOF_FAIL(DOF_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for a [%d, %d] view, pitch %d", (int)r, rows, cols, ld);
    if (g_tmaps.size() >= 256) g_tmaps.clear();
    g_tmaps.push_back(e);
    *out = e.m;
    return DOF_OK;
}

static bool g_gru_wgrad_dual = getenv("DOF_GRU_WGRAD_NODUAL") == nullptr;     // env switch for A/B measurements
static inline bool gru_wgrad_dual(int H) { return g_gru_wgrad_dual && H == 16; }

static size_t gru_wgrad_smem(int H, int I) {
    const bool dual = gru_wgrad_dual(H);
    const int nbp = (dual ? 2 : 1) * 4 * H / 32, nbq = cdiv(I, 32) + (dual ? 2 : 1), nb = nbp + nbq;
    return (size_t)GW_STAGES * ((size_t)(2 * nb + 1) * GW_BLK) + (3 * GW_STAGES + 1) * 8 + 16 + 1024 + 128;
}

static bool gru_wgrad_tc_eligible(int M, int H, int I, int ldx, const float* dG0, const float* dG1, const float* X, const float* Hout) {
    if (!tc_enabled() || !tmap_encoder()) return false;
    if (H != 16 && H != 32) return false;
    if (I < 4 || I > 64 || (I & 3) || (ldx & 3) || M < 4096) return false;
    if (!aligned16(dG0) || !aligned16(dG1) || !aligned16(X) || !aligned16(Hout)) return false;
    return gru_wgrad_smem(H, I) <= 227 * 1024;
}

// dG[d]: [M, 4H]; X: [M, I] with pitch ldx; Hout: [M, 2H] (forward | reversed halves); gradients are accumulated (+=)
static int launch_gru_wgrad_tc(float* const dG[2], const float* X, int ldx, const float* Hout, float* const dWih[2],
                               float* const dWhh[2], float* const dbih[2], float* const dbhh[2], int M, int T, int I, int H,
                               int sm_count, cudaStream_t st) {
    GruWgradMaps maps;
    GruWgradArgs a;
    memset(&a, 0, sizeof(a));
    const bool dual = gru_wgrad_dual(H);
    for (int d = 0; d < 2; d++) {
        DOF_TRY(tmap_rows32(dG[d], M, 4 * H, 4 * H, &maps.P[d]));
        DOF_TRY(tmap_rows32(X, M, I, ldx, &maps.X[d]));
        if (dual) DOF_TRY(tmap_rows32(Hout, M, 2 * H, 2 * H, &maps.Hs[d]));          // all 2H columns, both halves
        else DOF_TRY(tmap_rows32(Hout + d * H, M, H, 2 * H, &maps.Hs[d]));
        a.dWih[d] = dWih[d]; a.dWhh[d] = dWhh[d]; a.dbih[d] = dbih[d]; a.dbhh[d] = dbhh[d];
    }
    a.M = M; a.T = T; a.I = I; a.H = H; a.dual = dual ? 1 : 0;
    a.nbp = (dual ? 2 : 1) * 4 * H / 32; a.nbx = cdiv(I, 32);
    const size_t smem = gru_wgrad_smem(H, I);
    static bool attr = false;
    if (!attr) {
        DOF_CUDA(cudaFuncSetAttribute(gru_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr = true;
    }
    int ctas = dual ? sm_count : sm_count / 2;
    const int maxsplit = cdiv(M, 8 * GW_BM);
    if (ctas > maxsplit) ctas = maxsplit;
    if (ctas < 1) ctas = 1;
    ProfScope ps("gru_wgrad_tc", st, 2.0 * 2.0 * M * 3.0 * H * (I + H), 2.0 * 4.0 * M * (4.0 * H + I + H));
    gru_wgrad_tc_kernel<<<dim3(ctas, 1, dual ? 1 : 2), GW_THREADS, smem, st>>>(maps, a);
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}
