// Fused GRU layer backward (BPTT) on tcgen05: gate gradients + recurrent matmul + input gradient in ONE kernel.
//
// Per step (reverse of the forward order) and per tile of 128 sequences (= 128 TMEM lanes), one direction per CTA:
//   gate warps (2 threads / sequence):  d = dh + dOut_t;  dn = d(1-z); dz = d(h_prev - n); da_n = dn(1-n^2);
//       da_z = dz z(1-z); da_r = da_n hn r(1-r);  dG_t = [da_r | da_z | da_n r | da_n]   (the reference's autograd of
//       torch.nn.GRU, models_new.py:243-268) -> written ONCE into shared memory as the hi / lo operand planes
//   tensor core:  acc_dh = dG_t[:, 0:3H) . W_hh          (-> dh_{t-1} = d z + acc_dh, read back from TMEM)
//                 acc_dx = [da_r | da_z | da_n] . W_ih   (-> dX_t, added to HBM with vector reductions: the two
//                                                          directions of a layer accumulate into the same rows)
//   store warps:  dG_t rows -> HBM (the weight-gradient GEMMs read them), dX_t from TMEM -> HBM.
// The `lo` plane holds the EXACT remainder v - hi (the tensor core only looks at its upper 19 bits), so hi + lo
// reproduces the fp32 value and the operand tile doubles as the staging buffer of the dG rows.
// This replaces the SIMT BPTT kernel (gru.cuh) and the separate input-gradient GEMM of gru_param_grads.
//
// The saved gates come in the TILED layout the fused forward kernel writes when asked to
//   GtT[((tile*T + t) * H + c) * 128 + row][4 floats],  c = 16-byte chunk of the row r|z|n|hn (H chunks)
// so that a gate thread (= one row) reads them with fully coalesced 16-byte loads and no staging.
#pragma once
#include "common.cuh"
#include "tc_gemm.cuh"
#include "gru_tc.cuh"

struct GruBwdTcArgs {
    const float* Whh[2]; const float* Wih[2];
    const int* len;           // [S] or null
    const float* Hout;        // [S,T,2H] forward outputs (h_prev source)
    const float* GtT[2];      // tiled saved gates per direction
    const float* dOut;        // [S,T,2H] or null
    const float* dHn;         // [S,2H] or null
    float* dG[2];             // [S,T,4H] row-major
    float* dX;                // [S,T,I] (zeroed by the caller; both directions add) or null
    const float* dXmask;      // [S,T,I] or null: dX is kept only where mask > 0 (ReLU backward)
    int S, T, H, I;
};

struct GruBwdTcGeom { int whh_lbo, wih_lbo, tmem_cols; uint32_t whh_bytes, wih_bytes, a_bytes; };

#define GBT_THREADS 512          // warps 0-3 store warps | 4-11 gate warps | 12 MMA issuer | 13-15 dG-row drain helpers
                                 // (13 warps are charged registers like 16, so the 3 helpers are free; the gate warps were
                                 // stalled on the store warps finishing the dG rows of the previous step)
#define GBT_NDRAIN 7             // warps that stream the dG rows of a step: 0-3 and 13-15
#define GBT_MMA_WARP 12

__device__ __forceinline__ void split_tf32_exact(float v, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    lo = v - hi;                  // exact in fp32; the tensor core truncates it to tf32 itself
}

template <int H>
__global__ void __launch_bounds__(GBT_THREADS) gru_bwd_tc_kernel(const GruBwdTcArgs a, const GruBwdTcGeom geo) {
    constexpr int HC = H / 2;                                         // hidden units per gate thread
    extern __shared__ __align__(128) unsigned char bsm[];
    const int dir = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int I = a.I, T = a.T;
    unsigned char* Whh_hi = bsm;                                      // B operand [N = H][K = 3H]:  W_hh[c][j] at (n = j, k = c)
    unsigned char* Whh_lo = Whh_hi + geo.whh_bytes;
    unsigned char* Wih_hi = Whh_lo + geo.whh_bytes;                   // B operand [N = I][K = 3H]:  W_ih[c][i] at (n = i, k = c)
    unsigned char* Wih_lo = Wih_hi + geo.wih_bytes;
    unsigned char* A_hi = Wih_lo + geo.wih_bytes;                     // dG tile [128][4H], H chunks of 4 floats
    unsigned char* A_lo = A_hi + geo.a_bytes;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(A_lo + geo.a_bytes);
    // mbar: [0] a_full (256 gate threads) [1] dh_full (commit) [2] dx_full (commit) [3] st_done (128 store threads: dG rows
    // copied out of the operand tile) [4] dx_read (128 store threads: acc_dx read out of TMEM)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 6);
    int* lens_s = reinterpret_cast<int*>(tmem_slot + 4);              // [128]
    const int s0 = blockIdx.x * 128;
    const bool want_dx = a.dX != nullptr;

    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), geo.tmem_cols);
    if (tid == 32) {
        mbar_init(smem_u32(mbar + 0), 256); mbar_init(smem_u32(mbar + 1), 1);
        mbar_init(smem_u32(mbar + 2), 1); mbar_init(smem_u32(mbar + 3), GBT_NDRAIN * 32); mbar_init(smem_u32(mbar + 4), 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const float* wh = a.Whh[dir];
        for (int i = tid; i < 3 * H * H; i += GBT_THREADS) {
            const int c = i / H, j = i - c * H;                       // W_hh[c][j]
            float hi, lo;
            split_tf32(__ldg(wh + i), hi, lo);
            const uint32_t off = (uint32_t)j * 16 + (uint32_t)(c >> 2) * geo.whh_lbo + (c & 3) * 4;
            *reinterpret_cast<float*>(Whh_hi + off) = hi;
            *reinterpret_cast<float*>(Whh_lo + off) = lo;
        }
        if (want_dx) {
            const float* wi = a.Wih[dir];
            for (int i = tid; i < 3 * H * I; i += GBT_THREADS) {
                const int c = i / I, ii = i - c * I;                  // W_ih[c][ii]
                float hi, lo;
                split_tf32(__ldg(wi + i), hi, lo);
                const uint32_t off = (uint32_t)ii * 16 + (uint32_t)(c >> 2) * geo.wih_lbo + (c & 3) * 4;
                *reinterpret_cast<float*>(Wih_hi + off) = hi;
                *reinterpret_cast<float*>(Wih_lo + off) = lo;
            }
        }
        for (int i = tid; i < 128; i += GBT_THREADS) lens_s[i] = (s0 + i < a.S) ? (a.len ? __ldg(a.len + s0 + i) : T) : -1;
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t bar_afull = smem_u32(mbar), bar_dh = smem_u32(mbar + 1), bar_dx = smem_u32(mbar + 2), bar_st = smem_u32(mbar + 3),
                   bar_dxr = smem_u32(mbar + 4);
    const uint32_t acc_dh = tmem, acc_dx = tmem + H;

    if (warp >= 4 && warp < GBT_MMA_WARP) {
        // ===================== gate warps: two threads per sequence =====================
        const int ew = warp & 3, row = ew * 32 + lane, half = (warp - 4) >> 2;
        const int j0 = half * HC;
        const int s = s0 + row;
        const int len = lens_s[row] < 0 ? 0 : lens_s[row];
        const float* GtT = a.GtT[dir];
        float part[HC];                                               // dh before the recurrent term of the next step
#pragma unroll
        for (int j = 0; j < HC; j++) part[j] = 0.f;
        if (a.dHn && s < a.S) {
#pragma unroll
            for (int q = 0; q < HC / 4; q++) {
                const float4 d = __ldg(reinterpret_cast<const float4*>(a.dHn + (size_t)s * 2 * H + dir * H + j0 + q * 4));
                part[q * 4] = d.x; part[q * 4 + 1] = d.y; part[q * 4 + 2] = d.z; part[q * 4 + 3] = d.w;
            }
        }
        const size_t tile_base = (size_t)blockIdx.x * T;
        for (int step = 0; step < T; step++) {
            const int t = dir ? step : (T - 1 - step);
            const int tp = dir ? t + 1 : t - 1;
            const bool valid = t < len;
            // all global loads of this step first (they do not depend on dh): their latency hides behind the MMAs of
            // the previous step
            const float* gt = GtT + ((tile_base + t) * H) * 512 + row * 4;      // chunk c at gt + c*512
            float4 r4[HC / 4], z4[HC / 4], n4[HC / 4], q4[HC / 4], hp4[HC / 4], do4[HC / 4];
#pragma unroll
            for (int q = 0; q < HC / 4; q++) {
                const int cq = (j0 >> 2) + q;
                hp4[q] = make_float4(0.f, 0.f, 0.f, 0.f); do4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) {
                    r4[q] = __ldg(reinterpret_cast<const float4*>(gt + (size_t)cq * 512));
                    z4[q] = __ldg(reinterpret_cast<const float4*>(gt + (size_t)(H / 4 + cq) * 512));
                    n4[q] = __ldg(reinterpret_cast<const float4*>(gt + (size_t)(2 * (H / 4) + cq) * 512));
                    q4[q] = __ldg(reinterpret_cast<const float4*>(gt + (size_t)(3 * (H / 4) + cq) * 512));
                    if (tp >= 0 && tp < len) hp4[q] = __ldg(reinterpret_cast<const float4*>(a.Hout + ((size_t)s * T + tp) * 2 * H + dir * H + j0 + q * 4));
                    if (a.dOut) do4[q] = __ldg(reinterpret_cast<const float4*>(a.dOut + ((size_t)s * T + t) * 2 * H + dir * H + j0 + q * 4));
                } else {
                    r4[q] = z4[q] = n4[q] = q4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            // pull the next step's saved gates towards L2 while this step computes
            if (step + 1 < T) {
                const int tn = dir ? t + 1 : t - 1;
                const float* gn = GtT + ((tile_base + tn) * H + (j0 >> 2)) * 512 + row * 4;
                if ((lane & 7) == 0) {                                      // one lane per 128-byte line
#pragma unroll
                    for (int q = 0; q < 4; q++)
#pragma unroll
                        for (int qq = 0; qq < HC / 4; qq++)
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(gn + (size_t)(q * (H / 4) + qq) * 512));
                }
            }
            // recurrent term of the previous step: dh = part + dG_{prev} . W_hh
            float dh[HC];
            if (step > 0) {
                mbar_wait(bar_dh, (uint32_t)((step - 1) & 1));
                tc_fence_after();
                float v[HC];
                tmem_ld_hc<HC>(acc_dh + ((uint32_t)(ew * 32) << 16) + (uint32_t)j0, v);
#pragma unroll
                for (int j = 0; j < HC; j++) dh[j] = part[j] + v[j];
            } else {
#pragma unroll
                for (int j = 0; j < HC; j++) dh[j] = part[j];
            }
            // the operand tile is free again once the MMAs of the previous step retired and its rows were stored
            if (step > 0) {
                if (want_dx) mbar_wait(bar_dx, (uint32_t)((step - 1) & 1));
                mbar_wait(bar_st, (uint32_t)((step - 1) & 1));
            }
#pragma unroll
            for (int q = 0; q < HC / 4; q++) {
                const int cq = (j0 >> 2) + q;                                 // chunk inside a gate block
                float o_r[4] = {0.f, 0.f, 0.f, 0.f}, o_z[4] = {0.f, 0.f, 0.f, 0.f}, o_h[4] = {0.f, 0.f, 0.f, 0.f}, o_n[4] = {0.f, 0.f, 0.f, 0.f};
                if (valid) {
                    const float r[4] = {r4[q].x, r4[q].y, r4[q].z, r4[q].w}, z[4] = {z4[q].x, z4[q].y, z4[q].z, z4[q].w};
                    const float n[4] = {n4[q].x, n4[q].y, n4[q].z, n4[q].w}, hn[4] = {q4[q].x, q4[q].y, q4[q].z, q4[q].w};
                    const float hp[4] = {hp4[q].x, hp4[q].y, hp4[q].z, hp4[q].w}, dov[4] = {do4[q].x, do4[q].y, do4[q].z, do4[q].w};
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const float d = dh[q * 4 + e] + dov[e];
                        const float dn = d * (1.0f - z[e]);
                        const float dz = d * (hp[e] - n[e]);
                        const float dan = dn * (1.0f - n[e] * n[e]);
                        o_z[e] = dz * z[e] * (1.0f - z[e]);
                        o_r[e] = dan * hn[e] * r[e] * (1.0f - r[e]);
                        o_h[e] = dan * r[e];
                        o_n[e] = dan;
                        part[q * 4 + e] = d * z[e];
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 4; e++) part[q * 4 + e] = dh[q * 4 + e];
                }
                // operand planes: gate block g (0 r, 1 z, 2 n*r, 3 n), chunk cq -> tile chunk g*(H/4) + cq
                auto put = [&](int g, const float (&o)[4]) {
                    float4 hi, lo;
                    split_tf32_exact(o[0], hi.x, lo.x); split_tf32_exact(o[1], hi.y, lo.y);
                    split_tf32_exact(o[2], hi.z, lo.z); split_tf32_exact(o[3], hi.w, lo.w);
                    const uint32_t off = (uint32_t)row * 16 + (uint32_t)(g * (H / 4) + cq) * TC_A_LBO;
                    *reinterpret_cast<float4*>(A_hi + off) = hi;
                    *reinterpret_cast<float4*>(A_lo + off) = lo;
                };
                put(0, o_r); put(1, o_z); put(2, o_h); put(3, o_n);
            }
            fence_async_smem();
            tc_fence_before();
            mbar_arrive(bar_afull);
        }
    } else if (warp == GBT_MMA_WARP) {
        if (lane == 0) {
            // ===================== MMA issuer =====================
            const uint32_t id_h = umma_idesc_tf32(H, 0, 0), id_x = umma_idesc_tf32(I, 0, 0);
            const uint32_t a_hi = smem_u32(A_hi), a_lo = smem_u32(A_lo);
            const uint32_t whh_hi = smem_u32(Whh_hi), whh_lo = smem_u32(Whh_lo), wih_hi = smem_u32(Wih_hi), wih_lo = smem_u32(Wih_lo);
            for (int step = 0; step < T; step++) {
                mbar_wait(bar_afull, (uint32_t)(step & 1));
                tc_fence_after();
                // dh part: K = 3H = tile chunks [0, 3H/4)
                for (int ks = 0; ks < (3 * H) >> 3; ks++) {
                    const uint32_t ao = (uint32_t)ks * 2 * TC_A_LBO, wo = (uint32_t)ks * 2 * geo.whh_lbo;
                    const uint64_t dah = umma_desc(a_hi + ao, TC_A_LBO, 128), dal = umma_desc(a_lo + ao, TC_A_LBO, 128);
                    const uint64_t dbh = umma_desc(whh_hi + wo, geo.whh_lbo, 128), dbl = umma_desc(whh_lo + wo, geo.whh_lbo, 128);
                    umma_tf32(acc_dh, dah, dbh, id_h, ks > 0 ? 1u : 0u);
                    umma_tf32(acc_dh, dal, dbh, id_h, 1u);
                    umma_tf32(acc_dh, dah, dbl, id_h, 1u);
                }
                umma_commit(bar_dh);
                if (want_dx) {
                    if (step > 0) { mbar_wait(bar_dxr, (uint32_t)((step - 1) & 1)); tc_fence_after(); }   // acc_dx of the previous step read out
                    // dX part: K = 3H = tile chunks [0, 2H/4) (da_r, da_z) and [3H/4, H) (da_n) against W_ih rows r, z, n
                    for (int ks = 0; ks < (3 * H) >> 3; ks++) {
                        const int ca = ks < (2 * H) >> 3 ? 2 * ks : 2 * ks + (H >> 2);       // first A chunk of this K step
                        const uint32_t ao = (uint32_t)ca * TC_A_LBO, wo = (uint32_t)ks * 2 * geo.wih_lbo;
                        const uint64_t dah = umma_desc(a_hi + ao, TC_A_LBO, 128), dal = umma_desc(a_lo + ao, TC_A_LBO, 128);
                        const uint64_t dbh = umma_desc(wih_hi + wo, geo.wih_lbo, 128), dbl = umma_desc(wih_lo + wo, geo.wih_lbo, 128);
                        umma_tf32(acc_dx, dah, dbh, id_x, ks > 0 ? 1u : 0u);
                        umma_tf32(acc_dx, dal, dbh, id_x, 1u);
                        umma_tf32(acc_dx, dah, dbl, id_x, 1u);
                    }
                    umma_commit(bar_dx);
                }
            }
        }
    } else if (warp < 4 || warp > GBT_MMA_WARP) {
        // ===================== store warps: dG rows (from the operand planes) and dX rows (from TMEM) =====================
        // the dG rows of a step are dealt round-robin to the 7 draining warps; only warps 0-3 (TMEM lane groups) do dX
        float* dG = a.dG[dir];
        constexpr int LG = H, RG = 32 / LG;          // lanes per dG row (4H floats = H chunks), rows per warp instruction
        const int sw = warp < 4 ? warp : 4 + warp - (GBT_MMA_WARP + 1);
        for (int step = 0; step < T; step++) {
            const int t = dir ? step : (T - 1 - step);
            mbar_wait(bar_afull, (uint32_t)(step & 1));
#pragma unroll 4
            for (int q = sw; q < 128 / RG; q += GBT_NDRAIN) {
                const int r = q * RG + lane / LG, c = lane % LG;
                const uint32_t off = (uint32_t)r * 16 + (uint32_t)c * TC_A_LBO;
                const float4 hi = *reinterpret_cast<const float4*>(A_hi + off);
                const float4 lo = *reinterpret_cast<const float4*>(A_lo + off);
                if (lens_s[r] >= 0)
                    *reinterpret_cast<float4*>(dG + ((size_t)(s0 + r) * T + t) * 4 * H + c * 4) =
                        make_float4(hi.x + lo.x, hi.y + lo.y, hi.z + lo.z, hi.w + lo.w);
            }
            mbar_arrive(bar_st);                          // the operand tile may be refilled (its MMAs are tracked by dh / dx_full)
            if (want_dx && warp < 4) {
                mbar_wait(bar_dx, (uint32_t)(step & 1));
                tc_fence_after();
                const int r = warp * 32 + lane;
                const bool ok = lens_s[r] >= 0;
                float* xr = a.dX + ((size_t)(s0 + r) * T + t) * I;
                const float* mr = a.dXmask ? a.dXmask + ((size_t)(s0 + r) * T + t) * I : nullptr;
                for (int c0 = 0; c0 < I; c0 += 16) {
                    float v[16];
                    tmem_ld16(acc_dx + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
                    if (ok) {
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            if (c0 + q * 4 < I) {
                                float4 o = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                                if (mr) {
                                    const float4 m = __ldg(reinterpret_cast<const float4*>(mr + c0 + q * 4));
                                    o.x = m.x > 0.f ? o.x : 0.f; o.y = m.y > 0.f ? o.y : 0.f; o.z = m.z > 0.f ? o.z : 0.f; o.w = m.w > 0.f ? o.w : 0.f;
                                }
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(xr + c0 + q * 4), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
                            }
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(bar_dxr);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, geo.tmem_cols);
}

static bool gru_bwd_tc_geom(int H, int I, bool want_dx, GruBwdTcGeom& g, size_t& smem) {
    g.whh_lbo = H * 16 + 16;
    g.wih_lbo = I * 16 + 16;
    g.whh_bytes = (uint32_t)(3 * H / 4) * g.whh_lbo;
    g.wih_bytes = want_dx ? (uint32_t)(3 * H / 4) * g.wih_lbo : 0;
    g.a_bytes = (uint32_t)H * TC_A_LBO;                     // 4H / 4 = H chunks
    g.tmem_cols = tmem_cols_for(H + (want_dx ? I : 0));
    smem = 2 * (size_t)g.whh_bytes + 2 * (size_t)g.wih_bytes + 2 * (size_t)g.a_bytes + 6 * 8 + 16 + 128 * 4 + 128;
    return smem <= 227 * 1024;
}

static bool gru_bwd_tc_eligible(int H, int I) {
    GruBwdTcGeom g; size_t smem;
    return gru_tc_eligible(0, H, I) && (I % 16 == 0) && gru_bwd_tc_geom(H, I, true, g, smem);
}

template <int H>
static int gru_bwd_tc_launch_t(const GruBwdTcArgs& a, const GruBwdTcGeom& geo, size_t smem, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        DOF_CUDA(cudaFuncSetAttribute(gru_bwd_tc_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr = true;
    }
    dim3 grid(cdiv(a.S, 128), 2);
    gru_bwd_tc_kernel<H><<<grid, GBT_THREADS, smem, st>>>(a, geo);
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

static int launch_gru_bwd_tc(const GruBwdTcArgs& a, cudaStream_t st) {
    GruBwdTcGeom geo;
    size_t smem = 0;
    if (!gru_bwd_tc_geom(a.H, a.I, a.dX != nullptr, geo, smem)) DOF_FAIL(DOF_ERR_UNSUPPORTED, "fused GRU backward tile does not fit (H=%d I=%d)", a.H, a.I);
    const double rows = (double)a.S * a.T * 2;
    ProfScope ps(a.H == 32 ? "gru_bwd_tc_h32" : "gru_bwd_tc_h16", st, rows * 2.0 * 3 * a.H * (a.H + (a.dX ? a.I : 0)),
                 rows * 4.0 * a.H * (4 + 4 + 1 + (a.dOut ? 1 : 0)) + (a.dX ? rows * 4.0 * a.I * (a.dXmask ? 2 : 1) : 0.0));
    if (a.H == 32) return gru_bwd_tc_launch_t<32>(a, geo, smem, st);
    return gru_bwd_tc_launch_t<16>(a, geo, smem, st);
}
