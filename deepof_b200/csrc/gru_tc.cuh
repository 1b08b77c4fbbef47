// Fused GRU layer forward on tcgen05: input projection + recurrent matmul + gate math in ONE kernel.
//
// Replaces, for one direction of torch.nn.GRU (gate order r,z,n; reference RecurrentBlockPT / RecurrentDecoderPT,
// deepof/clustering/models_new.py:217-278, 326-373), the pair (tall-skinny GEMM  Gi = X.W_ih^T + b_ih  written
// to HBM, SIMT recurrent kernel reading it back).  Gi never exists: per time step the tensor core computes
//     acc[:, 0:3H)  = x_t . W_ih^T                      (issued one step AHEAD: it does not depend on h)
//     acc[:, 0:2H) += h_{t-1} . W_h{r,z}^T
//     acc[:, 3H:4H) = h_{t-1} . W_hn^T
// into TMEM (M = 128 sequences = 128 TMEM lanes, fp32 accumulate, 3xTF32 operand split = fp32-class accuracy),
// four gate warps (thread = sequence) read their TMEM row, apply sigmoid / tanh, keep h_t in registers, write the
// outputs, and publish h_t (hi / lo planes, canonical K-major UMMA layout) in shared memory as the A operand of
// the next step.  Weights (hi / lo) stay resident in shared memory for all T steps.
//
// Warp roles (416 threads): warps 0-3 stage x_t tiles (global -> registers -> hi/lo split -> shared) one step
// ahead and stream the staged gate / output rows of the finished step to HBM (one contiguous row per warp
// instruction, full 128-byte lines) so the writes are off the serial path; warps 4-11 gates (TMEM lane group =
// warp & 3; two threads per sequence, each half of the hidden units); warp 12 lane 0 issues the MMAs.  (256 small TMA bulk stores per step were
// tried first and measured 5x slower: the copy engine's per-copy overhead, not bytes, was the limit.)
// Packed-sequence semantics as in gru.cuh: on steps t >= len[s] the state is held and the output is zero.
#pragma once
#include "common.cuh"
#include "tc_gemm.cuh"
#include "tmap.cuh"

struct GruTcArgs {
    const float* X; long long x_ss; int x_st;     // x_t of sequence s: X + s*x_ss + t*x_st, I floats (x_st = 0: repeated input)
    const float* Wih[2]; const float* Whh[2]; const float* bih[2]; const float* bhh[2];
    const int* len;      // [S] or null (all T)
    float* Hout;         // [S,T,2H] (this direction's half) or null
    float* Gt[2];        // [S,T,4H] = r|z|n|hn per direction, or null
    float* Hn;           // [S,2H] final states [fwd|bwd] or null
    int S, T, H, I;
    long long* dbg;      // null, or [6 roles][T][4] clock64 stamps of CTA (0, 0) (tools/prof_gru.py)
    int gt_tiled;        // 1: Gt is written in the tiled layout the fused backward kernel reads (gru_bwd_tc.cuh):
                         //    Gt[((tile*T + t) * H + c) * 128 + row][4 floats], c = 16-byte chunk of r|z|n|hn
};

struct GruTcGeom { int wih_lbo, whh_lbo, tmem_cols, gs, os, xst; uint32_t wih_bytes, whh_bytes, x_bytes, h_bytes; };

// warps 0-3 x producers | 4-11 gate warps (two threads per sequence, each half of the hidden units) | 12 MMA issuer |
// 13-15 output store warps (staged rows -> HBM, one full row per warp instruction).  16 warps = 128 registers/thread.
#define GTC_THREADS 512
#define GTC_MMA_WARP 12
#define GTC_STORE_WARP 13
#define GTC_NSTORE 3
// 7 warps drain the staged rows of a step (the gate warps were stalled on the drain with 3).  H = 32: the 3 store warps +
// the 4 x-producer warps once x_{t+1} is staged (2.63 -> 1.89 ms per step).  H = 16: producers stage twice as many x
// columns per gate column and must not be delayed (helping measured 2.5x slower), but one gate thread per sequence is
// enough there, so warps 8-11 are store warps
#define GTC_NSTORE_ALL 7

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}
template <int HC>
__device__ __forceinline__ void tmem_ld_hc(uint32_t taddr, float (&v)[HC]) {
    if constexpr (HC == 16) tmem_ld16(taddr, v); else tmem_ld8(taddr, v);
}

template <int H, int KQM, bool TILED>
__global__ void __launch_bounds__(GTC_THREADS) gru_fwd_tc_kernel(const __grid_constant__ CUtensorMap hmap, const GruTcArgs a, const GruTcGeom geo) {
    // H = 32: 8 gate warps (two threads per sequence); H = 16: 4 gate warps (one thread per sequence) — warps 8-11 join
    // the store warps instead, because the drain of the staged gate rows is what the gate warps wait for
    constexpr int NGATE = (H == 32) ? 8 : 4;
    // H = 16, TILED (training): the drain is one thread, so warps 8-11 are free and join the x producers — at I = 64 four
    // warps needed 5.3 k cycles per step for x_{t+1} (16 loads + 32 shared stores per thread) and the MMA warp waited for them
    constexpr int NPROD = (TILED && H == 16) ? 256 : 128;
    constexpr int PJ = KQM * 128 / NPROD;
    constexpr int HC = H / (NGATE / 4);                               // hidden units per gate thread
    extern __shared__ __align__(128) unsigned char gsm[];
    const int dir = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int I = a.I, T = a.T, KQ = I >> 2;
    unsigned char* Wih_hi = gsm;
    unsigned char* Wih_lo = Wih_hi + geo.wih_bytes;
    unsigned char* Whh_hi = Wih_lo + geo.wih_bytes;
    unsigned char* Whh_lo = Whh_hi + geo.whh_bytes;
    unsigned char* Hs_hi = Whh_lo + geo.whh_bytes;
    unsigned char* Hs_lo = Hs_hi + geo.h_bytes;
    unsigned char* Xs = Hs_lo + geo.h_bytes;                         // 2 stages x (hi | lo)
    unsigned char* Gs = Xs + 2 * (size_t)geo.xst * geo.x_bytes;                // [128][gs] staged gate rows r|z|n|hn
    // TILED (training): Gs is CHUNK-major, [(chunk c of r|z|n|hn) * 128 + row][16 B] = exactly the 128 x 4H block of the tiled
    // gate layout, so ONE bulk copy drains it; Os is a dense [128][H] tile in the TMA swizzle of its row width (1024-byte
    // aligned), drained by ONE tensor store.  Otherwise: padded row-major tiles drained by the store warps.
    unsigned char* Os = TILED ? reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(Gs + 128 * (size_t)geo.gs) + 1023) & ~(uintptr_t)1023)
                              : Gs + 128 * (size_t)geo.gs;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(Os + 128 * (size_t)geo.os);
    // mbar: [0,2) x_full (128), [2,4) x_empty (1), [4,6) acc_full (1), [6] h_ready (gate threads), [7] stage_full (gate threads), [8] stage_free (7 draining warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 10);
    float* bs = reinterpret_cast<float*>(tmem_slot + 4);             // [4H]: b_ir+b_hr | b_iz+b_hz | b_in | b_hn
    int* lens_s = reinterpret_cast<int*>(bs + 4 * H);                // [128] valid length of every row, -1 outside the batch
    const int s0 = blockIdx.x * 128;
    const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
#define GTC_STAMP(role, step, slot) do { if (dbg_on) a.dbg[((role) * a.T + (step)) * 4 + (slot)] = clock64(); } while (0)

    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), geo.tmem_cols);
    if (tid == 32) {
        mbar_init(smem_u32(mbar + 0), NPROD); mbar_init(smem_u32(mbar + 1), NPROD);
        mbar_init(smem_u32(mbar + 2), 1); mbar_init(smem_u32(mbar + 3), 1);
        mbar_init(smem_u32(mbar + 4), 1); mbar_init(smem_u32(mbar + 5), 1);
        mbar_init(smem_u32(mbar + 6), NGATE * 32); mbar_init(smem_u32(mbar + 7), NGATE * 32); mbar_init(smem_u32(mbar + 8), TILED ? 1 : GTC_NSTORE_ALL * 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // weights: canonical K-major B operands, row n at n*16 B, K-chunk (4 floats) stride lbo
    {
        const float* wi = a.Wih[dir];
        for (int i = tid; i < 3 * H * I; i += GTC_THREADS) {
            const int n = i / I, k = i - n * I;
            float hi, lo;
            split_tf32(__ldg(wi + i), hi, lo);
            const uint32_t off = (uint32_t)n * 16 + (uint32_t)(k >> 2) * geo.wih_lbo + (k & 3) * 4;
            *reinterpret_cast<float*>(Wih_hi + off) = hi;
            *reinterpret_cast<float*>(Wih_lo + off) = lo;
        }
        const float* wh = a.Whh[dir];
        for (int i = tid; i < 3 * H * H; i += GTC_THREADS) {
            const int n = i / H, k = i - n * H;
            float hi, lo;
            split_tf32(__ldg(wh + i), hi, lo);
            const uint32_t off = (uint32_t)n * 16 + (uint32_t)(k >> 2) * geo.whh_lbo + (k & 3) * 4;
            *reinterpret_cast<float*>(Whh_hi + off) = hi;
            *reinterpret_cast<float*>(Whh_lo + off) = lo;
        }
        for (int i = tid; i < 4 * H; i += GTC_THREADS) {
            const int g = i / H, j = i - g * H;
            float v;
            if (g < 2) v = __ldg(a.bih[dir] + g * H + j) + __ldg(a.bhh[dir] + g * H + j);
            else if (g == 2) v = __ldg(a.bih[dir] + 2 * H + j);
            else v = __ldg(a.bhh[dir] + 2 * H + j);
            bs[i] = v;
        }
        for (int i = tid; i < 128; i += GTC_THREADS) lens_s[i] = (s0 + i < a.S) ? (a.len ? __ldg(a.len + s0 + i) : T) : -1;
        // h_0 = 0
        for (uint32_t i = tid * 16u; i < 2 * geo.h_bytes; i += GTC_THREADS * 16u)
            *reinterpret_cast<float4*>(Hs_hi + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_uniform(tmem_slot);
    const uint32_t bar_xfull = smem_u32(mbar), bar_xempty = smem_u32(mbar + 2), bar_acc = smem_u32(mbar + 4),
                   bar_h = smem_u32(mbar + 6), bar_sfull = smem_u32(mbar + 7), bar_sfree = smem_u32(mbar + 8);
    const bool want_g = a.Gt[dir] != nullptr, want_o = a.Hout != nullptr;

    // ===================== staged rows -> HBM (store warps + x producers) =====================
        // ===================== output store warps: staged rows -> HBM =====================
        // A gate row (4H floats) / an output row (H floats) is contiguous in HBM, so LG = H (resp. H/4) lanes move
        // one row with one 16-byte access each (full 128-byte lines); the warp instructions of a step are dealt
        // round-robin to the store warps.
        float* Gt = a.Gt[dir];
        constexpr int LG = H, RG = 32 / LG, NG = 128 / RG;       // lanes per gate row, rows / warp instruction, instructions / step
        constexpr int LO = H / 4, RO = 32 / LO, NO = 128 / RO;
        constexpr int CH = 8;
    auto store_step = [&](int step, int sw) {
            const int t = dir ? (T - 1 - step) : step;
            if (warp == GTC_STORE_WARP && lane == 0) GTC_STAMP(3, step, 0);
            mbar_wait(bar_sfull, (uint32_t)(step & 1));
            if (warp == GTC_STORE_WARP && lane == 0) GTC_STAMP(3, step, 1);
            if (want_g && TILED) {
                // chunk-major tiles: one warp instruction = one 16-byte chunk of 32 consecutive rows (512 B contiguous)
                float* gtile = Gt + ((size_t)blockIdx.x * T + t) * H * 512;
                for (int q0 = sw; q0 < 4 * H; q0 += GTC_NSTORE_ALL * CH) {
                    float4 v[CH];
#pragma unroll
                    for (int k = 0; k < CH; k++) {
                        const int q = q0 + k * GTC_NSTORE_ALL;
                        const int c = q >> 2, r = (q & 3) * 32 + lane;
                        if (q < 4 * H) v[k] = *reinterpret_cast<const float4*>(Gs + (size_t)r * geo.gs + c * 16);
                    }
#pragma unroll
                    for (int k = 0; k < CH; k++) {
                        const int q = q0 + k * GTC_NSTORE_ALL;
                        const int c = q >> 2, r = (q & 3) * 32 + lane;
                        if (q < 4 * H && t < lens_s[r]) *reinterpret_cast<float4*>(gtile + ((size_t)c * 128 + r) * 4) = v[k];
                    }
                }
            } else if (want_g && !TILED) {
                for (int c = sw; c < NG; c += GTC_NSTORE_ALL * CH) {
                    float4 v[CH];
#pragma unroll
                    for (int k = 0; k < CH; k++) {
                        const int q = c + k * GTC_NSTORE_ALL;
                        const int r = q * RG + lane / LG;
                        if (q < NG) v[k] = *reinterpret_cast<const float4*>(Gs + (size_t)r * geo.gs + (lane % LG) * 16);
                    }
#pragma unroll
                    for (int k = 0; k < CH; k++) {
                        const int q = c + k * GTC_NSTORE_ALL;
                        const int r = q * RG + lane / LG;
                        if (q < NG) {
                            const int rl = lens_s[r];
                            if (t < rl) *reinterpret_cast<float4*>(Gt + ((size_t)(s0 + r) * T + t) * 4 * H + (lane % LG) * 4) = v[k];
                        }
                    }
                }
            }
            if (want_o) {
                for (int c = sw; c < NO; c += GTC_NSTORE_ALL * CH) {
                    float4 v[CH];
#pragma unroll
                    for (int k = 0; k < CH; k++) {
                        const int q = c + k * GTC_NSTORE_ALL;
                        const int r = q * RO + lane / LO;
                        if (q < NO) v[k] = *reinterpret_cast<const float4*>(Os + (size_t)r * geo.os + (lane % LO) * 16);
                    }
#pragma unroll
                    for (int k = 0; k < CH; k++) {
                        const int q = c + k * GTC_NSTORE_ALL;
                        const int r = q * RO + lane / LO;
                        if (q < NO && lens_s[r] >= 0)
                            *reinterpret_cast<float4*>(a.Hout + ((size_t)(s0 + r) * T + t) * 2 * H + dir * H + (lane % LO) * 4) = v[k];
                    }
                }
            }
            mbar_arrive(bar_sfree);                       // staged rows consumed: the gate warps may refill them
            if (warp == GTC_STORE_WARP && lane == 0) GTC_STAMP(3, step, 2);
    };
    if (warp < 4 || (NPROD == 256 && warp >= 8 && warp < 12)) {
        const int ptid = warp < 4 ? tid : tid - 128;
        // ===================== x_t producers =====================
        // stage x_{step+1} (registers -> hi/lo split -> shared) while the global loads of x_{step+2} are in flight
        float4 pre[PJ];
        auto load_regs = [&](int step) {
            const int t = dir ? (T - 1 - step) : step;
#pragma unroll
            for (int j = 0; j < PJ; j++) {
                const int i = ptid + j * NPROD;
                const int row = i / KQ, kq = i - row * KQ;
                const int s = s0 + row;
                pre[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < 128 * KQ && s < a.S) pre[j] = __ldg(reinterpret_cast<const float4*>(a.X + (size_t)s * a.x_ss + (size_t)t * a.x_st + kq * 4));
            }
        };
        auto stage_x = [&](int step) {
            const int st = step % geo.xst;
            mbar_wait(bar_xempty + 8u * st, (uint32_t)(((step / geo.xst) & 1) ^ 1));
            unsigned char* X_hi = Xs + (size_t)st * 2 * geo.x_bytes;
            unsigned char* X_lo = X_hi + geo.x_bytes;
#pragma unroll
            for (int j = 0; j < PJ; j++) {
                const int i = ptid + j * NPROD;
                const int row = i / KQ, kq = i - row * KQ;
                if (i >= 128 * KQ) continue;
                float4 hi, lo;
                split_tf32x4(pre[j], hi, lo);
                const uint32_t off = (uint32_t)row * 16 + (uint32_t)kq * TC_A_LBO;
                *reinterpret_cast<float4*>(X_hi + off) = hi;
                *reinterpret_cast<float4*>(X_lo + off) = lo;
            }
            fence_async_smem();
            mbar_arrive(bar_xfull + 8u * st);
        };
        const bool help = !TILED && (H == 32) && (want_g || want_o);
        load_regs(0);
        stage_x(0);
        if (T > 1) load_regs(1);
        for (int step = 0; step + 1 < T; step++) {
            if (warp == 0 && lane == 0) GTC_STAMP(4, step, 0);
            stage_x(step + 1);
            if (warp == 0 && lane == 0) GTC_STAMP(4, step, 1);
            if (step + 2 < T) load_regs(step + 2);
            if (help) store_step(step, GTC_NSTORE + warp);         // x_{step+1} is staged: help draining step's rows
            if (warp == 0 && lane == 0) GTC_STAMP(4, step, 2);
        }
        if (help) store_step(T - 1, GTC_NSTORE + warp);
    } else if (warp < 4 + NGATE) {
        // ===================== gate warps: two threads (H = 32) / one thread (H = 16) per sequence =====================
        const int ew = warp & 3, row = ew * 32 + lane, half = (warp - 4) >> 2;
        const int j0 = half * HC;
        const int s = s0 + row;
        const int len = (s < a.S) ? (a.len ? __ldg(a.len + s) : T) : 0;
        float h[HC];
#pragma unroll
        for (int j = 0; j < HC; j++) h[j] = 0.f;
        float* grow = reinterpret_cast<float*>(Gs + (size_t)row * geo.gs);
        float* orow = reinterpret_cast<float*>(Os + (size_t)row * geo.os);
        const int ox = (H == 32) ? (row & 7) : ((row >> 1) & 3);   // TMA swizzle of the output tile: chunk c of row r sits at c ^ ox
        mbar_arrive(bar_h);                                   // h_0 (zeros) is in place
        for (int step = 0; step < T; step++) {
            const int t = dir ? (T - 1 - step) : step;
            const int buf = step & 1;
            const bool valid = t < len;
            if (warp == 4 && lane == 0) GTC_STAMP(0, step, 0);
            mbar_wait(bar_acc + 8u * buf, (uint32_t)((step >> 1) & 1));
            tc_fence_after();
            if (warp == 4 && lane == 0) GTC_STAMP(0, step, 1);
            const uint32_t trow = tmem + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * 4 * H + j0);
            float vr[HC], vz[HC], vn[HC], vh[HC];
            tmem_ld_hc<HC>(trow, vr);
            tmem_ld_hc<HC>(trow + H, vz);
            tmem_ld_hc<HC>(trow + 2 * H, vn);
            tmem_ld_hc<HC>(trow + 3 * H, vh);
#pragma unroll
            for (int j = 0; j < HC; j++) {
                const int jj = j0 + j;
                const float r = sigmoid_f(vr[j] + bs[jj]);
                const float z = sigmoid_f(vz[j] + bs[H + jj]);
                const float hn = vh[j] + bs[3 * H + jj];
                const float n = tanh_f(vn[j] + bs[2 * H + jj] + r * hn);
                const float hnew = (1.0f - z) * n + z * h[j];
                h[j] = valid ? hnew : h[j];
                vr[j] = r; vz[j] = z; vn[j] = n; vh[j] = hn;
            }
            if (warp == 4 && lane == 0) GTC_STAMP(0, step, 2);
            // publish h_t as the A operand of the next step (hi / lo planes) FIRST: the next step's h-part MMAs are the
            // serial chain of the layer, the staging of this step's gates / output rows below then overlaps them
#pragma unroll
            for (int q = 0; q < HC / 4; q++) {
                const float4 v = make_float4(h[q * 4], h[q * 4 + 1], h[q * 4 + 2], h[q * 4 + 3]);
                float4 hi, lo;
                split_tf32x4(v, hi, lo);
                const uint32_t off = (uint32_t)row * 16 + (uint32_t)(j0 / 4 + q) * TC_A_LBO;
                *reinterpret_cast<float4*>(Hs_hi + off) = hi;
                *reinterpret_cast<float4*>(Hs_lo + off) = lo;
            }
            fence_async_smem();
            tc_fence_before();
            mbar_arrive(bar_h);
            if (warp == 4 && lane == 0) GTC_STAMP(0, step, 3);
            if (want_g || want_o) {
                // the staging rows of the previous step must have left shared memory
                if (step > 0) mbar_wait(bar_sfree, (uint32_t)((step - 1) & 1));
                if (want_g && TILED) {
#pragma unroll
                    for (int q = 0; q < HC / 4; q++) {
                        float4* g4 = reinterpret_cast<float4*>(Gs) + (size_t)((j0 >> 2) + q) * 128 + row;      // chunk-major: 32 rows = 512 contiguous bytes
                        g4[0] = make_float4(vr[q * 4], vr[q * 4 + 1], vr[q * 4 + 2], vr[q * 4 + 3]);
                        g4[(H / 4) * 128] = make_float4(vz[q * 4], vz[q * 4 + 1], vz[q * 4 + 2], vz[q * 4 + 3]);
                        g4[2 * (H / 4) * 128] = make_float4(vn[q * 4], vn[q * 4 + 1], vn[q * 4 + 2], vn[q * 4 + 3]);
                        g4[3 * (H / 4) * 128] = make_float4(vh[q * 4], vh[q * 4 + 1], vh[q * 4 + 2], vh[q * 4 + 3]);
                    }
                } else if (want_g) {
#pragma unroll
                    for (int q = 0; q < HC / 4; q++) {
                        *reinterpret_cast<float4*>(grow + j0 + q * 4) = make_float4(vr[q * 4], vr[q * 4 + 1], vr[q * 4 + 2], vr[q * 4 + 3]);
                        *reinterpret_cast<float4*>(grow + H + j0 + q * 4) = make_float4(vz[q * 4], vz[q * 4 + 1], vz[q * 4 + 2], vz[q * 4 + 3]);
                        *reinterpret_cast<float4*>(grow + 2 * H + j0 + q * 4) = make_float4(vn[q * 4], vn[q * 4 + 1], vn[q * 4 + 2], vn[q * 4 + 3]);
                        *reinterpret_cast<float4*>(grow + 3 * H + j0 + q * 4) = make_float4(vh[q * 4], vh[q * 4 + 1], vh[q * 4 + 2], vh[q * 4 + 3]);
                    }
                }
                if (want_o) {
#pragma unroll
                    for (int q = 0; q < HC / 4; q++) {
                        const int oc = TILED ? (((j0 >> 2) + q) ^ ox) : ((j0 >> 2) + q);
                        *reinterpret_cast<float4*>(orow + oc * 4) =
                            valid ? make_float4(h[q * 4], h[q * 4 + 1], h[q * 4 + 2], h[q * 4 + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                fence_async_smem();
                mbar_arrive(bar_sfull);
            }
            if (warp == 4 && lane == 0) GTC_STAMP(1, step, 0);
        }
        if (a.Hn && s < a.S) {
#pragma unroll
            for (int q = 0; q < HC / 4; q++)
                *reinterpret_cast<float4*>(a.Hn + (size_t)s * 2 * H + dir * H + j0 + q * 4) = make_float4(h[q * 4], h[q * 4 + 1], h[q * 4 + 2], h[q * 4 + 3]);
        }
    } else if (warp == GTC_MMA_WARP) {
        {
            // ===================== MMA issuer: the whole warp runs the loop converged, one elected lane issues =====================
            // (inside an `if (lane == 0)` region every tcgen05.mma is wrapped in an ELECT / BRA.U.ANY waterfall loop)
            const uint32_t id_x = umma_idesc_tf32(3 * H, 0, 0), id_rz = umma_idesc_tf32(2 * H, 0, 0), id_n = umma_idesc_tf32(H, 0, 0);
            const uint32_t wih_hi = smem_u32(Wih_hi), wih_lo = smem_u32(Wih_lo), whh_hi = smem_u32(Whh_hi), whh_lo = smem_u32(Whh_lo);
            const uint32_t hs_hi = smem_u32(Hs_hi), hs_lo = smem_u32(Hs_lo);
            auto issue_x = [&](int step) {
                const int st = step % geo.xst;
                mbar_wait(bar_xfull + 8u * st, (uint32_t)((step / geo.xst) & 1));
                tc_fence_after();
                const uint32_t x_hi = smem_u32(Xs + (size_t)st * 2 * geo.x_bytes), x_lo = x_hi + geo.x_bytes;
                const uint32_t acc = tmem + (uint32_t)((step & 1) * 4 * H);
                if (elect_one_sync()) {
                    for (int ks = 0; ks < (I >> 3); ks++) {
                        const uint32_t ao = (uint32_t)ks * 2 * TC_A_LBO, wo = (uint32_t)ks * 2 * geo.wih_lbo;
                        const uint64_t dah = umma_desc(x_hi + ao, TC_A_LBO, 128), dal = umma_desc(x_lo + ao, TC_A_LBO, 128);
                        const uint64_t dbh = umma_desc(wih_hi + wo, geo.wih_lbo, 128), dbl = umma_desc(wih_lo + wo, geo.wih_lbo, 128);
                        umma_tf32(acc, dah, dbh, id_x, ks > 0 ? 1u : 0u);
                        umma_tf32(acc, dal, dbh, id_x, 1u);
                        umma_tf32(acc, dah, dbl, id_x, 1u);
                    }
                    umma_commit(bar_xempty + 8u * st);                 // x stage reusable once these MMAs retire
                }
                __syncwarp();
            };
            issue_x(0);
            for (int step = 0; step < T; step++) {
                const int buf = step & 1;
                if (lane == 0) GTC_STAMP(2, step, 0);
                mbar_wait(bar_h, (uint32_t)(step & 1));            // h_{step-1} published (phase step)
                tc_fence_after();
                if (lane == 0) GTC_STAMP(2, step, 1);
                const uint32_t acc = tmem + (uint32_t)(buf * 4 * H);
                if (elect_one_sync()) {
#pragma unroll
                    for (int ks = 0; ks < (H >> 3); ks++) {
                        const uint32_t ao = (uint32_t)ks * 2 * TC_A_LBO, wo = (uint32_t)ks * 2 * geo.whh_lbo;
                        const uint64_t dah = umma_desc(hs_hi + ao, TC_A_LBO, 128), dal = umma_desc(hs_lo + ao, TC_A_LBO, 128);
                        const uint64_t dbh = umma_desc(whh_hi + wo, geo.whh_lbo, 128), dbl = umma_desc(whh_lo + wo, geo.whh_lbo, 128);
                        // r,z gates accumulate on top of the x part
                        umma_tf32(acc, dah, dbh, id_rz, 1u);
                        umma_tf32(acc, dal, dbh, id_rz, 1u);
                        umma_tf32(acc, dah, dbl, id_rz, 1u);
                        // W_hn . h goes to its own columns [3H, 4H)
                        const uint64_t dnh = umma_desc(whh_hi + wo + 2 * H * 16, geo.whh_lbo, 128), dnl = umma_desc(whh_lo + wo + 2 * H * 16, geo.whh_lbo, 128);
                        umma_tf32(acc + 3 * H, dah, dnh, id_n, ks > 0 ? 1u : 0u);
                        umma_tf32(acc + 3 * H, dal, dnh, id_n, 1u);
                        umma_tf32(acc + 3 * H, dah, dnl, id_n, 1u);
                    }
                    umma_commit(bar_acc + 8u * buf);
                }
                __syncwarp();
                if (lane == 0) GTC_STAMP(2, step, 2);
                if (step + 1 < T) issue_x(step + 1);
                if (lane == 0) GTC_STAMP(2, step, 3);
            }
        }
    }
    if (TILED) {
        // ===================== drain (one thread): the staged gate block and output tile leave through the copy engine =====================
        if (tid == GTC_STORE_WARP * 32 && (want_g || want_o)) {
            for (int step = 0; step < T; step++) {
                const int t = dir ? (T - 1 - step) : step;
                GTC_STAMP(3, step, 0);
                mbar_wait(bar_sfull, (uint32_t)(step & 1));        // the gate threads fenced their writes for the async proxy
                GTC_STAMP(3, step, 1);
                if (want_g)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                 ::"l"(a.Gt[dir] + ((size_t)blockIdx.x * T + t) * H * 512), "r"(smem_u32(Gs)), "r"((uint32_t)(128 * 4 * H * 4)) : "memory");
                if (want_o)
                    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                                 ::"l"(&hmap), "r"(smem_u32(Os)), "r"(dir * H), "r"(t), "r"(s0) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the staging tiles have been read
                mbar_arrive(bar_sfree);
                GTC_STAMP(3, step, 2);
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");               // writes complete before the CTA exits
        }
    } else {
        const int swi = warp >= GTC_STORE_WARP ? warp - GTC_STORE_WARP : ((H == 16 && warp >= 4 + NGATE && warp < GTC_MMA_WARP) ? GTC_NSTORE + warp - (4 + NGATE) : -1);
        if (swi >= 0 && (want_g || want_o))
            for (int step = 0; step < T; step++) store_step(step, swi);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, geo.tmem_cols);
}

static bool gru_tc_geom(int H, int I, GruTcGeom& g, size_t& smem, bool tiled = false);
static bool gru_tc_eligible(int S, int H, int I) {
    (void)S;   // independent of the batch size: chunked and unchunked batches must take the same arithmetic path
    GruTcGeom g; size_t smem;
    return tc_enabled() && (H == 16 || H == 32) && (I % 8 == 0) && I >= 8 && I <= 64 && gru_tc_geom(H, I, g, smem);
}

static bool gru_tc_geom(int H, int I, GruTcGeom& g, size_t& smem, bool tiled) {
    g.wih_lbo = 3 * H * 16 + 16;
    g.whh_lbo = 3 * H * 16 + 16;
    g.wih_bytes = (uint32_t)(I / 4) * g.wih_lbo;
    g.whh_bytes = (uint32_t)(H / 4) * g.whh_lbo;
    g.x_bytes = (uint32_t)(I / 4) * TC_A_LBO;
    g.h_bytes = (uint32_t)(H / 4) * TC_A_LBO;
    g.tmem_cols = tmem_cols_for(8 * H);
    g.gs = 4 * H * 4 + (tiled ? 0 : 16);     // row-major staging: padded row strides (conflict-free 16-byte stores); tiled: dense
    g.os = H * 4 + (tiled ? 0 : 16);
    for (g.xst = 2; g.xst >= 1; g.xst--) {
        smem = 2 * (size_t)g.wih_bytes + 2 * (size_t)g.whh_bytes + 2 * (size_t)g.h_bytes + 2 * (size_t)g.xst * g.x_bytes +
               128 * (size_t)(g.gs + g.os) + 10 * 8 + 16 + (size_t)4 * H * 4 + 128 * 4 + 128 + (tiled ? 1024 : 0);
        if (smem <= 227 * 1024) return true;
    }
    return false;
}

template <int H, int KQM, bool TILED>
static int gru_tc_launch_tt(const CUtensorMap& hmap, const GruTcArgs& a, const GruTcGeom& geo, size_t smem, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        DOF_CUDA(cudaFuncSetAttribute(gru_fwd_tc_kernel<H, KQM, TILED>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr = true;
    }
    dim3 grid(cdiv(a.S, 128), 2);
    gru_fwd_tc_kernel<H, KQM, TILED><<<grid, GTC_THREADS, smem, st>>>(hmap, a, geo);
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}
template <int H, int KQM>
static int gru_tc_launch_t(const CUtensorMap& hmap, const GruTcArgs& a, const GruTcGeom& geo, size_t smem, cudaStream_t st) {
    return a.gt_tiled ? gru_tc_launch_tt<H, KQM, true>(hmap, a, geo, smem, st) : gru_tc_launch_tt<H, KQM, false>(hmap, a, geo, smem, st);
}

// one bidirectional GRU layer, both directions (blockIdx.y)
static int launch_gru_fwd_tc(const GruTcArgs& a, cudaStream_t st) {
    GruTcGeom geo;
    size_t smem = 0;
    if (!gru_tc_geom(a.H, a.I, geo, smem, a.gt_tiled != 0)) DOF_FAIL(DOF_ERR_UNSUPPORTED, "fused GRU tile does not fit (H=%d I=%d)", a.H, a.I);
    CUtensorMap hmap;
    memset(&hmap, 0, sizeof(hmap));
    if (a.gt_tiled) {
        if (!a.Gt[0] || !a.Gt[1] || !aligned16(a.Gt[0]) || !aligned16(a.Gt[1])) DOF_FAIL(DOF_ERR_ARG, "tiled gates need both 16-byte aligned gate buffers");
        if (a.Hout) {
            if (!aligned16(a.Hout)) DOF_FAIL(DOF_ERR_ARG, "fused GRU output must be 16-byte aligned");
            DOF_TRY(tmap_seq3d(a.Hout, a.S, a.T, a.H, 128, &hmap));
        }
    }
    if ((a.x_ss & 3) || (a.x_st & 3) || !aligned16(a.X)) DOF_FAIL(DOF_ERR_ARG, "fused GRU input must be 16-byte aligned");
    const double rows = (double)a.S * a.T * 2;
    ProfScope ps(a.H == 32 ? "gru_fwd_tc_h32" : "gru_fwd_tc_h16", st, rows * 2.0 * 3 * a.H * (a.I + a.H),
                 (double)a.S * a.T * 4.0 * a.I + rows * 4.0 * a.H * ((a.Hout ? 1 : 0) + (a.Gt[0] ? 4 : 0)));
    const int KQ = a.I / 4;
    if (a.H == 32) {
        if (KQ <= 4) return gru_tc_launch_t<32, 4>(hmap, a, geo, smem, st);
        if (KQ <= 8) return gru_tc_launch_t<32, 8>(hmap, a, geo, smem, st);
        return gru_tc_launch_t<32, 16>(hmap, a, geo, smem, st);
    }
    if (KQ <= 4) return gru_tc_launch_t<16, 4>(hmap, a, geo, smem, st);
    if (KQ <= 8) return gru_tc_launch_t<16, 8>(hmap, a, geo, smem, st);
    return gru_tc_launch_t<16, 16>(hmap, a, geo, smem, st);
}
