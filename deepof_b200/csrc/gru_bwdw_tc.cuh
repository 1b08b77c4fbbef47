// Fused GRU layer backward on tcgen05, SECOND generation: BPTT + input gradient + ALL FOUR PARAMETER GRADIENTS in one
// kernel, so the gate gradients dG (4H floats per row and direction) never reach HBM.
//
// Replaces gru_bwd_tc_kernel + gru_wgrad_tc_kernel for the layers whose input changes per step (reference: autograd of
// torch.nn.GRU inside RecurrentBlockPT / RecurrentDecoderPT, deepof/clustering/models_new.py:217-278, 326-373).
//
// Per step (reverse of the forward order) and per tile of R sequences, one direction per CTA:
//   gate warps (2 threads / sequence): dG_t = [da_r | da_z | da_n r | da_n] -> ONE shared-memory tile (hi / lo planes)
//   x-loader warps: x_t rows -> XH tile columns [0, I);  gate threads: h_prev rows -> XH tile columns [I, I+H)
//   tensor core, issuer A:  acc_dh = dG_t[:, 0:3H) . W_hh   (-> dh_{t-1}),   acc_dx = [da_r|da_z|da_n] . W_ih  (-> dX_t)
//   tensor core, issuer B:  acc_w[4H lanes][I | H | 1] += dG_t^T . [x_t | h_prev | 1]      (kept in TMEM for all T steps)
//   epilogue (once per CTA): acc_w -> dW_ih, dW_hh, db_ih, db_hh with vector reductions.
//
// The dG tile is used in BOTH operand roles from one copy: rows are 128 bytes (32 columns) in the SWIZZLE_128B_BASE32B
// arrangement (32-byte units XOR row & 3), which tcgen05 reads K-major (M = sequence, K = gate column: dh / dX) and
// MN-major (M = gate column, K = sequence: the weight gradients) — tools/probe_kmajor_sw32.cu pins both readings.
// All products are 3xTF32 (hi.hi + lo.hi + hi.lo, lo = exact remainder); the bias column multiplies an all-ones tile.
//
// Shared memory at H = 32, I = 32: dG 2 x 4 blocks + XH 2 x 2 blocks of R x 128 B, the two TMA staging tiles (h_prev, dOut)
// of R x 128 B, weights 50 KB -> R = 96 rows per tile; H = 16 uses R = 128.
#pragma once
#include "common.cuh"
#include "tc_gemm.cuh"
#include "gru_tc.cuh"
#include "gru_bwd_tc.cuh"
#include "gru_wgrad_tc.cuh"

struct GruBwdwArgs {
    const float* Whh[2]; const float* Wih[2];
    const int* len;           // [S] or null
    const float* X; long long x_ss; int x_st;   // layer input: x_t of sequence s at X + s*x_ss + t*x_st (I floats)
    const float* Hout;        // [S,T,2H] forward outputs (h_prev source)
    const float* GtT[2];      // saved gates, tiled layout of gru_fwd_tc_kernel (tiles of 128 sequences)
    const float* dOut;        // [S,T,2H] or null
    const float* dHn;         // [S,2H] or null
    float* dX;                // [S,T,I], zeroed by the caller; both directions add
    const float* dXmask;      // [S,T,I] or null: dX is kept only where mask > 0 (ReLU backward)
    float* dWih[2]; float* dWhh[2]; float* dbih[2]; float* dbhh[2];   // accumulated (+=)
    int S, T, H, I;
    long long* dbg;           // null, or [8 roles][T][4] clock64 stamps of CTA (0, 0) (tools/prof_gru_bwdw.py)
};

struct GruBwdwGeom { int whh_lbo, wih_lbo, tmem_cols; uint32_t whh_bytes, wih_bytes, dg_bytes, xh_bytes, st_bytes; };
// TMA views of Hout and dOut: [2H columns (contiguous), T steps, S sequences]; box = [H columns of one direction, 1 step, R sequences]
struct GruBwdwMaps { CUtensorMap Hout, dOut; };

#define GBW_THREADS 512          // warps 0-3 x loaders + dX drain + epilogue | 4-11 gate warps | 12 MMA issuer A (dh, dX) |
                                 // 13 MMA issuer B (weight gradients) | 14-15 x loaders
#define GBW_MMA_A 12
#define GBW_MMA_B 13
#define GBW_NXL 192              // x-loader threads (warps 0-3, 14, 15)

// K-major operand in the 128B_BASE32B arrangement: the swizzle repeats every FOUR rows and SBO is the stride between
// 4-row groups (512 B for rows packed at 128 B; tools/probe_kmajor_readout.cu read the addressing back from the tensor
// core: row m, K unit u -> m*128 + ((u ^ (m & 3)) * 32)); a K step of 8 columns = start address + 32 B; LBO is not used
__device__ __forceinline__ uint64_t umma_desc_k32(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(512u >> 4) << 32) | (1ull << 46) | (1ull << 61);
}
// byte offset of the 32-byte unit that holds columns [c8, c8 + 8) of row r in a tile of R rows
template <int R>
__device__ __forceinline__ uint32_t sw32_unit(int r, int c8) {
    return (uint32_t)(c8 >> 5) * (uint32_t)(R * 128) + (uint32_t)r * 128 + (uint32_t)((((c8 & 31) >> 3) ^ (r & 3)) << 5);
}
// the two 16-byte halves of one unit, hi / lo planes.  `sel` (= (row >> 2) & 1) swaps the order of the two stores so that
// rows r and r + 4 of a quarter-warp (same unit position) never hit the same banks in the same instruction
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, const float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// the two 16-byte halves of one 32-byte unit, hi / lo planes.  Threads with sel = (row >> 2) & 1 own their column chunks in
// swapped order inside every pair (logical chunk q = physical chunk q ^ sel, see the gate warps), so v0 goes to half `sel`
// and v1 to the other half: rows r and r + 4 of a quarter-warp (same unit position) never hit the same banks in one
// instruction, without any data-dependent select
__device__ __forceinline__ void st_unit(uint32_t hi_s, uint32_t lo_s, uint32_t off, const float (&v0)[4], const float (&v1)[4], int sel) {
    float4 h0, l0, h1, l1;
    split_tf32_exact(v0[0], h0.x, l0.x); split_tf32_exact(v0[1], h0.y, l0.y); split_tf32_exact(v0[2], h0.z, l0.z); split_tf32_exact(v0[3], h0.w, l0.w);
    split_tf32_exact(v1[0], h1.x, l1.x); split_tf32_exact(v1[1], h1.y, l1.y); split_tf32_exact(v1[2], h1.z, l1.z); split_tf32_exact(v1[3], h1.w, l1.w);
    const uint32_t oa = off + (sel ? 16u : 0u), ob = off + (sel ? 0u : 16u);
    sts128(hi_s + oa, h0);
    sts128(lo_s + oa, l0);
    sts128(hi_s + ob, h1);
    sts128(lo_s + ob, l1);
}

// register budget per role (setmaxnreg, per warpgroup): the gate warps keep a whole step of saved gates in flight
// (24 x 16-byte loads per thread at H = 32); 8 x 32 x GATE + 8 x 32 x OTHER = 64 K registers
template <int H> struct GbwRegs { static constexpr int GATE = H == 32 ? 184 : 136, OTHER = H == 32 ? 72 : 120; };

template <int H, int R, int KQM>
__global__ void __launch_bounds__(GBW_THREADS, 1) gru_bwdw_tc_kernel(const __grid_constant__ GruBwdwMaps maps, const GruBwdwArgs a, const GruBwdwGeom geo) {
    constexpr int HC = H / 2;                                         // hidden units per gate thread
    constexpr int NBP = 4 * H / 32;                                   // 32-column blocks of the dG tile
    extern __shared__ unsigned char bw_raw[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(bw_raw) + 1023) & ~(uintptr_t)1023);
    const int dir = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int I = a.I, T = a.T, KQ = I >> 2;
    const bool oxh = ((I + H) & 31) == 16;                           // a constant-one column fits into the XH tile (see below)
    unsigned char* DG_hi = sm;                                        // [NBP blocks][R rows][128 B]
    unsigned char* DG_lo = DG_hi + geo.dg_bytes;
    unsigned char* XH_hi = DG_lo + geo.dg_bytes;                      // columns [0, I) x_t | [I, I+H) h_prev
    unsigned char* XH_lo = XH_hi + geo.xh_bytes;
    // h_prev / dOut tiles of ONE step, [R rows][H floats], written by TMA (hardware swizzle: 16-byte chunk c of row r sits at
    // chunk c ^ (r & 7) for 128-byte rows, c ^ ((r >> 1) & 3) for 64-byte rows): a row-per-lane global load touches 32
    // lines per warp instruction and the gate warps spent a third of every step issuing them
    unsigned char* ST_hp = XH_lo + geo.xh_bytes;
    unsigned char* ST_do = ST_hp + geo.st_bytes;
    unsigned char* Whh_hi = ST_do + geo.st_bytes;                     // B operand [N = H][K = 3H]:  W_hh[c][j] at (n = j, k = c)
    unsigned char* Whh_lo = Whh_hi + geo.whh_bytes;
    unsigned char* Wih_hi = Whh_lo + geo.whh_bytes;                   // B operand [N = I][K = 3H]:  W_ih[c][i] at (n = i, k = c)
    unsigned char* Wih_lo = Wih_hi + geo.wih_bytes;
    float* ones = reinterpret_cast<float*>(Wih_lo + geo.wih_bytes);   // B operand [N = 16][K = 8] of 1.0f: every K step reads it
    uint64_t* mbar = reinterpret_cast<uint64_t*>(ones + 128);
    // mbar: [0] a_full (gate threads + x loaders) [1] dh_full (commit A) [2] dx_full (commit A) [3] w_done (commit B)
    //       [4] dx_read (128 drain threads: acc_dx read out of TMEM) [5] st_full (TMA bytes) [6] st_free (256 gate threads)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 8);
    int* lens_s = reinterpret_cast<int*>(tmem_slot + 4);              // [128]; -1 outside the tile / batch
    const int s0 = blockIdx.x * R;
    const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
#define GBW_STAMP(role, step, slot) do { if (dbg_on) a.dbg[((role) * T + (step)) * 4 + (slot)] = clock64(); } while (0)

    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), geo.tmem_cols);
    if (tid == 32) {
        mbar_init(smem_u32(mbar + 0), 256 + GBW_NXL); mbar_init(smem_u32(mbar + 1), 1);
        mbar_init(smem_u32(mbar + 2), 1); mbar_init(smem_u32(mbar + 3), 1); mbar_init(smem_u32(mbar + 4), 128);
        mbar_init(smem_u32(mbar + 5), 1); mbar_init(smem_u32(mbar + 6), 256);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const float* wh = a.Whh[dir];
        for (int i = tid; i < 3 * H * H; i += GBW_THREADS) {
            const int c = i / H, j = i - c * H;                       // W_hh[c][j]
            float hi, lo;
            split_tf32(__ldg(wh + i), hi, lo);
            const uint32_t off = (uint32_t)j * 16 + (uint32_t)(c >> 2) * geo.whh_lbo + (c & 3) * 4;
            *reinterpret_cast<float*>(Whh_hi + off) = hi;
            *reinterpret_cast<float*>(Whh_lo + off) = lo;
        }
        const float* wi = a.Wih[dir];
        for (int i = tid; i < 3 * H * I; i += GBW_THREADS) {
            const int c = i / I, ii = i - c * I;                      // W_ih[c][ii]
            float hi, lo;
            split_tf32(__ldg(wi + i), hi, lo);
            const uint32_t off = (uint32_t)ii * 16 + (uint32_t)(c >> 2) * geo.wih_lbo + (c & 3) * 4;
            *reinterpret_cast<float*>(Wih_hi + off) = hi;
            *reinterpret_cast<float*>(Wih_lo + off) = lo;
        }
        if (oxh) {
            // columns [I+H, I+H+16) of every XH row: 1, 0, 0, ... (hi plane), zeros (lo plane); written once, nobody else touches them
            for (int i = tid; i < R * 4; i += GBW_THREADS) {
                const int r = i >> 2, q = i & 3;
                const uint32_t off = sw32_unit<R>(r, (I + H + 4 * q) & ~7) + (uint32_t)(q & 1) * 16;
                *reinterpret_cast<float4*>(XH_hi + off) = make_float4(q == 0 ? 1.f : 0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(XH_lo + off) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        for (int i = tid; i < 128; i += GBW_THREADS) {
            ones[i] = 1.0f;
            lens_s[i] = (i < R && s0 + i < a.S) ? (a.len ? __ldg(a.len + s0 + i) : T) : -1;
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_uniform(tmem_slot);
    const uint32_t bar_afull = smem_u32(mbar), bar_dh = smem_u32(mbar + 1), bar_dx = smem_u32(mbar + 2), bar_w = smem_u32(mbar + 3),
                   bar_dxr = smem_u32(mbar + 4), bar_stfull = smem_u32(mbar + 5), bar_stfree = smem_u32(mbar + 6);
    // TMA of the h_prev / dOut tiles of `step_` (one thread); a step without either tile still completes the phase
    auto issue_stage = [&](int step_) {
        const int t = dir ? step_ : (T - 1 - step_);
        const int tp = dir ? t + 1 : t - 1;
        const bool hp_on = tp >= 0 && tp < T, do_on = a.dOut != nullptr;
        const uint32_t bytes = (hp_on ? geo.st_bytes : 0u) + (do_on ? geo.st_bytes : 0u);
        if (bytes) mbar_expect_tx(bar_stfull, bytes); else mbar_arrive(bar_stfull);
        if (hp_on) tma_load_3d(smem_u32(ST_hp), &maps.Hout, dir * H, tp, s0, bar_stfull);
        if (do_on) tma_load_3d(smem_u32(ST_do), &maps.dOut, dir * H, t, s0, bar_stfull);
    };
    if (tid == 14 * 32) issue_stage(0);
    // TMEM columns: acc_dh [H] | acc_dx [I] | weight-gradient accumulators.  The TMEM accumulator TRUNCATES on every
    // tcgen05.mma accumulate, a bias that grows with the number of accumulates (T * R/8 * 3 = 1050 at T = 25: 2e-5
    // relative, measured), so the hi.hi products alternate between two accumulators by step parity and the small lo
    // terms go to a third: 175 accumulates each on the large terms; the epilogue adds the three in fp32 registers.
    // When the XH tile has 16 spare columns in its last 32-column block (I + H = 16 mod 32: every H = 16 shape but I = 16) a
    // constant-one column rides along as column I + H of the B operand and the bias gradients come out of the same MMAs
    // (N = I + H + 16); otherwise two extra MMAs per K step multiply dG with a separate all-ones tile.
    const uint32_t NW = (uint32_t)(I + H) + (oxh ? 16u : 0u);
    const uint32_t acc_dh = tmem, acc_dx = tmem + H, acc_w = tmem + H + I, acc_wlo = acc_w + 2 * NW, acc_b = acc_w + 3 * NW, acc_blo = acc_b + 32;

    if (warp >= 4 && warp < GBW_MMA_A) {
        // ===================== gate warps: two threads per sequence =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(GbwRegs<H>::GATE));
        const int ew = warp & 3, row = ew * 32 + lane, half = (warp - 4) >> 2;
        const int j0 = half * HC;
        const int s = s0 + row;
        const int len = lens_s[row] < 0 ? 0 : lens_s[row];
        const int sel = (row >> 2) & 1;
        const bool in_tile = row < R;
        const float* GtT = a.GtT[dir] + (size_t)(s >> 7) * T * H * 512 + (size_t)(s & 127) * 4;     // + (t*H + chunk) * 512
        const uint32_t dg_hi = smem_u32(DG_hi), dg_lo = smem_u32(DG_lo), xh_hi = smem_u32(XH_hi), xh_lo = smem_u32(XH_lo);
        float part[HC];                                               // dh before the recurrent term of the next step
#pragma unroll
        for (int j = 0; j < HC; j++) part[j] = 0.f;
        if (a.dHn && len > 0) {
#pragma unroll
            for (int q = 0; q < HC / 4; q++) {
                const float4 d = __ldg(reinterpret_cast<const float4*>(a.dHn + (size_t)s * 2 * H + dir * H + j0 + (q ^ sel) * 4));
                part[q * 4] = d.x; part[q * 4 + 1] = d.y; part[q * 4 + 2] = d.z; part[q * 4 + 3] = d.w;
            }
        }
        // saved gates / h_prev / dOut of one step: 6 x HC/4 16-byte loads per thread, issued ONE STEP AHEAD (right after the
        // previous values of the same registers were consumed) so that their latency hides behind the stores, the MMAs
        // and the barrier round trip of the current step
        float4 r4[HC / 4], z4[HC / 4], n4[HC / 4], q4[HC / 4], hp4[HC / 4], do4[HC / 4];
        auto load_q = [&](int q, int step_) {
            const int t = dir ? step_ : (T - 1 - step_);
            const float* gt = GtT + (size_t)t * H * 512;
            const int pq = q ^ sel;                                   // physical chunk of logical chunk q
            const int cq = (j0 >> 2) + pq;
            if (t < len) {
                r4[q] = __ldg(reinterpret_cast<const float4*>(gt + (size_t)cq * 512));
                z4[q] = __ldg(reinterpret_cast<const float4*>(gt + (size_t)(H / 4 + cq) * 512));
                n4[q] = __ldg(reinterpret_cast<const float4*>(gt + (size_t)(2 * (H / 4) + cq) * 512));
                q4[q] = __ldg(reinterpret_cast<const float4*>(gt + (size_t)(3 * (H / 4) + cq) * 512));
            } else {
                r4[q] = z4[q] = n4[q] = q4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        // h_prev / dOut chunks of this thread's row from the TMA-staged tiles
        const uint32_t st_hp = smem_u32(ST_hp), st_do = smem_u32(ST_do);
        const uint32_t st_row = (uint32_t)row * (uint32_t)(H * 4);
        const int st_x = (H == 32) ? (row & 7) : ((row >> 1) & 3);
        auto read_stage = [&](int step_) {
            const int t = dir ? step_ : (T - 1 - step_);
            const int tp = dir ? t + 1 : t - 1;
            mbar_wait(bar_stfull, (uint32_t)(step_ & 1));
#pragma unroll
            for (int q = 0; q < HC / 4; q++) {
                const uint32_t off = st_row + (uint32_t)((((j0 >> 2) + (q ^ sel)) ^ st_x) << 4);
                hp4[q] = (t < len && tp >= 0 && tp < len) ? lds128(st_hp + off) : make_float4(0.f, 0.f, 0.f, 0.f);
                do4[q] = (t < len && a.dOut) ? lds128(st_do + off) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            fence_async_smem();          // generic-proxy reads above vs. the TMA (async proxy) that refills the tiles right after this arrive
            mbar_arrive(bar_stfree);
        };
        // DRAM -> L2 two steps ahead (no registers held): one SM can only keep ~30 KB of loads in flight, so a step's 100 KB
        // of saved gates / h_prev / dOut / x must already sit in L2 when the register loads above are issued
        auto prefetch_step = [&](int step_) {
            const int t = dir ? step_ : (T - 1 - step_);
            if (t >= len) return;
            if ((lane & 7) == 0) {                                          // one lane per 128-byte line of the tiled gates
                const float* gn = GtT + ((size_t)t * H + (j0 >> 2)) * 512;
#pragma unroll
                for (int g4 = 0; g4 < 4; g4++)
#pragma unroll
                    for (int qq = 0; qq < HC / 4; qq++)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(gn + (size_t)(g4 * (H / 4) + qq) * 512));
            }
        };
#pragma unroll
        for (int q = 0; q < HC / 4; q++) load_q(q, 0);
        if (H == 32 && T > 1) prefetch_step(1);       // measured: helps H = 32 (1.87 -> 1.81 ms), hurts H = 16 (1.12 -> 1.38 ms)
        for (int step = 0; step < T; step++) {
            const int t = dir ? step : (T - 1 - step);
            const bool valid = t < len;
            if (H == 32 && step + 2 < T) prefetch_step(step + 2);
            read_stage(step);
            // recurrent term of the previous step: dh = part + dG_{prev} . W_hh
            float dh[HC];
            const int grole = warp == 4 ? 0 : (warp == 11 ? 1 : -1);
            if (grole >= 0 && lane == 0) GBW_STAMP(grole, step, 0);
            if (step > 0) {
                mbar_wait(bar_dh, (uint32_t)((step - 1) & 1));
                tc_fence_after();
                float v[HC];
                tmem_ld_hc<HC>(acc_dh + ((uint32_t)(ew * 32) << 16) + (uint32_t)j0, v);
#pragma unroll
                for (int j = 0; j < HC; j++) dh[j] = part[j] + (sel ? v[j ^ 4] : v[j]);       // TMEM columns are in physical order
            } else {
#pragma unroll
                for (int j = 0; j < HC; j++) dh[j] = part[j];
            }
            if (grole >= 0 && lane == 0) GBW_STAMP(grole, step, 1);
            // gate gradients of the whole row segment into registers BEFORE waiting for the tiles: only the stores need them
            float o_r[HC / 4][4], o_z[HC / 4][4], o_h[HC / 4][4], o_n[HC / 4][4];
#pragma unroll
            for (int q = 0; q < HC / 4; q++) {
                const float r[4] = {r4[q].x, r4[q].y, r4[q].z, r4[q].w}, z[4] = {z4[q].x, z4[q].y, z4[q].z, z4[q].w};
                const float n[4] = {n4[q].x, n4[q].y, n4[q].z, n4[q].w}, hn[4] = {q4[q].x, q4[q].y, q4[q].z, q4[q].w};
                const float hp[4] = {hp4[q].x, hp4[q].y, hp4[q].z, hp4[q].w}, dov[4] = {do4[q].x, do4[q].y, do4[q].z, do4[q].w};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    if (valid) {
                        const float d = dh[q * 4 + e] + dov[e];
                        const float dn = d * (1.0f - z[e]);
                        const float dz = d * (hp[e] - n[e]);
                        const float dan = dn * (1.0f - n[e] * n[e]);
                        o_z[q][e] = dz * z[e] * (1.0f - z[e]);
                        o_r[q][e] = dan * hn[e] * r[e] * (1.0f - r[e]);
                        o_h[q][e] = dan * r[e];
                        o_n[q][e] = dan;
                        part[q * 4 + e] = d * z[e];
                    } else {
                        o_z[q][e] = o_r[q][e] = o_h[q][e] = o_n[q][e] = 0.f;
                        part[q * 4 + e] = dh[q * 4 + e];
                    }
                }
            }
            // the saved-gate registers are dead now: next step's loads go out here, a whole MMA round trip ahead of their use
            if (step + 1 < T) {
#pragma unroll
                for (int q = 0; q < HC / 4; q++) load_q(q, step + 1);
            }
            // the tiles are free again once every MMA of the previous step retired
            if (step > 0) {
                mbar_wait(bar_dx, (uint32_t)((step - 1) & 1));
                mbar_wait(bar_w, (uint32_t)((step - 1) & 1));
            }
            if (grole >= 0 && lane == 0) GBW_STAMP(grole, step, 2);
            if (in_tile) {
#pragma unroll
                for (int p = 0; p < HC / 8; p++) {
                    const int c8 = j0 + 8 * p;                               // column inside a gate block
                    const float hp0[4] = {hp4[2 * p].x, hp4[2 * p].y, hp4[2 * p].z, hp4[2 * p].w};
                    const float hp1[4] = {hp4[2 * p + 1].x, hp4[2 * p + 1].y, hp4[2 * p + 1].z, hp4[2 * p + 1].w};
                    st_unit(dg_hi, dg_lo, sw32_unit<R>(row, c8), o_r[2 * p], o_r[2 * p + 1], sel);
                    st_unit(dg_hi, dg_lo, sw32_unit<R>(row, H + c8), o_z[2 * p], o_z[2 * p + 1], sel);
                    st_unit(dg_hi, dg_lo, sw32_unit<R>(row, 2 * H + c8), o_h[2 * p], o_h[2 * p + 1], sel);
                    st_unit(dg_hi, dg_lo, sw32_unit<R>(row, 3 * H + c8), o_n[2 * p], o_n[2 * p + 1], sel);
                    st_unit(xh_hi, xh_lo, sw32_unit<R>(row, I + c8), hp0, hp1, sel);
                }
            }
            fence_async_smem();
            tc_fence_before();
            mbar_arrive(bar_afull);
            if (grole >= 0 && lane == 0) GBW_STAMP(grole, step, 3);
        }
    } else if (warp == GBW_MMA_A) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GbwRegs<H>::OTHER));
        {
            // ===================== MMA issuer A: dh and dX (whole warp converged, one elected lane issues) =====================
            const uint32_t id_h = umma_idesc_tf32(H, 0, 0), id_x = umma_idesc_tf32(I, 0, 0);
            const uint32_t a_hi = smem_u32(DG_hi), a_lo = smem_u32(DG_lo);
            const uint32_t whh_hi = smem_u32(Whh_hi), whh_lo = smem_u32(Whh_lo), wih_hi = smem_u32(Wih_hi), wih_lo = smem_u32(Wih_lo);
            for (int step = 0; step < T; step++) {
                if (lane == 0) GBW_STAMP(2, step, 0);
                // TMA of the NEXT step's h_prev / dOut tiles as soon as the gate threads have copied this step's (they do that at
                // the top of their step, long before a_full): this warp is idle until a_full anyway, and the tiles get a whole
                // step of lead (issued from a loader warp they arrived when they were needed: 4 - 7 k cycles exposed per step)
                if (step + 1 < T) {
                    mbar_wait(bar_stfree, (uint32_t)(step & 1));
                    if (elect_one_sync()) issue_stage(step + 1);
                    __syncwarp();
                }
                mbar_wait(bar_afull, (uint32_t)(step & 1));
                tc_fence_after();
                if (lane == 0) GBW_STAMP(2, step, 1);
                // dh: K = 3H = tile columns [0, 3H)
                if (elect_one_sync()) {
#pragma unroll
                    for (int ks = 0; ks < (3 * H) >> 3; ks++) {
                        const uint32_t ao = (uint32_t)(ks >> 2) * (uint32_t)(R * 128) + (uint32_t)(ks & 3) * 32, wo = (uint32_t)ks * 2 * geo.whh_lbo;
                        const uint64_t dah = umma_desc_k32(a_hi + ao), dal = umma_desc_k32(a_lo + ao);
                        const uint64_t dbh = umma_desc(whh_hi + wo, geo.whh_lbo, 128), dbl = umma_desc(whh_lo + wo, geo.whh_lbo, 128);
                        umma_tf32(acc_dh, dah, dbh, id_h, ks > 0 ? 1u : 0u);
                        umma_tf32(acc_dh, dal, dbh, id_h, 1u);
                        umma_tf32(acc_dh, dah, dbl, id_h, 1u);
                    }
                    umma_commit(bar_dh);
                }
                __syncwarp();
                if (lane == 0) GBW_STAMP(2, step, 2);
                if (step > 0) { mbar_wait(bar_dxr, (uint32_t)((step - 1) & 1)); tc_fence_after(); }   // acc_dx of the previous step read out
                // dX: tile columns [0, 2H) (da_r, da_z) and [3H, 4H) (da_n) against W_ih rows r, z, n
                if (elect_one_sync()) {
#pragma unroll
                    for (int ks = 0; ks < (3 * H) >> 3; ks++) {
                        const int kc = ks < ((2 * H) >> 3) ? ks : ks + (H >> 3);                  // 8-column group of the tile
                        const uint32_t ao = (uint32_t)(kc >> 2) * (uint32_t)(R * 128) + (uint32_t)(kc & 3) * 32, wo = (uint32_t)ks * 2 * geo.wih_lbo;
                        const uint64_t dah = umma_desc_k32(a_hi + ao), dal = umma_desc_k32(a_lo + ao);
                        const uint64_t dbh = umma_desc(wih_hi + wo, geo.wih_lbo, 128), dbl = umma_desc(wih_lo + wo, geo.wih_lbo, 128);
                        umma_tf32(acc_dx, dah, dbh, id_x, ks > 0 ? 1u : 0u);
                        umma_tf32(acc_dx, dal, dbh, id_x, 1u);
                        umma_tf32(acc_dx, dah, dbl, id_x, 1u);
                    }
                    umma_commit(bar_dx);
                }
                __syncwarp();
                if (lane == 0) GBW_STAMP(2, step, 3);
            }
        }
    } else if (warp == GBW_MMA_B) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GbwRegs<H>::OTHER));
        {
            // ===================== MMA issuer B: weight / bias gradients, accumulated over all T steps =====================
            const uint32_t id_w = umma_idesc_tf32((int)NW, 1, 1), id_b = umma_idesc_tf32(16, 1, 0);
            const uint32_t a_hi = smem_u32(DG_hi), a_lo = smem_u32(DG_lo), x_hi = smem_u32(XH_hi), x_lo = smem_u32(XH_lo);
            const uint64_t d1 = umma_desc(smem_u32(ones), 256, 128);
            for (int step = 0; step < T; step++) {
                mbar_wait(bar_afull, (uint32_t)(step & 1));
                mbar_wait(bar_dh, (uint32_t)(step & 1));                     // dh first: it is on the serial chain and the pipe is shared
                tc_fence_after();
                if (lane == 0) GBW_STAMP(3, step, 0);
                const uint32_t par = (uint32_t)(step & 1);
                if (elect_one_sync()) {
#pragma unroll
                    for (int ks = 0; ks < R / 8; ks++) {
                        const uint32_t o = (uint32_t)ks * 1024u;
                        const uint64_t dah = umma_desc_mn32(a_hi + o, R * 128), dal = umma_desc_mn32(a_lo + o, R * 128);
                        const uint64_t dbh = umma_desc_mn32(x_hi + o, R * 128), dbl = umma_desc_mn32(x_lo + o, R * 128);
                        const uint32_t first = (step == 0 && ks == 0) ? 0u : 1u, firstp = (step < 2 && ks == 0) ? 0u : 1u;
                        umma_tf32(acc_w + par * NW, dah, dbh, id_w, firstp);
                        umma_tf32(acc_wlo, dal, dbh, id_w, first);
                        umma_tf32(acc_wlo, dah, dbl, id_w, 1u);
                        if (!oxh) {
                            umma_tf32(acc_b + par * 16, dah, d1, id_b, firstp);
                            umma_tf32(acc_blo, dal, d1, id_b, first);
                        }
                    }
                    umma_commit(bar_w);
                }
                __syncwarp();
                if (lane == 0) GBW_STAMP(3, step, 1);
            }
        }
    } else {
        // ===================== x loaders (warps 0-3, 14, 15); warps 0-3 also drain dX and run the epilogue =====================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GbwRegs<H>::OTHER));
        const int ltid = warp < 4 ? tid : 128 + (tid - 14 * 32);
        const uint32_t xh_hi = smem_u32(XH_hi), xh_lo = smem_u32(XH_lo);
        float4 pre[KQM];
        auto load_x = [&](int step) {
            const int t = dir ? step : (T - 1 - step);
#pragma unroll
            for (int j = 0; j < KQM; j++) {
                const int i = ltid + j * GBW_NXL;
                const int r = i / KQ, kq = i - r * KQ;
                pre[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < R * KQ && s0 + r < a.S)
                    pre[j] = __ldg(reinterpret_cast<const float4*>(a.X + (size_t)(s0 + r) * a.x_ss + (size_t)t * a.x_st + kq * 4));
            }
        };
        auto stage_x = [&]() {
#pragma unroll
            for (int j = 0; j < KQM; j++) {
                const int i = ltid + j * GBW_NXL;
                const int r = i / KQ, kq = i - r * KQ;
                if (i >= R * KQ) continue;
                const float v[4] = {pre[j].x, pre[j].y, pre[j].z, pre[j].w};
                float4 hi, lo;
                split_tf32_exact(v[0], hi.x, lo.x); split_tf32_exact(v[1], hi.y, lo.y); split_tf32_exact(v[2], hi.z, lo.z); split_tf32_exact(v[3], hi.w, lo.w);
                const uint32_t off = sw32_unit<R>(r, (kq * 4) & ~7) + (uint32_t)(kq & 1) * 16;
                sts128(xh_hi + off, hi);
                sts128(xh_lo + off, lo);
            }
        };
        auto prefetch_x = [&](int step_) {
            const int t = dir ? step_ : (T - 1 - step_);
#pragma unroll
            for (int j = 0; j < KQM; j++) {
                const int i = ltid + j * GBW_NXL;
                const int r = i / KQ, kq = i - r * KQ;
                if (i < R * KQ && s0 + r < a.S && (kq & 7) == 0)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(a.X + (size_t)(s0 + r) * a.x_ss + (size_t)t * a.x_st + kq * 4));
            }
        };
        // x_{step+1} is staged BEFORE dX_step is drained: both wait for MMAs of `step` that retire at about the same time, and
        // only the staging is on the path to the next a_full (the drain has until the dX MMAs of step + 1)
        const int lrole = warp == 0 ? 4 : (warp == 14 ? 5 : -1);
        load_x(0);
        if (H == 32 && T > 1) prefetch_x(1);
        stage_x();
        fence_async_smem();
        mbar_arrive(bar_afull);
        for (int step = 0; step < T; step++) {
            const int t = dir ? step : (T - 1 - step);
            if (H == 32 && step + 2 < T) prefetch_x(step + 2);
            if (lrole >= 0 && lane == 0) GBW_STAMP(lrole, step, 0);
            if (step + 1 < T) load_x(step + 1);
            if (step + 1 < T) {
                mbar_wait(bar_w, (uint32_t)(step & 1));                      // the weight-gradient MMAs of this step read XH
                if (lrole >= 0 && lane == 0) GBW_STAMP(lrole, step, 1);
                stage_x();
                fence_async_smem();
                mbar_arrive(bar_afull);
                if (lrole >= 0 && lane == 0) GBW_STAMP(lrole, step, 2);
            }
            if (warp < 4) {
                // dX_t rows from TMEM -> HBM (vector reductions: both directions add into the same rows)
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
                            float4 o = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                            if (mr) {
                                const float4 m = __ldg(reinterpret_cast<const float4*>(mr + c0 + q * 4));
                                o.x = m.x > 0.f ? o.x : 0.f; o.y = m.y > 0.f ? o.y : 0.f; o.z = m.z > 0.f ? o.z : 0.f; o.w = m.w > 0.f ? o.w : 0.f;
                            }
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(xr + c0 + q * 4), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(bar_dxr);
                if (lrole >= 0 && lane == 0) GBW_STAMP(lrole, step, 3);
            }
        }
        if (warp < 4) {
            // ---- epilogue: thread = gate column n of dG; acc_w row n = [dW_ih row | dW_hh row | bias]
            mbar_wait(bar_w, (uint32_t)((T - 1) & 1));
            tc_fence_after();
            const int n = warp * 32 + lane;
            const bool valid = n < 4 * H;
            const int row_hh = n < 3 * H ? n : -1;
            const int row_ih = n < 2 * H ? n : (n >= 3 * H ? n - H : -1);
            const uint32_t trow = ((uint32_t)(warp * 32) << 16);
            const bool vih = ((reinterpret_cast<uintptr_t>(a.dWih[dir]) & 15) == 0), vhh = ((reinterpret_cast<uintptr_t>(a.dWhh[dir]) & 15) == 0);
            for (int c0 = 0; c0 < I + H + 8; c0 += 8) {
                float v[8], v1[8];
                const bool bias = !oxh && c0 >= I + H;
                tmem_ld8((bias ? acc_blo : acc_wlo + (uint32_t)c0) + trow, v);
                tmem_ld8((bias ? acc_b : acc_w + (uint32_t)c0) + trow, v1);
#pragma unroll
                for (int q = 0; q < 8; q++) v[q] += v1[q];
                if (T > 1) {                                                 // the odd-step accumulator exists from the second step on
                    tmem_ld8((bias ? acc_b + 16 : acc_w + NW + (uint32_t)c0) + trow, v1);
#pragma unroll
                    for (int q = 0; q < 8; q++) v[q] += v1[q];
                }
                if (!valid) continue;
                if (c0 < I) {
                    if (row_ih >= 0) {
                        float* o = a.dWih[dir] + (size_t)row_ih * I + c0;
#pragma unroll
                        for (int q = 0; q < 2; q++) {
                            if (vih)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4 * q), "f"(v[4 * q]), "f"(v[4 * q + 1]),
                                             "f"(v[4 * q + 2]), "f"(v[4 * q + 3]) : "memory");
                            else { atomicAdd(o + 4 * q, v[4 * q]); atomicAdd(o + 4 * q + 1, v[4 * q + 1]); atomicAdd(o + 4 * q + 2, v[4 * q + 2]); atomicAdd(o + 4 * q + 3, v[4 * q + 3]); }
                        }
                    }
                } else if (c0 < I + H) {
                    if (row_hh >= 0) {
                        float* o = a.dWhh[dir] + (size_t)row_hh * H + (c0 - I);
#pragma unroll
                        for (int q = 0; q < 2; q++) {
                            if (vhh)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4 * q), "f"(v[4 * q]), "f"(v[4 * q + 1]),
                                             "f"(v[4 * q + 2]), "f"(v[4 * q + 3]) : "memory");
                            else { atomicAdd(o + 4 * q, v[4 * q]); atomicAdd(o + 4 * q + 1, v[4 * q + 1]); atomicAdd(o + 4 * q + 2, v[4 * q + 2]); atomicAdd(o + 4 * q + 3, v[4 * q + 3]); }
                        }
                    }
                } else {
                    if (row_hh >= 0) atomicAdd(a.dbhh[dir] + row_hh, v[0]);
                    if (row_ih >= 0) atomicAdd(a.dbih[dir] + row_ih, v[0]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, geo.tmem_cols);
}

static inline int gru_bwdw_rows(int H) { return H == 32 ? 96 : 128; }

static bool gru_bwdw_geom(int H, int I, GruBwdwGeom& g, size_t& smem) {
    const int R = gru_bwdw_rows(H);
    g.whh_lbo = H * 16 + 16;
    g.wih_lbo = I * 16 + 16;
    g.whh_bytes = (uint32_t)(3 * H / 4) * g.whh_lbo;
    g.wih_bytes = (uint32_t)(3 * H / 4) * g.wih_lbo;
    g.dg_bytes = (uint32_t)(4 * H / 32) * R * 128;
    g.xh_bytes = (uint32_t)((I + H + 31) / 32) * R * 128;
    g.st_bytes = (uint32_t)R * H * 4;
    g.tmem_cols = tmem_cols_for(H + I + 3 * (I + H + 16) + 48);
    smem = 1024 + 2 * (size_t)g.dg_bytes + 2 * (size_t)g.xh_bytes + 2 * (size_t)g.st_bytes + 2 * (size_t)g.whh_bytes + 2 * (size_t)g.wih_bytes + 512 + 6 * 8 + 16 + 128 * 4 + 64;
    return smem <= 227 * 1024;
}

static bool g_gru_bwdw = getenv("DOF_GRU_BWDW_OFF") == nullptr;       // env switch for A/B measurements

static bool gru_bwdw_tc_eligible(int H, int I) {
    GruBwdwGeom g; size_t smem;
    return g_gru_bwdw && gru_bwd_tc_eligible(H, I) && ((I + H) % 16 == 0) && (I % 16 == 0) && gru_bwdw_geom(H, I, g, smem);
}

template <int H, int R, int KQM>
static int gru_bwdw_launch_t(const GruBwdwMaps& maps, const GruBwdwArgs& a, const GruBwdwGeom& geo, size_t smem, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        DOF_CUDA(cudaFuncSetAttribute(gru_bwdw_tc_kernel<H, R, KQM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr = true;
    }
    dim3 grid(cdiv(a.S, R), 2);
    gru_bwdw_tc_kernel<H, R, KQM><<<grid, GBW_THREADS, smem, st>>>(maps, a, geo);
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

static int launch_gru_bwdw_tc(const GruBwdwArgs& a, cudaStream_t st) {
    GruBwdwGeom geo;
    size_t smem = 0;
    if (!gru_bwdw_geom(a.H, a.I, geo, smem)) DOF_FAIL(DOF_ERR_UNSUPPORTED, "fused GRU backward + weight-gradient tile does not fit (H=%d I=%d)", a.H, a.I);
    if ((a.x_ss & 3) || (a.x_st & 3) || !aligned16(a.X) || !a.dX) DOF_FAIL(DOF_ERR_ARG, "fused GRU backward: input must be 16-byte aligned, dX required");
    const double rows = (double)a.S * a.T * 2;
    ProfScope ps(a.H == 32 ? "gru_bwdw_tc_h32" : "gru_bwdw_tc_h16", st, rows * 2.0 * 3 * a.H * 2.0 * (a.H + a.I),
                 rows * 4.0 * a.H * (4 + 1 + (a.dOut ? 1 : 0)) + rows * 4.0 * a.I * (a.dXmask ? 3 : 2));
    const int KQ = a.I / 4, R = gru_bwdw_rows(a.H);
    const int kqm = cdiv(R * KQ, GBW_NXL);
    if (!aligned16(a.Hout) || (a.dOut && !aligned16(a.dOut))) DOF_FAIL(DOF_ERR_ARG, "fused GRU backward: Hout / dOut must be 16-byte aligned");
    GruBwdwMaps maps;
    DOF_TRY(tmap_seq3d(a.Hout, a.S, a.T, a.H, R, &maps.Hout));
    if (a.dOut) DOF_TRY(tmap_seq3d(a.dOut, a.S, a.T, a.H, R, &maps.dOut)); else maps.dOut = maps.Hout;
    if (a.H == 32) {
        if (kqm <= 3) return gru_bwdw_launch_t<32, 96, 3>(maps, a, geo, smem, st);
        if (kqm <= 5) return gru_bwdw_launch_t<32, 96, 5>(maps, a, geo, smem, st);
        if (kqm <= 7) return gru_bwdw_launch_t<32, 96, 7>(maps, a, geo, smem, st);
        return gru_bwdw_launch_t<32, 96, 10>(maps, a, geo, smem, st);
    }
    if (kqm <= 3) return gru_bwdw_launch_t<16, 128, 3>(maps, a, geo, smem, st);
    if (kqm <= 6) return gru_bwdw_launch_t<16, 128, 6>(maps, a, geo, smem, st);
    if (kqm <= 8) return gru_bwdw_launch_t<16, 128, 8>(maps, a, geo, smem, st);
    return gru_bwdw_launch_t<16, 128, 11>(maps, a, geo, smem, st);
}
