// Tall-skinny fp32 GEMMs for the pose-window trainer.
//
// Every dense contraction on the hot path has a huge row count (sequences x time
// steps, up to 2.9 M rows per batch) and tiny N, K (<= 96 / <= 448).  Two kernels:
//   gemm_rows  : C[M,N]  = epi( A[M,K] . W^T + bias )           (row-parallel)
//   gemm_wgrad : dW[N,K] += P[M,N]^T . Q[M,K],  db[N] += sum_m P (row-reduction)
// The A / P / Q operands are read through a small "view" (plain / split columns /
// k=5 'same' im2col / time shift) so GRU, Linear and Conv1d forward and backward all
// go through the same two kernels without materialising im2col or shifted copies.
#pragma once
#include "common.cuh"

enum { A_PLAIN = 0, A_SPLIT = 1, A_CONV5 = 2, A_TSHIFT = 3, A_TAPS = 4 };

struct MatView {
    const float* p;
    int ld;
    int mode;
    int split, skip;  // A_SPLIT: col >= split reads col + skip
    int T;            // A_CONV5 / A_TSHIFT: rows are (seq, t) with t = row % T
    int sgn;          // A_CONV5: col = c*5+kk reads row (seq, t + sgn*(kk-2)), channel c
    int shift;        // A_TSHIFT: reads row (seq, t + shift); zero outside [0,T)
    // A_TAPS (dilated causal Conv1d as a GEMM, TCN family): column j*cc + c reads channel c of row
    // (seq, t + dil*(taps-1-j)), zero outside [0,T).  dil = -dilation: the convolution; dil = +dilation: its input gradient
    int taps, cc, dil;
    int off;          // A_TAPS: added to every tap's row shift (a "same"-padded Conv1d(k=5): dil = -1, off = +2)
};

static inline MatView mv_plain(const float* p, int ld) {
    MatView v; memset(&v, 0, sizeof(v)); v.p = p; v.ld = ld; v.mode = A_PLAIN; return v;
}
static inline MatView mv_split(const float* p, int ld, int split, int skip) {
    MatView v = mv_plain(p, ld); v.mode = A_SPLIT; v.split = split; v.skip = skip; return v;
}
static inline MatView mv_conv5(const float* p, int ld, int T, int sgn) {
    MatView v = mv_plain(p, ld); v.mode = A_CONV5; v.T = T; v.sgn = sgn; return v;
}
static inline MatView mv_tshift(const float* p, int ld, int T, int shift) {
    MatView v = mv_plain(p, ld); v.mode = A_TSHIFT; v.T = T; v.shift = shift; return v;
}
static inline MatView mv_taps(const float* p, int ld, int T, int taps, int cc, int dil, int off = 0) {
    MatView v = mv_plain(p, ld); v.mode = A_TAPS; v.T = T; v.taps = taps; v.cc = cc; v.dil = dil; v.off = off; return v;
}

__device__ __forceinline__ float mv_load(const MatView& a, int m, int k) {
    switch (a.mode) {
        case A_PLAIN:
            return __ldg(a.p + (size_t)m * a.ld + k);
        case A_SPLIT: {
            int c = (k >= a.split) ? k + a.skip : k;
            return __ldg(a.p + (size_t)m * a.ld + c);
        }
        case A_CONV5: {
            int c = k / 5, kk = k - c * 5;
            int t = m % a.T;
            int tt = t + a.sgn * (kk - 2);
            if (tt < 0 || tt >= a.T) return 0.f;
            return __ldg(a.p + (size_t)(m - t + tt) * a.ld + c);
        }
        case A_TAPS: {
            int j = k / a.cc, c = k - j * a.cc;
            int t = m % a.T;
            int sh = a.dil * (a.taps - 1 - j) + a.off;
            if (t + sh < 0 || t + sh >= a.T) return 0.f;
            return __ldg(a.p + (size_t)(m + sh) * a.ld + c);
        }
        default: {  // A_TSHIFT
            int t = m % a.T;
            int tt = t + a.shift;
            if (tt < 0 || tt >= a.T) return 0.f;
            return __ldg(a.p + (size_t)(m + a.shift) * a.ld + k);
        }
    }
}

struct GemmArgs {
    MatView A;
    const float* W; int ldw; int wT;   // wT=0: W[n*ldw+k] (torch Linear weight); wT=1: W[k*ldw+n]
    const float* bias;                 // [N] or null
    float* C; int ldc;
    int M, N, K;
    int relu;                          // relu(acc + bias (+C))
    int accum;                         // C += ...
    const float* mask; int ldmask;     // if set: out = mask[m,n] > 0 ? out : 0  (ReLU backward)
    // optional second K block: C = A.W + A2.W2 (same N, K) in one pass, e.g. the input gradient of a
    // bidirectional GRU  dX = dG_fwd.W_ih_fwd + dG_bwd.W_ih_bwd  without a read-modify-write of C
    int nkb;                           // 0/1: single block, 2: A2/W2 are valid
    MatView A2; const float* W2;
    int ksplit;                        // tensor-core kernel only: K of every matrix is staged in `ksplit` column
                                       // blocks (smaller smem stages -> double buffering for K > 64); 0/1 = off
    // W is a torch Conv1d weight [C_out, wcin, wtaps] (ldw = wcin * wtaps) addressed for an A_TAPS operand:
    //   wconv = 1 (forward):        n = c_out, k = j * wcin + c_in
    //   wconv = 2 (input gradient): n = c_in,  k = j * C_out + c_out   (C_out = K / wtaps)
    int wconv, wcin, wtaps;
    int wcpad;                         // wconv = 1: channels per tap of the A operand when its pitch is padded (> wcin; 0 = wcin)
};

__device__ __forceinline__ float gemm_w_at(const GemmArgs& g, const float* Wp, int n, int k) {
    if (g.wconv == 0) return g.wT == 0 ? __ldg(Wp + (size_t)n * g.ldw + k) : __ldg(Wp + (size_t)k * g.ldw + n);
    if (g.wconv == 1) {
        const int cp = g.wcpad > 0 ? g.wcpad : g.wcin;
        const int j = k / cp, ci = k - j * cp;
        return ci < g.wcin ? __ldg(Wp + (size_t)n * g.ldw + ci * g.wtaps + j) : 0.f;
    }
    const int co_n = g.K / g.wtaps;
    const int j = k / co_n, co = k - j * co_n;
    return __ldg(Wp + (size_t)co * g.ldw + n * g.wtaps + j);
}
struct GemmBatch { GemmArgs g[2]; };

#define GEMM_BM 128
#define GEMM_BK 32
#define GEMM_TM 8
#define GEMM_TN 4

template <int BN>
__global__ void __launch_bounds__(16 * (BN / GEMM_TN))
gemm_rows_kernel(const GemmBatch gb) {
    const GemmArgs& g = gb.g[blockIdx.z];
    constexpr int NT = 16 * (BN / GEMM_TN);
    constexpr int TXN = BN / GEMM_TN;
    __shared__ __align__(16) float As[GEMM_BK][GEMM_BM + 4];
    __shared__ __align__(16) float Ws[GEMM_BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid % TXN, ty = tid / TXN;
    const int m0 = blockIdx.x * GEMM_BM;
    const int n0 = blockIdx.y * BN;
    if (m0 >= g.M || n0 >= g.N) return;
    float acc[GEMM_TM][GEMM_TN];
#pragma unroll
    for (int i = 0; i < GEMM_TM; i++)
#pragma unroll
        for (int j = 0; j < GEMM_TN; j++) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < g.K; k0 += GEMM_BK) {
        for (int i = tid; i < GEMM_BM * GEMM_BK; i += NT) {
            int k = i % GEMM_BK, m = i / GEMM_BK;
            float v = 0.f;
            if (m0 + m < g.M && k0 + k < g.K) v = mv_load(g.A, m0 + m, k0 + k);
            As[k][m] = v;
        }
        if (g.wT == 0) {
            for (int i = tid; i < BN * GEMM_BK; i += NT) {
                int k = i % GEMM_BK, n = i / GEMM_BK;
                float v = 0.f;
                if (n0 + n < g.N && k0 + k < g.K) v = gemm_w_at(g, g.W, n0 + n, k0 + k);
                Ws[k][n] = v;
            }
        } else {
            for (int i = tid; i < BN * GEMM_BK; i += NT) {
                int n = i % BN, k = i / BN;
                float v = 0.f;
                if (n0 + n < g.N && k0 + k < g.K) v = gemm_w_at(g, g.W, n0 + n, k0 + k);
                Ws[k][n] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GEMM_BK; k++) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * GEMM_TM]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * GEMM_TM + 4]);
            float4 b = *reinterpret_cast<const float4*>(&Ws[k][tx * GEMM_TN]);
            float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < GEMM_TM; i++)
#pragma unroll
                for (int j = 0; j < GEMM_TN; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    const int nb = n0 + tx * GEMM_TN;
    const bool vec = ((g.ldc & 3) == 0) && (nb + 3 < g.N) && ((((uintptr_t)g.C) & 15) == 0) &&
                     (g.mask == nullptr || (((g.ldmask & 3) == 0) && ((((uintptr_t)g.mask) & 15) == 0)));
    float bv[4];
#pragma unroll
    for (int j = 0; j < 4; j++) bv[j] = (g.bias && nb + j < g.N) ? __ldg(g.bias + nb + j) : 0.f;
#pragma unroll
    for (int i = 0; i < GEMM_TM; i++) {
        int m = m0 + ty * GEMM_TM + i;
        if (m >= g.M) continue;
        float* cp = g.C + (size_t)m * g.ldc + nb;
        if (vec) {
            float4 o = make_float4(acc[i][0] + bv[0], acc[i][1] + bv[1], acc[i][2] + bv[2], acc[i][3] + bv[3]);
            if (g.accum) {
                float4 c = *reinterpret_cast<const float4*>(cp);
                o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w;
            }
            if (g.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            if (g.mask) {
                float4 mk = *reinterpret_cast<const float4*>(g.mask + (size_t)m * g.ldmask + nb);
                o.x = mk.x > 0.f ? o.x : 0.f; o.y = mk.y > 0.f ? o.y : 0.f;
                o.z = mk.z > 0.f ? o.z : 0.f; o.w = mk.w > 0.f ? o.w : 0.f;
            }
            *reinterpret_cast<float4*>(cp) = o;
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (nb + j >= g.N) continue;
                float o = acc[i][j] + bv[j];
                if (g.accum) o += cp[j];
                if (g.relu) o = fmaxf(o, 0.f);
                if (g.mask) o = g.mask[(size_t)m * g.ldmask + nb + j] > 0.f ? o : 0.f;
                cp[j] = o;
            }
        }
    }
}

static int launch_gemm_rows_simt(const GemmArgs* gs, int nbatch, cudaStream_t st) {
    GemmBatch gb;
    memset(&gb, 0, sizeof(gb));
    int M = 0, N = 0;
    for (int i = 0; i < nbatch; i++) {
        gb.g[i] = gs[i];
        if (gs[i].M > M) M = gs[i].M;
        if (gs[i].N > N) N = gs[i].N;
    }
    if (M <= 0 || N <= 0) return DOF_OK;
    int BN = N <= 16 ? 16 : N <= 32 ? 32 : N <= 48 ? 48 : N <= 64 ? 64 : 96;
    dim3 grid(cdiv(M, GEMM_BM), cdiv(N, BN), nbatch);
    double fl = 0.0;
    for (int i = 0; i < nbatch; i++) fl += 2.0 * gs[i].M * gs[i].N * gs[i].K;
    ProfScope ps("gemm_rows", st, fl);
    switch (BN) {
        case 16: gemm_rows_kernel<16><<<grid, 16 * 4, 0, st>>>(gb); break;
        case 32: gemm_rows_kernel<32><<<grid, 16 * 8, 0, st>>>(gb); break;
        case 48: gemm_rows_kernel<48><<<grid, 16 * 12, 0, st>>>(gb); break;
        case 64: gemm_rows_kernel<64><<<grid, 16 * 16, 0, st>>>(gb); break;
        default: gemm_rows_kernel<96><<<grid, 16 * 24, 0, st>>>(gb); break;
    }
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

static inline GemmArgs gemm_args(MatView A, const float* W, int ldw, int wT, const float* bias, float* C,
                                 int ldc, int M, int N, int K) {
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.A = A; g.W = W; g.ldw = ldw; g.wT = wT; g.bias = bias; g.C = C; g.ldc = ldc;
    g.M = M; g.N = N; g.K = K;
    return g;
}

// ---------------------------------------------------------------------------
// weight gradient: out[n,k] += sum_m P(m,n) * Q(m,k);  db[n] += sum_m P(m,n)
// ---------------------------------------------------------------------------
struct WGradArgs {
    MatView P, Q;
    float* dW; int ldo; int oT;   // oT=0: dW[n*ldo+k]; oT=1: dW[k*ldo+n]
    float* db;                    // [N] or null
    int M, N, K;
    int ks;                       // oT=0: column stride of the output, dW[n*ldo + k*ks] (0 = 1)
    int otaps, ocv;               // Q is an A_TAPS view (K = otaps * cc columns, column = tap * cc + channel) and dW a Conv1d weight
                                  // [N, ocv, otaps]: dW[n*ldo + channel*otaps + tap] for channel < ocv (0: plain output)
    int nv, kv;                   // only rows n < nv / columns k < kv of dW (and db) are written (0 = all): operands whose row
                                  // pitch was padded to a multiple of 4 floats with zero columns (N * F = 42 -> 44)
};
#define WG_MAXBATCH 5
struct WGradBatch { WGradArgs g[WG_MAXBATCH]; };

#define WG_RB 32
__global__ void __launch_bounds__(256) gemm_wgrad_kernel(const WGradBatch wb, int ktiles) {
    const WGradArgs& g = wb.g[blockIdx.z];
    __shared__ __align__(16) float Ps[WG_RB][64 + 4];
    __shared__ __align__(16) float Qs[WG_RB][64 + 4];
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;   // tx -> k, ty -> n
    const int nt = blockIdx.y / ktiles, kt = blockIdx.y % ktiles;
    const int n0 = nt * 64, k0 = kt * 64;
    if (n0 >= g.N || k0 >= g.K) return;
    int rows_per = (g.M + gridDim.x - 1) / gridDim.x;
    rows_per = (rows_per + WG_RB - 1) / WG_RB * WG_RB;
    const int mbeg = blockIdx.x * rows_per;
    const int mend = min(g.M, mbeg + rows_per);
    if (mbeg >= mend) return;
    float acc[4][4];
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    const bool do_bias = (g.db != nullptr) && kt == 0 && tx == 0;
    for (int mb = mbeg; mb < mend; mb += WG_RB) {
        for (int i = tid; i < WG_RB * 64; i += 256) {
            int c = i % 64, r = i / 64;
            int m = mb + r;
            float pv = 0.f, qv = 0.f;
            if (m < mend) {
                if (n0 + c < g.N) pv = mv_load(g.P, m, n0 + c);
                if (k0 + c < g.K) qv = mv_load(g.Q, m, k0 + c);
            }
            Ps[r][c] = pv;
            Qs[r][c] = qv;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < WG_RB; r++) {
            float4 p = *reinterpret_cast<const float4*>(&Ps[r][ty * 4]);
            float4 q = *reinterpret_cast<const float4*>(&Qs[r][tx * 4]);
            float pv[4] = {p.x, p.y, p.z, p.w}, qv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(pv[i], qv[j], acc[i][j]);
            }
            if (do_bias) {
#pragma unroll
                for (int i = 0; i < 4; i++) bsum[i] += pv[i];
            }
        }
        __syncthreads();
    }
    const int nvl = g.nv > 0 ? g.nv : g.N, kvl = g.kv > 0 ? g.kv : g.K;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int n = n0 + ty * 4 + i;
        if (n >= nvl) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int k = k0 + tx * 4 + j;
            if (k >= kvl) continue;
            float* o = g.oT ? g.dW + (size_t)k * g.ldo + n : g.dW + (size_t)n * g.ldo + (size_t)k * (g.ks > 0 ? g.ks : 1);
            if (g.otaps > 0) {
                const int cc = g.K / g.otaps, tj = k / cc, ch = k - tj * cc;
                if (ch >= g.ocv) continue;
                o = g.dW + (size_t)n * g.ldo + ch * g.otaps + tj;
            }
            atomicAdd(o, acc[i][j]);
        }
        if (do_bias) atomicAdd(g.db + n, bsum[i]);
    }
}

static int launch_gemm_wgrad_simt(const WGradArgs* gs, int nbatch, cudaStream_t st, int sm_count) {
    WGradBatch wb;
    memset(&wb, 0, sizeof(wb));
    int M = 0, N = 0, K = 0;
    for (int i = 0; i < nbatch; i++) {
        wb.g[i] = gs[i];
        if (gs[i].M > M) M = gs[i].M;
        if (gs[i].N > N) N = gs[i].N;
        if (gs[i].K > K) K = gs[i].K;
    }
    if (M <= 0 || N <= 0 || K <= 0) return DOF_OK;
    int ntiles = cdiv(N, 64), ktiles = cdiv(K, 64);
    int tiles = ntiles * ktiles * nbatch;
    int want = (4 * sm_count + tiles - 1) / tiles;     // ~4 CTAs per SM in total
    int maxsplit = cdiv(M, 4 * WG_RB);                  // at least 128 rows per CTA
    int msplit = want < maxsplit ? want : maxsplit;
    if (msplit < 1) msplit = 1;
    dim3 grid(msplit, ntiles * ktiles, nbatch);
    double fl = 0.0;
    for (int i = 0; i < nbatch; i++) fl += 2.0 * gs[i].M * gs[i].N * gs[i].K;
    ProfScope ps("gemm_wgrad", st, fl);
    gemm_wgrad_kernel<<<grid, 256, 0, st>>>(wb, ktiles);
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

static inline WGradArgs wgrad_args(MatView P, MatView Q, float* dW, int ldo, int oT, float* db, int M, int N,
                                   int K) {
    WGradArgs g;
    memset(&g, 0, sizeof(g));
    g.P = P; g.Q = Q; g.dW = dW; g.ldo = ldo; g.oT = oT; g.db = db; g.M = M; g.N = N; g.K = K;
    return g;
}
