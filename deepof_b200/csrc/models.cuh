// Kernels of the VQ-VAE and contrastive paths that sit around the shared recurrent encoder / decoder.
//
//   vq_fwd_kernel          VectorQuantizerPT.forward  (deepof/clustering/models_new.py:1358-1423):
//                          squared distances to the codebook, argmin, soft counts (1/d)^2 / sum, quantized
//                          latents, sum (q - z)^2 for the (gradient-free) vq loss, Gram matrix for the
//                          (gradient-free) kmeans loss, populated-code flags.  HBM-bound: D*4 B in,
//                          (4 + K*4 + D*4) B out per window; the codebook lives in shared memory.
//   vq_codebook_grad_kernel  d quantized -> d codebook (one-hot matmul backward, models_new.py:1382-1383)
//   vq_finalize_kernel     the 7 logged scalars of step_vqvae_distill (deepof/clustering/training.py:380-388)
//   views_kernel           step_contrastive_distill's inputs (training.py:497-525): recompute_edges,
//                          middle half-window, and the augmented view of _make_augmented_view
//                          (training.py:2128-2402: time shift, joint rotations, segment interpolation, offsets)
//   ntx_*                  normalize + cosine similarity / temperature + cross-entropy against the diagonal
//                          (training.py:532-545, deepof/clustering/losses.py:59-63, 130-141), value and gradient,
//                          without materialising the [B,B] similarity matrix.
#pragma once
#include "common.cuh"
#include "loss.cuh"
#include "loader.cuh"

#define VQ_MAXD 64
#define VQ_ST_SQ 0       // sum (q - z)^2
#define VQ_ST_REC_Q 1    // reconstruction NLL sum, decoder(quantized)
#define VQ_ST_REC_E 2    // reconstruction NLL sum, decoder(encoder output)
#define VQ_ST_GRAM 4     // D*D Gram sums, then K populated flags

struct VqArgs {
    const float* z;          // [B,D] encoder output
    const float* codebook;   // [D,K]
    float* quant;            // [B,D]
    float* soft;             // [B,K]
    int* idx;                // [B]
    double* stats;           // VQ_ST_* (zeroed by the caller)
    int B, D, K, want_gram;
};

// VectorQuantizerPT (models_new.py:1358-1423): one WARP per window, lanes own codes (k = lane + 32 c).  The codebook
// [D][K] sits in shared memory so a lane's reads are conflict-free and the soft-count stores are coalesced; the nearest
// code is a warp-min over the lanes' distances followed by a BALLOT on (distance == minimum) whose first set bit is the
// first minimum in code order (argmin keeps the first one).  Gram sums for the k-means term accumulate in a per-warp
// shared tile without atomics (one owner lane per entry) and are folded once per CTA.
#define VQ_WARPS 8
#define VQ_MAXK 256
__global__ void __launch_bounds__(VQ_WARPS * 32) vq_fwd_kernel(const VqArgs a) {
    extern __shared__ float vsm[];
    const int D = a.D, K = a.K, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* cb = vsm;                        // [D][K]
    float* ee = cb + D * K;                 // [K]
    float* zs = ee + K;                     // [VQ_WARPS][D]
    float* gram = zs + VQ_WARPS * D;        // [VQ_WARPS][D*D] (want_gram)
    for (int i = tid; i < D * K; i += blockDim.x) cb[i] = __ldg(a.codebook + i);
    if (a.want_gram) for (int i = tid; i < VQ_WARPS * D * D; i += blockDim.x) gram[i] = 0.f;
    __syncthreads();
    for (int k = tid; k < K; k += blockDim.x) {
        float s = 0.f;
        for (int d = 0; d < D; d++) s = fmaf(cb[d * K + k], cb[d * K + k], s);
        ee[k] = s;
    }
    __syncthreads();
    float* z = zs + warp * D;
    float* gw = gram + (size_t)warp * D * D;
    double sq = 0.0;
    const int nchunk = (K + 31) >> 5;
    for (int b = blockIdx.x * VQ_WARPS + warp; b < a.B; b += gridDim.x * VQ_WARPS) {
        float zz = 0.f;
        for (int d = lane; d < D; d += 32) { const float v = __ldg(a.z + (size_t)b * D + d); z[d] = v; zz = fmaf(v, v, zz); }
        zz = warp_sum(zz);
        __syncwarp();
        float dist[VQ_MAXK / 32];
        float mn = INFINITY, ssum = 0.f;
#pragma unroll
        for (int c = 0; c < VQ_MAXK / 32; c++) {
            dist[c] = INFINITY;
            const int k = lane + 32 * c;
            if (c < nchunk && k < K) {
                float dot = 0.f;
                for (int d = 0; d < D; d++) dot = fmaf(z[d], cb[d * K + k], dot);
                dist[c] = zz + ee[k] - 2.f * dot;                         // models_new.py:1407-1411
                mn = fminf(mn, dist[c]);
                const float inv = 1.f / dist[c];
                ssum = fmaf(inv, inv, ssum);                              // :1416
            }
        }
        mn = warp_min(mn);
        ssum = warp_sum(ssum);
        int best = -1;
#pragma unroll
        for (int c = 0; c < VQ_MAXK / 32; c++) {
            const unsigned hit = __ballot_sync(0xffffffffu, dist[c] == mn);
            if (best < 0 && hit) best = 32 * c + __ffs(hit) - 1;         // first minimum in code order
            const int k = lane + 32 * c;
            if (c < nchunk && k < K) { const float inv = 1.f / dist[c]; a.soft[(size_t)b * K + k] = inv * inv / ssum; }
        }
        if (lane == 0) {
            a.idx[b] = best;
            a.stats[VQ_ST_GRAM + D * D + best] = 1.0;                     // benign race: all writers store 1
        }
        for (int d = lane; d < D; d += 32) {
            const float q = cb[d * K + best];
            a.quant[(size_t)b * D + d] = q;
            const float df = q - z[d];
            sq += (double)(df * df);
        }
        if (a.want_gram)
            for (int e = lane; e < D * D; e += 32) gw[e] = fmaf(z[e / D], z[e % D], gw[e]);
        __syncwarp();
    }
    sq = warp_sum_d(sq);
    if (lane == 0 && sq != 0.0) atomicAdd(a.stats + VQ_ST_SQ, sq);
    if (a.want_gram) {
        __syncthreads();
        for (int i = tid; i < D * D; i += blockDim.x) {
            float s = 0.f;
            for (int w = 0; w < VQ_WARPS; w++) s += gram[(size_t)w * D * D + i];
            atomicAdd(a.stats + VQ_ST_GRAM + i, (double)s);
        }
    }
}

static inline size_t vq_fwd_smem(int D, int K, bool gram) {
    return ((size_t)D * K + K + (size_t)VQ_WARPS * D + (gram ? (size_t)VQ_WARPS * D * D : 0)) * 4;
}

__global__ void vq_codebook_grad_kernel(const float* __restrict__ dquant, const int* __restrict__ idx, float* __restrict__ gcb,
                                        int B, int D, int K) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * D) return;
    const int b = i / D, d = i - b * D;
    atomicAdd(gcb + (size_t)d * K + idx[b], dquant[i]);
}

#define VQ_ST_DISTILL 3  // sum_b w_b * soft-CE_b of the distillation head (0 without a teacher)
struct VqFinalArgs {
    double* stats; float* logs;
    int B, T, Dx, D, K;
    float beta, kmeans_w, lambda_distill;
};

// logs: 0 total 1 enc_rec 2 reconstruct 3 vq 4 kmeans 5 populated 6 distill (training.py:380-388)
__global__ void __launch_bounds__(32) vq_finalize_kernel(const VqFinalArgs a) {
    extern __shared__ __align__(16) double fsm[];
    const int D = a.D, lane = threadIdx.x;
    double* A = fsm;
    double* V = A + D * D;
    double km = 0.0;
    if (a.kmeans_w != 0.f) {
        for (int i = lane; i < D * D; i += 32) A[i] = a.stats[VQ_ST_GRAM + i] / (double)a.B;
        __syncwarp();
        jacobi_eig_warp(A, V, D);
        double s = 0.0;
        for (int i = lane; i < D; i += 32) { double ev = fabs(A[i * D + i]); s += sqrt(ev > 1e-9 ? ev : 1e-9); }
        s = warp_sum_d(s);
        km = (double)a.kmeans_w * s / D;                                   // losses.py:257-287
    }
    double pop = 0.0;
    for (int k = lane; k < a.K; k += 32) pop += a.stats[VQ_ST_GRAM + D * D + k];
    pop = warp_sum_d(pop);
    if (lane == 0) {
        const double cst = 0.5 * (double)a.Dx * 1.8378770664093453;        // Dx/2 * log(2 pi)
        const double bt = (double)a.B * a.T;
        const double encrec = a.stats[VQ_ST_REC_Q] / bt + cst, rec = a.stats[VQ_ST_REC_E] / bt + cst;
        const double vq = (1.0 + (double)a.beta) * a.stats[VQ_ST_SQ] / ((double)a.B * D);   // models_new.py:1387-1391
        const double dist = (double)a.lambda_distill * a.stats[VQ_ST_DISTILL] / (double)a.B;
        a.logs[0] = (float)(encrec + rec + vq + km + dist);
        a.logs[1] = (float)encrec; a.logs[2] = (float)rec; a.logs[3] = (float)vq; a.logs[4] = (float)km;
        a.logs[5] = (float)pop; a.logs[6] = (float)dist;
    }
}

// ---------------------------------------------------------------------------
// contrastive views
// ---------------------------------------------------------------------------
#define VW_MAXN 32
#define VW_MAXROT 8
struct ViewsArgs {
    const float* x_full;     // [B, Tf, N, 3]
    float* x2;               // [2B, Th, N, 3]: rows 0..B-1 the middle half window, B..2B-1 the augmented view
    float* a2;               // [2B, Th, E, 1]
    const int* start;        // [B] slice start of the augmented view
    const float* rot_theta;  // [n_rot, B] radians
    const int* interp_t0;    // [B] or null
    const int* interp_len;   // [B] (0 = off)
    const float* noise;      // [B, N, 3] or null
    int B, Tf, Th, N, E, n_rot, mid_start;
    int rot_pivot[VW_MAXROT];
    unsigned rot_mask[VW_MAXROT];
    short e0[LD_MAXE], e1[LD_MAXE];
};

// one row (all nodes of one time step) of the rotated view: training.py:2224-2249, rotations applied in order,
// each around the CURRENT position of its pivot
__device__ __forceinline__ void vw_rotated_row(const ViewsArgs& a, int b, int tsrc, float (&px)[VW_MAXN], float (&py)[VW_MAXN],
                                               float (&ps)[VW_MAXN]) {
    const float* r = a.x_full + ((size_t)b * a.Tf + tsrc) * a.N * 3;
    for (int n = 0; n < a.N; n++) { px[n] = r[3 * n]; py[n] = r[3 * n + 1]; ps[n] = r[3 * n + 2]; }
    for (int k = 0; k < a.n_rot; k++) {
        const float th = a.rot_theta[(size_t)k * a.B + b];
        float sn, cs;
        sincosf(th, &sn, &cs);
        const float cx = px[a.rot_pivot[k]], cy = py[a.rot_pivot[k]];
        const unsigned m = a.rot_mask[k];
        for (int n = 0; n < a.N; n++)
            if ((m >> n) & 1u) {
                const float rx = px[n] - cx, ry = py[n] - cy;
                px[n] = (rx * cs - ry * sn) + cx;
                py[n] = (rx * sn + ry * cs) + cy;
            }
    }
}

__device__ __forceinline__ void vw_write_row(const ViewsArgs& a, size_t row, const float (&px)[VW_MAXN], const float (&py)[VW_MAXN],
                                             const float (&ps)[VW_MAXN]) {
    float* xo = a.x2 + row * a.N * 3;
    for (int n = 0; n < a.N; n++) { xo[3 * n] = px[n]; xo[3 * n + 1] = py[n]; xo[3 * n + 2] = ps[n]; }
    float* ao = a.a2 + row * a.E;
    for (int e = 0; e < a.E; e++) {
        const float dx = px[a.e0[e]] - px[a.e1[e]], dy = py[a.e0[e]] - py[a.e1[e]];
        ao[e] = sqrtf(fmaxf(dx * dx + dy * dy, 1e-12f));                   // model_utils_new.py:359-362
    }
}

// one thread per (view, window, time step)
__global__ void __launch_bounds__(128) views_kernel(const ViewsArgs a) {
    const long long total = 2LL * a.B * a.Th;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i % a.Th);
        const int vb = (int)(i / a.Th);
        const int b = vb % a.B, view = vb / a.B;
        float px[VW_MAXN], py[VW_MAXN], ps[VW_MAXN];
        if (view == 0) {
            const float* r = a.x_full + ((size_t)b * a.Tf + a.mid_start + t) * a.N * 3;      // training.py:518-522
            for (int n = 0; n < a.N; n++) { px[n] = r[3 * n]; py[n] = r[3 * n + 1]; ps[n] = r[3 * n + 2]; }
        } else {
            const int st = a.start[b];
            const int L = a.interp_len ? a.interp_len[b] : 0;
            const int t0 = L > 0 ? a.interp_t0[b] : 0;
            if (L > 0 && t >= t0 && t < t0 + L) {                                          // training.py:2330-2357
                float qx[VW_MAXN], qy[VW_MAXN], qs[VW_MAXN];
                vw_rotated_row(a, b, st + t0 - 1, px, py, ps);
                vw_rotated_row(a, b, st + t0 + L, qx, qy, qs);
                float al = ((float)t - ((float)t0 - 1.f)) / ((float)L + 1.f);
                al = fminf(fmaxf(al, 0.f), 1.f);
                for (int n = 0; n < a.N; n++) {
                    px[n] = (1.f - al) * px[n] + al * qx[n];
                    py[n] = (1.f - al) * py[n] + al * qy[n];
                    ps[n] = (1.f - al) * ps[n] + al * qs[n];
                }
            } else {
                vw_rotated_row(a, b, st + t, px, py, ps);
            }
            if (a.noise) {
                const float* nz = a.noise + (size_t)b * a.N * 3;
                for (int n = 0; n < a.N; n++) { px[n] += nz[3 * n]; py[n] += nz[3 * n + 1]; ps[n] += nz[3 * n + 2]; }
            }
        }
        vw_write_row(a, (size_t)vb * a.Th + t, px, py, ps);
    }
}

// ---------------------------------------------------------------------------
// NT-Xent (nce, cosine)
// ---------------------------------------------------------------------------
#define NTX_ST_LOSS 0
#define NTX_ST_POS 1
#define NTX_ST_ALL 2
#define NTX_ST_DISTILL 7
#define NTX_TILE 128
#define NTX_WARPS 8
struct NtxArgs {
    const float* enc;   // [2B, D] encoder outputs: rows 0..B-1 = z, B..2B-1 = z_aug
    float* zn;          // [2B, D] row-normalised
    float* nrm;         // [2B]
    float* lse;         // [4,B] per-row coefficients of the gradient pass: m_i | A_i | P_i | Dg_i with
                        //   d loss_i / d s_ij = e_ij (A_i + P_i e_ij), e_ij = exp(s_ij - m_i), for j != i;  Dg_i for j == i
    int kind;           // 0 nce (losses.py:130-141), 1 dcl debiased (:144-173), 2 hard_dcl debiased (:213-249), 3 fc (:176-210)
    int sim;            // 0 cosine / dot on the normalised rows (losses.py:59-67), 1 euclidean / edit 1/(1+|x-y|) (:70-82)
    float tau_plus, beta, temperature;
    float* denc;        // [2B, D] gradient wrt enc
    double* stats;
    float* logs;
    int B, D;
    float inv_tau;
};

__global__ void ntx_norm_kernel(const NtxArgs a) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= 2 * a.B) return;
    float s = 0.f;
    for (int d = lane; d < a.D; d += 32) { float v = a.enc[(size_t)w * a.D + d]; s += v * v; }
    s = warp_sum(s);
    const float n = fmaxf(sqrtf(s), 1e-12f);                               // F.normalize eps
    for (int d = lane; d < a.D; d += 32) a.zn[(size_t)w * a.D + d] = a.enc[(size_t)w * a.D + d] / n;
    if (lane == 0) a.nrm[w] = n;
}

// PHASE 0: row statistics of sim = zn . an^T / tau (lse_i, s_ii, sum_j s_ij), one warp per row of z.
// PHASE 1: gradients; blocks [0, nb) own rows of z (d loss / d zn_i = sum_j g_ij an_j), blocks [nb, 2 nb) rows of
// z_aug (d loss / d an_j = sum_i g_ij zn_i), g_ij = (softmax_ij - delta_ij) / (B tau); then back through the row
// normalisation.  The opposite matrix streams through shared memory in 128-row tiles; lanes own rows of the tile
// and keep a private D-vector accumulator that is warp-reduced once at the end.
template <int PHASE, int DP>
__global__ void __launch_bounds__(NTX_WARPS * 32) ntx_kernel(const NtxArgs a) {
    extern __shared__ float nsm[];
    const int D = a.D, B = a.B, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* tile = nsm;                               // [NTX_TILE][DP+1]
    float* tl = tile + NTX_TILE * (DP + 1);          // [4][NTX_TILE] gradient coefficients of the tile rows (column blocks)
    float* mine = tl + 4 * NTX_TILE + warp * DP;     // this warp's own row
    const int nb = (B + NTX_WARPS - 1) / NTX_WARPS;
    const bool col = PHASE == 1 && (int)blockIdx.x >= nb;
    const int self = ((int)blockIdx.x % nb) * NTX_WARPS + warp;      // index inside its own matrix
    const bool active = self < B;
    const int w = self + (col ? B : 0);                               // row of enc / zn
    const float* other = a.zn + (col ? 0 : (size_t)B * D);
    if (active) for (int d = lane; d < D; d += 32) mine[d] = a.zn[(size_t)w * D + d];
    __syncwarp();
    float m = -INFINITY, l = 0.f, sall = 0.f, sdiag = 0.f;
    float acc[DP];
#pragma unroll
    for (int d = 0; d < DP; d++) acc[d] = 0.f;
    float my_m = 0.f, my_A = 0.f, my_P = 0.f, my_D = 0.f;
    if (PHASE == 1 && active && !col) { my_m = a.lse[self]; my_A = a.lse[B + self]; my_P = a.lse[2 * B + self]; my_D = a.lse[3 * B + self]; }
    float l2 = 0.f, wsum = 0.f;
    const float gs = a.inv_tau / (float)B;
    for (int j0 = 0; j0 < B; j0 += NTX_TILE) {
        __syncthreads();
        for (int i = threadIdx.x; i < NTX_TILE * D; i += blockDim.x) {
            const int r = i / D, d = i - r * D;
            tile[r * (DP + 1) + d] = (j0 + r < B) ? other[(size_t)(j0 + r) * D + d] : 0.f;
        }
        if (col)
            for (int i = threadIdx.x; i < 4 * NTX_TILE; i += blockDim.x) {
                const int q = i / NTX_TILE, r = i - q * NTX_TILE;
                tl[i] = (j0 + r < B) ? a.lse[q * B + j0 + r] : 0.f;
            }
        __syncthreads();
        if (!active) continue;
        for (int r = lane; r < NTX_TILE && j0 + r < B; r += 32) {
            const float* tr = tile + r * (DP + 1);
            float dot = 0.f, dist = 0.f, sv = 0.f;
            if (a.sim == 0) {
#pragma unroll
                for (int d = 0; d < DP; d++) if (d < D) dot += mine[d] * tr[d];
                sv = dot;
            } else {
#pragma unroll
                for (int d = 0; d < DP; d++) if (d < D) { const float df = mine[d] - tr[d]; dot += df * df; }
                dist = sqrtf(fmaxf(dot, 0.f));
                sv = 1.0f / (1.0f + dist);
            }
            const float s = sv * a.inv_tau;
            if (PHASE == 0) {
                sall += s;
                if (j0 + r == self) sdiag = s;
                const float mn = fmaxf(m, s);
                const float sc = __expf(m - mn), e = __expf(s - mn);
                l = l * sc + e;
                l2 = l2 * sc * sc + e * e;
                m = mn;
            } else {
                const float rm = col ? tl[r] : my_m, rA = col ? tl[NTX_TILE + r] : my_A, rP = col ? tl[2 * NTX_TILE + r] : my_P,
                            rD = col ? tl[3 * NTX_TILE + r] : my_D;
                const float e = __expf(s - rm);
                float g = (j0 + r == self) ? rD : e * (rA + rP * e);
                if (a.kind == 3) g = (j0 + r == self) ? rD : (s < rP ? e * rA : 0.f);     // fc: rP carries the row's cut value
                g *= gs;
                if (a.sim != 0) {
                    // d sim / d mine = -(mine - other) / ((1 + dist)^2 dist): accumulate w*other and sum(w)
                    g = dist > 0.f ? -g * sv * sv / dist : 0.f;
                    wsum += g;
                }
#pragma unroll
                for (int d = 0; d < DP; d++) if (d < D) acc[d] += g * tr[d];
            }
        }
    }
    if (!active) return;
    if (PHASE == 0) {
        const float mall = warp_max(m);
        const float rs = __expf(m - mall);
        l = warp_sum(l * rs);
        l2 = warp_sum(l2 * rs * rs);
        sall = warp_sum(sall);
        sdiag = warp_sum(sdiag);
        if (lane == 0) {
            // per-row loss and gradient coefficients (fp64: exp(s) reaches e^(1/temperature))
            const double E = exp((double)mall), pos = exp((double)sdiag), ed = exp((double)sdiag - (double)mall);
            const double Neff = (double)(B - 1), tp = (double)a.tau_plus;
            double loss, A, P = 0.0, Dg;
            if (a.kind == 0) {
                loss = (double)mall + log((double)l) - (double)sdiag;
                A = 1.0 / (double)l;
                Dg = ed / (double)l - 1.0;
            } else {
                const double S = fmax((double)l - ed, 0.0) * E;                      // sum of the negatives exp(s_ij)
                const double Q = fmax((double)l2 - ed * ed, 0.0) * E * E;            // sum of their squares
                double raw, clip;
                if (a.kind == 1) { raw = (-tp * Neff * pos + S) / (1.0 - tp); clip = Neff * exp(-1.0 / (double)a.temperature); }
                else {
                    const double R = a.beta != 0.f ? (double)a.beta * Neff * Q / fmax(S, 1e-300) : S;
                    raw = (-tp * Neff * pos + R) / (1.0 - tp); clip = exp(-1.0 / (double)a.temperature);
                }
                const double act = raw > clip ? 1.0 : 0.0, Ng = fmax(raw, clip);
                const double den = (1.0 - tp) * (pos + Ng);
                loss = log(pos + Ng) - (double)sdiag;
                Dg = (pos - act * tp * Neff * pos / (1.0 - tp)) / (pos + Ng) - 1.0;
                if (a.kind == 1 || a.beta == 0.f) A = act * E / den;
                else {
                    A = -act * (double)a.beta * Neff * Q * E / (fmax(S, 1e-300) * fmax(S, 1e-300)) / den;
                    P = act * 2.0 * (double)a.beta * Neff * E * E / fmax(S, 1e-300) / den;
                }
            }
            a.lse[self] = mall; a.lse[B + self] = (float)A; a.lse[2 * B + self] = (float)P; a.lse[3 * B + self] = (float)Dg;
            atomicAdd(a.stats + NTX_ST_LOSS, loss);
            atomicAdd(a.stats + NTX_ST_POS, (double)sdiag);
            atomicAdd(a.stats + NTX_ST_ALL, (double)sall);
        }
    } else {
        float dotg = 0.f;
        wsum = warp_sum(wsum);
#pragma unroll
        for (int d = 0; d < DP; d++) {
            if (d < D) {
                acc[d] = warp_sum(acc[d]);
                if (a.sim != 0) acc[d] = wsum * mine[d] - acc[d];
                dotg += acc[d] * mine[d];
            }
        }
        // through zn = e / max(|e|, eps): de = (dzn - (dzn . zn) zn) / |e|
        const float inv = 1.f / a.nrm[w];
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int d = 0; d < DP; d++) if (d < D) a.denc[(size_t)w * D + d] = (acc[d] - dotg * mine[d]) * inv;
        }
    }
}

// ---------------------------------------------------------------------------
// fc_loss_pt (losses.py:176-210): per row the k = ceil(0.1 N) LARGEST negatives are eliminated, the rest enter
// -log(pos / (pos + sum exp(neg))).  One warp per row keeps the row's N similarities in shared memory and finds the
// k-th largest by a 32-step bisection on the order-preserving integer image of the floats (exact, no sort); the
// gradient pass (ntx_kernel<1>, kind 3) recomputes s_ij with the same arithmetic and keeps s_ij < cut_i.
// Rows whose k-th and (k+1)-th largest negatives are bit-identical keep the loss exact (ties are counted) but give
// the tied entries no gradient.
// ---------------------------------------------------------------------------
__host__ __device__ inline int fc_drop_count(int N) {
    int k = (int)ceil(0.1 * (double)N);          // elimination_topk = 0.1 at the call site (training.py:545); Python float == double
    return k == 0 ? 1 : k;
}
__device__ __forceinline__ unsigned fc_key(float f) {
    const unsigned u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

template <int DP>
__global__ void __launch_bounds__(NTX_WARPS * 32) ntx_fc_rows_kernel(const NtxArgs a, int rw, int bpad) {
    extern __shared__ float nsm[];
    const int D = a.D, B = a.B, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* tile = nsm;                               // [NTX_TILE][DP+1]
    float* mine = tile + NTX_TILE * (DP + 1) + warp * DP;
    float* sv = tile + NTX_TILE * (DP + 1) + NTX_WARPS * DP + (size_t)warp * bpad;     // this warp's row of similarities
    const int self = (int)blockIdx.x * rw + warp;
    const bool active = warp < rw && self < B;
    const float* other = a.zn + (size_t)B * D;
    if (active) for (int d = lane; d < D; d += 32) mine[d] = a.zn[(size_t)self * D + d];
    __syncwarp();
    float sdiag = 0.f;
    for (int j0 = 0; j0 < B; j0 += NTX_TILE) {
        __syncthreads();
        for (int i = threadIdx.x; i < NTX_TILE * D; i += blockDim.x) {
            const int r = i / D, d = i - r * D;
            tile[r * (DP + 1) + d] = (j0 + r < B) ? other[(size_t)(j0 + r) * D + d] : 0.f;
        }
        __syncthreads();
        if (!active) continue;
        for (int r = lane; r < NTX_TILE && j0 + r < B; r += 32) {
            const float* tr = tile + r * (DP + 1);
            float dot = 0.f, sim;
            if (a.sim == 0) {
#pragma unroll
                for (int d = 0; d < DP; d++) if (d < D) dot += mine[d] * tr[d];
                sim = dot;
            } else {
#pragma unroll
                for (int d = 0; d < DP; d++) if (d < D) { const float df = mine[d] - tr[d]; dot += df * df; }
                sim = 1.0f / (1.0f + sqrtf(fmaxf(dot, 0.f)));
            }
            const float s = sim * a.inv_tau;
            if (j0 + r == self) { sdiag = s; sv[j0 + r] = -INFINITY; }     // the positive never competes with the negatives
            else sv[j0 + r] = s;
        }
    }
    if (!active) return;
    __syncwarp();
    sdiag = warp_sum(sdiag);
    const int kdrop = fc_drop_count(B), keep = B - 1 - kdrop;
    float cut = -INFINITY;                                                  // kept iff s < cut
    double negsum = 0.0, trim = 0.0;                                        // sum exp(s - m), sum s over the kept negatives
    float m = sdiag;
    if (keep > 0) {
        unsigned K = 0;
        for (int bit = 31; bit >= 0; bit--) {
            const unsigned cand = K | (1u << bit);
            int cnt = 0;
            for (int j = lane; j < B; j += 32) cnt += fc_key(sv[j]) >= cand;
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (cnt >= kdrop) K = cand;
        }
        int ngt = 0, neq = 0;
        float vcut = -INFINITY;
        for (int j = lane; j < B; j += 32) {
            const unsigned k = fc_key(sv[j]);
            ngt += k > K; neq += k == K;
            if (k == K) vcut = sv[j];
        }
        ngt = __reduce_add_sync(0xffffffffu, ngt);
        neq = __reduce_add_sync(0xffffffffu, neq);
        cut = warp_max(vcut);
        m = fmaxf(sdiag, cut);
        float ns = 0.f, ts = 0.f;
        for (int j = lane; j < B; j += 32) {
            const float s = sv[j];
            if (j != self && s < cut) { ns += __expf(s - m); ts += s; }
        }
        negsum = (double)warp_sum(ns);
        trim = (double)warp_sum(ts);
        const int keep_eq = neq - (kdrop - ngt);                            // copies of the cut value that survive
        negsum += (double)keep_eq * exp((double)cut - (double)m);
        trim += (double)keep_eq * (double)cut;
    }
    if (lane == 0) {
        const double pos = exp((double)sdiag - (double)m), den = pos + negsum;
        const double loss = (double)m + log(den) - (double)sdiag;
        a.lse[self] = m; a.lse[B + self] = (float)(1.0 / den); a.lse[2 * B + self] = cut; a.lse[3 * B + self] = (float)(-negsum / den);
        atomicAdd(a.stats + NTX_ST_LOSS, loss);
        atomicAdd(a.stats + NTX_ST_POS, (double)sdiag);
        atomicAdd(a.stats + NTX_ST_ALL, trim);
    }
}

// logs: 0 total 1 pos_similarity 2 neg_similarity 3 distill 4 seperability (training.py:582-588)
__global__ void ntx_finalize_kernel(const NtxArgs a, float tau, float lambda_distill) {
    const double B = (double)a.B;
    const double dist = (double)lambda_distill * a.stats[NTX_ST_DISTILL] / B;
    a.logs[0] = (float)(a.stats[NTX_ST_LOSS] / B + dist);
    a.logs[1] = (float)(a.stats[NTX_ST_POS] / B * tau);
    a.logs[2] = a.B > 1 ? (float)((a.stats[NTX_ST_ALL] - a.stats[NTX_ST_POS]) * tau / (B * (B - 1.0))) : 0.f;
    if (a.kind == 3) {                                                   // fc: mean of the KEPT negatives (losses.py:207)
        const int keep = a.B - 1 - fc_drop_count(a.B);
        a.logs[2] = keep > 0 ? (float)(a.stats[NTX_ST_ALL] * tau / (B * (double)keep)) : 0.f;
    }
    a.logs[3] = (float)dist;
    a.logs[4] = 0.f;
}

// ---------------------------------------------------------------------------
// distillation head of step_vqvae_distill / step_contrastive_distill (deepof/clustering/training.py:341-370, 550-578):
// logits = DiscriminativeHead(z) = z W^T + b (teacher_model.py:795-808); targets = softmax(log(clamp_min(tau, 1e-8)) / T)
// when T > 0; per-sample soft cross-entropy -sum_k clamp(t_k, 1e-8, 1) log_softmax(logits)_k (_soft_ce_logits,
// training.py:392-400), optionally weighted by the teacher confidence; loss = lambda * mean_b.  One thread per window:
// the head gradient is reduced in shared memory first, the encoder-output gradient is ADDED to denc.
// ---------------------------------------------------------------------------
#define DH_MAXK 64
#define DH_MAXD 64
struct DistillArgs {
    const float* z;          // [B, D] what the head sees: encoder output (VQ-VAE) or its row-normalised copy (contrastive)
    const float* nrm;        // [B] row norms when z is the normalised copy (the gradient goes back through F.normalize), else NULL
    const float* head;       // W [K, D] | b [K]
    float* head_grad;        // same layout, zeroed by the caller
    const float* tau;        // [B, K] teacher posteriors of the batch
    float* denc;             // [B, D] += d loss / d z
    double* stat;            // += sum_b w_b * ce_b (the caller scales by lambda / B)
    int B, D, K;
    float lambda, sharpen_T, conf_thresh;
    int conf_weight;
};

__global__ void __launch_bounds__(128) distill_head_kernel(const DistillArgs a) {
    extern __shared__ __align__(16) float dhsm[];
    const int D = a.D, K = a.K, tid = threadIdx.x;
    float* Ws = dhsm;                  // [K*D + K]
    float* Gs = Ws + K * D + K;       // [K*D + K] block-level gradient
    for (int i = tid; i < K * D + K; i += blockDim.x) { Ws[i] = a.head[i]; Gs[i] = 0.f; }
    __syncthreads();
    const int b = blockIdx.x * blockDim.x + tid;
    double mine = 0.0;
    if (b < a.B) {
        float z[DH_MAXD], t[DH_MAXK], lg[DH_MAXK];
        for (int d = 0; d < D; d++) z[d] = a.z[(size_t)b * D + d];
        float mx = -INFINITY;
        for (int k = 0; k < K; k++) {
            float s = Ws[K * D + k];
            for (int d = 0; d < D; d++) s += Ws[k * D + d] * z[d];
            lg[k] = s;
            mx = fmaxf(mx, s);
        }
        float se = 0.f;
        for (int k = 0; k < K; k++) se += expf(lg[k] - mx);
        const float lse = mx + logf(se);
        // teacher targets
        for (int k = 0; k < K; k++) t[k] = a.tau[(size_t)b * K + k];
        if (a.sharpen_T > 0.f) {
            float tm = -INFINITY;
            for (int k = 0; k < K; k++) { t[k] = logf(fmaxf(t[k], 1e-8f)) / a.sharpen_T; tm = fmaxf(tm, t[k]); }
            float ts = 0.f;
            for (int k = 0; k < K; k++) { t[k] = expf(t[k] - tm); ts += t[k]; }
            for (int k = 0; k < K; k++) t[k] /= ts;
        }
        float w = 1.f;
        if (a.conf_weight) {
            float conf = -INFINITY;
            for (int k = 0; k < K; k++) conf = fmaxf(conf, t[k]);
            w = fminf(fmaxf((conf - a.conf_thresh) / fmaxf(1e-6f, 1.0f - a.conf_thresh), 0.f), 1.f);
        }
        float ce = 0.f, tsum = 0.f;
        for (int k = 0; k < K; k++) { t[k] = fminf(fmaxf(t[k], 1e-8f), 1.0f); ce -= t[k] * (lg[k] - lse); tsum += t[k]; }
        mine = (double)(w * ce);
        const float gs = a.lambda * w / (float)a.B;
        float dz[DH_MAXD];
        for (int d = 0; d < D; d++) dz[d] = 0.f;
        for (int k = 0; k < K; k++) {
            const float dl = gs * (expf(lg[k] - lse) * tsum - t[k]);
            atomicAdd(Gs + K * D + k, dl);
            for (int d = 0; d < D; d++) { atomicAdd(Gs + k * D + d, dl * z[d]); dz[d] += dl * Ws[k * D + d]; }
        }
        if (a.nrm) {
            // z = e / |e|:  d loss / d e = (dz - (dz . z) z) / |e|
            float dot = 0.f;
            for (int d = 0; d < D; d++) dot += dz[d] * z[d];
            const float inv = 1.0f / a.nrm[b];
            for (int d = 0; d < D; d++) dz[d] = (dz[d] - dot * z[d]) * inv;
        }
        for (int d = 0; d < D; d++) a.denc[(size_t)b * D + d] += dz[d];
    }
    mine = warp_sum_d(mine);
    if ((tid & 31) == 0 && mine != 0.0) atomicAdd(a.stat, mine);
    __syncthreads();
    for (int i = tid; i < K * D + K; i += blockDim.x) if (Gs[i] != 0.f) atomicAdd(a.head_grad + i, Gs[i]);
}
