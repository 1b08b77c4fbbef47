// VadeLoss (losses.py:567-797) as block-reduction kernels, value AND analytic gradient.
//
//   recon_kernel       : reconstruction NLL sum + dloc                       (HBM-bound)
//   loss_stats_kernel  : per-window terms -> batch statistics (one warp per window)
//   loss_finalize_kernel (1 CTA): batch-level terms (non-empty floor, repel, Gram
//                        eigen-decomposition for the "kmeans" loss, MC-KL clamp), the 13
//                        logged scalars, and the coefficients the gradient pass needs
//   loss_grad_kernel   : per-window gradient wrt z / z_mean / softplus pre-activation
//                        and the q-path gradient of the GMM parameters
#pragma once
#include "common.cuh"
#include "../../include/deepof_b200.h"

#define LOSS_MAXK 32

// double stats layout
#define ST_RECON 0
#define ST_ACT 1
#define ST_KLPRE 2
#define ST_MCKL 3
#define ST_DWSUM 4
#define ST_DCESUM 5
#define ST_TEMPORAL 6
#define ST_NHEAD 8
struct StatsLayout {
    int qsum, qz, gram, gA, gMu, gLv, total;
};
static inline __host__ __device__ StatsLayout stats_layout(int D, int K) {
    StatsLayout s;
    s.qsum = ST_NHEAD; s.qz = s.qsum + K; s.gram = s.qz + K * D; s.gA = s.gram + D * D;
    s.gMu = s.gA + K; s.gLv = s.gMu + K * D; s.total = s.gLv + K * D;
    return s;
}
// float coefficient layout (written by finalize, read by grad)
#define CF_KLSCALE 0     // pretrain: klw/(D*B); main: flag*klw/(S*B)
#define CF_DISTSCALE 1   // lambda/(B*mean_w)
#define CF_ACT 2         // l1/B
#define CF_TEMPORAL 3    // rho/(B-1)
#define CF_NHEAD 8
struct CoefLayout {
    int dq, R, M2, total;
};
static inline __host__ __device__ CoefLayout coef_layout(int D, int K) {
    CoefLayout c;
    c.dq = CF_NHEAD; c.R = c.dq + K; c.M2 = c.R + K * D; c.total = c.M2 + D * D;
    return c;
}

// logs[] order == step_vade's log keys (training.py:292-306)
enum {
    LG_TOTAL = 0, LG_RECON, LG_KL, LG_CAT, LG_KMEANS, LG_ACT, LG_PRIOR, LG_DISTILL, LG_TFCLUST,
    LG_NONEMPTY, LG_TEMPORAL, LG_SCATTER, LG_REPEL, LG_KLWEIGHT, LG_MCKL_RAW, LG_N
};

__global__ void __launch_bounds__(256) recon_kernel(const float* __restrict__ loc, const float* __restrict__ x,
                                                    float* __restrict__ dloc, long long n, float inv_bt,
                                                    double* __restrict__ stats) {
    __shared__ double red[8];
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        float d = loc[i] - x[i];
        s += 0.5 * (double)(d * d);
        if (dloc) dloc[i] = d * inv_bt;
    }
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += red[i];
        atomicAdd(stats, t);          // `stats` points at the slot
    }
}

struct LossArgs {
    dof_vade_loss_cfg cfg;
    const float *z, *zm, *lv, *pre, *q;        // [B,D]x4, [B,K]
    const float *eps;                          // [B,D] reparam noise
    const float *mc_eps;                       // [S,B,D]
    unsigned long long noise_seed;             // eps / mc_eps == null: the Philox streams (seed, DOF_SITE_EPS / DOF_SITE_MC)
    const float *gmm_mu, *gmm_lv, *prior;
    const float *tau;                          // [B,K] tau_star rows of this batch, or null
    const float *class_weight;                 // [K] or null
    const float *floor_c;                      // [K] per-cluster non-empty floor
    double* stats; float* coef;
    float* dzm_kl; float* dlv_kl;              // [B,D] unscaled MC-KL gradients
    float* distw;                              // [B] per-window distillation weight (unnormalised)
    const float* dz_dec;                       // [B,D] gradient of z from the decoder
    float* dzm; float* dpre;                   // outputs [B,D]
    float* g_mu; float* g_lv;                  // gradient buffers of gmm_means / gmm_log_vars
    float* logs;                               // [LG_N]
    int B, D, K, T, Dx;
};

// one warp per window, 4 warps per CTA
#define LS_WARPS 4
__global__ void __launch_bounds__(LS_WARPS * 32) loss_stats_kernel(const LossArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int D = a.D, K = a.K, B = a.B;
    const StatsLayout SL = stats_layout(D, K);
    const bool main_mode = !a.cfg.pretrain_mode;
    const bool need_qz = a.cfg.repel_weight > 0.f;
    const bool need_gram = a.cfg.model_kmeans_weight > 0.f;
    // per-CTA accumulators
    float* acc = smem;                              // [SL.total] mirrors the stats layout
    float* cmu = acc + SL.total;                    // [K*D]
    float* cglv = cmu + K * D;                      // clamped gmm log-vars
    float* ciglv = cglv + K * D;                    // exp(-glv)
    float* clp = ciglv + K * D;                     // [K] log prior
    float* wsc = clp + K;                           // per-warp scratch
    const int per_warp = D * 3 + K + 2 * 32 * (D + 1) + 32 * (K + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* zsh = wsc + warp * per_warp;             // [D] z
    float* sqsh = zsh + D;                          // [D] exp(.5 lvc)
    float* lvcsh = sqsh + D;                        // [D] clamped lv
    float* qnsh = lvcsh + D;                        // [K]
    float* zs = qnsh + K;                           // [32][D+1]
    float* gam = zs + 32 * (D + 1);                 // [32][K+1]
    float* mcn = gam + 32 * (K + 1);                // [32][D+1] Monte-Carlo noise of the window
    for (int i = threadIdx.x; i < SL.total; i += blockDim.x) acc[i] = 0.f;
    for (int i = threadIdx.x; i < K * D; i += blockDim.x) {
        cmu[i] = __ldg(a.gmm_mu + i);
        float g = fminf(fmaxf(__ldg(a.gmm_lv + i), a.cfg.gmm_logvar_clamp_lo), a.cfg.gmm_logvar_clamp_hi);
        cglv[i] = g;
        ciglv[i] = expf(-g);
    }
    for (int i = threadIdx.x; i < K; i += blockDim.x) clp[i] = logf(fmaxf(__ldg(a.prior + i), 1e-8f));
    __syncthreads();
    for (int b = blockIdx.x * LS_WARPS + warp; b < B; b += gridDim.x * LS_WARPS) {
        // ---- normalised q (losses.py:589-592)
        float qr = (lane < K) ? a.q[(size_t)b * K + lane] : 0.f;
        float qc = (lane < K) ? fmaxf(qr, 1e-8f) : 0.f;
        float tot = warp_sum(qc);
        float qn = qc / tot;
        if (lane < K) { qnsh[lane] = qn; atomicAdd(&acc[SL.qsum + lane], qn); }
        float act = 0.f, klp = 0.f;
        for (int d = lane; d < D; d += 32) {
            float zv = a.z[(size_t)b * D + d];
            float lv = a.lv[(size_t)b * D + d];
            float zmv = a.zm[(size_t)b * D + d];
            float lvc = fminf(fmaxf(lv, -4.f), 2.f);
            zsh[d] = zv;
            lvcsh[d] = lvc;
            sqsh[d] = expf(0.5f * lvc);
            act += fabsf(lv);
            klp += 0.5f * (zmv * zmv + expf(lvc) - 1.0f - lvc);
        }
        act = warp_sum(act);
        klp = warp_sum(klp) / D;
        if (lane == 0) { atomicAdd(&acc[ST_ACT], act); atomicAdd(&acc[ST_KLPRE], klp); }
        __syncwarp();
        if (need_qz)
            for (int p = lane; p < K * D; p += 32) atomicAdd(&acc[SL.qz + p], qnsh[p / D] * zsh[p % D]);
        if (need_gram)
            for (int p = lane; p < D * D; p += 32) atomicAdd(&acc[SL.gram + p], zsh[p / D] * zsh[p % D]);
        // ---- temporal cohesion value (losses.py:709-712), adjacent batch rows
        if (main_mode && a.cfg.temporal_cohesion_weight > 0.f && b + 1 < B) {
            float qr2 = (lane < K) ? a.q[(size_t)(b + 1) * K + lane] : 0.f;
            float qc2 = (lane < K) ? fmaxf(qr2, 1e-8f) : 0.f;
            float tot2 = warp_sum(qc2);
            float dv = warp_sum((lane < K) ? fabsf(qc2 / tot2 - qn) : 0.f);
            if (lane == 0) atomicAdd(&acc[ST_TEMPORAL], dv);
        }
        // ---- distillation (losses.py:730-760)
        if (a.tau && a.cfg.lambda_distill > 0.f) {
            float tb = (lane < K) ? a.tau[(size_t)b * K + lane] : 0.f;
            if (a.cfg.distill_sharpen_T > 0.f) {
                float lg = (lane < K) ? logf(fmaxf(tb, 1e-8f)) / a.cfg.distill_sharpen_T : -INFINITY;
                float mx = warp_max(lg);
                float e = (lane < K) ? expf(lg - mx) : 0.f;
                tb = e / warp_sum(e);
            }
            float ce = -warp_sum((lane < K) ? tb * logf(fmaxf(qn, 1e-8f)) : 0.f);
            float wc = 1.f;
            if (a.class_weight) wc = warp_sum((lane < K) ? tb * __ldg(a.class_weight + lane) : 0.f);
            float wconf = 1.f;
            if (a.cfg.distill_conf_weight) {
                float conf = warp_max((lane < K) ? tb : 0.f);
                float thr = a.cfg.distill_conf_thresh;
                wconf = fminf(fmaxf((conf - thr) / fmaxf(1e-6f, 1.0f - thr), 0.f), 1.f);
            }
            if (lane == 0) {
                a.distw[b] = wc * wconf;
                atomicAdd(&acc[ST_DWSUM], wc);
                atomicAdd(&acc[ST_DCESUM], wc * wconf * ce);
            }
        }
        // ---- Monte-Carlo KL against the mixture prior (losses.py:506-545), lane = sample
        if (main_mode) {
            const int S = a.cfg.mc_samples;     // == 32 (checked on the host)
            float* myz = zs + lane * (D + 1);
            float* myg = gam + lane * (K + 1);
            const float* me = a.mc_eps ? a.mc_eps + ((size_t)lane * B + b) * D : nullptr;
            float* mye = mcn + lane * (D + 1);       // this sample's noise row (read twice)
            float logq = 0.f;
            for (int d = 0; d < D; d++) {
                mye[d] = me ? me[d] : philox_normal(a.noise_seed, DOF_SITE_MC, ((unsigned long long)lane * B + b) * D + d);
                float zmv = a.zm[(size_t)b * D + d];
                float zv = zmv + mye[d] * sqsh[d];
                myz[d] = zv;
                float df = zv - zmv;
                logq += LOG_2PI_F + lvcsh[d] + df * df * expf(-lvcsh[d]);
            }
            logq *= -0.5f;
            float mx = -INFINITY;
            for (int c = 0; c < K; c++) {
                float s = 0.f;
                for (int d = 0; d < D; d++) {
                    float df = myz[d] - cmu[c * D + d];
                    s += LOG_2PI_F + cglv[c * D + d] + df * df * ciglv[c * D + d];
                }
                float lp = clp[c] - 0.5f * s;
                myg[c] = lp;
                mx = fmaxf(mx, lp);
            }
            float se = 0.f;
            for (int c = 0; c < K; c++) se += expf(myg[c] - mx);
            float logp = mx + logf(se);
            for (int c = 0; c < K; c++) myg[c] = expf(myg[c] - logp);
            float kls = warp_sum(logq - logp);
            if (lane == 0) atomicAdd(&acc[ST_MCKL], kls);
            for (int d = 0; d < D; d++) {
                float g = 0.f;
                for (int c = 0; c < K; c++) g += myg[c] * (myz[d] - cmu[c * D + d]) * ciglv[c * D + d];
                float gz = warp_sum(g);
                float gl = warp_sum(-0.5f + 0.5f * g * mye[d] * sqsh[d]);
                if (lane == 0) { a.dzm_kl[(size_t)b * D + d] = gz; a.dlv_kl[(size_t)b * D + d] = gl; }
            }
            __syncwarp();
            for (int p = lane; p < K * D; p += 32) {
                int c = p / D, d = p % D;
                float m = cmu[p];
                float s1 = 0.f, s2 = 0.f;
                for (int s = 0; s < S; s++) {
                    float g = gam[s * (K + 1) + c];
                    float df = zs[s * (D + 1) + d] - m;
                    s1 = fmaf(g, df, s1);
                    s2 = fmaf(g * df, df, s2);
                }
                atomicAdd(&acc[SL.gMu + p], s1);
                atomicAdd(&acc[SL.gLv + p], s2);
            }
            if (lane < K) {
                float s0 = 0.f;
                for (int s = 0; s < S; s++) s0 += gam[s * (K + 1) + lane];
                atomicAdd(&acc[SL.gA + lane], s0);
            }
        }
        __syncwarp();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SL.total; i += blockDim.x)
        if (i != ST_RECON && acc[i] != 0.f) atomicAdd(a.stats + i, (double)acc[i]);
}

static inline size_t loss_stats_smem_floats(int D, int K) {
    StatsLayout SL = stats_layout(D, K);
    size_t per_warp = (size_t)D * 3 + K + 2 * 32 * (D + 1) + 32 * (K + 1);
    return (size_t)SL.total + 3 * (size_t)K * D + K + LS_WARPS * per_warp;
}

// ---------------------------------------------------------------------------
// finalize: single CTA of 256 threads.
// ---------------------------------------------------------------------------
__device__ void jacobi_eig_warp(double* A, double* V, int D) {
    // cyclic Jacobi on symmetric A[D][D]; eigenvalues on diag(A), vectors = columns of V.
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < D * D; i += 32) V[i] = (i / D == i % D) ? 1.0 : 0.0;
    __syncwarp();
    for (int sweep = 0; sweep < 40; sweep++) {
        double off = 0.0, dg = 0.0;
        for (int i = lane; i < D * D; i += 32) {
            int r = i / D, c = i % D;
            if (r < c) off += A[i] * A[i];
            if (r == c) dg += A[i] * A[i];
        }
        off = warp_sum_d(off);
        dg = warp_sum_d(dg);
        if (off <= 1e-30 * (dg + 1e-300)) break;
        for (int p = 0; p < D - 1; p++) {
            for (int q = p + 1; q < D; q++) {
                double apq = A[p * D + q];
                if (fabs(apq) < 1e-300) continue;
                double theta = (A[q * D + q] - A[p * D + p]) / (2.0 * apq);
                double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                __syncwarp();
                for (int k = lane; k < D; k += 32) {
                    double akp = A[k * D + p], akq = A[k * D + q];
                    A[k * D + p] = c * akp - s * akq;
                    A[k * D + q] = s * akp + c * akq;
                    double vkp = V[k * D + p], vkq = V[k * D + q];
                    V[k * D + p] = c * vkp - s * vkq;
                    V[k * D + q] = s * vkp + c * vkq;
                }
                __syncwarp();
                for (int k = lane; k < D; k += 32) {
                    double apk = A[p * D + k], aqk = A[q * D + k];
                    A[p * D + k] = c * apk - s * aqk;
                    A[q * D + k] = s * apk + c * aqk;
                }
                __syncwarp();
            }
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(256) loss_finalize_kernel(const LossArgs a) {
    extern __shared__ __align__(16) double dsm[];
    const int D = a.D, K = a.K, B = a.B;
    const StatsLayout SL = stats_layout(D, K);
    const CoefLayout CL = coef_layout(D, K);
    const dof_vade_loss_cfg& cfg = a.cfg;
    double* A = dsm;               // [D*D]
    double* V = A + D * D;         // [D*D]
    double* mean = V + D * D;      // [K*D] soft centroids
    double* Km = mean + K * D;     // [K*K]
    double* ew = Km + K * K;       // [D] eigen weights
    __shared__ double sval[LG_N];
    const int tid = threadIdx.x;
    if (tid < LG_N) sval[tid] = 0.0;
    __syncthreads();
    const bool main_mode = !cfg.pretrain_mode;
    // ---- scalar terms (thread 0)
    if (tid == 0) {
        double recon = a.stats[ST_RECON] / ((double)B * a.T) + 0.5 * a.Dx * 1.8378770664093453;
        sval[LG_RECON] = recon;
        sval[LG_ACT] = cfg.l1_activity_weight * a.stats[ST_ACT] / B;
        a.coef[CF_ACT] = cfg.l1_activity_weight / B;
        double kl;
        if (!main_mode) {
            kl = cfg.kl_weight * a.stats[ST_KLPRE] / B;
            a.coef[CF_KLSCALE] = cfg.kl_weight / ((float)D * B);
        } else {
            double raw = a.stats[ST_MCKL] / ((double)cfg.mc_samples * B);
            sval[LG_MCKL_RAW] = raw;
            kl = cfg.kl_weight * (raw > 0.0 ? raw : 0.0);
            a.coef[CF_KLSCALE] = raw > 0.0 ? cfg.kl_weight / ((float)cfg.mc_samples * B) : 0.f;
        }
        sval[LG_KL] = kl;
        sval[LG_KLWEIGHT] = cfg.kl_weight;
        // non-empty floor + categorical balance -> dq coefficients
        double ne = 0.0, cat = 0.0, qtot = 0.0;
        for (int c = 0; c < K; c++) {
            double qm = a.stats[SL.qsum + c] / B;
            qtot += qm;
            double dq = 0.0;
            if (cfg.nonempty_weight > 0.f) {
                double u = (double)a.floor_c[c] - qm;
                if (u > 0.0) {
                    ne += pow(u, (double)cfg.nonempty_p);
                    dq += -cfg.nonempty_weight * cfg.nonempty_p * pow(u, (double)cfg.nonempty_p - 1.0) / B;
                }
            }
            if (main_mode && cfg.reg_cat_clusters_weight > 0.f) {
                double uni = 1.0 / K;
                cat += uni * (log(uni) - log(qm + 1e-9)) / K;
                dq += -cfg.reg_cat_clusters_weight * uni / (K * (qm + 1e-9)) / B;
            }
            a.coef[CL.dq + c] = (float)dq;
        }
        sval[LG_NONEMPTY] = cfg.nonempty_weight * ne;
        sval[LG_CAT] = cfg.reg_cat_clusters_weight * cat;
        if (main_mode) {
            sval[LG_PRIOR] = log((double)(K > 1 ? K : 1)) * qtot;
            sval[LG_TFCLUST] = 0.0;
            if (cfg.temporal_cohesion_weight > 0.f && B > 1) {
                sval[LG_TEMPORAL] = cfg.temporal_cohesion_weight * a.stats[ST_TEMPORAL] / (B - 1);
                a.coef[CF_TEMPORAL] = cfg.temporal_cohesion_weight / (B - 1);
            } else {
                a.coef[CF_TEMPORAL] = 0.f;
            }
        } else {
            a.coef[CF_TEMPORAL] = 0.f;
        }
        // distillation normalisation
        if (a.tau && cfg.lambda_distill > 0.f) {
            double meanw = a.class_weight ? fmax(a.stats[ST_DWSUM] / B, 1e-8) : 1.0;
            sval[LG_DISTILL] = cfg.lambda_distill * a.stats[ST_DCESUM] / (B * meanw);
            a.coef[CF_DISTSCALE] = (float)(cfg.lambda_distill / (B * meanw));
        } else {
            a.coef[CF_DISTSCALE] = 0.f;
        }
    }
    // ---- repel (losses.py:647-664)
    if (cfg.repel_weight > 0.f) {
        for (int i = tid; i < K * D; i += blockDim.x) {
            double pi = fmax(a.stats[SL.qsum + i / D], 1e-8);
            mean[i] = a.stats[SL.qz + i] / pi;
        }
        __syncthreads();
        const double inv2l2 = 1.0 / fmax(1e-9, 2.0 * (double)cfg.repel_length_scale * cfg.repel_length_scale);
        for (int i = tid; i < K * K; i += blockDim.x) {
            int c = i / K, e = i % K;
            double d2 = 0.0;
            for (int d = 0; d < D; d++) { double df = mean[c * D + d] - mean[e * D + d]; d2 += df * df; }
            Km[i] = (c == e) ? 0.0 : exp(-d2 * inv2l2);
        }
        __syncthreads();
        const double denom = (double)(K * K - K > 1 ? K * K - K : 1);
        if (tid == 0) {
            double s = 0.0;
            for (int i = 0; i < K * K; i++) s += Km[i];
            sval[LG_REPEL] = cfg.repel_weight * s / denom;
        }
        for (int i = tid; i < K * D; i += blockDim.x) {
            int c = i / D, d = i % D;
            double g = 0.0;
            for (int e = 0; e < K; e++) g += Km[c * K + e] * (mean[c * D + d] - mean[e * D + d]);
            g *= -2.0 * (2.0 * inv2l2) * cfg.repel_weight / denom;    // d/dm_c of sum over ordered pairs
            double pi = fmax(a.stats[SL.qsum + c], 1e-8);
            a.coef[CL.R + i] = (float)(g / pi);
        }
    } else {
        for (int i = tid; i < K * D; i += blockDim.x) a.coef[CL.R + i] = 0.f;
    }
    // ---- "kmeans" loss: mean sqrt singular values of Gram z^T z / B (losses.py:257-287)
    // (skipped when its weight is zero, e.g. the main phase: value and gradient are exactly 0 then)
    if (cfg.model_kmeans_weight > 0.f && cfg.kmeans_loss_weight != 0.f) {
        for (int i = tid; i < D * D; i += blockDim.x) A[i] = (double)((float)(a.stats[SL.gram + i]) / (float)B);
        __syncthreads();
        if (tid < 32) {
            jacobi_eig_warp(A, V, D);
            double s = 0.0;
            for (int i = tid; i < D; i += 32) {
                double ev = fabs(A[i * D + i]);
                double cl = ev > 1e-9 ? ev : 1e-9;
                s += sqrt(cl);
                ew[i] = (ev >= 1e-9) ? 0.5 / sqrt(cl) : 0.0;
            }
            s = warp_sum_d(s);
            if (tid == 0) sval[LG_KMEANS] = (double)cfg.kmeans_loss_weight * cfg.model_kmeans_weight * s / D;
        }
        __syncthreads();
        const double sc = (double)cfg.kmeans_loss_weight * cfg.model_kmeans_weight / D * 2.0 / B;
        for (int i = tid; i < D * D; i += blockDim.x) {
            int r = i / D, c = i % D;
            double g = 0.0;
            for (int k = 0; k < D; k++) g += ew[k] * V[r * D + k] * V[c * D + k];
            a.coef[CL.M2 + i] = (float)(g * sc);
        }
    } else {
        for (int i = tid; i < D * D; i += blockDim.x) a.coef[CL.M2 + i] = 0.f;
    }
    __syncthreads();
    // ---- MC-KL gradient of the GMM parameters (needs the clamp flag)
    if (main_mode) {
        const float ksc = a.coef[CF_KLSCALE];
        if (ksc != 0.f && a.g_mu) {
            for (int i = tid; i < K * D; i += blockDim.x) {
                float raw = __ldg(a.gmm_lv + i);
                float glv = fminf(fmaxf(raw, cfg.gmm_logvar_clamp_lo), cfg.gmm_logvar_clamp_hi);
                double ig = exp(-(double)glv);
                // d(-log p)/dmu = -sum g (zs-mu) e^-glv ; d(-log p)/dglv = .5 sum g (1 - (zs-mu)^2 e^-glv)
                double gmu = -a.stats[SL.gMu + i] * ig;
                double glvg = 0.5 * (a.stats[SL.gA + i / D] - a.stats[SL.gLv + i] * ig);
                atomicAdd(a.g_mu + i, (float)(ksc * gmu));
                if (raw >= cfg.gmm_logvar_clamp_lo && raw <= cfg.gmm_logvar_clamp_hi)
                    atomicAdd(a.g_lv + i, (float)(ksc * glvg));
            }
        }
    }
    if (tid == 0) {
        double tot = sval[LG_RECON] + sval[LG_KL] + sval[LG_CAT] + sval[LG_TEMPORAL] + sval[LG_NONEMPTY] +
                     sval[LG_TFCLUST] + sval[LG_PRIOR] + sval[LG_KMEANS] + sval[LG_ACT] + sval[LG_SCATTER] +
                     sval[LG_REPEL] + sval[LG_DISTILL];
        sval[LG_TOTAL] = tot;
        for (int i = 0; i < LG_N; i++) a.logs[i] = (float)sval[i];
    }
}

static inline size_t loss_finalize_smem_bytes(int D, int K) {
    return sizeof(double) * ((size_t)2 * D * D + (size_t)K * D + (size_t)K * K + D);
}

// ---------------------------------------------------------------------------
// per-window gradient, one warp per window
// ---------------------------------------------------------------------------
__device__ __forceinline__ float qnorm_row(const float* q, int b, int K, int lane, float* qr_out, float* tot_out) {
    float qr = (lane < K) ? q[(size_t)b * K + lane] : 0.f;
    float qc = (lane < K) ? fmaxf(qr, 1e-8f) : 0.f;
    float tot = warp_sum(qc);
    if (qr_out) *qr_out = qr;
    if (tot_out) *tot_out = tot;
    return qc / tot;
}

__global__ void __launch_bounds__(LS_WARPS * 32) loss_grad_kernel(const LossArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int D = a.D, K = a.K, B = a.B;
    const CoefLayout CL = coef_layout(D, K);
    const dof_vade_loss_cfg& cfg = a.cfg;
    const bool main_mode = !cfg.pretrain_mode;
    float* amu = smem;                 // [K*D] per-CTA grad accumulators
    float* alv = amu + K * D;          // [K*D]
    float* cmu = alv + K * D;          // [K*D]
    float* cis2 = cmu + K * D;         // 1/sd^2
    float* cpass = cis2 + K * D;       // sd clamp pass flag
    float* M2 = cpass + K * D;         // [D*D]
    float* R = M2 + D * D;             // [K*D]
    float* cdq = R + K * D;            // [K]
    float* wsc = cdq + K;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* zsh = wsc + warp * (D + 2 * K);
    float* dlsh = zsh + D;
    float* qnsh = dlsh + K;
    for (int i = threadIdx.x; i < K * D; i += blockDim.x) {
        amu[i] = 0.f; alv[i] = 0.f;
        cmu[i] = __ldg(a.gmm_mu + i);
        float e = expf(0.5f * __ldg(a.gmm_lv + i));
        float sd = fmaxf(e, 1e-3f);
        cis2[i] = 1.0f / (sd * sd);
        cpass[i] = (e >= 1e-3f) ? 1.f : 0.f;
        R[i] = a.coef[CL.R + i];
    }
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) M2[i] = a.coef[CL.M2 + i];
    for (int i = threadIdx.x; i < K; i += blockDim.x) cdq[i] = a.coef[CL.dq + i];
    __syncthreads();
    const float klscale = a.coef[CF_KLSCALE];
    const float dscale = a.coef[CF_DISTSCALE];
    const float actc = a.coef[CF_ACT];
    const float tmpc = a.coef[CF_TEMPORAL];
    for (int b = blockIdx.x * LS_WARPS + warp; b < B; b += gridDim.x * LS_WARPS) {
        float qr, tot;
        float qn = qnorm_row(a.q, b, K, lane, &qr, &tot);
        float dq = (lane < K) ? cdq[lane] : 0.f;
        if (dscale != 0.f) {
            float tb = (lane < K) ? a.tau[(size_t)b * K + lane] : 0.f;
            if (cfg.distill_sharpen_T > 0.f) {
                float lg = (lane < K) ? logf(fmaxf(tb, 1e-8f)) / cfg.distill_sharpen_T : -INFINITY;
                float mx = warp_max(lg);
                float e = (lane < K) ? expf(lg - mx) : 0.f;
                tb = e / warp_sum(e);
            }
            if (lane < K && qn >= 1e-8f) dq += -dscale * a.distw[b] * tb / qn;
        }
        if (tmpc != 0.f) {
            if (b > 0) {
                float qp = qnorm_row(a.q, b - 1, K, lane, nullptr, nullptr);
                float df = qn - qp;
                if (lane < K) dq += tmpc * (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f));
            }
            if (b + 1 < B) {
                float qx = qnorm_row(a.q, b + 1, K, lane, nullptr, nullptr);
                float df = qx - qn;
                if (lane < K) dq -= tmpc * (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f));
            }
        }
        float inner = warp_sum((lane < K) ? dq * qn : 0.f);
        float dqc = (dq - inner) / tot;
        float dqr = (lane < K && qr >= 1e-8f) ? dqc : 0.f;
        float inner2 = warp_sum(qr * dqr);
        float dl = qr * (dqr - inner2);
        if (lane < K) { dlsh[lane] = dl; qnsh[lane] = qn; }
        for (int d = lane; d < D; d += 32) zsh[d] = a.z[(size_t)b * D + d];
        __syncwarp();
        for (int d = lane; d < D; d += 32) {
            float zv = zsh[d];
            float dz = a.dz_dec ? a.dz_dec[(size_t)b * D + d] : 0.f;
            for (int k = 0; k < D; k++) dz = fmaf(zsh[k], M2[k * D + d], dz);
            for (int c = 0; c < K; c++) {
                dz = fmaf(qnsh[c], R[c * D + d], dz);
                dz = fmaf(-dlsh[c] * (zv - cmu[c * D + d]), cis2[c * D + d], dz);
            }
            float lv = a.lv[(size_t)b * D + d];
            float pre = a.pre[(size_t)b * D + d];
            float zmv = a.zm[(size_t)b * D + d];
            float dzm = dz;
            const float epsv = a.eps ? a.eps[(size_t)b * D + d] : philox_normal(a.noise_seed, DOF_SITE_EPS, (unsigned long long)b * D + d);
            float dlv = dz * 0.5f * expf(0.5f * lv) * epsv;
            dlv += actc * (lv > 0.f ? 1.f : (lv < 0.f ? -1.f : 0.f));
            const bool inclamp = (lv >= -4.f && lv <= 2.f);
            if (!main_mode) {
                float lvc = fminf(fmaxf(lv, -4.f), 2.f);
                dzm += klscale * zmv;
                if (inclamp) dlv += klscale * 0.5f * (expf(lvc) - 1.0f);
            } else if (klscale != 0.f) {
                dzm += klscale * a.dzm_kl[(size_t)b * D + d];
                if (inclamp) dlv += klscale * a.dlv_kl[(size_t)b * D + d];
            }
            float sg = pre > 20.f ? 1.f : 1.0f / (1.0f + expf(-pre));
            a.dzm[(size_t)b * D + d] = dzm;
            a.dpre[(size_t)b * D + d] = dlv * sg;
        }
        for (int p = lane; p < K * D; p += 32) {
            int c = p / D, d = p % D;
            float df = zsh[d] - cmu[p];
            float dlc = dlsh[c];
            atomicAdd(&amu[p], dlc * df * cis2[p]);
            atomicAdd(&alv[p], 0.5f * dlc * (df * df * cis2[p] - 1.0f) * cpass[p]);
        }
        __syncwarp();
    }
    __syncthreads();
    if (a.g_mu) {
        for (int i = threadIdx.x; i < K * D; i += blockDim.x) {
            if (amu[i] != 0.f) atomicAdd(a.g_mu + i, amu[i]);
            if (alv[i] != 0.f) atomicAdd(a.g_lv + i, alv[i]);
        }
    }
}

static inline size_t loss_grad_smem_floats(int D, int K) {
    return (size_t)6 * K * D + (size_t)D * D + K + LS_WARPS * ((size_t)D + 2 * K);
}
