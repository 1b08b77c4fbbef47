// Element-wise / reduction kernels of the TCN model family, sm_100a.
// Reference: deepof/clustering/models_new.py  TemporalBlockPT :376-445, TCN1DPT :447-503.
//
// A temporal block is  conv1 -> BN -> ReLU -> conv2 -> BN -> ReLU (= skip) ;  out = ReLU(skip + residual).  The dilated causal
// convolutions run as GEMMs over time-shifted views of the [rows = (sequence, step), channels] activations (A_TAPS operand of
// gemm_rows / gemm_rows_tc, tcgen05 3xTF32) and their weight gradients as one time-shifted weight-gradient GEMM per tap.
// Train-mode BatchNorm couples ALL rows of a layer, so a layer cannot be fused across the statistics: the kernels here are the
// HBM-bound passes between the GEMMs — per-channel sums (fp32 partials, fp64 across threads / CTAs), normalise + ReLU,
// block output + skip accumulation, and the two-pass BatchNorm backward.  Rows are handled as float4 (C is a multiple of 4).
#pragma once
#include "common.cuh"

#define TCN_MAXC 128

// how a kernel obtains y = x * scale + shift of one BatchNorm layer
struct TcnBn {
    const double* stat;                     // train: [3C] sum (x - c) | sum (x - c)^2 | c over the rows of this pass (c = the first
                                            // row: a shifted one-pass variance, immune to mean^2 >> var); null: running buffers
    double inv_n;                           // 1 / rows
    const float* rmean; const float* rvar;  // eval
    const float* w; const float* b;
    float eps;
};

__device__ __forceinline__ void tcn_bn_load(const TcnBn& bn, int C, float* s_scale, float* s_shift, float* s_mean, float* s_rstd) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double m, v;
        if (bn.stat) {
            const double d = bn.stat[c] * bn.inv_n;
            v = bn.stat[C + c] * bn.inv_n - d * d;
            if (v < 0.0) v = 0.0;
            m = d + bn.stat[2 * C + c];
        } else {
            m = (double)bn.rmean[c];
            v = (double)bn.rvar[c];
        }
        const float rstd = (float)(1.0 / sqrt(v + (double)bn.eps));
        const float sc = bn.w[c] * rstd;
        s_scale[c] = sc;
        s_shift[c] = bn.b[c] - (float)m * sc;
        if (s_mean) { s_mean[c] = (float)m; s_rstd[c] = rstd; }
    }
    __syncthreads();
}

// thread -> (row lane, channel quad): q = C / 4 float4 per row, 256 / q rows per CTA pass
struct TcnLane { int cq, rl, rpb; bool live; };
__device__ __forceinline__ TcnLane tcn_lane(int C) {
    TcnLane l;
    const int q = C >> 2;
    l.rpb = blockDim.x / q;
    l.cq = threadIdx.x % q;
    l.rl = threadIdx.x / q;
    l.live = l.rl < l.rpb;
    return l;
}

// ---- per-channel shifted sums of A [R, C] -> stat [3C]: sum (x - c) (+=) | sum (x - c)^2 (+=) | c = A[0, :] ------------------
__global__ void __launch_bounds__(256) tcn_stats_kernel(const float* __restrict__ A, long long R, int C, double* __restrict__ stat) {
    __shared__ double s_acc[2 * TCN_MAXC];
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0.0;
    if (blockIdx.x == 0) for (int i = threadIdx.x; i < C; i += blockDim.x) stat[2 * C + i] = (double)__ldg(A + i);
    __syncthreads();
    const TcnLane l = tcn_lane(C);
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
    if (l.live) {
        const float4 c0 = __ldg(reinterpret_cast<const float4*>(A) + l.cq);
        const long long stride = (long long)gridDim.x * l.rpb;
        // four rows in flight per thread (one load per iteration left the kernel at 2.3 TB/s, long-scoreboard bound)
        for (long long r = (long long)blockIdx.x * l.rpb + l.rl; r < R; r += 4 * stride) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const long long ru = r + u * stride;
                v[u] = ru < R ? __ldg(reinterpret_cast<const float4*>(A + ru * C) + l.cq) : c0;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const float dx = v[u].x - c0.x, dy = v[u].y - c0.y, dz = v[u].z - c0.z, dw = v[u].w - c0.w;
                s[0] += dx; s[1] += dy; s[2] += dz; s[3] += dw;
                q[0] = fmaf(dx, dx, q[0]); q[1] = fmaf(dy, dy, q[1]); q[2] = fmaf(dz, dz, q[2]); q[3] = fmaf(dw, dw, q[3]);
            }
        }
    }
    if (l.live)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            atomicAdd(&s_acc[l.cq * 4 + j], (double)s[j]);
            atomicAdd(&s_acc[C + l.cq * 4 + j], (double)q[j]);
        }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(stat + i, s_acc[i]);
}

// ---- Y = relu(bn(A)) -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tcn_bnrelu_kernel(const float* __restrict__ A, float* __restrict__ Y, long long R, int C, const TcnBn bn) {
    __shared__ float s_scale[TCN_MAXC], s_shift[TCN_MAXC];
    tcn_bn_load(bn, C, s_scale, s_shift, nullptr, nullptr);
    const TcnLane l = tcn_lane(C);
    if (!l.live) return;
    const float4 sc = *reinterpret_cast<const float4*>(s_scale + l.cq * 4), sh = *reinterpret_cast<const float4*>(s_shift + l.cq * 4);
    for (long long r = (long long)blockIdx.x * l.rpb + l.rl; r < R; r += (long long)gridDim.x * l.rpb) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(A + r * C) + l.cq);
        float4 o;
        o.x = fmaxf(fmaf(v.x, sc.x, sh.x), 0.f); o.y = fmaxf(fmaf(v.y, sc.y, sh.y), 0.f);
        o.z = fmaxf(fmaf(v.z, sc.z, sh.z), 0.f); o.w = fmaxf(fmaf(v.w, sc.w, sh.w), 0.f);
        reinterpret_cast<float4*>(Y + r * C)[l.cq] = o;
    }
}

// ---- block output: y2 = relu(bn2(A2)); OUT = relu(y2 + RES); SKIP (+)= y2; FIN = relu(SKIP) on the last block ------------------
struct TcnOutArgs {
    const float* A2; const float* RES;      // [R, C]
    float* OUT;                             // [R, C] or null (last block: nothing reads it)
    float* SKIP;                            // [R, C], or [R / T, C] when last_only (only the final step of a sequence is used)
    float* FIN;                             // same shape as SKIP or null
    long long R; int C, T, first, last_only;
    TcnBn bn;
};

__global__ void __launch_bounds__(256) tcn_block_out_kernel(const TcnOutArgs a) {
    __shared__ float s_scale[TCN_MAXC], s_shift[TCN_MAXC];
    tcn_bn_load(a.bn, a.C, s_scale, s_shift, nullptr, nullptr);
    const TcnLane l = tcn_lane(a.C);
    if (!l.live) return;
    const int C = a.C;
    const float4 sc = *reinterpret_cast<const float4*>(s_scale + l.cq * 4), sh = *reinterpret_cast<const float4*>(s_shift + l.cq * 4);
    for (long long r = (long long)blockIdx.x * l.rpb + l.rl; r < a.R; r += (long long)gridDim.x * l.rpb) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(a.A2 + r * C) + l.cq);
        float4 y;
        y.x = fmaxf(fmaf(v.x, sc.x, sh.x), 0.f); y.y = fmaxf(fmaf(v.y, sc.y, sh.y), 0.f);
        y.z = fmaxf(fmaf(v.z, sc.z, sh.z), 0.f); y.w = fmaxf(fmaf(v.w, sc.w, sh.w), 0.f);
        if (a.OUT) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(a.RES + r * C) + l.cq);
            reinterpret_cast<float4*>(a.OUT + r * C)[l.cq] =
                make_float4(fmaxf(y.x + x.x, 0.f), fmaxf(y.y + x.y, 0.f), fmaxf(y.z + x.z, 0.f), fmaxf(y.w + x.w, 0.f));
        }
        long long sr = r;
        if (a.last_only) {
            if ((int)(r % a.T) != a.T - 1) continue;
            sr = r / a.T;
        }
        float4* sp = reinterpret_cast<float4*>(a.SKIP + sr * C) + l.cq;
        if (!a.first) { const float4 p = *sp; y.x += p.x; y.y += p.y; y.z += p.z; y.w += p.w; }
        *sp = y;
        if (a.FIN) reinterpret_cast<float4*>(a.FIN + sr * C)[l.cq] = make_float4(fmaxf(y.x, 0.f), fmaxf(y.y, 0.f), fmaxf(y.z, 0.f), fmaxf(y.w, 0.f));
    }
}

// ---- backward, pass 1: the gradient behind a BatchNorm + ReLU and its two per-channel sums -------------------------------------
// mode 2 (second conv of a block):  dS = dOUT * (OUT > 0)  -> DS;   dy = (dS + GS) * (bn(A) > 0) -> D
// mode 1 (first conv):              D already holds dy (the input-gradient GEMM applied the ReLU mask)
// bstat [2C] += sum dy | sum dy * xhat
struct TcnBwdArgs {
    const float* dOUT; const float* OUT;    // mode 2, both null for the last block
    const float* GS;                        // mode 2: gradient of the skip sum, [R, C] or [R / T, C] (gs_last_only)
    float* DS;                              // mode 2: [R, C] (written even when dOUT is null: zeros)
    float* D;                               // [R, C]
    const float* A;                         // pre-BatchNorm conv output
    double* bstat;
    long long R; int C, T, mode, gs_last_only;
    TcnBn bn;
};

__global__ void __launch_bounds__(256) tcn_bwd_reduce_kernel(const TcnBwdArgs a) {
    __shared__ float s_scale[TCN_MAXC], s_shift[TCN_MAXC], s_mean[TCN_MAXC], s_rstd[TCN_MAXC];
    __shared__ double s_acc[2 * TCN_MAXC];
    const int C = a.C;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0.0;
    tcn_bn_load(a.bn, C, s_scale, s_shift, s_mean, s_rstd);
    const TcnLane l = tcn_lane(C);
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    if (l.live) {
        float sc[4], sh[4], mu[4], rs[4];
#pragma unroll
        for (int j = 0; j < 4; j++) { sc[j] = s_scale[l.cq * 4 + j]; sh[j] = s_shift[l.cq * 4 + j]; mu[j] = s_mean[l.cq * 4 + j]; rs[j] = s_rstd[l.cq * 4 + j]; }
        for (long long r = (long long)blockIdx.x * l.rpb + l.rl; r < a.R; r += (long long)gridDim.x * l.rpb) {
            const float4 av4 = __ldg(reinterpret_cast<const float4*>(a.A + r * C) + l.cq);
            const float av[4] = {av4.x, av4.y, av4.z, av4.w};
            float d[4];
            if (a.mode == 2) {
                float ds[4] = {0.f, 0.f, 0.f, 0.f};
                if (a.dOUT) {
                    const float4 g4 = *(reinterpret_cast<const float4*>(a.dOUT + r * C) + l.cq);      // DS may alias dOUT: plain load
                    const float4 o4 = __ldg(reinterpret_cast<const float4*>(a.OUT + r * C) + l.cq);
                    ds[0] = o4.x > 0.f ? g4.x : 0.f; ds[1] = o4.y > 0.f ? g4.y : 0.f;
                    ds[2] = o4.z > 0.f ? g4.z : 0.f; ds[3] = o4.w > 0.f ? g4.w : 0.f;
                }
                reinterpret_cast<float4*>(a.DS + r * C)[l.cq] = make_float4(ds[0], ds[1], ds[2], ds[3]);
                float gs[4] = {0.f, 0.f, 0.f, 0.f};
                long long sr = r;
                bool has = true;
                if (a.gs_last_only) { has = (int)(r % a.T) == a.T - 1; sr = r / a.T; }
                if (has) {
                    const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.GS + sr * C) + l.cq);
                    gs[0] = g4.x; gs[1] = g4.y; gs[2] = g4.z; gs[3] = g4.w;
                }
#pragma unroll
                for (int j = 0; j < 4; j++) d[j] = fmaf(av[j], sc[j], sh[j]) > 0.f ? ds[j] + gs[j] : 0.f;
                reinterpret_cast<float4*>(a.D + r * C)[l.cq] = make_float4(d[0], d[1], d[2], d[3]);
            } else {
                const float4 d4 = __ldg(reinterpret_cast<const float4*>(a.D + r * C) + l.cq);
                d[0] = d4.x; d[1] = d4.y; d[2] = d4.z; d[3] = d4.w;
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                s1[j] += d[j];
                s2[j] = fmaf(d[j], (av[j] - mu[j]) * rs[j], s2[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            atomicAdd(&s_acc[l.cq * 4 + j], (double)s1[j]);
            atomicAdd(&s_acc[C + l.cq * 4 + j], (double)s2[j]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(a.bstat + i, s_acc[i]);
}

// ---- backward, pass 2: D <- w * rstd * (D - mean(dy) - xhat * mean(dy xhat)) in place; dw += sum dy xhat, db += sum dy -----------
__global__ void __launch_bounds__(256) tcn_bn_bwd_apply_kernel(float* __restrict__ D, const float* __restrict__ A, long long R, int C,
                                                               const TcnBn bn, const double* __restrict__ bstat, float* __restrict__ dw,
                                                               float* __restrict__ db) {
    __shared__ float s_scale[TCN_MAXC], s_shift[TCN_MAXC], s_mean[TCN_MAXC], s_rstd[TCN_MAXC], s_m1[TCN_MAXC], s_m2[TCN_MAXC];
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        s_m1[c] = (float)(bstat[c] * bn.inv_n);
        s_m2[c] = (float)(bstat[C + c] * bn.inv_n);
        if (blockIdx.x == 0) { atomicAdd(dw + c, (float)bstat[C + c]); atomicAdd(db + c, (float)bstat[c]); }
    }
    tcn_bn_load(bn, C, s_scale, s_shift, s_mean, s_rstd);
    const TcnLane l = tcn_lane(C);
    if (!l.live) return;
    float sc[4], mu[4], rs[4], m1[4], m2[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int c = l.cq * 4 + j;
        sc[j] = s_scale[c]; mu[j] = s_mean[c]; rs[j] = s_rstd[c]; m1[j] = s_m1[c]; m2[j] = s_m2[c];
    }
    for (long long r = (long long)blockIdx.x * l.rpb + l.rl; r < R; r += (long long)gridDim.x * l.rpb) {
        const float4 a4 = __ldg(reinterpret_cast<const float4*>(A + r * C) + l.cq);
        float4* dp = reinterpret_cast<float4*>(D + r * C) + l.cq;
        const float4 d4 = *dp;
        const float av[4] = {a4.x, a4.y, a4.z, a4.w}, dv[4] = {d4.x, d4.y, d4.z, d4.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) o[j] = sc[j] * (dv[j] - m1[j] - (av[j] - mu[j]) * rs[j] * m2[j]);
        *dp = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// ---- small helpers ---------------------------------------------------------------------------------------------------
// Xs[s, t, f] = x[b, gidx[g, t, f]]  (the reference's tf_style reshape of a window), s = b * G + g
// rows of Xs have pitch ldx >= F (padded to a multiple of 4 floats for the float4 GEMM producers; the pad columns are zeroed by the caller)
__global__ void tcn_gather_kernel(const float* __restrict__ x, const int* __restrict__ gidx, float* __restrict__ Xs, long long n, int G, int TF,
                                  int F, int ldx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long s = i / TF;
    const int rem = (int)(i - s * TF);
    const long long b = s / G;
    const int g = (int)(s - b * G);
    const long long row = i / F;
    const int f = (int)(i - row * F);
    Xs[row * ldx + f] = __ldg(x + b * (long long)G * TF + gidx[g * TF + rem]);
}

// out[b * T + t, :] = in[b, :]
__global__ void tcn_repeat_kernel(const float* __restrict__ in, float* __restrict__ out, long long n4, int T, int C4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const long long row = i / C4;
    const int c = (int)(i - row * C4);
    reinterpret_cast<float4*>(out)[i] = __ldg(reinterpret_cast<const float4*>(in) + (row / T) * C4 + c);
}

// out = ref > 0 ? g : 0
__global__ void tcn_relu_mask_kernel(const float* __restrict__ g, const float* __restrict__ ref, float* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = ref[i] > 0.f ? g[i] : 0.f;
}

// running buffers after a training step: for every pass g, running = (1 - m) * running + m * (mean, unbiased variance)
struct TcnBnDesc { long long mean, var, tracked; int stat, C, rows_per_window, decoder; };
__global__ void tcn_bn_update_kernel(float* __restrict__ state, const TcnBnDesc* __restrict__ desc, const double* __restrict__ stat,
                                     long long stat_stride, int enc_passes, int enc_windows, int dec_passes, int dec_windows, float momentum) {
    const TcnBnDesc d = desc[blockIdx.x];
    const int passes = d.decoder ? dec_passes : enc_passes;
    const double n = (double)(d.decoder ? dec_windows : enc_windows) * d.rows_per_window;
    if (passes < 1) return;
    for (int c = threadIdx.x; c < d.C; c += blockDim.x) {
        float m = state[d.mean + c], v = state[d.var + c];
        for (int g = 0; g < passes; g++) {
            const double* s = stat + (long long)g * stat_stride + d.stat;
            const double dm = s[c] / n;
            double var = s[d.C + c] / n - dm * dm;
            if (var < 0.0) var = 0.0;
            const double mu = dm + s[2 * d.C + c];
            m = (1.f - momentum) * m + momentum * (float)mu;
            v = (1.f - momentum) * v + momentum * (float)(var * (n / (n > 1.0 ? n - 1.0 : 1.0)));
        }
        state[d.mean + c] = m; state[d.var + c] = v;
    }
    if (threadIdx.x == 0) state[d.tracked] += (float)passes;
}
