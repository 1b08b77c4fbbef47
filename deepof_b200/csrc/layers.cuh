// Non-GEMM layers of the VaDE / recurrent path: the reference's group-reshape gather
// fused with the encoder Conv1d+ReLU, LayerNorm fwd/bwd, the CensNet graph aggregation,
// the Gaussian-mixture latent head, and fused clip+Adam.
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------
// Encoder stage 1: gather (SURVEY A.1 scramble; models_new.py:120-138) + Conv1d(F->C,
// k=5, 'same', no bias) + ReLU (models_new.py:228-230) + valid-length count (:233-234).
// One CTA per window.  Writes the gathered sequences Xs[S,T,F] (needed by the conv
// weight gradient), Cv[S,T,C] and len[S], with s = b*G + g.
// ---------------------------------------------------------------------------
struct EncConvArgs {
    const float* x;     // [B, T*G*F]
    const int* gidx;    // [G*T*F] index into a window for (g,t',f)
    const float* w;     // [C,F,5]
    float* Xs; float* Cv; int* len;
    int B, T, G, F, C;
};

__global__ void __launch_bounds__(256) enc_conv_kernel(const EncConvArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int T = a.T, G = a.G, F = a.F, C = a.C;
    const int W = T * G * F;
    float* seq = smem;                 // [G][T][F]
    float* w = seq + W;                // [C][F*5]
    int* flag = reinterpret_cast<int*>(w + C * F * 5);   // [G*T]
    const int b = blockIdx.x;
    const float* xw = a.x + (size_t)b * W;
    for (int i = threadIdx.x; i < W; i += blockDim.x) {
        float v = __ldg(xw + __ldg(a.gidx + i));
        seq[i] = v;
        a.Xs[(size_t)b * W + i] = v;   // [b*G+g][t][f] is exactly b*W + i
    }
    for (int i = threadIdx.x; i < C * F * 5; i += blockDim.x) w[i] = __ldg(a.w + i);
    for (int i = threadIdx.x; i < G * T; i += blockDim.x) flag[i] = 0;
    __syncthreads();
    float* out = a.Cv + (size_t)b * G * T * C;
    if (C == 32 && F <= 4) {
        // lane = output channel (its F*5 taps live in registers), warp = sequence g: the input window slides through
        // registers (F broadcast shared-memory loads per step), stores are one full 128-byte line per warp instruction,
        // the valid length comes from warp votes
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
        float wr[4][5];
#pragma unroll
        for (int f = 0; f < 4; f++)
#pragma unroll
            for (int kk = 0; kk < 5; kk++) wr[f][kk] = f < F ? w[lane * F * 5 + f * 5 + kk] : 0.f;
        for (int g = warp; g < G; g += nw) {
            const float* sq = seq + (size_t)g * T * F;
            float xw[4][5];
#pragma unroll
            for (int f = 0; f < 4; f++) {
                xw[f][0] = 0.f; xw[f][1] = 0.f;
                xw[f][2] = f < F ? sq[f] : 0.f;
                xw[f][3] = (f < F && T > 1) ? sq[F + f] : 0.f;
                xw[f][4] = (f < F && T > 2) ? sq[2 * F + f] : 0.f;
            }
            int n = 0;
            for (int t = 0; t < T; t++) {
                float acc = 0.f;
#pragma unroll
                for (int f = 0; f < 4; f++)
#pragma unroll
                    for (int kk = 0; kk < 5; kk++) acc = fmaf(wr[f][kk], xw[f][kk], acc);
                acc = fmaxf(acc, 0.f);
                out[((size_t)g * T + t) * C + lane] = acc;
                n += __any_sync(0xffffffffu, acc > 0.f) ? 1 : 0;
#pragma unroll
                for (int f = 0; f < 4; f++) {
                    xw[f][0] = xw[f][1]; xw[f][1] = xw[f][2]; xw[f][2] = xw[f][3]; xw[f][3] = xw[f][4];
                    xw[f][4] = (f < F && t + 3 < T) ? sq[(t + 3) * F + f] : 0.f;
                }
            }
            if (lane == 0) a.len[(size_t)b * G + g] = n;
        }
        return;
    }
    for (int i = threadIdx.x; i < G * T * C; i += blockDim.x) {
        int co = i % C, gt = i / C;
        int t = gt % T, g = gt / T;
        const float* sq = seq + (size_t)g * T * F;
        const float* wc = w + co * F * 5;
        float acc = 0.f;
        for (int f = 0; f < F; f++) {
#pragma unroll
            for (int kk = 0; kk < 5; kk++) {
                int tt = t + kk - 2;
                if (tt >= 0 && tt < T) acc = fmaf(wc[f * 5 + kk], sq[tt * F + f], acc);
            }
        }
        acc = fmaxf(acc, 0.f);
        out[i] = acc;
        if (acc > 0.f) flag[gt] = 1;
    }
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        int n = 0;
        for (int t = 0; t < T; t++) n += flag[g * T + t];
        a.len[(size_t)b * G + g] = n;
    }
}

// ---------------------------------------------------------------------------
// Weight gradient of the encoder Conv1d(F -> C, k=5, 'same', no bias):
//   dW[c][f][kk] += sum_{s,t} dCv[s,t,c] * Xs[s, t+kk-2, f]      (zero outside [0,T))
// One thread per (c, f) pair and sequence slot; the 5-tap window of Xs slides through registers, so a step costs
// one coalesced read of dCv (lanes = consecutive c) and one broadcast read of Xs for 5 FMAs.  No shared-memory
// staging, no tensor core: 32 x 15 outputs over 1.4 M rows is a pure stream of dCv (HBM-bound).
// ---------------------------------------------------------------------------
struct ConvWgradArgs { const float* dCv; const float* Xs; float* dW; int S, T, C, F; };

__global__ void __launch_bounds__(256) conv_wgrad_kernel(const ConvWgradArgs a) {
    __shared__ float red[256 * 5];
    const int pairs = a.C * a.F, slots = blockDim.x / pairs;
    const int tid = threadIdx.x;
    const int pair = tid % pairs, slot = tid / pairs;
    const int c = pair % a.C, f = pair / a.C;
    const int T = a.T, C = a.C, F = a.F;
    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (slot < slots) {
        for (long long seq = (long long)blockIdx.x * slots + slot; seq < a.S; seq += (long long)gridDim.x * slots) {
            const float* xs = a.Xs + seq * T * F + f;
            const float* dc = a.dCv + seq * T * C + c;
            float xm2 = 0.f, xm1 = 0.f, x0 = xs[0], xp1 = T > 1 ? xs[F] : 0.f, xp2 = T > 2 ? xs[2 * F] : 0.f;
#pragma unroll 5
            for (int t = 0; t < T; t++) {
                const float d = dc[(size_t)t * C];
                acc[0] = fmaf(d, xm2, acc[0]); acc[1] = fmaf(d, xm1, acc[1]); acc[2] = fmaf(d, x0, acc[2]);
                acc[3] = fmaf(d, xp1, acc[3]); acc[4] = fmaf(d, xp2, acc[4]);
                xm2 = xm1; xm1 = x0; x0 = xp1; xp1 = xp2;
                xp2 = (t + 3 < T) ? xs[(size_t)(t + 3) * F] : 0.f;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 5; k++) red[tid * 5 + k] = (slot < slots) ? acc[k] : 0.f;
    __syncthreads();
    if (tid < pairs) {
#pragma unroll
        for (int k = 0; k < 5; k++) {
            float s = 0.f;
            for (int q = 0; q < slots; q++) s += red[(q * pairs + tid) * 5 + k];
            atomicAdd(a.dW + (size_t)c * F * 5 + f * 5 + k, s);
        }
    }
}

// Same gradient, one WARP per sequence: the [T, C] block of dCv is one contiguous run that the warp streams with
// float4 loads.  C/4 lanes cover a time step; the 32/(C/4) lane groups of the warp each own a run of consecutive
// steps, so the 5-tap window of Xs SLIDES through registers (F new values per step instead of 5 F) and every warp load
// still covers whole 128-byte lines; four loads are in flight per lane.  Each lane keeps the 4 x F x 5 partial sums
// of its four channels.
template <int C, int F>
__global__ void __launch_bounds__(256) conv_wgrad_vec_kernel(const ConvWgradArgs a) {
    constexpr int LPR = C / 4, RPW = 32 / LPR, NW = C * F * 5;
    __shared__ float sdw[NW];
    const int lane = threadIdx.x & 31, c4 = (lane % LPR) * 4, grp = lane / LPR, T = a.T;
    const int chunk = (T + RPW - 1) / RPW, tb = grp * chunk, te = min(T, tb + chunk);
    for (int i = threadIdx.x; i < NW; i += blockDim.x) sdw[i] = 0.f;
    __syncthreads();
    float acc[4][F][5];
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int f = 0; f < F; f++)
#pragma unroll
            for (int k = 0; k < 5; k++) acc[j][f][k] = 0.f;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long seq = warp; seq < a.S; seq += nwarps) {
        const float* __restrict__ dc = a.dCv + seq * T * C + c4;
        const float* __restrict__ xs = a.Xs + seq * T * F;
        float w[5][F];              // at step t (after the shift): w[k] = Xs[t + k - 2]
#pragma unroll
        for (int k = 1; k < 5; k++) {
            const int tt = tb + k - 3;
#pragma unroll
            for (int f = 0; f < F; f++) w[k][f] = (tt >= 0 && tt < T) ? __ldg(xs + tt * F + f) : 0.f;
        }
        for (int t0 = tb; t0 < te; t0 += 4) {
            float4 d[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int t = t0 + q;
                d[q] = (t < te) ? __ldcs(reinterpret_cast<const float4*>(dc + (size_t)t * C)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int t = t0 + q;
                if (t >= te) break;
#pragma unroll
                for (int f = 0; f < F; f++) {
#pragma unroll
                    for (int k = 0; k < 4; k++) w[k][f] = w[k + 1][f];
                    w[4][f] = (t + 2 < T) ? __ldg(xs + (t + 2) * F + f) : 0.f;
                }
#pragma unroll
                for (int k = 0; k < 5; k++)
#pragma unroll
                    for (int f = 0; f < F; f++) {
                        const float xv = w[k][f];
                        acc[0][f][k] = fmaf(d[q].x, xv, acc[0][f][k]);
                        acc[1][f][k] = fmaf(d[q].y, xv, acc[1][f][k]);
                        acc[2][f][k] = fmaf(d[q].z, xv, acc[2][f][k]);
                        acc[3][f][k] = fmaf(d[q].w, xv, acc[3][f][k]);
                    }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int f = 0; f < F; f++)
#pragma unroll
            for (int k = 0; k < 5; k++) {
                float v = acc[j][f][k];
#pragma unroll
                for (int o = LPR; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (grp == 0) atomicAdd(&sdw[(c4 + j) * F * 5 + f * 5 + k], v);
            }
    __syncthreads();
    for (int i = threadIdx.x; i < NW; i += blockDim.x) atomicAdd(a.dW + i, sdw[i]);
}

template <int C>
static bool launch_conv_wgrad_vec(const ConvWgradArgs& ca, int sm, cudaStream_t st) {
    if ((((uintptr_t)ca.dCv) & 15) != 0) return false;
    long long want = ((long long)ca.S + 7) / 8;
    int grid = (int)(want < (long long)sm * 4 ? want : (long long)sm * 4);
    if (grid < 1) grid = 1;
    if (ca.F == 1) conv_wgrad_vec_kernel<C, 1><<<grid, 256, 0, st>>>(ca);
    else if (ca.F == 2) conv_wgrad_vec_kernel<C, 2><<<grid, 256, 0, st>>>(ca);
    else if (ca.F == 3) conv_wgrad_vec_kernel<C, 3><<<grid, 256, 0, st>>>(ca);
    else return false;
    return true;
}

// decoder validity: len[b] = #time steps whose data row is not all-zero (models_new.py:330-331).
// One warp per window; a time step's row (Dx floats, contiguous) is read by the whole warp.
__global__ void row_valid_len_kernel(const float* __restrict__ x, int* __restrict__ len, int B, int T, int Dx) {
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= B) return;
    int n = 0;
    for (int t = 0; t < T; t++) {
        const float* r = x + ((size_t)b * T + t) * Dx;
        bool any = false;
        for (int d = lane; d < Dx; d += 32) any |= (r[d] != 0.f);
        n += __any_sync(0xffffffffu, any) ? 1 : 0;
    }
    if (lane == 0) len[b] = n;
}

// ---------------------------------------------------------------------------
// LayerNorm over the last dim (biased variance), one warp per row, W <= 256.
// ---------------------------------------------------------------------------
#define LN_MAXV 8
// VPL values per lane (W <= 32*VPL), R rows per warp iteration: the R rows' loads are issued together so that
// each warp keeps R*W*4 bytes in flight (these kernels are pure HBM streams).
template <int VPL, int R>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                     const float* __restrict__ b, float eps,
                                                     float* __restrict__ y, float* __restrict__ mu_out,
                                                     float* __restrict__ rstd_out, long long Rows, int W) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    float wv[VPL], bv[VPL];
#pragma unroll
    for (int i = 0; i < VPL; i++) {
        const int c = lane + 32 * i;
        wv[i] = (c < W) ? __ldg(w + c) : 0.f;
        bv[i] = (c < W) ? __ldg(b + c) : 0.f;
    }
    for (long long r0 = warp; r0 < Rows; r0 += nwarps * R) {
        float v[R][VPL], s[R];
#pragma unroll
        for (int q = 0; q < R; q++) {
            const long long r = r0 + q * nwarps;
            s[q] = 0.f;
#pragma unroll
            for (int i = 0; i < VPL; i++) {
                const int c = lane + 32 * i;
                v[q][i] = (r < Rows && c < W) ? x[r * W + c] : 0.f;
                s[q] += v[q][i];
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int q = 0; q < R; q++) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
        float mu[R], var[R];
#pragma unroll
        for (int q = 0; q < R; q++) {
            mu[q] = s[q] / W;
            var[q] = 0.f;
#pragma unroll
            for (int i = 0; i < VPL; i++) {
                const int c = lane + 32 * i;
                const float d = (c < W) ? v[q][i] - mu[q] : 0.f;
                var[q] += d * d;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int q = 0; q < R; q++) var[q] += __shfl_xor_sync(0xffffffffu, var[q], o);
#pragma unroll
        for (int q = 0; q < R; q++) {
            const long long r = r0 + q * nwarps;
            if (r >= Rows) continue;
            const float vv = var[q] / W;
            float rstd = rsqrtf(vv + eps);
            // one Newton step: rsqrtf is ~2 ulp; keep LayerNorm at full fp32 accuracy
            rstd = rstd * (1.5f - 0.5f * (vv + eps) * rstd * rstd);
#pragma unroll
            for (int i = 0; i < VPL; i++) {
                const int c = lane + 32 * i;
                if (c < W) y[r * W + c] = (v[q][i] - mu[q]) * rstd * wv[i] + bv[i];
            }
            if (lane == 0 && mu_out) { mu_out[r] = mu[q]; rstd_out[r] = rstd; }
        }
    }
}

// dx = rstd*(g - mean(g) - xhat*mean(g*xhat)), g = dy*w; dw += dy*xhat; db += dy.
// relu_in: the LN input x is a ReLU output -> additionally mask dx by (x > 0).
template <int VPL, int R>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                     const float* __restrict__ mu_in,
                                                     const float* __restrict__ rstd_in,
                                                     const float* __restrict__ w, float* __restrict__ dx,
                                                     float* __restrict__ dw, float* __restrict__ db,
                                                     long long Rows, int W, int relu_in) {
    __shared__ float sdw[256], sdb[256];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) { sdw[i] = 0.f; sdb[i] = 0.f; }
    __syncthreads();
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    float adw[VPL], adb[VPL], wv[VPL];
#pragma unroll
    for (int i = 0; i < VPL; i++) {
        adw[i] = 0.f; adb[i] = 0.f;
        const int c = lane + 32 * i;
        wv[i] = (c < W) ? __ldg(w + c) : 0.f;
    }
    for (long long r0 = warp; r0 < Rows; r0 += nwarps * R) {
        float xv[R][VPL], dv[R][VPL], mu[R], rstd[R], s1[R], s2[R];
#pragma unroll
        for (int q = 0; q < R; q++) {
            const long long r = r0 + q * nwarps;
            const bool ok = r < Rows;
            mu[q] = ok ? mu_in[r] : 0.f;
            rstd[q] = ok ? rstd_in[r] : 0.f;
#pragma unroll
            for (int i = 0; i < VPL; i++) {
                const int c = lane + 32 * i;
                xv[q][i] = (ok && c < W) ? x[r * W + c] : 0.f;
                dv[q][i] = (ok && c < W) ? dy[r * W + c] : 0.f;
            }
        }
#pragma unroll
        for (int q = 0; q < R; q++) {
            s1[q] = 0.f; s2[q] = 0.f;
#pragma unroll
            for (int i = 0; i < VPL; i++) {
                const int c = lane + 32 * i;
                const float xh = (c < W) ? (xv[q][i] - mu[q]) * rstd[q] : 0.f;
                const float g = dv[q][i] * wv[i];
                s1[q] += g;
                s2[q] += g * xh;
                adw[i] += dv[q][i] * xh;
                adb[i] += dv[q][i];
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int q = 0; q < R; q++) {
                s1[q] += __shfl_xor_sync(0xffffffffu, s1[q], o);
                s2[q] += __shfl_xor_sync(0xffffffffu, s2[q], o);
            }
#pragma unroll
        for (int q = 0; q < R; q++) {
            const long long r = r0 + q * nwarps;
            if (r >= Rows) continue;
            const float m1 = s1[q] / W, m2 = s2[q] / W;
#pragma unroll
            for (int i = 0; i < VPL; i++) {
                const int c = lane + 32 * i;
                if (c < W) {
                    const float xh = (xv[q][i] - mu[q]) * rstd[q];
                    float o = rstd[q] * (dv[q][i] * wv[i] - m1 - xh * m2);
                    if (relu_in && !(xv[q][i] > 0.f)) o = 0.f;
                    dx[r * W + c] = o;
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < VPL; i++) {
        int c = lane + 32 * i;
        if (c < W) { atomicAdd(&sdw[c], adw[i]); atomicAdd(&sdb[c], adb[i]); }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < W; c += blockDim.x) {
        atomicAdd(dw + c, sdw[c]);
        atomicAdd(db + c, sdb[c]);
    }
}

// ---------------------------------------------------------------------------
// Vectorised LayerNorm for W in {32, 64, 128} and 16-byte aligned rows: W/4 lanes per row with one float4 each, so a
// warp load instruction covers 32/(W/4) CONSECUTIVE rows (512 contiguous bytes) and R of them are in flight per warp.
// Same arithmetic as the scalar kernels above (two-pass variance, Newton-refined rsqrt).
// ---------------------------------------------------------------------------
template <int W, int R>
__global__ void __launch_bounds__(256) ln_fwd_vec_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ b, float eps,
                                                         float* __restrict__ y, float* __restrict__ mu_out,
                                                         float* __restrict__ rstd_out, long long Rows) {
    constexpr int LPR = W / 4, RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, sub = lane / LPR, c4 = (lane % LPR) * 4;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long groups = (Rows + RPW - 1) / RPW;
    const float4 wv = __ldg(reinterpret_cast<const float4*>(w + c4));
    const float4 bv = __ldg(reinterpret_cast<const float4*>(b + c4));
    for (long long g0 = warp; g0 < groups; g0 += nwarps * R) {
        float4 v[R];
        float s[R];
#pragma unroll
        for (int q = 0; q < R; q++) {
            const long long r = (g0 + q * nwarps) * RPW + sub;
            v[q] = (r < Rows) ? __ldcs(reinterpret_cast<const float4*>(x + r * W + c4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            s[q] = (v[q].x + v[q].y) + (v[q].z + v[q].w);
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1)
#pragma unroll
            for (int q = 0; q < R; q++) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
        float mu[R], var[R];
#pragma unroll
        for (int q = 0; q < R; q++) {
            mu[q] = s[q] / W;
            const float d0 = v[q].x - mu[q], d1 = v[q].y - mu[q], d2 = v[q].z - mu[q], d3 = v[q].w - mu[q];
            var[q] = (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1)
#pragma unroll
            for (int q = 0; q < R; q++) var[q] += __shfl_xor_sync(0xffffffffu, var[q], o);
#pragma unroll
        for (int q = 0; q < R; q++) {
            const long long r = (g0 + q * nwarps) * RPW + sub;
            if (r >= Rows) continue;
            const float vv = var[q] / W;
            float rstd = rsqrtf(vv + eps);
            rstd = rstd * (1.5f - 0.5f * (vv + eps) * rstd * rstd);
            float4 o;
            o.x = (v[q].x - mu[q]) * rstd * wv.x + bv.x;
            o.y = (v[q].y - mu[q]) * rstd * wv.y + bv.y;
            o.z = (v[q].z - mu[q]) * rstd * wv.z + bv.z;
            o.w = (v[q].w - mu[q]) * rstd * wv.w + bv.w;
            *reinterpret_cast<float4*>(y + r * W + c4) = o;
            if (c4 == 0 && mu_out) { mu_out[r] = mu[q]; rstd_out[r] = rstd; }
        }
    }
}

template <int W, int R>
__global__ void __launch_bounds__(256) ln_bwd_vec_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                         const float* __restrict__ mu_in,
                                                         const float* __restrict__ rstd_in,
                                                         const float* __restrict__ w, float* __restrict__ dx,
                                                         float* __restrict__ dw, float* __restrict__ db,
                                                         long long Rows, int relu_in) {
    constexpr int LPR = W / 4, RPW = 32 / LPR;
    __shared__ float sdw[W], sdb[W];
    const int lane = threadIdx.x & 31, sub = lane / LPR, c4 = (lane % LPR) * 4;
    for (int i = threadIdx.x; i < W; i += blockDim.x) { sdw[i] = 0.f; sdb[i] = 0.f; }
    __syncthreads();
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long groups = (Rows + RPW - 1) / RPW;
    const float4 wv = __ldg(reinterpret_cast<const float4*>(w + c4));
    float4 adw = make_float4(0.f, 0.f, 0.f, 0.f), adb = adw;
    for (long long g0 = warp; g0 < groups; g0 += nwarps * R) {
        float4 xv[R], dv[R];
        float mu[R], rstd[R], s1[R], s2[R];
#pragma unroll
        for (int q = 0; q < R; q++) {
            const long long r = (g0 + q * nwarps) * RPW + sub;
            const bool ok = r < Rows;
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            xv[q] = ok ? __ldcs(reinterpret_cast<const float4*>(x + r * W + c4)) : z4;
            dv[q] = ok ? __ldcs(reinterpret_cast<const float4*>(dy + r * W + c4)) : z4;
            mu[q] = ok ? __ldg(mu_in + r) : 0.f;
            rstd[q] = ok ? __ldg(rstd_in + r) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < R; q++) {
            // xhat in place of x (rows beyond the end have x = mu = rstd = 0 -> xhat = 0, dy = 0)
            const float h0 = (xv[q].x - mu[q]) * rstd[q], h1 = (xv[q].y - mu[q]) * rstd[q];
            const float h2 = (xv[q].z - mu[q]) * rstd[q], h3 = (xv[q].w - mu[q]) * rstd[q];
            const float g0_ = dv[q].x * wv.x, g1 = dv[q].y * wv.y, g2 = dv[q].z * wv.z, g3 = dv[q].w * wv.w;
            s1[q] = (g0_ + g1) + (g2 + g3);
            s2[q] = (g0_ * h0 + g1 * h1) + (g2 * h2 + g3 * h3);
            adw.x += dv[q].x * h0; adw.y += dv[q].y * h1; adw.z += dv[q].z * h2; adw.w += dv[q].w * h3;
            adb.x += dv[q].x; adb.y += dv[q].y; adb.z += dv[q].z; adb.w += dv[q].w;
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1)
#pragma unroll
            for (int q = 0; q < R; q++) {
                s1[q] += __shfl_xor_sync(0xffffffffu, s1[q], o);
                s2[q] += __shfl_xor_sync(0xffffffffu, s2[q], o);
            }
#pragma unroll
        for (int q = 0; q < R; q++) {
            const long long r = (g0 + q * nwarps) * RPW + sub;
            if (r >= Rows) continue;
            const float m1 = s1[q] / W, m2 = s2[q] / W;
            const float h0 = (xv[q].x - mu[q]) * rstd[q], h1 = (xv[q].y - mu[q]) * rstd[q];
            const float h2 = (xv[q].z - mu[q]) * rstd[q], h3 = (xv[q].w - mu[q]) * rstd[q];
            float4 o;
            o.x = rstd[q] * (dv[q].x * wv.x - m1 - h0 * m2);
            o.y = rstd[q] * (dv[q].y * wv.y - m1 - h1 * m2);
            o.z = rstd[q] * (dv[q].z * wv.z - m1 - h2 * m2);
            o.w = rstd[q] * (dv[q].w * wv.w - m1 - h3 * m2);
            if (relu_in) {
                if (!(xv[q].x > 0.f)) o.x = 0.f;
                if (!(xv[q].y > 0.f)) o.y = 0.f;
                if (!(xv[q].z > 0.f)) o.z = 0.f;
                if (!(xv[q].w > 0.f)) o.w = 0.f;
            }
            *reinterpret_cast<float4*>(dx + r * W + c4) = o;
        }
    }
    // the RPW row slots of a warp hold the same columns: fold them with shuffles before touching shared memory
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) {
        adw.x += __shfl_xor_sync(0xffffffffu, adw.x, o); adw.y += __shfl_xor_sync(0xffffffffu, adw.y, o);
        adw.z += __shfl_xor_sync(0xffffffffu, adw.z, o); adw.w += __shfl_xor_sync(0xffffffffu, adw.w, o);
        adb.x += __shfl_xor_sync(0xffffffffu, adb.x, o); adb.y += __shfl_xor_sync(0xffffffffu, adb.y, o);
        adb.z += __shfl_xor_sync(0xffffffffu, adb.z, o); adb.w += __shfl_xor_sync(0xffffffffu, adb.w, o);
    }
    if (sub == 0) {
        atomicAdd(&sdw[c4 + 0], adw.x); atomicAdd(&sdw[c4 + 1], adw.y); atomicAdd(&sdw[c4 + 2], adw.z); atomicAdd(&sdw[c4 + 3], adw.w);
        atomicAdd(&sdb[c4 + 0], adb.x); atomicAdd(&sdb[c4 + 1], adb.y); atomicAdd(&sdb[c4 + 2], adb.z); atomicAdd(&sdb[c4 + 3], adb.w);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < W; c += blockDim.x) {
        atomicAdd(dw + c, sdw[c]);
        atomicAdd(db + c, sdb[c]);
    }
}

// ---------------------------------------------------------------------------
// CensNet aggregation (censNetConv_pt.py:92-136), one CTA per window:
//   Mv = (T diag(He.we) T^T) o lap ;  Pn = Mv.Hn        Me = (T^T diag(Hn.wn) T) o elap ; Pe = Me.He
// The dense kernels (Pn.node_kernel + bias, relu) go through gemm_rows.
// ---------------------------------------------------------------------------
struct CensArgs {
    const float* node; const float* edge;       // [B,N,C], [B,E,C]
    const float* lap; const float* elap; const float* inc;   // [N,N],[E,E],[N,E]
    const float* wn; const float* we;           // [C]
    float* Pn; float* Pe;                       // fwd outputs [B,N,C],[B,E,C]
    // backward
    const float* dPn; const float* dPe;
    float* dnode; float* dedge;                 // [B,N,C],[B,E,C]
    float* dwn; float* dwe;                     // [C] (atomic accumulate)
    int B, N, E, C;
};

__device__ __forceinline__ void cens_load(const CensArgs& a, int b, float* nd, float* ed, float* sinc,
                                          float* wev, float* wnv, float* Mv, float* Me) {
    const int N = a.N, E = a.E, C = a.C;
    for (int i = threadIdx.x; i < N * C; i += blockDim.x) nd[i] = __ldg(a.node + (size_t)b * N * C + i);
    for (int i = threadIdx.x; i < E * C; i += blockDim.x) ed[i] = __ldg(a.edge + (size_t)b * E * C + i);
    for (int i = threadIdx.x; i < N * E; i += blockDim.x) sinc[i] = __ldg(a.inc + i);
    __syncthreads();
    for (int i = threadIdx.x; i < N + E; i += blockDim.x) {
        float s = 0.f;
        if (i < E) {
            for (int c = 0; c < C; c++) s = fmaf(ed[i * C + c], __ldg(a.we + c), s);
            wev[i] = s;
        } else {
            int n = i - E;
            for (int c = 0; c < C; c++) s = fmaf(nd[n * C + c], __ldg(a.wn + c), s);
            wnv[n] = s;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N * N + E * E; i += blockDim.x) {
        if (i < N * N) {
            int r = i / N, c = i % N;
            float s = 0.f;
            for (int e = 0; e < E; e++) s = fmaf(sinc[r * E + e] * wev[e], sinc[c * E + e], s);
            Mv[i] = s * __ldg(a.lap + i);
        } else {
            int k = i - N * N;
            int e = k / E, f = k % E;
            float s = 0.f;
            for (int n = 0; n < N; n++) s = fmaf(sinc[n * E + e] * wnv[n], sinc[n * E + f], s);
            Me[k] = s * __ldg(a.elap + k);
        }
    }
    __syncthreads();
}

static inline size_t cens_smem_floats(int N, int E, int C) {
    return (size_t)N * C + (size_t)E * C + (size_t)N * E + E + N + (size_t)N * N + (size_t)E * E;
}

__global__ void __launch_bounds__(128) cens_fwd_kernel(const CensArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int N = a.N, E = a.E, C = a.C;
    float* nd = smem; float* ed = nd + N * C; float* sinc = ed + E * C;
    float* wev = sinc + N * E; float* wnv = wev + E; float* Mv = wnv + N; float* Me = Mv + N * N;
    const int b = blockIdx.x;
    cens_load(a, b, nd, ed, sinc, wev, wnv, Mv, Me);
    for (int i = threadIdx.x; i < (N + E) * C; i += blockDim.x) {
        if (i < N * C) {
            int r = i / C, c = i % C;
            float s = 0.f;
            for (int j = 0; j < N; j++) s = fmaf(Mv[r * N + j], nd[j * C + c], s);
            a.Pn[(size_t)b * N * C + i] = s;
        } else {
            int k = i - N * C;
            int r = k / C, c = k % C;
            float s = 0.f;
            for (int j = 0; j < E; j++) s = fmaf(Me[r * E + j], ed[j * C + c], s);
            a.Pe[(size_t)b * E * C + k] = s;
        }
    }
}

__global__ void __launch_bounds__(128) cens_bwd_kernel(const CensArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int N = a.N, E = a.E, C = a.C;
    float* nd = smem; float* ed = nd + N * C; float* sinc = ed + E * C;
    float* wev = sinc + N * E; float* wnv = wev + E; float* Mv = wnv + N; float* Me = Mv + N * N;
    float* dpn = Me + E * E;            // [N*C]
    float* dpe = dpn + N * C;           // [E*C]
    float* G1 = dpe + E * C;            // [N*N]
    float* G2 = G1 + N * N;             // [E*E]
    float* dwev = G2 + E * E;           // [E]
    float* dwnv = dwev + E;             // [N]
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < N * C; i += blockDim.x) dpn[i] = __ldg(a.dPn + (size_t)b * N * C + i);
    for (int i = threadIdx.x; i < E * C; i += blockDim.x) dpe[i] = __ldg(a.dPe + (size_t)b * E * C + i);
    cens_load(a, b, nd, ed, sinc, wev, wnv, Mv, Me);
    for (int i = threadIdx.x; i < N * N + E * E; i += blockDim.x) {
        if (i < N * N) {
            int r = i / N, c = i % N;
            float s = 0.f;
            for (int k = 0; k < C; k++) s = fmaf(dpn[r * C + k], nd[c * C + k], s);
            G1[i] = s * __ldg(a.lap + i);
        } else {
            int k2 = i - N * N;
            int e = k2 / E, f = k2 % E;
            float s = 0.f;
            for (int k = 0; k < C; k++) s = fmaf(dpe[e * C + k], ed[f * C + k], s);
            G2[k2] = s * __ldg(a.elap + k2);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N + E; i += blockDim.x) {
        float s = 0.f;
        if (i < E) {
            for (int r = 0; r < N; r++) {
                float ir = sinc[r * E + i];
                if (ir == 0.f) continue;
                for (int c = 0; c < N; c++) s = fmaf(ir * sinc[c * E + i], G1[r * N + c], s);
            }
            dwev[i] = s;
        } else {
            int n = i - E;
            for (int e = 0; e < E; e++) {
                float ie = sinc[n * E + e];
                if (ie == 0.f) continue;
                for (int f = 0; f < E; f++) s = fmaf(ie * sinc[n * E + f], G2[e * E + f], s);
            }
            dwnv[n] = s;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (N + E) * C; i += blockDim.x) {
        if (i < N * C) {
            int j = i / C, c = i % C;
            float s = dwnv[j] * __ldg(a.wn + c);
            for (int r = 0; r < N; r++) s = fmaf(Mv[r * N + j], dpn[r * C + c], s);
            a.dnode[(size_t)b * N * C + i] = s;
        } else {
            int k = i - N * C;
            int f = k / C, c = k % C;
            float s = dwev[f] * __ldg(a.we + c);
            for (int e = 0; e < E; e++) s = fmaf(Me[e * E + f], dpe[e * C + c], s);
            a.dedge[(size_t)b * E * C + k] = s;
        }
    }
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
        float s = 0.f;
        if (c < C) {
            for (int e = 0; e < E; e++) s = fmaf(dwev[e], ed[e * C + c], s);
            atomicAdd(a.dwe + c, s);
        } else {
            int cc = c - C;
            for (int n = 0; n < N; n++) s = fmaf(dwnv[n], nd[n * C + cc], s);
            atomicAdd(a.dwn + cc, s);
        }
    }
}

// ---------------------------------------------------------------------------
// Gaussian-mixture latent head (models_new.py:1761-1791, 1745-1759), one thread per window.
// ---------------------------------------------------------------------------
struct LatentArgs {
    const float* enc;                      // [B,D]
    const float *Wm, *bm, *Wv, *bv;        // encoder_mean / encoder_log_var Linear
    const float* eps;                      // [B,D] reparam noise; null => eval (z = z_mean) unless noise_seed != 0
    unsigned long long noise_seed;         // != 0 with eps == null: eps comes from the Philox stream (seed, DOF_SITE_EPS)
    const float *gmm_mu, *gmm_lv, *prior;  // [K,D],[K,D],[K]
    float *zm, *pre, *lv, *z, *q;          // [B,D] x4, [B,K]
    int B, D, K;
};

#define LOG_2PI_F 1.8378770664093453f
#define DOF_SITE_EPS 201u      // Philox sites of the VaDE noise: reparameterisation eps [B,D] ...
#define DOF_SITE_MC 200u       // ... and the Monte-Carlo KL samples [S,B,D]

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }

__global__ void __launch_bounds__(64) latent_fwd_kernel(const LatentArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int D = a.D, K = a.K;
    float* Wm = smem; float* Wv = Wm + D * D; float* bm = Wv + D * D; float* bv = bm + D;
    float* mu = bv + D; float* isd = mu + K * D; float* lcst = isd + K * D;   // lcst[K]
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) { Wm[i] = __ldg(a.Wm + i); Wv[i] = __ldg(a.Wv + i); }
    for (int i = threadIdx.x; i < D; i += blockDim.x) { bm[i] = __ldg(a.bm + i); bv[i] = __ldg(a.bv + i); }
    for (int i = threadIdx.x; i < K * D; i += blockDim.x) {
        mu[i] = __ldg(a.gmm_mu + i);
        float sd = fmaxf(expf(0.5f * __ldg(a.gmm_lv + i)), 1e-3f);
        isd[i] = 1.0f / sd;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < K; c += blockDim.x) {
        float s = logf(__ldg(a.prior + c) + 1e-9f);
        for (int d = 0; d < D; d++) s += logf(isd[c * D + d]) - 0.5f * LOG_2PI_F;   // -log(sd) - .5 log 2pi
        lcst[c] = s;
    }
    __syncthreads();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    const float* e = a.enc + (size_t)b * D;
    for (int d = 0; d < D; d++) {
        float s1 = bm[d], s2 = bv[d];
        for (int k = 0; k < D; k++) {
            float ev = e[k];
            s1 = fmaf(Wm[d * D + k], ev, s1);
            s2 = fmaf(Wv[d * D + k], ev, s2);
        }
        float lv = softplus_f(s2);
        float zz = s1;
        if (a.eps) zz = s1 + expf(0.5f * lv) * a.eps[(size_t)b * D + d];
        else if (a.noise_seed) zz = s1 + expf(0.5f * lv) * philox_normal(a.noise_seed, DOF_SITE_EPS, (unsigned long long)b * D + d);
        a.zm[(size_t)b * D + d] = s1;
        a.pre[(size_t)b * D + d] = s2;
        a.lv[(size_t)b * D + d] = lv;
        a.z[(size_t)b * D + d] = zz;
    }
    float mx = -INFINITY;
    float* qb = a.q + (size_t)b * K;
    for (int c = 0; c < K; c++) {
        float s = 0.f;
        for (int d = 0; d < D; d++) {
            float t = (a.z[(size_t)b * D + d] - mu[c * D + d]) * isd[c * D + d];
            s = fmaf(t, t, s);
        }
        float lg = lcst[c] - 0.5f * s;
        qb[c] = lg;
        mx = fmaxf(mx, lg);
    }
    float sum = 0.f;
    for (int c = 0; c < K; c++) { float v = expf(qb[c] - mx); qb[c] = v; sum += v; }
    float inv = 1.0f / sum;
    for (int c = 0; c < K; c++) qb[c] *= inv;
}

// out[b, :] = sum_t in[b, t, :]   (gradient of the decoder's RepeatVector)
__global__ void sum_over_t_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int T, int W,
                                  int accum) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * W) return;
    int b = (int)(i / W), c = (int)(i % W);
    float s = accum ? out[i] : 0.f;
    for (int t = 0; t < T; t++) s += in[((size_t)b * T + t) * W + c];
    out[i] = s;
}

// dst [R, ldd] = src [R, W] with the columns W .. ldd-1 set to zero: a row pitch that is a multiple of 4 floats makes the operand
// eligible for the float4 producers of the tensor-core GEMMs (N * F = 42 -> 44)
__global__ void pad_cols_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd, long long R, int W) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * ldd) return;
    const long long r = i / ldd;
    const int c = (int)(i - r * ldd);
    dst[i] = c < W ? __ldg(src + r * lds + c) : 0.f;
}

// dst[n][c*5+kk] = src[c][n][kk]   (Conv1d weight [Cout=c, Cin=n, 5] -> transposed-conv operand)
__global__ void conv_w_transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int Cout, int Cin) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Cout * Cin * 5) return;
    int kk = i % 5, n = (i / 5) % Cin, c = i / (5 * Cin);
    dst[(size_t)n * Cout * 5 + c * 5 + kk] = src[i];
}

// ---------------------------------------------------------------------------
// clip_grad_value_ + Adam (training.py:164-166, losses.py:817-833; torch.optim.Adam
// defaults beta=(0.9,0.999), eps=1e-8, no weight decay).  group[i]: 0 = not trained
// (buffers, dead parameters), 1.. = parameter group with its own lr / step count.
// ---------------------------------------------------------------------------
struct AdamArgs {
    float* p; const float* g; float* m; float* v; const unsigned char* group;
    long long n;
    float lr[4]; float wd[4]; float bc1[4]; float bc2_sqrt[4]; int active[4];
    float clip; float gscale; float beta1, beta2, eps;
};

__global__ void __launch_bounds__(256) clip_adam_kernel(const AdamArgs a) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    int grp = a.group ? a.group[i] : 1;          // no group map: one flat buffer in group 1
    if (grp == 0 || !a.active[grp]) return;
    float g = a.g[i] * a.gscale;
    if (a.clip > 0.f) g = fminf(fmaxf(g, -a.clip), a.clip);
    g += a.wd[grp] * a.p[i];          // torch.optim.Adam(weight_decay): L2 term added after clip_grad_value_
    float m = a.beta1 * a.m[i] + (1.0f - a.beta1) * g;
    float v = a.beta2 * a.v[i] + (1.0f - a.beta2) * g * g;
    a.m[i] = m;
    a.v[i] = v;
    float denom = sqrtf(v) / a.bc2_sqrt[grp] + a.eps;
    a.p[i] = a.p[i] - (a.lr[grp] / a.bc1[grp]) * (m / denom);
}
