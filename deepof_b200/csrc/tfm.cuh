// Transformer encoder of the path, EVAL-mode forward (reference deepof/clustering/models_new.py):
//   TransformerCorePT.forward :955-982, TransformerEncoderLayerPT :893-919 (post-LN, eps 1e-6), MultiHeadAttentionPT
//   :843-890 (bias-free projections, key-padding mask), sinusoidal_positional_encoding :832-840, TFMEncoderPT.forward
//   :1093-1164 (A.1 group scramble, CensNet, RMS normalisation, head MLP with BatchNorm running statistics).
//
// tfm_core_fwd_kernel: one CTA per (window, node) sequence at a time, persistent over the sequences; ALL weights of the
// core (2 layers: 136 KB at key_dim 40, dff 128) stay in shared memory, transposed ([K][N]: threads of a warp read
// consecutive output columns, 5 x 2 register tiles); the sequence (T x key_dim activations, q|k|v, FFN hidden) lives in shared memory, so HBM sees the
// raw window once and key_dim floats out.  Only the LAST time step leaves the core (:982), so the last layer computes
// queries, attention output, projections and FFN for that single row (keys / values for all rows).  fp32 SIMT: the
// per-sequence products are 25 x 40 x {120, 40, 128}: far below a tcgen05 tile, and this is the inference path
// (embedding_per_video); training runs the composed tcgen05 path in tfm_step.cuh / tfm_train.cuh instead.
#pragma once
#include "common.cuh"
#include "tfm_train.cuh"

#define TFM_MAXT 64
#define TFM_THREADS 512

struct TfmCoreArgs {
    const float* x;        // [B, T, G, F] window tensor (x or a)
    const float* params;   // this core's parameters, reference order (embed.weight first)
    float* out;            // [B*G, dk]
    float* ybuf;           // [B*G, T, dk] activations between launches when the layers do not fit one launch, else NULL
    int B, T, G, F, dk, heads, dff, layers;
    int l_begin, l_end;    // layers [l_begin, l_end) run in this launch (their weights are the ones held in shared memory)
};

__host__ __device__ inline int tfm_layer_floats(int dk, int dff) { return 4 * dk * dk + 2 * dk + dff * dk + dff + dk * dff + dk + 2 * dk; }
__host__ __device__ inline int tfm_core_floats(int F, int dk, int dff, int layers) { return dk * F + dk + layers * tfm_layer_floats(dk, dff); }
// shared-memory copy: weight matrices TRANSPOSED ([K][N], q|k|v side by side as one [dk][3dk] matrix)
__host__ __device__ inline int tfm_layer_smem_floats(int dk, int dff) { return tfm_layer_floats(dk, dff); }
static inline size_t tfm_core_smem_bytes(int T, int F, int dk, int dff, int layers) {      // layers = layers per launch
    size_t fl = (size_t)dk * F + dk + (size_t)layers * tfm_layer_smem_floats(dk, dff)   // weights
              + (size_t)T * dk                       // positional encoding
              + 2 * ((size_t)T * dk * 2              // per group: y, att
                     + (size_t)T * (3 * dk + 1)      // q | k | v, odd pitch (keys of one head: conflict-free across steps)
                     + (size_t)T * dff               // FFN hidden
                     + (size_t)T);                   // key mask
    return fl * 4 + 64;
}

// out[t][n] = sum_k in[t][k] * Wt[k][n] (+ bias) (relu) for t in [t0, T); Wt is the transposed weight with row pitch ldw.
// Register tile: 5 rows x 2 columns per thread (one 8-byte weight load + 5 broadcast activation loads per 10 FMAs).
#define TFM_TT 5
__device__ __forceinline__ void tfm_linear(int ltid, int nthr, const float* in, int ldin, const float* Wt, int ldw, const float* bias,
                                           float* out, int ldout, int t0, int T, int N, int K, bool relu) {
    const int nc = N >> 1, ng = (T - t0 + TFM_TT - 1) / TFM_TT;
    for (int tile = ltid; tile < nc * ng; tile += nthr) {
        const int cn = (tile % nc) * 2, tb = t0 + (tile / nc) * TFM_TT;
        float acc[TFM_TT][2];
        const float bx = bias ? bias[cn] : 0.f, by = bias ? bias[cn + 1] : 0.f;
#pragma unroll
        for (int i = 0; i < TFM_TT; i++) { acc[i][0] = bx; acc[i][1] = by; }
        const float* v[TFM_TT];
#pragma unroll
        for (int i = 0; i < TFM_TT; i++) v[i] = in + (size_t)min(tb + i, T - 1) * ldin;
        const float* w = Wt + cn;
#pragma unroll 4
        for (int k = 0; k < K; k++) {
            const float2 ww = *reinterpret_cast<const float2*>(w + (size_t)k * ldw);
#pragma unroll
            for (int i = 0; i < TFM_TT; i++) { const float x = v[i][k]; acc[i][0] += x * ww.x; acc[i][1] += x * ww.y; }
        }
#pragma unroll
        for (int i = 0; i < TFM_TT; i++) {
            if (tb + i < T) {
                float2 o = make_float2(acc[i][0], acc[i][1]);
                if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); }
                out[(size_t)(tb + i) * ldout + cn] = o.x;            // scalar stores: the q|k|v pitch is odd
                out[(size_t)(tb + i) * ldout + cn + 1] = o.y;
            }
        }
    }
}

// y[t] = LayerNorm(y[t] + r[t]) for t in [t0, T): one warp per row
__device__ __forceinline__ void tfm_add_ln(int ltid, int nthr, float* y, const float* r, int ldr, const float* w, const float* b, int t0,
                                           int T, int dk) {
    const int warp = ltid >> 5, lane = ltid & 31, nw = nthr >> 5;
    for (int t = t0 + warp; t < T; t += nw) {
        float s = 0.f;
        for (int d = lane; d < dk; d += 32) s += y[t * dk + d] + r[(size_t)t * ldr + d];
        const float mu = warp_sum(s) / dk;
        float q = 0.f;
        for (int d = lane; d < dk; d += 32) { const float v = y[t * dk + d] + r[(size_t)t * ldr + d] - mu; q += v * v; }
        const float rs = rsqrtf(warp_sum(q) / dk + 1e-6f);
        for (int d = lane; d < dk; d += 32) y[t * dk + d] = (y[t * dk + d] + r[(size_t)t * ldr + d] - mu) * rs * w[d] + b[d];
    }
}

#define TFM_GROUPS 2                                   // independent 256-thread groups per CTA, one sequence each
__device__ __forceinline__ void tfm_group_sync(int grp) {
    asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(TFM_THREADS / TFM_GROUPS) : "memory");
}

__global__ void __launch_bounds__(TFM_THREADS, 1) tfm_core_fwd_kernel(const TfmCoreArgs a) {
    extern __shared__ __align__(16) float tfsm[];
    const int T = a.T, G = a.G, F = a.F, dk = a.dk, dff = a.dff, heads = a.heads, hd = a.dk / a.heads;
    const int tid = threadIdx.x;
    float* We = tfsm;                    // [dk][F]
    float* be = We + dk * F;              // [dk]
    float* Wl = be + dk;                  // layers x padded layer block
    const int lsm = tfm_layer_smem_floats(dk, dff);
    const int nl = a.l_end - a.l_begin;
    float* pe = Wl + (size_t)nl * lsm;    // [T][dk]
    const int grp = tid / (TFM_THREADS / TFM_GROUPS), ltid = tid % (TFM_THREADS / TFM_GROUPS), nthr = TFM_THREADS / TFM_GROUPS;
    const int ldq = 3 * dk + 1;
    const size_t gfl = (size_t)T * dk * 2 + (size_t)T * ldq + (size_t)T * dff + T;
    float* y = pe + T * dk + grp * gfl;   // [T][dk]      (this group's sequence)
    float* att = y + T * dk;              // [T][dk]
    float* qkv = att + T * dk;            // [T][3dk + 1]
    float* ff = qkv + T * ldq;            // [T][dff]
    float* kmask = ff + T * dff;          // [T] 1 = padded key
    // ---- weights -> shared memory (padded rows), positional encoding
    for (int i = tid; i < dk * F + dk; i += blockDim.x) We[i] = __ldg(a.params + i);
    for (int l = a.l_begin; l < a.l_end; l++) {
        const float* src = a.params + dk * F + dk + (size_t)l * tfm_layer_floats(dk, dff);
        float* dst = Wl + (size_t)(l - a.l_begin) * lsm;
        // q | k | v: three [dk][dk] (out, in) matrices -> one [in = dk][3dk] matrix; out_proj -> [dk][dk] transposed
        for (int i = tid; i < 3 * dk * dk; i += blockDim.x) { const int m = i / (dk * dk), r = (i / dk) % dk, c = i % dk; dst[c * 3 * dk + m * dk + r] = __ldg(src + i); }
        src += 3 * dk * dk; dst += 3 * dk * dk;
        for (int i = tid; i < dk * dk; i += blockDim.x) { const int r = i / dk, c = i % dk; dst[c * dk + r] = __ldg(src + i); }
        src += dk * dk; dst += dk * dk;
        for (int i = tid; i < 2 * dk; i += blockDim.x) dst[i] = __ldg(src + i);                      // norm1 w | b
        src += 2 * dk; dst += 2 * dk;
        for (int i = tid; i < dff * dk; i += blockDim.x) { const int r = i / dk, c = i % dk; dst[c * dff + r] = __ldg(src + i); }   // ffn.0.weight -> [dk][dff]
        src += dff * dk; dst += dff * dk;
        for (int i = tid; i < dff; i += blockDim.x) dst[i] = __ldg(src + i);                         // ffn.0.bias
        src += dff; dst += dff;
        for (int i = tid; i < dk * dff; i += blockDim.x) { const int r = i / dff, c = i % dff; dst[c * dk + r] = __ldg(src + i); }  // ffn.2.weight -> [dff][dk]
        src += dk * dff; dst += dk * dff;
        for (int i = tid; i < 3 * dk; i += blockDim.x) dst[i] = __ldg(src + i);                      // ffn.2.bias | norm2 w | b
    }
    for (int i = tid; i < T * dk; i += blockDim.x) {
        const int t = i / dk, d = i % dk;
        const float div = expf((float)(d & ~1) * (-logf(10000.0f) / (float)dk));                     // :835
        pe[i] = (d & 1) ? cosf((float)t * div) : sinf((float)t * div);
    }
    __syncthreads();
    const float scale = sqrtf((float)dk), qs = rsqrtf((float)hd);
    const int S = a.B * G;
    for (int s = blockIdx.x * TFM_GROUPS + grp; s < S; s += gridDim.x * TFM_GROUPS) {
        const int b = s / G, g = s % G;
        const float* xw = a.x + (size_t)b * T * G * F;
        // ---- A.1 gather + padding mask + embedding: y = relu(x We^T + be) * sqrt(dk) + PE
        for (int i = ltid; i < T * dk; i += nthr) {
            const int t = i / dk, d = i % dk;
            float acc = be[d];
            bool allz = true;
            for (int f = 0; f < F; f++) {
                const int lin = (f * T + t) * G + g;
                const float v = __ldg(xw + (size_t)(lin % T) * G * F + lin / T);
                allz = allz && (v == 0.0f);
                acc += v * We[d * F + f];
            }
            y[i] = a.l_begin == 0 ? fmaxf(acc, 0.f) * scale + pe[i] : a.ybuf[(size_t)s * T * dk + i];
            if (d == 0) kmask[t] = allz ? 1.f : 0.f;
        }
        tfm_group_sync(grp);
        for (int l = a.l_begin; l < a.l_end; l++) {
            const float* P = Wl + (size_t)(l - a.l_begin) * lsm;
            const float* Wq = P;                               // [dk][3dk]: q | k | v columns
            const float* Wo = P + 3 * dk * dk;                 // [dk][dk]
            const float* n1 = Wo + dk * dk;
            const float* W1 = n1 + 2 * dk;                     // [dk][dff]
            const float* b1 = W1 + dff * dk;
            const float* W2 = b1 + dff;                        // [dff][dk]
            const float* b2 = W2 + dk * dff;
            const float* n2 = b2 + dk;
            const int t0 = (l == a.layers - 1) ? T - 1 : 0;    // only the last step leaves the core
            // keys / values for every step, queries from t0 on (q rows < t0 are computed too when t0 == 0 only)
            if (t0 == 0) tfm_linear(ltid, nthr, y, dk, Wq, 3 * dk, nullptr, qkv, ldq, 0, T, 3 * dk, dk, false);    // q | k | v
            else {
                tfm_linear(ltid, nthr, y, dk, Wq + dk, 3 * dk, nullptr, qkv + dk, ldq, 0, T, 2 * dk, dk, false);   // k | v for every step
                tfm_linear(ltid, nthr, y, dk, Wq, 3 * dk, nullptr, qkv, ldq, t0, T, dk, dk, false);                 // q for the last step
            }
            tfm_group_sync(grp);
            // attention: one WARP per (head, query step); lanes own keys (scores, softmax by shuffles), then head dims
            {
                const int warp = ltid >> 5, lane = ltid & 31, nw = nthr >> 5;
                for (int o = warp; o < heads * (T - t0); o += nw) {
                    const int hh = o / (T - t0), tq = t0 + o % (T - t0);
                    const float* q = qkv + (size_t)tq * ldq + hh * hd;
                    float sc[TFM_MAXT / 32];
                    float mx = -INFINITY;
#pragma unroll
                    for (int c = 0; c < TFM_MAXT / 32; c++) {
                        const int tk = lane + 32 * c;
                        float dot = -INFINITY;
                        if (tk < T && kmask[tk] == 0.f) {
                            const float* kk = qkv + (size_t)tk * ldq + dk + hh * hd;
                            dot = 0.f;
                            for (int d = 0; d < hd; d++) dot += q[d] * kk[d];
                            dot *= qs;
                        }
                        sc[c] = dot;
                        mx = fmaxf(mx, dot);
                    }
                    mx = warp_max(mx);
                    float sum = 0.f;
#pragma unroll
                    for (int c = 0; c < TFM_MAXT / 32; c++) {
                        const int tk = lane + 32 * c;
                        sc[c] = tk < T ? expf(sc[c] - mx) : 0.f;          // all keys masked: exp(-inf + inf) = NaN like the reference
                        sum += sc[c];
                    }
                    sum = warp_sum(sum);
                    const float inv = 1.0f / sum;
                    for (int d0 = 0; d0 < hd; d0 += 32) {
                        const int d = d0 + lane;
                        float acc = 0.f;
#pragma unroll
                        for (int c = 0; c < TFM_MAXT / 32; c++) {
                            for (int j = 0; j < 32 && j + 32 * c < T; j++) {
                                const float pj = __shfl_sync(0xffffffffu, sc[c], j);
                                if (d < hd) acc += pj * qkv[(size_t)(j + 32 * c) * ldq + 2 * dk + hh * hd + d];
                            }
                        }
                        if (d < hd) att[(size_t)tq * dk + hh * hd + d] = acc * inv;
                    }
                }
            }
            tfm_group_sync(grp);
            tfm_linear(ltid, nthr, att, dk, Wo, dk, nullptr, qkv, ldq, t0, T, dk, dk, false);          // out projection -> qkv[:, 0:dk] (q is dead)
            tfm_group_sync(grp);
            tfm_add_ln(ltid, nthr, y, qkv, ldq, n1, n1 + dk, t0, T, dk);        // y = LN(y + out_proj(attention))
            tfm_group_sync(grp);
            tfm_linear(ltid, nthr, y, dk, W1, dff, b1, ff, dff, t0, T, dff, dk, true);
            tfm_group_sync(grp);
            tfm_linear(ltid, nthr, ff, dff, W2, dk, b2, att, dk, t0, T, dk, dff, false);
            tfm_group_sync(grp);
            tfm_add_ln(ltid, nthr, y, att, dk, n2, n2 + dk, t0, T, dk);
            tfm_group_sync(grp);
        }
        if (a.l_end == a.layers) {
            for (int d = ltid; d < dk; d += nthr) a.out[(size_t)s * dk + d] = y[(size_t)(T - 1) * dk + d];
        } else {
            for (int i = ltid; i < T * dk; i += nthr) a.ybuf[(size_t)s * T * dk + i] = y[i];
        }
        tfm_group_sync(grp);
    }
}

// ---- head: RMS normalisation + Linear/ReLU/BN(eval) x2 + Linear; one warp per window --------------------------
struct TfmHeadArgs {
    const float* on; const float* oe;      // CensNet outputs [B, N*D], [B, E*D] (already >= 0)
    const float* w0; const float* b0; const float* bn2;   // bn: weight | bias | running_mean | running_var
    const float* w3; const float* b3; const float* bn5;
    const float* w6; const float* b6;
    float* out;                            // [B, D]
    int B, ND, ED, D;
};

__global__ void __launch_bounds__(128) tfm_head_fwd_kernel(const TfmHeadArgs a) {
    extern __shared__ __align__(16) float hsm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * 4 + warp;
    if (b >= a.B) return;
    const int KD = a.ND + a.ED, D = a.D;
    float* v = hsm + (size_t)warp * (KD + 3 * D);      // input | h1 [2D] | h2 [D]
    float* h1 = v + KD;
    float* h2 = h1 + 2 * D;
    float ss = 0.f;
    for (int k = lane; k < KD; k += 32) {
        const float x = k < a.ND ? a.on[(size_t)b * a.ND + k] : a.oe[(size_t)b * a.ED + (k - a.ND)];
        v[k] = x;
        ss += x * x;
    }
    const float rms = sqrtf(warp_sum(ss) / KD);
    const float inv = 1.0f / fmaxf(rms, 1.0f);                                         // :1150-1151
    for (int k = lane; k < KD; k += 32) {
        float x = fminf(fmaxf(v[k] * inv, -1e4f), 1e4f);                               // :1152
        if (x != x) x = 0.f;                                                           // nan_to_num :1153
        v[k] = x;
    }
    __syncwarp();
    for (int n = 0; n < 2 * D; n++) {
        float acc = 0.f;
        for (int k = lane; k < KD; k += 32) acc += v[k] * __ldg(a.w0 + (size_t)n * KD + k);
        acc = warp_sum(acc);
        if (lane == 0) {
            const float r = fmaxf(acc + a.b0[n], 0.f);
            h1[n] = (r - a.bn2[2 * 2 * D + n]) * rsqrtf(a.bn2[3 * 2 * D + n] + 1e-3f) * a.bn2[n] + a.bn2[2 * D + n];
        }
    }
    __syncwarp();
    for (int n = lane; n < D; n += 32) {
        float acc = a.b3[n];
        for (int k = 0; k < 2 * D; k++) acc += h1[k] * a.w3[(size_t)n * 2 * D + k];
        const float r = fmaxf(acc, 0.f);
        h2[n] = (r - a.bn5[2 * D + n]) * rsqrtf(a.bn5[3 * D + n] + 1e-3f) * a.bn5[n] + a.bn5[D + n];
    }
    __syncwarp();
    for (int n = lane; n < D; n += 32) {
        float acc = a.b6[n];
        for (int k = 0; k < D; k++) acc += h2[k] * a.w6[(size_t)n * D + k];
        a.out[(size_t)b * D + n] = acc;
    }
}


// ---------------------------------------------------------------------------
// transformer decoder (row a13), eval forward, from library GEMM / LayerNorm kernels plus three small kernels
// (TFMDecoderPT.forward models_new.py:1232-1266, CausalSelfAttentionLayer :1270-1327)
// ---------------------------------------------------------------------------
__global__ void gelu_kernel(float* __restrict__ x, long long n) {          // nn.GELU() default: exact erf form
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const float v = x[i]; x[i] = 0.5f * v * (1.0f + erff(v * 0.70710678118654752f)); }
}

// h[b, t, :] = g[b, :] + PE[t, :]  (:1247-1255)
__global__ void tfm_dec_input_kernel(const float* __restrict__ g, float* __restrict__ h, int B, int T, int dm) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * T * dm) return;
    const int d = (int)(i % dm), t = (int)((i / dm) % T), b = (int)(i / ((long long)dm * T));
    const float div = expf((float)(d & ~1) * (-logf(10000.0f) / (float)dm));
    h[i] = g[(size_t)b * dm + d] + ((d & 1) ? cosf((float)t * div) : sinf((float)t * div));
}

// softmax(q k^T / sqrt(hd) + causal mask) v for one sequence per CTA; qkv rows are q | k | v (3 dm floats)
__global__ void __launch_bounds__(128) tfm_causal_attn_kernel(const float* __restrict__ qkv, float* __restrict__ out, int T, int dm, int heads) {
    extern __shared__ __align__(16) float asm_[];
    const int b = blockIdx.x, hd = dm / heads;
    const float* src = qkv + (size_t)b * T * 3 * dm;
    for (int i = threadIdx.x; i < T * 3 * dm; i += blockDim.x) asm_[i] = src[i];
    __syncthreads();
    const float qs = rsqrtf((float)hd);
    for (int o = threadIdx.x; o < heads * T; o += blockDim.x) {
        const int hh = o / T, tq = o % T;
        const float* q = asm_ + (size_t)tq * 3 * dm + hh * hd;
        float sc[TFM_MAXT];
        float mx = -INFINITY;
        for (int tk = 0; tk <= tq; tk++) {
            const float* kk = asm_ + (size_t)tk * 3 * dm + dm + hh * hd;
            float dot = 0.f;
            for (int d = 0; d < hd; d++) dot += q[d] * kk[d];
            sc[tk] = dot * qs;
            mx = fmaxf(mx, sc[tk]);
        }
        float sum = 0.f;
        for (int tk = 0; tk <= tq; tk++) { sc[tk] = expf(sc[tk] - mx); sum += sc[tk]; }
        const float inv = 1.0f / sum;
        for (int d = 0; d < hd; d++) {
            float acc = 0.f;
            for (int tk = 0; tk <= tq; tk++) acc += sc[tk] * asm_[(size_t)tk * 3 * dm + 2 * dm + hh * hd + d];
            out[((size_t)b * T + tq) * dm + hh * hd + d] = acc * inv;
        }
    }
}

// ---------------------------------------------------------------------------
// Building blocks of the transformer TRAINING step (rows a12 / a13; not wired into a product path yet, op-level
// parity-tested against autograd): multi-head attention over one sequence per CTA with an optional key-padding mask,
// optional causal mask and inverted dropout on the attention weights (scaled_dot_product_attention with dropout_p,
// models_new.py:880-884, 1311-1315).  The dropout decisions are an INPUT (keep [S,H,T,T] bytes, 1 = kept) so that the
// forward and the backward agree and tests can inject the reference's masks; nothing but q|k|v is saved: the
// backward recomputes the probabilities.
//   forward : P = softmax(q k^T / sqrt(hd) + masks);  Pd = P * keep / (1 - p);  out = Pd v
//   backward: dV = Pd^T dOut;  dPd = dOut v^T;  dP = dPd * keep / (1 - p);  dS = P * (dP - rowsum(dP * P));
//             dQ = dS k / sqrt(hd);  dK = dS^T q / sqrt(hd)
// ---------------------------------------------------------------------------
struct TfmAttnArgs {
    const float* qkv;            // [S, T, 3 dm]  rows q | k | v
    const unsigned char* kpad;   // [S, T] 1 = padded key, or null
    DropSite drop;               // dropout on the attention weights, element index ((s * H + h) * T + tq) * T + tk
    float* out;                  // forward: [S, T, dm]   (q_from > 0: [S, T - q_from, dm], rows tq - q_from)
    const float* dout;           // backward: same shape as out
    float* dqkv;                 // backward: [S, T, 3 dm]
    int S, T, dm, heads, causal;
    int q_from;                  // only queries tq >= q_from are computed (the last layer of the encoder core needs tq = T - 1)
};

template <bool BWD>
__global__ void __launch_bounds__(128) tfm_attn_train_kernel(const TfmAttnArgs a) {
    extern __shared__ __align__(16) float atsm[];
    const int T = a.T, dm = a.dm, heads = a.heads, hd = dm / heads, s = blockIdx.x;
    const int ldq = 3 * dm + 1;
    float* sq = atsm;                          // [T][3dm + 1]
    float* sdo = sq + (size_t)T * ldq;         // BWD: dOut [T][dm + 1]
    float* sdq = sdo + (BWD ? (size_t)T * (dm + 1) : 0);   // BWD: dqkv accumulator [T][3dm + 1]
    const float* src = a.qkv + (size_t)s * T * 3 * dm;
    // a thread owns columns of the [T, 3 dm] block: no per-element division, consecutive threads read consecutive floats of a step
    for (int c = threadIdx.x; c < 3 * dm; c += blockDim.x) {
#pragma unroll 5
        for (int t = 0; t < T; t++) sq[t * ldq + c] = __ldg(src + (size_t)t * 3 * dm + c);
    }
    if (BWD) {
        const int TQl = T - a.q_from;
        const float* dsrc = a.dout + (size_t)s * TQl * dm;
        for (int c = threadIdx.x; c < dm; c += blockDim.x)
            for (int t = 0; t < TQl; t++) sdo[(a.q_from + t) * (dm + 1) + c] = __ldg(dsrc + (size_t)t * dm + c);
        for (int i = threadIdx.x; i < T * ldq; i += blockDim.x) sdq[i] = 0.f;
    }
    __syncthreads();
    const float qs = rsqrtf((float)hd);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int q0 = a.q_from, TQ = T - q0;
    // one warp per (head, query step); lanes own keys
    for (int o = warp; o < heads * TQ; o += nw) {
        const int hh = o / TQ, tq = q0 + o % TQ;
        const float* q = sq + (size_t)tq * ldq + hh * hd;
        float p[TFM_MAXT / 32], pd[TFM_MAXT / 32];
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < TFM_MAXT / 32; c++) {
            const int tk = lane + 32 * c;
            float dot = -INFINITY;
            const bool vis = tk < T && !(a.causal && tk > tq) && !(a.kpad && a.kpad[(size_t)s * T + tk]);
            if (vis) {
                const float* kk = sq + (size_t)tk * ldq + dm + hh * hd;
                dot = 0.f;
                for (int d = 0; d < hd; d++) dot += q[d] * kk[d];
                dot *= qs;
            }
            p[c] = dot;
            mx = fmaxf(mx, dot);
        }
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < TFM_MAXT / 32; c++) {
            const int tk = lane + 32 * c;
            p[c] = tk < T ? expf(p[c] - mx) : 0.f;
            sum += p[c];
        }
        sum = warp_sum(sum);
#pragma unroll
        for (int c = 0; c < TFM_MAXT / 32; c++) {
            const int tk = lane + 32 * c;
            p[c] /= sum;
            float kp = 1.f;
            if (tk < T) kp = drop_mul1(a.drop, (((unsigned long long)s * heads + hh) * T + tq) * T + tk);
            pd[c] = p[c] * kp;                 // dropped-out, rescaled weights
        }
        if (!BWD) {
            for (int d0 = 0; d0 < hd; d0 += 32) {
                const int d = d0 + lane;
                float acc = 0.f;
#pragma unroll
                for (int c = 0; c < TFM_MAXT / 32; c++)
                    for (int j = 0; j < 32 && j + 32 * c < T; j++) {
                        const float pj = __shfl_sync(0xffffffffu, pd[c], j);
                        if (d < hd) acc += pj * sq[(size_t)(j + 32 * c) * ldq + 2 * dm + hh * hd + d];
                    }
                if (d < hd) a.out[((size_t)s * TQ + (tq - q0)) * dm + hh * hd + d] = acc;
            }
        } else {
            const float* dor = sdo + (size_t)tq * (dm + 1) + hh * hd;
            // dPd[tk] = dOut[tq] . v[tk];  dP = dPd * keep / (1 - p);  dS = P (dP - sum_k dP_k P_k)
            float dp[TFM_MAXT / 32];
            float dot_pp = 0.f;
#pragma unroll
            for (int c = 0; c < TFM_MAXT / 32; c++) {
                const int tk = lane + 32 * c;
                float v = 0.f;
                if (tk < T) {
                    const float* vv = sq + (size_t)tk * ldq + 2 * dm + hh * hd;
                    for (int d = 0; d < hd; d++) v += dor[d] * vv[d];
                    // dV[tk] += Pd[tq, tk] * dOut[tq]   (different warps share tk: shared-memory atomics)
                    for (int d = 0; d < hd; d++) atomicAdd(sdq + (size_t)tk * ldq + 2 * dm + hh * hd + d, pd[c] * dor[d]);
                    v *= drop_mul1(a.drop, (((unsigned long long)s * heads + hh) * T + tq) * T + tk);
                }
                dp[c] = v;
                dot_pp += v * p[c];
            }
            dot_pp = warp_sum(dot_pp);
#pragma unroll
            for (int c = 0; c < TFM_MAXT / 32; c++) {
                const int tk = lane + 32 * c;
                const float ds = (tk < T) ? p[c] * (dp[c] - dot_pp) * qs : 0.f;
                dp[c] = ds;
                if (tk < T && ds != 0.f) {
                    // dK[tk] += dS[tq, tk] * q[tq]
                    for (int d = 0; d < hd; d++) atomicAdd(sdq + (size_t)tk * ldq + dm + hh * hd + d, ds * q[d]);
                }
            }
            // dQ[tq] = sum_tk dS[tq, tk] * k[tk]: lanes over head dims, keys by shuffle
            for (int d0 = 0; d0 < hd; d0 += 32) {
                const int d = d0 + lane;
                float acc = 0.f;
#pragma unroll
                for (int c = 0; c < TFM_MAXT / 32; c++)
                    for (int j = 0; j < 32 && j + 32 * c < T; j++) {
                        const float dsj = __shfl_sync(0xffffffffu, dp[c], j);
                        if (d < hd) acc += dsj * sq[(size_t)(j + 32 * c) * ldq + dm + hh * hd + d];
                    }
                if (d < hd) sdq[(size_t)tq * ldq + hh * hd + d] = acc;      // each (head, tq) is owned by one warp
            }
        }
    }
    if (BWD) {
        __syncthreads();
        float* dst = a.dqkv + (size_t)s * T * 3 * dm;
        for (int c = threadIdx.x; c < 3 * dm; c += blockDim.x) {
#pragma unroll 5
            for (int t = 0; t < T; t++) dst[(size_t)t * 3 * dm + c] = sdq[t * ldq + c];
        }
    }
}

// ---------------------------------------------------------------------------
// Attention of the training step, second generation: one CTA per sequence, one WARP per head, one LANE per query step
// (T <= 32).  q | k | v are staged head-major with the head dimension padded to HDP floats, so a lane walks the keys
// with broadcast LDS.128 loads and keeps its whole row (scores, probabilities, output accumulator) in registers: no
// shuffles, no shared-memory atomics.  Backward: phase A (lane = query) recomputes the probabilities, forms dS and dQ and
// parks Pd / dS in shared memory; phase B (lane = key) reduces dV = Pd^T dO and dK = dS^T q over the queries.
// Dropout: explicit keep masks use the reference's dense [S,H,T,T] element order; the Philox stream numbers a row's
// keys from a 32-aligned base (element (row, tk) -> row * 32 + tk), one counter block per 4 keys.
// ---------------------------------------------------------------------------
#define TFM2_MAXT 32

__device__ __forceinline__ uint32_t attn_keep_bits(const DropSite& d, unsigned long long row, int T) {
    if (d.rate <= 0.f) return 0xffffffffu;
    uint32_t bits = 0;
    if (d.keep) {
        const unsigned char* k = d.keep + row * T;
        for (int tk = 0; tk < T; tk++) bits |= (k[tk] ? 1u : 0u) << tk;
        return bits;
    }
    const uint32_t thr = (uint32_t)((double)d.rate * 4294967296.0);
    const uint2 key = make_uint2((uint32_t)d.seed, (uint32_t)(d.seed >> 32));
    for (int b4 = 0; b4 * 4 < T; b4++) {
        const unsigned long long blk = row * 8 + b4;
        const uint4 r = philox4x32_10(make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), d.site, 1u), key);
        bits |= ((r.x >= thr ? 1u : 0u) | (r.y >= thr ? 2u : 0u) | (r.z >= thr ? 4u : 0u) | (r.w >= thr ? 8u : 0u)) << (4 * b4);
    }
    return bits;
}

template <int HDP, bool BWD>
__global__ void __launch_bounds__(256) tfm_attn2_kernel(const TfmAttnArgs a) {
    extern __shared__ __align__(16) float a2sm[];
    const int T = a.T, dm = a.dm, H = a.heads, hd = dm / H, s = blockIdx.x;
    const int nthr = blockDim.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = nthr >> 5;
    const int HT = H * T;
    float* sQ = a2sm;                       // [H][T][HDP]
    float* sK = sQ + (size_t)HT * HDP;
    float* sV = sK + (size_t)HT * HDP;
    float* sdO = sV + (size_t)HT * HDP;     // BWD
    float* sP = sdO + (BWD ? (size_t)HT * HDP : 0);      // BWD: [H][T][T + 1]  dropped probabilities
    float* sdS = sP + (BWD ? (size_t)HT * (T + 1) : 0);  // BWD: [H][T][T + 1]
    const int nz = (BWD ? 4 : 3) * HT * HDP;
    if (hd < HDP) { for (int i = tid; i < nz; i += nthr) a2sm[i] = 0.f; __syncthreads(); }
    // staging: a thread owns COLUMNS of the [T, 3 dm] block (its head / offset decomposition is computed once, not per element —
    // the per-element version spent five integer divisions on every one of the 3000 floats of a cfg3 sequence); for a fixed step
    // consecutive threads read consecutive floats
    const float* src = a.qkv + (size_t)s * T * 3 * dm;
    for (int c = tid; c < 3 * dm; c += nthr) {
        const int m = c / dm, cc = c - m * dm, h = cc / hd, d = cc - h * hd;
        float* dst = (m == 0 ? sQ : m == 1 ? sK : sV) + (size_t)h * T * HDP + d;
        const float* sp = src + c;
#pragma unroll 5
        for (int t = 0; t < T; t++) dst[t * HDP] = __ldg(sp + (size_t)t * 3 * dm);
    }
    if (BWD) {
        const float* dsrc = a.dout + (size_t)s * T * dm;
        for (int c = tid; c < dm; c += nthr) {
            const int h = c / hd, d = c - h * hd;
            float* dst = sdO + (size_t)h * T * HDP + d;
            const float* sp = dsrc + c;
#pragma unroll 5
            for (int t = 0; t < T; t++) dst[t * HDP] = __ldg(sp + (size_t)t * dm);
        }
    }
    __syncthreads();
    const float qs = rsqrtf((float)hd);
    const float inv_keep = a.drop.rate > 0.f ? 1.0f / (1.0f - a.drop.rate) : 1.0f;
    uint32_t padbits = 0;                                  // bit tk = key tk is padded
    if (a.kpad) for (int tk = 0; tk < T; tk++) padbits |= (a.kpad[(size_t)s * T + tk] ? 1u : 0u) << tk;
    for (int h = warp; h < H; h += nw) {
        const int tq = lane;
        const bool live = tq < T;
        const int tqc = live ? tq : T - 1;
        const float* qrow = sQ + ((size_t)h * T + tqc) * HDP;
        float q[HDP];
#pragma unroll
        for (int d = 0; d < HDP; d += 4) { const float4 v = *reinterpret_cast<const float4*>(qrow + d); q[d] = v.x; q[d + 1] = v.y; q[d + 2] = v.z; q[d + 3] = v.w; }
        float p[TFM2_MAXT];
        float mx = -INFINITY;
        uint32_t vis = ~padbits;
        if (a.causal) vis &= (tqc >= 31 ? 0xffffffffu : ((2u << tqc) - 1u));
#pragma unroll
        for (int tk = 0; tk < TFM2_MAXT; tk++) {
            p[tk] = -INFINITY;
            if (tk < T) {
                const float* kr = sK + ((size_t)h * T + tk) * HDP;
                float dot = 0.f;
#pragma unroll
                for (int d = 0; d < HDP; d += 4) {
                    const float4 kv = *reinterpret_cast<const float4*>(kr + d);
                    dot = fmaf(q[d], kv.x, dot); dot = fmaf(q[d + 1], kv.y, dot); dot = fmaf(q[d + 2], kv.z, dot); dot = fmaf(q[d + 3], kv.w, dot);
                }
                if ((vis >> tk) & 1u) { p[tk] = dot * qs; mx = fmaxf(mx, p[tk]); }
            }
        }
        float sum = 0.f;
#pragma unroll
        for (int tk = 0; tk < TFM2_MAXT; tk++) {
            if (tk < T) { p[tk] = expf(p[tk] - mx); sum += p[tk]; } else p[tk] = 0.f;     // all keys masked: NaN like the reference
        }
        const float isum = 1.0f / sum;
        const uint32_t keep = attn_keep_bits(a.drop, ((unsigned long long)s * H + h) * T + tqc, T);
        if (!BWD) {
            float acc[HDP];
#pragma unroll
            for (int d = 0; d < HDP; d++) acc[d] = 0.f;
#pragma unroll
            for (int tk = 0; tk < TFM2_MAXT; tk++) {
                if (tk < T) {
                    const float w = ((keep >> tk) & 1u) ? p[tk] * isum * inv_keep : 0.f;
                    const float* vr = sV + ((size_t)h * T + tk) * HDP;
#pragma unroll
                    for (int d = 0; d < HDP; d += 4) {
                        const float4 vv = *reinterpret_cast<const float4*>(vr + d);
                        acc[d] = fmaf(w, vv.x, acc[d]); acc[d + 1] = fmaf(w, vv.y, acc[d + 1]); acc[d + 2] = fmaf(w, vv.z, acc[d + 2]); acc[d + 3] = fmaf(w, vv.w, acc[d + 3]);
                    }
                }
            }
            if (live) {
                float* o = a.out + ((size_t)s * T + tq) * dm + h * hd;
#pragma unroll
                for (int d = 0; d < HDP; d++) if (d < hd) o[d] = acc[d];
            }
        } else {
            const float* dor = sdO + ((size_t)h * T + tqc) * HDP;
            float dO[HDP];
#pragma unroll
            for (int d = 0; d < HDP; d += 4) { const float4 v = *reinterpret_cast<const float4*>(dor + d); dO[d] = v.x; dO[d + 1] = v.y; dO[d + 2] = v.z; dO[d + 3] = v.w; }
            float dp[TFM2_MAXT];
            float delta = 0.f;
#pragma unroll
            for (int tk = 0; tk < TFM2_MAXT; tk++) {
                dp[tk] = 0.f;
                if (tk < T) {
                    p[tk] *= isum;
                    const float* vr = sV + ((size_t)h * T + tk) * HDP;
                    float dot = 0.f;
#pragma unroll
                    for (int d = 0; d < HDP; d += 4) {
                        const float4 vv = *reinterpret_cast<const float4*>(vr + d);
                        dot = fmaf(dO[d], vv.x, dot); dot = fmaf(dO[d + 1], vv.y, dot); dot = fmaf(dO[d + 2], vv.z, dot); dot = fmaf(dO[d + 3], vv.w, dot);
                    }
                    const float m = ((keep >> tk) & 1u) ? inv_keep : 0.f;
                    dp[tk] = dot * m;
                    delta = fmaf(dp[tk], p[tk], delta);
                    if (live) sP[((size_t)h * T + tq) * (T + 1) + tk] = p[tk] * m;
                }
            }
            float dq[HDP];
#pragma unroll
            for (int d = 0; d < HDP; d++) dq[d] = 0.f;
#pragma unroll
            for (int tk = 0; tk < TFM2_MAXT; tk++) {
                if (tk < T) {
                    const float ds = p[tk] * (dp[tk] - delta) * qs;
                    if (live) sdS[((size_t)h * T + tq) * (T + 1) + tk] = ds;
                    const float* kr = sK + ((size_t)h * T + tk) * HDP;
#pragma unroll
                    for (int d = 0; d < HDP; d += 4) {
                        const float4 kv = *reinterpret_cast<const float4*>(kr + d);
                        dq[d] = fmaf(ds, kv.x, dq[d]); dq[d + 1] = fmaf(ds, kv.y, dq[d + 1]); dq[d + 2] = fmaf(ds, kv.z, dq[d + 2]); dq[d + 3] = fmaf(ds, kv.w, dq[d + 3]);
                    }
                }
            }
            float* go = a.dqkv + ((size_t)s * T + tqc) * 3 * dm + h * hd;
            if (live) {
#pragma unroll
                for (int d = 0; d < HDP; d++) if (d < hd) go[d] = dq[d];
            }
            __syncwarp();
            // phase B: lane = key
            float dk[HDP], dv[HDP];
#pragma unroll
            for (int d = 0; d < HDP; d++) dk[d] = dv[d] = 0.f;
            for (int t2 = 0; t2 < T; t2++) {
                const float pd = sP[((size_t)h * T + t2) * (T + 1) + tqc], ds = sdS[((size_t)h * T + t2) * (T + 1) + tqc];
                const float* d2 = sdO + ((size_t)h * T + t2) * HDP;
                const float* q2 = sQ + ((size_t)h * T + t2) * HDP;
#pragma unroll
                for (int d = 0; d < HDP; d += 4) {
                    const float4 ov = *reinterpret_cast<const float4*>(d2 + d), qv = *reinterpret_cast<const float4*>(q2 + d);
                    dv[d] = fmaf(pd, ov.x, dv[d]); dv[d + 1] = fmaf(pd, ov.y, dv[d + 1]); dv[d + 2] = fmaf(pd, ov.z, dv[d + 2]); dv[d + 3] = fmaf(pd, ov.w, dv[d + 3]);
                    dk[d] = fmaf(ds, qv.x, dk[d]); dk[d + 1] = fmaf(ds, qv.y, dk[d + 1]); dk[d + 2] = fmaf(ds, qv.z, dk[d + 2]); dk[d + 3] = fmaf(ds, qv.w, dk[d + 3]);
                }
            }
            if (live) {
#pragma unroll
                for (int d = 0; d < HDP; d++) if (d < hd) { go[dm + d] = dk[d]; go[2 * dm + d] = dv[d]; }
            }
            __syncwarp();
        }
    }
}

static inline size_t tfm_attn2_smem(int T, int heads, int hdp, bool bwd) {
    return ((size_t)(bwd ? 4 : 3) * heads * T * hdp + (bwd ? (size_t)2 * heads * T * (T + 1) : 0)) * 4;
}
