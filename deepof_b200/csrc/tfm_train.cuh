// Kernels of the transformer-family TRAINING step that are not GEMMs:
//   dropout decisions (explicit keep masks for parity runs, counter-based Philox4x32-10 otherwise),
//   embedding + positional encoding (TransformerCorePT.forward, deepof/clustering/models_new.py:955-982),
//   residual + dropout + LayerNorm forward / backward as sub-warp shuffle reductions (TransformerEncoderLayerPT
//   :893-919 post-LN; CausalSelfAttentionLayer :1270-1327 pre-LN), exact-erf GELU, and the encoder head of
//   TFMEncoderPT.forward (:1146-1164): RMS normalisation, train-mode BatchNorm1d (batch statistics, eps 1e-3,
//   momentum 0.01), batch standardisation with the unbiased std clamped at 0.1.
// All dense products of the step go through the tcgen05 GEMM kernels of tc_gemm.cuh; attention is tfm.cuh.
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------
// dropout
// ---------------------------------------------------------------------------
struct DropSite {
    const unsigned char* keep;   // explicit keep mask (1 = kept) indexed by element, or null -> Philox
    unsigned long long seed;     // Philox key
    unsigned int site;           // Philox counter word 2: one value per dropout site of the step
    float rate;                  // 0: no dropout (eval)
};

static inline DropSite drop_none() { DropSite d; d.keep = nullptr; d.seed = 0; d.site = 0; d.rate = 0.f; return d; }

// multipliers (0 or 1/(1-rate)) of the 4 consecutive elements idx4 .. idx4+3 (idx4 % 4 == 0)
__device__ __forceinline__ void drop_mul4(const DropSite& d, unsigned long long idx4, float (&m)[4]) {
    if (d.rate <= 0.f) { m[0] = m[1] = m[2] = m[3] = 1.f; return; }
    const float inv = 1.0f / (1.0f - d.rate);
    if (d.keep) {
        const uchar4 k = *reinterpret_cast<const uchar4*>(d.keep + idx4);
        m[0] = k.x ? inv : 0.f; m[1] = k.y ? inv : 0.f; m[2] = k.z ? inv : 0.f; m[3] = k.w ? inv : 0.f;
        return;
    }
    const uint32_t thr = (uint32_t)((double)d.rate * 4294967296.0);
    const uint4 r = philox4x32_10(make_uint4((uint32_t)(idx4 >> 2), (uint32_t)(idx4 >> 34), d.site, 0u),
                                  make_uint2((uint32_t)d.seed, (uint32_t)(d.seed >> 32)));
    m[0] = r.x >= thr ? inv : 0.f; m[1] = r.y >= thr ? inv : 0.f; m[2] = r.z >= thr ? inv : 0.f; m[3] = r.w >= thr ? inv : 0.f;
}
__device__ __forceinline__ float drop_mul1(const DropSite& d, unsigned long long idx) {
    if (d.rate <= 0.f) return 1.f;
    const float inv = 1.0f / (1.0f - d.rate);
    if (d.keep) return d.keep[idx] ? inv : 0.f;
    const uint32_t thr = (uint32_t)((double)d.rate * 4294967296.0);
    const uint4 r = philox4x32_10(make_uint4((uint32_t)(idx >> 2), (uint32_t)(idx >> 34), d.site, 0u),
                                  make_uint2((uint32_t)d.seed, (uint32_t)(d.seed >> 32)));
    const uint32_t w = (idx & 3) == 0 ? r.x : (idx & 3) == 1 ? r.y : (idx & 3) == 2 ? r.z : r.w;
    return w >= thr ? inv : 0.f;
}

// ---------------------------------------------------------------------------
// sinusoidal positional encoding table [T, dm] (models_new.py:832-840)
// ---------------------------------------------------------------------------
__global__ void tfm_pe_kernel(float* __restrict__ pe, int T, int dm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T * dm) return;
    const int t = i / dm, d = i % dm;
    const float div = expf((float)(d & ~1) * (-logf(10000.0f) / (float)dm));
    pe[i] = (d & 1) ? cosf((float)t * div) : sinf((float)t * div);
}

// ---------------------------------------------------------------------------
// embedding: A.1 gather -> Xs [R, F], key-padding flag, y0 = drop(relu(xs We^T + be) * sqrt(dk) + PE)   (:966-973)
// one thread per (row, 4 columns)
// ---------------------------------------------------------------------------
struct TfmEmbedArgs {
    const float* x;        // [B, T, G, F] windows
    const int* gidx;       // [G, T, F] offsets inside a window (A.1 scramble)
    const float* We; const float* be;   // [dk, F], [dk]
    const float* pe;       // [T, dk]
    float* Xs;             // [R, F]
    unsigned char* kpad;   // [R] 1 = all-zero step
    float* Y0;             // [R, dk]
    const float* dY0;      // backward
    float* dWe; float* dbe;
    DropSite drop;
    int B, T, G, F, dk;
};

__global__ void __launch_bounds__(256) tfm_embed_fwd_kernel(const TfmEmbedArgs a) {
    const int cg = a.dk >> 2;
    const long long total = (long long)a.B * a.G * a.T * cg;
    const float scale = sqrtf((float)a.dk);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / cg;
        const int c4 = (int)(i % cg) * 4;
        const int t = (int)(row % a.T);
        const long long s = row / a.T;
        const int g = (int)(s % a.G);
        const long long b = s / a.G;
        const float* xw = a.x + b * (long long)a.T * a.G * a.F;
        const int* gi = a.gidx + ((size_t)g * a.T + t) * a.F;
        float xs[4];
        bool allz = true;
        for (int f = 0; f < a.F; f++) {
            const float v = __ldg(xw + gi[f]);
            if (f < 4) xs[f] = v;
            allz = allz && (v == 0.0f);
        }
        float m[4];
        drop_mul4(a.drop, (unsigned long long)row * a.dk + c4, m);
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float acc = __ldg(a.be + c4 + j);
            for (int f = 0; f < a.F; f++) acc = fmaf(f < 4 ? xs[f] : __ldg(xw + gi[f]), __ldg(a.We + (size_t)(c4 + j) * a.F + f), acc);
            o[j] = (fmaxf(acc, 0.f) * scale + __ldg(a.pe + (size_t)t * a.dk + c4 + j)) * m[j];
        }
        *reinterpret_cast<float4*>(a.Y0 + row * a.dk + c4) = make_float4(o[0], o[1], o[2], o[3]);
        if (c4 == 0) {
            a.kpad[row] = allz ? 1 : 0;
            for (int f = 0; f < a.F; f++) a.Xs[row * a.F + f] = f < 4 ? xs[f] : __ldg(xw + gi[f]);
        }
    }
}

// dpre = dY0 * drop * sqrt(dk) * (pre > 0);  dWe[d, f] += dpre xs[f];  dbe[d] += dpre.  F <= 4.
__global__ void __launch_bounds__(256) tfm_embed_bwd_kernel(const TfmEmbedArgs a) {
    extern __shared__ float esm[];                 // [dk * (F + 1)]
    const int cg = a.dk >> 2, F = a.F;
    const int rows_per = blockDim.x / cg;
    const int lane_row = threadIdx.x / cg, c4 = (threadIdx.x % cg) * 4;
    const long long R = (long long)a.B * a.G * a.T;
    const float scale = sqrtf((float)a.dk);
    for (int i = threadIdx.x; i < a.dk * (F + 1); i += blockDim.x) esm[i] = 0.f;
    __syncthreads();
    float acc[4][5];
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int f = 0; f < 5; f++) acc[j][f] = 0.f;
    float w[4][4], bb[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        bb[j] = __ldg(a.be + c4 + j);
#pragma unroll
        for (int f = 0; f < 4; f++) w[j][f] = f < F ? __ldg(a.We + (size_t)(c4 + j) * F + f) : 0.f;
    }
    if (lane_row < rows_per) {
        for (long long row = (long long)blockIdx.x * rows_per + lane_row; row < R; row += (long long)gridDim.x * rows_per) {
            float xs[4] = {0.f, 0.f, 0.f, 0.f};
            for (int f = 0; f < F; f++) xs[f] = a.Xs[row * F + f];
            const float4 dy = *reinterpret_cast<const float4*>(a.dY0 + row * a.dk + c4);
            float m[4];
            drop_mul4(a.drop, (unsigned long long)row * a.dk + c4, m);
            const float d4[4] = {dy.x, dy.y, dy.z, dy.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float pre = bb[j];
#pragma unroll
                for (int f = 0; f < 4; f++) pre = fmaf(xs[f], w[j][f], pre);
                const float dp = pre > 0.f ? d4[j] * m[j] * scale : 0.f;
#pragma unroll
                for (int f = 0; f < 4; f++) acc[j][f] = fmaf(dp, xs[f], acc[j][f]);
                acc[j][4] += dp;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            for (int f = 0; f < F; f++) atomicAdd(esm + (c4 + j) * (F + 1) + f, acc[j][f]);
            atomicAdd(esm + (c4 + j) * (F + 1) + F, acc[j][4]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < a.dk * (F + 1); i += blockDim.x) {
        const int d = i / (F + 1), f = i % (F + 1);
        if (f < F) atomicAdd(a.dWe + (size_t)d * F + f, esm[i]);
        else atomicAdd(a.dbe + d, esm[i]);
    }
}

// ---------------------------------------------------------------------------
// (residual + dropout) + LayerNorm, forward and backward.  LPR lanes share a row (LPR * 4 * V4 >= W), a warp covers
// 32 / LPR consecutive rows with float4 accesses; statistics by xor-shuffles inside the LPR-lane group.
//   forward : u = res[map(r)] + branch[r] * drop(map(r));  out = LN(u)       (res == null: u = branch, no dropout)
//   backward: dx = LN'(dy);  dx_out (+)= dx;  ddrop = dx * drop(map(r))       (either output may be null)
// map(r) = r * row_mul + row_add selects the residual row, r * drp_mul + drp_add the row of the dropout elements (layers that
// only keep the last step of every sequence).
// ---------------------------------------------------------------------------
struct TfmLnArgs {
    const float* res;      // [*, W] residual stream, row map(r)
    const float* branch;   // [R, W]
    float* u;              // [R, W] pre-LN sum (may alias branch) or null
    float* out;            // [R, W]
    float* mu; float* rs;  // [R]
    const float* w; const float* b;
    // backward
    const float* dy;       // [R, W]
    const float* x;        // [R, W] = u of the forward
    float* dx; int dx_accum;
    float* ddrop;          // [R, W] or null
    float* dw; float* db;  // [W] atomically accumulated
    DropSite drop;
    long long R, row_mul, row_add, drp_mul, drp_add;
    int W;
    float eps;
};

template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int LPR, int V4>
__global__ void __launch_bounds__(256) tfm_ln_fwd_kernel(const TfmLnArgs a) {
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, sub = lane % LPR, rw = lane / LPR;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const int W = a.W;
    const float invW = 1.0f / (float)W;
    for (long long r0 = gw * RPW; r0 < a.R; r0 += nw * RPW) {
        const long long r = r0 + rw;
        const bool live = r < a.R;
        const long long mr = r * a.row_mul + a.row_add, md = r * a.drp_mul + a.drp_add;
        float v[V4][4];
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < V4; q++) {
            const int c = (sub + q * LPR) * 4;
            v[q][0] = v[q][1] = v[q][2] = v[q][3] = 0.f;
            if (live && c < W) {
                const float4 bv = *reinterpret_cast<const float4*>(a.branch + r * W + c);
                v[q][0] = bv.x; v[q][1] = bv.y; v[q][2] = bv.z; v[q][3] = bv.w;
                if (a.res) {
                    float m[4];
                    drop_mul4(a.drop, (unsigned long long)md * W + c, m);
                    const float4 rv = *reinterpret_cast<const float4*>(a.res + mr * W + c);
                    v[q][0] = fmaf(v[q][0], m[0], rv.x); v[q][1] = fmaf(v[q][1], m[1], rv.y);
                    v[q][2] = fmaf(v[q][2], m[2], rv.z); v[q][3] = fmaf(v[q][3], m[3], rv.w);
                }
                if (a.u) *reinterpret_cast<float4*>(a.u + r * W + c) = make_float4(v[q][0], v[q][1], v[q][2], v[q][3]);
                s += v[q][0] + v[q][1] + v[q][2] + v[q][3];
            }
        }
        const float mu = group_sum<LPR>(s) * invW;
        float sq = 0.f;
#pragma unroll
        for (int q = 0; q < V4; q++) {
            const int c = (sub + q * LPR) * 4;
            if (c < W) {
#pragma unroll
                for (int j = 0; j < 4; j++) { const float d = v[q][j] - mu; sq = fmaf(d, d, sq); }
            }
        }
        const float rs = rsqrtf(group_sum<LPR>(sq) * invW + a.eps);
        if (live) {
#pragma unroll
            for (int q = 0; q < V4; q++) {
                const int c = (sub + q * LPR) * 4;
                if (c < W) {
                    // the affine parameters sit at arbitrary float offsets of the flat state: scalar loads
                    *reinterpret_cast<float4*>(a.out + r * W + c) =
                        make_float4((v[q][0] - mu) * rs * __ldg(a.w + c) + __ldg(a.b + c), (v[q][1] - mu) * rs * __ldg(a.w + c + 1) + __ldg(a.b + c + 1),
                                    (v[q][2] - mu) * rs * __ldg(a.w + c + 2) + __ldg(a.b + c + 2), (v[q][3] - mu) * rs * __ldg(a.w + c + 3) + __ldg(a.b + c + 3));
                }
            }
            if (sub == 0) { a.mu[r] = mu; a.rs[r] = rs; }
        }
    }
}

template <int LPR, int V4>
__global__ void __launch_bounds__(256) tfm_ln_bwd_kernel(const TfmLnArgs a) {
    constexpr int RPW = 32 / LPR;
    __shared__ float sacc[2 * 32 * 4 * V4 * 1];     // [2][W <= 128 * V4]
    const int lane = threadIdx.x & 31, sub = lane % LPR, rw = lane / LPR;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const int W = a.W;
    const float invW = 1.0f / (float)W;
    for (int i = threadIdx.x; i < 2 * W; i += blockDim.x) sacc[i] = 0.f;
    __syncthreads();
    float aw[V4][4], ab[V4][4], wv[V4][4];
#pragma unroll
    for (int q = 0; q < V4; q++) {
        const int c = (sub + q * LPR) * 4;
#pragma unroll
        for (int j = 0; j < 4; j++) { aw[q][j] = ab[q][j] = 0.f; wv[q][j] = c < W ? __ldg(a.w + c + j) : 0.f; }
    }
    for (long long r0 = gw * RPW; r0 < a.R; r0 += nw * RPW) {
        const long long r = r0 + rw;
        const bool live = r < a.R;
        const long long md = r * a.drp_mul + a.drp_add;
        float xh[V4][4], g[V4][4];
        float s1 = 0.f, s2 = 0.f;
        float mu = 0.f, rs = 0.f;
        if (live) { mu = a.mu[r]; rs = a.rs[r]; }
#pragma unroll
        for (int q = 0; q < V4; q++) {
            const int c = (sub + q * LPR) * 4;
#pragma unroll
            for (int j = 0; j < 4; j++) xh[q][j] = g[q][j] = 0.f;
            if (live && c < W) {
                const float4 xv = *reinterpret_cast<const float4*>(a.x + r * W + c);
                const float4 dv = *reinterpret_cast<const float4*>(a.dy + r * W + c);
                const float xx[4] = {xv.x, xv.y, xv.z, xv.w}, dd[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    xh[q][j] = (xx[j] - mu) * rs;
                    g[q][j] = dd[j] * wv[q][j];
                    aw[q][j] = fmaf(dd[j], xh[q][j], aw[q][j]);
                    ab[q][j] += dd[j];
                    s1 += g[q][j];
                    s2 = fmaf(g[q][j], xh[q][j], s2);
                }
            }
        }
        const float m1 = group_sum<LPR>(s1) * invW, m2 = group_sum<LPR>(s2) * invW;
        if (live) {
#pragma unroll
            for (int q = 0; q < V4; q++) {
                const int c = (sub + q * LPR) * 4;
                if (c < W) {
                    float d[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) d[j] = rs * (g[q][j] - m1 - xh[q][j] * m2);
                    if (a.ddrop) {
                        float m[4];
                        drop_mul4(a.drop, (unsigned long long)md * W + c, m);
                        *reinterpret_cast<float4*>(a.ddrop + r * W + c) = make_float4(d[0] * m[0], d[1] * m[1], d[2] * m[2], d[3] * m[3]);
                    }
                    if (a.dx) {
                        float4* p = reinterpret_cast<float4*>(a.dx + r * W + c);
                        if (a.dx_accum) { const float4 o = *p; d[0] += o.x; d[1] += o.y; d[2] += o.z; d[3] += o.w; }
                        *p = make_float4(d[0], d[1], d[2], d[3]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < V4; q++) {
        const int c = (sub + q * LPR) * 4;
        if (c < W) {
#pragma unroll
            for (int j = 0; j < 4; j++) { atomicAdd(sacc + c + j, aw[q][j]); atomicAdd(sacc + W + c + j, ab[q][j]); }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < W; i += blockDim.x) { atomicAdd(a.dw + i, sacc[i]); atomicAdd(a.db + i, sacc[W + i]); }
}

// ---------------------------------------------------------------------------
// element-wise pieces (float4; n % 4 == 0)
// ---------------------------------------------------------------------------
__device__ __forceinline__ float gelu_f(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float v) {
    return 0.5f * (1.0f + erff(v * 0.70710678118654752f)) + v * 0.3989422804014327f * __expf(-0.5f * v * v);
}

// mode 0: out = a + b * drop            (residual add)
// mode 1: out = a * drop                (gradient through a dropout)
// mode 2: out = gelu(a) * drop          (FFN hidden: GELU then dropout)
// mode 3: out = a * drop * gelu'(b)     (gradient through dropout and GELU; b = pre-activation)
struct TfmEwArgs { const float* a; const float* b; float* out; DropSite drop; long long n; int mode; };

__global__ void __launch_bounds__(256) tfm_ew_kernel(const TfmEwArgs e) {
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < e.n; i += (long long)gridDim.x * blockDim.x * 4) {
        float m[4];
        drop_mul4(e.drop, (unsigned long long)i, m);
        const float4 av = *reinterpret_cast<const float4*>(e.a + i);
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e.mode == 0 || e.mode == 3) bv = *reinterpret_cast<const float4*>(e.b + i);
        float4 o;
        if (e.mode == 0) o = make_float4(fmaf(bv.x, m[0], av.x), fmaf(bv.y, m[1], av.y), fmaf(bv.z, m[2], av.z), fmaf(bv.w, m[3], av.w));
        else if (e.mode == 1) o = make_float4(av.x * m[0], av.y * m[1], av.z * m[2], av.w * m[3]);
        else if (e.mode == 2) o = make_float4(gelu_f(av.x) * m[0], gelu_f(av.y) * m[1], gelu_f(av.z) * m[2], gelu_f(av.w) * m[3]);
        else o = make_float4(av.x * m[0] * gelu_grad_f(bv.x), av.y * m[1] * gelu_grad_f(bv.y), av.z * m[2] * gelu_grad_f(bv.z),
                             av.w * m[3] * gelu_grad_f(bv.w));
        *reinterpret_cast<float4*>(e.out + i) = o;
    }
}

// scalar tail-safe variants for buffers whose length is not a multiple of 4 (latent expansion of odd sizes)
__global__ void tfm_gelu_fwd_kernel(const float* __restrict__ pre, float* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = gelu_f(pre[i]);
}
__global__ void tfm_gelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ pre, float* __restrict__ dpre, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dpre[i] = dy[i] * gelu_grad_f(pre[i]);
}

// dst[(r * mul + add) * W + c] += src[r * W + c]   (gradient of "take the last step")
__global__ void tfm_row_scatter_add_kernel(const float* __restrict__ src, float* __restrict__ dst, long long R, int W, long long mul,
                                           long long add) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * W) return;
    const long long r = i / W;
    const int c = (int)(i % W);
    dst[(r * mul + add) * W + c] += src[i];
}

// ---------------------------------------------------------------------------
// encoder head (TFMEncoderPT.forward :1146-1164)
// ---------------------------------------------------------------------------
// h = clamp(concat(on, oe) / max(rms, 1), +-1e4), NaN -> 0; one warp per window.  Backward (dh given):
//   d enc = rms > 1 ? (dh - h * sum(dh h) / n) / rms : dh,  then the ReLU mask of the CensNet outputs (on, oe > 0).
struct TfmRmsArgs {
    const float* on; const float* oe;   // [B, ND], [B, ED]  (>= 0: ReLU already applied)
    float* h;                           // [B, ND + ED]
    float* rms;                         // [B]
    const float* dh;                    // backward
    float* don; float* doe;
    int B, ND, ED;
    int norelu;                         // backward: the inputs are NOT ReLU outputs (TCN decoder latent): no mask
};

__global__ void __launch_bounds__(256) tfm_rms_fwd_kernel(const TfmRmsArgs a) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= a.B) return;
    const int KD = a.ND + a.ED;
    float ss = 0.f;
    for (int k = lane; k < KD; k += 32) {
        const float x = k < a.ND ? a.on[(size_t)warp * a.ND + k] : a.oe[(size_t)warp * a.ED + (k - a.ND)];
        ss = fmaf(x, x, ss);
    }
    const float rms = sqrtf(warp_sum(ss) / (float)KD);
    const float inv = 1.0f / fmaxf(rms, 1.0f);
    for (int k = lane; k < KD; k += 32) {
        const float x = k < a.ND ? a.on[(size_t)warp * a.ND + k] : a.oe[(size_t)warp * a.ED + (k - a.ND)];
        float v = fminf(fmaxf(x * inv, -1e4f), 1e4f);
        if (v != v) v = 0.f;
        a.h[(size_t)warp * KD + k] = v;
    }
    if (lane == 0) a.rms[warp] = rms;
}

__global__ void __launch_bounds__(256) tfm_rms_bwd_kernel(const TfmRmsArgs a) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= a.B) return;
    const int KD = a.ND + a.ED;
    const float rms = a.rms[warp];
    float dot = 0.f;
    if (rms > 1.0f)
        for (int k = lane; k < KD; k += 32) dot = fmaf(a.dh[(size_t)warp * KD + k], a.h[(size_t)warp * KD + k], dot);
    dot = warp_sum(dot) / (float)KD;
    const float inv = rms > 1.0f ? 1.0f / rms : 1.0f;
    for (int k = lane; k < KD; k += 32) {
        const float hv = a.h[(size_t)warp * KD + k];
        float d = a.dh[(size_t)warp * KD + k];
        if (fabsf(hv) > 1e4f) d = 0.f;
        if (rms > 1.0f) d = (d - hv * dot) * inv;
        if (k < a.ND) a.don[(size_t)warp * a.ND + k] = (a.norelu || a.on[(size_t)warp * a.ND + k] > 0.f) ? d : 0.f;
        else a.doe[(size_t)warp * a.ED + (k - a.ND)] = a.oe[(size_t)warp * a.ED + (k - a.ND)] > 0.f ? d : 0.f;
    }
}

// Column statistics over the batch.  One CTA per (32 columns, statistics group); blockDim = (32, 32): threadIdx.x =
// column, threadIdx.y = row lane.  The batch is split into `groups` equal row ranges with separate statistics (the two
// encoder passes of the contrastive step see B rows each, training.py:527-531).
//   kind 0  BatchNorm1d, train: y = (x - mean) * rsqrt(var_biased + eps) * w + b;  stat_out = mean | unbiased var
//   kind 1  BatchNorm1d, eval : y = (x - running_mean) * rsqrt(running_var + eps) * w + b
//   kind 2  batch standardisation: y = (x - mean) / max(std_unbiased, 0.1)
struct TfmColArgs {
    const float* x; float* y;                 // [B, C]
    const float* w; const float* b;           // [C]
    const float* run_mean; const float* run_var;
    float* mean; float* scale;                // [groups, C] saved for the backward: mean, rstd (kind 0) / mean, 1/s (kind 2)
    float* stat_out;                          // [groups, 2, C] kind 0: batch mean | unbiased variance (running-buffer update)
    unsigned char* clamped;                   // [groups, C] kind 2: 1 = std was clamped
    // backward
    const float* dy; float* dx; float* dw; float* db;
    const float* relu_ref;                    // kind 0 backward: dx = relu_ref > 0 ? dx : 0  (x is a ReLU output) or null
    int B, C, groups, kind;
    float eps;
};

__device__ __forceinline__ float col_reduce(float v, float (*sm)[33]) {
    sm[threadIdx.y][threadIdx.x] = v;
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) {
        if (threadIdx.y < o) sm[threadIdx.y][threadIdx.x] += sm[threadIdx.y + o][threadIdx.x];
        __syncthreads();
    }
    const float r = sm[0][threadIdx.x];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(1024) tfm_col_fwd_kernel(const TfmColArgs a) {
    __shared__ float sm[32][33];
    const int c = blockIdx.x * 32 + threadIdx.x, grp = blockIdx.y;
    const int Bg = a.B / a.groups, r0 = grp * Bg;
    const bool live = c < a.C;
    float mean = 0.f, scale = 1.f;
    if (a.kind == 1) {
        if (live) { mean = a.run_mean[c]; scale = rsqrtf(a.run_var[c] + a.eps); }
    } else {
        float s = 0.f;
        if (live) for (int r = threadIdx.y; r < Bg; r += 32) s += a.x[(size_t)(r0 + r) * a.C + c];
        mean = col_reduce(s, sm) / (float)Bg;
        float q = 0.f;
        if (live) for (int r = threadIdx.y; r < Bg; r += 32) { const float d = a.x[(size_t)(r0 + r) * a.C + c] - mean; q = fmaf(d, d, q); }
        q = col_reduce(q, sm);
        if (a.kind == 0) {
            scale = rsqrtf(q / (float)Bg + a.eps);
            if (live && threadIdx.y == 0 && a.stat_out) {
                a.stat_out[((size_t)grp * 2) * a.C + c] = mean;
                a.stat_out[((size_t)grp * 2 + 1) * a.C + c] = q / (float)(Bg > 1 ? Bg - 1 : 1);
            }
        } else {
            const float sd = sqrtf(q / (float)(Bg - 1));
            scale = 1.0f / fmaxf(sd, 0.1f);
            if (live && threadIdx.y == 0) a.clamped[(size_t)grp * a.C + c] = sd < 0.1f ? 1 : 0;
        }
        if (live && threadIdx.y == 0) { a.mean[(size_t)grp * a.C + c] = mean; a.scale[(size_t)grp * a.C + c] = scale; }
    }
    if (!live) return;
    const float w = a.kind == 2 ? 1.f : a.w[c], b = a.kind == 2 ? 0.f : a.b[c];
    for (int r = threadIdx.y; r < Bg; r += 32) a.y[(size_t)(r0 + r) * a.C + c] = (a.x[(size_t)(r0 + r) * a.C + c] - mean) * scale * w + b;
}

// kind 0: dx = w * rstd * (dy - mean(dy) - xhat * mean(dy xhat));  dw += sum dy xhat;  db += sum dy
// kind 2: dx = (dy - mean(dy) - [not clamped] y * sum(dy y) / (B - 1)) / s
__global__ void __launch_bounds__(1024) tfm_col_bwd_kernel(const TfmColArgs a) {
    __shared__ float sm[32][33];
    const int c = blockIdx.x * 32 + threadIdx.x, grp = blockIdx.y;
    const int Bg = a.B / a.groups, r0 = grp * Bg;
    const bool live = c < a.C;
    const float mean = live ? a.mean[(size_t)grp * a.C + c] : 0.f, scale = live ? a.scale[(size_t)grp * a.C + c] : 0.f;
    float s1 = 0.f, s2 = 0.f;
    if (live)
        for (int r = threadIdx.y; r < Bg; r += 32) {
            const float d = a.dy[(size_t)(r0 + r) * a.C + c];
            const float xh = (a.x[(size_t)(r0 + r) * a.C + c] - mean) * scale;
            s1 += d;
            s2 = fmaf(d, xh, s2);
        }
    s1 = col_reduce(s1, sm);
    s2 = col_reduce(s2, sm);
    if (!live) return;
    if (a.kind == 0) {
        const float w = a.w[c];
        if (threadIdx.y == 0) { atomicAdd(a.dw + c, s2); atomicAdd(a.db + c, s1); }
        const float m1 = s1 / (float)Bg, m2 = s2 / (float)Bg;
        for (int r = threadIdx.y; r < Bg; r += 32) {
            const size_t i = (size_t)(r0 + r) * a.C + c;
            const float xh = (a.x[i] - mean) * scale;
            float d = w * scale * (a.dy[i] - m1 - xh * m2);
            if (a.relu_ref && !(a.relu_ref[i] > 0.f)) d = 0.f;
            a.dx[i] = d;
        }
    } else {
        const float m1 = s1 / (float)Bg;
        const float m2 = a.clamped[(size_t)grp * a.C + c] ? 0.f : s2 / (float)(Bg - 1);
        for (int r = threadIdx.y; r < Bg; r += 32) {
            const size_t i = (size_t)(r0 + r) * a.C + c;
            const float yh = (a.x[i] - mean) * scale;
            a.dx[i] = (a.dy[i] - m1 - yh * m2) * scale;
        }
    }
}

// running = (1 - momentum) * running + momentum * batch statistic, one group after the other (models_new.py:508-516)
__global__ void tfm_bn_update_kernel(float* __restrict__ run_mean, float* __restrict__ run_var, float* __restrict__ tracked,
                                     const float* __restrict__ stat, int C, int groups, float momentum) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && tracked) *tracked += (float)groups;
    if (c >= C) return;
    float m = run_mean[c], v = run_var[c];
    for (int g = 0; g < groups; g++) {
        m = (1.f - momentum) * m + momentum * stat[((size_t)g * 2) * C + c];
        v = (1.f - momentum) * v + momentum * stat[((size_t)g * 2 + 1) * C + c];
    }
    run_mean[c] = m; run_var[c] = v;
}
