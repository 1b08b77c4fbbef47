// Recurrent part of torch.nn.GRU (gate order r,z,n), one direction per blockIdx.y.
//
// The input projection  Gi = X.W_ih^T + b_ih  for all time steps is a tall-skinny GEMM
// (gemm.cuh) and is NOT on the serial path.  These kernels run the T serial steps for a
// tile of MT sequences per CTA:  gh = h.W_hh^T + b_hh  (register-tiled, W_hh and h in
// shared memory), the gate math, and the write of h_t (and, for training, of the gates
// r,z,n and hn = W_hn.h+b_hn that the backward pass needs).
//
// Packed-sequence semantics of the reference (models_new.py:233-249, oracle gru_direction):
// a sequence has a valid prefix len[s]; on invalid steps the state is held and the output
// is zero; the reverse direction starts at T-1 with h=0, so it only "starts" at len-1.
#pragma once
#include "common.cuh"

struct GruFwdArgs {
    const float* Gi[2]; long long gi_ss; int gi_st;  // Gi[dir][s*gi_ss + t*gi_st + c]
    const float* Whh[2]; const float* bhh[2];
    const int* len;      // [S] or null (all T)
    float* Hout;         // [S,T,2H] (this direction's half) or null
    float* Gt[2];        // [S,T,4H] = r|z|n|hn per direction, or null
    float* Hn;           // [S,2H] final states [fwd|bwd] or null
    int S, T, H;
};

__device__ __forceinline__ float4 ld4g(const float* p, int j0, int H, bool vec) {
    if (vec) return __ldg(reinterpret_cast<const float4*>(p + j0));
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j0 + 0 < H) v.x = __ldg(p + j0 + 0);
    if (j0 + 1 < H) v.y = __ldg(p + j0 + 1);
    if (j0 + 2 < H) v.z = __ldg(p + j0 + 2);
    if (j0 + 3 < H) v.w = __ldg(p + j0 + 3);
    return v;
}
__device__ __forceinline__ void st4g(float* p, int j0, int H, bool vec, float4 v) {
    if (vec) { *reinterpret_cast<float4*>(p + j0) = v; return; }
    if (j0 + 0 < H) p[j0 + 0] = v.x;
    if (j0 + 1 < H) p[j0 + 1] = v.y;
    if (j0 + 2 < H) p[j0 + 2] = v.z;
    if (j0 + 3 < H) p[j0 + 3] = v.w;
}

#define GRU_TM 4

template <int HP, int MT>
__global__ void __launch_bounds__((MT / GRU_TM) * (HP / 4))
gru_fwd_kernel(const GruFwdArgs a) {
    constexpr int NJ = HP / 4;
    constexpr int NT = (MT / GRU_TM) * NJ;
    extern __shared__ __align__(16) float smem[];
    float* WT = smem;                    // [HP][3*HP]   WT[k][g*HP+j] = Whh[(g*H+j)*H+k]
    float* bh = WT + HP * 3 * HP;        // [3*HP]
    float* hT = bh + 3 * HP;             // [2][HP][MT]
    const int dir = blockIdx.y;
    const int H = a.H, T = a.T;
    const int tid = threadIdx.x;
    const int jg = tid % NJ, mg = tid / NJ;
    const int j0 = jg * 4;
    const int s0 = blockIdx.x * MT + mg * GRU_TM;
    const bool vec = ((H & 3) == 0) && (j0 + 3 < H);   // HP may exceed H: idle j-groups use the guarded path
    const float* Whh = a.Whh[dir];
    const float* bhh = a.bhh[dir];
    for (int i = tid; i < HP * 3 * HP; i += NT) {
        int k = i / (3 * HP), c = i % (3 * HP);
        int g = c / HP, j = c % HP;
        WT[i] = (k < H && j < H) ? __ldg(Whh + (size_t)(g * H + j) * H + k) : 0.f;
    }
    for (int i = tid; i < 3 * HP; i += NT) {
        int g = i / HP, j = i % HP;
        bh[i] = (j < H) ? __ldg(bhh + g * H + j) : 0.f;
    }
    for (int i = tid; i < 2 * HP * MT; i += NT) hT[i] = 0.f;
    int len[GRU_TM];
#pragma unroll
    for (int m = 0; m < GRU_TM; m++) {
        int s = s0 + m;
        len[m] = (s < a.S) ? (a.len ? __ldg(a.len + s) : T) : 0;
    }
    float h[GRU_TM][4];
#pragma unroll
    for (int m = 0; m < GRU_TM; m++)
#pragma unroll
        for (int j = 0; j < 4; j++) h[m][j] = 0.f;
    __syncthreads();
    const float* Gi = a.Gi[dir];
    float* Gt = a.Gt[dir];
    int cur = 0;
    for (int step = 0; step < T; step++) {
        const int t = dir ? (T - 1 - step) : step;
        float4 gi[3][GRU_TM];
#pragma unroll
        for (int m = 0; m < GRU_TM; m++) {
            int s = s0 + m;
            if (s < a.S) {
                const float* gp = Gi + (size_t)s * a.gi_ss + (size_t)t * a.gi_st;
#pragma unroll
                for (int g = 0; g < 3; g++) gi[g][m] = ld4g(gp + g * H, j0, H, vec);
            } else {
#pragma unroll
                for (int g = 0; g < 3; g++) gi[g][m] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        float acc[3][GRU_TM][4];
#pragma unroll
        for (int g = 0; g < 3; g++) {
            float4 b = *reinterpret_cast<const float4*>(&bh[g * HP + j0]);
#pragma unroll
            for (int m = 0; m < GRU_TM; m++) {
                acc[g][m][0] = b.x; acc[g][m][1] = b.y; acc[g][m][2] = b.z; acc[g][m][3] = b.w;
            }
        }
        const float* hc = hT + cur * HP * MT + mg * GRU_TM;
#pragma unroll 4
        for (int k = 0; k < HP; k++) {
            float4 hv4 = *reinterpret_cast<const float4*>(hc + k * MT);
            float hv[4] = {hv4.x, hv4.y, hv4.z, hv4.w};
#pragma unroll
            for (int g = 0; g < 3; g++) {
                float4 w4 = *reinterpret_cast<const float4*>(&WT[k * 3 * HP + g * HP + j0]);
                float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int m = 0; m < GRU_TM; m++)
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[g][m][j] = fmaf(hv[m], wv[j], acc[g][m][j]);
            }
        }
        float* hn = hT + (cur ^ 1) * HP * MT + mg * GRU_TM;
#pragma unroll
        for (int m = 0; m < GRU_TM; m++) {
            int s = s0 + m;
            const bool valid = t < len[m];
            float gir[4] = {gi[0][m].x, gi[0][m].y, gi[0][m].z, gi[0][m].w};
            float giz[4] = {gi[1][m].x, gi[1][m].y, gi[1][m].z, gi[1][m].w};
            float gin[4] = {gi[2][m].x, gi[2][m].y, gi[2][m].z, gi[2][m].w};
            float r[4], z[4], n[4], o[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                r[j] = sigmoid_f(gir[j] + acc[0][m][j]);
                z[j] = sigmoid_f(giz[j] + acc[1][m][j]);
                n[j] = tanh_f(gin[j] + r[j] * acc[2][m][j]);
                float hnew = (1.0f - z[j]) * n[j] + z[j] * h[m][j];
                h[m][j] = valid ? hnew : h[m][j];
                o[j] = valid ? h[m][j] : 0.f;
            }
            if (s < a.S) {
                if (a.Hout)
                    st4g(a.Hout + ((size_t)s * T + t) * 2 * H + dir * H, j0, H, vec,
                         make_float4(o[0], o[1], o[2], o[3]));
                if (Gt && valid) {
                    float* gp = Gt + ((size_t)s * T + t) * 4 * H;
                    st4g(gp, j0, H, vec, make_float4(r[0], r[1], r[2], r[3]));
                    st4g(gp + H, j0, H, vec, make_float4(z[0], z[1], z[2], z[3]));
                    st4g(gp + 2 * H, j0, H, vec, make_float4(n[0], n[1], n[2], n[3]));
                    st4g(gp + 3 * H, j0, H, vec,
                         make_float4(acc[2][m][0], acc[2][m][1], acc[2][m][2], acc[2][m][3]));
                }
            }
        }
        // publish h (transposed) for the next step
#pragma unroll
        for (int j = 0; j < 4; j++)
            *reinterpret_cast<float4*>(hn + (j0 + j) * MT) = make_float4(h[0][j], h[1][j], h[2][j], h[3][j]);
        __syncthreads();
        cur ^= 1;
    }
    if (a.Hn) {
#pragma unroll
        for (int m = 0; m < GRU_TM; m++) {
            int s = s0 + m;
            if (s < a.S)
                st4g(a.Hn + (size_t)s * 2 * H + dir * H, j0, H, vec, make_float4(h[m][0], h[m][1], h[m][2], h[m][3]));
        }
    }
}

// ---------------------------------------------------------------------------
// BPTT.  Per step (reverse of the forward order):
//   dh += dOut[t];  dn = dh(1-z); dz = dh(hprev-n); da_n = dn(1-n^2); da_z = dz z(1-z);
//   da_r = da_n*hn * r(1-r);  dGi = [da_r,da_z,da_n];  dGh = [da_r,da_z,da_n*r];
//   dh_prev = dh*z + dGh.W_hh
// dG is written as [S,T,4H] = da_r | da_z | da_n*r | da_n  so that dGh = cols [0,3H) and
// dGi = cols [0,2H) u [3H,4H) (MatView A_SPLIT) feed the weight/input-gradient GEMMs.
// ---------------------------------------------------------------------------
struct GruBwdArgs {
    const float* Whh[2];
    const int* len;
    const float* Hout;        // [S,T,2H] forward outputs (h_prev source)
    const float* Gt[2];       // saved gates
    const float* dOut;        // [S,T,2H] or null
    const float* dHn;         // [S,2H] or null
    float* dG[2];             // [S,T,4H]
    int S, T, H;
};

template <int HP, int MT>
__global__ void __launch_bounds__((MT / GRU_TM) * (HP / 4))
gru_bwd_kernel(const GruBwdArgs a) {
    constexpr int NJ = HP / 4;
    constexpr int NT = (MT / GRU_TM) * NJ;
    extern __shared__ __align__(16) float smem[];
    float* W = smem;                     // [3*HP][HP]  W[(g*HP+j)][k] = Whh[(g*H+j)*H+k]
    float* dGT = W + 3 * HP * HP;        // [2][3*HP][MT]
    const int dir = blockIdx.y;
    const int H = a.H, T = a.T;
    const int tid = threadIdx.x;
    const int jg = tid % NJ, mg = tid / NJ;
    const int j0 = jg * 4;
    const int s0 = blockIdx.x * MT + mg * GRU_TM;
    const bool vec = ((H & 3) == 0) && (j0 + 3 < H);   // HP may exceed H: idle j-groups use the guarded path
    const float* Whh = a.Whh[dir];
    for (int i = tid; i < 3 * HP * HP; i += NT) {
        int c = i / HP, k = i % HP;
        int g = c / HP, j = c % HP;
        W[i] = (k < H && j < H) ? __ldg(Whh + (size_t)(g * H + j) * H + k) : 0.f;
    }
    int len[GRU_TM];
    float dh[GRU_TM][4];
#pragma unroll
    for (int m = 0; m < GRU_TM; m++) {
        int s = s0 + m;
        len[m] = (s < a.S) ? (a.len ? __ldg(a.len + s) : T) : 0;
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.dHn && s < a.S) d = ld4g(a.dHn + (size_t)s * 2 * H + dir * H, j0, H, vec);
        dh[m][0] = d.x; dh[m][1] = d.y; dh[m][2] = d.z; dh[m][3] = d.w;
    }
    __syncthreads();
    const float* Gt = a.Gt[dir];
    float* dG = a.dG[dir];
    int cur = 0;
    for (int step = 0; step < T; step++) {
        const int t = dir ? step : (T - 1 - step);
        const int tp = dir ? t + 1 : t - 1;
        float* dgs = dGT + cur * 3 * HP * MT + mg * GRU_TM;
        float part[GRU_TM][4];
        float dar[GRU_TM][4], daz[GRU_TM][4], dah[GRU_TM][4];
        // pull the rows of the NEXT step towards L2 while this step computes (the saved gates were written a whole
        // forward pass ago and are cold); one thread per 128-byte line
        if (step + 1 < T && (j0 & 31) == 0) {
            const int tn = dir ? t + 1 : t - 1;
            const int tpn = dir ? tn + 1 : tn - 1;
#pragma unroll
            for (int m = 0; m < GRU_TM; m++) {
                const int s = s0 + m;
                if (s < a.S && tn < len[m]) {
                    const float* gp = Gt + ((size_t)s * T + tn) * 4 * H + j0;
#pragma unroll
                    for (int q = 0; q < 4; q++) asm volatile("prefetch.global.L2 [%0];" ::"l"(gp + q * H));
                    if (tpn >= 0 && tpn < T) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.Hout + ((size_t)s * T + tpn) * 2 * H + dir * H + j0));
                    if (a.dOut) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.dOut + ((size_t)s * T + tn) * 2 * H + dir * H + j0));
                }
            }
        }
#pragma unroll
        for (int m = 0; m < GRU_TM; m++) {
            int s = s0 + m;
            const bool valid = t < len[m];
            float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            float o_r[4] = {0, 0, 0, 0}, o_z[4] = {0, 0, 0, 0}, o_h[4] = {0, 0, 0, 0}, o_n[4] = {0, 0, 0, 0};
            if (valid) {
                const float* gp = Gt + ((size_t)s * T + t) * 4 * H;
                float4 r4 = ld4g(gp, j0, H, vec);
                z4 = ld4g(gp + H, j0, H, vec);
                float4 n4 = ld4g(gp + 2 * H, j0, H, vec);
                float4 q4 = ld4g(gp + 3 * H, j0, H, vec);
                float4 hp4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (tp >= 0 && tp < len[m]) hp4 = ld4g(a.Hout + ((size_t)s * T + tp) * 2 * H + dir * H, j0, H, vec);
                float4 do4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (a.dOut) do4 = ld4g(a.dOut + ((size_t)s * T + t) * 2 * H + dir * H, j0, H, vec);
                float r[4] = {r4.x, r4.y, r4.z, r4.w}, z[4] = {z4.x, z4.y, z4.z, z4.w};
                float n[4] = {n4.x, n4.y, n4.z, n4.w}, q[4] = {q4.x, q4.y, q4.z, q4.w};
                float hp[4] = {hp4.x, hp4.y, hp4.z, hp4.w}, dov[4] = {do4.x, do4.y, do4.z, do4.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    float d = dh[m][j] + dov[j];
                    float dn = d * (1.0f - z[j]);
                    float dz = d * (hp[j] - n[j]);
                    float dan = dn * (1.0f - n[j] * n[j]);
                    float dazv = dz * z[j] * (1.0f - z[j]);
                    float darv = dan * q[j] * r[j] * (1.0f - r[j]);
                    o_r[j] = darv; o_z[j] = dazv; o_h[j] = dan * r[j]; o_n[j] = dan;
                    part[m][j] = d * z[j];
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) part[m][j] = dh[m][j];
            }
#pragma unroll
            for (int j = 0; j < 4; j++) { dar[m][j] = o_r[j]; daz[m][j] = o_z[j]; dah[m][j] = o_h[j]; }
            if (s < a.S) {
                float* op = dG + ((size_t)s * T + t) * 4 * H;
                st4g(op, j0, H, vec, make_float4(o_r[0], o_r[1], o_r[2], o_r[3]));
                st4g(op + H, j0, H, vec, make_float4(o_z[0], o_z[1], o_z[2], o_z[3]));
                st4g(op + 2 * H, j0, H, vec, make_float4(o_h[0], o_h[1], o_h[2], o_h[3]));
                st4g(op + 3 * H, j0, H, vec, make_float4(o_n[0], o_n[1], o_n[2], o_n[3]));
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            *reinterpret_cast<float4*>(dgs + (0 * HP + j0 + j) * MT) = make_float4(dar[0][j], dar[1][j], dar[2][j], dar[3][j]);
            *reinterpret_cast<float4*>(dgs + (1 * HP + j0 + j) * MT) = make_float4(daz[0][j], daz[1][j], daz[2][j], daz[3][j]);
            *reinterpret_cast<float4*>(dgs + (2 * HP + j0 + j) * MT) = make_float4(dah[0][j], dah[1][j], dah[2][j], dah[3][j]);
        }
        __syncthreads();
        if (step + 1 < T) {   // the recurrent term is not needed after the last step
#pragma unroll 4
            for (int c = 0; c < 3 * HP; c++) {
                float4 g4 = *reinterpret_cast<const float4*>(dgs + c * MT);
                float4 w4 = *reinterpret_cast<const float4*>(&W[c * HP + j0]);
                float gv[4] = {g4.x, g4.y, g4.z, g4.w}, wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int m = 0; m < GRU_TM; m++)
#pragma unroll
                    for (int j = 0; j < 4; j++) part[m][j] = fmaf(gv[m], wv[j], part[m][j]);
            }
        }
#pragma unroll
        for (int m = 0; m < GRU_TM; m++)
#pragma unroll
            for (int j = 0; j < 4; j++) dh[m][j] = part[m][j];
        cur ^= 1;
    }
}

// ---- dispatch ---------------------------------------------------------------
template <int HP, int MT>
static int gru_fwd_launch_t(const GruFwdArgs& a, cudaStream_t st) {
    size_t smem = (size_t)(HP * 3 * HP + 3 * HP + 2 * HP * MT) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        DOF_CUDA(cudaFuncSetAttribute(gru_fwd_kernel<HP, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    dim3 grid(cdiv(a.S, MT), 2);
    ProfScope ps(HP == 32 ? "gru_fwd_h32" : HP == 16 ? "gru_fwd_h16" : "gru_fwd_other", st,
                 2.0 * 2.0 * a.S * a.T * 3.0 * a.H * a.H);
    gru_fwd_kernel<HP, MT><<<grid, (MT / GRU_TM) * (HP / 4), smem, st>>>(a);
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}
template <int HP, int MT>
static int gru_bwd_launch_t(const GruBwdArgs& a, cudaStream_t st) {
    size_t smem = (size_t)(3 * HP * HP + 2 * 3 * HP * MT) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        DOF_CUDA(cudaFuncSetAttribute(gru_bwd_kernel<HP, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    dim3 grid(cdiv(a.S, MT), 2);
    ProfScope ps(HP == 32 ? "gru_bwd_h32" : HP == 16 ? "gru_bwd_h16" : "gru_bwd_other", st,
                 2.0 * 2.0 * a.S * a.T * 3.0 * a.H * a.H);
    gru_bwd_kernel<HP, MT><<<grid, (MT / GRU_TM) * (HP / 4), smem, st>>>(a);
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

static int launch_gru_fwd(const GruFwdArgs& a, cudaStream_t st) {
    if (a.S <= 0) return DOF_OK;
    const int H = a.H;
    if (H <= 8) return gru_fwd_launch_t<8, 128>(a, st);
    if (H <= 12) return gru_fwd_launch_t<12, 128>(a, st);
    if (H <= 16) return gru_fwd_launch_t<16, 128>(a, st);
    if (H <= 24) return gru_fwd_launch_t<24, 64>(a, st);
    if (H <= 32) return gru_fwd_launch_t<32, 64>(a, st);
    if (H <= 48) return gru_fwd_launch_t<48, 64>(a, st);
    if (H <= 64) return gru_fwd_launch_t<64, 64>(a, st);
    DOF_FAIL(DOF_ERR_UNSUPPORTED, "GRU hidden size %d > 64 is not supported yet", H);
}
static int launch_gru_bwd(const GruBwdArgs& a, cudaStream_t st) {
    if (a.S <= 0) return DOF_OK;
    const int H = a.H;
    if (H <= 8) return gru_bwd_launch_t<8, 128>(a, st);
    if (H <= 12) return gru_bwd_launch_t<12, 128>(a, st);
    if (H <= 16) return gru_bwd_launch_t<16, 128>(a, st);
    if (H <= 24) return gru_bwd_launch_t<24, 64>(a, st);
    if (H <= 32) return gru_bwd_launch_t<32, 64>(a, st);
    if (H <= 48) return gru_bwd_launch_t<48, 64>(a, st);
    if (H <= 64) return gru_bwd_launch_t<64, 32>(a, st);
    DOF_FAIL(DOF_ERR_UNSUPPORTED, "GRU hidden size %d > 64 is not supported yet", H);
}
