// Orchestration of the transformer-family forward / backward on a dof_handle whose config says
// encoder == DOF_ENCODER_TRANSFORMER.  Included by api.cu after the handle, the GEMM / LayerNorm launchers and the
// CensNet helpers are defined.  Reference: deepof/clustering/models_new.py
//   TFMEncoderPT.forward :1093-1164, TransformerCorePT.forward :955-982, TransformerEncoderLayerPT :893-919,
//   MultiHeadAttentionPT :843-890, TFMDecoderPT.forward :1232-1266, CausalSelfAttentionLayer :1270-1327.
// Every dense product runs through launch_gemm_rows / launch_gemm_wgrad (tcgen05, 3xTF32, where the shape is eligible);
// only the LAST time step of the last encoder layer leaves a core (:982), so that layer computes queries, attention
// output, projections, FFN and both LayerNorms for one row per sequence (keys / values for all rows).
#pragma once

#define TFM_ENC_RATE 0.1f
#define TFM_DEC_RATE 0.2f

// ---- dropout plan: byte offsets of the explicit masks (see dof_set_dropout in the header) and Philox site numbers
struct DropPlan {
    size_t enc[2][1 + 3 * TFM_MAXL];     // per core: embed, then per layer attn | drop1 | drop2
    size_t dec[2][4 * TFM_MAXL];         // per pass, per layer: attn | drop_o | drop_h | drop_f
    size_t total;
};

static DropPlan drop_plan(const dof_config& c, const Layout& L, int Bw, int B, int passes) {
    DropPlan p;
    memset(&p, 0, sizeof(p));
    size_t off = 0;
    const size_t T = (size_t)c.T;
    for (int b = 0; b < 2; b++) {
        const size_t S = (size_t)Bw * (b == 0 ? c.N : c.E);
        p.enc[b][0] = off; off += S * T * L.dk;
        for (int l = 0; l < L.layers; l++) {
            p.enc[b][1 + 3 * l] = off; off += S * L.heads * T * T;
            p.enc[b][2 + 3 * l] = off; off += S * T * L.dk;
            p.enc[b][3 + 3 * l] = off; off += S * T * L.dk;
        }
    }
    const size_t dm = 4 * (size_t)c.D;
    for (int ps = 0; ps < passes && ps < 2; ps++)
        for (int l = 0; l < L.dec_layers; l++) {
            p.dec[ps][4 * l] = off; off += (size_t)B * L.dec_heads * T * T;
            p.dec[ps][4 * l + 1] = off; off += (size_t)B * T * dm;
            p.dec[ps][4 * l + 2] = off; off += (size_t)B * T * L.dec_dff;
            p.dec[ps][4 * l + 3] = off; off += (size_t)B * T * dm;
        }
    p.total = off;
    return p;
}

static DropSite drop_site(const dof_handle* h, bool train, float rate, size_t mask_off, unsigned int site) {
    DropSite d;
    d.keep = (train && h->drop_masks) ? h->drop_masks + mask_off : nullptr;
    d.seed = h->drop_seed; d.site = site; d.rate = train ? rate : 0.f;
    return d;
}

// ---- launch helpers ------------------------------------------------------------------------------------------------
static int tfm_ln_launch(const TfmLnArgs& a, bool bwd, int sm, cudaStream_t st) {
    if (a.R <= 0) return DOF_OK;
    if ((a.W & 3) || a.W > 256) DOF_FAIL(DOF_ERR_UNSUPPORTED, "LayerNorm width %d (multiple of 4, <= 256)", a.W);
    const int lpr = a.W <= 32 ? 8 : a.W <= 64 ? 16 : 32;
    const long long warps = (a.R + (32 / lpr) - 1) / (32 / lpr);
    long long blocks = (warps + 7) / 8;
    const long long cap = (long long)sm * (bwd ? 8 : 16);
    const int grid = (int)(blocks < cap ? blocks : cap);
    ProfScope ps(bwd ? "tfm_ln_bwd" : "tfm_ln_fwd", st, 0.0, (bwd ? 16.0 : 16.0) * a.R * a.W);
    if (!bwd) {
        if (a.W <= 32) tfm_ln_fwd_kernel<8, 1><<<grid, 256, 0, st>>>(a);
        else if (a.W <= 64) tfm_ln_fwd_kernel<16, 1><<<grid, 256, 0, st>>>(a);
        else if (a.W <= 128) tfm_ln_fwd_kernel<32, 1><<<grid, 256, 0, st>>>(a);
        else tfm_ln_fwd_kernel<32, 2><<<grid, 256, 0, st>>>(a);
    } else {
        if (a.W <= 32) tfm_ln_bwd_kernel<8, 1><<<grid, 256, 0, st>>>(a);
        else if (a.W <= 64) tfm_ln_bwd_kernel<16, 1><<<grid, 256, 0, st>>>(a);
        else if (a.W <= 128) tfm_ln_bwd_kernel<32, 1><<<grid, 256, 0, st>>>(a);
        else tfm_ln_bwd_kernel<32, 2><<<grid, 256, 0, st>>>(a);
    }
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

static int tfm_ln_fwd(const float* res, long long rmul, long long radd, const float* branch, float* u, float* out, float* mu, float* rs,
                      const float* w, const float* b, DropSite drop, long long dmul, long long dadd, long long R, int W, int sm,
                      cudaStream_t st) {
    TfmLnArgs a;
    memset(&a, 0, sizeof(a));
    a.res = res; a.branch = branch; a.u = u; a.out = out; a.mu = mu; a.rs = rs; a.w = w; a.b = b; a.drop = drop;
    a.R = R; a.row_mul = rmul; a.row_add = radd; a.drp_mul = dmul; a.drp_add = dadd; a.W = W; a.eps = 1e-6f;
    return tfm_ln_launch(a, false, sm, st);
}

static int tfm_ln_bwd(const float* dy, const float* x, const float* mu, const float* rs, const float* w, float* dx, int dx_accum,
                      float* ddrop, DropSite drop, long long dmul, long long dadd, float* dw, float* db, long long R, int W, int sm,
                      cudaStream_t st) {
    TfmLnArgs a;
    memset(&a, 0, sizeof(a));
    a.dy = dy; a.x = x; a.mu = const_cast<float*>(mu); a.rs = const_cast<float*>(rs); a.w = w; a.dx = dx; a.dx_accum = dx_accum; a.ddrop = ddrop; a.drop = drop;
    a.dw = dw; a.db = db; a.R = R; a.drp_mul = dmul; a.drp_add = dadd; a.row_mul = 1; a.W = W; a.eps = 1e-6f;
    return tfm_ln_launch(a, true, sm, st);
}

static int tfm_ew(int mode, const float* a, const float* b, float* out, DropSite drop, long long n, int sm, cudaStream_t st) {
    if (n <= 0) return DOF_OK;
    if (n & 3) DOF_FAIL(DOF_ERR_UNSUPPORTED, "element-wise length %lld is not a multiple of 4", n);
    TfmEwArgs e;
    e.a = a; e.b = b; e.out = out; e.drop = drop; e.n = n; e.mode = mode;
    const long long blocks = (n / 4 + 255) / 256;
    const int grid = (int)(blocks < (long long)sm * 16 ? blocks : (long long)sm * 16);
    ProfScope ps("tfm_ew", st, 0.0, (mode == 0 || mode == 3 ? 12.0 : 8.0) * n);
    tfm_ew_kernel<<<grid, 256, 0, st>>>(e);
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

// C = act(A . W^T + b) for any K / N: K > 128 is staged in column blocks when the tensor-core kernel takes it
static int tfm_gemm(const float* A, int lda, const float* W, int ldw, int wT, const float* bias, float* C, int ldc, int M, int N, int K,
                    int relu, int accum, const float* mask, int ldmask, cudaStream_t st) {
    if (M <= 0) return DOF_OK;
    for (int n0 = 0; n0 < N; n0 += 256) {
        const int nn = N - n0 < 256 ? N - n0 : 256;
        GemmArgs g = gemm_args(mv_plain(A, lda), wT ? W + n0 : W + (size_t)n0 * ldw, ldw, wT, bias ? bias + n0 : nullptr, C + n0, ldc, M, nn, K);
        g.relu = relu; g.accum = accum;
        if (mask) { g.mask = mask + n0; g.ldmask = ldmask; }
        if (K > 128) {
            for (int ks = 2; ks <= 8; ks++)
                if (K % ks == 0 && K / ks <= 128 && ((K / ks) & 3) == 0) { g.ksplit = ks; break; }
            if (g.ksplit > 1 && !tc_rows_eligible(g)) g.ksplit = 0;
        }
        DOF_TRY(launch_gemm_rows(&g, 1, st));
    }
    return DOF_OK;
}

// dW[N, K] += P^T Q, db[N] += sum P, chunked so that the tensor-core kernel (N <= 128, K <= 252) takes every piece
// nv / kv > 0: only the first nv rows / kv columns of dW exist (P / Q carry zero columns up to N / K, a pitch padded to 4 floats)
static int tfm_wgrad(const float* P, int ldp, const float* Q, int ldq, float* dW, int ldo, int oT, float* db, int M, int N, int K, int sm,
                     cudaStream_t st, int nv = 0, int kv = 0) {
    if (M <= 0) return DOF_OK;
    int nchunk = N <= 128 ? N : 128, kchunk = K <= 252 ? K : 128;
    // the tensor-core kernel prefetches (N/4 + K/4) / 4 float4 column blocks per 8-row group and holds at most 12 of them in its
    // producers' registers: 128 x 128 pieces do not fit (they fell back to the SIMT kernel: 8.5 ms per cfg5 step), 128 x 64 do
    auto fits = [](int nn, int kk) { return cdiv(nn >> 2, 4) + cdiv(kk >> 2, 4) <= (TCW_PW * TCW_MAXPRE) / (TCW_BM / 8); };
    if (!fits(nchunk, kchunk)) {
        if (kchunk > 64 && fits(nchunk, 64)) kchunk = 64;
        else if (nchunk > 64 && fits(64, kchunk)) nchunk = 64;
        else if (nchunk > 64 && kchunk > 64) nchunk = kchunk = 64;
    }
    std::vector<WGradArgs> v;
    for (int n0 = 0; n0 < N; n0 += nchunk)
        for (int k0 = 0; k0 < K; k0 += kchunk) {
            const int nn = N - n0 < nchunk ? N - n0 : nchunk, kk = K - k0 < kchunk ? K - k0 : kchunk;
            float* o = oT ? dW + (size_t)k0 * ldo + n0 : dW + (size_t)n0 * ldo + k0;
            WGradArgs wa = wgrad_args(mv_plain(P + n0, ldp), mv_plain(Q + k0, ldq), o, ldo, oT, (db && k0 == 0) ? db + n0 : nullptr, M, nn, kk);
            if (nv > 0) { wa.nv = nv - n0 < nn ? nv - n0 : nn; if (wa.nv <= 0) continue; }
            if (kv > 0) { wa.kv = kv - k0 < kk ? kv - k0 : kk; if (wa.kv <= 0) continue; }
            v.push_back(wa);
        }
    size_t i = 0;
    while (i < v.size()) {                                   // batch runs of equal (N, K)
        size_t j = i + 1;
        while (j < v.size() && j - i < WG_MAXBATCH && v[j].N == v[i].N && v[j].K == v[i].K) j++;
        DOF_TRY(launch_gemm_wgrad(v.data() + i, (int)(j - i), st, sm));
        i = j;
    }
    return DOF_OK;
}

template <int HDP>
static int tfm_attn2_launch(const TfmAttnArgs& a, bool bwd, cudaStream_t st) {
    const size_t smem = tfm_attn2_smem(a.T, a.heads, HDP, bwd);
    static size_t attr[2] = {48 * 1024, 48 * 1024};
    if (smem > attr[bwd ? 1 : 0]) {
        if (bwd) DOF_CUDA(cudaFuncSetAttribute(tfm_attn2_kernel<HDP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else DOF_CUDA(cudaFuncSetAttribute(tfm_attn2_kernel<HDP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr[bwd ? 1 : 0] = smem;
    }
    const int threads = 32 * (a.heads < 8 ? a.heads : 8);
    if (bwd) tfm_attn2_kernel<HDP, true><<<a.S, threads, smem, st>>>(a);
    else tfm_attn2_kernel<HDP, false><<<a.S, threads, smem, st>>>(a);
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

static int tfm_attention(const float* qkv, const unsigned char* kpad, DropSite drop, int causal, int S, int T, int dm, int heads,
                         int q_from, float* out, const float* dout, float* dqkv, cudaStream_t st) {
    const bool bwd = dout != nullptr;
    TfmAttnArgs a;
    a.qkv = qkv; a.kpad = kpad; a.drop = drop; a.out = out; a.dout = dout; a.dqkv = dqkv; a.S = S; a.T = T; a.dm = dm; a.heads = heads;
    a.causal = causal; a.q_from = q_from;
    const int TQ = T - q_from, hd = dm / heads;
    const double fl = (bwd ? 2.5 : 1.0) * 4.0 * (double)S * TQ * T * dm / (causal ? 2.0 : 1.0);
    const double by = 4.0 * (double)S * (bwd ? (6.0 * T * dm + TQ * dm) : (3.0 * T * dm + TQ * dm));
    // all queries, T <= 32: lane-per-query kernel; otherwise (last-step-only layers, long windows) the warp-per-query kernel
    const int hdp = hd <= 4 ? 4 : hd <= 8 ? 8 : hd <= 12 ? 12 : hd <= 16 ? 16 : hd <= 32 ? 32 : 0;
    if (q_from == 0 && T <= TFM2_MAXT && hdp && tfm_attn2_smem(T, heads, hdp, bwd) <= 200 * 1024) {
        ProfScope ps(bwd ? "tfm_attn_bwd" : "tfm_attn_fwd", st, fl, by);
        switch (hdp) {
            case 4: return tfm_attn2_launch<4>(a, bwd, st);
            case 8: return tfm_attn2_launch<8>(a, bwd, st);
            case 12: return tfm_attn2_launch<12>(a, bwd, st);
            case 16: return tfm_attn2_launch<16>(a, bwd, st);
            default: return tfm_attn2_launch<32>(a, bwd, st);
        }
    }
    const size_t smem = ((size_t)T * (3 * dm + 1) * (bwd ? 2 : 1) + (bwd ? (size_t)T * (dm + 1) : 0)) * 4;
    if (smem > 200 * 1024) DOF_FAIL(DOF_ERR_UNSUPPORTED, "sequence of %d x %d does not fit shared memory", T, 3 * dm);
    static bool attr = false;
    if (!attr) {
        DOF_CUDA(cudaFuncSetAttribute(tfm_attn_train_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        DOF_CUDA(cudaFuncSetAttribute(tfm_attn_train_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    if (bwd) {
        ProfScope ps("tfm_attn_last_bwd", st, fl, by);
        tfm_attn_train_kernel<true><<<S, 128, smem, st>>>(a);
    } else {
        ProfScope ps("tfm_attn_last_fwd", st, fl, by);
        tfm_attn_train_kernel<false><<<S, 128, smem, st>>>(a);
    }
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

// ---- one transformer core (node or edge sequences) -----------------------------------------------------------------
static int tfm_core_forward(dof_handle* h, int b, const float* state, const float* xin, int Bw, bool train, const DropPlan& dp,
                            cudaStream_t st) {
    const dof_config& c = h->cfg;
    const Layout& L = h->L;
    TfmCoreWS& w = h->tc[b];
    const TfmCoreP& P = L.tcore[b];
    const int T = c.T, dk = L.dk, dff = L.dff, S = Bw * w.G, sm = h->sm_count;
    const long long R = (long long)S * T;
    const unsigned int site0 = 1 + 32 * b;
    TfmEmbedArgs ea;
    memset(&ea, 0, sizeof(ea));
    ea.x = xin; ea.gidx = w.gidx; ea.We = state + P.We; ea.be = state + P.be; ea.pe = h->pe_enc; ea.Xs = w.Xs; ea.kpad = w.kpad; ea.Y0 = w.Y0;
    ea.drop = drop_site(h, train, TFM_ENC_RATE, dp.enc[b][0], site0);
    ea.B = Bw; ea.T = T; ea.G = w.G; ea.F = w.Fin; ea.dk = dk;
    {
        const long long blocks = (R * (dk / 4) + 255) / 256;
        ProfScope ps("tfm_embed_fwd", st, 2.0 * R * dk * w.Fin, 4.0 * R * (w.Fin + dk));
        tfm_embed_fwd_kernel<<<(int)(blocks < (long long)sm * 16 ? blocks : (long long)sm * 16), 256, 0, st>>>(ea);
    }
    DOF_LAUNCH_CHECK();
    const float* Yin = w.Y0;
    for (int l = 0; l < L.layers; l++) {
        const bool last = l == L.layers - 1;
        const TfmLayerWS& q = w.l[l];
        const TfmLayerP& Q = P.l[l];
        const long long Rl = last ? S : R, mul = last ? T : 1, add = last ? T - 1 : 0;
        DOF_TRY(tfm_gemm(Yin, dk, state + Q.Wqkv, dk, 0, nullptr, q.QKV, 3 * dk, (int)R, 3 * dk, dk, 0, 0, nullptr, 0, st));
        DOF_TRY(tfm_attention(q.QKV, w.kpad, drop_site(h, train, TFM_ENC_RATE, dp.enc[b][1 + 3 * l], site0 + 1 + 3 * l), 0, S, T, dk,
                              L.heads, last ? T - 1 : 0, q.ATT, nullptr, nullptr, st));
        DOF_TRY(tfm_gemm(q.ATT, dk, state + Q.Wo, dk, 0, nullptr, q.U1, dk, (int)Rl, dk, dk, 0, 0, nullptr, 0, st));
        DOF_TRY(tfm_ln_fwd(Yin, mul, add, q.U1, q.U1, q.Y1, q.mu1, q.rs1, state + Q.n1w, state + Q.n1b,
                           drop_site(h, train, TFM_ENC_RATE, dp.enc[b][2 + 3 * l], site0 + 2 + 3 * l), mul, add, Rl, dk, sm, st));
        DOF_TRY(tfm_gemm(q.Y1, dk, state + Q.W1, dk, 0, state + Q.b1, q.FF, dff, (int)Rl, dff, dk, 1, 0, nullptr, 0, st));
        DOF_TRY(tfm_gemm(q.FF, dff, state + Q.W2, dff, 0, state + Q.b2, q.U2, dk, (int)Rl, dk, dff, 0, 0, nullptr, 0, st));
        DOF_TRY(tfm_ln_fwd(q.Y1, 1, 0, q.U2, q.U2, q.Y2, q.mu2, q.rs2, state + Q.n2w, state + Q.n2b,
                           drop_site(h, train, TFM_ENC_RATE, dp.enc[b][3 + 3 * l], site0 + 3 + 3 * l), mul, add, Rl, dk, sm, st));
        Yin = q.Y2;
    }
    return DOF_OK;
}

// backward of one core from w.dOut = d(loss)/d(core output) [S, dk]
static int tfm_core_backward(dof_handle* h, int b, const float* state, float* grad, int Bw, const DropPlan& dp, cudaStream_t st) {
    const dof_config& c = h->cfg;
    const Layout& L = h->L;
    TfmCoreWS& w = h->tc[b];
    const TfmCoreP& P = L.tcore[b];
    const int T = c.T, dk = L.dk, dff = L.dff, S = Bw * w.G, sm = h->sm_count;
    const long long R = (long long)S * T;
    const unsigned int site0 = 1 + 32 * b;
    for (int l = L.layers - 1; l >= 0; l--) {
        const bool last = l == L.layers - 1;
        const TfmLayerWS& q = w.l[l];
        const TfmLayerP& Q = P.l[l];
        const float* Yin = l == 0 ? w.Y0 : w.l[l - 1].Y2;
        const long long Rl = last ? S : R, mul = last ? T : 1, add = last ? T - 1 : 0;
        const float* dYout = last ? w.dOut : w.dA;
        float* bufX = last ? w.sA : w.dB;       // d Y1, then d ATT
        float* bufD = last ? w.sB : w.dC;       // gradients behind a dropout
        float* bufF = last ? w.sFF : w.dFF;
        float* bufR = last ? w.sC : w.dA;       // residual gradient into the layer input
        DOF_TRY(tfm_ln_bwd(dYout, q.U2, q.mu2, q.rs2, state + Q.n2w, bufX, 0, bufD,
                           drop_site(h, true, TFM_ENC_RATE, dp.enc[b][3 + 3 * l], site0 + 3 + 3 * l), mul, add, grad + Q.n2w, grad + Q.n2b,
                           Rl, dk, sm, st));
        DOF_TRY(tfm_wgrad(bufD, dk, q.FF, dff, grad + Q.W2, dff, 0, grad + Q.b2, (int)Rl, dk, dff, sm, st));
        DOF_TRY(tfm_gemm(bufD, dk, state + Q.W2, dff, 1, nullptr, bufF, dff, (int)Rl, dff, dk, 0, 0, q.FF, dff, st));
        DOF_TRY(tfm_wgrad(bufF, dff, q.Y1, dk, grad + Q.W1, dk, 0, grad + Q.b1, (int)Rl, dff, dk, sm, st));
        DOF_TRY(tfm_gemm(bufF, dff, state + Q.W1, dk, 1, nullptr, bufX, dk, (int)Rl, dk, dff, 0, 1, nullptr, 0, st));
        DOF_TRY(tfm_ln_bwd(bufX, q.U1, q.mu1, q.rs1, state + Q.n1w, bufR, 0, bufD,
                           drop_site(h, true, TFM_ENC_RATE, dp.enc[b][2 + 3 * l], site0 + 2 + 3 * l), mul, add, grad + Q.n1w, grad + Q.n1b,
                           Rl, dk, sm, st));
        DOF_TRY(tfm_wgrad(bufD, dk, q.ATT, dk, grad + Q.Wo, dk, 0, nullptr, (int)Rl, dk, dk, sm, st));
        DOF_TRY(tfm_gemm(bufD, dk, state + Q.Wo, dk, 1, nullptr, bufX, dk, (int)Rl, dk, dk, 0, 0, nullptr, 0, st));
        DOF_TRY(tfm_attention(q.QKV, w.kpad, drop_site(h, true, TFM_ENC_RATE, dp.enc[b][1 + 3 * l], site0 + 1 + 3 * l), 0, S, T, dk,
                              L.heads, last ? T - 1 : 0, nullptr, bufX, w.dQKV, st));
        for (int m = 0; m < 3; m++)       // q | k | v weights are consecutive [dk, dk] blocks
            DOF_TRY(tfm_wgrad(w.dQKV + m * dk, 3 * dk, Yin, dk, grad + Q.Wqkv + (size_t)m * dk * dk, dk, 0, nullptr, (int)R, dk, dk, sm, st));
        DOF_TRY(tfm_gemm(w.dQKV, 3 * dk, state + Q.Wqkv, dk, 1, nullptr, w.dA, dk, (int)R, dk, 3 * dk, 0, last ? 0 : 1, nullptr, 0, st));
        if (last) {
            ProfScope ps("tfm_row_scatter_add", st);
            tfm_row_scatter_add_kernel<<<cdiv((long long)S * dk, 256), 256, 0, st>>>(w.sC, w.dA, S, dk, T, T - 1);
            DOF_LAUNCH_CHECK();
        }
    }
    TfmEmbedArgs ea;
    memset(&ea, 0, sizeof(ea));
    ea.We = state + P.We; ea.be = state + P.be; ea.Xs = w.Xs; ea.dY0 = w.dA; ea.dWe = grad + P.We; ea.dbe = grad + P.be;
    ea.drop = drop_site(h, true, TFM_ENC_RATE, dp.enc[b][0], site0);
    ea.B = Bw; ea.T = T; ea.G = w.G; ea.F = w.Fin; ea.dk = dk;
    {
        const int rows_per = 256 / (dk / 4);
        const long long blocks = (R + rows_per - 1) / rows_per;
        ProfScope ps("tfm_embed_bwd", st, 2.0 * R * dk * w.Fin, 4.0 * R * (w.Fin + dk));
        tfm_embed_bwd_kernel<<<(int)(blocks < (long long)sm * 8 ? blocks : (long long)sm * 8), 256, (size_t)dk * (w.Fin + 1) * 4, st>>>(ea);
    }
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

// ---- encoder head helpers --------------------------------------------------------------------------------------------
static int tfm_col(dof_handle* h, bool bwd, int kind, const float* x, float* y, const float* w, const float* b, const float* rmean,
                   const float* rvar, int slot, float* stat_out, const float* dy, float* dx, float* dw, float* db, const float* relu_ref,
                   int B, int C, int groups, cudaStream_t st) {
    TfmColArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x; a.y = y; a.w = w; a.b = b; a.run_mean = rmean; a.run_var = rvar; a.mean = h->bnm[slot]; a.scale = h->bns[slot];
    a.stat_out = stat_out; a.clamped = h->bclamp; a.dy = dy; a.dx = dx; a.dw = dw; a.db = db; a.relu_ref = relu_ref;
    a.B = B; a.C = C; a.groups = groups; a.kind = kind; a.eps = 1e-3f;
    dim3 grid(cdiv(C, 32), groups), block(32, 32);
    ProfScope ps(bwd ? "tfm_col_bwd" : "tfm_col_fwd", st);
    if (bwd) tfm_col_bwd_kernel<<<grid, block, 0, st>>>(a);
    else tfm_col_fwd_kernel<<<grid, block, 0, st>>>(a);
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

static CensArgs tfm_cens_args(dof_handle* h, const float* state, const float* node, const float* edge, int B) {
    const dof_config& c = h->cfg;
    const Layout& L = h->L;
    CensArgs a;
    memset(&a, 0, sizeof(a));
    a.node = node; a.edge = edge;
    a.lap = state + L.lap; a.elap = state + L.elap; a.inc = state + L.inc;
    a.wn = state + L.node_weights; a.we = state + L.edge_weights;
    a.Pn = h->Pn; a.Pe = h->Pe;
    a.B = B; a.N = c.N; a.E = c.E; a.C = L.dk;
    return a;
}

static int enc_tail_forward(dof_handle* h, const float* state, const float* node, const float* edge, int Bw, bool train, int groups,
                            bool standardise, cudaStream_t st);
// TFMEncoderPT.forward: Bw windows -> h->enc [Bw, D].  `groups`: row ranges with separate batch statistics (train only).
static int tfm_encoder_forward(dof_handle* h, const float* state, const float* x, const float* a, int Bw, bool train, int groups,
                               cudaStream_t st) {
    const dof_config& c = h->cfg;
    const Layout& L = h->L;
    if (groups < 1 || Bw % groups) DOF_FAIL(DOF_ERR_ARG, "batch %d is not a multiple of %d statistics groups", Bw, groups);
    if (train && Bw / groups < 2) DOF_FAIL(DOF_ERR_UNSUPPORTED, "train-mode BatchNorm needs at least 2 windows per pass");
    const DropPlan dp = drop_plan(c, L, Bw, 0, 0);
    if (train && h->drop_masks && h->drop_mask_bytes < dp.total) DOF_FAIL(DOF_ERR_ARG, "dropout masks: %zu bytes < %zu", h->drop_mask_bytes, dp.total);
    DOF_TRY(fork_join_blocks(h, st, [&](int b, cudaStream_t s) { return tfm_core_forward(h, b, state, b == 0 ? x : a, Bw, train, dp, s); }));
    return enc_tail_forward(h, state, h->tc[0].l[L.layers - 1].Y2, h->tc[1].l[L.layers - 1].Y2, Bw, train, groups, true, st);
}

// What follows the per-node / per-edge sequence models in TFMEncoderPT (:1123-1164) AND TCNEncoderPT (:630-657): CensNet, ReLU,
// per-sample RMS normalisation, the head MLP with two BatchNorm layers, and (transformer only) the batch standardisation.
// node [Bw * N, L.dk], edge [Bw * E, L.dk] -> h->enc [Bw, D]
static int enc_tail_forward(dof_handle* h, const float* state, const float* node, const float* edge, int Bw, bool train, int groups,
                            bool standardise, cudaStream_t st) {
    const dof_config& c = h->cfg;
    const Layout& L = h->L;
    const int N = c.N, E = c.E, D = c.D, dk = L.dk, KD = (N + E) * D;
    CensArgs ca = tfm_cens_args(h, state, node, edge, Bw);
    const size_t smem = cens_smem_floats(N, E, dk) * 4;
    if (smem > 200 * 1024) DOF_FAIL(DOF_ERR_UNSUPPORTED, "graph too large for the CensNet kernel (%zu B smem)", smem);
    static bool attr = false;
    if (!attr) {
        DOF_CUDA(cudaFuncSetAttribute(cens_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        DOF_CUDA(cudaFuncSetAttribute(cens_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr = true;
    }
    { ProfScope ps("cens_fwd", st);
    cens_fwd_kernel<<<Bw, 128, smem, st>>>(ca); }
    DOF_LAUNCH_CHECK();
    DOF_TRY(tfm_gemm(h->Pn, dk, state + L.node_kernel, D, 1, state + L.node_bias, h->On, D, Bw * N, D, dk, 1, 0, nullptr, 0, st));
    DOF_TRY(tfm_gemm(h->Pe, dk, state + L.edge_kernel, D, 1, state + L.edge_bias, h->Oe, D, Bw * E, D, dk, 1, 0, nullptr, 0, st));
    TfmRmsArgs ra;
    memset(&ra, 0, sizeof(ra));
    ra.on = h->On; ra.oe = h->Oe; ra.h = h->hIn; ra.rms = h->hRms; ra.B = Bw; ra.ND = N * D; ra.ED = E * D;
    { ProfScope ps("tfm_rms_fwd", st);
    tfm_rms_fwd_kernel<<<cdiv((long long)Bw * 32, 256), 256, 0, st>>>(ra); }
    DOF_LAUNCH_CHECK();
    const int kind = train ? 0 : 1;
    DOF_TRY(tfm_gemm(h->hIn, KD, state + L.h_w0, KD, 0, state + L.h_b0, h->h1r, 2 * D, Bw, 2 * D, KD, 1, 0, nullptr, 0, st));
    DOF_TRY(tfm_col(h, false, kind, h->h1r, h->h1, state + L.bn2.w, state + L.bn2.b, state + L.bn2.mean, state + L.bn2.var, 0, h->bnstat[0],
                    nullptr, nullptr, nullptr, nullptr, nullptr, Bw, 2 * D, groups, st));
    DOF_TRY(tfm_gemm(h->h1, 2 * D, state + L.h_w3, 2 * D, 0, state + L.h_b3, h->h2r, D, Bw, D, 2 * D, 1, 0, nullptr, 0, st));
    DOF_TRY(tfm_col(h, false, kind, h->h2r, h->h2, state + L.bn5.w, state + L.bn5.b, state + L.bn5.mean, state + L.bn5.var, 1, h->bnstat[1],
                    nullptr, nullptr, nullptr, nullptr, nullptr, Bw, D, groups, st));
    DOF_TRY(tfm_gemm(h->h2, D, state + L.h_w6, D, 0, state + L.h_b6, (train && standardise) ? h->h3 : h->enc, D, Bw, D, D, 0, 0, nullptr, 0, st));
    if (train && standardise)      // batch standardisation (:1161-1162)
        DOF_TRY(tfm_col(h, false, 2, h->h3, h->enc, nullptr, nullptr, nullptr, nullptr, 2, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                        Bw, D, groups, st));
    h->enc_groups = groups;
    if (train) h->bn_pending = 1;
    return DOF_OK;
}

// backward of the transformer encoder from h->denc [Bw, D]
static int enc_tail_backward(dof_handle* h, const float* state, float* grad, const float* node, const float* edge, float* dnode, float* dedge,
                             int Bw, bool standardise, cudaStream_t st);
static int tfm_encoder_backward(dof_handle* h, const float* state, float* grad, int Bw, cudaStream_t st) {
    const dof_config& c = h->cfg;
    const Layout& L = h->L;
    const DropPlan dp = drop_plan(c, L, Bw, 0, 0);
    DOF_TRY(enc_tail_backward(h, state, grad, h->tc[0].l[L.layers - 1].Y2, h->tc[1].l[L.layers - 1].Y2, h->tc[0].dOut, h->tc[1].dOut, Bw, true, st));
    DOF_TRY(fork_join_blocks(h, st, [&](int b, cudaStream_t s) { return tfm_core_backward(h, b, state, grad, Bw, dp, s); }));
    return DOF_OK;
}

// backward of enc_tail_forward from h->denc [Bw, D]: parameter gradients of the head and CensNet, dnode [Bw * N, dk], dedge
static int enc_tail_backward(dof_handle* h, const float* state, float* grad, const float* node, const float* edge, float* dnode, float* dedge,
                             int Bw, bool standardise, cudaStream_t st) {
    const dof_config& c = h->cfg;
    const Layout& L = h->L;
    const int N = c.N, E = c.E, D = c.D, dk = L.dk, KD = (N + E) * D, sm = h->sm_count, groups = h->enc_groups;
    const float* dh3 = h->denc;
    if (standardise) {
        DOF_TRY(tfm_col(h, true, 2, h->h3, nullptr, nullptr, nullptr, nullptr, nullptr, 2, nullptr, h->denc, h->dh3, nullptr, nullptr, nullptr, Bw, D,
                        groups, st));
        dh3 = h->dh3;
    }
    DOF_TRY(tfm_wgrad(dh3, D, h->h2, D, grad + L.h_w6, D, 0, grad + L.h_b6, Bw, D, D, sm, st));
    DOF_TRY(tfm_gemm(dh3, D, state + L.h_w6, D, 1, nullptr, h->dh2, D, Bw, D, D, 0, 0, nullptr, 0, st));
    DOF_TRY(tfm_col(h, true, 0, h->h2r, nullptr, state + L.bn5.w, nullptr, nullptr, nullptr, 1, nullptr, h->dh2, h->dh2r, grad + L.bn5.w,
                    grad + L.bn5.b, h->h2r, Bw, D, groups, st));
    DOF_TRY(tfm_wgrad(h->dh2r, D, h->h1, 2 * D, grad + L.h_w3, 2 * D, 0, grad + L.h_b3, Bw, D, 2 * D, sm, st));
    DOF_TRY(tfm_gemm(h->dh2r, D, state + L.h_w3, 2 * D, 1, nullptr, h->dh1, 2 * D, Bw, 2 * D, D, 0, 0, nullptr, 0, st));
    DOF_TRY(tfm_col(h, true, 0, h->h1r, nullptr, state + L.bn2.w, nullptr, nullptr, nullptr, 0, nullptr, h->dh1, h->dh1r, grad + L.bn2.w,
                    grad + L.bn2.b, h->h1r, Bw, 2 * D, groups, st));
    DOF_TRY(tfm_wgrad(h->dh1r, 2 * D, h->hIn, KD, grad + L.h_w0, KD, 0, grad + L.h_b0, Bw, 2 * D, KD, sm, st));
    DOF_TRY(tfm_gemm(h->dh1r, 2 * D, state + L.h_w0, KD, 1, nullptr, h->dhIn, KD, Bw, KD, 2 * D, 0, 0, nullptr, 0, st));
    TfmRmsArgs ra;
    memset(&ra, 0, sizeof(ra));
    ra.on = h->On; ra.oe = h->Oe; ra.h = h->hIn; ra.rms = h->hRms; ra.dh = h->dhIn; ra.don = h->dOn; ra.doe = h->dOe;
    ra.B = Bw; ra.ND = N * D; ra.ED = E * D;
    { ProfScope ps("tfm_rms_bwd", st);
    tfm_rms_bwd_kernel<<<cdiv((long long)Bw * 32, 256), 256, 0, st>>>(ra); }
    DOF_LAUNCH_CHECK();
    // CensNet dense kernels: O = relu(P . kernel + bias), kernel [dk, D]
    DOF_TRY(tfm_wgrad(h->dOn, D, h->Pn, dk, grad + L.node_kernel, D, 1, grad + L.node_bias, Bw * N, D, dk, sm, st));
    DOF_TRY(tfm_wgrad(h->dOe, D, h->Pe, dk, grad + L.edge_kernel, D, 1, grad + L.edge_bias, Bw * E, D, dk, sm, st));
    DOF_TRY(tfm_gemm(h->dOn, D, state + L.node_kernel, D, 0, nullptr, h->dPn, dk, Bw * N, dk, D, 0, 0, nullptr, 0, st));
    DOF_TRY(tfm_gemm(h->dOe, D, state + L.edge_kernel, D, 0, nullptr, h->dPe, dk, Bw * E, dk, D, 0, 0, nullptr, 0, st));
    CensArgs ca = tfm_cens_args(h, state, node, edge, Bw);
    ca.dPn = h->dPn; ca.dPe = h->dPe; ca.dnode = dnode; ca.dedge = dedge;
    ca.dwn = grad + L.node_weights; ca.dwe = grad + L.edge_weights;
    const size_t smem = (cens_smem_floats(N, E, dk) + (size_t)(N + E) * dk + (size_t)N * N + (size_t)E * E + N + E) * 4;
    if (smem > 220 * 1024) DOF_FAIL(DOF_ERR_UNSUPPORTED, "graph too large for the CensNet backward kernel");
    { ProfScope ps("cens_bwd", st);
    cens_bwd_kernel<<<Bw, 128, smem, st>>>(ca); }
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

// ---- decoder -------------------------------------------------------------------------------------------------------------
// TFMDecoderPT.forward: zin [B, D] -> h->loc [B, T, Dx].  `pass` selects the dropout sites (VQ-VAE decodes twice).
static int tfm_decoder_forward(dof_handle* h, const float* state, const float* zin, int B, bool train, int pass, cudaStream_t st) {
    const dof_config& c = h->cfg;
    const Layout& L = h->L;
    TfmDecWS& w = h->td;
    const int T = c.T, D = c.D, dm = 4 * D, dff = L.dec_dff, Dx = c.N * c.F, sm = h->sm_count;
    const long long R = (long long)B * T;
    const DropPlan dp = drop_plan(c, L, B, B, pass + 1);
    if (train && h->drop_masks && h->drop_mask_bytes < dp.total) DOF_FAIL(DOF_ERR_ARG, "dropout masks: %zu bytes < %zu", h->drop_mask_bytes, dp.total);
    const unsigned int site0 = 100 + 32 * pass;
    const int ein[3] = {D, D, 2 * D}, eout[3] = {D, 2 * D, 4 * D};
    const float* cur = zin;
    for (int i = 0; i < 3; i++) {                                      // latent expansion (:1192-1199)
        DOF_TRY(tfm_gemm(cur, ein[i], state + L.dex_w[i], ein[i], 0, state + L.dex_b[i], w.P[i], eout[i], B, eout[i], ein[i], 0, 0, nullptr, 0, st));
        { ProfScope ps("tfm_gelu_fwd", st);
        tfm_gelu_fwd_kernel<<<cdiv((long long)B * eout[i], 256), 256, 0, st>>>(w.P[i], w.G[i], (long long)B * eout[i]); }
        DOF_LAUNCH_CHECK();
        cur = w.G[i];
    }
    { ProfScope ps("tfm_dec_input", st);
    tfm_dec_input_kernel<<<cdiv(R * dm, 256), 256, 0, st>>>(cur, w.H0, B, T, dm); }
    DOF_LAUNCH_CHECK();
    const float* Hin = w.H0;
    for (int l = 0; l < L.dec_layers; l++) {
        const TfmDecLayerWS& q = w.l[l];
        const TfmLayerP& Q = L.tdl[l];
        const size_t* mo = dp.dec[pass] + 4 * l;
        DOF_TRY(tfm_ln_fwd(nullptr, 1, 0, Hin, nullptr, q.Xn1, q.muA, q.rsA, state + Q.n1w, state + Q.n1b, drop_none(), 1, 0, R, dm, sm, st));
        for (int m = 0; m < 3; m++)
            DOF_TRY(tfm_gemm(q.Xn1, dm, state + Q.Wqkv + (size_t)m * dm * dm, dm, 0, nullptr, q.QKV + m * dm, 3 * dm, (int)R, dm, dm, 0, 0, nullptr, 0, st));
        DOF_TRY(tfm_attention(q.QKV, nullptr, drop_site(h, train, TFM_DEC_RATE, mo[0], site0 + 4 * l), 1, B, T, dm, L.dec_heads, 0, q.ATT,
                              nullptr, nullptr, st));
        DOF_TRY(tfm_gemm(q.ATT, dm, state + Q.Wo, dm, 0, nullptr, w.tmpA, dm, (int)R, dm, dm, 0, 0, nullptr, 0, st));
        DOF_TRY(tfm_ew(0, Hin, w.tmpA, q.Hmid, drop_site(h, train, TFM_DEC_RATE, mo[1], site0 + 4 * l + 1), R * dm, sm, st));
        DOF_TRY(tfm_ln_fwd(nullptr, 1, 0, q.Hmid, nullptr, q.Xn2, q.muB, q.rsB, state + Q.n2w, state + Q.n2b, drop_none(), 1, 0, R, dm, sm, st));
        DOF_TRY(tfm_gemm(q.Xn2, dm, state + Q.W1, dm, 0, state + Q.b1, q.PRE, dff, (int)R, dff, dm, 0, 0, nullptr, 0, st));
        DOF_TRY(tfm_ew(2, q.PRE, nullptr, q.FH, drop_site(h, train, TFM_DEC_RATE, mo[2], site0 + 4 * l + 2), R * dff, sm, st));
        DOF_TRY(tfm_gemm(q.FH, dff, state + Q.W2, dff, 0, state + Q.b2, w.tmpA, dm, (int)R, dm, dff, 0, 0, nullptr, 0, st));
        DOF_TRY(tfm_ew(0, q.Hmid, w.tmpA, q.Hout, drop_site(h, train, TFM_DEC_RATE, mo[3], site0 + 4 * l + 3), R * dm, sm, st));
        Hin = q.Hout;
    }
    const int DxP = round_up(Dx, 4);       // pitch of Y / dY: zero pad columns keep their weight gradients on the tensor-core kernel
    if (train && DxP != Dx) DOF_CUDA(cudaMemsetAsync(w.Y, 0, (size_t)R * DxP * 4, st));
    DOF_TRY(tfm_gemm(Hin, dm, state + L.dout_w, dm, 0, state + L.dout_b, w.Y, DxP, (int)R, Dx, dm, 0, 0, nullptr, 0, st));
    DOF_TRY(tfm_gemm(w.Y, DxP, state + L.loc_w, Dx, 0, state + L.loc_b, h->loc, Dx, (int)R, Dx, Dx, 0, 0, nullptr, 0, st));
    return DOF_OK;
}

// backward of the transformer decoder from h->dloc; writes d(loss)/d(zin) into h->dz_dec
static int tfm_decoder_backward(dof_handle* h, const float* state, float* grad, const float* zin, int B, int pass, cudaStream_t st) {
    const dof_config& c = h->cfg;
    const Layout& L = h->L;
    TfmDecWS& w = h->td;
    const int T = c.T, D = c.D, dm = 4 * D, dff = L.dec_dff, Dx = c.N * c.F, sm = h->sm_count;
    const long long R = (long long)B * T;
    const DropPlan dp = drop_plan(c, L, B, B, pass + 1);
    const unsigned int site0 = 100 + 32 * pass;
    const float* Hlast = w.l[L.dec_layers - 1].Hout;
    const int DxP = round_up(Dx, 4);
    const float* dlp = padded_dloc(h, R, st);                       // [R, DxP], zero pad columns
    if (DxP != Dx) DOF_CUDA(cudaMemsetAsync(w.dY, 0, (size_t)R * DxP * 4, st));
    DOF_TRY(tfm_wgrad(dlp, DxP, w.Y, DxP, grad + L.loc_w, Dx, 0, grad + L.loc_b, (int)R, DxP, DxP, sm, st, Dx, Dx));
    DOF_TRY(tfm_gemm(h->dloc, Dx, state + L.loc_w, Dx, 1, nullptr, w.dY, DxP, (int)R, Dx, Dx, 0, 0, nullptr, 0, st));
    DOF_TRY(tfm_wgrad(w.dY, DxP, Hlast, dm, grad + L.dout_w, dm, 0, grad + L.dout_b, (int)R, DxP, dm, sm, st, Dx, 0));
    DOF_TRY(tfm_gemm(w.dY, DxP, state + L.dout_w, dm, 1, nullptr, w.dH, dm, (int)R, dm, Dx, 0, 0, nullptr, 0, st));
    for (int l = L.dec_layers - 1; l >= 0; l--) {
        const TfmDecLayerWS& q = w.l[l];
        const TfmLayerP& Q = L.tdl[l];
        const size_t* mo = dp.dec[pass] + 4 * l;
        const float* Hin = l == 0 ? w.H0 : w.l[l - 1].Hout;
        // FFN branch
        DOF_TRY(tfm_ew(1, w.dH, nullptr, w.tmpA, drop_site(h, true, TFM_DEC_RATE, mo[3], site0 + 4 * l + 3), R * dm, sm, st));
        DOF_TRY(tfm_wgrad(w.tmpA, dm, q.FH, dff, grad + Q.W2, dff, 0, grad + Q.b2, (int)R, dm, dff, sm, st));
        DOF_TRY(tfm_gemm(w.tmpA, dm, state + Q.W2, dff, 1, nullptr, w.tmpF, dff, (int)R, dff, dm, 0, 0, nullptr, 0, st));
        DOF_TRY(tfm_ew(3, w.tmpF, q.PRE, w.tmpF, drop_site(h, true, TFM_DEC_RATE, mo[2], site0 + 4 * l + 2), R * dff, sm, st));
        DOF_TRY(tfm_wgrad(w.tmpF, dff, q.Xn2, dm, grad + Q.W1, dm, 0, grad + Q.b1, (int)R, dff, dm, sm, st));
        DOF_TRY(tfm_gemm(w.tmpF, dff, state + Q.W1, dm, 1, nullptr, w.tmpA, dm, (int)R, dm, dff, 0, 0, nullptr, 0, st));
        DOF_TRY(tfm_ln_bwd(w.tmpA, q.Hmid, q.muB, q.rsB, state + Q.n2w, w.dH, 1, nullptr, drop_none(), 1, 0, grad + Q.n2w, grad + Q.n2b, R, dm,
                           sm, st));
        // attention branch
        DOF_TRY(tfm_ew(1, w.dH, nullptr, w.tmpA, drop_site(h, true, TFM_DEC_RATE, mo[1], site0 + 4 * l + 1), R * dm, sm, st));
        DOF_TRY(tfm_wgrad(w.tmpA, dm, q.ATT, dm, grad + Q.Wo, dm, 0, nullptr, (int)R, dm, dm, sm, st));
        DOF_TRY(tfm_gemm(w.tmpA, dm, state + Q.Wo, dm, 1, nullptr, w.tmpB, dm, (int)R, dm, dm, 0, 0, nullptr, 0, st));
        DOF_TRY(tfm_attention(q.QKV, nullptr, drop_site(h, true, TFM_DEC_RATE, mo[0], site0 + 4 * l), 1, B, T, dm, L.dec_heads, 0, nullptr,
                              w.tmpB, w.dQKV, st));
        for (int m = 0; m < 3; m++) {
            DOF_TRY(tfm_wgrad(w.dQKV + m * dm, 3 * dm, q.Xn1, dm, grad + Q.Wqkv + (size_t)m * dm * dm, dm, 0, nullptr, (int)R, dm, dm, sm, st));
            DOF_TRY(tfm_gemm(w.dQKV + m * dm, 3 * dm, state + Q.Wqkv + (size_t)m * dm * dm, dm, 1, nullptr, w.tmpA, dm, (int)R, dm, dm, 0, m > 0,
                             nullptr, 0, st));
        }
        DOF_TRY(tfm_ln_bwd(w.tmpA, Hin, q.muA, q.rsA, state + Q.n1w, w.dH, 1, nullptr, drop_none(), 1, 0, grad + Q.n1w, grad + Q.n1b, R, dm, sm,
                           st));
    }
    const int ein[3] = {D, D, 2 * D}, eout[3] = {D, 2 * D, 4 * D};
    { ProfScope ps("sum_over_t", st);
    sum_over_t_kernel<<<cdiv((long long)B * dm, 256), 256, 0, st>>>(w.dH, w.dG[2], B, T, dm, 0); }
    DOF_LAUNCH_CHECK();
    for (int i = 2; i >= 0; i--) {
        { ProfScope ps("tfm_gelu_bwd", st);
        tfm_gelu_bwd_kernel<<<cdiv((long long)B * eout[i], 256), 256, 0, st>>>(w.dG[i], w.P[i], w.dP[i], (long long)B * eout[i]); }
        DOF_LAUNCH_CHECK();
        const float* in = i == 0 ? zin : w.G[i - 1];
        DOF_TRY(tfm_wgrad(w.dP[i], eout[i], in, ein[i], grad + L.dex_w[i], ein[i], 0, grad + L.dex_b[i], B, eout[i], ein[i], sm, st));
        DOF_TRY(tfm_gemm(w.dP[i], eout[i], state + L.dex_w[i], ein[i], 1, nullptr, i == 0 ? h->dz_dec : w.dG[i - 1], ein[i], B, ein[i], eout[i], 0, 0,
                         nullptr, 0, st));
    }
    return DOF_OK;
}

// running statistics of the two BatchNorm layers after a training step (called from dof_clip_adam)
static int tfm_bn_apply(dof_handle* h, float* state, cudaStream_t st) {
    const Layout& L = h->L;
    const int D = h->cfg.D;
    ProfScope ps("tfm_bn_update", st);
    tfm_bn_update_kernel<<<cdiv(2 * D, 128), 128, 0, st>>>(state + L.bn2.mean, state + L.bn2.var, state + L.bn2.tracked, h->bnstat[0], 2 * D,
                                                         h->enc_groups, 0.01f);
    tfm_bn_update_kernel<<<cdiv(D, 128), 128, 0, st>>>(state + L.bn5.mean, state + L.bn5.var, state + L.bn5.tracked, h->bnstat[1], D,
                                                     h->enc_groups, 0.01f);
    DOF_LAUNCH_CHECK();
    h->bn_pending = 0;
    return DOF_OK;
}
