// Orchestration of the TCN-family forward / backward on a dof_handle whose config says
// encoder == DOF_ENCODER_TCN.  Included by api.cu after tfm_step.cuh (shares the encoder tail: CensNet, RMS normalisation, head).
// Reference: deepof/clustering/models_new.py  TemporalBlockPT :376-445, TCN1DPT :447-503, TCNEncoderPT.forward :609-657,
// TCNDecoderPT.forward :792-819.
//
// Rows are (sequence, step) pairs, row-major [rows, channels].  A dilated causal Conv1d (kernel 4) is ONE GEMM whose A operand is
// the A_TAPS view of the input (column block j = the row (3 - j) * dilation steps earlier, zero before the window starts) against
// the Conv1d weight addressed in place (GemmArgs.wconv): tcgen05 3xTF32 when the channel count is a multiple of 4 and the
// operands fit the weight-resident kernel, the SIMT kernel otherwise (first encoder block: 3 / 1 input channels).  Its input
// gradient is the same GEMM with the taps pointing forward in time, its weight gradient one time-shifted weight-gradient GEMM
// per tap (batched in one launch).  Train-mode BatchNorm needs the statistics of ALL rows of a layer before the next layer can
// start, so every convolution is followed by a statistics pass and a normalise + ReLU pass (tcn.cuh).
#pragma once

static int tcn_ew_grid(long long R, int C, int sm, int per_sm) {
    const int rpb = 256 / (C >> 2);
    long long blocks = (R + rpb - 1) / rpb;
    const long long cap = (long long)sm * per_sm;
    if (blocks > cap) blocks = cap;
    return (int)(blocks < 1 ? 1 : blocks);
}

static int tcn_check_c(int C) {
    if ((C & 3) || C > TCN_MAXC || C < 4) DOF_FAIL(DOF_ERR_UNSUPPORTED, "TCN channel count %d (multiple of 4, <= %d)", C, TCN_MAXC);
    return DOF_OK;
}

static TcnBn tcn_bn_ref(const dof_handle* h, const float* state, const TfmBnP& bn, bool train, int pass, int st_off, long long rows) {
    TcnBn r;
    r.stat = train ? h->tstat + (long long)pass * h->tstat_stride + st_off : nullptr;
    r.inv_n = 1.0 / (double)rows;
    r.rmean = state + bn.mean; r.rvar = state + bn.var; r.w = state + bn.w; r.b = state + bn.b;
    r.eps = 1e-3f;
    return r;
}

// K = 4 * channels > 128 (the decoder's 64-channel convolutions, K = 256) only fits the weight-resident tensor-core kernel when it is
// staged in 32-column blocks (GemmArgs.ksplit).  For K = 128 the single 128-column stage is faster than four 32-column stages
// (measured at 1.4 M rows: 0.82 vs 0.95 ms per launch — the per-stage barrier round trips outweigh two CTAs per SM), so the
// encoder convolutions keep one stage.  DOF_TCN_KSPLIT=<columns> forces a block width for A/B runs (0: never split).
static void tcn_pick_ksplit(GemmArgs& g) {
    static const int forced = getenv("DOF_TCN_KSPLIT") ? atoi(getenv("DOF_TCN_KSPLIT")) : -1;
    const int width = forced >= 0 ? forced : (g.K > 128 ? 32 : 0);
    if (width <= 0 || g.K <= width || (g.K % width) || (g.A.cc & 3)) return;
    g.ksplit = g.K / width;
    if (!tc_rows_eligible(g)) g.ksplit = 0;
}

// A [R, C] = conv(X [R, cin]) + bias, dilated causal
// (ldx = row pitch of X; when it exceeds cin the pad columns are zero and every tap spans ldx columns of the operand)
static int tcn_conv_fwd(const float* X, int ldx, int cin, int T, int dil, const float* W, const float* bias, float* A, int C, long long R,
                        cudaStream_t st) {
    GemmArgs g = gemm_args(mv_taps(X, ldx, T, TCN_TAPS, ldx, -dil), W, cin * TCN_TAPS, 0, bias, A, C, (int)R, C, TCN_TAPS * ldx);
    g.wconv = 1; g.wcin = cin; g.wtaps = TCN_TAPS; g.wcpad = ldx;
    tcn_pick_ksplit(g);
    return launch_gemm_rows(&g, 1, st);
}

// dX [R, cin] (+)= conv^T(dA [R, C]); mask: dX = mask > 0 ? dX : 0
static int tcn_conv_dgrad(const float* dA, int C, int T, int dil, const float* W, int cin, float* dX, long long R, int accum,
                          const float* mask, cudaStream_t st) {
    GemmArgs g = gemm_args(mv_taps(dA, C, T, TCN_TAPS, C, +dil), W, cin * TCN_TAPS, 0, nullptr, dX, cin, (int)R, cin, TCN_TAPS * C);
    g.wconv = 2; g.wcin = cin; g.wtaps = TCN_TAPS; g.accum = accum;
    if (mask) { g.mask = mask; g.ldmask = cin; }
    tcn_pick_ksplit(g);
    return launch_gemm_rows(&g, 1, st);
}

// dW [C, cin, 4] += sum_rows dA[(s, t), :]^T X[(s, t - (3 - j) dil), :];  db += sum_rows dA
static int tcn_conv_wgrad(const float* dA, int C, const float* X, int ldx, int cin, int T, int dil, float* dW, float* db, long long R, int sm,
                          cudaStream_t st) {
    // all four taps in ONE weight-gradient GEMM when K = 4 * cin fits the tensor-core kernel: Q = the A_TAPS view of X (dA is read
    // once instead of once per tap), the epilogue scatters column (tap, channel) to the Conv1d layout [C, cin, 4]
    static const bool fused = !(getenv("DOF_TCN_WGRAD_FUSED") && getenv("DOF_TCN_WGRAD_FUSED")[0] == '0');
    if (fused && (ldx & 3) == 0) {
        WGradArgs f = wgrad_args(mv_plain(dA, C), mv_taps(X, ldx, T, TCN_TAPS, ldx, -dil), dW, cin * TCN_TAPS, 0, db, (int)R, C, TCN_TAPS * ldx);
        f.otaps = TCN_TAPS; f.ocv = cin;
        if (tc_enabled() && tc_wgrad_eligible(f)) return launch_gemm_wgrad_tc(&f, 1, st, sm);
    }
    WGradArgs w[TCN_TAPS];
    for (int j = 0; j < TCN_TAPS; j++) {
        w[j] = wgrad_args(mv_plain(dA, C), mv_tshift(X, ldx, T, -(TCN_TAPS - 1 - j) * dil), dW + j, cin * TCN_TAPS, 0,
                          j == TCN_TAPS - 1 ? db : nullptr, (int)R, C, cin);
        w[j].ks = TCN_TAPS;
    }
    return launch_gemm_wgrad(w, TCN_TAPS, st, sm);
}

static int tcn_stats(const float* A, long long R, int C, double* stat, int sm, cudaStream_t st) {
    ProfScope ps("tcn_stats", st, 0.0, 4.0 * R * C);
    tcn_stats_kernel<<<tcn_ew_grid(R, C, sm, 8), 256, 0, st>>>(A, R, C, stat);
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

// one TCN1DPT over the rows [r0, r0 + R) of the stack's buffers (r0 a multiple of T): X0 -> SKIP / FIN
static int tcn_stack_forward(dof_handle* h, int si, const float* state, long long r0, long long R, bool train, int pass, bool last_only,
                             cudaStream_t st) {
    const Layout& L = h->L;
    const TcnStackP& P = L.tstack[si];
    const TcnStackWS& w = h->ts[si];
    const int C = P.C, T = h->cfg.T, sm = h->sm_count;
    DOF_TRY(tcn_check_c(C));
    if (R <= 0) return DOF_OK;
    if (train) {
        double* s0 = h->tstat + (long long)pass * h->tstat_stride + w.b[0].st1;
        DOF_CUDA(cudaMemsetAsync(s0, 0, (size_t)P.nb * 6 * C * sizeof(double), st));
    }
    const long long sr0 = last_only ? r0 / T : r0;
    for (int i = 0; i < P.nb; i++) {
        const TcnBlockP& B = P.blk[i];
        const TcnBlockWS& q = w.b[i];
        const int ldx = i == 0 ? w.x0_ld : C;
        const float* Xin = (i == 0 ? w.X0 : w.b[i - 1].OUT) + r0 * ldx;
        float *A1 = q.A1 + r0 * C, *Y1 = q.Y1 + r0 * C, *A2 = q.A2 + r0 * C;
        DOF_TRY(tcn_conv_fwd(Xin, ldx, B.cin, T, B.dil, state + B.c1w, state + B.c1b, A1, C, R, st));
        const TcnBn bn1 = tcn_bn_ref(h, state, B.bn1, train, pass, q.st1, R);
        if (train) DOF_TRY(tcn_stats(A1, R, C, const_cast<double*>(bn1.stat), sm, st));
        { ProfScope ps("tcn_bnrelu", st, 0.0, 8.0 * R * C);
        tcn_bnrelu_kernel<<<tcn_ew_grid(R, C, sm, 8), 256, 0, st>>>(A1, Y1, R, C, bn1); }
        DOF_LAUNCH_CHECK();
        DOF_TRY(tcn_conv_fwd(Y1, C, C, T, B.dil, state + B.c2w, state + B.c2b, A2, C, R, st));
        const TcnBn bn2 = tcn_bn_ref(h, state, B.bn2, train, pass, q.st2, R);
        if (train) DOF_TRY(tcn_stats(A2, R, C, const_cast<double*>(bn2.stat), sm, st));
        const float* res = Xin;
        if (B.has_ds) {                                   // 1x1 residual projection (:418, :442)
            GemmArgs g = gemm_args(mv_plain(Xin, ldx), state + B.dsw, B.cin, 0, state + B.dsb, q.RES + r0 * C, C, (int)R, C, B.cin);
            DOF_TRY(launch_gemm_rows(&g, 1, st));
            res = q.RES + r0 * C;
        }
        TcnOutArgs o;
        memset(&o, 0, sizeof(o));
        const bool last = i == P.nb - 1;
        o.A2 = A2; o.RES = res; o.OUT = last ? nullptr : q.OUT + r0 * C;
        o.SKIP = w.SKIP + sr0 * C; o.FIN = last ? w.FIN + sr0 * C : nullptr;
        o.R = R; o.C = C; o.T = T; o.first = i == 0; o.last_only = last_only ? 1 : 0; o.bn = bn2;
        { ProfScope ps("tcn_block_out", st, 0.0, (last ? 4.0 : 12.0) * R * C);
        tcn_block_out_kernel<<<tcn_ew_grid(R, C, sm, 8), 256, 0, st>>>(o); }
        DOF_LAUNCH_CHECK();
    }
    return DOF_OK;
}

// backward of tcn_stack_forward.  In: GS (gradient of the skip sum, already masked by FIN > 0) at [sr0, ...).  Parameter gradients
// are accumulated into grad; *dx0 receives d(loss)/d(X0) [R, cin0] when need_dx0.
static int tcn_stack_backward(dof_handle* h, int si, const float* state, float* grad, long long r0, long long R, int pass, bool last_only,
                              bool need_dx0, const float** dx0, cudaStream_t st) {
    const Layout& L = h->L;
    const TcnStackP& P = L.tstack[si];
    const TcnStackWS& w = h->ts[si];
    const int C = P.C, T = h->cfg.T, sm = h->sm_count;
    if (R <= 0) return DOF_OK;
    DOF_CUDA(cudaMemsetAsync(h->tbstat + w.b[0].st1, 0, (size_t)P.nb * 6 * C * sizeof(double), st));
    const long long sr0 = last_only ? r0 / T : r0;
    float *DA = w.DA + r0 * C, *DB = w.DB + r0 * C, *DX = w.DX + r0 * C;
    for (int i = P.nb - 1; i >= 0; i--) {
        const TcnBlockP& B = P.blk[i];
        const TcnBlockWS& q = w.b[i];
        const int ldx = i == 0 ? w.x0_ld : C;
        const float* Xin = (i == 0 ? w.X0 : w.b[i - 1].OUT) + r0 * ldx;
        const float *A1 = q.A1 + r0 * C, *Y1 = q.Y1 + r0 * C, *A2 = q.A2 + r0 * C;
        const bool last = i == P.nb - 1;
        const TcnBn bn1 = tcn_bn_ref(h, state, B.bn1, true, pass, q.st1, R), bn2 = tcn_bn_ref(h, state, B.bn2, true, pass, q.st2, R);
        double *bs1 = h->tbstat + q.st1, *bs2 = h->tbstat + q.st2;
        TcnBwdArgs a;
        memset(&a, 0, sizeof(a));
        a.dOUT = last ? nullptr : DX; a.OUT = last ? nullptr : q.OUT + r0 * C;
        a.GS = w.GS + sr0 * C; a.DS = DX; a.D = DA; a.A = A2; a.bstat = bs2;
        a.R = R; a.C = C; a.T = T; a.mode = 2; a.gs_last_only = last_only ? 1 : 0; a.bn = bn2;
        { ProfScope ps("tcn_bwd_reduce", st, 0.0, (last ? 12.0 : 20.0) * R * C);
        tcn_bwd_reduce_kernel<<<tcn_ew_grid(R, C, sm, 8), 256, 0, st>>>(a); }
        DOF_LAUNCH_CHECK();
        { ProfScope ps("tcn_bn_bwd_apply", st, 0.0, 12.0 * R * C);
        tcn_bn_bwd_apply_kernel<<<tcn_ew_grid(R, C, sm, 8), 256, 0, st>>>(DA, A2, R, C, bn2, bs2, grad + B.bn2.w, grad + B.bn2.b); }
        DOF_LAUNCH_CHECK();
        DOF_TRY(tcn_conv_wgrad(DA, C, Y1, C, C, T, B.dil, grad + B.c2w, grad + B.c2b, R, sm, st));
        DOF_TRY(tcn_conv_dgrad(DA, C, T, B.dil, state + B.c2w, C, DB, R, 0, Y1, st));     // mask: relu(bn1(A1)) > 0
        a.mode = 1; a.D = DB; a.A = A1; a.bstat = bs1; a.bn = bn1; a.dOUT = a.OUT = a.GS = nullptr; a.DS = nullptr;
        { ProfScope ps("tcn_bwd_reduce", st, 0.0, 8.0 * R * C);
        tcn_bwd_reduce_kernel<<<tcn_ew_grid(R, C, sm, 8), 256, 0, st>>>(a); }
        DOF_LAUNCH_CHECK();
        { ProfScope ps("tcn_bn_bwd_apply", st, 0.0, 12.0 * R * C);
        tcn_bn_bwd_apply_kernel<<<tcn_ew_grid(R, C, sm, 8), 256, 0, st>>>(DB, A1, R, C, bn1, bs1, grad + B.bn1.w, grad + B.bn1.b); }
        DOF_LAUNCH_CHECK();
        DOF_TRY(tcn_conv_wgrad(DB, C, Xin, ldx, B.cin, T, B.dil, grad + B.c1w, grad + B.c1b, R, sm, st));
        if (B.has_ds) {
            WGradArgs wd = wgrad_args(mv_plain(DX, C), mv_plain(Xin, ldx), grad + B.dsw, B.cin, 0, grad + B.dsb, (int)R, C, ldx);
            wd.kv = B.cin;
            DOF_TRY(launch_gemm_wgrad(&wd, 1, st, sm));
        }
        if (i > 0 || need_dx0) {
            if (B.has_ds) {
                if (i > 0) DOF_FAIL(DOF_ERR_UNSUPPORTED, "residual projection inside a TCN stack");
                float* D0 = w.DX0 + r0 * B.cin;
                DOF_TRY(tcn_conv_dgrad(DB, C, T, B.dil, state + B.c1w, B.cin, D0, R, 0, nullptr, st));
                GemmArgs g = gemm_args(mv_plain(DX, C), state + B.dsw, B.cin, 1, nullptr, D0, B.cin, (int)R, B.cin, C);
                g.accum = 1;
                DOF_TRY(launch_gemm_rows(&g, 1, st));
                if (dx0) *dx0 = D0;
            } else {
                DOF_TRY(tcn_conv_dgrad(DB, C, T, B.dil, state + B.c1w, B.cin, DX, R, 1, nullptr, st));
                if (i == 0 && dx0) *dx0 = DX;
            }
        }
    }
    return DOF_OK;
}

// ---- encoder ------------------------------------------------------------------------------------------------------------
static int tcn_core_forward(dof_handle* h, int b, const float* state, const float* xin, int Bw, bool train, int groups, cudaStream_t st) {
    const dof_config& c = h->cfg;
    const TcnStackWS& w = h->ts[b];
    const int T = c.T, TF = T * w.Fin;
    const long long S = (long long)Bw * w.G, n = S * TF;
    if (w.x0_ld != w.Fin) DOF_CUDA(cudaMemsetAsync(w.X0, 0, (size_t)S * T * w.x0_ld * 4, st));
    { ProfScope ps("tcn_gather", st, 0.0, 8.0 * n);
    tcn_gather_kernel<<<cdiv(n, 256), 256, 0, st>>>(xin, w.gidx, w.X0, n, w.G, TF, w.Fin, w.x0_ld); }
    DOF_LAUNCH_CHECK();
    const long long Rg = S / groups * T;
    for (int g = 0; g < groups; g++) DOF_TRY(tcn_stack_forward(h, b, state, g * Rg, Rg, train, g, true, st));
    return DOF_OK;
}

static int tcn_core_backward(dof_handle* h, int b, const float* state, float* grad, const float* dfin, int Bw, int groups, cudaStream_t st) {
    const dof_config& c = h->cfg;
    const TcnStackWS& w = h->ts[b];
    const int C = h->L.tstack[b].C;
    const long long S = (long long)Bw * w.G, n = S * C;
    { ProfScope ps("tcn_relu_mask", st, 0.0, 12.0 * n);
    tcn_relu_mask_kernel<<<cdiv(n, 256), 256, 0, st>>>(dfin, w.FIN, w.GS, n); }
    DOF_LAUNCH_CHECK();
    const long long Rg = S / groups * c.T;
    for (int g = 0; g < groups; g++) DOF_TRY(tcn_stack_backward(h, b, state, grad, g * Rg, Rg, g, true, false, nullptr, st));
    return DOF_OK;
}

// TCNEncoderPT.forward: Bw windows -> h->enc [Bw, D]; `groups` row ranges with separate batch statistics (train only)
static int tcn_encoder_forward(dof_handle* h, const float* state, const float* x, const float* a, int Bw, bool train, int groups,
                               cudaStream_t st) {
    if (groups < 1 || groups > 2 || Bw % groups) DOF_FAIL(DOF_ERR_ARG, "batch %d is not a multiple of %d statistics groups (<= 2)", Bw, groups);
    if (train && Bw / groups < 2) DOF_FAIL(DOF_ERR_UNSUPPORTED, "train-mode BatchNorm needs at least 2 windows per pass");
    DOF_TRY(fork_join_blocks(h, st, [&](int b, cudaStream_t s) { return tcn_core_forward(h, b, state, b == 0 ? x : a, Bw, train, groups, s); }));
    if (train) h->tcn_enc_windows = Bw / groups;
    return enc_tail_forward(h, state, h->ts[0].FIN, h->ts[1].FIN, Bw, train, groups, false, st);
}

// the CensNet backward writes d(loss)/d(node), d(loss)/d(edge) into the DA scratch of the two stacks ([S, C] fits in [R, C])
static int tcn_encoder_backward(dof_handle* h, const float* state, float* grad, int Bw, cudaStream_t st) {
    const int groups = h->enc_groups;
    float *dn = h->ts[0].DB, *de = h->ts[1].DB;
    DOF_TRY(enc_tail_backward(h, state, grad, h->ts[0].FIN, h->ts[1].FIN, dn, de, Bw, false, st));
    DOF_TRY(fork_join_blocks(h, st, [&](int b, cudaStream_t s) { return tcn_core_backward(h, b, state, grad, b == 0 ? dn : de, Bw, groups, s); }));
    return DOF_OK;
}

// ---- decoder ------------------------------------------------------------------------------------------------------------
static int tcn_col(bool bwd, int kind, const float* x, float* y, const float* w, const float* b, const float* rmean, const float* rvar,
                   float* mean, float* scale, float* stat_out, const float* dy, float* dx, float* dw, float* db, const float* relu_ref,
                   int B, int C, cudaStream_t st) {
    TfmColArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x; a.y = y; a.w = w; a.b = b; a.run_mean = rmean; a.run_var = rvar; a.mean = mean; a.scale = scale;
    a.stat_out = stat_out; a.dy = dy; a.dx = dx; a.dw = dw; a.db = db; a.relu_ref = relu_ref;
    a.B = B; a.C = C; a.groups = 1; a.kind = kind; a.eps = 1e-3f;
    dim3 grid(cdiv(C, 32), 1), block(32, 32);
    ProfScope ps(bwd ? "tfm_col_bwd" : "tfm_col_fwd", st);
    if (bwd) tfm_col_bwd_kernel<<<grid, block, 0, st>>>(a);
    else tfm_col_fwd_kernel<<<grid, block, 0, st>>>(a);
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

// TCNDecoderPT.forward: zin [B, D] -> h->loc [B, T, N * F].  `pass`: which decoder pass of the step (VQ-VAE decodes twice; each pass
// has its own batch statistics and moves the running buffers once).
static int tcn_decoder_forward(dof_handle* h, const float* state, const float* zin, int B, bool train, int pass, cudaStream_t st) {
    const dof_config& c = h->cfg;
    const Layout& L = h->L;
    const int T = c.T, D = c.D, Dx = c.N * c.F, C = L.tstack[2].C;
    const long long R = (long long)B * T;
    if (pass < 0 || pass > 1) DOF_FAIL(DOF_ERR_ARG, "decoder pass %d", pass);
    if (train && B < 2) DOF_FAIL(DOF_ERR_UNSUPPORTED, "train-mode BatchNorm needs at least 2 windows per pass");
    TfmRmsArgs ra;
    memset(&ra, 0, sizeof(ra));
    ra.on = zin; ra.h = h->dgz; ra.rms = h->drms; ra.B = B; ra.ND = D; ra.ED = 0; ra.norelu = 1;
    { ProfScope ps("tfm_rms_fwd", st);
    tfm_rms_fwd_kernel<<<cdiv((long long)B * 32, 256), 256, 0, st>>>(ra); }
    DOF_LAUNCH_CHECK();
    const int fin[3] = {D, D, 2 * D}, fout[3] = {D, 2 * D, 4 * D};
    const float* cur = h->dgz;
    for (int i = 0; i < 3; i++) {                          // z = bn0(fc0(g)); z = bn1(relu(fc1(z))); z = bn2(relu(fc2(z)))  (:811-813)
        DOF_TRY(tfm_gemm(cur, fin[i], state + L.dfc_w[i], fin[i], 0, state + L.dfc_b[i], h->dfo[i], fout[i], B, fout[i], fin[i], i > 0 ? 1 : 0, 0,
                         nullptr, 0, st));
        const TfmBnP& bn = L.dbn[i];
        DOF_TRY(tcn_col(false, train ? 0 : 1, h->dfo[i], h->dzo[i], state + bn.w, state + bn.b, state + bn.mean, state + bn.var, h->dbnm[i],
                        h->dbns[i], h->dbnstat[i] + (size_t)pass * 2 * fout[i], nullptr, nullptr, nullptr, nullptr, nullptr, B, fout[i], st));
        cur = h->dzo[i];
    }
    { ProfScope ps("tcn_repeat", st, 0.0, 4.0 * R * 4 * D);
    tcn_repeat_kernel<<<cdiv(R * D, 256), 256, 0, st>>>(cur, h->ts[2].X0, R * D, T, D); }
    DOF_LAUNCH_CHECK();
    DOF_TRY(tcn_stack_forward(h, 2, state, 0, R, train, pass, false, st));
    DOF_TRY(tfm_gemm(h->ts[2].FIN, C, state + L.loc_w, C, 0, state + L.loc_b, h->loc, Dx, (int)R, Dx, C, 0, 0, nullptr, 0, st));
    if (train) {
        h->tcn_dec_windows = B;
        if (pass + 1 > h->tcn_dec_passes) h->tcn_dec_passes = pass + 1;
        h->bn_pending = 1;
    }
    return DOF_OK;
}

// backward of the TCN decoder from h->dloc; writes d(loss)/d(zin) into h->dz_dec
static int tcn_decoder_backward(dof_handle* h, const float* state, float* grad, const float* zin, int B, int pass, cudaStream_t st) {
    const dof_config& c = h->cfg;
    const Layout& L = h->L;
    const TcnStackWS& w = h->ts[2];
    const int T = c.T, D = c.D, Dx = c.N * c.F, C = L.tstack[2].C, sm = h->sm_count;
    const long long R = (long long)B * T;
    const int DxP = round_up(Dx, 4);
    DOF_TRY(tfm_wgrad(padded_dloc(h, R, st), DxP, w.FIN, C, grad + L.loc_w, C, 0, grad + L.loc_b, (int)R, DxP, C, sm, st, Dx, 0));
    DOF_TRY(tfm_gemm(h->dloc, Dx, state + L.loc_w, C, 1, nullptr, w.GS, C, (int)R, C, Dx, 0, 0, w.FIN, C, st));      // mask: FIN > 0
    const float* dx0 = nullptr;
    DOF_TRY(tcn_stack_backward(h, 2, state, grad, 0, R, pass, false, true, &dx0, st));
    { ProfScope ps("sum_over_t", st);
    sum_over_t_kernel<<<cdiv((long long)B * 4 * D, 256), 256, 0, st>>>(dx0, h->ddz[2], B, T, 4 * D, 0); }
    DOF_LAUNCH_CHECK();
    const int fin[3] = {D, D, 2 * D}, fout[3] = {D, 2 * D, 4 * D};
    for (int i = 2; i >= 0; i--) {
        const TfmBnP& bn = L.dbn[i];
        DOF_TRY(tcn_col(true, 0, h->dfo[i], nullptr, state + bn.w, nullptr, nullptr, nullptr, h->dbnm[i], h->dbns[i], nullptr, h->ddz[i], h->ddf[i],
                        grad + bn.w, grad + bn.b, i > 0 ? h->dfo[i] : nullptr, B, fout[i], st));
        const float* in = i == 0 ? h->dgz : h->dzo[i - 1];
        DOF_TRY(tfm_wgrad(h->ddf[i], fout[i], in, fin[i], grad + L.dfc_w[i], fin[i], 0, grad + L.dfc_b[i], B, fout[i], fin[i], sm, st));
        DOF_TRY(tfm_gemm(h->ddf[i], fout[i], state + L.dfc_w[i], fin[i], 1, nullptr, i == 0 ? h->ddg : h->ddz[i - 1], fin[i], B, fin[i], fout[i], 0, 0,
                         nullptr, 0, st));
    }
    TfmRmsArgs ra;
    memset(&ra, 0, sizeof(ra));
    ra.on = zin; ra.h = h->dgz; ra.rms = h->drms; ra.dh = h->ddg; ra.don = h->dz_dec; ra.B = B; ra.ND = D; ra.ED = 0; ra.norelu = 1;
    { ProfScope ps("tfm_rms_bwd", st);
    tfm_rms_bwd_kernel<<<cdiv((long long)B * 32, 256), 256, 0, st>>>(ra); }
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

// running statistics of every BatchNorm of the TCN family after a training step (called from dof_clip_adam): the two head layers
// through tfm_bn_apply, the three decoder front layers, and all TemporalBlockPT layers in one table-driven launch
static int tcn_bn_apply(dof_handle* h, float* state, cudaStream_t st) {
    const Layout& L = h->L;
    const int D = h->cfg.D;
    const int fout[3] = {D, 2 * D, 4 * D};
    ProfScope ps("tcn_bn_update", st);
    if (h->tcn_dec_passes > 0 && h->cfg.model != DOF_MODEL_CONTRASTIVE)
        for (int i = 0; i < 3; i++)
            tfm_bn_update_kernel<<<cdiv(fout[i], 128), 128, 0, st>>>(state + L.dbn[i].mean, state + L.dbn[i].var, state + L.dbn[i].tracked,
                                                                   h->dbnstat[i], fout[i], h->tcn_dec_passes, 0.01f);
    const int enc_passes = h->tcn_enc_windows > 0 ? h->enc_groups : 0;
    tcn_bn_update_kernel<<<h->tn_desc, 64, 0, st>>>(state, h->tdesc, h->tstat, h->tstat_stride, enc_passes, h->tcn_enc_windows,
                                                  h->tcn_dec_passes, h->tcn_dec_windows, 0.1f);
    DOF_LAUNCH_CHECK();
    h->tcn_enc_windows = 0; h->tcn_dec_passes = 0; h->tcn_dec_windows = 0;
    return DOF_OK;
}
