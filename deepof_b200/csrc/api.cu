// deepof_b200 C-ABI: state layout, workspace, and the orchestration of the VaDE / recurrent
// forward, loss, backward and optimizer kernels.  See include/deepof_b200.h.
#include <string>
#include <vector>
#include <map>

#include "../../include/deepof_b200.h"
#include "common.cuh"
#include "gemm.cuh"
#include "tc_gemm.cuh"
#include "gru.cuh"
#include "gru_tc.cuh"
#include "gru_bwd_tc.cuh"
#include "gru_wgrad_tc.cuh"
#include "gru_bwdw_tc.cuh"
#include "peer.cuh"
#include "layers.cuh"
#include "loss.cuh"
#include "loader.cuh"
#include "models.cuh"
#include "tfm.cuh"
#include "tcn.cuh"

thread_local char g_dof_err[512] = {0};
DofProf g_prof;
extern "C" { static int loader_device_check(); }

// ---------------------------------------------------------------------------
// state layout  (reference VaDEPT.state_dict() order)
// ---------------------------------------------------------------------------
struct Entry {
    std::string name;
    int64_t off, numel;
    int ndim;
    int shape[4];
    int group;
};

struct GruP { int64_t w_ih[2], w_hh[2], b_ih[2], b_hh[2]; };
struct BlockP { int64_t conv; GruP g1, g2; int64_t n1w, n1b, n2w, n2b, pw, pb; };
#define TFM_MAXL 4
struct TfmLayerP { int64_t Wqkv, Wo, n1w, n1b, W1, b1, W2, b2, n2w, n2b; };
struct TfmCoreP { int64_t We, be; TfmLayerP l[TFM_MAXL]; };
struct TfmBnP { int64_t w, b, mean, var, tracked; };
// TCN family: one TemporalBlockPT (two dilated causal Conv1d + BatchNorm, optional 1x1 residual projection) and one TCN1DPT
#define TCN_MAXB 8
#define TCN_TAPS 4
struct TcnBlockP { int64_t c1w, c1b, c2w, c2b, dsw, dsb; TfmBnP bn1, bn2; int cin, dil, has_ds; };
struct TcnStackP { int nb, C, cin0; TcnBlockP blk[TCN_MAXB]; };
struct Layout {
    std::vector<Entry> e;
    int64_t total = 0;
    int64_t lap, elap, inc;
    BlockP blk[2];
    // transformer family (encoder == DOF_ENCODER_TRANSFORMER)
    int dk = 0, heads = 4, dff = 128, layers = 2, dec_heads = 8, dec_dff = 128, dec_layers = 2;
    TfmCoreP tcore[2];
    int64_t h_w0, h_b0, h_w3, h_b3, h_w6, h_b6;
    TfmBnP bn2, bn5;
    int64_t dex_w[3], dex_b[3], dout_w, dout_b;
    TfmLayerP tdl[TFM_MAXL];
    // TCN family (encoder == DOF_ENCODER_TCN): node / edge / decoder stacks, decoder front MLP (fc0..2 + bn0..2)
    TcnStackP tstack[3];
    int64_t dfc_w[3], dfc_b[3];
    TfmBnP dbn[3];
    int64_t node_kernel, edge_kernel, node_weights, edge_weights, node_bias, edge_bias;
    int64_t final_w, final_b;
    GruP dg1, dg2;
    int64_t dn1w, dn1b, dn2w, dn2b, dconv, dn3w, dn3b, loc_w, loc_b;
    int64_t gmm_mu, gmm_lv, prior, pretrain, Wm, bm, Wv, bv, lens_w, lens_b;
    int64_t codebook;
};

static int64_t add_entry(Layout& L, const std::string& name, int group, int d0, int d1 = -1, int d2 = -1) {
    Entry en;
    en.name = name;
    en.off = L.total;
    en.group = group;
    en.shape[0] = en.shape[1] = en.shape[2] = en.shape[3] = 0;
    if (d0 < 0) { en.ndim = 0; en.numel = 1; }
    else if (d1 < 0) { en.ndim = 1; en.shape[0] = d0; en.numel = d0; }
    else if (d2 < 0) { en.ndim = 2; en.shape[0] = d0; en.shape[1] = d1; en.numel = (int64_t)d0 * d1; }
    else { en.ndim = 3; en.shape[0] = d0; en.shape[1] = d1; en.shape[2] = d2; en.numel = (int64_t)d0 * d1 * d2; }
    L.total += en.numel;
    L.e.push_back(en);
    return en.off;
}

static void add_gru(Layout& L, const std::string& p, int group, int I, int H, GruP& g) {
    const char* suf[2] = {"", "_reverse"};
    for (int d = 0; d < 2; d++) {
        g.w_ih[d] = add_entry(L, p + "weight_ih_l0" + suf[d], group, 3 * H, I);
        g.w_hh[d] = add_entry(L, p + "weight_hh_l0" + suf[d], group, 3 * H, H);
        g.b_ih[d] = add_entry(L, p + "bias_ih_l0" + suf[d], group, 3 * H);
        g.b_hh[d] = add_entry(L, p + "bias_hh_l0" + suf[d], group, 3 * H);
    }
}

static int dint(const dof_config& c) { return c.D < 64 ? c.D : 64; }

static int tfm_key_dim(const dof_config& c) {                 // models_new.py:1014-1019
    int kd = c.N * c.F < 64 ? c.N * c.F : 64;
    kd = (kd / 4) * 4;
    return kd < 4 ? 4 : kd;
}

static void add_latent_tail(Layout& L, const dof_config& c) {
    const int D = c.D, K = c.K;
    if (c.model == DOF_MODEL_VQVAE) {                        // VQVAEPT: vq_layer.codebook [D,K] (models_new.py:1349-1351)
        L.codebook = add_entry(L, "vq_layer.codebook", 3, D, K);
        return;
    }
    L.gmm_mu = add_entry(L, "latent_space.gmm_means", 3, K, D);
    L.gmm_lv = add_entry(L, "latent_space.gmm_log_vars", 3, K, D);
    L.prior = add_entry(L, "latent_space.prior", 0, K);
    L.pretrain = add_entry(L, "latent_space.pretrain", 0, -1);
    L.Wm = add_entry(L, "latent_space.encoder_mean.weight", 1, D, D);
    L.bm = add_entry(L, "latent_space.encoder_mean.bias", 1, D);
    L.Wv = add_entry(L, "latent_space.encoder_log_var.weight", 1, D, D);
    L.bv = add_entry(L, "latent_space.encoder_log_var.bias", 1, D);
    L.lens_w = add_entry(L, "latent_space.lens.weight", 0, D, D);   // never receives a gradient
    L.lens_b = add_entry(L, "latent_space.lens.bias", 0, D);
}

// state_dict order of the reference models built with encoder_type="transformer" (TFMEncoderPT after its first
// forward — the CensNet parameters are created lazily —, TFMDecoderPT, then the latent space / codebook)
static Layout build_layout_tfm(const dof_config& c) {
    Layout L;
    const int N = c.N, E = c.E, D = c.D, dk = tfm_key_dim(c);
    L.dk = dk;
    L.lap = add_entry(L, "encoder.laplacian", 0, N, N);
    L.elap = add_entry(L, "encoder.edge_laplacian", 0, E, E);
    L.inc = add_entry(L, "encoder.incidence", 0, N, E);
    const char* cn[2] = {"encoder.node_tf.", "encoder.edge_tf."};
    for (int b = 0; b < 2; b++) {
        std::string p = cn[b];
        TfmCoreP& P = L.tcore[b];
        P.We = add_entry(L, p + "embed.weight", 1, dk, b == 0 ? c.F : c.Fe);
        P.be = add_entry(L, p + "embed.bias", 1, dk);
        for (int l = 0; l < L.layers; l++) {
            std::string q = p + "layers." + std::to_string(l) + ".";
            TfmLayerP& Q = P.l[l];
            Q.Wqkv = add_entry(L, q + "mha.q_proj.weight", 1, dk, dk);
            add_entry(L, q + "mha.k_proj.weight", 1, dk, dk);
            add_entry(L, q + "mha.v_proj.weight", 1, dk, dk);
            Q.Wo = add_entry(L, q + "mha.out_proj.weight", 1, dk, dk);
            Q.n1w = add_entry(L, q + "norm1.weight", 1, dk); Q.n1b = add_entry(L, q + "norm1.bias", 1, dk);
            Q.W1 = add_entry(L, q + "ffn.0.weight", 1, L.dff, dk); Q.b1 = add_entry(L, q + "ffn.0.bias", 1, L.dff);
            Q.W2 = add_entry(L, q + "ffn.2.weight", 1, dk, L.dff); Q.b2 = add_entry(L, q + "ffn.2.bias", 1, dk);
            Q.n2w = add_entry(L, q + "norm2.weight", 1, dk); Q.n2b = add_entry(L, q + "norm2.bias", 1, dk);
        }
    }
    std::string g = "encoder.spatial_gnn_block.";
    L.node_kernel = add_entry(L, g + "node_kernel", 1, dk, D);
    L.edge_kernel = add_entry(L, g + "edge_kernel", 1, dk, D);
    L.node_weights = add_entry(L, g + "node_weights", 1, dk, 1);
    L.edge_weights = add_entry(L, g + "edge_weights", 1, dk, 1);
    L.node_bias = add_entry(L, g + "node_bias", 1, D);
    L.edge_bias = add_entry(L, g + "edge_bias", 1, D);
    auto add_bn = [&](const std::string& p, int C, TfmBnP& bn) {
        bn.w = add_entry(L, p + "weight", 1, C); bn.b = add_entry(L, p + "bias", 1, C);
        bn.mean = add_entry(L, p + "running_mean", 0, C); bn.var = add_entry(L, p + "running_var", 0, C);
        bn.tracked = add_entry(L, p + "num_batches_tracked", 0, -1);
    };
    L.h_w0 = add_entry(L, "encoder.head.0.weight", 1, 2 * D, (N + E) * D);
    L.h_b0 = add_entry(L, "encoder.head.0.bias", 1, 2 * D);
    add_bn("encoder.head.2.", 2 * D, L.bn2);
    L.h_w3 = add_entry(L, "encoder.head.3.weight", 1, D, 2 * D);
    L.h_b3 = add_entry(L, "encoder.head.3.bias", 1, D);
    add_bn("encoder.head.5.", D, L.bn5);
    L.h_w6 = add_entry(L, "encoder.head.6.weight", 1, D, D);
    L.h_b6 = add_entry(L, "encoder.head.6.bias", 1, D);
    if (c.model == DOF_MODEL_CONTRASTIVE) return L;
    const int dm = 4 * D, Dx = N * c.F;
    const int ein[3] = {D, D, 2 * D}, eout[3] = {D, 2 * D, 4 * D};
    for (int i = 0; i < 3; i++) {
        L.dex_w[i] = add_entry(L, "decoder.latent_expand." + std::to_string(2 * i) + ".weight", 2, eout[i], ein[i]);
        L.dex_b[i] = add_entry(L, "decoder.latent_expand." + std::to_string(2 * i) + ".bias", 2, eout[i]);
    }
    for (int l = 0; l < L.dec_layers; l++) {
        std::string q = "decoder.layers." + std::to_string(l) + ".";
        TfmLayerP& Q = L.tdl[l];
        Q.Wqkv = add_entry(L, q + "q_proj.weight", 2, dm, dm); add_entry(L, q + "k_proj.weight", 2, dm, dm);
        add_entry(L, q + "v_proj.weight", 2, dm, dm);
        Q.Wo = add_entry(L, q + "out_proj.weight", 2, dm, dm);
        Q.n1w = add_entry(L, q + "norm1.weight", 2, dm); Q.n1b = add_entry(L, q + "norm1.bias", 2, dm);
        Q.n2w = add_entry(L, q + "norm2.weight", 2, dm); Q.n2b = add_entry(L, q + "norm2.bias", 2, dm);
        Q.W1 = add_entry(L, q + "ffn.0.weight", 2, L.dec_dff, dm); Q.b1 = add_entry(L, q + "ffn.0.bias", 2, L.dec_dff);
        Q.W2 = add_entry(L, q + "ffn.3.weight", 2, dm, L.dec_dff); Q.b2 = add_entry(L, q + "ffn.3.bias", 2, dm);
    }
    L.dout_w = add_entry(L, "decoder.output_proj.weight", 2, Dx, dm);
    L.dout_b = add_entry(L, "decoder.output_proj.bias", 2, Dx);
    L.loc_w = add_entry(L, "decoder.prob_decoder.loc_projection.weight", 2, Dx, Dx);
    L.loc_b = add_entry(L, "decoder.prob_decoder.loc_projection.bias", 2, Dx);
    add_latent_tail(L, c);
    return L;
}

static void add_bn_entries(Layout& L, const std::string& p, int group, int C, TfmBnP& bn) {
    bn.w = add_entry(L, p + "weight", group, C); bn.b = add_entry(L, p + "bias", group, C);
    bn.mean = add_entry(L, p + "running_mean", 0, C); bn.var = add_entry(L, p + "running_var", 0, C);
    bn.tracked = add_entry(L, p + "num_batches_tracked", 0, -1);
}

static void add_tcn_stack(Layout& L, const std::string& p, int group, int cin0, int C, const int* dils, int nb, TcnStackP& S) {
    S.nb = nb; S.C = C; S.cin0 = cin0;
    int cin = cin0;
    for (int i = 0; i < nb; i++) {
        const std::string q = p + "blocks." + std::to_string(i) + ".";
        TcnBlockP& B = S.blk[i];
        B.cin = cin; B.dil = dils[i]; B.has_ds = cin != C ? 1 : 0;
        B.c1w = add_entry(L, q + "conv1.weight", group, C, cin, TCN_TAPS); B.c1b = add_entry(L, q + "conv1.bias", group, C);
        add_bn_entries(L, q + "bn1.", group, C, B.bn1);
        B.c2w = add_entry(L, q + "conv2.weight", group, C, C, TCN_TAPS); B.c2b = add_entry(L, q + "conv2.bias", group, C);
        add_bn_entries(L, q + "bn2.", group, C, B.bn2);
        B.dsw = B.dsb = -1;
        if (B.has_ds) { B.dsw = add_entry(L, q + "downsample.weight", group, C, cin, 1); B.dsb = add_entry(L, q + "downsample.bias", group, C); }
        cin = C;
    }
}

// state_dict order of the reference models built with encoder_type="TCN" (TCNEncoderPT models_new.py:574-607 after its first
// forward, TCNDecoderPT :745-775, then the latent space / codebook)
static Layout build_layout_tcn(const dof_config& c) {
    Layout L;
    const int N = c.N, E = c.E, D = c.D;
    const int enc_d[8] = {1, 2, 4, 8, 1, 2, 4, 8}, dec_d[4] = {8, 4, 2, 1};
    L.dk = 32;                                            // conv_filters of the encoder = CensNet input channels
    L.lap = add_entry(L, "encoder.laplacian", 0, N, N);
    L.elap = add_entry(L, "encoder.edge_laplacian", 0, E, E);
    L.inc = add_entry(L, "encoder.incidence", 0, N, E);
    add_tcn_stack(L, "encoder.node_tcn.", 1, c.F, 32, enc_d, 8, L.tstack[0]);
    add_tcn_stack(L, "encoder.edge_tcn.", 1, c.Fe, 32, enc_d, 8, L.tstack[1]);
    std::string g = "encoder.spatial_gnn_block.";
    L.node_kernel = add_entry(L, g + "node_kernel", 1, L.dk, D);
    L.edge_kernel = add_entry(L, g + "edge_kernel", 1, L.dk, D);
    L.node_weights = add_entry(L, g + "node_weights", 1, L.dk, 1);
    L.edge_weights = add_entry(L, g + "edge_weights", 1, L.dk, 1);
    L.node_bias = add_entry(L, g + "node_bias", 1, D);
    L.edge_bias = add_entry(L, g + "edge_bias", 1, D);
    L.h_w0 = add_entry(L, "encoder.head.0.weight", 1, 2 * D, (N + E) * D);
    L.h_b0 = add_entry(L, "encoder.head.0.bias", 1, 2 * D);
    add_bn_entries(L, "encoder.head.2.", 1, 2 * D, L.bn2);
    L.h_w3 = add_entry(L, "encoder.head.3.weight", 1, D, 2 * D);
    L.h_b3 = add_entry(L, "encoder.head.3.bias", 1, D);
    add_bn_entries(L, "encoder.head.5.", 1, D, L.bn5);
    L.h_w6 = add_entry(L, "encoder.head.6.weight", 1, D, D);
    L.h_b6 = add_entry(L, "encoder.head.6.bias", 1, D);
    if (c.model == DOF_MODEL_CONTRASTIVE) return L;
    const int fin[3] = {D, D, 2 * D}, fout[3] = {D, 2 * D, 4 * D};
    for (int i = 0; i < 3; i++) {
        const std::string s = std::to_string(i);
        L.dfc_w[i] = add_entry(L, "decoder.fc" + s + ".weight", 2, fout[i], fin[i]);
        L.dfc_b[i] = add_entry(L, "decoder.fc" + s + ".bias", 2, fout[i]);
        add_bn_entries(L, "decoder.bn" + s + ".", 2, fout[i], L.dbn[i]);
    }
    add_tcn_stack(L, "decoder.tcn.", 2, 4 * D, 64, dec_d, 4, L.tstack[2]);
    L.loc_w = add_entry(L, "decoder.prob_decoder.loc_projection.weight", 2, N * c.F, 64);
    L.loc_b = add_entry(L, "decoder.prob_decoder.loc_projection.bias", 2, N * c.F);
    add_latent_tail(L, c);
    return L;
}

static Layout build_layout(const dof_config& c) {
    if (c.encoder == DOF_ENCODER_TRANSFORMER) return build_layout_tfm(c);
    if (c.encoder == DOF_ENCODER_TCN) return build_layout_tcn(c);
    Layout L;
    const int N = c.N, E = c.E, D = c.D, di = dint(c);
    L.lap = add_entry(L, "encoder.laplacian", 0, N, N);
    L.elap = add_entry(L, "encoder.edge_laplacian", 0, E, E);
    L.inc = add_entry(L, "encoder.incidence", 0, N, E);
    const char* bn[2] = {"encoder.node_recurrent_block.", "encoder.edge_recurrent_block."};
    for (int b = 0; b < 2; b++) {
        std::string p = bn[b];
        int Fin = b == 0 ? c.F : c.Fe;
        BlockP& B = L.blk[b];
        B.conv = add_entry(L, p + "conv1d.weight", 1, 2 * di, Fin, 5);
        add_gru(L, p + "gru1.", 1, 2 * di, 2 * di, B.g1);
        B.n1w = add_entry(L, p + "norm1.weight", 1, 4 * di);
        B.n1b = add_entry(L, p + "norm1.bias", 1, 4 * di);
        add_gru(L, p + "gru2.", 1, 4 * di, di, B.g2);
        B.n2w = add_entry(L, p + "norm2.weight", 1, 2 * di);
        B.n2b = add_entry(L, p + "norm2.bias", 1, 2 * di);
        int pg = (di != D) ? 1 : 0;   // dead parameter when internal_dim == latent_dim (models_new.py:274)
        B.pw = add_entry(L, p + "projection.weight", pg, 2 * D, 2 * di);
        B.pb = add_entry(L, p + "projection.bias", pg, 2 * D);
    }
    std::string g = "encoder.spatial_gnn_block.";
    L.node_kernel = add_entry(L, g + "node_kernel", 1, 2 * D, D);
    L.edge_kernel = add_entry(L, g + "edge_kernel", 1, 2 * D, D);
    L.node_weights = add_entry(L, g + "node_weights", 1, 2 * D, 1);
    L.edge_weights = add_entry(L, g + "edge_weights", 1, 2 * D, 1);
    L.node_bias = add_entry(L, g + "node_bias", 1, D);
    L.edge_bias = add_entry(L, g + "edge_bias", 1, D);
    L.final_w = add_entry(L, "encoder.final_dense.weight", 1, D, (N + E) * D);
    L.final_b = add_entry(L, "encoder.final_dense.bias", 1, D);
    if (c.model == DOF_MODEL_CONTRASTIVE) return L;          // ContrastivePT: encoder only (models_new.py:2030-2057)
    add_gru(L, "decoder.gru1.", 2, D, D, L.dg1);
    L.dn1w = add_entry(L, "decoder.norm1.weight", 2, 2 * D);
    L.dn1b = add_entry(L, "decoder.norm1.bias", 2, 2 * D);
    add_gru(L, "decoder.gru2.", 2, 2 * D, 2 * D, L.dg2);
    L.dn2w = add_entry(L, "decoder.norm2.weight", 2, 4 * D);
    L.dn2b = add_entry(L, "decoder.norm2.bias", 2, 4 * D);
    L.dconv = add_entry(L, "decoder.conv1d.weight", 2, 2 * D, 4 * D, 5);
    L.dn3w = add_entry(L, "decoder.norm3.weight", 2, 2 * D);
    L.dn3b = add_entry(L, "decoder.norm3.bias", 2, 2 * D);
    L.loc_w = add_entry(L, "decoder.prob_decoder.loc_projection.weight", 2, N * c.F, 2 * D);
    L.loc_b = add_entry(L, "decoder.prob_decoder.loc_projection.bias", 2, N * c.F);
    add_latent_tail(L, c);
    return L;
}

static int check_cfg(const dof_config* c) {
    if (!c) DOF_FAIL(DOF_ERR_ARG, "null config");
    if (c->T < 1 || c->N < 1 || c->E < 1 || c->F < 1 || c->Fe < 1 || c->D < 1 || c->K < 1)
        DOF_FAIL(DOF_ERR_ARG, "bad geometry T=%d N=%d E=%d F=%d Fe=%d D=%d K=%d", c->T, c->N, c->E, c->F,
                 c->Fe, c->D, c->K);
    if (c->model < DOF_MODEL_VADE || c->model > DOF_MODEL_CONTRASTIVE) DOF_FAIL(DOF_ERR_ARG, "unknown model kind %d", c->model);
    if (c->model == DOF_MODEL_VADE && c->K > LOSS_MAXK) DOF_FAIL(DOF_ERR_UNSUPPORTED, "n_components %d > %d", c->K, LOSS_MAXK);
    if (c->model == DOF_MODEL_VQVAE && (size_t)c->D * c->K * 4 > 96 * 1024) DOF_FAIL(DOF_ERR_UNSUPPORTED, "codebook %d x %d does not fit in shared memory", c->D, c->K);
    if (c->encoder != DOF_ENCODER_RECURRENT && c->encoder != DOF_ENCODER_TRANSFORMER && c->encoder != DOF_ENCODER_TCN)
        DOF_FAIL(DOF_ERR_ARG, "unknown encoder kind %d", c->encoder);
    if (c->encoder == DOF_ENCODER_TCN && c->D > 64) DOF_FAIL(DOF_ERR_UNSUPPORTED, "latent_dim %d > 64 not supported by the TCN path", c->D);
    if (c->encoder == DOF_ENCODER_RECURRENT && c->D > 32) DOF_FAIL(DOF_ERR_UNSUPPORTED, "latent_dim %d > 32 not supported by the GRU kernels yet", c->D);
    if (c->encoder == DOF_ENCODER_TRANSFORMER) {
        if (c->D > 64) DOF_FAIL(DOF_ERR_UNSUPPORTED, "latent_dim %d > 64 not supported by the transformer path", c->D);
        if (c->T > TFM_MAXT) DOF_FAIL(DOF_ERR_UNSUPPORTED, "window length %d > %d", c->T, TFM_MAXT);
        if (c->F > 4 || c->Fe > 4) DOF_FAIL(DOF_ERR_UNSUPPORTED, "more than 4 features per node / edge");
        if ((4 * c->D) % 8) DOF_FAIL(DOF_ERR_ARG, "decoder model_dim %d is not a multiple of 8 heads", 4 * c->D);
    }
    return DOF_OK;
}

// ---------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------
struct Bump {
    char* base; size_t off, cap;
    bool dry;
    template <typename T> T* get(size_t n) {
        size_t bytes = (n * sizeof(T) + 255) / 256 * 256;
        T* p = dry ? nullptr : reinterpret_cast<T*>(base + off);
        off += bytes;
        return p;
    }
};

struct BlockWS {
    int S, Fin, G;
    int* gidx;
    float *Xs, *Cv; int* len;
    float *Gi1[2], *H1, *Gt1[2], *mu1, *rs1, *Y1;
    float *Gi2[2], *H2, *Gt2[2], *Hn, *mu2, *rs2, *Y2, *P;
    float *dP, *dY2, *dHn, *dG2[2], *dY1, *dH1, *dG1[2], *dCv;
};

// transformer family: per (node | edge) core activations.  R = S * T rows; the last layer keeps its post-attention
// tensors for the final step of every sequence only ([S, *]).
struct TfmLayerWS { float *QKV, *ATT, *U1, *mu1, *rs1, *Y1, *FF, *U2, *mu2, *rs2, *Y2; };
struct TfmCoreWS {
    int S, G, Fin;
    int* gidx;
    float* Xs; unsigned char* kpad; float* Y0;
    TfmLayerWS l[TFM_MAXL];
    float *dA, *dB, *dC, *dFF, *dQKV;        // [R, *] gradient scratch
    float *sA, *sB, *sC, *sFF, *dOut;        // [S, *] gradient scratch of the last layer, dOut = d(core output)
};
struct TfmDecLayerWS { float *Xn1, *QKV, *ATT, *Hmid, *Xn2, *PRE, *FH, *Hout, *muA, *rsA, *muB, *rsB; };
struct TfmDecWS {
    float *P[3], *G[3];                      // latent expansion: pre-activations and GELU outputs [B, *]
    float *H0, *tmpA, *tmpB, *tmpF, *Y;      // [R, dm] x3, [R, dff], [R, Dx]
    TfmDecLayerWS l[TFM_MAXL];
    float *dH, *dQKV, *dY, *dP[3], *dG[3];
};

// TCN family: activations of one TCN1DPT stack.  R = sequences * T rows; A1 / A2 are the pre-BatchNorm convolution outputs,
// Y1 = relu(bn1(A1)), OUT the block output (input of the next block), RES the 1x1 residual projection of block 0.
struct TcnBlockWS { float *A1, *Y1, *A2, *OUT, *RES; int st1, st2; };     // st*: offsets of the BatchNorm sums in a statistics pass
struct TcnStackWS {
    int S, G, Fin;                           // encoder stacks: sequences at max_batch, graph size, features
    int* gidx;
    long long rows;                          // rows at max_batch
    int x0_ld;                               // row pitch of X0: cin0 rounded up to a multiple of 4 (zero pad columns)
    float* X0;                               // [rows, x0_ld]
    TcnBlockWS b[TCN_MAXB];
    float *SKIP, *FIN;                       // skip sum and relu(skip sum): [S, C] (encoder, last step only) or [rows, C]
    float *GS, *DA, *DB, *DX, *DX0;          // backward scratch
};

struct dof_handle {
    dof_config cfg;
    Layout L;
    // TCN family
    TcnStackWS ts[3];
    double *tstat = nullptr, *tbstat = nullptr;   // [2 passes][tstat_stride] sums | sums of squares; backward sums
    long long tstat_stride = 0;
    TcnBnDesc* tdesc = nullptr; int tn_desc = 0;
    int tcn_enc_windows = 0, tcn_dec_passes = 0, tcn_dec_windows = 0;
    float *dgz, *drms, *dfo[3], *dzo[3], *ddz[3], *ddf[3], *ddg;   // decoder front MLP: fc outputs, BatchNorm outputs, their gradients
    float *dbnm[3], *dbns[3], *dbnstat[3];
    // transformer family
    TfmCoreWS tc[2];
    TfmDecWS td;
    float *pe_enc, *pe_dec;                  // positional encodings [T, dk], [T, 4D]
    float *hIn, *hRms, *h1r, *h1, *h2r, *h2, *h3, *dh3, *dh2, *dh2r, *dh1, *dh1r, *dhIn;   // encoder head
    float *bnm[3], *bns[3], *bnstat[2]; unsigned char* bclamp;
    int enc_groups = 1;                      // statistics groups of the last encoder pass (contrastive: 2)
    int next_groups = 1;                     // statistics groups of the NEXT training-mode encoder pass
    int dec_pass = 0;                        // which decoder pass of the step comes next (VQ-VAE decodes twice)
    int bn_pending = 0;                      // batch statistics of the last training step wait for dof_clip_adam
    unsigned long long drop_seed = 0;
    const unsigned char* drop_masks = nullptr;
    size_t drop_mask_bytes = 0;
    unsigned long long noise_seed = 0x9E3779B97F4A7C15ull;   // Philox key of the VaDE noise when eps / mc_eps are NULL
    int device, sm_count, max_batch, training;
    int di, H1, H2, C1;
    BlockWS blk[2];
    unsigned char* group;   // [state numel]
    // mid
    float *Pn, *Pe, *On, *Oe, *enc, *zm, *pre, *lv, *z, *q;
    float *dOn, *dOe, *dPn, *dPe, *denc, *dzm, *dpre, *dz_dec;
    // decoder
    int* lenD;
    float *GiD1[2], *HD1, *GtD1[2], *muD1, *rsD1, *YD1;
    float *GiD2[2], *HD2, *GtD2[2], *muD2, *rsD2, *YD2, *Cd, *muD3, *rsD3, *YD3, *loc;
    float *dlocP = nullptr;    // dloc with its row pitch padded to a multiple of 4 floats (tensor-core weight gradients), or null
    float *dloc, *dYD3, *dCd, *dYD2, *dHD2, *dGD2[2], *dYD1, *dHD1, *dGD1[2], *dGs[2], *Wt;
    // loss
    double* stats; float *coef, *dzm_kl, *dlv_kl, *distw;
    // VQ-VAE
    float *quant, *soft; int* vidx; double* vstats;
    // contrastive
    float *zn, *nrm, *lse; double* nstats;
    int lastB;
    // the node and the edge recurrent blocks are independent until CensNet: the edge block runs on a second stream
    // so that its kernels fill the SMs the node block's last wave leaves idle (896 CTAs on 148 SMs)
    cudaStream_t aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    std::map<std::string, std::pair<const void*, int64_t>> dbg;
};

static bool g_concurrent = true;

template <typename F>
static int fork_join_blocks(dof_handle* h, cudaStream_t st, F fn) {
    if (!h->aux || !g_concurrent) { DOF_TRY(fn(0, st)); return fn(1, st); }
    DOF_CUDA(cudaEventRecord(h->ev_fork, st));
    DOF_CUDA(cudaStreamWaitEvent(h->aux, h->ev_fork, 0));
    int r0 = fn(0, st);
    int r1 = fn(1, h->aux);
    DOF_CUDA(cudaEventRecord(h->ev_join, h->aux));
    DOF_CUDA(cudaStreamWaitEvent(st, h->ev_join, 0));
    return r0 != DOF_OK ? r0 : r1;
}

static void plan_workspace_head(dof_handle* h, Bump& bp);
static void plan_workspace_tfm_encoder(dof_handle* h, Bump& bp) {
    const dof_config& c = h->cfg;
    const Layout& L = h->L;
    const int B = h->max_batch, T = c.T, dk = L.dk, dff = L.dff;
    const bool tr = h->training != 0;
    h->pe_enc = bp.get<float>((size_t)T * dk);
    for (int b = 0; b < 2; b++) {
        TfmCoreWS& w = h->tc[b];
        w.G = b == 0 ? c.N : c.E;
        w.Fin = b == 0 ? c.F : c.Fe;
        w.S = B * w.G;
        const size_t S = (size_t)w.S, R = S * T;
        w.gidx = bp.get<int>((size_t)w.G * T * w.Fin);
        w.Xs = bp.get<float>(R * w.Fin);
        w.kpad = bp.get<unsigned char>(R);
        w.Y0 = bp.get<float>(R * dk);
        for (int l = 0; l < L.layers; l++) {
            TfmLayerWS& q = w.l[l];
            const size_t Rl = (l == L.layers - 1) ? S : R;       // rows of the post-attention tensors
            q.QKV = bp.get<float>(R * 3 * dk);
            q.ATT = bp.get<float>(Rl * dk);
            q.U1 = bp.get<float>(Rl * dk); q.mu1 = bp.get<float>(Rl); q.rs1 = bp.get<float>(Rl);
            q.Y1 = bp.get<float>(Rl * dk);
            q.FF = bp.get<float>(Rl * dff);
            q.U2 = bp.get<float>(Rl * dk); q.mu2 = bp.get<float>(Rl); q.rs2 = bp.get<float>(Rl);
            q.Y2 = bp.get<float>(Rl * dk);
        }
        if (tr) {
            w.dA = bp.get<float>(R * dk); w.dB = bp.get<float>(R * dk); w.dC = bp.get<float>(R * dk);
            w.dFF = bp.get<float>(R * dff); w.dQKV = bp.get<float>(R * 3 * dk);
            w.sA = bp.get<float>(S * dk); w.sB = bp.get<float>(S * dk); w.sC = bp.get<float>(S * dk);
            w.sFF = bp.get<float>(S * dff); w.dOut = bp.get<float>(S * dk);
        }
    }
    plan_workspace_head(h, bp);
}

// encoder head shared by the transformer and the TCN encoders (RMS normalisation, Dense + BatchNorm x 2, Dense)
static void plan_workspace_head(dof_handle* h, Bump& bp) {
    const dof_config& c = h->cfg;
    const int B = h->max_batch, D = c.D;
    const bool tr = h->training != 0;
    const size_t Bz = (size_t)B, KD = (size_t)(c.N + c.E) * D;
    h->hIn = bp.get<float>(Bz * KD); h->hRms = bp.get<float>(Bz);
    h->h1r = bp.get<float>(Bz * 2 * D); h->h1 = bp.get<float>(Bz * 2 * D);
    h->h2r = bp.get<float>(Bz * D); h->h2 = bp.get<float>(Bz * D); h->h3 = bp.get<float>(Bz * D);
    const int cw[3] = {2 * D, D, D};
    for (int i = 0; i < 3; i++) { h->bnm[i] = bp.get<float>((size_t)2 * cw[i]); h->bns[i] = bp.get<float>((size_t)2 * cw[i]); }
    h->bnstat[0] = bp.get<float>((size_t)2 * 2 * 2 * D); h->bnstat[1] = bp.get<float>((size_t)2 * 2 * D);
    h->bclamp = bp.get<unsigned char>((size_t)2 * D);
    if (tr) {
        h->dh3 = bp.get<float>(Bz * D); h->dh2 = bp.get<float>(Bz * D); h->dh2r = bp.get<float>(Bz * D);
        h->dh1 = bp.get<float>(Bz * 2 * D); h->dh1r = bp.get<float>(Bz * 2 * D); h->dhIn = bp.get<float>(Bz * KD);
    }
}

static void plan_workspace_tfm_decoder(dof_handle* h, Bump& bp) {
    const dof_config& c = h->cfg;
    const Layout& L = h->L;
    const int B = h->max_batch, T = c.T, D = c.D, dm = 4 * D, dff = L.dec_dff, Dx = c.N * c.F;
    const bool tr = h->training != 0;
    const size_t Bz = (size_t)B, R = Bz * T;
    TfmDecWS& w = h->td;
    const int eout[3] = {D, 2 * D, 4 * D};
    h->pe_dec = bp.get<float>((size_t)T * dm);
    for (int i = 0; i < 3; i++) { w.P[i] = bp.get<float>(Bz * eout[i]); w.G[i] = bp.get<float>(Bz * eout[i]); }
    w.H0 = bp.get<float>(R * dm); w.tmpA = bp.get<float>(R * dm); w.tmpB = bp.get<float>(R * dm);
    w.tmpF = bp.get<float>(R * dff); w.Y = bp.get<float>(R * round_up(Dx, 4));
    for (int l = 0; l < L.dec_layers; l++) {
        TfmDecLayerWS& q = w.l[l];
        const bool keep = tr || l == 0;                          // eval: every layer reuses layer 0's buffers
        if (!keep) { q = w.l[0]; continue; }
        q.Xn1 = bp.get<float>(R * dm); q.QKV = bp.get<float>(R * 3 * dm); q.ATT = bp.get<float>(R * dm);
        q.Hmid = bp.get<float>(R * dm); q.Xn2 = bp.get<float>(R * dm); q.PRE = bp.get<float>(R * dff);
        q.FH = bp.get<float>(R * dff); q.Hout = bp.get<float>(R * dm);
        q.muA = bp.get<float>(R); q.rsA = bp.get<float>(R); q.muB = bp.get<float>(R); q.rsB = bp.get<float>(R);
    }
    if (tr) {
        w.dH = bp.get<float>(R * dm); w.dQKV = bp.get<float>(R * 3 * dm); w.dY = bp.get<float>(R * round_up(Dx, 4));
        for (int i = 0; i < 3; i++) { w.dP[i] = bp.get<float>(Bz * eout[i]); w.dG[i] = bp.get<float>(Bz * eout[i]); }
    }
}

// one TCN1DPT stack: eval keeps one set of block buffers (every block reuses them), training keeps all of them
static void plan_tcn_stack(dof_handle* h, Bump& bp, int si, long long rows, long long seqs, bool last_only, bool need_dx0, int& stat_off) {
    const TcnStackP& P = h->L.tstack[si];
    TcnStackWS& w = h->ts[si];
    const bool tr = h->training != 0;
    const size_t R = (size_t)rows, C = (size_t)P.C;
    w.rows = rows;
    w.x0_ld = round_up(P.cin0, 4);
    w.X0 = bp.get<float>(R * w.x0_ld);
    for (int i = 0; i < P.nb; i++) {
        TcnBlockWS& q = w.b[i];
        if (!tr && i >= 2) { q = w.b[i & 1]; }
        else {
            if (!tr && i == 1) { q.A1 = w.b[0].A1; q.Y1 = w.b[0].Y1; q.A2 = w.b[0].A2; q.RES = nullptr; }
            else {
                q.A1 = bp.get<float>(R * C); q.Y1 = bp.get<float>(R * C); q.A2 = bp.get<float>(R * C);
                q.RES = P.blk[i].has_ds ? bp.get<float>(R * C) : nullptr;
            }
            q.OUT = bp.get<float>(R * C);
        }
        q.st1 = stat_off; stat_off += 3 * P.C;
        q.st2 = stat_off; stat_off += 3 * P.C;
    }
    const size_t SR = last_only ? (size_t)seqs : R;
    w.SKIP = bp.get<float>(SR * C); w.FIN = bp.get<float>(SR * C);
    w.GS = w.DA = w.DB = w.DX = w.DX0 = nullptr;
    if (tr) {
        w.GS = bp.get<float>(SR * C);
        w.DA = bp.get<float>(R * C); w.DB = bp.get<float>(R * C); w.DX = bp.get<float>(R * C);
        if (need_dx0) w.DX0 = bp.get<float>(R * P.cin0);
    }
}

static void plan_workspace_tcn(dof_handle* h, Bump& bp) {
    const dof_config& c = h->cfg;
    const int B = h->max_batch, T = c.T, D = c.D;
    const bool tr = h->training != 0;
    int stat_off = 0;
    for (int b = 0; b < 2; b++) {
        TcnStackWS& w = h->ts[b];
        w.G = b == 0 ? c.N : c.E;
        w.Fin = b == 0 ? c.F : c.Fe;
        w.S = B * w.G;
        w.gidx = bp.get<int>((size_t)w.G * T * w.Fin);
        plan_tcn_stack(h, bp, b, (long long)w.S * T, w.S, true, false, stat_off);
    }
    if (c.model != DOF_MODEL_CONTRASTIVE) {
        h->ts[2].S = B; h->ts[2].G = 1; h->ts[2].Fin = 4 * D; h->ts[2].gidx = nullptr;
        plan_tcn_stack(h, bp, 2, (long long)B * T, B, false, true, stat_off);
        const size_t Bz = (size_t)B;
        const int fout[3] = {D, 2 * D, 4 * D};
        h->dgz = bp.get<float>(Bz * D); h->drms = bp.get<float>(Bz);
        for (int i = 0; i < 3; i++) {
            h->dfo[i] = bp.get<float>(Bz * fout[i]); h->dzo[i] = bp.get<float>(Bz * fout[i]);
            h->dbnm[i] = bp.get<float>((size_t)fout[i]); h->dbns[i] = bp.get<float>((size_t)fout[i]);
            h->dbnstat[i] = bp.get<float>((size_t)2 * 2 * fout[i]);
            if (tr) { h->ddz[i] = bp.get<float>(Bz * fout[i]); h->ddf[i] = bp.get<float>(Bz * fout[i]); }
        }
        if (tr) h->ddg = bp.get<float>(Bz * D);
    }
    h->tstat_stride = stat_off;
    h->tstat = bp.get<double>((size_t)2 * stat_off);
    h->tbstat = bp.get<double>((size_t)stat_off);
    int nd = 0;
    for (int si = 0; si < 3; si++) if (si < 2 || c.model != DOF_MODEL_CONTRASTIVE) nd += 2 * h->L.tstack[si].nb;
    h->tn_desc = nd;
    h->tdesc = bp.get<TcnBnDesc>((size_t)nd);
    plan_workspace_head(h, bp);
}

static void plan_workspace(dof_handle* h, Bump& bp) {
    const dof_config& c = h->cfg;
    const int B = h->max_batch, T = c.T, D = c.D, K = c.K, N = c.N, E = c.E;
    const int H1 = h->H1, H2 = h->H2, C1 = h->C1;
    const bool tr = h->training != 0;
    const bool tcn = c.encoder == DOF_ENCODER_TCN;
    const bool tfm = c.encoder == DOF_ENCODER_TRANSFORMER || tcn;     // "not recurrent": the transformer-style head and buffers
    const int CC = tfm ? h->L.dk : 2 * D;                        // CensNet input channels
    h->group = bp.get<unsigned char>((size_t)h->L.total);
    if (tcn) plan_workspace_tcn(h, bp);
    else if (tfm) plan_workspace_tfm_encoder(h, bp);
    for (int b = 0; b < 2 && !tfm; b++) {
        BlockWS& w = h->blk[b];
        w.G = b == 0 ? N : E;
        w.Fin = b == 0 ? c.F : c.Fe;
        w.S = B * w.G;
        const size_t S = (size_t)w.S, ST = S * T;
        w.gidx = bp.get<int>((size_t)w.G * T * w.Fin);
        w.Xs = bp.get<float>(ST * w.Fin);
        w.Cv = bp.get<float>(ST * C1);
        w.len = bp.get<int>(S);
        for (int d = 0; d < 2; d++) w.Gi1[d] = bp.get<float>(ST * 3 * H1);
        w.H1 = bp.get<float>(ST * 2 * H1);
        for (int d = 0; d < 2; d++) w.Gt1[d] = tr ? bp.get<float>((size_t)round_up(w.S, 128) * T * 4 * H1) : nullptr;   // tiled layout: whole tiles
        w.mu1 = bp.get<float>(ST); w.rs1 = bp.get<float>(ST);
        w.Y1 = bp.get<float>(ST * 2 * H1);
        for (int d = 0; d < 2; d++) w.Gi2[d] = bp.get<float>(ST * 3 * H2);
        w.H2 = tr ? bp.get<float>(ST * 2 * H2) : nullptr;
        for (int d = 0; d < 2; d++) w.Gt2[d] = tr ? bp.get<float>((size_t)round_up(w.S, 128) * T * 4 * H2) : nullptr;
        w.Hn = bp.get<float>(S * 2 * H2);
        w.mu2 = bp.get<float>(S); w.rs2 = bp.get<float>(S);
        w.Y2 = bp.get<float>(S * 2 * H2);
        w.P = (h->di != D) ? bp.get<float>(S * 2 * D) : w.Y2;
        if (tr) {
            w.dP = bp.get<float>(S * 2 * D);
            w.dY2 = (h->di != D) ? bp.get<float>(S * 2 * H2) : w.dP;
            w.dHn = bp.get<float>(S * 2 * H2);
            for (int d = 0; d < 2; d++) w.dG2[d] = bp.get<float>(ST * 4 * H2);
            w.dY1 = bp.get<float>(ST * 2 * H1);
            w.dH1 = bp.get<float>(ST * 2 * H1);
            for (int d = 0; d < 2; d++) w.dG1[d] = bp.get<float>(ST * 4 * H1);
            w.dCv = bp.get<float>(ST * C1);
        }
    }
    const size_t Bz = (size_t)B, BT = Bz * T;
    const int model = c.model;
    h->Pn = bp.get<float>(Bz * N * CC); h->Pe = bp.get<float>(Bz * E * CC);
    h->On = bp.get<float>(Bz * N * D); h->Oe = bp.get<float>(Bz * E * D);
    h->enc = bp.get<float>(Bz * D);
    if (tr) {
        h->dOn = bp.get<float>(Bz * N * D); h->dOe = bp.get<float>(Bz * E * D);
        h->dPn = bp.get<float>(Bz * N * CC); h->dPe = bp.get<float>(Bz * E * CC);
        h->denc = bp.get<float>(Bz * D);
    }
    if (model == DOF_MODEL_CONTRASTIVE) {
        h->zn = bp.get<float>(Bz * D); h->nrm = bp.get<float>(Bz); h->lse = bp.get<float>(4 * Bz);
        h->nstats = bp.get<double>(8);
        return;
    }
    if (model == DOF_MODEL_VADE) {
        h->zm = bp.get<float>(Bz * D); h->pre = bp.get<float>(Bz * D);
        h->lv = bp.get<float>(Bz * D); h->z = bp.get<float>(Bz * D); h->q = bp.get<float>(Bz * K);
    } else {
        h->quant = bp.get<float>(Bz * D); h->soft = bp.get<float>(Bz * K); h->vidx = bp.get<int>(Bz);
        h->vstats = bp.get<double>((size_t)VQ_ST_GRAM + (size_t)D * D + K);
    }
    h->lenD = bp.get<int>(Bz);
    if (tcn) {}
    else if (tfm) plan_workspace_tfm_decoder(h, bp);
    else {
    for (int d = 0; d < 2; d++) h->GiD1[d] = bp.get<float>(Bz * 3 * D);
    h->HD1 = bp.get<float>(BT * 2 * D);
    for (int d = 0; d < 2; d++) h->GtD1[d] = tr ? bp.get<float>((size_t)round_up(B, 128) * T * 4 * D) : nullptr;
    h->muD1 = bp.get<float>(BT); h->rsD1 = bp.get<float>(BT);
    h->YD1 = bp.get<float>(BT * 2 * D);
    for (int d = 0; d < 2; d++) h->GiD2[d] = bp.get<float>(BT * 6 * D);
    h->HD2 = bp.get<float>(BT * 4 * D);
    for (int d = 0; d < 2; d++) h->GtD2[d] = tr ? bp.get<float>((size_t)round_up(B, 128) * T * 8 * D) : nullptr;
    h->muD2 = bp.get<float>(BT); h->rsD2 = bp.get<float>(BT);
    h->YD2 = bp.get<float>(BT * 4 * D);
    h->Cd = bp.get<float>(BT * 2 * D);
    h->muD3 = bp.get<float>(BT); h->rsD3 = bp.get<float>(BT);
    h->YD3 = bp.get<float>(BT * 2 * D);
    }
    h->loc = bp.get<float>(BT * N * c.F);
    if (tr) {
        if (model == DOF_MODEL_VADE) { h->dzm = bp.get<float>(Bz * D); h->dpre = bp.get<float>(Bz * D); }
        h->dz_dec = bp.get<float>(Bz * D);
        h->dloc = bp.get<float>(BT * N * c.F);
        h->dlocP = (N * c.F) & 3 ? bp.get<float>(BT * round_up(N * c.F, 4)) : nullptr;
        if (!tfm) {
        h->dYD3 = bp.get<float>(BT * 2 * D); h->dCd = bp.get<float>(BT * 2 * D);
        h->dYD2 = bp.get<float>(BT * 4 * D); h->dHD2 = bp.get<float>(BT * 4 * D);
        for (int d = 0; d < 2; d++) h->dGD2[d] = bp.get<float>(BT * 8 * D);
        h->dYD1 = bp.get<float>(BT * 2 * D); h->dHD1 = bp.get<float>(BT * 2 * D);
        for (int d = 0; d < 2; d++) h->dGD1[d] = bp.get<float>(BT * 4 * D);
        for (int d = 0; d < 2; d++) h->dGs[d] = bp.get<float>(Bz * 4 * D);
        h->Wt = bp.get<float>((size_t)4 * D * 2 * D * 5);
        }
        if (model != DOF_MODEL_VADE) return;
        StatsLayout SL = stats_layout(D, K);
        CoefLayout CL = coef_layout(D, K);
        h->stats = bp.get<double>(SL.total);
        h->coef = bp.get<float>(CL.total);
        h->dzm_kl = bp.get<float>(Bz * D); h->dlv_kl = bp.get<float>(Bz * D);
        h->distw = bp.get<float>(Bz);
    }
}

static void init_derived(dof_handle* h) {
    h->di = dint(h->cfg);
    h->H1 = 2 * h->di; h->H2 = h->di; h->C1 = 2 * h->di;
}

// index of (g,t',f) inside one window flattened [T,G,F]
static void build_gidx(int T, int G, int F, std::vector<int>& out) {
    out.resize((size_t)G * T * F);
    for (int g = 0; g < G; g++)
        for (int t2 = 0; t2 < T; t2++)
            for (int f = 0; f < F; f++) {
                long long lin = ((long long)f * T + t2) * G + g;
                int j = (int)(lin / T), t = (int)(lin % T);
                out[((size_t)g * T + t2) * F + f] = t * (G * F) + j;
            }
}

extern "C" {

int dof_abi_version(void) { return DOF_ABI_VERSION; }
#ifndef DOF_SOURCE_HASH
#define DOF_SOURCE_HASH "unknown"
#endif
// sha256 (first 16 hex digits) of csrc/* and include/deepof_b200.h at build time: __graft_entry__.build() rebuilds when the
// sources on disk hash differently, so a prebuilt library can never silently disagree with the tree it ships in
static const char g_source_hash[] = "DOF_SOURCE_HASH=" DOF_SOURCE_HASH;      // the marker lets build() read it without dlopen
const char* dof_source_hash(void) { return g_source_hash + 16; }
const char* dof_last_error(void) { return g_dof_err; }

int64_t dof_state_numel(const dof_config* cfg) {
    if (check_cfg(cfg) != DOF_OK) return -1;
    return build_layout(*cfg).total;
}
int dof_state_num_entries(const dof_config* cfg) {
    if (check_cfg(cfg) != DOF_OK) return -1;
    return (int)build_layout(*cfg).e.size();
}
int dof_state_entry(const dof_config* cfg, int index, char* name_out, int64_t* offset_out, int64_t* numel_out,
                    int* ndim_out, int* shape_out, int* group_out) {
    DOF_TRY(check_cfg(cfg));
    Layout L = build_layout(*cfg);
    if (index < 0 || index >= (int)L.e.size()) DOF_FAIL(DOF_ERR_ARG, "entry index %d out of range", index);
    const Entry& e = L.e[index];
    if (name_out) { strncpy(name_out, e.name.c_str(), 127); name_out[127] = 0; }
    if (offset_out) *offset_out = e.off;
    if (numel_out) *numel_out = e.numel;
    if (ndim_out) *ndim_out = e.ndim;
    if (shape_out) for (int i = 0; i < 4; i++) shape_out[i] = e.shape[i];
    if (group_out) *group_out = e.group;
    return DOF_OK;
}

// censNetConv_pt.py:160-175 (gcn_filter :182-243, incidence :296-370, line graph :258-279)
static void gcn_filter_host(const std::vector<double>& A, int n, std::vector<double>& out) {
    std::vector<double> Ah(A), deg(n, 0.0);
    for (int i = 0; i < n; i++) Ah[(size_t)i * n + i] += 1.0;
    for (int i = 0; i < n; i++) {
        double s = 0.0;
        for (int j = 0; j < n; j++) s += Ah[(size_t)i * n + j];
        deg[i] = (s == 0.0) ? 1.0 : s;
    }
    out.resize((size_t)n * n);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) out[(size_t)i * n + j] = pow(deg[i], -0.5) * Ah[(size_t)i * n + j] * pow(deg[j], -0.5);
}

int dof_graph_operators(const double* adjacency, int N, int max_edges, float* laplacian, float* edge_laplacian,
                        float* incidence, int* n_edges_out) {
    if (!adjacency || N < 1) DOF_FAIL(DOF_ERR_ARG, "bad adjacency");
    std::vector<std::pair<int, int>> edges;
    for (int i = 0; i < N; i++)
        for (int j = i; j < N; j++)
            if (adjacency[(size_t)i * N + j] != 0.0) edges.push_back({i, j});
    const int E = (int)edges.size();
    if (n_edges_out) *n_edges_out = E;
    if (!laplacian) return DOF_OK;   // query only
    if (E > max_edges) DOF_FAIL(DOF_ERR_ARG, "%d edges > max_edges %d", E, max_edges);
    std::vector<double> A(adjacency, adjacency + (size_t)N * N), lap;
    gcn_filter_host(A, N, lap);
    for (size_t i = 0; i < lap.size(); i++) laplacian[i] = (float)lap[i];
    std::vector<double> inc((size_t)N * E, 0.0);
    for (int e = 0; e < E; e++) { inc[(size_t)edges[e].first * E + e] = 1.0; inc[(size_t)edges[e].second * E + e] = 1.0; }
    for (size_t i = 0; i < inc.size(); i++) incidence[i] = (float)inc[i];
    std::vector<double> line((size_t)E * E, 0.0), elap;
    for (int e = 0; e < E; e++)
        for (int f = 0; f < E; f++) {
            double s = 0.0;
            for (int n = 0; n < N; n++) s += inc[(size_t)n * E + e] * inc[(size_t)n * E + f];
            line[(size_t)e * E + f] = s - (e == f ? 2.0 : 0.0);
        }
    gcn_filter_host(line, E, elap);
    for (size_t i = 0; i < elap.size(); i++) edge_laplacian[i] = (float)elap[i];
    return DOF_OK;
}

size_t dof_workspace_bytes(const dof_config* cfg, int max_batch, int training) {
    if (check_cfg(cfg) != DOF_OK || max_batch < 1) return 0;
    dof_handle h;
    h.cfg = *cfg; h.L = build_layout(*cfg); h.max_batch = max_batch; h.training = training;
    init_derived(&h);
    Bump bp{nullptr, 0, 0, true};
    plan_workspace(&h, bp);
    return bp.off + 256;
}

int dof_create(const dof_config* cfg, int device, int max_batch, int training, void* workspace,
               size_t workspace_bytes, dof_handle** out) {
    DOF_TRY(check_cfg(cfg));
    if (!out || !workspace || max_batch < 1) DOF_FAIL(DOF_ERR_ARG, "null output / workspace or bad max_batch");
    int ndev = 0;
    DOF_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) DOF_FAIL(DOF_ERR_ARG, "device %d not available (%d devices)", device, ndev);
    cudaDeviceProp prop;
    DOF_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) DOF_FAIL(DOF_ERR_UNSUPPORTED, "deepof_b200 needs an sm_100-class GPU, found sm_%d%d", prop.major, prop.minor);
    DOF_CUDA(cudaSetDevice(device));
    dof_handle* h = new dof_handle();
    h->cfg = *cfg; h->L = build_layout(*cfg); h->device = device; h->sm_count = prop.multiProcessorCount;
    h->max_batch = max_batch; h->training = training; h->lastB = 0;
    g_sm_count = h->sm_count;
    init_derived(h);
    size_t need = dof_workspace_bytes(cfg, max_batch, training);
    if (workspace_bytes < need) { delete h; DOF_FAIL(DOF_ERR_WORKSPACE, "workspace %zu bytes < required %zu", workspace_bytes, need); }
    uintptr_t base = ((uintptr_t)workspace + 255) / 256 * 256;
    Bump bp{reinterpret_cast<char*>(base), 0, workspace_bytes, false};
    plan_workspace(h, bp);
    // constant tables
    std::vector<unsigned char> grp((size_t)h->L.total, 0);
    for (const Entry& e : h->L.e)
        for (int64_t i = 0; i < e.numel; i++) grp[(size_t)(e.off + i)] = (unsigned char)e.group;
    cudaError_t ce = cudaMemcpy(h->group, grp.data(), grp.size(), cudaMemcpyHostToDevice);
    const bool tfm = cfg->encoder == DOF_ENCODER_TRANSFORMER, tcn = cfg->encoder == DOF_ENCODER_TCN;
    for (int b = 0; b < 2 && ce == cudaSuccess; b++) {
        std::vector<int> gi;
        const int G = tcn ? h->ts[b].G : tfm ? h->tc[b].G : h->blk[b].G, Fin = tcn ? h->ts[b].Fin : tfm ? h->tc[b].Fin : h->blk[b].Fin;
        build_gidx(cfg->T, G, Fin, gi);
        ce = cudaMemcpy(tcn ? h->ts[b].gidx : tfm ? h->tc[b].gidx : h->blk[b].gidx, gi.data(), gi.size() * sizeof(int), cudaMemcpyHostToDevice);
    }
    if (tcn && ce == cudaSuccess) {                      // BatchNorm descriptor table of the TCN blocks (running-buffer update)
        std::vector<TcnBnDesc> dv;
        for (int si = 0; si < 3; si++) {
            if (si == 2 && cfg->model == DOF_MODEL_CONTRASTIVE) break;
            const TcnStackP& P = h->L.tstack[si];
            for (int i = 0; i < P.nb; i++)
                for (int k = 0; k < 2; k++) {
                    const TfmBnP& bn = k == 0 ? P.blk[i].bn1 : P.blk[i].bn2;
                    TcnBnDesc d;
                    d.mean = bn.mean; d.var = bn.var; d.tracked = bn.tracked; d.C = P.C; d.decoder = si == 2 ? 1 : 0;
                    d.stat = k == 0 ? h->ts[si].b[i].st1 : h->ts[si].b[i].st2;
                    d.rows_per_window = (si == 0 ? cfg->N : si == 1 ? cfg->E : 1) * cfg->T;
                    dv.push_back(d);
                }
        }
        ce = cudaMemcpy(h->tdesc, dv.data(), dv.size() * sizeof(TcnBnDesc), cudaMemcpyHostToDevice);
    }
    if (tfm && ce == cudaSuccess) {
        tfm_pe_kernel<<<cdiv(cfg->T * h->L.dk, 256), 256>>>(h->pe_enc, cfg->T, h->L.dk);
        if (cfg->model != DOF_MODEL_CONTRASTIVE) tfm_pe_kernel<<<cdiv(cfg->T * 4 * cfg->D, 256), 256>>>(h->pe_dec, cfg->T, 4 * cfg->D);
        ce = cudaDeviceSynchronize();
    }
    if (ce != cudaSuccess) { delete h; DOF_FAIL(DOF_ERR_CUDA, "table upload failed: %s", cudaGetErrorString(ce)); }
    const char* ss = getenv("DOF_SINGLE_STREAM");
    if (!(ss && ss[0] == '1')) {
        if (cudaStreamCreateWithFlags(&h->aux, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess) {
            delete h;
            DOF_FAIL(DOF_ERR_CUDA, "could not create the auxiliary stream");
        }
    }
    *out = h;
    return DOF_OK;
}

int dof_destroy(dof_handle* h) {
    if (h) {
        if (h->aux) { cudaStreamSynchronize(h->aux); cudaStreamDestroy(h->aux); }
        if (h->ev_fork) cudaEventDestroy(h->ev_fork);
        if (h->ev_join) cudaEventDestroy(h->ev_join);
    }
    delete h;
    return DOF_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
static int ln_fwd(const float* x, const float* w, const float* b, float* y, float* mu, float* rs, long long R,
                  int W, int sm, cudaStream_t st, float eps = 1e-3f) {
    if (W > 32 * LN_MAXV) DOF_FAIL(DOF_ERR_UNSUPPORTED, "LayerNorm width %d > %d", W, 32 * LN_MAXV);
    long long blocks = (R + 7) / 8;
    int grid = (int)(blocks < (long long)sm * 16 ? blocks : (long long)sm * 16);
    if (grid < 1) return DOF_OK;
    { ProfScope ps("ln_fwd", st, 0.0, 8.0 * R * W);
    const bool vec = ((((uintptr_t)x) | ((uintptr_t)y) | ((uintptr_t)w) | ((uintptr_t)b)) & 15) == 0;
    if (vec && W == 64) ln_fwd_vec_kernel<64, 4><<<grid, 256, 0, st>>>(x, w, b, eps, y, mu, rs, R);
    else if (vec && W == 32) ln_fwd_vec_kernel<32, 4><<<grid, 256, 0, st>>>(x, w, b, eps, y, mu, rs, R);
    else if (vec && W == 128) ln_fwd_vec_kernel<128, 4><<<grid, 256, 0, st>>>(x, w, b, eps, y, mu, rs, R);
    else if (W <= 32) ln_fwd_kernel<1, 4><<<grid, 256, 0, st>>>(x, w, b, eps, y, mu, rs, R, W);
    else if (W <= 64) ln_fwd_kernel<2, 4><<<grid, 256, 0, st>>>(x, w, b, eps, y, mu, rs, R, W);
    else if (W <= 128) ln_fwd_kernel<4, 2><<<grid, 256, 0, st>>>(x, w, b, eps, y, mu, rs, R, W);
    else ln_fwd_kernel<8, 1><<<grid, 256, 0, st>>>(x, w, b, eps, y, mu, rs, R, W); }
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}
static int ln_bwd(const float* dy, const float* x, const float* mu, const float* rs, const float* w, float* dx,
                  float* dw, float* db, long long R, int W, int relu_in, int sm, cudaStream_t st) {
    long long blocks = (R + 7) / 8;
    int grid = (int)(blocks < (long long)sm * 8 ? blocks : (long long)sm * 8);
    if (grid < 1) return DOF_OK;
    { ProfScope ps("ln_bwd", st, 0.0, 12.0 * R * W);
    const bool vec = ((((uintptr_t)dy) | ((uintptr_t)x) | ((uintptr_t)dx) | ((uintptr_t)w)) & 15) == 0;
    if (vec && W == 64) ln_bwd_vec_kernel<64, 4><<<grid, 256, 0, st>>>(dy, x, mu, rs, w, dx, dw, db, R, relu_in);
    else if (vec && W == 32) ln_bwd_vec_kernel<32, 4><<<grid, 256, 0, st>>>(dy, x, mu, rs, w, dx, dw, db, R, relu_in);
    else if (vec && W == 128) ln_bwd_vec_kernel<128, 2><<<grid, 256, 0, st>>>(dy, x, mu, rs, w, dx, dw, db, R, relu_in);
    else if (W <= 32) ln_bwd_kernel<1, 4><<<grid, 256, 0, st>>>(dy, x, mu, rs, w, dx, dw, db, R, W, relu_in);
    else if (W <= 64) ln_bwd_kernel<2, 4><<<grid, 256, 0, st>>>(dy, x, mu, rs, w, dx, dw, db, R, W, relu_in);
    else if (W <= 128) ln_bwd_kernel<4, 2><<<grid, 256, 0, st>>>(dy, x, mu, rs, w, dx, dw, db, R, W, relu_in);
    else ln_bwd_kernel<8, 1><<<grid, 256, 0, st>>>(dy, x, mu, rs, w, dx, dw, db, R, W, relu_in); }
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

// two-direction input projection Gi[dir] = A . W_ih[dir]^T + b_ih[dir]
static int gru_input_proj(const float* state, const GruP& g, MatView A, int M, int I, int H, float* const Gi[2],
                          cudaStream_t st) {
    GemmArgs ga[2];
    for (int d = 0; d < 2; d++)
        ga[d] = gemm_args(A, state + g.w_ih[d], I, 0, state + g.b_ih[d], Gi[d], 3 * H, M, 3 * H, I);
    return launch_gemm_rows(ga, 2, st);
}

// One bidirectional GRU layer.  Eligible shapes run the fused tcgen05 kernel (input projection + recurrence +
// gates, gru_tc.cuh); the rest fall back to the projection GEMM (into Gi) + the SIMT recurrent kernel.
// X rows: sequence s, step t at X + s*x_ss + t*x_st (x_st = 0: the same input at every step).
static int gru_layer_forward(const float* state, const GruP& g, const float* X, long long x_ss, int x_st, int I, int H, int S,
                             int T, const int* len, float* Hout, float* const Gt[2], float* Hn, float* const Gi[2],
                             cudaStream_t st) {
    if (gru_tc_eligible(S, H, I)) {
        GruTcArgs a;
        memset(&a, 0, sizeof(a));
        a.X = X; a.x_ss = x_ss; a.x_st = x_st;
        for (int d = 0; d < 2; d++) {
            a.Wih[d] = state + g.w_ih[d]; a.Whh[d] = state + g.w_hh[d]; a.bih[d] = state + g.b_ih[d]; a.bhh[d] = state + g.b_hh[d];
            a.Gt[d] = Gt ? Gt[d] : nullptr;
        }
        a.len = len; a.Hout = Hout; a.Hn = Hn; a.S = S; a.T = T; a.H = H; a.I = I;
        a.gt_tiled = (Gt && gru_bwd_tc_eligible(H, I)) ? 1 : 0;     // the fused backward kernel reads tiled gates
        return launch_gru_fwd_tc(a, st);
    }
    const int M = x_st == 0 ? S : S * T;
    DOF_TRY(gru_input_proj(state, g, mv_plain(X, I), M, I, H, Gi, st));
    GruFwdArgs f;
    memset(&f, 0, sizeof(f));
    for (int d = 0; d < 2; d++) {
        f.Gi[d] = Gi[d]; f.Whh[d] = state + g.w_hh[d]; f.bhh[d] = state + g.b_hh[d];
        f.Gt[d] = Gt ? Gt[d] : nullptr;
    }
    f.gi_ss = x_st == 0 ? 3 * H : (long long)T * 3 * H; f.gi_st = x_st == 0 ? 0 : 3 * H;
    f.len = len; f.Hout = Hout; f.Hn = Hn; f.S = S; f.T = T; f.H = H;
    return launch_gru_fwd(f, st);
}

static int enc_block_forward(dof_handle* h, int b, const float* state, const float* xin, int B, bool train,
                             cudaStream_t st) {
    const dof_config& c = h->cfg;
    BlockWS& w = h->blk[b];
    const BlockP& P = h->L.blk[b];
    const int T = c.T, S = B * w.G, H1 = h->H1, H2 = h->H2, C1 = h->C1;
    const int M = S * T;
    EncConvArgs ca;
    ca.x = xin; ca.gidx = w.gidx; ca.w = state + P.conv; ca.Xs = w.Xs; ca.Cv = w.Cv; ca.len = w.len;
    ca.B = B; ca.T = T; ca.G = w.G; ca.F = w.Fin; ca.C = C1;
    size_t smem = ((size_t)T * w.G * w.Fin + (size_t)C1 * w.Fin * 5 + (size_t)w.G * T) * 4;
    { ProfScope ps("enc_conv", st);
    enc_conv_kernel<<<B, 256, smem, st>>>(ca); }
    DOF_LAUNCH_CHECK();
    DOF_TRY(gru_layer_forward(state, P.g1, w.Cv, (long long)T * C1, C1, C1, H1, S, T, w.len, w.H1, train ? w.Gt1 : nullptr, nullptr,
                              w.Gi1, st));
    DOF_TRY(ln_fwd(w.H1, state + P.n1w, state + P.n1b, w.Y1, w.mu1, w.rs1, M, 2 * H1, h->sm_count, st));
    DOF_TRY(gru_layer_forward(state, P.g2, w.Y1, (long long)T * 2 * H1, 2 * H1, 2 * H1, H2, S, T, w.len, train ? w.H2 : nullptr,
                              train ? w.Gt2 : nullptr, w.Hn, w.Gi2, st));
    DOF_TRY(ln_fwd(w.Hn, state + P.n2w, state + P.n2b, w.Y2, w.mu2, w.rs2, S, 2 * H2, h->sm_count, st));
    if (h->di != c.D) {
        GemmArgs g = gemm_args(mv_plain(w.Y2, 2 * H2), state + P.pw, 2 * H2, 0, state + P.pb, w.P, 2 * c.D, S,
                               2 * c.D, 2 * H2);
        DOF_TRY(launch_gemm_rows(&g, 1, st));
    }
    return DOF_OK;
}

static CensArgs cens_args(dof_handle* h, const float* state, int B) {
    const dof_config& c = h->cfg;
    CensArgs a;
    memset(&a, 0, sizeof(a));
    a.node = h->blk[0].P; a.edge = h->blk[1].P;
    a.lap = state + h->L.lap; a.elap = state + h->L.elap; a.inc = state + h->L.inc;
    a.wn = state + h->L.node_weights; a.we = state + h->L.edge_weights;
    a.Pn = h->Pn; a.Pe = h->Pe;
    a.B = B; a.N = c.N; a.E = c.E; a.C = 2 * c.D;
    return a;
}

static int rec_encoder_forward(dof_handle* h, const float* state, const float* x, const float* a, int B, bool train,
                               cudaStream_t st) {
    const dof_config& c = h->cfg;
    const Layout& L = h->L;
    const int N = c.N, E = c.E, D = c.D;
    DOF_TRY(fork_join_blocks(h, st, [&](int b, cudaStream_t s) { return enc_block_forward(h, b, state, b == 0 ? x : a, B, train, s); }));
    CensArgs ca = cens_args(h, state, B);
    size_t smem = cens_smem_floats(N, E, 2 * D) * 4;
    if (smem > 200 * 1024) DOF_FAIL(DOF_ERR_UNSUPPORTED, "graph too large for the CensNet kernel (%zu B smem)", smem);
    static bool attr = false;
    if (!attr) {
        DOF_CUDA(cudaFuncSetAttribute(cens_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        DOF_CUDA(cudaFuncSetAttribute(cens_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr = true;
    }
    { ProfScope ps("cens_fwd", st);
    cens_fwd_kernel<<<B, 128, smem, st>>>(ca); }
    DOF_LAUNCH_CHECK();
    GemmArgs g[2];
    g[0] = gemm_args(mv_plain(h->Pn, 2 * D), state + L.node_kernel, D, 1, state + L.node_bias, h->On, D, B * N, D, 2 * D);
    g[1] = gemm_args(mv_plain(h->Pe, 2 * D), state + L.edge_kernel, D, 1, state + L.edge_bias, h->Oe, D, B * E, D, 2 * D);
    g[0].relu = g[1].relu = 1;
    DOF_TRY(launch_gemm_rows(g, 2, st));
    // final dense over the concatenated node / edge embeddings.  K = G*D (224 at cfg2) is split into two K blocks
    // so the problem fits the tensor-core kernel's tile (the SIMT kernel would run it on B/128 = 32 CTAs).
    for (int part = 0; part < 2; part++) {
        const float* X = part == 0 ? h->On : h->Oe;
        const int GD = (part == 0 ? N : E) * D;
        const float* W = state + L.final_w + (part == 0 ? 0 : (size_t)N * D);
        GemmArgs f;
        if (GD % 8 == 0) {
            f = gemm_args(mv_plain(X, GD), W, (N + E) * D, 0, part == 0 ? state + L.final_b : nullptr, h->enc, D, B, D, GD / 2);
            f.nkb = 2;
            f.A2 = mv_plain(X + GD / 2, GD);
            f.W2 = W + GD / 2;
        } else {
            f = gemm_args(mv_plain(X, GD), W, (N + E) * D, 0, part == 0 ? state + L.final_b : nullptr, h->enc, D, B, D, GD);
        }
        f.accum = part;
        DOF_TRY(launch_gemm_rows(&f, 1, st));
    }
    return DOF_OK;
}

// GaussianMixtureLatentPT (models_new.py:1679-1791): latent heads, reparameterisation, GMM posterior
static int vade_latent_forward(dof_handle* h, const float* state, int B, const float* eps, cudaStream_t st, bool sample = false) {
    const dof_config& c = h->cfg;
    const Layout& L = h->L;
    const int D = c.D, K = c.K;
    LatentArgs la;
    la.enc = h->enc; la.Wm = state + L.Wm; la.bm = state + L.bm; la.Wv = state + L.Wv; la.bv = state + L.bv;
    la.eps = eps; la.noise_seed = (sample && !eps) ? h->noise_seed : 0ull;
    la.gmm_mu = state + L.gmm_mu; la.gmm_lv = state + L.gmm_lv; la.prior = state + L.prior;
    la.zm = h->zm; la.pre = h->pre; la.lv = h->lv; la.z = h->z; la.q = h->q; la.B = B; la.D = D; la.K = K;
    size_t lsm = ((size_t)2 * D * D + 2 * D + 2 * (size_t)K * D + K) * 4;
    { ProfScope ps("latent_fwd", st);
    latent_fwd_kernel<<<cdiv(B, 64), 64, lsm, st>>>(la); }
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

static int rec_decoder_forward(dof_handle* h, const float* state, const float* zin, const float* x, int B, bool train,
                               cudaStream_t st) {
    const dof_config& c = h->cfg;
    const Layout& L = h->L;
    const int T = c.T, D = c.D, NF = c.N * c.F, M = B * T;
    { ProfScope ps("row_valid_len", st);
    row_valid_len_kernel<<<cdiv((long long)B * 32, 256), 256, 0, st>>>(x, h->lenD, B, T, NF); }
    DOF_LAUNCH_CHECK();
    DOF_TRY(gru_layer_forward(state, L.dg1, zin, D, 0, D, D, B, T, h->lenD, h->HD1, train ? h->GtD1 : nullptr, nullptr, h->GiD1, st));
    DOF_TRY(ln_fwd(h->HD1, state + L.dn1w, state + L.dn1b, h->YD1, h->muD1, h->rsD1, M, 2 * D, h->sm_count, st));
    DOF_TRY(gru_layer_forward(state, L.dg2, h->YD1, (long long)T * 2 * D, 2 * D, 2 * D, 2 * D, B, T, h->lenD, h->HD2,
                              train ? h->GtD2 : nullptr, nullptr, h->GiD2, st));
    DOF_TRY(ln_fwd(h->HD2, state + L.dn2w, state + L.dn2b, h->YD2, h->muD2, h->rsD2, M, 4 * D, h->sm_count, st));
    // Conv1d(k = 5, "same") + ReLU: on the tensor-core kernel as a 5-tap A_TAPS operand staged one tap (4 D columns) per K block,
    // the torch weight [2D, 4D, 5] addressed in place; the im2col view on the SIMT kernel for shapes that kernel does not take
    GemmArgs gc = gemm_args(mv_taps(h->YD2, 4 * D, T, 5, 4 * D, -1, +2), state + L.dconv, 20 * D, 0, nullptr, h->Cd, 2 * D, M, 2 * D, 20 * D);
    gc.relu = 1; gc.wconv = 1; gc.wcin = 4 * D; gc.wtaps = 5; gc.ksplit = 5;
    static const bool conv_tc = !(getenv("DOF_DEC_CONV_TC") && getenv("DOF_DEC_CONV_TC")[0] == '0');
    if (!(conv_tc && tc_enabled() && tc_rows_eligible(gc))) {
        gc = gemm_args(mv_conv5(h->YD2, 4 * D, T, +1), state + L.dconv, 20 * D, 0, nullptr, h->Cd, 2 * D, M, 2 * D, 20 * D);
        gc.relu = 1;
    }
    DOF_TRY(launch_gemm_rows(&gc, 1, st));
    DOF_TRY(ln_fwd(h->Cd, state + L.dn3w, state + L.dn3b, h->YD3, h->muD3, h->rsD3, M, 2 * D, h->sm_count, st));
    GemmArgs gl = gemm_args(mv_plain(h->YD3, 2 * D), state + L.loc_w, 2 * D, 0, state + L.loc_b, h->loc, NF, M, NF, 2 * D);
    DOF_TRY(launch_gemm_rows(&gl, 1, st));
    return DOF_OK;
}

static void register_debug(dof_handle* h, int B) {
    const dof_config& c = h->cfg;
    h->dbg.clear();
    const bool tcn = c.encoder == DOF_ENCODER_TCN;
    const bool tfm = c.encoder == DOF_ENCODER_TRANSFORMER || tcn;
    if (tcn) {
        const int C = h->L.dk;
        h->dbg["node_out"] = {h->ts[0].FIN, (int64_t)B * c.N * C};
        h->dbg["edge_out"] = {h->ts[1].FIN, (int64_t)B * c.E * C};
        h->dbg["node_a1"] = {h->ts[0].b[0].A1, (int64_t)B * c.N * c.T * C};
        h->dbg["node_out0"] = {h->ts[0].b[0].OUT, (int64_t)B * c.N * c.T * C};
        h->dbg["edge_a1"] = {h->ts[1].b[0].A1, (int64_t)B * c.E * c.T * C};
        h->dbg["edge_a2"] = {h->ts[1].b[0].A2, (int64_t)B * c.E * c.T * C};
        h->dbg["head_in"] = {h->hIn, (int64_t)B * (c.N + c.E) * c.D};
        h->dbg["tcn_stat"] = {h->tstat, (int64_t)4 * h->tstat_stride};      // doubles viewed as floats
        if (c.model != DOF_MODEL_CONTRASTIVE) {
            h->dbg["dec_fin"] = {h->ts[2].FIN, (int64_t)B * c.T * h->L.tstack[2].C};
            h->dbg["dec_x0"] = {h->ts[2].X0, (int64_t)B * c.T * 4 * c.D};
        }
    } else if (tfm) {
        const int ll = h->L.layers - 1, dk = h->L.dk;
        h->dbg["node_out"] = {h->tc[0].l[ll].Y2, (int64_t)B * c.N * dk};
        h->dbg["edge_out"] = {h->tc[1].l[ll].Y2, (int64_t)B * c.E * dk};
        h->dbg["node_y0"] = {h->tc[0].Y0, (int64_t)B * c.N * c.T * dk};
        h->dbg["node_l0"] = {h->tc[0].l[0].Y2, (int64_t)B * c.N * (ll == 0 ? 1 : c.T) * dk};
        h->dbg["head_in"] = {h->hIn, (int64_t)B * (c.N + c.E) * c.D};
        h->dbg["head_h1"] = {h->h1, (int64_t)B * 2 * c.D};
        h->dbg["head_h3"] = {h->h3, (int64_t)B * c.D};
        h->dbg["bn2_stat"] = {h->bnstat[0], (int64_t)4 * c.D};
        h->dbg["bn5_stat"] = {h->bnstat[1], (int64_t)2 * c.D};
        if (h->training) {
            h->dbg["dnode_out"] = {h->tc[0].dOut, (int64_t)B * c.N * dk};
            h->dbg["dedge_out"] = {h->tc[1].dOut, (int64_t)B * c.E * dk};
            h->dbg["dhead_in"] = {h->dhIn, (int64_t)B * (c.N + c.E) * c.D};
            h->dbg["dnode_y0"] = {h->tc[0].dA, (int64_t)B * c.N * c.T * dk};
        }
    } else {
    h->dbg["node_out"] = {h->blk[0].P, (int64_t)B * c.N * 2 * c.D};
    h->dbg["edge_out"] = {h->blk[1].P, (int64_t)B * c.E * 2 * c.D};
    h->dbg["node_conv"] = {h->blk[0].Cv, (int64_t)B * c.N * c.T * h->C1};
    h->dbg["node_gru1"] = {h->blk[0].H1, (int64_t)B * c.N * c.T * 2 * h->H1};
    h->dbg["node_hn"] = {h->blk[0].Hn, (int64_t)B * c.N * 2 * h->H2};
    h->dbg["len_node"] = {h->blk[0].len, (int64_t)B * c.N};
    h->dbg["len_edge"] = {h->blk[1].len, (int64_t)B * c.E};
    }
    h->dbg["cens_node"] = {h->On, (int64_t)B * c.N * c.D};
    h->dbg["cens_edge"] = {h->Oe, (int64_t)B * c.E * c.D};
    h->dbg["enc"] = {h->enc, (int64_t)B * c.D};
    h->dbg["quant"] = {h->quant, (int64_t)B * c.D};
    h->dbg["soft"] = {h->soft, (int64_t)B * c.K};
    h->dbg["z"] = {h->z, (int64_t)B * c.D};
    h->dbg["z_mean"] = {h->zm, (int64_t)B * c.D};
    h->dbg["z_log_var"] = {h->lv, (int64_t)B * c.D};
    h->dbg["q"] = {h->q, (int64_t)B * c.K};
    h->dbg["loc"] = {h->loc, (int64_t)B * c.T * c.N * c.F};
    if (h->training) {
        h->dbg["dloc"] = {h->dloc, (int64_t)B * c.T * c.N * c.F};
        h->dbg["dz_dec"] = {h->dz_dec, (int64_t)B * c.D};
        h->dbg["dzm"] = {h->dzm, (int64_t)B * c.D};
        h->dbg["dpre"] = {h->dpre, (int64_t)B * c.D};
        h->dbg["denc"] = {h->denc, (int64_t)B * c.D};
        if (!tfm) {
            h->dbg["dnode_out"] = {h->blk[0].dP, (int64_t)B * c.N * 2 * c.D};
            h->dbg["dedge_out"] = {h->blk[1].dP, (int64_t)B * c.E * 2 * c.D};
        }
    }
}

// ---------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------
// gradients of one bidirectional GRU given dG[dir] [M rows x 4H] (layout da_r|da_z|da_n*r|da_n):
//   dW_ih, db_ih (from X), dW_hh, db_hh (from time-shifted H), and optionally dX.
static bool g_gru_wgrad_merged = getenv("DOF_GRU_WGRAD_SPLIT") == nullptr;   // env switch for A/B measurements

static int gru_param_grads(dof_handle* h, const GruP& g, float* grad, const float* state, float* const dG[2],
                           MatView X, int Mx, float* const dGx[2],   // dGx: rows matching X (== dG unless repeated input)
                           const float* Hout, int M, int T, int I, int H, float* dX, const float* dXmask,
                           cudaStream_t st) {
    if (g_gru_wgrad_merged && Mx == M && dGx[0] == dG[0] && dGx[1] == dG[1] && X.mode == A_PLAIN &&
        gru_wgrad_tc_eligible(M, H, I, X.ld, dG[0], dG[1], X.p, Hout)) {
        // all four parameter gradients of both directions in one pass over dG (gru_wgrad_tc.cuh)
        float* dWih[2] = {grad + g.w_ih[0], grad + g.w_ih[1]};
        float* dWhh[2] = {grad + g.w_hh[0], grad + g.w_hh[1]};
        float* dbih[2] = {grad + g.b_ih[0], grad + g.b_ih[1]};
        float* dbhh[2] = {grad + g.b_hh[0], grad + g.b_hh[1]};
        DOF_TRY(launch_gru_wgrad_tc(dG, X.p, X.ld, Hout, dWih, dWhh, dbih, dbhh, M, T, I, H, h->sm_count, st));
    } else {
        WGradArgs wa[2];
        for (int d = 0; d < 2; d++)
            wa[d] = wgrad_args(mv_split(dGx[d], 4 * H, 2 * H, H), X, grad + g.w_ih[d], I, 0, grad + g.b_ih[d], Mx, 3 * H, I);
        DOF_TRY(launch_gemm_wgrad(wa, 2, st, h->sm_count));
        for (int d = 0; d < 2; d++)
            wa[d] = wgrad_args(mv_plain(dG[d], 4 * H), mv_tshift(Hout + d * H, 2 * H, T, d ? +1 : -1), grad + g.w_hh[d], H, 0,
                               grad + g.b_hh[d], M, 3 * H, H);
        DOF_TRY(launch_gemm_wgrad(wa, 2, st, h->sm_count));
    }
    if (dX) {
        // dX = dGi_fwd . W_ih_fwd + dGi_bwd . W_ih_bwd as ONE two-K-block GEMM (no read-modify-write of dX)
        GemmArgs ga = gemm_args(mv_split(dGx[0], 4 * H, 2 * H, H), state + g.w_ih[0], I, 1, nullptr, dX, I, Mx, I, 3 * H);
        ga.nkb = 2;
        ga.A2 = mv_split(dGx[1], 4 * H, 2 * H, H);
        ga.W2 = state + g.w_ih[1];
        if (dXmask) { ga.mask = dXmask; ga.ldmask = I; }
        DOF_TRY(launch_gemm_rows(&ga, 1, st));
    }
    return DOF_OK;
}

// BPTT of one bidirectional GRU layer -> dG[dir] (and, on the fused path, the input gradient dX).  Returns through
// *dx_done whether dX has been produced (otherwise gru_param_grads runs the input-gradient GEMM) and through *w_done
// whether the four parameter gradients have been accumulated as well (second-generation fused kernel, gru_bwdw_tc.cuh:
// dG never reaches HBM; taken when the layer input changes per step and `grad` is given).
static int gru_layer_backward(const float* state, const GruP& g, int S, int T, int I, int H, const int* len, const float* Hout,
                              float* const Gt[2], const float* dOut, const float* dHn, float* const dG[2], float* dX,
                              const float* dXmask, bool* dx_done, cudaStream_t st, const float* X = nullptr, long long x_ss = 0,
                              int x_st = 0, float* grad = nullptr, bool* w_done = nullptr) {
    *dx_done = false;
    if (w_done) *w_done = false;
    if (w_done && grad && X && x_st != 0 && dX && gru_bwdw_tc_eligible(H, I)) {
        GruBwdwArgs b;
        memset(&b, 0, sizeof(b));
        for (int d = 0; d < 2; d++) {
            b.Whh[d] = state + g.w_hh[d]; b.Wih[d] = state + g.w_ih[d]; b.GtT[d] = Gt[d];
            b.dWih[d] = grad + g.w_ih[d]; b.dWhh[d] = grad + g.w_hh[d]; b.dbih[d] = grad + g.b_ih[d]; b.dbhh[d] = grad + g.b_hh[d];
        }
        b.len = len; b.X = X; b.x_ss = x_ss; b.x_st = x_st; b.Hout = Hout; b.dOut = dOut; b.dHn = dHn; b.dX = dX; b.dXmask = dXmask;
        b.S = S; b.T = T; b.H = H; b.I = I;
        DOF_CUDA(cudaMemsetAsync(dX, 0, (size_t)S * T * I * 4, st));
        *dx_done = true; *w_done = true;
        return launch_gru_bwdw_tc(b, st);
    }
    if (gru_bwd_tc_eligible(H, I)) {
        GruBwdTcArgs b;
        memset(&b, 0, sizeof(b));
        for (int d = 0; d < 2; d++) { b.Whh[d] = state + g.w_hh[d]; b.Wih[d] = state + g.w_ih[d]; b.GtT[d] = Gt[d]; b.dG[d] = dG[d]; }
        b.len = len; b.Hout = Hout; b.dOut = dOut; b.dHn = dHn; b.dX = dX; b.dXmask = dXmask; b.S = S; b.T = T; b.H = H; b.I = I;
        if (dX) { DOF_CUDA(cudaMemsetAsync(dX, 0, (size_t)S * T * I * 4, st)); *dx_done = true; }
        return launch_gru_bwd_tc(b, st);
    }
    GruBwdArgs b;
    memset(&b, 0, sizeof(b));
    for (int d = 0; d < 2; d++) { b.Whh[d] = state + g.w_hh[d]; b.Gt[d] = Gt[d]; b.dG[d] = dG[d]; }
    b.len = len; b.Hout = Hout; b.dOut = dOut; b.dHn = dHn; b.S = S; b.T = T; b.H = H;
    return launch_gru_bwd(b, st);
}

// h->dloc [R, N * F] with its row pitch padded to a multiple of 4 floats (zero columns); h->dloc itself when it already is
static const float* padded_dloc(dof_handle* h, long long R, cudaStream_t st) {
    const int Dx = h->cfg.N * h->cfg.F, DxP = round_up(Dx, 4);
    if (DxP == Dx) return h->dloc;
    ProfScope ps("pad_cols", st, 0.0, 4.0 * R * (Dx + DxP));
    pad_cols_kernel<<<cdiv(R * DxP, 256), 256, 0, st>>>(h->dloc, Dx, h->dlocP, DxP, R, Dx);
    return h->dlocP;
}

static int rec_decoder_backward(dof_handle* h, const float* state, float* grad, const float* zin, int B, cudaStream_t st) {
    const dof_config& c = h->cfg;
    const Layout& L = h->L;
    const int T = c.T, D = c.D, NF = c.N * c.F, M = B * T, sm = h->sm_count;
    const int NFP = round_up(NF, 4);
    WGradArgs w0 = wgrad_args(mv_plain(padded_dloc(h, M, st), NFP), mv_plain(h->YD3, 2 * D), grad + L.loc_w, 2 * D, 0, grad + L.loc_b, M, NFP, 2 * D);
    w0.nv = NF;
    DOF_TRY(launch_gemm_wgrad(&w0, 1, st, sm));
    GemmArgs g0 = gemm_args(mv_plain(h->dloc, NF), state + L.loc_w, 2 * D, 1, nullptr, h->dYD3, 2 * D, M, 2 * D, NF);
    DOF_TRY(launch_gemm_rows(&g0, 1, st));
    DOF_TRY(ln_bwd(h->dYD3, h->Cd, h->muD3, h->rsD3, state + L.dn3w, h->dCd, grad + L.dn3w, grad + L.dn3b, M, 2 * D, 1, sm, st));
    // conv weight gradient dW[o][c][kk] = sum_rows dCd[(s,t), o] * YD2[(s, t+kk-2), c]: one time-shifted tensor-core
    // weight-gradient GEMM per tap (K = 4D, output column stride 5), the five taps batched in one launch; the im2col
    // GEMM (K = 20 D, SIMT) only for shapes the tensor-core kernel does not take
    WGradArgs w5[5];
    bool tc5 = tc_enabled() && (L.dconv & 3) == 0;
    for (int kk = 0; kk < 5; kk++) {
        w5[kk] = wgrad_args(mv_plain(h->dCd, 2 * D), mv_tshift(h->YD2, 4 * D, T, kk - 2), grad + L.dconv + kk, 20 * D, 0, nullptr, M, 2 * D, 4 * D);
        w5[kk].ks = 5;
        tc5 = tc5 && tc_wgrad_eligible(w5[kk]);
    }
    if (tc5) {
        DOF_TRY(launch_gemm_wgrad_tc(w5, 5, st, sm));
    } else {
        WGradArgs w1 = wgrad_args(mv_plain(h->dCd, 2 * D), mv_conv5(h->YD2, 4 * D, T, +1), grad + L.dconv, 20 * D, 0, nullptr, M, 2 * D, 20 * D);
        DOF_TRY(launch_gemm_wgrad(&w1, 1, st, sm));
    }
    // input gradient of the Conv1d: the same 5-tap operand over dCd with the taps mirrored (row t + 2 - j), weight addressed in place
    GemmArgs g1 = gemm_args(mv_taps(h->dCd, 2 * D, T, 5, 2 * D, +1, -2), state + L.dconv, 20 * D, 0, nullptr, h->dYD2, 4 * D, M, 4 * D, 10 * D);
    g1.wconv = 2; g1.wcin = 4 * D; g1.wtaps = 5; g1.ksplit = 5;
    static const bool conv_tc = !(getenv("DOF_DEC_CONV_TC") && getenv("DOF_DEC_CONV_TC")[0] == '0');
    if (!(conv_tc && tc_enabled() && tc_rows_eligible(g1))) {
        { ProfScope ps("conv_w_transpose", st);
        conv_w_transpose_kernel<<<cdiv(2 * D * 4 * D * 5, 256), 256, 0, st>>>(state + L.dconv, h->Wt, 2 * D, 4 * D); }
        DOF_LAUNCH_CHECK();
        g1 = gemm_args(mv_conv5(h->dCd, 2 * D, T, -1), h->Wt, 10 * D, 0, nullptr, h->dYD2, 4 * D, M, 4 * D, 10 * D);
    }
    DOF_TRY(launch_gemm_rows(&g1, 1, st));
    DOF_TRY(ln_bwd(h->dYD2, h->HD2, h->muD2, h->rsD2, state + L.dn2w, h->dHD2, grad + L.dn2w, grad + L.dn2b, M, 4 * D, 0, sm, st));
    bool dxd = false;
    bool wd = false;
    DOF_TRY(gru_layer_backward(state, L.dg2, B, T, 2 * D, 2 * D, h->lenD, h->HD2, h->GtD2, h->dHD2, nullptr, h->dGD2, h->dYD1, nullptr, &dxd, st,
                               h->YD1, (long long)T * 2 * D, 2 * D, grad, &wd));
    if (!wd)
        DOF_TRY(gru_param_grads(h, L.dg2, grad, state, h->dGD2, mv_plain(h->YD1, 2 * D), M, h->dGD2, h->HD2, M, T, 2 * D, 2 * D,
                                dxd ? nullptr : h->dYD1, nullptr, st));
    DOF_TRY(ln_bwd(h->dYD1, h->HD1, h->muD1, h->rsD1, state + L.dn1w, h->dHD1, grad + L.dn1w, grad + L.dn1b, M, 2 * D, 0, sm, st));
    // the decoder's first GRU sees the SAME input (z) at every step: its input gradient is the sum over time and
    // stays on the GEMM path below (dGs), only dG comes from the fused kernel
    DOF_TRY(gru_layer_backward(state, L.dg1, B, T, D, D, h->lenD, h->HD1, h->GtD1, h->dHD1, nullptr, h->dGD1, nullptr, nullptr, &dxd, st));
    for (int d = 0; d < 2; d++) {
        { ProfScope ps("sum_over_t", st);
        sum_over_t_kernel<<<cdiv((long long)B * 4 * D, 256), 256, 0, st>>>(h->dGD1[d], h->dGs[d], B, T, 4 * D, 0); }
        DOF_LAUNCH_CHECK();
    }
    DOF_TRY(gru_param_grads(h, L.dg1, grad, state, h->dGD1, mv_plain(zin, D), B, h->dGs, h->HD1, M, T, D, D, h->dz_dec,
                            nullptr, st));
    return DOF_OK;
}

static int enc_block_backward(dof_handle* h, int bi, const float* state, float* grad, int B, cudaStream_t st) {
    const dof_config& c = h->cfg;
    BlockWS& w = h->blk[bi];
    const BlockP& P = h->L.blk[bi];
    const int T = c.T, S = B * w.G, H1 = h->H1, H2 = h->H2, C1 = h->C1, M = S * T, sm = h->sm_count;
    if (h->di != c.D) {
        WGradArgs wp = wgrad_args(mv_plain(w.dP, 2 * c.D), mv_plain(w.Y2, 2 * H2), grad + P.pw, 2 * H2, 0, grad + P.pb, S, 2 * c.D, 2 * H2);
        DOF_TRY(launch_gemm_wgrad(&wp, 1, st, sm));
        GemmArgs gp = gemm_args(mv_plain(w.dP, 2 * c.D), state + P.pw, 2 * H2, 1, nullptr, w.dY2, 2 * H2, S, 2 * H2, 2 * c.D);
        DOF_TRY(launch_gemm_rows(&gp, 1, st));
    }
    DOF_TRY(ln_bwd(w.dY2, w.Hn, w.mu2, w.rs2, state + P.n2w, w.dHn, grad + P.n2w, grad + P.n2b, S, 2 * H2, 0, sm, st));
    bool dxd = false;
    bool wd = false;
    DOF_TRY(gru_layer_backward(state, P.g2, S, T, 2 * H1, H2, w.len, w.H2, w.Gt2, nullptr, w.dHn, w.dG2, w.dY1, nullptr, &dxd, st,
                               w.Y1, (long long)T * 2 * H1, 2 * H1, grad, &wd));
    if (!wd)
        DOF_TRY(gru_param_grads(h, P.g2, grad, state, w.dG2, mv_plain(w.Y1, 2 * H1), M, w.dG2, w.H2, M, T, 2 * H1, H2,
                                dxd ? nullptr : w.dY1, nullptr, st));
    DOF_TRY(ln_bwd(w.dY1, w.H1, w.mu1, w.rs1, state + P.n1w, w.dH1, grad + P.n1w, grad + P.n1b, M, 2 * H1, 0, sm, st));
    DOF_TRY(gru_layer_backward(state, P.g1, S, T, C1, H1, w.len, w.H1, w.Gt1, w.dH1, nullptr, w.dG1, w.dCv, w.Cv, &dxd, st,
                               w.Cv, (long long)T * C1, C1, grad, &wd));
    if (!wd)
        DOF_TRY(gru_param_grads(h, P.g1, grad, state, w.dG1, mv_plain(w.Cv, C1), M, w.dG1, w.H1, M, T, C1, H1, dxd ? nullptr : w.dCv, w.Cv, st));
    if (C1 * w.Fin <= 256) {
        ConvWgradArgs ca;
        ca.dCv = w.dCv; ca.Xs = w.Xs; ca.dW = grad + P.conv; ca.S = S; ca.T = T; ca.C = C1; ca.F = w.Fin;
        const int pairs = C1 * w.Fin, slots = 256 / pairs, threads = pairs * slots;
        int grid = cdiv(S, slots) < sm * 16 ? cdiv(S, slots) : sm * 16;
        { ProfScope ps("conv_wgrad", st, 2.0 * M * C1 * w.Fin * 5, 4.0 * M * (C1 + w.Fin));
        bool vec = false;
        if (C1 == 32) vec = launch_conv_wgrad_vec<32>(ca, sm, st);
        else if (C1 == 16) vec = launch_conv_wgrad_vec<16>(ca, sm, st);
        else if (C1 == 64) vec = launch_conv_wgrad_vec<64>(ca, sm, st);
        else if (C1 == 128) vec = launch_conv_wgrad_vec<128>(ca, sm, st);
        if (!vec) conv_wgrad_kernel<<<grid, threads, 0, st>>>(ca); }
        DOF_LAUNCH_CHECK();
    } else {
        WGradArgs wc = wgrad_args(mv_plain(w.dCv, C1), mv_conv5(w.Xs, w.Fin, T, +1), grad + P.conv, w.Fin * 5, 0, nullptr, M, C1, w.Fin * 5);
        DOF_TRY(launch_gemm_wgrad(&wc, 1, st, sm));
    }
    return DOF_OK;
}

// gradient of the latent heads: dzm, dpre -> parameter gradients and d enc
static int vade_latent_backward(dof_handle* h, const float* state, float* grad, int B, cudaStream_t st) {
    const dof_config& c = h->cfg;
    const Layout& L = h->L;
    const int D = c.D, sm = h->sm_count;
    WGradArgs wl[2];
    wl[0] = wgrad_args(mv_plain(h->dzm, D), mv_plain(h->enc, D), grad + L.Wm, D, 0, grad + L.bm, B, D, D);
    wl[1] = wgrad_args(mv_plain(h->dpre, D), mv_plain(h->enc, D), grad + L.Wv, D, 0, grad + L.bv, B, D, D);
    DOF_TRY(launch_gemm_wgrad(wl, 2, st, sm));
    GemmArgs ge = gemm_args(mv_plain(h->dzm, D), state + L.Wm, D, 1, nullptr, h->denc, D, B, D, D);
    DOF_TRY(launch_gemm_rows(&ge, 1, st));
    ge = gemm_args(mv_plain(h->dpre, D), state + L.Wv, D, 1, nullptr, h->denc, D, B, D, D);
    ge.accum = 1;
    DOF_TRY(launch_gemm_rows(&ge, 1, st));
    return DOF_OK;
}

// backward of RecurrentEncoderPT from h->denc (gradient wrt the encoder output)
static int rec_encoder_backward(dof_handle* h, const float* state, float* grad, int B, cudaStream_t st) {
    const dof_config& c = h->cfg;
    const Layout& L = h->L;
    const int N = c.N, E = c.E, D = c.D, sm = h->sm_count;
    // final dense
    WGradArgs wf = wgrad_args(mv_plain(h->denc, D), mv_plain(h->On, N * D), grad + L.final_w, (N + E) * D, 0, grad + L.final_b, B, D, N * D);
    DOF_TRY(launch_gemm_wgrad(&wf, 1, st, sm));
    wf = wgrad_args(mv_plain(h->denc, D), mv_plain(h->Oe, E * D), grad + L.final_w + (size_t)N * D, (N + E) * D, 0, nullptr, B, D, E * D);
    DOF_TRY(launch_gemm_wgrad(&wf, 1, st, sm));
    GemmArgs gf = gemm_args(mv_plain(h->denc, D), state + L.final_w, (N + E) * D, 1, nullptr, h->dOn, N * D, B, N * D, D);
    gf.mask = h->On; gf.ldmask = N * D;
    DOF_TRY(launch_gemm_rows(&gf, 1, st));
    gf = gemm_args(mv_plain(h->denc, D), state + L.final_w + (size_t)N * D, (N + E) * D, 1, nullptr, h->dOe, E * D, B, E * D, D);
    gf.mask = h->Oe; gf.ldmask = E * D;
    DOF_TRY(launch_gemm_rows(&gf, 1, st));
    // CensNet dense kernels
    WGradArgs wk[2];
    wk[0] = wgrad_args(mv_plain(h->dOn, D), mv_plain(h->Pn, 2 * D), grad + L.node_kernel, D, 1, grad + L.node_bias, B * N, D, 2 * D);
    wk[1] = wgrad_args(mv_plain(h->dOe, D), mv_plain(h->Pe, 2 * D), grad + L.edge_kernel, D, 1, grad + L.edge_bias, B * E, D, 2 * D);
    DOF_TRY(launch_gemm_wgrad(wk, 2, st, sm));
    GemmArgs gk[2];
    gk[0] = gemm_args(mv_plain(h->dOn, D), state + L.node_kernel, D, 0, nullptr, h->dPn, 2 * D, B * N, 2 * D, D);
    gk[1] = gemm_args(mv_plain(h->dOe, D), state + L.edge_kernel, D, 0, nullptr, h->dPe, 2 * D, B * E, 2 * D, D);
    DOF_TRY(launch_gemm_rows(gk, 2, st));
    CensArgs ca = cens_args(h, state, B);
    ca.dPn = h->dPn; ca.dPe = h->dPe; ca.dnode = h->blk[0].dP; ca.dedge = h->blk[1].dP;
    ca.dwn = grad + L.node_weights; ca.dwe = grad + L.edge_weights;
    size_t smem = (cens_smem_floats(N, E, 2 * D) + (size_t)(N + E) * 2 * D + (size_t)N * N + (size_t)E * E + N + E) * 4;
    if (smem > 220 * 1024) DOF_FAIL(DOF_ERR_UNSUPPORTED, "graph too large for the CensNet backward kernel");
    { ProfScope ps("cens_bwd", st);
    cens_bwd_kernel<<<B, 128, smem, st>>>(ca); }
    DOF_LAUNCH_CHECK();
    DOF_TRY(fork_join_blocks(h, st, [&](int b, cudaStream_t s) { return enc_block_backward(h, b, state, grad, B, s); }));
    return DOF_OK;
}

#include "tfm_step.cuh"
#include "tcn_step.cuh"

// ---- encoder / decoder family dispatch (dof_config.encoder) ----------------------------------------------------------
static inline bool is_tfm(const dof_handle* h) { return h->cfg.encoder == DOF_ENCODER_TRANSFORMER; }
static inline bool is_tcn(const dof_handle* h) { return h->cfg.encoder == DOF_ENCODER_TCN; }
static int encoder_forward(dof_handle* h, const float* state, const float* x, const float* a, int B, bool train, cudaStream_t st) {
    if (is_tfm(h)) return tfm_encoder_forward(h, state, x, a, B, train, train ? h->next_groups : 1, st);
    if (is_tcn(h)) return tcn_encoder_forward(h, state, x, a, B, train, train ? h->next_groups : 1, st);
    return rec_encoder_forward(h, state, x, a, B, train, st);
}
static int encoder_backward(dof_handle* h, const float* state, float* grad, int B, cudaStream_t st) {
    if (is_tfm(h)) return tfm_encoder_backward(h, state, grad, B, st);
    if (is_tcn(h)) return tcn_encoder_backward(h, state, grad, B, st);
    return rec_encoder_backward(h, state, grad, B, st);
}
static int decoder_forward(dof_handle* h, const float* state, const float* zin, const float* x, int B, bool train, cudaStream_t st) {
    if (is_tfm(h)) return tfm_decoder_forward(h, state, zin, B, train, h->dec_pass, st);
    if (is_tcn(h)) return tcn_decoder_forward(h, state, zin, B, train, h->dec_pass, st);
    return rec_decoder_forward(h, state, zin, x, B, train, st);
}
static int decoder_backward(dof_handle* h, const float* state, float* grad, const float* zin, int B, cudaStream_t st) {
    if (is_tfm(h)) return tfm_decoder_backward(h, state, grad, zin, B, h->dec_pass, st);
    if (is_tcn(h)) return tcn_decoder_backward(h, state, grad, zin, B, h->dec_pass, st);
    return rec_decoder_backward(h, state, grad, zin, B, st);
}

static int check_batch(dof_handle* h, int B) {
    if (!h) DOF_FAIL(DOF_ERR_ARG, "null handle");
    if (B < 1 || B > h->max_batch) DOF_FAIL(DOF_ERR_ARG, "batch %d outside [1, %d]", B, h->max_batch);
    return DOF_OK;
}

extern "C" {

int dof_vade_forward_eval(dof_handle* h, const float* state, const float* x, const float* a, int B, float* enc,
                          float* emb, float* q, float* loc, void* stream) {
    DOF_TRY(check_batch(h, B));
    if (!state || !x || !a) DOF_FAIL(DOF_ERR_ARG, "null input");
    cudaStream_t st = (cudaStream_t)stream;
    const dof_config& c = h->cfg;
    if (c.model != DOF_MODEL_VADE) DOF_FAIL(DOF_ERR_ARG, "handle is not a VaDE model");
    DOF_TRY(encoder_forward(h, state, x, a, B, false, st));
    DOF_TRY(vade_latent_forward(h, state, B, nullptr, st));
    if (loc) DOF_TRY(decoder_forward(h, state, h->z, x, B, false, st));
    if (enc) DOF_CUDA(cudaMemcpyAsync(enc, h->enc, (size_t)B * c.D * 4, cudaMemcpyDeviceToDevice, st));
    if (emb) DOF_CUDA(cudaMemcpyAsync(emb, h->zm, (size_t)B * c.D * 4, cudaMemcpyDeviceToDevice, st));
    if (q) DOF_CUDA(cudaMemcpyAsync(q, h->q, (size_t)B * c.K * 4, cudaMemcpyDeviceToDevice, st));
    if (loc) DOF_CUDA(cudaMemcpyAsync(loc, h->loc, (size_t)B * c.T * c.N * c.F * 4, cudaMemcpyDeviceToDevice, st));
    h->lastB = B;
    register_debug(h, B);
    return DOF_OK;
}

int dof_vade_embed(dof_handle* h, const float* state, const float* x, const float* a, int B, float* emb, float* q,
                   void* stream) {
    return dof_vade_forward_eval(h, state, x, a, B, nullptr, emb, q, nullptr, stream);
}

static int vade_step(dof_handle* h, const float* state, float* grad, const float* x, const float* a, int B,
                     const float* eps, const float* mc_eps, const float* tau_batch, const float* class_weight,
                     const float* floor_c, const dof_vade_loss_cfg* loss, float* logs, void* stream, bool train) {
    DOF_TRY(check_batch(h, B));
    if (!h->training) DOF_FAIL(DOF_ERR_ARG, "handle was created with training=0");
    if (!state || (train && !grad) || !x || !a || !loss || !logs || !floor_c) DOF_FAIL(DOF_ERR_ARG, "null argument");
    if (!loss->pretrain_mode && loss->mc_samples != 32) DOF_FAIL(DOF_ERR_UNSUPPORTED, "mc_samples must be 32, got %d", loss->mc_samples);
    if (loss->tf_cluster_weight != 0.f) DOF_FAIL(DOF_ERR_UNSUPPORTED, "tf_cluster_weight != 0 is not supported");
    if (loss->reg_scatter_weight != 0.f) DOF_FAIL(DOF_ERR_UNSUPPORTED, "reg_scatter_weight != 0 is not supported");
    cudaStream_t st = (cudaStream_t)stream;
    const dof_config& c = h->cfg;
    const Layout& L = h->L;
    const int D = c.D, K = c.K, T = c.T, NF = c.N * c.F;
    if (train) DOF_CUDA(cudaMemsetAsync(grad, 0, (size_t)L.total * 4, st));
    if (c.model != DOF_MODEL_VADE) DOF_FAIL(DOF_ERR_ARG, "handle is not a VaDE model");
    DOF_TRY(encoder_forward(h, state, x, a, B, train, st));
    DOF_TRY(vade_latent_forward(h, state, B, train ? eps : nullptr, st, train));
    DOF_TRY(decoder_forward(h, state, h->z, x, B, train, st));
    // ---- loss
    StatsLayout SL = stats_layout(D, K);
    DOF_CUDA(cudaMemsetAsync(h->stats, 0, (size_t)SL.total * sizeof(double), st));
    const long long nrec = (long long)B * T * NF;
    int rgrid = (int)((nrec + 255) / 256 < (long long)h->sm_count * 8 ? (nrec + 255) / 256 : (long long)h->sm_count * 8);
    { ProfScope ps("recon", st, 0.0, 12.0 * nrec);
    recon_kernel<<<rgrid, 256, 0, st>>>(h->loc, x, train ? h->dloc : nullptr, nrec, 1.0f / ((float)B * T), h->stats + ST_RECON); }
    DOF_LAUNCH_CHECK();
    LossArgs la;
    memset(&la, 0, sizeof(la));
    la.cfg = *loss;
    la.z = h->z; la.zm = h->zm; la.lv = h->lv; la.pre = h->pre; la.q = h->q; la.eps = eps; la.mc_eps = mc_eps;
    la.noise_seed = h->noise_seed;
    la.gmm_mu = state + L.gmm_mu; la.gmm_lv = state + L.gmm_lv; la.prior = state + L.prior;
    la.tau = (loss->lambda_distill > 0.f) ? tau_batch : nullptr;
    la.class_weight = class_weight; la.floor_c = floor_c;
    la.stats = h->stats; la.coef = h->coef; la.dzm_kl = h->dzm_kl; la.dlv_kl = h->dlv_kl; la.distw = h->distw;
    la.dz_dec = h->dz_dec; la.dzm = h->dzm; la.dpre = h->dpre; la.g_mu = train ? grad + L.gmm_mu : nullptr; la.g_lv = train ? grad + L.gmm_lv : nullptr;
    la.logs = logs; la.B = B; la.D = D; la.K = K; la.T = T; la.Dx = NF;
    static bool attr = false;
    if (!attr) {
        DOF_CUDA(cudaFuncSetAttribute(loss_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        DOF_CUDA(cudaFuncSetAttribute(loss_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        DOF_CUDA(cudaFuncSetAttribute(loss_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr = true;
    }
    int lgrid = cdiv(B, LS_WARPS) < h->sm_count * 4 ? cdiv(B, LS_WARPS) : h->sm_count * 4;
    { ProfScope ps("loss_stats", st);
    loss_stats_kernel<<<lgrid, LS_WARPS * 32, loss_stats_smem_floats(D, K) * 4, st>>>(la); }
    DOF_LAUNCH_CHECK();
    { ProfScope ps("loss_finalize", st);
    loss_finalize_kernel<<<1, 256, loss_finalize_smem_bytes(D, K), st>>>(la); }
    DOF_LAUNCH_CHECK();
    if (!train) { h->lastB = B; register_debug(h, B); return DOF_OK; }
    // ---- backward
    DOF_TRY(decoder_backward(h, state, grad, h->z, B, st));
    { ProfScope ps("loss_grad", st);
    loss_grad_kernel<<<lgrid, LS_WARPS * 32, loss_grad_smem_floats(D, K) * 4, st>>>(la); }
    DOF_LAUNCH_CHECK();
    DOF_TRY(vade_latent_backward(h, state, grad, B, st));
    DOF_TRY(encoder_backward(h, state, grad, B, st));
    h->lastB = B;
    register_debug(h, B);
    return DOF_OK;
}

int dof_vade_loss_grad(dof_handle* h, const float* state, float* grad, const float* x, const float* a, int B,
                       const float* eps, const float* mc_eps, const float* tau_batch, const float* class_weight,
                       const float* floor_c, const dof_vade_loss_cfg* loss, float* logs, void* stream) {
    return vade_step(h, state, grad, x, a, B, eps, mc_eps, tau_batch, class_weight, floor_c, loss, logs, stream, true);
}

// validate_one_epoch_indexed's step (training.py:190-229): model.eval() -> z = z_mean, no dropout, BatchNorm running
// statistics; the criterion's terms and the 13 logs, no gradient
int dof_vade_loss_eval(dof_handle* h, const float* state, const float* x, const float* a, int B, const float* mc_eps,
                       const float* tau_batch, const float* class_weight, const float* floor_c, const dof_vade_loss_cfg* loss,
                       float* logs, void* stream) {
    return vade_step(h, state, nullptr, x, a, B, nullptr, mc_eps, tau_batch, class_weight, floor_c, loss, logs, stream, false);
}

// ---- encoder only: model.encoder(x, a)  (RecurrentEncoderPT.forward, models_new.py:140-181) -------------
int dof_encode(dof_handle* h, const float* state, const float* x, const float* a, int B, float* enc, void* stream) {
    DOF_TRY(check_batch(h, B));
    if (!state || !x || !a || !enc) DOF_FAIL(DOF_ERR_ARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    DOF_TRY(encoder_forward(h, state, x, a, B, false, st));
    DOF_CUDA(cudaMemcpyAsync(enc, h->enc, (size_t)B * h->cfg.D * 4, cudaMemcpyDeviceToDevice, st));
    h->lastB = B;
    register_debug(h, B);
    return DOF_OK;
}

// ---- VQ-VAE --------------------------------------------------------------------------------------------
static int vq_forward(dof_handle* h, const float* state, int B, bool want_gram, cudaStream_t st) {
    const dof_config& c = h->cfg;
    VqArgs v;
    v.z = h->enc; v.codebook = state + h->L.codebook; v.quant = h->quant; v.soft = h->soft; v.idx = h->vidx;
    v.stats = h->vstats; v.B = B; v.D = c.D; v.K = c.K; v.want_gram = want_gram ? 1 : 0;
    if (c.D > VQ_MAXD || c.K > VQ_MAXK) DOF_FAIL(DOF_ERR_UNSUPPORTED, "latent_dim %d > %d or codebook %d > %d", c.D, VQ_MAXD, c.K, VQ_MAXK);
    DOF_CUDA(cudaMemsetAsync(h->vstats, 0, ((size_t)VQ_ST_GRAM + (size_t)c.D * c.D + c.K) * sizeof(double), st));
    const size_t smem = vq_fwd_smem(c.D, c.K, want_gram);
    if (smem > 200 * 1024) DOF_FAIL(DOF_ERR_UNSUPPORTED, "codebook %d x %d does not fit in shared memory", c.D, c.K);
    static size_t attr = 48 * 1024;
    if (smem > attr) { DOF_CUDA(cudaFuncSetAttribute(vq_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = smem; }
    const int grid = cdiv(B, VQ_WARPS) < 2 * h->sm_count ? cdiv(B, VQ_WARPS) : 2 * h->sm_count;
    { ProfScope ps("vq_fwd", st, 0.0, (double)B * (8.0 * c.D + 4.0 + 4.0 * c.K));
    vq_fwd_kernel<<<grid, VQ_WARPS * 32, smem, st>>>(v); }
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

int dof_vqvae_forward_eval(dof_handle* h, const float* state, const float* x, const float* a, int B, float* enc, float* quant,
                           float* soft, int* idx, float* loc_q, float* loc_e, void* stream) {
    DOF_TRY(check_batch(h, B));
    const dof_config& c = h->cfg;
    if (c.model != DOF_MODEL_VQVAE) DOF_FAIL(DOF_ERR_ARG, "handle is not a VQ-VAE model");
    if (!state || !x || !a) DOF_FAIL(DOF_ERR_ARG, "null input");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nl = (size_t)B * c.T * c.N * c.F * 4;
    DOF_TRY(encoder_forward(h, state, x, a, B, false, st));
    DOF_TRY(vq_forward(h, state, B, false, st));
    if (enc) DOF_CUDA(cudaMemcpyAsync(enc, h->enc, (size_t)B * c.D * 4, cudaMemcpyDeviceToDevice, st));
    if (quant) DOF_CUDA(cudaMemcpyAsync(quant, h->quant, (size_t)B * c.D * 4, cudaMemcpyDeviceToDevice, st));
    if (soft) DOF_CUDA(cudaMemcpyAsync(soft, h->soft, (size_t)B * c.K * 4, cudaMemcpyDeviceToDevice, st));
    if (idx) DOF_CUDA(cudaMemcpyAsync(idx, h->vidx, (size_t)B * 4, cudaMemcpyDeviceToDevice, st));
    if (loc_q) {
        DOF_TRY(decoder_forward(h, state, h->quant, x, B, false, st));
        DOF_CUDA(cudaMemcpyAsync(loc_q, h->loc, nl, cudaMemcpyDeviceToDevice, st));
    }
    if (loc_e) {
        DOF_TRY(decoder_forward(h, state, h->enc, x, B, false, st));
        DOF_CUDA(cudaMemcpyAsync(loc_e, h->loc, nl, cudaMemcpyDeviceToDevice, st));
    }
    h->lastB = B;
    register_debug(h, B);
    return DOF_OK;
}

// step_vqvae_distill forward + backward without teacher (training.py:312-389): loss = NLL(decoder(quantized)) +
// NLL(decoder(encoder output)) + float(vq_loss) + float(kmeans_loss).  The encoder learns through the bypass
// decoder only; the codebook through the one-hot matmul of the quantized path (there is no straight-through).
// the distillation head on the first B rows of the encoder output: adds its gradient to denc, its loss sum to *stat
static int distill_head_step(dof_handle* h, const dof_distill_cfg* dc, int B, double* stat, bool normalised, cudaStream_t st) {
    const int D = h->cfg.D, K = dc->K;
    if (!dc->head || !dc->head_grad || !dc->tau_batch) DOF_FAIL(DOF_ERR_ARG, "distillation: null head / head_grad / tau_batch");
    if (K < 1 || K > DH_MAXK || D > DH_MAXD) DOF_FAIL(DOF_ERR_UNSUPPORTED, "distillation head: K=%d (<= %d), D=%d (<= %d)", K, DH_MAXK, D, DH_MAXD);
    DOF_CUDA(cudaMemsetAsync(dc->head_grad, 0, ((size_t)K * D + K) * 4, st));
    DistillArgs a;
    a.z = normalised ? h->zn : h->enc; a.nrm = normalised ? h->nrm : nullptr; a.head = dc->head; a.head_grad = dc->head_grad; a.tau = dc->tau_batch; a.denc = h->denc; a.stat = stat;
    a.B = B; a.D = D; a.K = K; a.lambda = dc->lambda; a.sharpen_T = dc->sharpen_T; a.conf_thresh = dc->conf_thresh;
    a.conf_weight = dc->conf_weight;
    { ProfScope ps("distill_head", st);
    distill_head_kernel<<<cdiv(B, 128), 128, (size_t)2 * (K * D + K) * 4, st>>>(a); }
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

int dof_vqvae_loss_grad(dof_handle* h, const float* state, float* grad, const float* x, const float* a, int B, float beta,
                        float kmeans_weight, float* logs, void* stream) {
    return dof_vqvae_loss_grad_distill(h, state, grad, x, a, B, beta, kmeans_weight, nullptr, logs, stream);
}

static int vqvae_step(dof_handle* h, const float* state, float* grad, const float* x, const float* a, int B, float beta,
                      float kmeans_weight, const dof_distill_cfg* distill, float* logs, void* stream, bool train) {
    DOF_TRY(check_batch(h, B));
    const dof_config& c = h->cfg;
    if (c.model != DOF_MODEL_VQVAE) DOF_FAIL(DOF_ERR_ARG, "handle is not a VQ-VAE model");
    if (!h->training) DOF_FAIL(DOF_ERR_ARG, "handle was created with training=0");
    if (!state || (train && !grad) || !x || !a || !logs) DOF_FAIL(DOF_ERR_ARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const Layout& L = h->L;
    const int D = c.D, K = c.K, T = c.T, NF = c.N * c.F;
    if (train) DOF_CUDA(cudaMemsetAsync(grad, 0, (size_t)L.total * 4, st));
    DOF_TRY(encoder_forward(h, state, x, a, B, train, st));
    DOF_TRY(vq_forward(h, state, B, kmeans_weight != 0.f, st));
    const long long nrec = (long long)B * T * NF;
    int rgrid = (int)((nrec + 255) / 256 < (long long)h->sm_count * 8 ? (nrec + 255) / 256 : (long long)h->sm_count * 8);
    for (int pass = 0; pass < 2; pass++) {
        const float* zin = pass == 0 ? h->quant : h->enc;
        h->dec_pass = pass;
        DOF_TRY(decoder_forward(h, state, zin, x, B, train, st));
        { ProfScope ps("recon", st, 0.0, 12.0 * nrec);
        recon_kernel<<<rgrid, 256, 0, st>>>(h->loc, x, train ? h->dloc : nullptr, nrec, 1.0f / ((float)B * T), h->vstats + (pass == 0 ? VQ_ST_REC_Q : VQ_ST_REC_E)); }
        DOF_LAUNCH_CHECK();
        if (!train) continue;
        DOF_TRY(decoder_backward(h, state, grad, zin, B, st));
        if (pass == 0) {
            { ProfScope ps("vq_codebook_grad", st);
            vq_codebook_grad_kernel<<<cdiv((long long)B * D, 256), 256, 0, st>>>(h->dz_dec, h->vidx, grad + L.codebook, B, D, K); }
            DOF_LAUNCH_CHECK();
        }
    }
    h->dec_pass = 0;
    const bool dist_on = train && distill && distill->lambda > 0.f;      // training.py:346
    if (train) {
        DOF_CUDA(cudaMemcpyAsync(h->denc, h->dz_dec, (size_t)B * D * 4, cudaMemcpyDeviceToDevice, st));
        if (dist_on) DOF_TRY(distill_head_step(h, distill, B, h->vstats + VQ_ST_DISTILL, false, st));
        DOF_TRY(encoder_backward(h, state, grad, B, st));
    }
    VqFinalArgs f;
    f.stats = h->vstats; f.logs = logs; f.B = B; f.T = T; f.Dx = NF; f.D = D; f.K = K; f.beta = beta; f.kmeans_w = kmeans_weight;
    f.lambda_distill = dist_on ? distill->lambda : 0.f;
    { ProfScope ps("vq_finalize", st);
    vq_finalize_kernel<<<1, 32, (size_t)2 * D * D * sizeof(double), st>>>(f); }
    DOF_LAUNCH_CHECK();
    h->lastB = B;
    register_debug(h, B);
    return DOF_OK;
}

int dof_vqvae_loss_grad_distill(dof_handle* h, const float* state, float* grad, const float* x, const float* a, int B, float beta,
                                float kmeans_weight, const dof_distill_cfg* distill, float* logs, void* stream) {
    return vqvae_step(h, state, grad, x, a, B, beta, kmeans_weight, distill, logs, stream, true);
}

// the validation step of fit_VQVAE (training.py:1165-1170): eval-mode forward, the loss terms, no gradient, teacher off
int dof_vqvae_loss_eval(dof_handle* h, const float* state, const float* x, const float* a, int B, float beta, float kmeans_weight,
                        float* logs, void* stream) {
    return vqvae_step(h, state, nullptr, x, a, B, beta, kmeans_weight, nullptr, logs, stream, false);
}

// ---- contrastive -------------------------------------------------------------------------------------------
int dof_contrastive_views(const dof_views_cfg* v, const float* x_full, int B, float* x2, float* a2, void* stream) {
    if (!v || !x_full || !x2 || !a2 || !v->start || B < 1) DOF_FAIL(DOF_ERR_ARG, "null / bad argument");
    if (v->T_full < 2 || v->N < 1 || v->N > VW_MAXN || v->E < 0 || v->E > LD_MAXE || v->n_rot < 0 || v->n_rot > VW_MAXROT)
        DOF_FAIL(DOF_ERR_UNSUPPORTED, "views: T_full=%d N=%d (<= %d) E=%d (<= %d) n_rot=%d (<= %d)", v->T_full, v->N, VW_MAXN, v->E,
                 LD_MAXE, v->n_rot, VW_MAXROT);
    if (v->E > 0 && !v->edges) DOF_FAIL(DOF_ERR_ARG, "null edges");
    if (v->n_rot > 0 && !v->rot_theta) DOF_FAIL(DOF_ERR_ARG, "null rot_theta");
    if ((v->interp_t0 == nullptr) != (v->interp_len == nullptr)) DOF_FAIL(DOF_ERR_ARG, "interp_t0 / interp_len must both be given");
    DOF_TRY(loader_device_check());
    ViewsArgs a;
    memset(&a, 0, sizeof(a));
    a.x_full = x_full; a.x2 = x2; a.a2 = a2; a.start = v->start; a.rot_theta = v->rot_theta; a.interp_t0 = v->interp_t0;
    a.interp_len = v->interp_len; a.noise = v->noise; a.B = B; a.Tf = v->T_full; a.Th = v->T_full / 2; a.N = v->N; a.E = v->E;
    a.n_rot = v->n_rot;
    a.mid_start = a.Th / 2;                                            // training.py:518-519
    for (int k = 0; k < v->n_rot; k++) {
        if (v->rot_pivot[k] < 0 || v->rot_pivot[k] >= v->N) DOF_FAIL(DOF_ERR_ARG, "rotation pivot out of range");
        a.rot_pivot[k] = v->rot_pivot[k]; a.rot_mask[k] = v->rot_mask[k];
    }
    for (int e = 0; e < v->E; e++) {
        if (v->edges[2 * e] < 0 || v->edges[2 * e] >= v->N || v->edges[2 * e + 1] < 0 || v->edges[2 * e + 1] >= v->N)
            DOF_FAIL(DOF_ERR_ARG, "edge %d out of range", e);
        a.e0[e] = (short)v->edges[2 * e]; a.e1[e] = (short)v->edges[2 * e + 1];
    }
    const long long total = 2LL * B * a.Th;
    int grid = cdiv(total, 128) < g_sm_count * 16 ? cdiv(total, 128) : g_sm_count * 16;
    { ProfScope ps("views", (cudaStream_t)stream, 0.0, (double)B * (1.5 * a.Tf * a.N * 3 * 4 + 2.0 * a.Th * (3 * a.N + a.E) * 4));
    views_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(a); }
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

}  // extern "C"

// ---- transformer encoder, eval forward (tfm.cuh) ---------------------------------------------------------------
struct TfmLayout {
    std::vector<Entry> e;
    int64_t total = 0;
    int64_t lap, elap, inc, core[2], node_kernel, edge_kernel, node_weights, edge_weights, node_bias, edge_bias;
    int64_t w0, b0, bn2, w3, b3, bn5, w6, b6;
};

static int64_t tfm_add(TfmLayout& L, const std::string& name, int d0, int d1 = -1) {
    Entry en;
    en.name = name; en.off = L.total; en.group = 1;
    en.shape[0] = d0; en.shape[1] = d1 < 0 ? 0 : d1; en.shape[2] = en.shape[3] = 0;
    en.ndim = d1 < 0 ? 1 : 2;
    en.numel = d1 < 0 ? d0 : (int64_t)d0 * d1;
    L.total += en.numel;
    L.e.push_back(en);
    return en.off;
}

static int check_tfm_cfg(const dof_tfm_cfg* c) {
    if (!c) DOF_FAIL(DOF_ERR_ARG, "null config");
    if (c->T < 1 || c->N < 1 || c->E < 1 || c->F < 1 || c->Fe < 1 || c->D < 1 || c->key_dim < 1 || c->heads < 1 || c->dff < 1 || c->layers < 1)
        DOF_FAIL(DOF_ERR_ARG, "bad transformer geometry");
    if (c->key_dim % c->heads) DOF_FAIL(DOF_ERR_ARG, "key_dim %d is not a multiple of heads %d", c->key_dim, c->heads);
    if (c->T > TFM_MAXT) DOF_FAIL(DOF_ERR_UNSUPPORTED, "window length %d > %d", c->T, TFM_MAXT);
    if (tfm_core_smem_bytes(c->T, c->F > c->Fe ? c->F : c->Fe, c->key_dim, c->dff, 1) > 227 * 1024)
        DOF_FAIL(DOF_ERR_UNSUPPORTED, "one transformer layer (key_dim %d, dff %d, T %d) does not fit shared memory", c->key_dim, c->dff, c->T);
    return DOF_OK;
}

// reference state_dict order of TFMEncoderPT after its first forward (the CensNet parameters are built lazily);
// the integer num_batches_tracked buffers are not part of the flat float state
static TfmLayout build_tfm_layout(const dof_tfm_cfg& c) {
    TfmLayout L;
    const int N = c.N, E = c.E, D = c.D, dk = c.key_dim;
    L.lap = tfm_add(L, "laplacian", N, N);
    L.elap = tfm_add(L, "edge_laplacian", E, E);
    L.inc = tfm_add(L, "incidence", N, E);
    const char* cn[2] = {"node_tf.", "edge_tf."};
    for (int b = 0; b < 2; b++) {
        std::string p = cn[b];
        L.core[b] = tfm_add(L, p + "embed.weight", dk, b == 0 ? c.F : c.Fe);
        tfm_add(L, p + "embed.bias", dk);
        for (int l = 0; l < c.layers; l++) {
            std::string q = p + "layers." + std::to_string(l) + ".";
            tfm_add(L, q + "mha.q_proj.weight", dk, dk); tfm_add(L, q + "mha.k_proj.weight", dk, dk);
            tfm_add(L, q + "mha.v_proj.weight", dk, dk); tfm_add(L, q + "mha.out_proj.weight", dk, dk);
            tfm_add(L, q + "norm1.weight", dk); tfm_add(L, q + "norm1.bias", dk);
            tfm_add(L, q + "ffn.0.weight", c.dff, dk); tfm_add(L, q + "ffn.0.bias", c.dff);
            tfm_add(L, q + "ffn.2.weight", dk, c.dff); tfm_add(L, q + "ffn.2.bias", dk);
            tfm_add(L, q + "norm2.weight", dk); tfm_add(L, q + "norm2.bias", dk);
        }
    }
    std::string g = "spatial_gnn_block.";
    L.node_kernel = tfm_add(L, g + "node_kernel", dk, D);
    L.edge_kernel = tfm_add(L, g + "edge_kernel", dk, D);
    L.node_weights = tfm_add(L, g + "node_weights", dk, 1);
    L.edge_weights = tfm_add(L, g + "edge_weights", dk, 1);
    L.node_bias = tfm_add(L, g + "node_bias", D);
    L.edge_bias = tfm_add(L, g + "edge_bias", D);
    L.w0 = tfm_add(L, "head.0.weight", 2 * D, (N + E) * D);
    L.b0 = tfm_add(L, "head.0.bias", 2 * D);
    L.bn2 = tfm_add(L, "head.2.weight", 2 * D); tfm_add(L, "head.2.bias", 2 * D);
    tfm_add(L, "head.2.running_mean", 2 * D); tfm_add(L, "head.2.running_var", 2 * D);
    L.w3 = tfm_add(L, "head.3.weight", D, 2 * D);
    L.b3 = tfm_add(L, "head.3.bias", D);
    L.bn5 = tfm_add(L, "head.5.weight", D); tfm_add(L, "head.5.bias", D);
    tfm_add(L, "head.5.running_mean", D); tfm_add(L, "head.5.running_var", D);
    L.w6 = tfm_add(L, "head.6.weight", D, D);
    L.b6 = tfm_add(L, "head.6.bias", D);
    return L;
}

extern "C" {

int64_t dof_tfm_numel(const dof_tfm_cfg* cfg) {
    if (check_tfm_cfg(cfg) != DOF_OK) return -1;
    return build_tfm_layout(*cfg).total;
}

int dof_tfm_num_entries(const dof_tfm_cfg* cfg) {
    if (check_tfm_cfg(cfg) != DOF_OK) return -1;
    return (int)build_tfm_layout(*cfg).e.size();
}

int dof_tfm_entry(const dof_tfm_cfg* cfg, int index, char* name_out, int64_t* offset_out, int64_t* numel_out, int* ndim_out,
                  int* shape_out) {
    DOF_TRY(check_tfm_cfg(cfg));
    TfmLayout L = build_tfm_layout(*cfg);
    if (index < 0 || index >= (int)L.e.size()) DOF_FAIL(DOF_ERR_ARG, "entry index %d out of range", index);
    const Entry& e = L.e[index];
    if (name_out) { strncpy(name_out, e.name.c_str(), 127); name_out[127] = 0; }
    if (offset_out) *offset_out = e.off;
    if (numel_out) *numel_out = e.numel;
    if (ndim_out) *ndim_out = e.ndim;
    if (shape_out) for (int i = 0; i < 4; i++) shape_out[i] = e.shape[i];
    return DOF_OK;
}

static size_t tfm_plan(const dof_tfm_cfg& c, int B, char* base, float** nodes, float** edges, float** Pn, float** Pe, float** On,
                       float** Oe, float** ybuf, bool* one_launch) {
    Bump bp{base, 0, 0, base == nullptr};
    const int dk = c.key_dim, G = c.N > c.E ? c.N : c.E;
    float* p;
    p = bp.get<float>((size_t)B * c.N * dk); if (nodes) *nodes = p;
    p = bp.get<float>((size_t)B * c.E * dk); if (edges) *edges = p;
    p = bp.get<float>((size_t)B * c.N * dk); if (Pn) *Pn = p;
    p = bp.get<float>((size_t)B * c.E * dk); if (Pe) *Pe = p;
    p = bp.get<float>((size_t)B * c.N * c.D); if (On) *On = p;
    p = bp.get<float>((size_t)B * c.E * c.D); if (Oe) *Oe = p;
    const bool one = tfm_core_smem_bytes(c.T, c.F > c.Fe ? c.F : c.Fe, dk, c.dff, c.layers) <= 227 * 1024;
    if (one_launch) *one_launch = one;
    p = one ? nullptr : bp.get<float>((size_t)B * G * c.T * dk);
    if (ybuf) *ybuf = p;
    return bp.off;
}

size_t dof_tfm_workspace_bytes(const dof_tfm_cfg* cfg, int B) {
    if (check_tfm_cfg(cfg) != DOF_OK || B < 1) return 0;
    return tfm_plan(*cfg, B, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
}

// ---- read-outs on an encoder output (the transformer family's embedding path) ----------------------------------
// GaussianMixtureLatentPT in eval mode (models_new.py:1745-1791): emb = z_mean [B,D], q = GMM posterior [B,K].
// scratch: 3*B*D floats (softplus pre-activation, log-variance, z — outputs of the shared kernel that eval does not need)
int dof_latent_eval(const float* enc, const float* Wm, const float* bm, const float* Wv, const float* bv, const float* gmm_mu,
                    const float* gmm_lv, const float* prior, int B, int D, int K, float* emb, float* q, float* scratch, void* stream) {
    if (!enc || !Wm || !bm || !Wv || !bv || !gmm_mu || !gmm_lv || !prior || !emb || !q || !scratch || B < 1)
        DOF_FAIL(DOF_ERR_ARG, "null / bad argument");
    if (K > LOSS_MAXK || D > 64) DOF_FAIL(DOF_ERR_UNSUPPORTED, "latent read-out: K=%d (<= %d), D=%d (<= 64)", K, LOSS_MAXK, D);
    LatentArgs la;
    la.enc = enc; la.Wm = Wm; la.bm = bm; la.Wv = Wv; la.bv = bv; la.eps = nullptr; la.noise_seed = 0ull; la.gmm_mu = gmm_mu; la.gmm_lv = gmm_lv; la.prior = prior;
    la.zm = emb; la.pre = scratch; la.lv = scratch + (size_t)B * D; la.z = scratch + (size_t)2 * B * D; la.q = q; la.B = B; la.D = D; la.K = K;
    const size_t lsm = ((size_t)2 * D * D + 2 * D + 2 * (size_t)K * D + K) * 4;
    if (lsm > 48 * 1024) DOF_FAIL(DOF_ERR_UNSUPPORTED, "latent head does not fit shared memory");
    { ProfScope ps("latent_fwd", (cudaStream_t)stream);
    latent_fwd_kernel<<<cdiv(B, 64), 64, lsm, (cudaStream_t)stream>>>(la); }
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

// VectorQuantizerPT read-outs (models_new.py:1358-1423): nearest code, soft counts, quantized latents.
// scratch: (4 + D*D + K) doubles
int dof_vq_eval(const float* enc, const float* codebook, int B, int D, int K, float* quant, float* soft, int* idx, double* scratch,
                void* stream) {
    if (!enc || !codebook || !quant || !soft || !idx || !scratch || B < 1) DOF_FAIL(DOF_ERR_ARG, "null / bad argument");
    if (D > VQ_MAXD || K > VQ_MAXK) DOF_FAIL(DOF_ERR_UNSUPPORTED, "latent_dim %d > %d or codebook %d > %d", D, VQ_MAXD, K, VQ_MAXK);
    cudaStream_t st = (cudaStream_t)stream;
    VqArgs v;
    v.z = enc; v.codebook = codebook; v.quant = quant; v.soft = soft; v.idx = idx; v.stats = scratch; v.B = B; v.D = D; v.K = K; v.want_gram = 0;
    DOF_CUDA(cudaMemsetAsync(scratch, 0, ((size_t)VQ_ST_GRAM + (size_t)D * D + K) * sizeof(double), st));
    const size_t smem = vq_fwd_smem(D, K, false);
    if (smem > 96 * 1024) DOF_FAIL(DOF_ERR_UNSUPPORTED, "codebook %d x %d does not fit in shared memory", D, K);
    if (smem > 48 * 1024) DOF_CUDA(cudaFuncSetAttribute(vq_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    const int grid = cdiv(B, VQ_WARPS) < 2 * g_sm_count ? cdiv(B, VQ_WARPS) : 2 * g_sm_count;
    { ProfScope ps("vq_fwd", st, 0.0, (double)B * (8.0 * D + 4.0 + 4.0 * K));
    vq_fwd_kernel<<<grid, VQ_WARPS * 32, smem, st>>>(v); }
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

// TFMEncoderPT.forward in eval mode (models_new.py:1093-1164): x [B,T,N,F], a [B,T,E,Fe] -> enc_out [B,D]; nodes_out
// [B*N,key_dim] / edges_out [B*E,key_dim] (may be NULL) receive the last-step outputs of the two transformer cores
int dof_tfm_encode(const dof_tfm_cfg* cfg, const float* state, const float* x, const float* a, int B, void* workspace,
                   size_t workspace_bytes, float* enc_out, float* nodes_out, float* edges_out, void* stream) {
    DOF_TRY(check_tfm_cfg(cfg));
    if (!state || !x || !a || !workspace || !enc_out || B < 1) DOF_FAIL(DOF_ERR_ARG, "null / bad argument");
    const dof_tfm_cfg& c = *cfg;
    float *nodes, *edges, *Pn, *Pe, *On, *Oe, *ybuf;
    bool one;
    const size_t need = tfm_plan(c, B, (char*)workspace, &nodes, &edges, &Pn, &Pe, &On, &Oe, &ybuf, &one);
    if (need > workspace_bytes) DOF_FAIL(DOF_ERR_ARG, "workspace too small: %zu < %zu bytes", workspace_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    const TfmLayout L = build_tfm_layout(c);
    const int dk = c.key_dim, N = c.N, E = c.E, D = c.D;
    static bool attr = false;
    if (!attr) {
        DOF_CUDA(cudaFuncSetAttribute(tfm_core_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        DOF_CUDA(cudaFuncSetAttribute(cens_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    for (int b = 0; b < 2; b++) {
        TfmCoreArgs t;
        memset(&t, 0, sizeof(t));
        t.x = b == 0 ? x : a; t.params = state + L.core[b]; t.out = b == 0 ? nodes : edges; t.ybuf = ybuf;
        t.B = B; t.T = c.T; t.G = b == 0 ? N : E; t.F = b == 0 ? c.F : c.Fe; t.dk = dk; t.heads = c.heads; t.dff = c.dff; t.layers = c.layers;
        const int S = B * t.G;
        const int grid = cdiv(S, TFM_GROUPS) < g_sm_count ? cdiv(S, TFM_GROUPS) : g_sm_count;
        const int per = one ? c.layers : 1;
        const double fl = (double)S * c.T * 2.0 * c.layers * (4.0 * dk * dk + 2.0 * dk * c.dff + 2.0 * c.T * dk);
        for (int l0 = 0; l0 < c.layers; l0 += per) {
            t.l_begin = l0; t.l_end = l0 + per;
            ProfScope ps("tfm_core_fwd", st, fl * per / c.layers, (double)S * (c.T * t.F + dk) * 4.0);
            tfm_core_fwd_kernel<<<grid, TFM_THREADS, tfm_core_smem_bytes(c.T, t.F, dk, c.dff, per), st>>>(t);
            DOF_LAUNCH_CHECK();
        }
    }
    if (nodes_out) DOF_CUDA(cudaMemcpyAsync(nodes_out, nodes, (size_t)B * N * dk * 4, cudaMemcpyDeviceToDevice, st));
    if (edges_out) DOF_CUDA(cudaMemcpyAsync(edges_out, edges, (size_t)B * E * dk * 4, cudaMemcpyDeviceToDevice, st));
    CensArgs ca;
    memset(&ca, 0, sizeof(ca));
    ca.node = nodes; ca.edge = edges; ca.lap = state + L.lap; ca.elap = state + L.elap; ca.inc = state + L.inc;
    ca.wn = state + L.node_weights; ca.we = state + L.edge_weights; ca.Pn = Pn; ca.Pe = Pe; ca.B = B; ca.N = N; ca.E = E; ca.C = dk;
    const size_t csm = cens_smem_floats(N, E, dk) * 4;
    if (csm > 200 * 1024) DOF_FAIL(DOF_ERR_UNSUPPORTED, "graph too large for the CensNet kernel (%zu B smem)", csm);
    { ProfScope ps("cens_fwd", st);
    cens_fwd_kernel<<<B, 128, csm, st>>>(ca); }
    DOF_LAUNCH_CHECK();
    GemmArgs g[2];
    g[0] = gemm_args(mv_plain(Pn, dk), state + L.node_kernel, D, 1, state + L.node_bias, On, D, B * N, D, dk);
    g[1] = gemm_args(mv_plain(Pe, dk), state + L.edge_kernel, D, 1, state + L.edge_bias, Oe, D, B * E, D, dk);
    g[0].relu = g[1].relu = 1;
    if (N == E) DOF_TRY(launch_gemm_rows(g, 2, st));
    else { DOF_TRY(launch_gemm_rows(g, 1, st)); DOF_TRY(launch_gemm_rows(g + 1, 1, st)); }
    TfmHeadArgs hargs;
    hargs.on = On; hargs.oe = Oe; hargs.w0 = state + L.w0; hargs.b0 = state + L.b0; hargs.bn2 = state + L.bn2; hargs.w3 = state + L.w3;
    hargs.b3 = state + L.b3; hargs.bn5 = state + L.bn5; hargs.w6 = state + L.w6; hargs.b6 = state + L.b6; hargs.out = enc_out;
    hargs.B = B; hargs.ND = N * D; hargs.ED = E * D; hargs.D = D;
    const size_t hsm = (size_t)4 * ((N + E) * D + 3 * D) * 4;
    if (hsm > 48 * 1024) DOF_FAIL(DOF_ERR_UNSUPPORTED, "head input of %d floats does not fit the head kernel", (N + E) * D);
    { ProfScope ps("tfm_head_fwd", st);
    tfm_head_fwd_kernel<<<cdiv(B, 4), 128, hsm, st>>>(hargs); }
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

}  // extern "C"

// ---- transformer decoder, eval forward ---------------------------------------------------------------------------
struct TfmDecLayout {
    std::vector<Entry> e;
    int64_t total = 0;
    int64_t ex_w[3], ex_b[3], layer[8], out_w, out_b, loc_w, loc_b;
};

static int check_tfm_dec_cfg(const dof_tfm_dec_cfg* c) {
    if (!c) DOF_FAIL(DOF_ERR_ARG, "null config");
    if (c->T < 1 || c->Dx < 1 || c->D < 1 || c->heads < 1 || c->dff < 1 || c->layers < 1 || c->layers > 8) DOF_FAIL(DOF_ERR_ARG, "bad decoder geometry");
    if ((4 * c->D) % c->heads) DOF_FAIL(DOF_ERR_ARG, "model_dim %d is not a multiple of heads %d", 4 * c->D, c->heads);
    if (c->T > TFM_MAXT) DOF_FAIL(DOF_ERR_UNSUPPORTED, "window length %d > %d", c->T, TFM_MAXT);
    if ((size_t)c->T * 12 * c->D * 4 > 200 * 1024) DOF_FAIL(DOF_ERR_UNSUPPORTED, "sequence of %d x %d does not fit the attention kernel", c->T, 12 * c->D);
    return DOF_OK;
}

static TfmDecLayout build_tfm_dec_layout(const dof_tfm_dec_cfg& c) {
    TfmDecLayout L;
    auto add = [&](const std::string& name, int d0, int d1 = -1) {
        Entry en;
        en.name = name; en.off = L.total; en.group = 2;
        en.shape[0] = d0; en.shape[1] = d1 < 0 ? 0 : d1; en.shape[2] = en.shape[3] = 0;
        en.ndim = d1 < 0 ? 1 : 2;
        en.numel = d1 < 0 ? d0 : (int64_t)d0 * d1;
        L.total += en.numel;
        L.e.push_back(en);
        return en.off;
    };
    const int D = c.D, dm = 4 * D;
    const int ein[3] = {D, D, 2 * D}, eout[3] = {D, 2 * D, 4 * D};
    for (int i = 0; i < 3; i++) {
        L.ex_w[i] = add("latent_expand." + std::to_string(2 * i) + ".weight", eout[i], ein[i]);
        L.ex_b[i] = add("latent_expand." + std::to_string(2 * i) + ".bias", eout[i]);
    }
    for (int l = 0; l < c.layers; l++) {
        std::string q = "layers." + std::to_string(l) + ".";
        L.layer[l] = add(q + "q_proj.weight", dm, dm); add(q + "k_proj.weight", dm, dm); add(q + "v_proj.weight", dm, dm);
        add(q + "out_proj.weight", dm, dm);
        add(q + "norm1.weight", dm); add(q + "norm1.bias", dm); add(q + "norm2.weight", dm); add(q + "norm2.bias", dm);
        add(q + "ffn.0.weight", c.dff, dm); add(q + "ffn.0.bias", c.dff); add(q + "ffn.3.weight", dm, c.dff); add(q + "ffn.3.bias", dm);
    }
    L.out_w = add("output_proj.weight", c.Dx, dm); L.out_b = add("output_proj.bias", c.Dx);
    L.loc_w = add("prob_decoder.loc_projection.weight", c.Dx, c.Dx); L.loc_b = add("prob_decoder.loc_projection.bias", c.Dx);
    return L;
}

extern "C" {

int64_t dof_tfm_dec_numel(const dof_tfm_dec_cfg* cfg) {
    if (check_tfm_dec_cfg(cfg) != DOF_OK) return -1;
    return build_tfm_dec_layout(*cfg).total;
}
int dof_tfm_dec_num_entries(const dof_tfm_dec_cfg* cfg) {
    if (check_tfm_dec_cfg(cfg) != DOF_OK) return -1;
    return (int)build_tfm_dec_layout(*cfg).e.size();
}
int dof_tfm_dec_entry(const dof_tfm_dec_cfg* cfg, int index, char* name_out, int64_t* offset_out, int64_t* numel_out, int* ndim_out,
                      int* shape_out) {
    DOF_TRY(check_tfm_dec_cfg(cfg));
    TfmDecLayout L = build_tfm_dec_layout(*cfg);
    if (index < 0 || index >= (int)L.e.size()) DOF_FAIL(DOF_ERR_ARG, "entry index %d out of range", index);
    const Entry& e = L.e[index];
    if (name_out) { strncpy(name_out, e.name.c_str(), 127); name_out[127] = 0; }
    if (offset_out) *offset_out = e.off;
    if (numel_out) *numel_out = e.numel;
    if (ndim_out) *ndim_out = e.ndim;
    if (shape_out) for (int i = 0; i < 4; i++) shape_out[i] = e.shape[i];
    return DOF_OK;
}

struct TfmDecWs { float *g0, *g1, *H, *Xn, *QKV, *A, *Fh, *Y, *mu, *rs; };
static size_t tfm_dec_plan(const dof_tfm_dec_cfg& c, int B, char* base, TfmDecWs* w) {
    Bump bp{base, 0, 0, base == nullptr};
    const size_t R = (size_t)B * c.T, dm = 4 * (size_t)c.D;
    TfmDecWs t;
    t.g0 = bp.get<float>((size_t)B * dm); t.g1 = bp.get<float>((size_t)B * dm);
    t.H = bp.get<float>(R * dm); t.Xn = bp.get<float>(R * dm); t.QKV = bp.get<float>(R * 3 * dm); t.A = bp.get<float>(R * dm);
    t.Fh = bp.get<float>(R * c.dff); t.Y = bp.get<float>(R * c.Dx); t.mu = bp.get<float>(R); t.rs = bp.get<float>(R);
    if (w) *w = t;
    return bp.off;
}

size_t dof_tfm_dec_workspace_bytes(const dof_tfm_dec_cfg* cfg, int B) {
    if (check_tfm_dec_cfg(cfg) != DOF_OK || B < 1) return 0;
    return tfm_dec_plan(*cfg, B, nullptr, nullptr);
}

static int gelu_inplace(float* x, long long n, cudaStream_t st) {
    ProfScope ps("gelu", st, 0.0, 8.0 * n);
    gelu_kernel<<<cdiv(n, 256), 256, 0, st>>>(x, n);
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

// TFMDecoderPT.forward in eval mode: z [B,D] -> loc [B,T,Dx] (the mean of the reconstruction distribution; the validity
// mask of x_target only scales the distribution, ProbabilisticDecoderPT models_new.py:677-710)
int dof_tfm_decode(const dof_tfm_dec_cfg* cfg, const float* state, const float* z, int B, void* workspace, size_t workspace_bytes,
                   float* loc_out, void* stream) {
    DOF_TRY(check_tfm_dec_cfg(cfg));
    if (!state || !z || !workspace || !loc_out || B < 1) DOF_FAIL(DOF_ERR_ARG, "null / bad argument");
    const dof_tfm_dec_cfg& c = *cfg;
    TfmDecWs w;
    const size_t need = tfm_dec_plan(c, B, (char*)workspace, &w);
    if (need > workspace_bytes) DOF_FAIL(DOF_ERR_ARG, "workspace too small: %zu < %zu bytes", workspace_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    const TfmDecLayout L = build_tfm_dec_layout(c);
    const int D = c.D, dm = 4 * D, T = c.T, R = B * T;
    // latent expansion: Linear + GELU three times (:1192-1199)
    const int ein[3] = {D, D, 2 * D}, eout[3] = {D, 2 * D, 4 * D};
    const float* cur = z;
    float* bufs[2] = {w.g0, w.g1};
    for (int i = 0; i < 3; i++) {
        float* o = bufs[i & 1];
        GemmArgs g = gemm_args(mv_plain(cur, ein[i]), state + L.ex_w[i], ein[i], 0, state + L.ex_b[i], o, eout[i], B, eout[i], ein[i]);
        DOF_TRY(launch_gemm_rows(&g, 1, st));
        DOF_TRY(gelu_inplace(o, (long long)B * eout[i], st));
        cur = o;
    }
    { ProfScope ps("tfm_dec_input", st);
    tfm_dec_input_kernel<<<cdiv((long long)R * dm, 256), 256, 0, st>>>(cur, w.H, B, T, dm); }
    DOF_LAUNCH_CHECK();
    const size_t asmem = (size_t)T * 3 * dm * 4;
    if (asmem > 48 * 1024) DOF_CUDA(cudaFuncSetAttribute(tfm_causal_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int l = 0; l < c.layers; l++) {
        const float* P = state + L.layer[l];
        const float *Wq = P, *Wo = P + (size_t)3 * dm * dm, *n1 = Wo + (size_t)dm * dm, *n2 = n1 + 2 * dm;
        const float *W1 = n2 + 2 * dm, *b1 = W1 + (size_t)c.dff * dm, *W2 = b1 + c.dff, *b2 = W2 + (size_t)dm * c.dff;
        DOF_TRY(ln_fwd(w.H, n1, n1 + dm, w.Xn, w.mu, w.rs, R, dm, g_sm_count, st, 1e-6f));
        if (3 * dm <= 256) {                                   // q | k | v weights are consecutive [dm, dm] tensors = one [3dm, dm]
            GemmArgs g = gemm_args(mv_plain(w.Xn, dm), Wq, dm, 0, nullptr, w.QKV, 3 * dm, R, 3 * dm, dm);
            DOF_TRY(launch_gemm_rows(&g, 1, st));
        } else {
            for (int m = 0; m < 3; m++) {
                GemmArgs g = gemm_args(mv_plain(w.Xn, dm), Wq + (size_t)m * dm * dm, dm, 0, nullptr, w.QKV + (size_t)m * dm, 3 * dm, R, dm, dm);
                DOF_TRY(launch_gemm_rows(&g, 1, st));
            }
        }
        { ProfScope ps("tfm_causal_attn", st, 4.0 * B * T * T * dm / 2.0, (double)R * 4 * dm * 4);
        tfm_causal_attn_kernel<<<B, 128, asmem, st>>>(w.QKV, w.A, T, dm, c.heads); }
        DOF_LAUNCH_CHECK();
        GemmArgs go = gemm_args(mv_plain(w.A, dm), Wo, dm, 0, nullptr, w.H, dm, R, dm, dm);
        go.accum = 1;                                          // x = x + out_proj(attn)
        DOF_TRY(launch_gemm_rows(&go, 1, st));
        DOF_TRY(ln_fwd(w.H, n2, n2 + dm, w.Xn, w.mu, w.rs, R, dm, g_sm_count, st, 1e-6f));
        GemmArgs g1 = gemm_args(mv_plain(w.Xn, dm), W1, dm, 0, b1, w.Fh, c.dff, R, c.dff, dm);
        DOF_TRY(launch_gemm_rows(&g1, 1, st));
        DOF_TRY(gelu_inplace(w.Fh, (long long)R * c.dff, st));
        GemmArgs g2 = gemm_args(mv_plain(w.Fh, c.dff), W2, c.dff, 0, b2, w.H, dm, R, dm, c.dff);
        g2.accum = 1;                                          // x = x + ffn(norm2(x))
        DOF_TRY(launch_gemm_rows(&g2, 1, st));
    }
    GemmArgs gy = gemm_args(mv_plain(w.H, dm), state + L.out_w, dm, 0, state + L.out_b, w.Y, c.Dx, R, c.Dx, dm);
    DOF_TRY(launch_gemm_rows(&gy, 1, st));
    GemmArgs gl = gemm_args(mv_plain(w.Y, c.Dx), state + L.loc_w, c.Dx, 0, state + L.loc_b, loc_out, c.Dx, R, c.Dx, c.Dx);
    DOF_TRY(launch_gemm_rows(&gl, 1, st));
    return DOF_OK;
}

}  // extern "C"

template <int DP>
static int ntx_launch(const NtxArgs& n, int B, cudaStream_t st) {
    const size_t smem = ((size_t)NTX_TILE * (DP + 1) + 4 * NTX_TILE + (size_t)NTX_WARPS * DP) * 4;
    static bool attr = false;
    if (!attr && smem > 48 * 1024) {
        DOF_CUDA(cudaFuncSetAttribute(ntx_kernel<0, DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DOF_CUDA(cudaFuncSetAttribute(ntx_kernel<1, DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    const int nb = cdiv(B, NTX_WARPS);
    const double fl = 2.0 * B * (double)B * n.D;
    if (n.kind == 3) {
        // fc: every active warp keeps its row of B similarities in shared memory
        const int bpad = round_up(B, 32);
        const size_t base = ((size_t)NTX_TILE * (DP + 1) + (size_t)NTX_WARPS * DP) * 4;
        int rw = (int)((200 * 1024 - base) / ((size_t)bpad * 4));
        if (rw > NTX_WARPS) rw = NTX_WARPS;
        if (rw < 1) DOF_FAIL(DOF_ERR_UNSUPPORTED, "fc loss: batch %d does not fit the per-row shared-memory buffer", B);
        const size_t fsmem = base + (size_t)rw * bpad * 4;
        DOF_CUDA(cudaFuncSetAttribute(ntx_fc_rows_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        { ProfScope ps("ntxent_fc_rows", st, fl, 0.0);
        ntx_fc_rows_kernel<DP><<<cdiv(B, rw), NTX_WARPS * 32, fsmem, st>>>(n, rw, bpad); }
        DOF_LAUNCH_CHECK();
    } else {
        { ProfScope ps("ntxent_rows", st, fl, 0.0);
        ntx_kernel<0, DP><<<nb, NTX_WARPS * 32, smem, st>>>(n); }
        DOF_LAUNCH_CHECK();
    }
    { ProfScope ps("ntxent_grad", st, 4.0 * fl, 0.0);
    ntx_kernel<1, DP><<<2 * nb, NTX_WARPS * 32, smem, st>>>(n); }
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

extern "C" {

// step_contrastive_distill forward + backward without teacher (training.py:527-545, 159-163) on the two views
// produced by dof_contrastive_views: x2 / a2 rows 0..B-1 = main view, B..2B-1 = augmented view.
int dof_contrastive_loss_grad(dof_handle* h, const float* state, float* grad, const float* x2, const float* a2, int B,
                              int loss_kind, int sim_kind, float temperature, float tau_plus, float beta, float* logs,
                              float* z_out, void* stream) {
    return dof_contrastive_loss_grad_distill(h, state, grad, x2, a2, B, loss_kind, sim_kind, temperature, tau_plus, beta, nullptr,
                                             logs, z_out, stream);
}

static int contrastive_step(dof_handle* h, const float* state, float* grad, const float* x2, const float* a2, int B,
                            int loss_kind, int sim_kind, float temperature, float tau_plus, float beta,
                            const dof_distill_cfg* distill, float* logs, float* z_out, void* stream, bool train);

int dof_contrastive_loss_grad_distill(dof_handle* h, const float* state, float* grad, const float* x2, const float* a2, int B,
                                      int loss_kind, int sim_kind, float temperature, float tau_plus, float beta,
                                      const dof_distill_cfg* distill, float* logs, float* z_out, void* stream) {
    return contrastive_step(h, state, grad, x2, a2, B, loss_kind, sim_kind, temperature, tau_plus, beta, distill, logs, z_out, stream, true);
}

// the validation step of fit_contrastive: eval-mode encoder on both views, the loss, no gradient, teacher off
int dof_contrastive_loss_eval(dof_handle* h, const float* state, const float* x2, const float* a2, int B, int loss_kind, int sim_kind,
                              float temperature, float tau_plus, float beta, float* logs, float* z_out, void* stream) {
    return contrastive_step(h, state, nullptr, x2, a2, B, loss_kind, sim_kind, temperature, tau_plus, beta, nullptr, logs, z_out, stream, false);
}

static int contrastive_step(dof_handle* h, const float* state, float* grad, const float* x2, const float* a2, int B,
                            int loss_kind, int sim_kind, float temperature, float tau_plus, float beta,
                            const dof_distill_cfg* distill, float* logs, float* z_out, void* stream, bool train) {
    DOF_TRY(check_batch(h, 2 * B));
    const dof_config& c = h->cfg;
    if (c.model != DOF_MODEL_CONTRASTIVE) DOF_FAIL(DOF_ERR_ARG, "handle is not a contrastive model");
    if (!h->training) DOF_FAIL(DOF_ERR_ARG, "handle was created with training=0");
    if (!state || (train && !grad) || !x2 || !a2 || !logs || !(temperature > 0.f)) DOF_FAIL(DOF_ERR_ARG, "null / bad argument");
    if (sim_kind < 0 || sim_kind > 1) DOF_FAIL(DOF_ERR_UNSUPPORTED, "similarity kind %d (0 cosine / dot, 1 euclidean / edit)", sim_kind);
    if (loss_kind < 0 || loss_kind > 3) DOF_FAIL(DOF_ERR_UNSUPPORTED, "contrastive loss kind %d (0 nce, 1 dcl, 2 hard_dcl, 3 fc)", loss_kind);
    if ((loss_kind == 1 || loss_kind == 2) && !(tau_plus >= 0.f && tau_plus < 1.f)) DOF_FAIL(DOF_ERR_ARG, "tau_plus must be in [0, 1)");
    if (c.D > 64) DOF_FAIL(DOF_ERR_UNSUPPORTED, "latent_dim %d > 64", c.D);
    cudaStream_t st = (cudaStream_t)stream;
    if (train) DOF_CUDA(cudaMemsetAsync(grad, 0, (size_t)h->L.total * 4, st));
    h->next_groups = 2;                 // the reference encodes the two views in two passes: separate batch statistics
    int erc = encoder_forward(h, state, x2, a2, 2 * B, train, st);
    h->next_groups = 1;
    DOF_TRY(erc);
    if (z_out) DOF_CUDA(cudaMemcpyAsync(z_out, h->enc, (size_t)2 * B * c.D * 4, cudaMemcpyDeviceToDevice, st));
    NtxArgs n;
    n.enc = h->enc; n.zn = h->zn; n.nrm = h->nrm; n.lse = h->lse; n.denc = h->denc; n.stats = h->nstats; n.logs = logs;
    n.B = B; n.D = c.D; n.inv_tau = 1.0f / temperature;
    n.kind = loss_kind; n.sim = sim_kind; n.tau_plus = tau_plus; n.beta = beta; n.temperature = temperature;
    DOF_CUDA(cudaMemsetAsync(h->nstats, 0, 8 * sizeof(double), st));
    { ProfScope ps("ntxent_norm", st);
    ntx_norm_kernel<<<cdiv(2LL * B * 32, 256), 256, 0, st>>>(n); }
    DOF_LAUNCH_CHECK();
    if (c.D <= 8) DOF_TRY(ntx_launch<8>(n, B, st));
    else if (c.D <= 16) DOF_TRY(ntx_launch<16>(n, B, st));
    else if (c.D <= 32) DOF_TRY(ntx_launch<32>(n, B, st));
    else DOF_TRY(ntx_launch<64>(n, B, st));
    // training.py:533-557: the head sees the row-NORMALISED embedding of the MAIN view (z is reassigned before z_main = z)
    const bool dist_on = train && distill && distill->lambda > 0.f;
    if (dist_on) DOF_TRY(distill_head_step(h, distill, B, h->nstats + NTX_ST_DISTILL, true, st));
    { ProfScope ps("ntxent_finalize", st);
    ntx_finalize_kernel<<<1, 1, 0, st>>>(n, temperature, dist_on ? distill->lambda : 0.f); }
    DOF_LAUNCH_CHECK();
    if (train) DOF_TRY(encoder_backward(h, state, grad, 2 * B, st));
    h->lastB = 2 * B;
    register_debug(h, 2 * B);
    return DOF_OK;
}

int dof_clip_adam(dof_handle* h, float* state, const float* grad, float* adam_m, float* adam_v, const dof_adam_cfg* opt,
                  void* stream) {
    if (!h || !state || !grad || !adam_m || !adam_v || !opt) DOF_FAIL(DOF_ERR_ARG, "null argument");
    AdamArgs a;
    memset(&a, 0, sizeof(a));
    a.p = state; a.g = grad; a.m = adam_m; a.v = adam_v; a.group = h->group; a.n = h->L.total;
    for (int g = 0; g < 4; g++) {
        a.lr[g] = opt->lr[g];
        a.wd[g] = opt->weight_decay[g];
        a.active[g] = opt->active[g];
        int s = opt->step[g] > 0 ? opt->step[g] : 1;
        a.bc1[g] = (float)(1.0 - pow((double)opt->beta1, s));
        a.bc2_sqrt[g] = (float)sqrt(1.0 - pow((double)opt->beta2, s));
    }
    a.clip = opt->clip_value; a.gscale = opt->grad_scale; a.beta1 = opt->beta1; a.beta2 = opt->beta2; a.eps = opt->eps;
    { ProfScope ps("clip_adam", (cudaStream_t)stream, 0.0, 28.0 * a.n);
    clip_adam_kernel<<<cdiv(a.n, 256), 256, 0, (cudaStream_t)stream>>>(a); }
    DOF_LAUNCH_CHECK();
    if (is_tcn(h) && h->bn_pending) DOF_TRY(tcn_bn_apply(h, state, (cudaStream_t)stream));
    if ((is_tfm(h) || is_tcn(h)) && h->bn_pending) DOF_TRY(tfm_bn_apply(h, state, (cudaStream_t)stream));
    return DOF_OK;
}

size_t dof_dropout_mask_bytes(const dof_config* cfg, int Bw, int B, int dec_passes) {
    if (check_cfg(cfg) != DOF_OK || cfg->encoder != DOF_ENCODER_TRANSFORMER || Bw < 1 || B < 0 || dec_passes < 0 || dec_passes > 2) return 0;
    const Layout L = build_layout(*cfg);
    return drop_plan(*cfg, L, Bw, B, dec_passes).total;
}

int dof_set_noise_seed(dof_handle* h, unsigned long long seed) {
    if (!h) DOF_FAIL(DOF_ERR_ARG, "null handle");
    h->noise_seed = seed ? seed : 0x9E3779B97F4A7C15ull;
    return DOF_OK;
}

int dof_set_dropout(dof_handle* h, unsigned long long seed, const unsigned char* masks, size_t mask_bytes) {
    if (!h) DOF_FAIL(DOF_ERR_ARG, "null handle");
    if (!is_tfm(h)) DOF_FAIL(DOF_ERR_ARG, "dropout only exists in the transformer family");
    h->drop_seed = seed; h->drop_masks = masks; h->drop_mask_bytes = masks ? mask_bytes : 0;
    return DOF_OK;
}

// Adam on a caller-owned flat buffer (the distillation head): no clipping (clip_grad_value_ only sees the model,
// training.py:165), weight decay as in build_optimizer_generic (losses.py:805-814)
int dof_adam_flat(float* param, const float* grad, float* adam_m, float* adam_v, long long n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int step, float grad_scale, void* stream) {
    if (!param || !grad || !adam_m || !adam_v || n < 1) DOF_FAIL(DOF_ERR_ARG, "null / bad argument");
    AdamArgs a;
    memset(&a, 0, sizeof(a));
    a.p = param; a.g = grad; a.m = adam_m; a.v = adam_v; a.group = nullptr; a.n = n;
    const int s = step > 0 ? step : 1;
    a.lr[1] = lr; a.wd[1] = weight_decay; a.active[1] = 1;
    a.bc1[1] = (float)(1.0 - pow((double)beta1, s));
    a.bc2_sqrt[1] = (float)sqrt(1.0 - pow((double)beta2, s));
    a.clip = 0.f; a.gscale = grad_scale; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
    { ProfScope ps("clip_adam", (cudaStream_t)stream, 0.0, 28.0 * n);
    clip_adam_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(a); }
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

// ---- gradient exchange over NVLink peer memory (peer.cuh) ----
int dof_peer_barrier(const void* const* pads, int world, int rank, int epoch, void* stream) {
    if (!pads || world < 1 || world > DOF_PEER_MAX || rank < 0 || rank >= world) DOF_FAIL(DOF_ERR_ARG, "bad peer barrier arguments (world %d rank %d)", world, rank);
    PeerPtrs pp;
    memset(&pp, 0, sizeof(pp));
    for (int i = 0; i < world; i++) {
        if (!pads[i]) DOF_FAIL(DOF_ERR_ARG, "null signal pad of peer %d", i);
        pp.p[i] = const_cast<void*>(pads[i]);
    }
    { ProfScope ps("peer_barrier", (cudaStream_t)stream);
    peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pp, world, rank, epoch); }
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

int dof_peer_reduce(const void* const* peers, int world, float* out, long long n, void* stream) {
    if (!peers || !out || world < 1 || world > DOF_PEER_MAX || n < 4 || (n & 3)) DOF_FAIL(DOF_ERR_ARG, "bad peer reduce arguments (world %d n %lld)", world, n);
    if (!aligned16(out)) DOF_FAIL(DOF_ERR_ARG, "peer reduce output must be 16-byte aligned");
    PeerPtrs pp;
    memset(&pp, 0, sizeof(pp));
    for (int i = 0; i < world; i++) {
        if (!peers[i] || !aligned16(peers[i])) DOF_FAIL(DOF_ERR_ARG, "peer %d: null or misaligned gradient buffer", i);
        pp.p[i] = const_cast<void*>(peers[i]);
    }
    const long long n4 = n / 4;
    int grid = (int)((n4 + 255) / 256);
    if (grid > 4 * g_sm_count) grid = 4 * g_sm_count;
    { ProfScope ps("peer_reduce", (cudaStream_t)stream, 0.0, 4.0 * n * (world + 1));
    peer_reduce_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pp, world, reinterpret_cast<float4*>(out), n4); }
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

const void* dof_debug_tensor(dof_handle* h, const char* name, int64_t* numel_out) {
    if (!h || !name) return nullptr;
    auto it = h->dbg.find(name);
    if (it == h->dbg.end()) return nullptr;
    if (numel_out) *numel_out = it->second.second;
    return it->second.first;
}

long long dof_launch_count(void) { return g_prof.launches; }

// 1 = use the tcgen05 GEMM kernels where eligible (default), 0 = fp32 SIMT kernels only
int dof_set_tensor_cores(int enable) {
    tc_enabled();
    int old = g_tc_enabled ? 1 : 0;
    g_tc_enabled = enable != 0;
    return old;
}

// 1 (default): the node and edge recurrent blocks of the encoder run concurrently on two streams; 0: one stream
// (per-kernel event timing is only meaningful when kernels do not overlap).  Returns the previous setting.
int dof_set_concurrency(int enable) {
    int old = g_concurrent ? 1 : 0;
    g_concurrent = enable != 0;
    return old;
}

int dof_profile_begin(void) {
    for (auto& r : g_prof.recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.recs.clear();
    g_prof.enabled = true;
    return DOF_OK;
}

// Synchronises the device, aggregates the recorded launches per kernel class and writes
// "name count total_ms algorithmic_flops algorithmic_bytes\n" lines into out.
int dof_profile_end(char* out, size_t cap) {
    g_prof.enabled = false;
    DOF_CUDA(cudaDeviceSynchronize());
    struct Agg { long long n = 0; double ms = 0, flops = 0, bytes = 0; };
    std::map<std::string, Agg> agg;
    for (auto& r : g_prof.recs) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        Agg& e = agg[r.name];
        e.n++; e.ms += ms; e.flops += r.flops; e.bytes += r.bytes;
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    g_prof.recs.clear();
    std::string txt;
    char line[160];
    for (auto& kv : agg) {
        snprintf(line, sizeof(line), "%s %lld %.6f %.6e %.6e\n", kv.first.c_str(), kv.second.n, kv.second.ms,
                 kv.second.flops, kv.second.bytes);
        txt += line;
    }
    if (out && cap > 0) { strncpy(out, txt.c_str(), cap - 1); out[cap - 1] = 0; }
    return DOF_OK;
}

// ---- window loader -----------------------------------------
static int loader_params(const dof_loader_cfg* c, const float* frames, long long n_frames, LoaderP& p) {
    if (!c || !frames) DOF_FAIL(DOF_ERR_ARG, "null loader config / frame table");
    if (c->T < 1 || c->step < 1 || c->N < 1 || c->E < 0 || n_frames < 0)
        DOF_FAIL(DOF_ERR_ARG, "bad loader geometry T=%d step=%d N=%d E=%d frames=%lld", c->T, c->step, c->N, c->E, n_frames);
    if (c->N > LD_MAXN || c->E > LD_MAXE) DOF_FAIL(DOF_ERR_UNSUPPORTED, "loader supports N <= %d, E <= %d (got %d, %d)", LD_MAXN, LD_MAXE, c->N, c->E);
    if (c->center_node >= c->N || c->align_node >= c->N) DOF_FAIL(DOF_ERR_ARG, "centre / align node out of range");
    if (!c->speed_scale || !c->speed_shift || (c->E > 0 && (!c->dist_div || !c->dist_scale || !c->dist_shift || !c->edges)))
        DOF_FAIL(DOF_ERR_ARG, "null per-column constants");
    memset(&p, 0, sizeof(p));
    p.frames = frames; p.n_frames = n_frames;
    p.T = c->T; p.step = c->step; p.N = c->N; p.E = c->E; p.center_node = c->center_node; p.align_node = c->align_node;
    p.cx = c->cx; p.cy = c->cy; p.fps = c->fps; p.clip = c->clip; p.coord_scale = c->coord_scale; p.coord_shift = c->coord_shift;
    for (int n = 0; n < c->N; n++) { p.speed_scale[n] = c->speed_scale[n]; p.speed_shift[n] = c->speed_shift[n]; }
    for (int e = 0; e < c->E; e++) {
        if (c->edges[2 * e] < 0 || c->edges[2 * e] >= c->N || c->edges[2 * e + 1] < 0 || c->edges[2 * e + 1] >= c->N)
            DOF_FAIL(DOF_ERR_ARG, "edge %d = (%d, %d) out of range", e, c->edges[2 * e], c->edges[2 * e + 1]);
        if (!(c->dist_div[e] > 0.0)) DOF_FAIL(DOF_ERR_ARG, "dist_div[%d] must be > 0", e);
        p.dist_div[e] = c->dist_div[e]; p.dist_scale[e] = c->dist_scale[e]; p.dist_shift[e] = c->dist_shift[e];
        p.e0[e] = (short)c->edges[2 * e]; p.e1[e] = (short)c->edges[2 * e + 1];
    }
    return DOF_OK;
}

static int loader_device_check() {
    int dev = 0;
    DOF_CUDA(cudaGetDevice(&dev));
    int major = 0;
    DOF_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major < 10) DOF_FAIL(DOF_ERR_UNSUPPORTED, "deepof_b200 needs an sm_100-class GPU");
    if (g_sm_count <= 0) DOF_CUDA(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
    return DOF_OK;
}

long long dof_loader_num_windows(long long n_frames, int T, int step) {
    if (T < 1 || step < 1 || n_frames < T) return 0;
    return (n_frames - T) / step + 1;                       // utils.py:3354-3377
}

int dof_load_windows(const dof_loader_cfg* cfg, const float* frames, long long n_frames, long long first_window,
                     int count, float* x, float* a, void* stream) {
    LoaderP p;
    DOF_TRY(loader_params(cfg, frames, n_frames, p));
    if (count == 0) return DOF_OK;
    if (!x || (cfg->E > 0 && !a)) DOF_FAIL(DOF_ERR_ARG, "null output");
    const long long nw = dof_loader_num_windows(n_frames, cfg->T, cfg->step);
    if (first_window < 0 || count < 0 || first_window + count > nw)
        DOF_FAIL(DOF_ERR_ARG, "windows [%lld, %lld) outside [0, %lld)", first_window, first_window + count, nw);
    DOF_TRY(loader_device_check());
    int wpb = 8;
    LoaderSmem so = loader_smem(p.T, p.step, p.N, p.E, wpb);
    while (so.total > 100 * 1024 && wpb > 4) { wpb -= 4; so = loader_smem(p.T, p.step, p.N, p.E, wpb); }
    if (so.total > 220 * 1024) DOF_FAIL(DOF_ERR_UNSUPPORTED, "window tile needs %zu B of shared memory", so.total);
    p.w_start = first_window; p.B = count; p.wpb = wpb; p.x = x; p.a = a;
    p.bulk_in = ((uintptr_t)frames % 16 == 0) ? 1 : 0;
    p.bulk_out = ((uintptr_t)x % 16 == 0 && (uintptr_t)a % 16 == 0) ? 1 : 0;
    static size_t attr_set = 0;
    if (so.total > 48 * 1024 && so.total > attr_set) {
        DOF_CUDA(cudaFuncSetAttribute(load_windows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)so.total));
        attr_set = so.total;
    }
    const double C = 3.0 * p.N + p.E;
    { ProfScope ps("load_windows", (cudaStream_t)stream, 0.0, ((double)(count - 1) * p.step + p.T) * 2 * p.N * 4 + (double)count * p.T * C * 4);
    load_windows_kernel<<<cdiv(count, wpb), LD_THREADS, so.total, (cudaStream_t)stream>>>(p, so); }
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

int dof_loader_pair_length(const float* frames, long long n_frames, int N, int node_a, int node_b, double* out, void* stream) {
    if (!frames || !out || N < 1 || node_a < 0 || node_a >= N || node_b < 0 || node_b >= N) DOF_FAIL(DOF_ERR_ARG, "bad argument");
    if (n_frames <= 0) return DOF_OK;
    DOF_TRY(loader_device_check());
    int grid = cdiv(n_frames, 256) < g_sm_count * 8 ? cdiv(n_frames, 256) : g_sm_count * 8;
    { ProfScope ps("loader_pair_length", (cudaStream_t)stream);
    loader_pair_length_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(frames, n_frames, N, node_a, node_b, out); }
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

int dof_loader_moments(const dof_loader_cfg* cfg, const float* frames, long long n_frames, const double* shift3, double* out9,
                       void* stream) {
    LoaderP p;
    DOF_TRY(loader_params(cfg, frames, n_frames, p));
    if (!shift3 || !out9) DOF_FAIL(DOF_ERR_ARG, "null argument");
    if (n_frames <= 0) return DOF_OK;
    DOF_TRY(loader_device_check());
    const long long total = n_frames * (3LL * p.N + p.E);
    int grid = cdiv(total, 256) < g_sm_count * 8 ? cdiv(total, 256) : g_sm_count * 8;
    { ProfScope ps("loader_moments", (cudaStream_t)stream);
    loader_moments_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, shift3[0], shift3[1], shift3[2], out9); }
    DOF_LAUNCH_CHECK();
    return DOF_OK;
}

// ---- op-level test hooks -----------------------------------------------------
static MatView make_view(const float* p, int ld, int mode, int p0, int p1) {
    switch (mode) {
        case A_SPLIT: return mv_split(p, ld, p0, p1);
        case A_CONV5: return mv_conv5(p, ld, p0, p1);
        case A_TSHIFT: return mv_tshift(p, ld, p0, p1);
        default: return mv_plain(p, ld);
    }
}

int dof_test_gemm_rows(const float* A, int lda, int mode, int p0, int p1, int p2, const float* W, int ldw, int wT,
                       const float* bias, float* C, int ldc, int M, int N, int K, int relu, int accum,
                       const float* mask, void* stream) {
    (void)p2;
    GemmArgs g = gemm_args(make_view(A, lda, mode, p0, p1), W, ldw, wT, bias, C, ldc, M, N, K);
    g.relu = relu; g.accum = accum; g.mask = mask; g.ldmask = ldc;
    return launch_gemm_rows(&g, 1, (cudaStream_t)stream);
}

int dof_test_gemm_wgrad(const float* P, int ldp, int pmode, int pp0, int pp1, const float* Q, int ldq, int qmode,
                        int qp0, int qp1, float* dW, int ldo, int oT, float* db, int M, int N, int K, void* stream) {
    WGradArgs g = wgrad_args(make_view(P, ldp, pmode, pp0, pp1), make_view(Q, ldq, qmode, qp0, qp1), dW, ldo, oT, db, M, N, K);
    int dev = 0, sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
    return launch_gemm_wgrad(&g, 1, (cudaStream_t)stream, sm);
}

int dof_test_gru_fwd(const float* gi_f, const float* gi_b, long long gi_ss, int gi_st, const float* whh_f,
                     const float* whh_b, const float* bhh_f, const float* bhh_b, const int* len, float* hout,
                     float* gt_f, float* gt_b, float* hn, int S, int T, int H, void* stream) {
    GruFwdArgs f;
    memset(&f, 0, sizeof(f));
    f.Gi[0] = gi_f; f.Gi[1] = gi_b; f.gi_ss = gi_ss; f.gi_st = gi_st;
    f.Whh[0] = whh_f; f.Whh[1] = whh_b; f.bhh[0] = bhh_f; f.bhh[1] = bhh_b;
    f.len = len; f.Hout = hout; f.Gt[0] = gt_f; f.Gt[1] = gt_b; f.Hn = hn; f.S = S; f.T = T; f.H = H;
    return launch_gru_fwd(f, (cudaStream_t)stream);
}

static long long* g_gru_bwdw_dbg = nullptr;
// fused tcgen05 GRU layer (input projection + recurrence + gates); fails if the shape is not eligible
int dof_test_gru_layer_fwd(const float* X, long long x_ss, int x_st, const float* const* w8, const int* len, float* hout,
                           float* gt_f, float* gt_b, float* hn, int S, int T, int H, int I, int gt_tiled, void* stream) {
    if (!gru_tc_eligible(S, H, I)) DOF_FAIL(DOF_ERR_UNSUPPORTED, "shape S=%d H=%d I=%d is not eligible for the fused GRU kernel", S, H, I);
    GruTcArgs a;
    memset(&a, 0, sizeof(a));
    a.X = X; a.x_ss = x_ss; a.x_st = x_st;
    for (int d = 0; d < 2; d++) { a.Wih[d] = w8[d]; a.Whh[d] = w8[2 + d]; a.bih[d] = w8[4 + d]; a.bhh[d] = w8[6 + d]; }
    a.len = len; a.Hout = hout; a.Gt[0] = gt_f; a.Gt[1] = gt_b; a.Hn = hn; a.S = S; a.T = T; a.H = H; a.I = I;
    a.gt_tiled = gt_tiled;
    a.dbg = g_gru_bwdw_dbg;                 // the profiling hook of dof_test_gru_bwdw_timeline serves both fused kernels
    return launch_gru_fwd_tc(a, (cudaStream_t)stream);
}

// fused tcgen05 BPTT (gate gradients + recurrent matmul + input gradient); gates in the tiled layout
int dof_test_gru_layer_bwd(const float* const* w8, const int* len, const float* hout, const float* gtT_f, const float* gtT_b,
                           const float* dout, const float* dhn, float* dg_f, float* dg_b, float* dx, const float* dxmask,
                           int S, int T, int H, int I, void* stream) {
    if (!gru_bwd_tc_eligible(H, I)) DOF_FAIL(DOF_ERR_UNSUPPORTED, "shape H=%d I=%d is not eligible for the fused GRU backward kernel", H, I);
    GruBwdTcArgs b;
    memset(&b, 0, sizeof(b));
    for (int d = 0; d < 2; d++) { b.Wih[d] = w8[d]; b.Whh[d] = w8[2 + d]; }
    b.GtT[0] = gtT_f; b.GtT[1] = gtT_b; b.dG[0] = dg_f; b.dG[1] = dg_b;
    b.len = len; b.Hout = hout; b.dOut = dout; b.dHn = dhn; b.dX = dx; b.dXmask = dxmask; b.S = S; b.T = T; b.H = H; b.I = I;
    if (dx) DOF_CUDA(cudaMemsetAsync(dx, 0, (size_t)S * T * I * 4, (cudaStream_t)stream));
    return launch_gru_bwd_tc(b, (cudaStream_t)stream);
}

// second-generation fused backward: BPTT + dX + the four parameter gradients of both directions in one kernel.
// out = per direction [dW_ih (3H x I) | dW_hh (3H x H) | db_ih (3H) | db_hh (3H)], accumulated into (zero it first)
int dof_test_gru_layer_bwdw(const float* X, const float* const* w8, const int* len, const float* hout, const float* gtT_f,
                            const float* gtT_b, const float* dout, const float* dhn, float* dx, const float* dxmask, float* out,
                            int S, int T, int H, int I, void* stream) {
    if (!gru_bwdw_tc_eligible(H, I)) DOF_FAIL(DOF_ERR_UNSUPPORTED, "shape H=%d I=%d is not eligible for the fused GRU backward + weight-gradient kernel", H, I);
    GruBwdwArgs b;
    memset(&b, 0, sizeof(b));
    const size_t per = (size_t)3 * H * I + (size_t)3 * H * H + 6 * H;
    for (int d = 0; d < 2; d++) {
        b.Wih[d] = w8[d]; b.Whh[d] = w8[2 + d];
        b.dWih[d] = out + d * per; b.dWhh[d] = b.dWih[d] + 3 * H * I; b.dbih[d] = b.dWhh[d] + 3 * H * H; b.dbhh[d] = b.dbih[d] + 3 * H;
    }
    b.GtT[0] = gtT_f; b.GtT[1] = gtT_b; b.X = X; b.x_ss = (long long)T * I; b.x_st = I;
    b.len = len; b.Hout = hout; b.dOut = dout; b.dHn = dhn; b.dX = dx; b.dXmask = dxmask; b.S = S; b.T = T; b.H = H; b.I = I;
    b.dbg = g_gru_bwdw_dbg;
    DOF_CUDA(cudaMemsetAsync(dx, 0, (size_t)S * T * I * 4, (cudaStream_t)stream));
    return launch_gru_bwdw_tc(b, (cudaStream_t)stream);
}

// test / profiling hook: device buffer [8][T][4] of clock64 stamps written by CTA (0, 0) of the next dof_test_gru_layer_bwdw
// launches (NULL = off)
int dof_test_gru_bwdw_timeline(long long* dbg) { g_gru_bwdw_dbg = dbg; return DOF_OK; }

// test hook: the encoder alone in TRAIN mode, backward from a given d(loss)/d(encoder output) (any model kind, any
// encoder family): grad is overwritten, enc_out [B, D] receives the train-mode encoder output.
int dof_test_encoder_grad(dof_handle* h, const float* state, float* grad, const float* x, const float* a, int B, int groups,
                          const float* denc, float* enc_out, void* stream) {
    DOF_TRY(check_batch(h, B));
    if (!h->training) DOF_FAIL(DOF_ERR_ARG, "handle was created with training=0");
    if (!state || !grad || !x || !a || !denc) DOF_FAIL(DOF_ERR_ARG, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    DOF_CUDA(cudaMemsetAsync(grad, 0, (size_t)h->L.total * 4, st));
    h->next_groups = groups > 0 ? groups : 1;
    int rc = encoder_forward(h, state, x, a, B, true, st);
    h->next_groups = 1;
    DOF_TRY(rc);
    if (enc_out) DOF_CUDA(cudaMemcpyAsync(enc_out, h->enc, (size_t)B * h->cfg.D * 4, cudaMemcpyDeviceToDevice, st));
    DOF_CUDA(cudaMemcpyAsync(h->denc, denc, (size_t)B * h->cfg.D * 4, cudaMemcpyDeviceToDevice, st));
    DOF_TRY(encoder_backward(h, state, grad, B, st));
    h->lastB = B;
    register_debug(h, B);
    return DOF_OK;
}

// test hook: multi-head attention with key-padding / causal masks and dropout on the weights, forward or backward
// (building block of the transformer training step, tfm.cuh).  backward = (dout != NULL): writes dqkv.
int dof_test_tfm_attention(const float* qkv, const unsigned char* kpad, const unsigned char* keep, float rate, int causal, int S,
                           int T, int dm, int heads, float* out, const float* dout, float* dqkv, void* stream) {
    if (!qkv || S < 1 || T < 1 || T > TFM_MAXT || heads < 1 || dm % heads) DOF_FAIL(DOF_ERR_ARG, "bad attention arguments");
    if (!(rate >= 0.f && rate < 1.f)) DOF_FAIL(DOF_ERR_ARG, "dropout rate must be in [0, 1)");
    const bool bwd = dout != nullptr;
    if (bwd ? !dqkv : !out) DOF_FAIL(DOF_ERR_ARG, "null output");
    DropSite drop = drop_none();
    drop.keep = keep; drop.rate = keep ? rate : 0.f;
    return tfm_attention(qkv, kpad, drop, causal, S, T, dm, heads, 0, out, dout, dqkv, (cudaStream_t)stream);
}

int dof_test_tcn_conv(int mode, const float* X, int ldx, int cin, int T, int dilation, const float* W, const float* bias, float* A,
                      int C, long long R, float* dX, float* dW, float* db, void* stream) {
    if (!X || !W || !A || ldx < cin || cin < 1 || T < 1 || dilation < 1 || C < 1 || R < 1 || R % T) DOF_FAIL(DOF_ERR_ARG, "bad convolution arguments");
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
    g_sm_count = sm;
    if (mode == 0) return tcn_conv_fwd(X, ldx, cin, T, dilation, W, bias, A, C, R, st);
    if (mode == 1) {
        if (!dX || ldx != cin) DOF_FAIL(DOF_ERR_ARG, "input gradient: dX required, ldx == cin");
        return tcn_conv_dgrad(A, C, T, dilation, W, cin, dX, R, 0, nullptr, st);
    }
    if (mode == 2) {
        if (!dW) DOF_FAIL(DOF_ERR_ARG, "weight gradient: dW required");
        return tcn_conv_wgrad(A, C, X, ldx, cin, T, dilation, dW, db, R, sm, st);
    }
    DOF_FAIL(DOF_ERR_ARG, "mode %d", mode);
}

int dof_test_gru_wgrad(const float* dg_f, const float* dg_b, const float* x, int ldx, const float* hout, float* out, int M, int T,
                       int I, int H, void* stream) {
    if (!gru_wgrad_tc_eligible(M, H, I, ldx, dg_f, dg_b, x, hout))
        DOF_FAIL(DOF_ERR_UNSUPPORTED, "merged GRU weight gradient: shape M=%d H=%d I=%d pitch=%d is not eligible", M, H, I, ldx);
    float* dG[2] = {const_cast<float*>(dg_f), const_cast<float*>(dg_b)};
    const size_t per = (size_t)3 * H * I + (size_t)3 * H * H + 6 * H;
    float *dWih[2], *dWhh[2], *dbih[2], *dbhh[2];
    for (int d = 0; d < 2; d++) {
        dWih[d] = out + d * per; dWhh[d] = dWih[d] + 3 * H * I; dbih[d] = dWhh[d] + 3 * H * H; dbhh[d] = dbih[d] + 3 * H;
    }
    int dev = 0, sm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
    return launch_gru_wgrad_tc(dG, x, ldx, hout, dWih, dWhh, dbih, dbhh, M, T, I, H, sm, (cudaStream_t)stream);
}

int dof_test_gru_bwd(const float* whh_f, const float* whh_b, const int* len, const float* hout, const float* gt_f,
                     const float* gt_b, const float* dout, const float* dhn, float* dg_f, float* dg_b, int S, int T,
                     int H, void* stream) {
    GruBwdArgs b;
    memset(&b, 0, sizeof(b));
    b.Whh[0] = whh_f; b.Whh[1] = whh_b; b.len = len; b.Hout = hout; b.Gt[0] = gt_f; b.Gt[1] = gt_b;
    b.dOut = dout; b.dHn = dhn; b.dG[0] = dg_f; b.dG[1] = dg_b; b.S = S; b.T = T; b.H = H;
    return launch_gru_bwd(b, (cudaStream_t)stream);
}

int dof_test_layernorm(const float* x, const float* w, const float* b, float eps, float* y, float* mu, float* rstd,
                       const float* dy, float* dx, float* dw, float* db, long long R, int W, int relu_in,
                       void* stream) {
    (void)eps;
    cudaStream_t st = (cudaStream_t)stream;
    DOF_TRY(ln_fwd(x, w, b, y, mu, rstd, R, W, 148, st));
    if (dy) DOF_TRY(ln_bwd(dy, x, mu, rstd, w, dx, dw, db, R, W, relu_in, 148, st));
    return DOF_OK;
}

}  // extern "C"
