"""Parameter dictionaries (reference state_dict names and shapes) for the CPU oracle, built WITHOUT the product library.

TEST INFRASTRUCTURE ONLY (tests/, bench.py's CPU arm).  Names and shapes follow the reference modules
(deepof/clustering/models_new.py): RecurrentEncoderPT :37-118, RecurrentBlockPT :184-222, RecurrentDecoderPT :281-324,
TFMEncoderPT :991-1089, TransformerCorePT :922-953, TFMDecoderPT :1176-1230, CausalSelfAttentionLayer :1272-1300,
GaussianMixtureLatentPT :1679-1743, VectorQuantizerPT :1330-1356, CensNetConvPT (censNetConv_pt.py:40-90).
"""
from collections import OrderedDict

import torch


def _gru(d, pre, I, H):
    for suf in ("", "_reverse"):
        d[pre + "weight_ih_l0" + suf] = (3 * H, I)
        d[pre + "weight_hh_l0" + suf] = (3 * H, H)
        d[pre + "bias_ih_l0" + suf] = (3 * H,)
        d[pre + "bias_hh_l0" + suf] = (3 * H,)


def _tail(d, kind, D, K):
    if kind == "vqvae":
        d["vq_layer.codebook"] = (D, K)
    elif kind == "vade":
        d["latent_space.gmm_means"] = (K, D)
        d["latent_space.gmm_log_vars"] = (K, D)
        d["latent_space.prior"] = (K,)
        d["latent_space.pretrain"] = ()
        d["latent_space.encoder_mean.weight"] = (D, D)
        d["latent_space.encoder_mean.bias"] = (D,)
        d["latent_space.encoder_log_var.weight"] = (D, D)
        d["latent_space.encoder_log_var.bias"] = (D,)
        d["latent_space.lens.weight"] = (D, D)
        d["latent_space.lens.bias"] = (D,)


def recurrent_shapes(kind, N, E, F, Fe, D, K):
    d = OrderedDict()
    di = min(64, D)
    d["encoder.laplacian"], d["encoder.edge_laplacian"], d["encoder.incidence"] = (N, N), (E, E), (N, E)
    for blk, Fin in (("encoder.node_recurrent_block.", F), ("encoder.edge_recurrent_block.", Fe)):
        d[blk + "conv1d.weight"] = (2 * di, Fin, 5)
        _gru(d, blk + "gru1.", 2 * di, 2 * di)
        d[blk + "norm1.weight"], d[blk + "norm1.bias"] = (4 * di,), (4 * di,)
        _gru(d, blk + "gru2.", 4 * di, di)
        d[blk + "norm2.weight"], d[blk + "norm2.bias"] = (2 * di,), (2 * di,)
        d[blk + "projection.weight"], d[blk + "projection.bias"] = (2 * D, 2 * di), (2 * D,)
    g = "encoder.spatial_gnn_block."
    d[g + "node_kernel"], d[g + "edge_kernel"] = (2 * D, D), (2 * D, D)
    d[g + "node_weights"], d[g + "edge_weights"] = (2 * D, 1), (2 * D, 1)
    d[g + "node_bias"], d[g + "edge_bias"] = (D,), (D,)
    d["encoder.final_dense.weight"], d["encoder.final_dense.bias"] = (D, (N + E) * D), (D,)
    if kind == "contrastive":
        return d
    _gru(d, "decoder.gru1.", D, D)
    d["decoder.norm1.weight"], d["decoder.norm1.bias"] = (2 * D,), (2 * D,)
    _gru(d, "decoder.gru2.", 2 * D, 2 * D)
    d["decoder.norm2.weight"], d["decoder.norm2.bias"] = (4 * D,), (4 * D,)
    d["decoder.conv1d.weight"] = (2 * D, 4 * D, 5)
    d["decoder.norm3.weight"], d["decoder.norm3.bias"] = (2 * D,), (2 * D,)
    d["decoder.prob_decoder.loc_projection.weight"], d["decoder.prob_decoder.loc_projection.bias"] = (N * F, 2 * D), (N * F,)
    _tail(d, kind, D, K)
    return d


def transformer_shapes(kind, N, E, F, Fe, D, K, heads=4, dff=128, layers=2, dec_dff=128, dec_layers=2):
    d = OrderedDict()
    dk = max((min(64, N * F) // heads) * heads, heads)
    d["encoder.laplacian"], d["encoder.edge_laplacian"], d["encoder.incidence"] = (N, N), (E, E), (N, E)
    for core, Fin in (("encoder.node_tf.", F), ("encoder.edge_tf.", Fe)):
        d[core + "embed.weight"], d[core + "embed.bias"] = (dk, Fin), (dk,)
        for l in range(layers):
            q = core + f"layers.{l}."
            for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
                d[q + f"mha.{nm}.weight"] = (dk, dk)
            d[q + "norm1.weight"], d[q + "norm1.bias"] = (dk,), (dk,)
            d[q + "ffn.0.weight"], d[q + "ffn.0.bias"] = (dff, dk), (dff,)
            d[q + "ffn.2.weight"], d[q + "ffn.2.bias"] = (dk, dff), (dk,)
            d[q + "norm2.weight"], d[q + "norm2.bias"] = (dk,), (dk,)
    g = "encoder.spatial_gnn_block."
    d[g + "node_kernel"], d[g + "edge_kernel"] = (dk, D), (dk, D)
    d[g + "node_weights"], d[g + "edge_weights"] = (dk, 1), (dk, 1)
    d[g + "node_bias"], d[g + "edge_bias"] = (D,), (D,)
    d["encoder.head.0.weight"], d["encoder.head.0.bias"] = (2 * D, (N + E) * D), (2 * D,)
    for i, C in ((2, 2 * D), (5, D)):
        if i == 5:
            d["encoder.head.3.weight"], d["encoder.head.3.bias"] = (D, 2 * D), (D,)
        for nm in ("weight", "bias", "running_mean", "running_var"):
            d[f"encoder.head.{i}.{nm}"] = (C,)
        d[f"encoder.head.{i}.num_batches_tracked"] = ()
    d["encoder.head.6.weight"], d["encoder.head.6.bias"] = (D, D), (D,)
    if kind == "contrastive":
        return d
    dm, Dx = 4 * D, N * F
    for i, (o, k) in enumerate(((D, D), (2 * D, D), (4 * D, 2 * D))):
        d[f"decoder.latent_expand.{2 * i}.weight"], d[f"decoder.latent_expand.{2 * i}.bias"] = (o, k), (o,)
    for l in range(dec_layers):
        q = f"decoder.layers.{l}."
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            d[q + nm + ".weight"] = (dm, dm)
        for nm in ("norm1", "norm2"):
            d[q + nm + ".weight"], d[q + nm + ".bias"] = (dm,), (dm,)
        d[q + "ffn.0.weight"], d[q + "ffn.0.bias"] = (dec_dff, dm), (dec_dff,)
        d[q + "ffn.3.weight"], d[q + "ffn.3.bias"] = (dm, dec_dff), (dm,)
    d["decoder.output_proj.weight"], d["decoder.output_proj.bias"] = (Dx, dm), (Dx,)
    d["decoder.prob_decoder.loc_projection.weight"], d["decoder.prob_decoder.loc_projection.bias"] = (Dx, Dx), (Dx,)
    _tail(d, kind, D, K)
    return d


def _bn(d, pre, C):
    for nm in ("weight", "bias", "running_mean", "running_var"):
        d[pre + nm] = (C,)
    d[pre + "num_batches_tracked"] = ()


def _tcn_stack(d, pre, cin, C, nb):
    for i in range(nb):
        q = pre + f"blocks.{i}."
        d[q + "conv1.weight"], d[q + "conv1.bias"] = (C, cin, 4), (C,)
        _bn(d, q + "bn1.", C)
        d[q + "conv2.weight"], d[q + "conv2.bias"] = (C, C, 4), (C,)
        _bn(d, q + "bn2.", C)
        if cin != C:
            d[q + "downsample.weight"], d[q + "downsample.bias"] = (C, cin, 1), (C,)
        cin = C


def tcn_shapes(kind, N, E, F, Fe, D, K):
    """TCNEncoderPT :574-607, TCNDecoderPT :745-775 (models_new.py)."""
    d = OrderedDict()
    d["encoder.laplacian"], d["encoder.edge_laplacian"], d["encoder.incidence"] = (N, N), (E, E), (N, E)
    _tcn_stack(d, "encoder.node_tcn.", F, 32, 8)
    _tcn_stack(d, "encoder.edge_tcn.", Fe, 32, 8)
    g = "encoder.spatial_gnn_block."
    d[g + "node_kernel"], d[g + "edge_kernel"] = (32, D), (32, D)
    d[g + "node_weights"], d[g + "edge_weights"] = (32, 1), (32, 1)
    d[g + "node_bias"], d[g + "edge_bias"] = (D,), (D,)
    d["encoder.head.0.weight"], d["encoder.head.0.bias"] = (2 * D, (N + E) * D), (2 * D,)
    _bn(d, "encoder.head.2.", 2 * D)
    d["encoder.head.3.weight"], d["encoder.head.3.bias"] = (D, 2 * D), (D,)
    _bn(d, "encoder.head.5.", D)
    d["encoder.head.6.weight"], d["encoder.head.6.bias"] = (D, D), (D,)
    if kind == "contrastive":
        return d
    for i, (o, k) in enumerate(((D, D), (2 * D, D), (4 * D, 2 * D))):
        d[f"decoder.fc{i}.weight"], d[f"decoder.fc{i}.bias"] = (o, k), (o,)
        _bn(d, f"decoder.bn{i}.", o)
    _tcn_stack(d, "decoder.tcn.", 4 * D, 64, 4)
    d["decoder.prob_decoder.loc_projection.weight"], d["decoder.prob_decoder.loc_projection.bias"] = (N * F, 64), (N * F,)
    _tail(d, kind, D, K)
    return d


def random_params(kind, encoder, N, E, F, Fe, D, K, graph, seed=0, scale=0.1):
    """A parameter dictionary with N(0, scale^2) weights, unit LayerNorm / BatchNorm scales, the graph operators of
    `graph` = (laplacian, edge_laplacian, incidence) and a uniform GMM prior: enough for timing and property tests."""
    shapes = {"transformer": transformer_shapes, "TCN": tcn_shapes}.get(encoder, recurrent_shapes)(kind, N, E, F, Fe, D, K)
    g = torch.Generator().manual_seed(seed)
    p = OrderedDict()
    for name, shape in shapes.items():
        if name.endswith("num_batches_tracked"):
            p[name] = torch.zeros((), dtype=torch.int64)
        elif name.endswith("running_mean"):
            p[name] = torch.zeros(shape)
        elif name.endswith("running_var") or ((".norm" in name or ".head.2." in name or ".head.5." in name or ".bn" in name) and name.endswith("weight") and len(shape) == 1):
            p[name] = torch.ones(shape)
        elif name == "latent_space.prior":
            p[name] = torch.full(shape, 1.0 / K)
        elif name == "vq_layer.codebook":
            p[name] = torch.rand(shape, generator=g)
        else:
            p[name] = torch.randn(shape, generator=g) * scale
    p["encoder.laplacian"], p["encoder.edge_laplacian"], p["encoder.incidence"] = (t.float() for t in graph)
    return p
