"""CPU: the transformer-encoder oracle (oracle/tfm_oracle.py, eval mode) vs golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden_tfm.py)."""
import pytest
import torch

from oracle import tfm_oracle as TO
from oracle import vade_oracle as O
from helpers import golden_cases_of, load_golden_of, sub, rel_l2

TFM = golden_cases_of("tfm")


def test_goldens_present():
    assert len(TFM) >= 3


@pytest.mark.parametrize("case", TFM)
def test_tfm_encoder_eval(case):
    g = load_golden_of("tfm", case)
    p = sub(g, "p/")
    x, a = torch.from_numpy(g["x"]), torch.from_numpy(g["a"])
    graph = O.graph_operators(g["adjacency"])
    heads = int(g["meta"][6])
    with torch.no_grad():
        out = TO.encoder_forward_eval(x, a, p, graph, heads)
    B, N, E = x.shape[0], x.shape[2], a.shape[2]
    assert rel_l2(out["nodes"].reshape(B * N, -1), g["eval/nodes"]) < 5e-6
    assert rel_l2(out["edges"].reshape(B * E, -1), g["eval/edges"]) < 5e-6
    assert rel_l2(out["out"], g["eval/out"]) < 1e-5
    if case == "padded":      # the key-padding mask is really exercised
        xs = O.group_reshape(x).reshape(B * N, x.shape[1], 3)
        assert bool((xs == 0).all(-1).any())


@pytest.mark.parametrize("case", [c for c in golden_cases_of("tfmmodel") if c != "contrastive"])
def test_tfm_decoder_eval(case):
    """TFMDecoderPT in eval mode on the latent the reference fed it (VaDE: z_mean, VQ-VAE: encoder output / codes)."""
    g = load_golden_of("tfmmodel", case)
    p = sub(g, "p/")
    x = torch.from_numpy(g["x"])
    B, T, N, F = x.shape
    xf = x.reshape(B, T, N * F)
    with torch.no_grad():
        loc, mask = TO.decoder_forward_eval(torch.from_numpy(g["eval/emb"]), xf, p)
        assert rel_l2(loc, g["eval/loc"]) < 1e-5
        if "eval/loc_q" in g:
            loc_q, _ = TO.decoder_forward_eval(torch.from_numpy(g["eval/quant"]), xf, p)
            assert rel_l2(loc_q, g["eval/loc_q"]) < 1e-5
    assert bool(mask.all())
