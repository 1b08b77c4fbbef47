"""GPU: the transformer encoder's eval forward (dof_tfm_encode through TFMEncoderB200) vs (1) golden vectors produced by
the UNMODIFIED reference (tests/golden/tfm_*.npz) and (2) the CPU oracle on a larger batch."""
import numpy as np
import pytest
import torch

from helpers import golden_cases_of, load_golden_of, sub, rel_l2

pytestmark = pytest.mark.gpu

TFM = golden_cases_of("tfm")


def _model(g, **kw):
    from deepof_b200 import TFMEncoderB200
    T, N, E, D, B, dk, heads, dff, layers = (int(v) for v in g["meta"])
    m = TFMEncoderB200((T, N, 3), (T, E, 1), g["adjacency"], D, num_layers=layers, num_heads=heads, dff=dff, **kw)
    assert m.key_dim == dk                                              # models_new.py:1014-1019
    assert [k for k in m.state_dict()] == [k[2:] for k in g if k.startswith("p/") and "num_batches_tracked" not in k]
    # the graph operators the library derives from the adjacency matrix == the reference's buffers
    for k in ("laplacian", "edge_laplacian", "incidence"):
        np.testing.assert_allclose(m.state_dict()[k].cpu().numpy(), g["p/" + k], rtol=0, atol=1e-6)
    m.load_state_dict(sub(g, "p/"))
    return m


@pytest.mark.parametrize("case", TFM)
def test_tfm_encoder_eval_vs_reference_golden(case):
    g = load_golden_of("tfm", case)
    m = _model(g)
    out, nodes, edges = m.encode(torch.from_numpy(g["x"]), torch.from_numpy(g["a"]), return_cores=True)
    errs = dict(nodes=rel_l2(nodes.cpu(), g["eval/nodes"]), edges=rel_l2(edges.cpu(), g["eval/edges"]), out=rel_l2(out.cpu(), g["eval/out"]))
    print(case, errs)
    assert errs["nodes"] < 2e-5 and errs["edges"] < 2e-5 and errs["out"] < 1e-4, errs
    assert float((out.cpu() - torch.from_numpy(g["eval/out"])).abs().max()) < 1e-4


def test_tfm_encoder_vs_oracle_chunked_batch():
    """B = 300 in chunks of 128: equals the oracle; chunking only changes which GEMM kernel the CensNet projection takes
    (tensor-core kernel from 2048 rows on), i.e. fp32 rounding."""
    from oracle import tfm_oracle as TO
    from oracle import vade_oracle as O
    g = load_golden_of("tfm", "cfg5")
    T, N, E, D = (int(v) for v in g["meta"][:4])
    x, a = O.synthetic_windows(300, T, g["adjacency"], seed=123)
    x[5, 7:] = 0.0
    a[5, 7:] = 0.0
    m = _model(g, max_batch=128)
    out = m(x, a)
    with torch.no_grad():
        ref = TO.encoder_forward_eval(x, a, sub(g, "p/"), O.graph_operators(g["adjacency"]), int(g["meta"][6]))
    assert rel_l2(out.cpu(), ref["out"]) < 1e-4
    m2 = _model(g, max_batch=300)
    assert rel_l2(m2(x, a).cpu(), out.cpu()) < 2e-6


def test_tfm_encoder_wide_key_dim_runs_layer_by_layer():
    """22 nodes -> key_dim 64: two layers do not fit one launch's shared memory, the core runs one layer per launch."""
    from deepof_b200 import TFMEncoderB200
    from oracle import tfm_oracle as TO
    from oracle import vade_oracle as O
    N, T, D, B = 22, 25, 8, 40
    adj = O.default_adjacency(N)
    E = int(np.count_nonzero(np.triu(adj)))
    m = TFMEncoderB200((T, N, 3), (T, E, 1), adj, D, seed=5)
    assert m.key_dim == 64
    x, a = O.synthetic_windows(B, T, adj, seed=9)
    out = m(x, a)
    p = {k: v.cpu() for k, v in m.state_dict().items()}
    with torch.no_grad():
        ref = TO.encoder_forward_eval(x, a, p, O.graph_operators(adj), 4)
    assert rel_l2(out.cpu(), ref["out"]) < 1e-4
    with pytest.raises(NotImplementedError):
        m.train()


TFMM = golden_cases_of("tfmmodel")


@pytest.mark.parametrize("case", TFMM)
def test_transformer_model_embeddings_from_reference_checkpoint(case, tmp_path):
    """A reference checkpoint of a transformer-encoder model (state_dict + rebuild_spec, model_utils_new.py:263-329)
    loads through load_model_from_ckpt and yields the embeddings / soft assignments embedding_per_video reads."""
    from deepof_b200 import load_model_from_ckpt
    g = load_golden_of("tfmmodel", case)
    T, N, E, D, K, B = (int(v) for v in g["meta"])
    name = str(g["model"])
    spec = dict(model_name=name, x_shape=(T, N, 3), a_shape=(T, E, 1), adjacency_matrix=g["adjacency"], latent_dim=D, n_components=K,
                encoder_type="transformer", use_gnn=True, kmeans_loss=0.0, interaction_regularization=0.0, lens_enabled=False)
    path = tmp_path / "ckpt.pth"
    torch.save({"state_dict": {k[2:]: torch.from_numpy(np.asarray(g[k])) for k in g if k.startswith("p/")}, "rebuild_spec": spec,
                "log_summary": {"ok": 1}}, path)
    m, summary = load_model_from_ckpt(str(path), max_batch=64)
    assert summary == {"ok": 1} and m.window_size == (T // 2 if name == "contrastive" else T)
    emb, q = m.embed(torch.from_numpy(g["x"]), torch.from_numpy(g["a"]))
    assert rel_l2(emb.cpu(), g["eval/emb"]) < 1e-4, rel_l2(emb.cpu(), g["eval/emb"])
    if name == "contrastive":
        assert q is None
    else:
        assert rel_l2(q.cpu(), g["eval/q"]) < 1e-4
        assert torch.equal(q.argmax(1).cpu(), torch.from_numpy(g["eval/q"]).argmax(1))
    # every tensor of the checkpoint survives a state_dict round trip, in the reference's key order and dtypes
    sd = m.state_dict()
    assert list(sd) == [k[2:] for k in g if k.startswith("p/")]
    for k in g:
        if k.startswith("p/"):
            assert sd[k[2:]].dtype == torch.from_numpy(np.asarray(g[k])).dtype, k
            assert torch.equal(sd[k[2:]].cpu(), torch.from_numpy(np.asarray(g[k]))), k
    # the reconstruction means through the same (trainable) model object
    if name == "vade":
        loc = m(torch.from_numpy(g["x"]), torch.from_numpy(g["a"]))[0]
        assert rel_l2(loc.cpu(), g["eval/loc"]) < 1e-4, rel_l2(loc.cpu(), g["eval/loc"])
    elif name == "vqvae":
        lq, le = m(torch.from_numpy(g["x"]), torch.from_numpy(g["a"]))[:2]
        assert rel_l2(le.cpu(), g["eval/loc"]) < 1e-4 and rel_l2(lq.cpu(), g["eval/loc_q"]) < 1e-4


def test_tfm_decoder_vs_oracle_large_batch():
    """B = 200 (5000 rows: the projections take the tensor-core GEMM kernel) vs the oracle."""
    from oracle import tfm_oracle as TO
    g = load_golden_of("tfmmodel", "vade")
    T, N, E, D, K, B = (int(v) for v in g["meta"])
    from deepof_b200 import TFMDecoderB200
    dec = TFMDecoderB200((T, N * 3), D, max_batch=128)
    dec.load_state_dict({k[2:]: g[k] for k in g if k.startswith("p/decoder.")})
    z = torch.randn(200, D, generator=torch.Generator().manual_seed(3))
    loc = dec(z)
    with torch.no_grad():
        ref, _ = TO.decoder_forward_eval(z, torch.ones(200, T, N * 3), sub(g, "p/"))
    assert rel_l2(loc.cpu(), ref) < 1e-4
