"""ctypes binding of libdeepof_b200.so (include/deepof_b200.h).

There is NO fallback: if the shared library is missing or fails to load, importing the
product API raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"``
(nvcc, sm_100a).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdeepof_b200.so")

DOF_N_LOGS = 16
LOG_KEYS = ("total_loss", "reconstruct_loss", "kl_div", "cat_clust_loss", "kmeans_loss", "activity_l1",
            "prior_loss", "distill_loss", "tf_clust_loss", "nonempty_loss", "temporal_loss", "scatter_loss",
            "repel_loss")  # order of step_vade's log dict (reference training.py:292-306)


class DofConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("T", "N", "E", "F", "Fe", "D", "K", "model", "encoder")]


MODEL_VADE, MODEL_VQVAE, MODEL_CONTRASTIVE = 0, 1, 2
ENCODER_RECURRENT, ENCODER_TRANSFORMER, ENCODER_TCN = 0, 1, 2
ENCODER_KINDS = {"recurrent": ENCODER_RECURRENT, "transformer": ENCODER_TRANSFORMER, "TCN": ENCODER_TCN}
ABI_VERSION = 3


class DofVadeLossCfg(C.Structure):
    _fields_ = [
        ("pretrain_mode", C.c_int), ("kl_weight", C.c_float), ("l1_activity_weight", C.c_float),
        ("kmeans_loss_weight", C.c_float), ("model_kmeans_weight", C.c_float),
        ("repel_weight", C.c_float), ("repel_length_scale", C.c_float),
        ("nonempty_weight", C.c_float), ("nonempty_floor", C.c_float), ("nonempty_p", C.c_int),
        ("tf_cluster_weight", C.c_float), ("reg_cat_clusters_weight", C.c_float),
        ("temporal_cohesion_weight", C.c_float), ("reg_scatter_weight", C.c_float),
        ("reg_scatter_beta", C.c_float), ("gmm_logvar_clamp_lo", C.c_float),
        ("gmm_logvar_clamp_hi", C.c_float), ("mc_samples", C.c_int), ("lambda_distill", C.c_float),
        ("distill_sharpen_T", C.c_float), ("distill_conf_weight", C.c_int),
        ("distill_conf_thresh", C.c_float),
    ]


class DofAdamCfg(C.Structure):
    _fields_ = [("lr", C.c_float * 4), ("step", C.c_int * 4), ("active", C.c_int * 4),
                ("weight_decay", C.c_float * 4), ("clip_value", C.c_float), ("grad_scale", C.c_float), ("beta1", C.c_float),
                ("beta2", C.c_float), ("eps", C.c_float)]


class DofLoaderCfg(C.Structure):
    _fields_ = [("T", C.c_int), ("step", C.c_int), ("N", C.c_int), ("E", C.c_int), ("center_node", C.c_int),
                ("align_node", C.c_int), ("cx", C.c_double), ("cy", C.c_double), ("fps", C.c_double),
                ("clip", C.c_double), ("coord_scale", C.c_double), ("coord_shift", C.c_double),
                ("speed_scale", C.POINTER(C.c_double)), ("speed_shift", C.POINTER(C.c_double)),
                ("dist_div", C.POINTER(C.c_double)), ("dist_scale", C.POINTER(C.c_double)),
                ("dist_shift", C.POINTER(C.c_double)), ("edges", C.POINTER(C.c_int))]


class DofViewsCfg(C.Structure):
    _fields_ = [("T_full", C.c_int), ("N", C.c_int), ("E", C.c_int), ("edges", C.POINTER(C.c_int)),
                ("start", C.c_void_p), ("n_rot", C.c_int), ("rot_pivot", C.c_int * 8), ("rot_mask", C.c_uint * 8),
                ("rot_theta", C.c_void_p), ("interp_t0", C.c_void_p), ("interp_len", C.c_void_p),
                ("noise", C.c_void_p)]


class DofDistillCfg(C.Structure):
    _fields_ = [("head", C.c_void_p), ("head_grad", C.c_void_p), ("tau_batch", C.c_void_p), ("K", C.c_int),
                ("lambda_", C.c_float), ("sharpen_T", C.c_float), ("conf_weight", C.c_int), ("conf_thresh", C.c_float)]


class DofTfmCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("T", "N", "E", "F", "Fe", "D", "key_dim", "heads", "dff", "layers")]


class DofTfmDecCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("T", "Dx", "D", "heads", "dff", "layers")]


class DofError(RuntimeError):
    pass


_lib = None

_P = C.c_void_p
_SIGS = {
    "dof_abi_version": (C.c_int, []),
    "dof_source_hash": (C.c_char_p, []),
    "dof_last_error": (C.c_char_p, []),
    "dof_state_numel": (C.c_int64, [C.POINTER(DofConfig)]),
    "dof_state_num_entries": (C.c_int, [C.POINTER(DofConfig)]),
    "dof_state_entry": (C.c_int, [C.POINTER(DofConfig), C.c_int, C.c_char_p, C.POINTER(C.c_int64),
                                  C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                  C.POINTER(C.c_int)]),
    "dof_graph_operators": (C.c_int, [C.POINTER(C.c_double), C.c_int, C.c_int, C.POINTER(C.c_float),
                                      C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "dof_workspace_bytes": (C.c_size_t, [C.POINTER(DofConfig), C.c_int, C.c_int]),
    "dof_create": (C.c_int, [C.POINTER(DofConfig), C.c_int, C.c_int, C.c_int, _P, C.c_size_t, C.POINTER(_P)]),
    "dof_destroy": (C.c_int, [_P]),
    "dof_vade_embed": (C.c_int, [_P, _P, _P, _P, C.c_int, _P, _P, _P]),
    "dof_vade_forward_eval": (C.c_int, [_P, _P, _P, _P, C.c_int, _P, _P, _P, _P, _P]),
    "dof_vade_loss_grad": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, _P, _P, _P, _P, _P,
                                     C.POINTER(DofVadeLossCfg), _P, _P]),
    "dof_vade_loss_eval": (C.c_int, [_P, _P, _P, _P, C.c_int, _P, _P, _P, _P, C.POINTER(DofVadeLossCfg), _P, _P]),
    "dof_vqvae_loss_eval": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_float, C.c_float, _P, _P]),
    "dof_contrastive_loss_eval": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, _P, _P, _P]),
    "dof_clip_adam": (C.c_int, [_P, _P, _P, _P, _P, C.POINTER(DofAdamCfg), _P]),
    "dof_encode": (C.c_int, [_P, _P, _P, _P, C.c_int, _P, _P]),
    "dof_dropout_mask_bytes": (C.c_size_t, [C.POINTER(DofConfig), C.c_int, C.c_int, C.c_int]),
    "dof_set_noise_seed": (C.c_int, [_P, C.c_ulonglong]),
    "dof_set_dropout": (C.c_int, [_P, C.c_ulonglong, _P, C.c_size_t]),
    "dof_vqvae_forward_eval": (C.c_int, [_P, _P, _P, _P, C.c_int, _P, _P, _P, _P, _P, _P, _P]),
    "dof_vqvae_loss_grad": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_float, C.c_float, _P, _P]),
    "dof_contrastive_views": (C.c_int, [C.POINTER(DofViewsCfg), _P, C.c_int, _P, _P, _P]),
    "dof_vqvae_loss_grad_distill": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_float, C.c_float, C.POINTER(DofDistillCfg), _P, _P]),
    "dof_contrastive_loss_grad_distill": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                                    C.POINTER(DofDistillCfg), _P, _P, _P]),
    "dof_tfm_numel": (C.c_int64, [C.POINTER(DofTfmCfg)]),
    "dof_tfm_num_entries": (C.c_int, [C.POINTER(DofTfmCfg)]),
    "dof_tfm_entry": (C.c_int, [C.POINTER(DofTfmCfg), C.c_int, C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "dof_tfm_workspace_bytes": (C.c_size_t, [C.POINTER(DofTfmCfg), C.c_int]),
    "dof_tfm_encode": (C.c_int, [C.POINTER(DofTfmCfg), _P, _P, _P, C.c_int, _P, C.c_size_t, _P, _P, _P, _P]),
    "dof_latent_eval": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P]),
    "dof_vq_eval": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "dof_tfm_dec_numel": (C.c_int64, [C.POINTER(DofTfmDecCfg)]),
    "dof_tfm_dec_num_entries": (C.c_int, [C.POINTER(DofTfmDecCfg)]),
    "dof_tfm_dec_entry": (C.c_int, [C.POINTER(DofTfmDecCfg), C.c_int, C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                    C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "dof_tfm_dec_workspace_bytes": (C.c_size_t, [C.POINTER(DofTfmDecCfg), C.c_int]),
    "dof_tfm_decode": (C.c_int, [C.POINTER(DofTfmDecCfg), _P, _P, C.c_int, _P, C.c_size_t, _P, _P]),
    "dof_adam_flat": (C.c_int, [_P, _P, _P, _P, C.c_longlong, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int,
                                C.c_float, _P]),
    "dof_contrastive_loss_grad": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, _P, _P,
                                            _P]),
    "dof_loader_num_windows": (C.c_longlong, [C.c_longlong, C.c_int, C.c_int]),
    "dof_load_windows": (C.c_int, [C.POINTER(DofLoaderCfg), _P, C.c_longlong, C.c_longlong, C.c_int, _P, _P, _P]),
    "dof_loader_pair_length": (C.c_int, [_P, C.c_longlong, C.c_int, C.c_int, C.c_int, _P, _P]),
    "dof_loader_moments": (C.c_int, [C.POINTER(DofLoaderCfg), _P, C.c_longlong, C.POINTER(C.c_double), _P, _P]),
    "dof_debug_tensor": (_P, [_P, C.c_char_p, C.POINTER(C.c_int64)]),
    "dof_launch_count": (C.c_longlong, []),
    "dof_set_tensor_cores": (C.c_int, [C.c_int]),
    "dof_set_concurrency": (C.c_int, [C.c_int]),
    "dof_profile_begin": (C.c_int, []),
    "dof_profile_end": (C.c_int, [C.c_char_p, C.c_size_t]),
    "dof_test_gemm_rows": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.c_int, _P,
                                     _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "dof_test_gemm_wgrad": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.c_int, C.c_int,
                                      C.c_int, _P, C.c_int, C.c_int, _P, C.c_int, C.c_int, C.c_int, _P]),
    "dof_test_gru_fwd": (C.c_int, [_P, _P, C.c_longlong, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int,
                                   C.c_int, C.c_int, _P]),
    "dof_test_gru_layer_fwd": (C.c_int, [_P, C.c_longlong, C.c_int, C.POINTER(C.c_void_p), _P, _P, _P, _P, _P, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_int, _P]),
    "dof_test_gru_layer_bwd": (C.c_int, [C.POINTER(C.c_void_p), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int,
                                         C.c_int, C.c_int, _P]),
    "dof_test_gru_layer_bwdw": (C.c_int, [_P, C.POINTER(C.c_void_p), _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int,
                                          C.c_int, C.c_int, _P]),
    "dof_peer_barrier": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, _P]),
    "dof_peer_reduce": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, _P, C.c_longlong, _P]),
    "dof_test_gru_bwdw_timeline": (C.c_int, [_P]),
    "dof_test_gru_wgrad": (C.c_int, [_P, _P, _P, C.c_int, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "dof_test_tfm_attention": (C.c_int, [_P, _P, _P, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P]),
    "dof_test_encoder_grad": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, _P, _P, _P]),
    "dof_test_gru_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "dof_test_layernorm": (C.c_int, [_P, _P, _P, C.c_float, _P, _P, _P, _P, _P, _P, _P, C.c_longlong, C.c_int,
                                     C.c_int, _P]),
    "dof_test_tcn_conv": (C.c_int, [C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_int, C.c_longlong, _P, _P, _P, _P]),
}
EXPORTS = tuple(_SIGS.keys())


def lib():
    """Load (once) and return the shared library; raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DofError(
                f"{LIB_PATH} not found: the CUDA library is not built (run __graft_entry__.build()); "
                "deepof_b200 has no CPU fallback")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise DofError(f"deepof_b200 error {rc}: {lib().dof_last_error().decode()}")


def ptr(t):
    """Device (or host) pointer of a torch tensor / None."""
    if t is None:
        return None
    assert t.is_contiguous(), "deepof_b200 needs contiguous tensors"
    return C.c_void_p(t.data_ptr())
