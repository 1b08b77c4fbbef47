"""deepof_b200 — B200-native (sm_100a) trainer for DeepOF's VaDE pose-window embeddings.

Only the hot path lives here: hand-written CUDA kernels behind a C-ABI
(``include/deepof_b200.h``, ``libdeepof_b200.so``) and a thin Python host that mirrors the
reference's model / step interface.  There is no CPU fallback.
"""
from ._lib import DofError, LIB_PATH, LOG_KEYS  # noqa: F401
from .vade import VaDEB200, VadeLossCfg, graph_operators, state_layout  # noqa: F401
from .tfm import TFMDecoderB200, TFMEncoderB200, TFMModelB200  # noqa: F401
from .models import (VQVAEB200, ContrastiveB200, ContrastiveAugCfg, AugParams, RotationTable, DistillHeadB200,  # noqa: F401
                     Distillation)
from .inference import embedding_per_video  # noqa: F401
from .api import (StepResult, step_vade, step_vqvae_distill, step_contrastive_distill, train_one_epoch_indexed,  # noqa: F401
                  train_deepof_model, save_model_info, load_model_from_ckpt)
from .teacher import (initialize_gmm_from_teacher, gmm_moments_from_teacher, TurtleTeacherB200,  # noqa: F401
                      run_turtle_teacher_on_views, IncrementalPCAB200, teacher_views_from_windows, build_turtle_teacher)
from .loader import WindowLoader, GlobalScalers, VideoConstants, batch_starts, reference_divisors  # noqa: F401

__all__ = ["VaDEB200", "VadeLossCfg", "DofError", "graph_operators", "state_layout", "LIB_PATH", "LOG_KEYS",
           "TFMEncoderB200", "TFMDecoderB200", "TFMModelB200", "VQVAEB200", "ContrastiveB200", "DistillHeadB200", "Distillation", "ContrastiveAugCfg", "AugParams", "RotationTable", "embedding_per_video", "StepResult", "step_vade", "step_vqvae_distill", "step_contrastive_distill",
           "train_one_epoch_indexed", "train_deepof_model", "save_model_info", "load_model_from_ckpt", "WindowLoader", "GlobalScalers", "VideoConstants", "batch_starts", "reference_divisors",
           "initialize_gmm_from_teacher", "gmm_moments_from_teacher", "TurtleTeacherB200", "run_turtle_teacher_on_views",
           "IncrementalPCAB200", "teacher_views_from_windows", "build_turtle_teacher"]
