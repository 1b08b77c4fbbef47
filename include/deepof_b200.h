/* deepof_b200 — C ABI of the B200-native pose-window embedding trainer.
 *
 * This is the drop-in boundary for the ONE hot path of mlfpm/deepof that this library
 * replaces: forward / backward / optimizer step of the VaDE model with the recurrent
 * (GRU + CensNet) encoder-decoder over sliding windows of keypoint trajectories, plus the
 * eval-mode embedding used by inference.  The reference has no native boundary (its
 * "operator API" is Python); each entry point cites the reference
 * code it stands in for (paths relative to the reference repository root).
 *
 * Conventions
 *  - plain C: pointers and sizes only, no torch / C++ types.
 *  - every `float*` / `int*` data argument is a DEVICE pointer to a contiguous row-major
 *    array that the CALLER owns (the Python host allocates them as torch tensors); the
 *    library owns nothing but the handle and carves its activations out of the
 *    caller-provided workspace.
 *  - all work is enqueued on the `stream` argument (a cudaStream_t passed as void*); no
 *    call synchronises the device.
 *  - returns 0 on success, a negative DOF_ERR_* code otherwise; dof_last_error() gives
 *    the message (thread-local).  Nothing throws across the boundary.
 *  - one handle per (process, device); not thread-safe.
 *  - there is no CPU fallback: every entry point fails if no sm_100-class device is usable.
 */
#ifndef DEEPOF_B200_H
#define DEEPOF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DOF_ABI_VERSION 3

/* Model geometry.  Mirrors the constructor arguments of VaDEPT / RecurrentEncoderPT /
 * RecurrentDecoderPT (deepof/clustering/models_new.py:37-105, 281-324, 1794-1839). */
typedef struct {
    int T;   /* window length (time steps)                          */
    int N;   /* graph nodes (body parts)                            */
    int E;   /* graph edges                                         */
    int F;   /* features per node  (x, y, speed) = 3                */
    int Fe;  /* features per edge  (log1p length) = 1               */
    int D;   /* latent_dim                                          */
    int K;   /* n_components (GMM clusters / VQ codebook size)      */
    int model; /* DOF_MODEL_*: which reference model the state buffer lays out */
    int encoder; /* DOF_ENCODER_*: encoder_type of the reference model builders (models_new.py:1432-1508) */
} dof_config;

/* Encoder / decoder family (the reference's `encoder_type`):
 *  RECURRENT    RecurrentEncoderPT + RecurrentDecoderPT (models_new.py:37-373): Conv1d + 2 BiGRU + LayerNorm per node / edge
 *  TRANSFORMER  TFMEncoderPT + TFMDecoderPT (models_new.py:985-1327): per node / edge TransformerCorePT (key_dim =
 *               min(64, N*F) rounded down to a multiple of 4 heads, dff 128, 2 post-LN layers, dropout 0.1), CensNet, RMS
 *               normalisation, head MLP with BatchNorm (eps 1e-3, momentum 0.01), train-mode batch standardisation;
 *               decoder: latent expansion MLP (GELU) to 4*D, 2 pre-LN causal self-attention layers (8 heads, dff 128,
 *               dropout 0.2).  The state layout follows the reference state_dict of that model; the integer
 *               num_batches_tracked buffers are carried as one float each. */
/*  TCN          TCNEncoderPT + TCNDecoderPT (models_new.py:376-657, 713-819): per node / edge TCN1DPT (2 stacks of dilations
 *               1,2,4,8; 32 filters, kernel 4, causal padding, skip connections, train-mode BatchNorm eps 1e-3 momentum 0.1,
 *               no dropout), last step -> CensNet -> the same RMS normalisation and head MLP as the transformer encoder (no
 *               batch standardisation); decoder: RMS normalisation, three Dense + BatchNorm (momentum 0.01) layers to 4*D,
 *               repeated over the window, TCN1DPT (dilations 8,4,2,1; 64 filters) and the probabilistic head.  The dilated
 *               causal convolutions run as tcgen05 GEMMs over time-shifted views; BatchNorm running buffers are updated by
 *               dof_clip_adam from the batch statistics of every encoder / decoder pass of the preceding step. */
#define DOF_ENCODER_RECURRENT 0
#define DOF_ENCODER_TRANSFORMER 1
#define DOF_ENCODER_TCN 2

/* Model kinds (all with encoder_type="recurrent", use_gnn=True):
 *  VADE         VaDEPT        deepof/clustering/models_new.py:1794-1976  encoder + decoder + latent_space
 *  VQVAE        VQVAEPT       models_new.py:1510-1635                    encoder + decoder + vq_layer.codebook [D,K]
 *  CONTRASTIVE  ContrastivePT models_new.py:1978-2069                    encoder only; T is the HALF window the
 *                                                                         encoder sees (window_size = T_full // 2) */
#define DOF_MODEL_VADE 0
#define DOF_MODEL_VQVAE 1
#define DOF_MODEL_CONTRASTIVE 2

/* One phase of VadeLoss (deepof/clustering/losses.py:383-457, 567-797).  Field names follow
 * the reference attributes.  tf_cluster_weight and reg_scatter_weight must be 0 (their
 * reference defaults); a non-zero value is rejected with DOF_ERR_UNSUPPORTED. */
typedef struct {
    int pretrain_mode;
    float kl_weight;               /* Dynamic_weight_manager.get_weight() for this step */
    float l1_activity_weight;
    float kmeans_loss_weight;      /* VadeLoss.kmeans_loss_weight of the active mode    */
    float model_kmeans_weight;     /* GaussianMixtureLatentPT.kmeans_weight             */
    float repel_weight, repel_length_scale;
    float nonempty_weight, nonempty_floor;
    int nonempty_p;
    float tf_cluster_weight;
    float reg_cat_clusters_weight;
    float temporal_cohesion_weight;
    float reg_scatter_weight, reg_scatter_beta;
    float gmm_logvar_clamp_lo, gmm_logvar_clamp_hi;
    int mc_samples;                /* must be 32 (losses.py:526)                        */
    float lambda_distill, distill_sharpen_T;
    int distill_conf_weight;
    float distill_conf_thresh;
} dof_vade_loss_cfg;

/* Per-step optimizer scalars: clip_grad_value_ + Adam with the parameter groups of
 * build_optimizer_vade (deepof/clustering/losses.py:817-833; training.py:164-166).
 * Groups: 1 = encoder + latent heads, 2 = decoder, 3 = GMM means/log-vars.  `step[g]` is the
 * 1-based Adam step count of the group AFTER this update; `active[g]`=0 leaves the group
 * untouched (requires_grad=False in the reference, training.py:1746-1767). */
typedef struct {
    float lr[4];
    int step[4];
    int active[4];
    float weight_decay[4]; /* torch.optim.Adam L2 term per group: 0 for build_optimizer_vade, 1e-4 for
                            * build_optimizer_generic (losses.py:805-814), added after the clip     */
    float clip_value;   /* 0.75 in the reference; <=0 disables clipping              */
    float grad_scale;   /* 1/world_size after a summed all-reduce, else 1             */
    float beta1, beta2, eps;
} dof_adam_cfg;

#define DOF_N_LOGS 16
/* logs[] (device, DOF_N_LOGS floats) in the order of step_vade's log dict
 * (deepof/clustering/training.py:292-306):
 *  0 total_loss 1 reconstruct_loss 2 kl_div 3 cat_clust_loss 4 kmeans_loss 5 activity_l1
 *  6 prior_loss 7 distill_loss 8 tf_clust_loss 9 nonempty_loss 10 temporal_loss
 *  11 scatter_loss 12 repel_loss 13 kl_weight 14 raw MC-KL mean (before the clamp) */

typedef struct dof_handle dof_handle;

int dof_abi_version(void);
/* first 16 hex digits of the sha256 over csrc/ and this header the library was built from (build provenance) */
const char* dof_source_hash(void);
const char* dof_last_error(void);

/* ---- flat state buffer ----------------------------------------------------------------
 * All parameters and buffers of the model live in ONE flat fp32 buffer whose segments are,
 * in order, the entries of the reference VaDEPT.state_dict(), so a
 * reference checkpoint maps onto it by name and vice versa (model_utils_new.py:263-329). */
int64_t dof_state_numel(const dof_config* cfg);
int dof_state_num_entries(const dof_config* cfg);
/* name_out: caller buffer of >=128 bytes; shape_out: 4 ints (unused dims = 0);
 * group_out: 0 buffer / never-trained parameter, 1..3 optimizer group. */
int dof_state_entry(const dof_config* cfg, int index, char* name_out, int64_t* offset_out,
                    int64_t* numel_out, int* ndim_out, int* shape_out, int* group_out);

/* Graph operators of CensNetConvPT.preprocess (deepof/clustering/censNetConv_pt.py:160-175):
 * HOST arrays. adjacency [N*N] -> laplacian [N*N], edge_laplacian [E*E], incidence [N*E].
 * Returns E through n_edges_out (edges = non-zeros of triu(adjacency), row-major). */
int dof_graph_operators(const double* adjacency, int N, int max_edges, float* laplacian,
                        float* edge_laplacian, float* incidence, int* n_edges_out);

/* ---- handle ---------------------------------------------------------------------------- */
size_t dof_workspace_bytes(const dof_config* cfg, int max_batch, int training);
int dof_create(const dof_config* cfg, int device, int max_batch, int training, void* workspace,
               size_t workspace_bytes, dof_handle** out);
int dof_destroy(dof_handle* h);

/* ---- eval-mode embedding: what embedding_per_video reads as model(x,a)[1], [2]
 * (deepof/clustering/model_utils_new.py:610-617; VaDEPT.forward models_new.py:1841-1891).
 * x [B,T,N,F], a [B,T,E,Fe] -> emb [B,D] (= z_mean), q [B,K]. */
int dof_vade_embed(dof_handle* h, const float* state, const float* x, const float* a, int B,
                   float* emb, float* q, void* stream);

/* Full eval forward incl. decoder mean (tests / reconstruction): also enc [B,D] and
 * loc [B,T,N*F]; any output pointer may be NULL. */
int dof_vade_forward_eval(dof_handle* h, const float* state, const float* x, const float* a, int B,
                          float* enc, float* emb, float* q, float* loc, void* stream);

/* ---- one training step, split at the gradient all-reduce
 * dof_vade_loss_grad  = step_vade forward + criterion + loss.backward()
 *                       (deepof/clustering/training.py:231-309, 159-163)
 * dof_clip_adam       = clip_grad_value_(0.75) + optimizer.step()  (training.py:164-166)
 * grad is overwritten (zeroed first).  eps [B,D] is the reparameterisation noise, mc_eps
 * [32,B,D] the Monte-Carlo KL noise (main mode only).  Either may be NULL: the kernels then draw
 * it themselves from a counter-based Philox4x32-10 stream (standard normal by Box-Muller) keyed by
 * dof_set_noise_seed — nothing is materialised in HBM; pass explicit tensors to reproduce a
 * reference run (parity tests).
 * tau_batch [B,K] = tau_star[batch_indices] or NULL, class_weight [K] or NULL,
 * floor_c [K] = per-cluster non-empty floor (losses.py:672-680). */
int dof_vade_loss_grad(dof_handle* h, const float* state, float* grad, const float* x, const float* a,
                       int B, const float* eps, const float* mc_eps, const float* tau_batch,
                       const float* class_weight, const float* floor_c, const dof_vade_loss_cfg* loss,
                       float* logs, void* stream);
/* The step validate_one_epoch_indexed runs (deepof/clustering/training.py:190-229): the model in eval() — z = z_mean, no
 * dropout, BatchNorm running statistics, no batch standardisation — and the criterion's terms; logs as above, no
 * gradient, parameters and optimizer state untouched.  The handle must have been created with training=1 (the loss
 * workspace). */
int dof_vade_loss_eval(dof_handle* h, const float* state, const float* x, const float* a, int B, const float* mc_eps,
                       const float* tau_batch, const float* class_weight, const float* floor_c, const dof_vade_loss_cfg* loss,
                       float* logs, void* stream);
int dof_clip_adam(dof_handle* h, float* state, const float* grad, float* adam_m, float* adam_v,
                  const dof_adam_cfg* opt, void* stream);

/* ---- window loader: raw pose frames -> x [B,T,N,3], a [B,T,E,1]   -------------
 * Replaces, for one video, the reference's CPU chain between a pose table and its window store:
 * centre (deepof/data.py:1844-1869), align (data.py:1878-1928 -> deepof/utils.py:2097-2142, 1298-1319),
 * rolling_speed (utils.py:3788-3857), edge lengths (utils.py:863-881), scale_table (utils.py:2425-2566),
 * _pp_apply_global (utils.py:2866-2921), clip + interpolate + sanitize (utils.py:2990-3004, 2577-2583),
 * rolling_window (utils.py:3354-3377) and reorder_and_reshape (deepof/clustering/dataset.py:16-26).
 * Every scaler is affine, so the host folds them into one (scale, shift) per column:
 *   coords  z = r * coord_scale + coord_shift          r = centred, aligned coordinate
 *   speeds  z = v * speed_scale[n] + speed_shift[n]    v = round(mean3(|p_t - p_{t-2}| / 2), 3) * fps
 *   edges   z = log1p(max(d / dist_div[e], 0)) * dist_scale[e] + dist_shift[e]
 * then |z| > clip -> linear interpolation along the frame axis (nearest valid frames on both sides,
 * edge-filled at the ends, 0 if a column has no valid frame).  The per-column arrays and `edges` are
 * HOST arrays (copied into the launch); `frames` [n_frames,N,2], x, a are DEVICE arrays.
 * Window w covers frames w*step .. w*step + T - 1. */
typedef struct {
    int T, step, N, E;
    int center_node;           /* body part to centre on, or -1: arena centre (cx, cy) */
    int align_node;            /* body part rotated onto +y, or -1: no alignment       */
    double cx, cy, fps;
    double clip;               /* 10 in the reference (interpolate_normalized); <= 0 disables */
    double coord_scale, coord_shift;
    const double* speed_scale; const double* speed_shift;                       /* [N] */
    const double* dist_div; const double* dist_scale; const double* dist_shift; /* [E] */
    const int* edges;                                                           /* [E,2] */
} dof_loader_cfg;

long long dof_loader_num_windows(long long n_frames, int T, int step);
int dof_load_windows(const dof_loader_cfg* cfg, const float* frames, long long n_frames, long long first_window,
                     int count, float* x, float* a, void* stream);
/* out[f] = |p_a(f) - p_b(f)| (fp64, DEVICE [n_frames]); the size factor of scale_table is its nanmedian
 * (utils.py:2478-2489). */
int dof_loader_pair_length(const float* frames, long long n_frames, int N, int node_a, int node_b, double* out,
                           void* stream);
/* Moments of the columns as configured by cfg (clip is honoured; pass clip <= 0 for scaler fits), NaNs
 * skipped, per group g in {coords, speeds, edges}: out9[3g] += count, out9[3g+1] += sum(z - shift3[g]),
 * out9[3g+2] += sum((z - shift3[g])^2).  out9 (DEVICE, fp64) is accumulated into, not cleared: this is
 * what the groupwise StandardScaler fits of scale_table / _pp_fit_global_scaler see. */
int dof_loader_moments(const dof_loader_cfg* cfg, const float* frames, long long n_frames, const double* shift3,
                       double* out9, void* stream);

/* ---- encoder only: model.encoder(x, a) -> enc [B,D] (RecurrentEncoderPT.forward, models_new.py:140-181);
 * this is ContrastivePT.forward (models_new.py:2063-2069) and VQVAEPT.encode (:1637-1640).  Any model kind. */
int dof_encode(dof_handle* h, const float* state, const float* x, const float* a, int B, float* enc, void* stream);

/* ---- dropout and BatchNorm state of the transformer family (DOF_ENCODER_TRANSFORMER) ------------------------------
 * Every dropout decision of a training step (nn.Dropout / scaled_dot_product_attention(dropout_p), models_new.py:880-884,
 * 914-917, 972, 1311-1325) comes from a counter-based Philox4x32-10 stream keyed by `seed` (one counter block per
 * dropout site and element, so forward and backward regenerate the same decisions and nothing is stored), or — for
 * parity runs against the reference — from explicit keep masks (`masks` != NULL, DEVICE bytes, 1 = kept), concatenated in
 * the order the reference draws them for ONE encoder pass over the Bw windows handed to the step and `dec_passes`
 * decoder passes over B windows (VaDE 1, VQ-VAE 2, contrastive 0 with Bw = 2B: rows 0..B-1 the main view):
 *   for core in (node, edge), S = Bw * (N or E):  embed [S,T,key_dim];  per layer: attention weights [S,4,T,T],
 *       dropout1 [S,T,key_dim], dropout2 [S,T,key_dim]
 *   per decoder pass, per layer: attention weights [B,8,T,T], out-projection dropout [B,T,4D], FFN hidden [B,T,128],
 *       FFN output [B,T,4D].
 * dof_dropout_mask_bytes returns the total for Bw encoder windows and dec_passes passes over B decoder windows.  The
 * setting persists until the next call; the default is Philox with seed 0 — callers advance the seed every step.
 * BatchNorm running statistics (BatchNorm1dKerasFP32, models_new.py:508-516) are updated by dof_clip_adam from the
 * batch statistics of the preceding *_loss_grad call (rank-local, as under the reference's DDP with
 * broadcast_buffers=False). */
/* Philox key of the in-kernel VaDE noise (eps, MC-KL samples) used when the noise pointers are NULL; callers advance
 * it every step.  0 selects the default key. */
int dof_set_noise_seed(dof_handle* h, unsigned long long seed);
size_t dof_dropout_mask_bytes(const dof_config* cfg, int Bw, int B, int dec_passes);
int dof_set_dropout(dof_handle* h, unsigned long long seed, const unsigned char* masks, size_t mask_bytes);

/* ---- VQ-VAE (DOF_MODEL_VQVAE) ------------------------------------------------------------------------
 * Eval forward of VQVAEPT(x, a, return_all_outputs=True) (models_new.py:1575-1635): enc [B,D] encoder output,
 * quant [B,D] quantized latents, soft [B,K] soft counts, idx [B] code indices (VectorQuantizerPT.get_code_indices,
 * :1406-1423), loc_q / loc_e [B,T,N*F] decoder means from the quantized / the encoder latents.  Outputs may be NULL. */
int dof_vqvae_forward_eval(dof_handle* h, const float* state, const float* x, const float* a, int B, float* enc,
                           float* quant, float* soft, int* idx, float* loc_q, float* loc_e, void* stream);
/* step_vqvae_distill forward + loss.backward() with the teacher off (deepof/clustering/training.py:312-389, 159-163).
 * grad is overwritten.  logs (DEVICE, DOF_N_LOGS floats): 0 total_loss 1 enc_rec_loss 2 reconstruct_loss 3 vq_loss
 * 4 kmeans_loss 5 number_of_populated_clusters 6 distill_loss (= 0). */
int dof_vqvae_loss_grad(dof_handle* h, const float* state, float* grad, const float* x, const float* a, int B,
                        float beta, float kmeans_weight, float* logs, void* stream);

/* validation step of fit_VQVAE (training.py:1165-1170): eval-mode forward + loss terms, teacher off, no gradient */
int dof_vqvae_loss_eval(dof_handle* h, const float* state, const float* x, const float* a, int B, float beta, float kmeans_weight,
                        float* logs, void* stream);

/* ---- distillation head of step_vqvae_distill / step_contrastive_distill -------------------------------
 * (deepof/clustering/training.py:341-370, 550-578): DiscriminativeHead (teacher_model.py:795-808) = one Linear(D, K) on
 * the encoder output (VQ-VAE: z_e; contrastive: the row-normalised z of the MAIN view, training.py:533, 556), soft cross-entropy (_soft_ce_logits,
 * training.py:392-400) against tau_batch [B,K] = tau_star[batch_indices], sharpened with temperature sharpen_T when
 * > 0 and optionally confidence-weighted; loss += lambda * mean_b.  head = W [K,D] | b [K] is CALLER-owned (it is not
 * part of the model's state_dict in the reference either); head_grad has the same layout and is overwritten.
 * lambda <= 0 or distill == NULL is the teacher-off step.  logs slot distill_loss receives lambda * mean_b. */
typedef struct dof_distill_cfg {
    const float* head;
    float* head_grad;
    const float* tau_batch;
    int K;
    float lambda;
    float sharpen_T;
    int conf_weight;
    float conf_thresh;
} dof_distill_cfg;
int dof_vqvae_loss_grad_distill(dof_handle* h, const float* state, float* grad, const float* x, const float* a, int B,
                                float beta, float kmeans_weight, const dof_distill_cfg* distill, float* logs,
                                void* stream);
/* Adam on a caller-owned flat buffer (the head): torch.optim.Adam(lr, weight_decay) of build_optimizer_generic
 * (losses.py:805-814) WITHOUT clip_grad_value_ (training.py:165 clips model.parameters() only); step >= 1 is the
 * bias-correction count, grad_scale = 1 / world after a sum all-reduce. */
int dof_adam_flat(float* param, const float* grad, float* adam_m, float* adam_v, long long n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream);

/* ---- data-parallel gradient exchange over NVLink peer memory (csrc/peer.cuh) -------------------------
 * Replaces the per-step gradient all-reduce DistributedDataParallel runs for the reference (deepof/clustering/
 * training.py:1567 DDP wrap, :164 loss.backward()).  Every rank owns one symmetric buffer [gradient | signal pad] that
 * is mapped into all peers; `pads` / `peers` are HOST arrays of `world` DEVICE pointers as seen from THIS process
 * (pads[r] = rank r's signal pad, >= world int32, zero-initialised; peers[r] = rank r's gradient, 16-byte aligned).
 * dof_peer_barrier: rank-to-rank barrier on the stream; `epoch` must increase by one with every call on all ranks.
 * dof_peer_reduce:  out[i] = sum over r = 0..world-1 (in that order: bit-identical on every rank) of peers[r][i];
 *                   n floats, multiple of 4.  Call order per step: barrier, reduce, barrier. */
int dof_peer_barrier(const void* const* pads, int world, int rank, int epoch, void* stream);
int dof_peer_reduce(const void* const* peers, int world, float* out, long long n, void* stream);

/* ---- contrastive (DOF_MODEL_CONTRASTIVE) --------------------------------------------------------------
 * dof_contrastive_views builds what step_contrastive_distill feeds its encoder (training.py:497-525): from
 * x_full [B,T_full,N,3] the middle half window and the augmented view of _make_augmented_view (training.py:2128-2402)
 * with edge lengths recomputed from the coordinates (recompute_edges, model_utils_new.py:332-364):
 *   x2 [2B,T_full/2,N,3], a2 [2B,T_full/2,E,1]; rows 0..B-1 the main view, B..2B-1 the augmented view.
 * The random decisions of the augmentation are INPUTS (drawn by the host RNG): start [B] slice start of the
 * augmented view; n_rot rotations applied in order, rotation k turning the nodes of bit mask rot_mask[k] around node
 * rot_pivot[k] by rot_theta[k*B + b] radians; interp_t0 / interp_len [B]: frames t0 .. t0+len-1 of the half window are
 * replaced by the linear interpolation between frames t0-1 and t0+len (len 0 = off; both NULL = off); noise [B,N,3]
 * additive per-node offsets (NULL = off).  start, rot_theta, interp_*, noise are DEVICE arrays; edges is HOST. */
typedef struct {
    int T_full, N, E;
    const int* edges;          /* HOST [E,2] */
    const int* start;          /* DEVICE [B] */
    int n_rot;
    int rot_pivot[8];
    unsigned int rot_mask[8];
    const float* rot_theta;    /* DEVICE [n_rot,B] */
    const int* interp_t0;      /* DEVICE [B] or NULL */
    const int* interp_len;     /* DEVICE [B] or NULL */
    const float* noise;        /* DEVICE [B,N,3] or NULL */
} dof_views_cfg;
int dof_contrastive_views(const dof_views_cfg* v, const float* x_full, int B, float* x2, float* a2, void* stream);
/* step_contrastive_distill forward + loss.backward() with the teacher off (training.py:527-545, 159-163):
 * z = encoder(main view), z_aug = encoder(augmented view) (one pass over the 2B windows: the handle needs
 * max_batch >= 2B), F.normalize, cosine similarity / temperature, cross-entropy against the diagonal
 * (loss_kind 0: nce_loss_pt, deepof/clustering/losses.py:130-141), or the debiased losses on the same similarities:
 * loss_kind 1 dcl_loss_pt (losses.py:144-173, tau_plus), 2 hard_loss_pt (losses.py:213-249, tau_plus, beta),
 * 3 fc_loss_pt (losses.py:176-210, the ceil(0.1 B) largest negatives of every row eliminated);
 * sim_kind 0 cosine / dot (losses.py:59-67), 1 euclidean / edit = 1 / (1 + |x - y|) (losses.py:70-82).  logs: 0 total_loss 1 pos_similarity 2 neg_similarity
 * 3 distill_loss (= 0) 4 seperability (= 0).  z_out [2B,D] (may be NULL) receives the raw encoder outputs. */
int dof_contrastive_loss_grad(dof_handle* h, const float* state, float* grad, const float* x2, const float* a2, int B,
                              int loss_kind, int sim_kind, float temperature, float tau_plus, float beta, float* logs,
                              float* z_out, void* stream);
/* the same step with the distillation head (see dof_distill_cfg) */
int dof_contrastive_loss_grad_distill(dof_handle* h, const float* state, float* grad, const float* x2, const float* a2,
                                      int B, int loss_kind, int sim_kind, float temperature, float tau_plus, float beta,
                                      const dof_distill_cfg* distill, float* logs, float* z_out, void* stream);

/* validation step of fit_contrastive: eval-mode encoder on both views + loss, teacher off, no gradient */
int dof_contrastive_loss_eval(dof_handle* h, const float* state, const float* x2, const float* a2, int B, int loss_kind, int sim_kind,
                              float temperature, float tau_plus, float beta, float* logs, float* z_out, void* stream);

/* ---- transformer encoder, EVAL-mode forward -------------------------------------------------
 * TFMEncoderPT.forward with the module in eval() (deepof/clustering/models_new.py:985-1164): the embedding path
 * `model.encoder(x, a)` of the transformer model family (embedding_per_video, model_utils_new.py:545-621).  Stateless
 * calls: `state` is the flat float parameter vector in the reference's state_dict order (dof_tfm_entry enumerates names,
 * offsets and shapes; the integer num_batches_tracked buffers are not part of it), `workspace` is caller-owned device
 * memory of dof_tfm_workspace_bytes(cfg, B).  The training step of this encoder is NOT built yet. */
typedef struct dof_tfm_cfg {
    int T, N, E, F, Fe, D;
    int key_dim;   /* min(64, N*F) rounded down to a multiple of heads (models_new.py:1014-1019) */
    int heads;     /* 4 */
    int dff;       /* 128 */
    int layers;    /* 2 */
} dof_tfm_cfg;
int64_t dof_tfm_numel(const dof_tfm_cfg* cfg);
int dof_tfm_num_entries(const dof_tfm_cfg* cfg);
int dof_tfm_entry(const dof_tfm_cfg* cfg, int index, char* name_out, int64_t* offset_out, int64_t* numel_out,
                  int* ndim_out, int* shape_out);
size_t dof_tfm_workspace_bytes(const dof_tfm_cfg* cfg, int B);
int dof_tfm_encode(const dof_tfm_cfg* cfg, const float* state, const float* x, const float* a, int B, void* workspace,
                   size_t workspace_bytes, float* enc_out, float* nodes_out, float* edges_out, void* stream);

/* ---- transformer decoder, EVAL-mode forward -------------------------------------------------
 * TFMDecoderPT.forward in eval() (models_new.py:1167-1266; CausalSelfAttentionLayer :1270-1327): latent z [B,D] ->
 * loc [B,T,Dx], the mean of the reconstruction distribution (Dx = N*F; model_dim = 4*D, heads 8, dff 128, 2 layers in
 * the reference).  State = the decoder's float parameters in state_dict order (dof_tfm_dec_entry), names without the
 * "decoder." prefix.  Built from the library's GEMM / LayerNorm kernels plus GELU, positional input and causal
 * attention kernels.  The training step is NOT built yet. */
typedef struct dof_tfm_dec_cfg { int T, Dx, D, heads, dff, layers; } dof_tfm_dec_cfg;
int64_t dof_tfm_dec_numel(const dof_tfm_dec_cfg* cfg);
int dof_tfm_dec_num_entries(const dof_tfm_dec_cfg* cfg);
int dof_tfm_dec_entry(const dof_tfm_dec_cfg* cfg, int index, char* name_out, int64_t* offset_out, int64_t* numel_out,
                      int* ndim_out, int* shape_out);
size_t dof_tfm_dec_workspace_bytes(const dof_tfm_dec_cfg* cfg, int B);
int dof_tfm_decode(const dof_tfm_dec_cfg* cfg, const float* state, const float* z, int B, void* workspace,
                   size_t workspace_bytes, float* loc_out, void* stream);

/* Read-outs on an encoder output enc [B,D] (what embedding_per_video needs from the transformer model family):
 * dof_latent_eval = GaussianMixtureLatentPT in eval mode (models_new.py:1745-1791): emb = z_mean, q = GMM posterior;
 * scratch 3*B*D floats.  dof_vq_eval = VectorQuantizerPT (models_new.py:1358-1423): quantized latents, soft counts,
 * code indices; scratch (4 + D*D + K) doubles. */
int dof_latent_eval(const float* enc, const float* Wm, const float* bm, const float* Wv, const float* bv,
                    const float* gmm_mu, const float* gmm_lv, const float* prior, int B, int D, int K, float* emb,
                    float* q, float* scratch, void* stream);
int dof_vq_eval(const float* enc, const float* codebook, int B, int D, int K, float* quant, float* soft, int* idx,
                double* scratch, void* stream);

/* Debug / test access to intermediate activations of the last forward (device pointers into
 * the workspace; NULL if unknown).  Names: "node_out","edge_out","enc","z","z_mean",
 * "z_log_var","q","loc","len_node","len_edge". */
const void* dof_debug_tensor(dof_handle* h, const char* name, int64_t* numel_out);

/* ---- measurement support: number of kernels this library has launched so far, and optional
 * per-kernel-class CUDA-event timing (events bracket each launch on its stream). */
long long dof_launch_count(void);
/* 1 (default): GEMMs run on tcgen05 tensor cores (3xTF32, fp32-class accuracy) where eligible;
 * 0: fp32 SIMT kernels only.  Returns the previous setting.  Env DOF_DISABLE_TC=1 sets 0. */
int dof_set_tensor_cores(int enable);
/* 1 (default): the two independent recurrent blocks of the encoder (nodes, edges) run concurrently on the caller's
 * stream and a library-owned auxiliary stream (fork / join with events, nothing synchronises the host); 0: everything
 * on the caller's stream.  Env DOF_SINGLE_STREAM=1 disables the auxiliary stream at dof_create.  Returns the previous
 * setting. */
int dof_set_concurrency(int enable);
int dof_profile_begin(void);
int dof_profile_end(char* out, size_t cap);

/* ---- op-level test hooks (used by tests/ to check single kernels against torch) ---------- */
int dof_test_gemm_rows(const float* A, int lda, int mode, int p0, int p1, int p2, const float* W, int ldw,
                       int wT, const float* bias, float* C, int ldc, int M, int N, int K, int relu,
                       int accum, const float* mask, void* stream);
int dof_test_gemm_wgrad(const float* P, int ldp, int pmode, int pp0, int pp1, const float* Q, int ldq,
                        int qmode, int qp0, int qp1, float* dW, int ldo, int oT, float* db, int M, int N,
                        int K, void* stream);
int dof_test_gru_fwd(const float* gi_f, const float* gi_b, long long gi_ss, int gi_st, const float* whh_f,
                     const float* whh_b, const float* bhh_f, const float* bhh_b, const int* len,
                     float* hout, float* gt_f, float* gt_b, float* hn, int S, int T, int H, void* stream);
/* w8 (HOST array of 8 DEVICE pointers): W_ih fwd, W_ih bwd, W_hh fwd, W_hh bwd, b_ih fwd, b_ih bwd, b_hh fwd, b_hh bwd */
int dof_test_gru_layer_fwd(const float* X, long long x_ss, int x_st, const float* const* w8, const int* len,
                           float* hout, float* gt_f, float* gt_b, float* hn, int S, int T, int H, int I, int gt_tiled,
                           void* stream);
/* fused BPTT: gates in the tiled layout written by dof_test_gru_layer_fwd(gt_tiled = 1) (buffers of
 * ceil(S/128)*128*T*4H floats); dx [S,T,I] may be NULL, dxmask [S,T,I] may be NULL */
int dof_test_gru_layer_bwd(const float* const* w8, const int* len, const float* hout, const float* gtT_f,
                           const float* gtT_b, const float* dout, const float* dhn, float* dg_f, float* dg_b, float* dx,
                           const float* dxmask, int S, int T, int H, int I, void* stream);
/* second-generation fused backward (gru_bwdw_tc.cuh): BPTT + input gradient + the four parameter gradients of both
 * directions in ONE kernel (the gate gradients never reach HBM).  X [S,T,I] layer input; gates tiled as above; dx [S,T,I]
 * (required); out (zeroed by the caller, accumulated into) laid out like dof_test_gru_wgrad's.  Replaces autograd of
 * torch.nn.GRU (deepof/clustering/models_new.py:217-278, 326-373). */
int dof_test_gru_layer_bwdw(const float* X, const float* const* w8, const int* len, const float* hout, const float* gtT_f,
                            const float* gtT_b, const float* dout, const float* dhn, float* dx, const float* dxmask,
                            float* out, int S, int T, int H, int I, void* stream);
/* profiling hook: dbg = device buffer of 8 * T * 4 int64 clock stamps written by CTA (0, 0) of the following
 * dof_test_gru_layer_bwdw launches (roles: two gate warps, the two MMA issuers, two loader warps); NULL switches it off */
int dof_test_gru_bwdw_timeline(long long* dbg);
/* merged GRU parameter gradients (gru_wgrad_tc.cuh): dg_f / dg_b [M,4H] = [dr, dz, dn*r, dn], x [M,I] (pitch ldx),
 * hout [M,2H]; out (zeroed by the caller, accumulated into) = per direction dW_ih [3H,I] | dW_hh [3H,H] | db_ih [3H] |
 * db_hh [3H].  Returns DOF_ERR_UNSUPPORTED when the shape is not eligible for the tensor-core kernel. */
int dof_test_gru_wgrad(const float* dg_f, const float* dg_b, const float* x, int ldx, const float* hout, float* out,
                       int M, int T, int I, int H, void* stream);
/* multi-head attention of the transformer training step (tfm.cuh): qkv [S,T,3*dm] rows q|k|v, kpad [S,T] bytes (1 =
 * padded key) or NULL, keep [S,heads,T,T] bytes (1 = kept; inverted dropout with `rate`) or NULL, causal 0/1.
 * dout == NULL: forward, out [S,T,dm].  dout != NULL: backward, dqkv [S,T,3*dm] (the probabilities are recomputed). */
int dof_test_tfm_attention(const float* qkv, const unsigned char* kpad, const unsigned char* keep, float rate,
                           int causal, int S, int T, int dm, int heads, float* out, const float* dout, float* dqkv,
                           void* stream);
/* the encoder alone in train mode + its backward from denc [B,D] = d(loss)/d(encoder output); `groups` = row ranges with
 * separate batch statistics (transformer head; 1 otherwise).  grad is overwritten, enc_out [B,D] may be NULL. */
int dof_test_encoder_grad(dof_handle* h, const float* state, float* grad, const float* x, const float* a, int B, int groups,
                          const float* denc, float* enc_out, void* stream);
int dof_test_gru_bwd(const float* whh_f, const float* whh_b, const int* len, const float* hout,
                     const float* gt_f, const float* gt_b, const float* dout, const float* dhn,
                     float* dg_f, float* dg_b, int S, int T, int H, void* stream);
int dof_test_layernorm(const float* x, const float* w, const float* b, float eps, float* y, float* mu,
                       float* rstd, const float* dy, float* dx, float* dw, float* db, long long R, int W,
                       int relu_in, void* stream);

/* test hook: one dilated causal Conv1d(k = 4) of a TemporalBlockPT (models_new.py:408-437) as the library runs it — rows are
 * (sequence, step) pairs, X [R, ldx] (ldx >= cin, pad columns zero), W the torch weight [C, cin, 4], T steps per sequence.
 *   mode 0: A [R, C] = conv(X) + bias                       (tensor-core kernel when R >= 2048 and ldx % 4 == 0)
 *   mode 1: dX [R, cin] = conv^T(A)   (A holds d(loss)/d(conv output); requires ldx == cin, cin % 4 == 0 for the fast path)
 *   mode 2: dW [C, cin, 4] += A^T (taps of X), db [C] += sum A   (one fused tensor-core GEMM when R >= 4096) */
int dof_test_tcn_conv(int mode, const float* X, int ldx, int cin, int T, int dilation, const float* W, const float* bias, float* A,
                      int C, long long R, float* dX, float* dW, float* db, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPOF_B200_H */
