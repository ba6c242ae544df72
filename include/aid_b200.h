/* aid_b200.h — C ABI of libaid_sm100.so: the B200-native (sm_100a) hot path of
 * neuronphysics/active-inference-diffusion.
 *
 * The reference has NO native/FFI layer: its boundary for this path is the
 * Python nn.Module surface.  Each entry point below names the reference method it replaces;
 * the Python mirror classes in active_inference_diffusion_b200/ bind these through ctypes
 * (see INTEGRATION.md for the stub a maintainer would add to the reference itself).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - the library never allocates, frees or retains device memory: packed weights and
 *     workspace are caller-owned buffers whose sizes come from the *_bytes() queries;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*), nothing synchronises;
 *   - return 0 on success, negative on error; aid_last_error() gives the thread-local message;
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef AID_B200_H
#define AID_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AID_ABI_VERSION 4

/* LatentScoreNetwork dimensions — models/score_networks.py:20-29 */
typedef struct AidScoreDims {
  int32_t latent_dim;     /* L */
  int32_t obs_dim;        /* O (observation_dim seen by obs_encoder) */
  int32_t hidden_dim;     /* H, multiple of 64 */
  int32_t time_embed_dim; /* E, multiple of 64 (reference default 128) */
  int32_t num_blocks;     /* DiT blocks (reference default 6) */
} AidScoreDims;

/* Index of each fp32 parameter pointer in the `params` table given to aid_score_pack.
 * Names are the reference state_dict keys.  Per-block entries start at
 * AID_SP_BLOCK0 with stride AID_SP_BLOCK_STRIDE. */
enum AidScoreParam {
  AID_SP_TIME_SCALE = 0,          /* time_scale                                  [] */
  AID_SP_OUTPUT_MULTIPLIER,       /* output_multiplier                           [1] */
  AID_SP_FREQ_SCALE,              /* time_embed.0.freq_scale                     [1] */
  AID_SP_TE1_W, AID_SP_TE1_B,     /* time_embed.1            [2H,E] [2H] */
  AID_SP_TE3_W, AID_SP_TE3_B,     /* time_embed.3            [H,2H] [H] */
  AID_SP_OE0_W, AID_SP_OE0_B,     /* obs_encoder.0           [H,O] [H] */
  AID_SP_OE1_G, AID_SP_OE1_B,     /* obs_encoder.1 (LayerNorm) [H] [H] */
  AID_SP_OE4_W, AID_SP_OE4_B,     /* obs_encoder.4           [H,H] [H] */
  AID_SP_OE5_G, AID_SP_OE5_B,     /* obs_encoder.5 (LayerNorm) */
  AID_SP_OE7_W, AID_SP_OE7_B,     /* obs_encoder.7           [H,H] [H] */
  AID_SP_OE8_G, AID_SP_OE8_B,     /* obs_encoder.8 (LayerNorm) */
  AID_SP_CE0_W, AID_SP_CE0_B,     /* continuous_time_embed.0 [E,1] [E] */
  AID_SP_CE2_W, AID_SP_CE2_B,     /* continuous_time_embed.2 [E,E] [E] */
  AID_SP_CE4_W, AID_SP_CE4_B,     /* continuous_time_embed.4 [H,E] [H] */
  AID_SP_LP_W, AID_SP_LP_B,       /* latent_proj             [H,L] [H] */
  AID_SP_NF_W, AID_SP_NF_B,       /* norm_final.adaLN_modulation.1 [2H,H] [2H] */
  AID_SP_OUT0_W, AID_SP_OUT0_B,   /* output_proj.0           [H/2,H] [H/2] */
  AID_SP_OUT2_W,                  /* output_proj.2.weight    [L,H/2] */
  AID_SP_BLOCK0                   /* first per-block entry */
};
/* per block i: index = AID_SP_BLOCK0 + i*AID_SP_BLOCK_STRIDE + one of: */
enum AidScoreBlockParam {
  AID_SPB_N1_W = 0, AID_SPB_N1_B, /* norm1.adaLN_modulation.1 [2H,H] [2H] */
  AID_SPB_N2_W, AID_SPB_N2_B,     /* norm2.adaLN_modulation.1 */
  AID_SPB_INPROJ_W, AID_SPB_INPROJ_B, /* attention.in_proj_weight [3H,H], in_proj_bias [3H] */
  AID_SPB_OUTPROJ_W, AID_SPB_OUTPROJ_B, /* attention.out_proj [H,H] [H] */
  AID_SPB_FC1_W, AID_SPB_FC1_B,   /* mlp.0 [4H,H] [4H] */
  AID_SPB_FC2_W, AID_SPB_FC2_B,   /* mlp.2 [H,4H] [H] */
  AID_SP_BLOCK_STRIDE
};

int32_t aid_abi_version(void);
/* tensor-core operand element type this build of the library packs and multiplies: 0 = bf16
 * (libaid_sm100.so), 1 = IEEE fp16 (libaid_sm100_f16.so: 11-bit significand, the TF32-class mode of
 * the rel-1e-3 contract).  Packed buffers are only valid for the library that wrote them. */
int32_t aid_operand_type(void);
const char* aid_last_error(void);
/* number of CUDA devices visible; <= 0 means the product path cannot run */
int32_t aid_device_count(void);
/* kernels launched by this library on the calling thread since the last reset (bench evidence) */
int64_t aid_launch_count(void);
void aid_reset_launch_count(void);

/* Device timing of one GEMM class for the roofline line of bench.py: while selected, CUDA events
 * bracket every tcgen05 GEMM launch with this epilogue kind (0 pack, 1 f32/residual, 2 adaLN,
 * 3 score/step), reduction length k and output width n.  epi < 0 switches it off.
 * aid_profile_collect synchronises on the recorded events and returns their summed duration. */
int32_t aid_profile_select(int32_t epi, int32_t k, int32_t n);
int32_t aid_profile_collect(double* total_ms_host, int64_t* launches_host);
/* the same with up to 4 independent selections (slot 0 = the pair above) */
int32_t aid_profile_select_slot(int32_t slot, int32_t epi, int32_t k, int32_t n);
int32_t aid_profile_collect_slot(int32_t slot, double* total_ms_host, int64_t* launches_host);

/* ---- weights ------------------------------------------------------------------------------
 * Derived cache of LatentScoreNetwork parameters: 16-bit tcgen05 operand tiles (128 rows x 64
 * columns, K-major NO-swizzle canonical layout: [16-byte K chunk][row][8 elements], one bulk copy
 * per tile), the single-token attention folded to one HxH matrix, adaLN modulation rows
 * interleaved per 64 hidden columns, fp32 biases/LayerNorm affine/scalars.
 * Rebuild whenever the parameters change (optimizer.step / load_state_dict). */
size_t aid_score_packed_bytes(const AidScoreDims* dims);
int32_t aid_score_num_params(const AidScoreDims* dims);
int32_t aid_score_pack(const AidScoreDims* dims, const float* const* params_host_table,
                       int32_t num_params, void* packed, size_t packed_bytes, void* stream);

/* ---- LatentScoreNetwork.forward — models/score_networks.py:101-171 -------------------------
 * z_t [B,L], time [B], observation [B,O] or NULL (-> zero embedding :146-149), score_out [B,L],
 * all fp32 row-major contiguous.  `continuous` is the caller-resolved batch-global branch of
 * :121 (time.max()<=1 && time.min()>=0). */
size_t aid_score_workspace_bytes(const AidScoreDims* dims, int32_t batch, int32_t table_rows);
int32_t aid_score_forward(const AidScoreDims* dims, const void* packed, void* workspace,
                          size_t workspace_bytes, const float* z_t, const float* time,
                          const float* observation, int32_t batch, int32_t continuous,
                          float* score_out, void* stream);

/* ---- reverse diffusion — core/diffusion.py:176-255 and utils/async_collector.py:530-595 -----
 * Runs `n_steps` denoising steps.  Step s calls the score net with the batch-constant time
 * step_time_host[s] (T-1..0 as floats for generate_latent_trajectory; step/(T-1) for the
 * collector loop) and applies p_sample with coefficient index step_index_host[s].
 * coef_host = 5 arrays of length T, concatenated: sqrt_one_minus_alphas_cumprod,
 * 1/sqrt(alphas), posterior_mean_coef1, posterior_mean_coef2, sqrt(posterior_variance)
 * (evaluated by the caller with the same fp32 torch expressions as the reference).
 * noise: [n_noise, B, L] standard normals in draw order (one per step whose index != 0), or NULL
 * for deterministic=True.  traj_out (optional): [(n_steps+1), B, L] receives z after every step
 * (slot 0 = z_init).  z_out [B,L] receives the final latent. */
int32_t aid_sample(const AidScoreDims* dims, const void* packed, void* workspace,
                   size_t workspace_bytes, int32_t batch, int32_t n_steps,
                   const float* step_time_host, const int32_t* step_index_host,
                   const float* coef_host, int32_t T, const float* observation,
                   const float* z_init, const float* noise, float* z_out, float* traj_out,
                   void* stream);

/* Same loop with the noise source made explicit ("noise = injected pointer or Philox
 * seed/offset").  z_init / noise as above when given.  With `philox` != NULL (DEVICE pointer to
 * {uint64 seed, uint64 call offset}) every draw that was not injected is generated inside the
 * kernels (csrc/philox.cuh: Philox4x32-10 + Box-Muller keyed by seed, call offset, draw index,
 * GLOBAL row = row_offset + row, column): z_T by a fill kernel, the per-step eps in the epilogue
 * of the last GEMM, so no [T-1,B,L] noise tensor exists.  The caller advances the device-side
 * offset between calls (a replayed CUDA graph then draws fresh noise).  deterministic != 0: no
 * per-step noise (core/diffusion.py:233).  The call enqueues only kernels, device-to-device
 * copies and memsets on `stream` (no host-memory copies), so it can be captured in a CUDA graph.
 * Batches of at most 256 rows (the sizes of the reference's callers: act() one row, the collector one
 * row per environment, train_step 256) run all n_steps as ONE persistent kernel (csrc/small.inc: up to 128
 * co-resident CTAs, grid barrier between layers; needs hidden_dim % 64 == 0 and <= 512, latent_dim % 32
 * == 0, num_blocks <= 8, otherwise and for larger batches the per-layer tcgen05 launch chain runs).
 * Same arguments, same noise streams, results equal to the chain's up to 16-bit operand rounding; the
 * environment variable AID_SMALL_MAX (0..256, read once) moves the switch-over, 0 disables the kernel. */
typedef struct AidSampleNoise {
  const float* z_init;        /* [B,L] or NULL (requires philox) */
  const float* noise;         /* [n_draws,B,L] or NULL */
  const void* philox;         /* device {uint64 seed, uint64 offset} or NULL */
  int64_t row_offset;         /* global index of row 0 */
  int32_t deterministic;
} AidSampleNoise;
int32_t aid_sample_ex(const AidScoreDims* dims, const void* packed, void* workspace,
                      size_t workspace_bytes, int32_t batch, int32_t n_steps,
                      const float* step_time_host, const int32_t* step_index_host,
                      const float* coef_host, int32_t T, const float* observation,
                      const AidSampleNoise* noise, float* z_out, float* traj_out, void* stream);
/* out[rows, cols] (cols % 4 == 0) = the standard normals draw `draw` of the Philox stream above */
int32_t aid_philox_normal(const void* philox, uint32_t draw, int64_t row_offset, float* out, int32_t rows,
                          int32_t cols, void* stream);

/* ---- training: score matching + gradient penalty through the trunk of the score net -----------
 * Reference: core/active_inference.py:584-606 (s_theta on the noised latent, score-matching
 * term) and :709-729 (_compute_gradient_penalty: g = d(sum s)/dz with create_graph=True and the
 * double backward of mean((|g|_2 - 1)^2)); network models/score_networks.py:151-171.
 * The trunk = latent_proj, the DiT blocks (single-token attention folded to W_f = W_o W_v by the
 * caller), norm_final and output_proj, clamp, output_multiplier and time weight.  Its conditioning
 * input `cond` [B,H] (time embedding + observation embedding, pre-SiLU) and the fold stay with the
 * caller's autograd; this library returns d cond and the gradients of the folded tensors.
 * Four passes share one caller-owned workspace (csrc/train.inc; derivation oracle/manual_score_grad.py):
 *   aid_dsm_forward                    s [B,L]
 *   aid_gp_forward_backward(phase 0)   g [B,L] = d(sum s)/dz
 *   aid_gp_forward_backward(phase 1)   adjoint of phase 0 for g_bar = dL/dg: WRITES the penalty's part
 *                                      of the weight gradients into `grads`, leaves the second-order
 *                                      terms in the workspace (s_bar is read for the common stream scale)
 *   aid_dsm_backward                   backward for s_bar = dL/ds (+ the terms of phase 1 when
 *                                      with_penalty != 0, accumulating into `grads`): every trunk
 *                                      gradient, d cond [B,H], and dz [B,L] (score-matching stream
 *                                      only: the reference detaches the penalty's input) unless NULL.
 * params / grads: tables of DEVICE pointers, fp32 row-major, in this order:
 *   latent_proj.weight [H,L], .bias [H], output_proj.0.weight [H/2,H], .bias [H/2],
 *   output_proj.2.weight [L,H/2], output_multiplier [1],
 *   W_mod [(2*blocks+1)*2H, H], b_mod [(2*blocks+1)*2H]   (adaLN_modulation.1 of norm1, norm2 of
 *   every block in order, then norm_final, concatenated),
 *   then per block: W_f [H,H], b_f [H], mlp.0.weight [4H,H], .bias [4H], mlp.2.weight [H,4H], .bias [H].
 * aid_dsm_backward runs stages [stage_begin, stage_end) of 2 + num_blocks (0: output head; 1+k: block
 * num_blocks-1-k; last: latent_proj, dz, modulation Linear, d cond): a data-parallel caller all-reduces
 * a finished stage's gradients on another stream while the next stage computes.  (0, 2+num_blocks) = all.
 * tw: per-row time weight [B] of the continuous-time branch (models/score_networks.py:137,170) or NULL.
 * AidScoreDims: latent_dim % 8 == 0, hidden_dim % 128 == 0 (obs_dim / time_embed_dim unused). */
size_t aid_train_packed_bytes(const AidScoreDims* dims);
int32_t aid_train_num_params(const AidScoreDims* dims);
int32_t aid_train_pack(const AidScoreDims* dims, const float* const* params_host_table, int32_t num_params,
                       void* packed, size_t packed_bytes, void* stream);
size_t aid_train_workspace_bytes(const AidScoreDims* dims, int32_t batch);
int32_t aid_dsm_forward(const AidScoreDims* dims, const void* packed, void* workspace, size_t workspace_bytes,
                        int32_t batch, const float* z, const float* cond, const float* tw, float* score_out,
                        void* stream);
int32_t aid_gp_forward_backward(const AidScoreDims* dims, const void* packed, void* workspace,
                                size_t workspace_bytes, int32_t batch, int32_t phase, const float* tw,
                                float* g_out, const float* g_bar, const float* s_bar,
                                float* const* grads_host_table, void* stream);
int32_t aid_dsm_backward(const AidScoreDims* dims, const void* packed, void* workspace, size_t workspace_bytes,
                         int32_t batch, const float* s_bar, const float* tw, const float* cond,
                         int32_t with_penalty, float* const* grads_host_table, float* dz, float* dcond,
                         int32_t stage_begin, int32_t stage_end, void* stream);
/* Primitive of the weight gradients: out[N,K] = dy[rows,N]^T x[rows,K], fp32 row-major in and out.
 * Both operands are packed row-major (as the training passes' producers leave them) and read as
 * MN-major tcgen05 operands: no transposed pack. */
size_t aid_wgrad_workspace_bytes(int32_t rows, int32_t N, int32_t K);
int32_t aid_wgrad(const float* dy, const float* x, float* out, int32_t rows, int32_t N, int32_t K,
                  void* workspace, size_t workspace_bytes, void* stream);
/* byte offset of a named saved tensor inside the workspace (tests compare them with the
 * specification); -1 for an unknown name */
int64_t aid_train_debug_offset(const AidScoreDims* dims, int32_t batch, const char* name, int32_t index);

/* ---- EFE heads: policy / dynamics / value / reward -----------------------------------------
 * DiffusionConditionedPolicy (models/policy_networks.py:12-146, num_layers=3, state-dependent std),
 * LatentDynamicsModel (models/dynamics_models.py:9-67, num_layers=3), ValueNetwork
 * (models/value_networks.py:9-60, num_layers=3) and the reward predictor
 * (core/active_inference.py:160-167) as DiffusionActiveInference builds them (:83-107). */
typedef struct AidHeadsDims {
  int32_t latent_dim;      /* L */
  int32_t action_dim;      /* A */
  int32_t hidden_dim;      /* H, multiple of 64 */
  int32_t time_embed_dim;  /* E of ValueNetwork.time_embed (reference: 128), multiple of 64 */
} AidHeadsDims;

/* fp32 parameter table for aid_heads_pack: state_dict order of policy_network, latent_dynamics,
 * value_network, reward_predictor (W = weight, B = bias, G = LayerNorm weight). */
enum AidHeadsParam {
  AID_HP_PE0_W = 0, AID_HP_PE0_B, AID_HP_PE1_G, AID_HP_PE1_B, AID_HP_PE3_W, AID_HP_PE3_B, /* latent_encoder.{0,1,3} */
  AID_HP_TR0_W, AID_HP_TR0_B, AID_HP_TR1_G, AID_HP_TR1_B,                                 /* trunk.{0,1} */
  AID_HP_TR3_W, AID_HP_TR3_B, AID_HP_TR4_G, AID_HP_TR4_B,                                 /* trunk.{3,4} */
  AID_HP_TR6_W, AID_HP_TR6_B, AID_HP_TR7_G, AID_HP_TR7_B,                                 /* trunk.{6,7} */
  AID_HP_MH0_W, AID_HP_MH0_B, AID_HP_MH2_W, AID_HP_MH2_B,                                 /* mean_head.{0,2} */
  AID_HP_LS0_W, AID_HP_LS0_B, AID_HP_LS2_W, AID_HP_LS2_B,                                 /* log_std_head.{0,2} */
  AID_HP_DY0_W, AID_HP_DY0_B, AID_HP_DY1_G, AID_HP_DY1_B,                                 /* dynamics network.{0,1} */
  AID_HP_DY3_W, AID_HP_DY3_B, AID_HP_DY4_G, AID_HP_DY4_B,
  AID_HP_DY6_W, AID_HP_DY6_B, AID_HP_DY7_G, AID_HP_DY7_B,
  AID_HP_DY9_W, AID_HP_DY9_B,
  AID_HP_VFREQ, AID_HP_VT1_W, AID_HP_VT1_B,                                               /* value time_embed.{0,1} */
  AID_HP_V0_W, AID_HP_V0_B, AID_HP_V1_G, AID_HP_V1_B,                                     /* value network.{0,1} */
  AID_HP_V3_W, AID_HP_V3_B, AID_HP_V4_G, AID_HP_V4_B,
  AID_HP_V6_W, AID_HP_V6_B, AID_HP_V7_G, AID_HP_V7_B,
  AID_HP_V9_W, AID_HP_V9_B,
  AID_HP_R0_W, AID_HP_R0_B, AID_HP_R1_G, AID_HP_R1_B, AID_HP_R3_W, AID_HP_R3_B, AID_HP_R5_W, AID_HP_R5_B, /* reward_predictor.{0,1,3,5} */
  AID_HP_COUNT
};

typedef struct AidEfeConfig {   /* configs/config.py:49-54 */
  float epistemic_weight, pragmatic_weight, consistency_weight, discount_factor;
} AidEfeConfig;

size_t aid_heads_packed_bytes(const AidHeadsDims* dims);
int32_t aid_heads_pack(const AidHeadsDims* dims, const float* const* params_host_table,
                       int32_t num_params, void* packed, size_t packed_bytes, void* stream);
size_t aid_heads_workspace_bytes(const AidHeadsDims* dims, int32_t batch);

/* compute_expected_free_energy_diffusion — core/active_inference.py:314-396.
 * For k < num_trajectories, t < horizon (draw index d = k*horizon + t):
 *   a = policy(z) with rsample noise policy_noise[d] [B,A]; mu' = 2z + f(z,a) (:447-464);
 *   z' = mu' + reparam_noise[d] * exp(0.5 ln 0.1); prag = pw * r(z')/tau + V(z', t);
 *   cons = -sum_a H[pi]; G_k += gamma^t (ew * epi[d] + pw * prag + cw * cons); efe = mean_k G_k.
 * epistemic: device array [K*h] of batch-constant MINE values (the estimator returns one scalar
 * per call, :1050-1053) or NULL for 0.  preference_temperature: device scalar (module buffer).
 * Outputs: efe [B]; optional first_action [B,A] (trajectory 0, step 0), pragmatic_last / consistency_last
 * [K,B] (last-step terms used for the info dict, :381-389). */
int32_t aid_efe_rollout(const AidHeadsDims* dims, const void* packed, void* workspace,
                        size_t workspace_bytes, int32_t batch, int32_t horizon,
                        int32_t num_trajectories, const AidEfeConfig* cfg,
                        const float* preference_temperature, const float* latent,
                        const float* policy_noise, const float* reparam_noise,
                        const float* epistemic, float* efe_out, float* first_action_out,
                        float* pragmatic_last, float* consistency_last, void* stream);

/* Stand-alone head forwards (called outside the rollout by agents/collector, e.g.
 * core/active_inference.py:507-510).  which: 0 policy -> out [B,2A] = (mean | raw log_std),
 * 1 dynamics(z,a) -> out [B,L] = state + net([z,a]) (models/dynamics_models.py:64-67),
 * 2 value(z,t) -> out [B,1], 3 reward -> out [B,2].  aux = action [B,A] (dynamics) or time [B] (value). */
int32_t aid_head_forward(const AidHeadsDims* dims, const void* packed, void* workspace,
                         size_t workspace_bytes, int32_t which, int32_t batch, const float* z,
                         const float* aux, float* out, void* stream);

/* ---- BeliefDynamics.update, diagonal covariance — core/belief_dynamics.py:97-172 (fp64) ----
 * Batched over `rows` independent beliefs, all arrays [rows, latent_dim] float64:
 *   g = -(mu - o)/ns^2 - mu + s;  mu' = mu - lr*g*dt/(1 + 0.1|g|) + sqrt(2 D dt)*ns*eps;
 *   v' = clamp(v * exp((2(1/ns^2 + 1) + 2D) dt), max(min_variance,1e-8), max_variance);  p' = 1/v'.
 * noise may be NULL (eps = 0).  The reference method crashes as shipped; this is
 * the closed-form restatement for the default Gaussian observation model. */
int32_t aid_fp_belief_update(const double* mean, const double* variance, const double* observation,
                             const double* score, const double* noise, int32_t rows, int32_t latent_dim,
                             double dt, double diffusion_coefficient, double learning_rate,
                             double noise_scale, double min_variance, double max_variance,
                             double* mean_out, double* variance_out, double* precision_out, void* stream);

/* ---- bias gradients of the training graph ------------------------------------------------------
 * out[n] = sum_m x[m*row_stride + n] (fp32, deterministic two-stage reduction): the `grad.sum(0)` of
 * every nn.Linear bias in the backward of compute_diffusion_elbo (core/active_inference.py:584-606;
 * autograd's generic reduction ran at ~1 TB/s on these [batch, N] tensors).  workspace: caller-owned,
 * aid_colsum_workspace_bytes(M, N) bytes. */
size_t aid_colsum_workspace_bytes(int32_t M, int32_t N);
int32_t aid_colsum(const float* x, int64_t row_stride, int32_t M, int32_t N, float* out, void* workspace,
                   size_t workspace_bytes, void* stream);

/* out[i] = gg[i] * g[i] * phi(x[i]) * (2 - x[i]^2), phi = standard normal density: the derivative of
 * gelu_backward(g, x) = g * d/dx[x Phi(x)] with respect to x, contracted with gg -- the one term of the
 * gradient penalty's double backward (core/active_inference.py:709-729) through nn.GELU
 * (models/score_networks.py:199) that autograd otherwise evaluates as ~10 element-wise kernels. */
int32_t aid_gelu_double_backward(const float* gg, const float* g, const float* x, float* out, int64_t n,
                                 void* stream);

/* ---- _update_time_importance — core/active_inference.py:750-771 -----------------------------
 * weights[bin(t_i)] <- 0.99*w + 0.01*loss_i for i = 0..n-1 in batch order (double arithmetic, fp32
 * storage after every step, exactly as the reference's .item() loop), bin(t) =
 * clamp((int64)(t*99), 0, n_bins-1).  weights [n_bins] fp32 updated in place; bins_out (optional)
 * [n] int64 receives the bin indices. */
int32_t aid_time_importance_update(const float* t, const float* loss, int32_t n, float* weights,
                                   int32_t n_bins, int64_t* bins_out, void* stream);

/* ---- compute_lambda_returns — core/active_inference.py:638-707 -----------------
 * Dreamer-style lambda-returns over the batch axis exactly as the reference's Python loop computes
 * them (fp32, same operation order; bit-identical to it on CPU tensors): out[i] mixes the n-step
 * returns n = 1..min(n_steps, batch-1-i) with weights (1-lambda)lambda^(n-1), the last one taking
 * lambda^(N-1), normalised by their sum + 1e-8; indices with no n-step return get the one-step TD
 * target.  rewards, next_values, out: [batch] fp32; dones: [batch] uint8 (non-zero = terminal).
 * The reference's `values` argument is unused by its body (:641) and has no counterpart here. */
int32_t aid_lambda_returns(const float* rewards, const float* next_values, const uint8_t* dones,
                           int32_t batch, double discount_factor, double lambda, int32_t n_steps,
                           int32_t exclude_immediate_rewards, float* out, void* stream);

/* ---- DrQV2Encoder.forward — encoder/visual_encoders.py:13-189 (+ SpatialAttention :192-224) ---
 * SURVEY.md §8 f-1.  Eval-mode forward: 3x3 convs (first stride 2) with spectral-norm scaling
 * W/sigma, sigma = u^T W v from the stored buffers (no power iteration), GroupNorm(min(32,C/4)) +
 * Mish, spatial attention x*(1+sigmoid(conv7x7([mean_c, max_c])/temperature)), flatten,
 * LayerNorm(D), Linear(D,2F), LayerNorm, Mish, Linear(2F,F), LayerNorm, tanh.  Dropout = identity.
 * Convolutions and the D->2F projection run on the tcgen05 GEMM kernels (bf16 operands, fp32
 * accumulation; precision 1 = bf16x3 operand split, fp32-grade).
 * params (aid_encoder_pack), device fp32 pointers in this order:
 *   per conv layer i: convs.i.weight_orig [Cout,Cin,3,3], convs.i.weight_u [Cout] (NULL: no spectral
 *   norm), convs.i.weight_v [Cin*9] (NULL likewise), norms.i.weight [Cout], norms.i.bias [Cout];
 *   attention.spatial_conv.weight [1,2,7,7], .bias [1], attention.temperature [1] (NULL when
 *   use_attention == 0); ln.weight [D], ln.bias [D]; output_layers.0.weight [2F,D], .bias [2F];
 *   output_layers.1.weight/.bias [2F]; output_layers.4.weight [F,2F], .bias [F];
 *   output_layers.5.weight/.bias [F].
 * pixels: [batch, in_channels, height, width] contiguous, uint8 (scaled by 1/255 as :165-166) when
 * is_u8 != 0, else fp32.  features: [batch, feature_dim] fp32. */
typedef struct AidEncoderDims {
  int32_t in_channels;   /* channels x frame_stack */
  int32_t height, width;
  int32_t num_filters;   /* multiple of 8; layer i has num_filters * 2^min(i,3) channels */
  int32_t num_layers;
  int32_t feature_dim;
  int32_t use_attention;
  int32_t precision;     /* 0 bf16, 1 bf16x3 (fixed at pack time) */
} AidEncoderDims;
size_t aid_encoder_packed_bytes(const AidEncoderDims* dims);
int32_t aid_encoder_num_params(const AidEncoderDims* dims);
int32_t aid_encoder_pack(const AidEncoderDims* dims, const float* const* params, int32_t num_params,
                         void* packed, size_t packed_bytes, void* stream);
size_t aid_encoder_workspace_bytes(const AidEncoderDims* dims, int32_t batch);
int32_t aid_encoder_forward(const AidEncoderDims* dims, const void* packed, void* workspace,
                            size_t workspace_bytes, int32_t batch, const void* pixels, int32_t is_u8,
                            float* features, void* stream);

/* ---- function-space epistemic (MINE) estimator, state observations -------------------------------
 * Replaces FunctionSpaceEpistemicEstimator.forward / compute_jacobian_features
 * (core/active_inference.py:940-1063) in eval mode (Dropout = identity), with the decoder applied as
 * decode_observation does (:237-242): observation_decoder = [Linear(L,2H) LN SiLU; Linear(2H,2H) LN
 * SiLU (+skip); Linear(2H,H) LN SiLU; Linear(H,O)] (:109-131).  Fixed widths of the reference:
 * feature_extractor O-128-256-128, jacobian_projector 512-512(LN)-J, latent_processor L-128-128,
 * mine_network (J+128)-512-512-1, ntk_samples 4.
 * One call: N = num_samples*batch rows; the decoder runs once over 5N stacked rows (base + 4 perturbed),
 * the feature extractor once over 4N, the MINE network once over 2N (joint | marginal).
 *   mean, logvar [batch,L]; z_noise [num_samples,batch,L]; dir_noise [4,N,L] (the reference's randn draws,
 *   in its order); perm_idx [N] int64 = i*batch + randperm_i(batch) (GLOBAL rows of the marginal gather);
 *   perturbation_scale: device scalar (the module parameter); running_mean: device scalar, updated in
 *   place by ema_loss's rule (:828-836: first call takes the value, later alpha*t + (1-alpha)*old).
 *   stats_out [4] (device) = mi, joint term, marginal term, running mean (the reference's four metrics);
 *   t_out [2N] or NULL = T_joint | T_marg;  partial_out [4] doubles or NULL = sum T_joint,
 *   max T_marg, sum exp(T_marg - max), N  (what a sharded caller all-reduces). */
typedef struct AidEpistemicDims {
  int32_t latent_dim;       /* L, multiple of 8 */
  int32_t hidden_dim;       /* H of the observation decoder, multiple of 64 */
  int32_t observation_dim;  /* O */
  int32_t jacobian_dim;     /* J = spatial_aggregator_output_dim (reference: 256) */
} AidEpistemicDims;
enum AidEpistemicParam {   /* W = weight, B = bias, G / BETA = LayerNorm weight / bias */
  AID_EP_D0_W = 0, AID_EP_D0_B, AID_EP_D0_G, AID_EP_D0_BETA,     /* observation_decoder.0.{0,1} */
  AID_EP_D1_W, AID_EP_D1_B, AID_EP_D1_G, AID_EP_D1_BETA,         /* observation_decoder.1.{0,1} */
  AID_EP_D2_W, AID_EP_D2_B, AID_EP_D2_G, AID_EP_D2_BETA,         /* observation_decoder.2.{0,1} */
  AID_EP_D3_W, AID_EP_D3_B,                                      /* observation_decoder.3 */
  AID_EP_FE0_W, AID_EP_FE0_B, AID_EP_FE2_W, AID_EP_FE2_B, AID_EP_FE4_W, AID_EP_FE4_B,   /* feature_extractor.{0,2,4} */
  AID_EP_JP0_W, AID_EP_JP0_B, AID_EP_JP1_G, AID_EP_JP1_BETA, AID_EP_JP4_W, AID_EP_JP4_B, /* jacobian_projector.{0,1,4} */
  AID_EP_LP0_W, AID_EP_LP0_B, AID_EP_LP2_W, AID_EP_LP2_B,        /* latent_processor.{0,2} */
  AID_EP_MN0_W, AID_EP_MN0_B, AID_EP_MN3_W, AID_EP_MN3_B, AID_EP_MN6_W, AID_EP_MN6_B,   /* mine_network.{0,3,6} */
  AID_EP_COUNT
};
size_t aid_epistemic_packed_bytes(const AidEpistemicDims* dims);
int32_t aid_epistemic_pack(const AidEpistemicDims* dims, const float* const* params, int32_t num_params,
                           void* packed, size_t packed_bytes, void* stream);
size_t aid_epistemic_workspace_bytes(const AidEpistemicDims* dims, int32_t batch, int32_t num_samples);
int32_t aid_epistemic_forward(const AidEpistemicDims* dims, const void* packed, void* workspace,
                              size_t workspace_bytes, int32_t batch, int32_t num_samples,
                              const float* mean, const float* logvar, const float* z_noise,
                              const float* dir_noise, const int64_t* perm_idx,
                              const float* perturbation_scale, float alpha, float* running_mean,
                              float* stats_out, float* t_out, double* partial_out, void* stream);
/* The same estimator over `groups` independent row sets at once: the K rollouts of one EFE evaluation
 * (core/active_inference.py:337-378 calls the estimator once per (k, t): 50 calls of ~40 launches for an
 * act() on one observation; per step t the K calls see independent rows).  batch = groups * Bg rows, group g
 * = rows g*Bg .. g*Bg+Bg-1 of mean / logvar and of every sample block of z_noise [S, batch, L]; perm_idx must
 * permute within (sample, group) blocks.  group_stats [groups, 4] = (mi, joint term, marginal term,
 * exp(marginal term)) per group; the running mean is not touched -- aid_ema_sequence applies the calls'
 * exp(marginal term) values in the reference's (k, t) order (:828-836). */
int32_t aid_epistemic_forward_grouped(const AidEpistemicDims* dims, const void* packed, void* workspace,
                                      size_t workspace_bytes, int32_t batch, int32_t num_samples, int32_t groups,
                                      const float* mean, const float* logvar, const float* z_noise,
                                      const float* dir_noise, const int64_t* perm_idx,
                                      const float* perturbation_scale, float* group_stats, void* stream);
int32_t aid_ema_sequence(const float* values, int32_t n, float alpha, float* running_mean, void* stream);

/* ---- 3x3 convolution (padding 1, stride 1 or 2, no bias) for the encoder's TRAINING graph --------
 * Replaces the nn.Conv2d calls of DrQV2Encoder.forward in training mode and their backward
 * (encoder/visual_encoders.py:56-76,166-176): forward, input gradient and weight gradient as tcgen05
 * GEMMs over an explicit im2col operand (k = tap*Cin + c), NCHW fp32 tensors in and out.
 *   forward : y [n,Cout,Ho,Wo] = conv(x [n,Cin,H,W], w [Cout,Cin,3,3] / *sigma)  (sigma: device scalar or NULL)
 *   dgrad   : stride 1 = aid_conv3x3_forward(dy, flipped + transposed w); stride 2 = aid_conv3x3_dgrad_direct
 *   wgrad   : dw [Cout,Cin,3,3] = dy_rows^T * im2col(x) through the MN-major weight-gradient GEMM;
 *             dy_scale: optional device scalar multiplied into dy before the operand rounding (keeps small
 *             cotangents inside the fp16 range; the result carries the factor).
 * precision: 0 = one pass on the library's operand type, 1 = hi/lo split (3x the reduction length). */
size_t aid_conv3x3_workspace_bytes(int32_t n, int32_t cin, int32_t cout, int32_t H, int32_t W, int32_t stride,
                                   int32_t precision);
int32_t aid_conv3x3_forward(const float* x, const float* w, const float* sigma, int32_t n, int32_t cin,
                            int32_t cout, int32_t H, int32_t W, int32_t stride, int32_t precision, float* y,
                            void* workspace, size_t workspace_bytes, void* stream);
int32_t aid_conv3x3_wgrad(const float* x, const float* dy, const float* dy_scale, int32_t n, int32_t cin,
                          int32_t cout, int32_t H, int32_t W, int32_t stride, float* dw, void* workspace,
                          size_t workspace_bytes, void* stream);
int32_t aid_conv3x3_dgrad_direct(const float* dy, const float* w, int32_t n, int32_t cin, int32_t cout,
                                 int32_t H, int32_t W, int32_t stride, float* dx, void* stream);

/* ---- primitive exposed for tests: y = act(x W^T + b) through the tcgen05 path --------------
 * x [M,K], w [N,K], bias [N] or NULL, y [M,N]; act: 0 none, 1 SiLU, 2 ReLU, 3 GELU(erf).
 * via_packed != 0 routes the result through the bf16 packed epilogue and back (tests EPI_PACK). */
size_t aid_linear_workspace_bytes(int32_t M, int32_t N, int32_t K);
int32_t aid_linear(const float* x, const float* w, const float* bias, float* y, int32_t M,
                   int32_t N, int32_t K, int32_t act, int32_t via_packed, void* workspace,
                   size_t workspace_bytes, void* stream);

/* ---- training graph GEMM: out[M,N] = A[M,K] * B[N,K]^T (+ bias[N]) ---------------------------
 * Replaces the cuBLAS calls behind every nn.Linear / F.linear of the score network inside
 * compute_diffusion_elbo and its backward / double backward (core/active_inference.py:584-606,
 * 709-729; models/score_networks.py:41-99,189-202).  Element (i,k) of A is a[i*a_rs + k*a_cs]
 * (same for B), so x W^T (forward), dY W (input gradient) and dY^T X (weight gradient) are one
 * entry point.  fp32 in / fp32 out, bf16 tensor-core operands, fp32 accumulation; long reductions
 * with few output tiles are split over K internally.
 * precision: 0 = bf16 operands (error ~2^-9 per product);  1 = bf16x3: each operand split into
 * hi + lo bf16 halves and hi*hi + hi*lo + lo*hi accumulated in one GEMM of 3x the reduction
 * length -- products exact to ~2^-16, the mode that meets the rel-1e-3 gradient contract. */
size_t aid_gemm_nt_workspace_bytes(int32_t M, int32_t N, int32_t K, int32_t precision);
int32_t aid_gemm_nt(const float* a, int64_t a_rs, int64_t a_cs, const float* b, int64_t b_rs,
                    int64_t b_cs, const float* bias, float* out, int32_t M, int32_t N, int32_t K,
                    int32_t precision, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AID_B200_H */
