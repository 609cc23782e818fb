/* ldt_b200 -- C ABI of the B200 (sm_100a) kernels behind the LDT sampling hot path.
 *
 * The reference (Negai-98/LDT) has no FFI for its model path: the seams are Python call signatures
 * (SURVEY.md section 8b).  Its one native seam is the StructuralLosses torch extension.  Every entry
 * point below names the reference interface it replaces; the Python host mirror in ldt_b200/ binds
 * them with ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name says host;
 *   - the caller owns every buffer including workspaces; nothing here allocates device memory,
 *     synchronises the device, or throws;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); work is enqueued on it;
 *   - return 0 on success, a negative LDT_ERR_* code otherwise; ldt_last_error_string() describes the
 *     most recent failure on the calling thread;
 *   - tensors are dense row-major; "token-major" means [batch * tokens, channels].
 */
#ifndef LDT_B200_H_
#define LDT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDT_ABI_VERSION 2   /* 2: ldt_sde_step takes rng_state; ldt_pairwise_cd_upper added */

#define LDT_OK 0
#define LDT_ERR_INVALID (-1)
#define LDT_ERR_CUDA (-2)
#define LDT_ERR_UNSUPPORTED (-3)
#define LDT_ERR_WORKSPACE (-4)

int ldt_abi_version(void);
const char* ldt_last_error_string(void);

/* ------------------------------------------------------------------------------------------------
 * Metrics
 * ------------------------------------------------------------------------------------------------ */

/* Nearest-neighbour squared distances, both directions, with argmin indices.
 * Replaces nndistance() -- evaluation/pytorch_structural_losses/src/nndistance.cu:125-128, bound as
 * StructuralLossesBackend.NNDistance (src/structural_loss.cpp:80-99, pybind/bind.cpp:14).
 *   xyz1 [b,n,3] f32, xyz2 [b,m,3] f32 -> dist1 [b,n] f32, idx1 [b,n] i32, dist2 [b,m], idx2 [b,m].
 * Distances and indices are bit-identical to the reference kernel (lowest index wins ties). */
int ldt_nn_distance(int b, int n, const float* xyz1, int m, const float* xyz2, float* dist1, int* idx1,
                    float* dist2, int* idx2, void* stream);

/* Rows [row_begin,row_end) of the pairwise Chamfer matrix
 *   M[i,j] = mean_p min_q |a_i[p]-b_j[q]|^2 + mean_q min_p |a_i[p]-b_j[q]|^2
 * Replaces the Python double loop _pairwise_CD_ (evaluation/evaluation_metrics.py:165-198), i.e.
 * 2*na*ceil(nb/batch) launches of NmDistanceKernel plus expand/mean/cat, by one launch.
 *   a [na,pa,3] f32, b [nb,pb,3] f32 -> out [(row_end-row_begin), nb] f32 (row-major). */
int ldt_pairwise_cd(int na, int nb, int pa, int pb, const float* a, const float* b, int row_begin, int row_end,
                    float* out, void* stream);

/* The same matrix for ONE cloud set against itself (M_rr / M_ss of compute_CD_metrics, evaluation/evaluation_metrics.py:
 * 311-312, which the reference evaluates in full): only entries on or above the diagonal are computed, for the rows
 * i = row_first, row_first + row_step, ... < n (interleaved rows balance the triangle over row_step ranks; 1 rank:
 * row_first 0, row_step 1).  The kernel's value for (i, j) is bit-identical to its value for (j, i), so mirroring the
 * result reproduces the full matrix exactly at half the pair evaluations.
 *   a [n,p,3] f32 -> out [n,n] f32 row-major, FULL-matrix indexing: out[i*n + j] written for owned i and j >= i only. */
int ldt_pairwise_cd_upper(int n, int p, const float* a, int row_first, int row_step, float* out, void* stream);

/* Approximate earth-mover distance (forward only).  Replaces approxmatch() + matchcost()
 * (evaluation/pytorch_structural_losses/src/approxmatch.cu:299-316), bound as StructuralLossesBackend.ApproxMatch /
 * MatchCost (src/structural_loss.cpp:14-78) and used through StructuralLosses.match_cost (match_cost.py:6-45).
 *   xyz1 [b,n,3] f32, xyz2 [b,m,3] f32; n, m <= 4096.
 * ldt_match_cost          cost [b]  = MatchCost(xyz1, xyz2, ApproxMatch(xyz1, xyz2)) without materialising the match
 * ldt_approx_match        match [b,m,n] f32 (dense; the caller provides b*m*n floats), as ApproxMatch returns it
 * ldt_match_cost_from_match  cost [b] from a given match, as MatchCost
 * ldt_pairwise_emd        rows [row_begin,row_end) of out[i,j] = match_cost(a_i, b_j) / p -- what _pairwise_EMD_CD_
 *                         (evaluation/evaluation_metrics.py:112-162) assembles with one kernel-launch pair per
 *                         (row, column batch); a [na,p,3], b [nb,p,3], out [(row_end-row_begin), nb] row-major. */
int ldt_match_cost(int b, int n, int m, const float* xyz1, const float* xyz2, float* cost, void* stream);
int ldt_approx_match(int b, int n, int m, const float* xyz1, const float* xyz2, float* match, void* stream);
int ldt_match_cost_from_match(int b, int n, int m, const float* xyz1, const float* xyz2, const float* match, float* cost,
                              void* stream);
int ldt_pairwise_emd(int na, int nb, int p, const float* a, const float* b, int row_begin, int row_end, float* out,
                     void* stream);

/* ------------------------------------------------------------------------------------------------
 * Dense contraction core (used by the score net and the decoder; exported for parity tests)
 *   C[M,N] = epilogue( A[M,K] . W[N,K]^T + bias[N] )
 * A, W are bf16 K-major (row-major with K contiguous, leading dimensions lda/ldw in elements,
 * multiples of 8; K a multiple of 64) -- or f32 / TF32 with operand_type = 1.  W is exactly a Conv1d(k=1)/Linear weight [out,in].
 * This is what replaces every nn.Conv1d(k=1)/nn.Linear on the path (model/layers.py:159-161,
 * 120-124, 172, 238; model/scorenet/score.py:95; model/Compressor/Network.py:61,153).
 * ------------------------------------------------------------------------------------------------ */
enum ldt_epilogue {
  LDT_EPI_BIAS_F32 = 0,       /* out f32  = acc + bias                                            */
  LDT_EPI_BIAS_BF16 = 1,      /* out bf16 = acc + bias                                            */
  LDT_EPI_BIAS_GELU_BF16 = 2, /* out bf16 = gelu_erf(acc + bias)       (MLP fc, layers.py:127-128) */
  LDT_EPI_GATE_RESID_F32 = 3, /* out f32  = resid + gate[row/rows_per_gate] * (acc + bias); gate may
                                 be NULL (= 1): x + gate*f(x) of layers.py:218-219,225-226           */
  LDT_EPI_BIAS_GELU_F32 = 4,  /* out f32  = tf32_round(gelu_erf(acc + bias)): the MLP hidden of the TF32 parity mode
                                 (operand_type 1), rounded because it is the A operand of the next contraction */
  LDT_EPI_BIAS_RELU_F32 = 5,  /* out f32  = max(acc + bias, 0): Conv1d(k=1) + BatchNorm1d (folded into W and
                                 bias by the host) + ReLU of the encoder prologue (ConvBNReLU1D, Compressor/layers.py:115-127;
                                 MiniPointnet, Network.py:94-95)                                     */
  LDT_EPI_RESID_RELU_F32 = 6, /* out f32  = max(resid + acc + bias, 0): ConvBNReLURes1D's act(net2(net1(x)) + x),
                                 Compressor/layers.py:159-160; no gate                               */
};

typedef struct ldt_gemm_args {
  int M, N, K;
  const void* A;   /* bf16 [M, lda]   */
  int lda;
  const void* W;   /* bf16 [N, ldw]   */
  int ldw;
  const float* bias; /* [N] or NULL   */
  void* out;       /* f32 or bf16 [M, ldo] */
  int ldo;
  int epilogue;    /* enum ldt_epilogue */
  const float* resid; /* f32 [M, ldo] (may alias out) for LDT_EPI_GATE_RESID_F32 */
  const float* gate;  /* f32 rows of length >= N; row r of the output uses gate + (r / rows_per_gate) *
                         gate_stride; gate_stride == 0 broadcasts one row */
  long long gate_stride;
  int rows_per_gate;
  int backend;     /* 0 = tcgen05/TMEM/TMA kernel, tile shape chosen by the library (product path);
                      1 = naive SIMT cross-check kernel (tests only);
                      2 = force single-CTA 128-row tiles; 3 = force CTA-pair (cta_group::2) 256-row tiles */
  int operand_type; /* 0 = A, W are bf16 (tcgen05 kind::f16; the product path);
                       1 = A, W are f32 holding TF32-rounded values (tcgen05 kind::tf32, same pipeline, CTA-pair kernel
                           only): the parity mode that matches the precision of the reference's own GPU arithmetic
                           (cuDNN TF32 convolutions).  K a multiple of 32, lda / ldw multiples of 4; epilogues 0, 3, 4, 5, 6;
                       2 = as 1 for operands laid out by ldt_split_tf32 (3xTF32, fp32-grade): LDT_EPI_BIAS_GELU_F32 then does
                           NOT round its output (the consumer splits it again) */
} ldt_gemm_args;

int ldt_gemm_bf16(const ldt_gemm_args* args, void* stream);

/* Fused MLP half of a transformer block, one launch:
 *     out[M,C] = resid + gate * ( GELU(A[M,C] . W1[inner,C]^T + bias1) . W2[C,inner]^T + bias2 )
 * = MLP.forward (model/layers.py:110-133: Conv1d(C->4C), nn.GELU() exact erf, Conv1d(4C->C)) followed by the gated
 * residual add of ResidualBlock.forward (layers.py:219); bit-identical to ldt_gemm_bf16(LDT_EPI_BIAS_GELU_BF16) followed by
 * ldt_gemm_bf16(LDT_EPI_GATE_RESID_F32).  fc2 tiles start as soon as the fc1 tiles of their 256 rows are stored.
 *   A, W1, W2 bf16 row-major; `hidden` = caller-owned bf16 scratch [M, ldh >= inner] (the GELU output lives there);
 *   resid/out f32 [M, ldo] (may alias); gate as in ldt_gemm_args (NULL = 1).
 *   sync = caller-owned device words, ldt_mlp_sync_words(M) of them, ZEROED ONCE before the first call: per-256-row
 *   completion counters.  The kernel leaves them zero again (self-cleaning), so the same buffer serves every later
 *   launch on the same stream; two launches in flight at once need two buffers.
 *   Needs C % 256 == 0 and inner % 256 == 0 (LDT_ERR_UNSUPPORTED otherwise: use the two GEMM calls) and a grid of
 *   co-resident CTA pairs (checked with cudaOccupancyMaxActiveClusters). */
typedef struct {
  int M, C, inner;
  const void* A;      int lda;
  const void* W1;     int ldw1;   const float* bias1;
  void* hidden;       int ldh;
  const void* W2;     int ldw2;   const float* bias2;
  const float* resid; float* out; int ldo;
  const float* gate;  long long gate_stride; int rows_per_gate;
  unsigned int* sync;
} ldt_mlp_args;
int ldt_mlp_bf16(const ldt_mlp_args* args, void* stream);
int ldt_mlp_sync_words(int M);
/* The kernel's static tile schedule, host-side (tests): work item `it` of CTA pair `p` of `pairs`; *phase = 1 (fc1 tile),
 * 2 (fc2 tile) or 0 (surplus slot / out of range), with its 256-row block and 256-column tile. */
int ldt_mlp_schedule_item(int tiles_m, int tn1, int tn2, int kb1, int kb2, int pairs, int p, int it, int* num_items,
                          int* phase, int* mb, int* nt);

/* Diagnostics: when dev_buf != NULL, every CTA of the CTA-pair GEMM kernel writes 8 u64 stall counters
 * (clock64 ticks) at dev_buf[8*blockIdx.x ...]: [0] MMA thread total, [1] its wait for TMA data, [2] its wait for
 * a free accumulator, [3] epilogue warp total, [4] its wait for the accumulator, [5] TMA producer wait for a free
 * stage, [6] producer total, [7] tiles.  dev_buf must hold 8 * gridDim u64 (<= 8 * SM count).  NULL switches it off. */
int ldt_debug_set_gemm_counters(unsigned long long* dev_buf);

/* Diagnostics (scripts/exp_gemm_limits.py only; results are WRONG while set): bit 0 = the CTA-pair GEMM skips its A
 * loads, bit 1 = skips its W loads (half the operand traffic either way), bit 2 = skips the epilogue, bit 3 = the fused
 * MLP kernel ignores its completion counters, bits 3-6 (with the counters on, GELU epilogue) = epilogue parts removed,
 * bit 8 = bf16 outputs through per-lane st.global instead of bulk tensor stores (results stay correct).  0 = off. */
int ldt_debug_set_gemm_mode(int mode);
int ldt_debug_get_gemm_mode(void);

/* Diagnostics: a measured FP32-FMA rate for the Chamfer kernel's roofline (bench.py).  Launches blocks_per_sm * SMs blocks
 * of 256 threads, each thread running `iters` rounds of 8 independent FMA chains: scalar fma.rn.f32 (packed = 0) or
 * fma.rn.f32x2 (packed = 1).  *flop (host) receives the floating-point operations the launch performs; the caller times
 * it with CUDA events.  out: any device float (never written in practice). */
int ldt_debug_fma_peak(int iters, int packed, int blocks_per_sm, float* out, long long* flop, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Element-wise / normalisation kernels of the score net and decoder
 * ------------------------------------------------------------------------------------------------ */

/* f32 [rows, cols] -> bf16 [rows, ld_out] with zero padding of columns cols..ld_out-1. */
int ldt_cast_pad_bf16(int rows, int cols, const float* in, int ld_in, void* out, int ld_out, void* stream);

/* Weight packing: f32 [rows, cols] -> bf16 [rows, ld_out] (zero padded), same kernel, named for the
 * pack step SURVEY.md 8b lists (ldt_pack_weights). */
int ldt_pack_weights(int rows, int cols, const float* w, int ld_in, void* out, int ld_out, void* stream);

/* LayerNorm over channels (eps 1e-6, tools/utils.py:127-133) fused with either AdaLN modulation
 *   y = LN(x) * (1 + scale[g]) + shift[g]     (model/layers.py:136-137, 218-219; g = row / rows_per_mod)
 * or an element-wise affine (weight, bias) (decoder blocks, layers.py:163-164 with dim_c None);
 * pass shift/scale or weight/bias, the other pair NULL.  x f32 [rows, C] -> y bf16 [rows, C]. */
int ldt_layernorm_mod_bf16(int rows, int C, const float* x, const float* shift, const float* scale,
                           long long mod_stride, int rows_per_mod, const float* weight, const float* bias,
                           float eps, void* y, void* stream);

/* Sinusoidal time features -> Linear -> SiLU -> Linear (+ optional additive embedding) -> SiLU.
 * Replaces TimeEmbedding.forward (model/layers.py:14-41) and the leading SiLU of every adaLN
 * (layers.py:172,237).  t [R] f32, freq [half] f32 (host-computed exactly as layers.py:28-30),
 * w0 [D, 2*half], w1 [D, D] f32 -> c [R, D] f32 (pre-SiLU, what Score.forward calls `c`) and
 * silu_c bf16 [R_pad, D] (the A operand of the adaLN GEMMs).  extra [R, D] f32 or NULL is added to c
 * (label / image-condition embedding, score.py:135). */
int ldt_time_embedding(int R, int half, int D, const float* t, const float* freq, const float* w0,
                       const float* b0, const float* w1, const float* b1, const float* extra, float* c,
                       void* silu_c, float* scratch /* [R, D + 2*half] f32 */, void* stream);

/* ------------------------------------------------------------------------------------------------
 * TF32 parity mode (Score.precision = "tf32"): fp32 activations, kind::tf32 contractions (ldt_gemm_bf16 with
 * operand_type 1) -- the precision of the reference's own GPU path (cuDNN TF32 convolutions).  Plain SIMT kernels.
 *   ldt_round_pad_tf32      out f32 [rows, ld_out] = tf32_round(silu ? SiLU(in) : in), zero-padded columns cols..ld_out
 *   ldt_layernorm_mod_f32   ldt_layernorm_mod_bf16 with an f32 output (any C), TF32-rounded when round_tf32_out != 0
 *   ldt_attention_nk32_f32  ldt_attention_nk32 on f32 q/k/v (dh in {32, 64}), fp32 softmax, f32 output (rounded likewise)
 * ------------------------------------------------------------------------------------------------ */
int ldt_round_pad_tf32(long long rows, int cols, const float* in, int ld_in, float* out, int ld_out, int silu, void* stream);

/* Error-compensated TF32 operands ("3xTF32") for fp32-grade contractions on the tensor cores: in f32 [rows, ld_in] ->
 * out f32 [rows, 3*ld_part], v = hi + lo with hi = tf32(v), lo = tf32(v - hi); an activation row becomes [hi | hi | lo], a
 * weight row (weight_side = 1) [hi | lo | hi], each part zero-padded to ld_part columns, so that ONE ldt_gemm_bf16 call with
 * operand_type 1 and K = 3*ld_part computes a_hi.w_hi + a_hi.w_lo + a_lo.w_hi.  Used for the fp32 Conv1d / Linear layers of
 * the encoder and condition prologues (model/Compressor/Network.py:192, layers.py:115-160, scorenet/score.py:36-41), which
 * the reference evaluates in fp32, and for Score.precision = "fp32" (the fp32-grade parity mode of the score net: the producers
 * above then run with round_tf32_out = 0 and every contraction operand goes through this split; silu = 1 applies SiLU first,
 * the adaLN input of layers.py:172). */
int ldt_split_tf32(long long rows, int cols, const float* in, int ld_in, float* out, int ld_part, int weight_side, int silu,
                   void* stream);
int ldt_layernorm_mod_f32(int rows, int C, const float* x, const float* shift, const float* scale, long long mod_stride,
                          int rows_per_mod, const float* weight, const float* bias, float eps, float* y, int round_tf32_out,
                          void* stream);
int ldt_attention_nk32_f32(int B, int H, int Nq, int dh, const float* q, int ldq, const float* k, const float* v, int ldkv,
                           float* o, int round_tf32_out, void* stream);
/* ldt_attention_longkv on f32 q / k / v with an f32 output (the fp32 parity mode of Compressor.forward's posterior blocks). */
int ldt_attention_longkv_f32(int B, int H, int Nq, int Nk, int dh, const float* q, int ldq, const float* k, const float* v,
                             int ldkv, float* o, void* stream);

/* Multi-head attention over a short key set, one (batch, head) pair per warp group.
 *   q  bf16 [B*Nq, ldq]  (head h uses columns h*dh..h*dh+dh-1 -- contiguous channel groups,
 *                          layers.py:192-194)
 *   k,v bf16 [B*Nk, ldkv] likewise
 *   o  bf16: written as the reference's (w@v).reshape(B,N,C) does (layers.py:197): the [B,H,Nq,dh]
 *      result buffer is stored contiguously and re-read as token-major [B*Nq, H*dh] WITHOUT permuting
 *      heads back.  This quirk is part of the trained weights' meaning and is reproduced on purpose.
 * Nk must be 32 (z_scale latent tokens); dh in {8, 16, 32, 64} (8 / 16: the 128-wide, 16-head score net of
 * experiments/Hybrid_Trainer/airplane/config.yaml, served by a one-lane-per-query SIMT kernel). */
int ldt_attention_nk32(int B, int H, int Nq, int dh, const void* q, int ldq, const void* k, const void* v,
                       int ldkv, void* o, void* stream);
/* Which kernel serves ldt_attention_nk32: 0 (default) = S = Q K^T and O = P V as tcgen05.mma with TMEM accumulators
 * (csrc/attention_tc.cu) for Nq == 32 or Nq >= 128 with dh in {32, 64}; 1 = the warp-level mma.sync kernels for every
 * shape (the cross-check in tests).  Both round the un-normalised probabilities to bf16 at the same point.  The fused
 * projection + attention kernel (ldt_qkv_attention_bf16) follows the same switch. */
int ldt_debug_set_attention_backend(int backend);
int ldt_debug_get_attention_backend(void);

/* The transposed shape: a SHORT query set over a LONG key set (Nq = 32 latent tokens attending to the Nk = 2048 decoded
 * points in DecoderBlock.compute_posterior, model/Compressor/Network.py:62-77), online softmax over key chunks.
 * Same operand and output layouts as ldt_attention_nk32 (q [B*Nq, ldq], k/v [B*Nk, ldkv], o = [B,H,Nq,dh] contiguous).
 * dh in {32, 64}. */
int ldt_attention_longkv(int B, int H, int Nq, int Nk, int dh, const void* q, int ldq, const void* k, const void* v,
                         int ldkv, void* o, void* stream);

/* Fused Q/K/V projection + self-attention of one score-net block (fc_q, fc_kv and compute_attention of
 * model/layers.py:186-197 in one kernel; Q, K, V stay on chip).
 *   A   bf16 [B*32, lda]   LayerNorm'd + modulated activations (K = hidden columns used)
 *   Wp  bf16 [H*192, ldw]  projection weights packed HEAD-MAJOR: rows h*192+[0,64) = fc_q rows of head h,
 *                          +[64,128) = the K rows of fc_kv, +[128,192) = its V rows (heads are contiguous
 *                          64-channel groups, layers.py:192-194)
 *   bias_p f32 [H*192] packed the same way, or NULL
 *   out bf16 [B, H, 32, 64] contiguous == the reference's (w@v).reshape(B,N,C) buffer (layout quirk, :197)
 * 32 tokens per sample, head dim 64 (the shipped score configuration); K a multiple of 64. */
int ldt_qkv_attention_bf16(int B, int H, int K, const void* A, int lda, const void* Wp, int ldw, const float* bias_p,
                           void* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Reverse-SDE predictor updates (diffusion/diffusion_continuous.py:141-191) fused with the score
 * conversion of Trainer.score_fn (trainer/Latent_SDE_Trainer.py:57-61):  score = -params / sqrt(var)
 * All arrays [numel] f32.  Per-step scalars come from a device table so the same launch can sit in
 * a CUDA graph: coef points at 8 floats for this step (see ldt_sde_coef).  Noise: z != NULL uses
 * caller-provided normals (teacher-forced parity); z == NULL draws Philox4x32-10 normals with
 * (seed, offset) laid out exactly like torch.randn_like on a CUDA generator.
 * ------------------------------------------------------------------------------------------------ */
enum ldt_predictor {
  LDT_PRED_ANCESTRAL = 0,         /* :152-162 */
  LDT_PRED_REVERSE_DIFFUSION = 1, /* :141-150 */
  LDT_PRED_EULER_MARUYAMA = 2,    /* :182-191 */
  LDT_PRED_DDIM = 3,              /* :164-180 */
  LDT_PRED_CORRECTOR = 4,         /* the update of LangevinCorrector / AncestralCorrector, :193-229 */
};
/* coef layout per step (8 floats): [0]=sqrt(var(t)) [1..7] predictor specific, see ldt_b200/sde.py */
#define LDT_SDE_COEF_STRIDE 8

/* offset_per_step: Philox offset advance per step (what torch adds per randn_like call), so step i of a
 * replayed CUDA graph uses offset + i*offset_per_step.  rng_grid: number of 256-thread blocks torch would
 * launch for `numel` elements on this device (<= 0 lets the library pick; only matters when z == NULL).
 * rng_state (device u64[2] = {seed, base offset}, or NULL): added to the by-value seed / offset inside the kernel, so a
 * captured CUDA graph can be replayed from any generator position.  x_next may alias x (in-place update). */
int ldt_sde_step(int predictor, long long numel, const float* x, const float* params, const float* z,
                 const float* coef_table, const int* step_index /* device int, or NULL = 0 */,
                 unsigned long long seed, unsigned long long offset, unsigned long long offset_per_step,
                 const unsigned long long* rng_state, int rng_grid, float* x_next, float* x_mean, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Whole-path entry points: one call per score-net evaluation / per N-step sampling loop (csrc/path.cu).  Pure
 * orchestration of the entry points above in the order ldt_b200/score.py::run_tokens and sampler.py::StepGraph issue
 * them, so results are bit-identical to the Python-orchestrated path.  Plain (non-UNet) AdaLN score net, self-attention
 * blocks with head dim 64 (the shipped experiments/Latent_Diffusion_Trainer config); other variants are orchestrated by
 * the host from the kernel-level entry points.  All pointers are device pointers except `blocks` (host array) and the
 * structs themselves.
 * ------------------------------------------------------------------------------------------------ */
typedef struct ldt_score_block {       /* packed as ldt_b200/score.py::packed does */
  const void* w_qkv_packed;  /* bf16 [heads*192, hidden]: per head [q_h | k_h | v_h] rows (fc_q / fc_kv, layers.py:159-160) */
  const float* b_qkv_packed; /* f32 [heads*192] */
  const void* w_o;           /* bf16 [hidden, hidden]      fc_o                */
  const float* b_o;
  const void* w_fc1;         /* bf16 [mlp_hidden, hidden]  mlp.fc.0.0          */
  const float* b_fc1;
  const void* w_fc2;         /* bf16 [hidden, mlp_hidden]  mlp.out             */
  const float* b_fc2;
} ldt_score_block;

typedef struct ldt_score_plan {
  int batch, tokens /* 32 */, z_dim, z_pad /* z_dim padded to a multiple of 64 */, hidden, heads, mlp_hidden, num_blocks;
  const void* w_in;   /* bf16 [hidden, z_pad]  ln_in   */
  const float* b_in;
  const void* w_out;  /* bf16 [z_dim, hidden]  ln_out.ln */
  const float* b_out;
  const ldt_score_block* blocks;   /* HOST array [num_blocks] */
  /* caller-owned device workspace for `batch` samples (M = batch * tokens rows) */
  void* ws_xa;    /* bf16 [M, z_pad]      */
  float* ws_h;    /* f32  [M, hidden]     residual stream */
  void* ws_a;     /* bf16 [M, hidden]     LayerNorm output */
  void* ws_att;   /* bf16 [M, hidden]     attention output, [B,H,32,64] contiguous (layers.py:197) */
  void* ws_hid;   /* bf16 [M, mlp_hidden] */
} ldt_score_plan;

/* out f32 [M, z_dim] = the score net's token path on x_tokens f32 [M, z_dim] given the AdaLN rows `mod`
 * (f32, 6*hidden per block then 2*hidden for the final layer; one row per sample at stride mod_stride, or one broadcast
 * row with mod_stride 0): score.py:136-150.  The AdaLN rows themselves come from ldt_time_embedding + ldt_gemm_bf16. */
int ldt_score_forward(const ldt_score_plan* plan, const float* x_tokens, const float* mod, long long mod_stride, float* out,
                      void* stream);

typedef struct ldt_decoder_layer {     /* one DecoderBlock as ldt_b200/compressor.py::packed lays it out */
  const void* w_ln;  const float* b_ln;     /* bf16 [hidden, z_pad]  decoder.l.ln (Conv1d z_dim -> hidden)     */
  const void* w_kv;  const float* b_kv;     /* bf16 [2*hidden, hidden] att1.fc_kv                               */
  const void* w_q;   const float* b_q;      /* bf16 [hidden, hidden]   att1.fc_q                                */
  const void* w_o;   const float* b_o;      /* att1.fc_o */
  const void* w_fc1; const float* b_fc1;    /* bf16 [mlp_hidden, hidden] att1.mlp.fc.0.0                        */
  const void* w_fc2; const float* b_fc2;    /* bf16 [hidden, mlp_hidden] att1.mlp.out                           */
  const float* norm1_w; const float* norm1_b; const float* norm2_w; const float* norm2_b;   /* affine LayerNorms */
} ldt_decoder_layer;

typedef struct ldt_decoder_plan {
  int batch, num_points, z_dim, z_pad, hidden, heads, mlp_hidden, n_layers;
  const ldt_decoder_layer* layers;   /* HOST array [n_layers], in module order decoder.0 .. decoder.(n-1) */
  const void* w_out;                 /* bf16 [8, hidden]: Conv1d(hidden -> 3) padded to 8 output rows */
  const float* b_out;                /* f32 [8] */
  /* caller-owned device workspace: MT = batch*32 token rows, MQ = batch*num_points point rows */
  void* ws_e;    /* bf16 [MT, z_pad]    */
  void* ws_x;    /* bf16 [MT, hidden]   */
  void* ws_kv;   /* bf16 [MT, 2*hidden] */
  void* ws_a;    /* bf16 [MQ, hidden]   */
  void* ws_q;    /* bf16 [MQ, hidden]   */
  void* ws_att;  /* bf16 [MQ, hidden]   */
  void* ws_hid;  /* bf16 [MQ, mlp_hidden] */
} ldt_decoder_plan;

/* Compressor.sample's decoder (model/Compressor/Network.py:261-266): eps f32 [batch*32, n_layers*z_dim] (given_eps,
 * token-major); o f32 [MQ, hidden] holds InitialSet's rows on entry (the residual stream, updated in place);
 * points8 f32 [MQ, 8]: columns 0..2 are the generated points. */
int ldt_decoder_forward(const ldt_decoder_plan* plan, const float* eps, float* o, float* points8, void* stream);

typedef struct ldt_sample_args {
  const ldt_score_plan* score;
  int predictor;            /* enum ldt_predictor, 0..3 */
  int num_steps;            /* steps to run from the current *step */
  int use_graph;            /* 1: capture one step on `stream` and launch it num_steps - 1 times; 0: plain launches */
  const float* mod_table;   /* f32 [N, mod_len]: AdaLN rows of every timestep (unconditional sampling: batch-invariant) */
  long long mod_len;
  float* mod_cur;           /* f32 [mod_len] scratch: the current step's row */
  const float* coef;        /* f32 [N, 8] per-step predictor scalars (LDT_SDE_COEF_STRIDE) */
  int* step;                /* device step counter, advanced by one per step */
  const unsigned long long* rng_state;   /* device {seed, offset} of the Philox stream (ldt_sde_step) */
  unsigned long long offset_per_step;
  int rng_grid;
  float* x;                 /* f32 [batch, 32, z_dim] loop state, updated in place */
  float* x_mean;            /* f32 same shape: the denoised mean of the last step run */
  float* params;            /* f32 same shape scratch: the score net's output */
} ldt_sample_args;

/* The reverse-SDE loop of pc_sampling (diffusion_continuous.py:242-249): num_steps x (select AdaLN row, score net,
 * fused predictor update with in-kernel noise, step counter + 1). */
int ldt_sample_loop(const ldt_sample_args* args, void* stream);

/* PNDM pieces (diffusion_continuous.py:260-316).
 * ldt_pndm_transfer: out = x + coef[0] * (coef[1] * x - coef[2] * et)  -- transfer() :264-274; coef is a DEVICE array of
 *   three floats (at_next - at, 1/(sqrt(at)(sqrt(at)+sqrt(at_next))), 1/(sqrt(at)(sqrt((1-at_next)at)+sqrt((1-at)at_next)))).
 * ldt_lincomb4: out = scale * (((c0*a0 + c1*a1) + c2*a2) + c3*a3) -- the Runge-Kutta (:284) and linear multistep (:300)
 *   noise combinations.  Both are rounded like the reference's torch expression. */
int ldt_pndm_transfer(long long numel, const float* x, const float* et, const float* coef, float* out, void* stream);
int ldt_lincomb4(long long numel, float c0, const float* a0, float c1, const float* a1, float c2, const float* a2,
                 float c3, const float* a3, float scale, float* out, void* stream);

/* out[0] = mean over rows of ||x[row, 0:row_len]||_2 (norms [rows] is scratch): the grad / noise norms of the Langevin
 * corrector, diffusion_continuous.py:205-206.  Fixed summation order (deterministic). */
int ldt_batch_mean_norm(int rows, long long row_len, const float* x, float* norms, float* out, void* stream);

/* *step_index += 1 (device side), so a captured step graph can be replayed N times. */
int ldt_advance_step(int* step_index, void* stream);

/* out[0:row_len] = table[*step_index, 0:row_len]  (f32, row_len % 4 == 0).  Used to pull the current
 * step's AdaLN modulation rows out of the per-timestep table ("batch-invariant AdaLN"). */
int ldt_select_row(const float* table, long long row_len, const int* step_index, float* out, void* stream);

/* Per-step conditioning vector of CONDITIONAL sampling (completion / class-conditional):
 *   c[r,:] = table[*step_index,:] (+ extra[r,:]);  silu_out[r,:] = bf16(SiLU(c[r,:]))
 * table [N,D] f32 holds TimeEmbedding(t_i) for every step (batch-invariant), extra [R,D] f32 (or NULL) the per-sample
 * image / label embedding: model/scorenet/score.py:134-135 `c = t_emb + condition[1]`, followed by the SiLU of every
 * adaLN branch (model/layers.py:172).  c_out [R,D] f32 may be NULL.  step_index NULL = row 0. */
int ldt_cond_silu(int R, int D, const float* table, const int* step_index, const float* extra, float* c_out,
                  void* silu_out /* bf16 [R,D] */, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Point-set prologue of the completion path
 * ------------------------------------------------------------------------------------------------ */

/* Furthest point sampling: idx [b,m] i32 = indices of m points of each cloud xyz [b,n,3] f32 (n <= 8192), starting
 * from index 0, each next point the one farthest from the selected set (ties -> lowest index).  Replaces
 * pointnet2_utils.furthest_point_sample (un-vendored dependency; call sites model/Compressor/layers.py:106 and
 * completion_trainer/Latent_SDE_Trainer.py:182-183).  Points with |p|^2 <= min_sq_norm are never selected
 * (pointnet2_ops uses 1e-3; a negative value disables the rule, matching the reference's in-tree
 * model/functional/src/sampling/sampling.cu:86-167). */
int ldt_furthest_point_sample(int b, int n, int m, const float* xyz, float min_sq_norm, int* idx, void* stream);

/* k nearest neighbours: idx [b,s,k] i32 = the k points of xyz [b,n,3] closest to each of centers [b,s,3], in order of
 * increasing squared distance (ties -> lowest index).  Replaces knn_point = square_distance + torch.topk
 * (model/Compressor/layers.py:63-98), which builds a dense [b,s,n] matrix and returns the set in unspecified order. */
int ldt_knn_indices(int b, int n, int s, int k, const float* xyz, const float* centers, int* idx, void* stream);

/* LocalGrouper's normalised group features (model/Compressor/layers.py:300-317), written as the A operand of
 * PreExtraction's first 1x1 convolution: out [b*s*k, ld_out] f32 (columns >= 2d+3 zero), row (b,s,j) =
 *     cat( alpha * ((g - mean) / (std_b + 1e-5)) + beta ,  fea[b, center_idx[b,s], :] )
 * with g = cat(fea[b, group_idx[b,s,j], :], xyz[b, group_idx[b,s,j], :]) (use_xyz=True), mean = the anchor's own
 * cat(fea, xyz) (normalize 2, "anchor"), the mean of g over the k neighbours (1, "center"), or no normalisation at all (0),
 * and std_b the unbiased standard deviation of all (g - mean) of sample b.  xyz [b,n,3], fea [b,n,d] f32, indices i32;
 * alpha / beta [d+3]; partial = 2*b*s doubles of scratch.  Deterministic (fixed-order sums, no atomics). */
int ldt_group_features(int b, int n, int s, int k, int d, const float* xyz, const float* fea, const int* center_idx,
                       const int* group_idx, int normalize, const float* alpha, const float* beta, double* partial, float* out,
                       int ld_out, void* stream);

/* out[g, c] = max over j < k of x[g*k + j, c]  (f32; x [groups*k, ldx], out [groups, ldo]): the max over a group's neighbours
 * that ends PreExtraction (model/Compressor/layers.py:189-190) and MiniPointnet (model/Compressor/Network.py:97). */
int ldt_group_max(int groups, int k, int c, const float* x, int ldx, float* out, int ldo, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Device properties needed by the host mirror
 * ------------------------------------------------------------------------------------------------ */
int ldt_device_sm_count(void);

/* Programmatic dependent launch for the per-step kernels (default OFF; ldt_set_pdl(1) or the environment variable
 * LDT_PDL=1 switches it on).  With it, consecutive kernels on one stream overlap launch latency and prologue with
 * the previous kernel's tail; results are unchanged.  Measured on B200 inside the replayed step graph it is neutral
 * (44.98 vs 45.20 clouds/s): graph-internal kernel->kernel gaps are already ~1 us.  Takes effect for subsequent
 * launches (and graph captures). */
int ldt_set_pdl(int enable);

#ifdef __cplusplus
}
#endif
#endif /* LDT_B200_H_ */
