// Whole-path entry points of the C ABI (SURVEY.md 8b export list): the score-net token pass and the N-step reverse-SDE
// loop as ONE call each, for hosts that do not want to orchestrate ~150 kernel launches per step themselves.
//   ldt_score_forward  = Score.forward's token path (model/scorenet/score.py:136-150 -> model/layers.py:202-229,240-245)
//   ldt_decoder_forward = Compressor.sample's decoder (model/Compressor/Network.py:261-266 -> DecoderBlock.forward :80-83)
//   ldt_sample_loop    = pc_sampling's loop (diffusion/diffusion_continuous.py:242-249) with the Ancestral / ReverseDiffusion /
//                        EulerMaruyama / DDIM predictor and score = -params / sqrt(var) (trainer/Latent_SDE_Trainer.py:57-61)
// Pure orchestration over the kernels of this library (same launches, same order, same arguments as ldt_b200/score.py::
// run_tokens and sampler.py::StepGraph, so results are bit-identical to the Python-orchestrated path); no allocation: the
// caller owns weights, workspace and state.  The loop captures ONE step into a CUDA graph on the caller's stream and
// launches it N times (eager launches when the stream cannot capture, e.g. the legacy default stream).
#include "common.cuh"
#include "ldt_b200.h"

using namespace ldt;

static int score_forward_impl(const ldt_score_plan& p, const float* x_tokens, const float* mod, long long mod_stride, float* out,
                              void* stream) {
  const int M = p.batch * p.tokens, Hd = p.hidden;
  int rc = ldt_cast_pad_bf16(M, p.z_dim, x_tokens, p.z_dim, p.ws_xa, p.z_pad, stream);
  if (rc) return rc;
  ldt_gemm_args g = {};
  g.M = M; g.rows_per_gate = p.tokens; g.gate_stride = mod_stride;
  // ln_in: Conv1d(z_dim -> hidden)
  g.N = Hd; g.K = p.z_pad; g.A = p.ws_xa; g.lda = p.z_pad; g.W = p.w_in; g.ldw = p.z_pad; g.bias = p.b_in;
  g.out = p.ws_h; g.ldo = Hd; g.epilogue = LDT_EPI_BIAS_F32; g.resid = nullptr; g.gate = nullptr;
  rc = ldt_gemm_bf16(&g, stream);
  if (rc) return rc;
  for (int i = 0; i < p.num_blocks; ++i) {
    const ldt_score_block& b = p.blocks[i];
    const float* m = mod + static_cast<long long>(i) * 6 * Hd;   // shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp
    rc = ldt_layernorm_mod_bf16(M, Hd, p.ws_h, m, m + Hd, mod_stride, p.tokens, nullptr, nullptr, 1e-6f, p.ws_a, stream);
    if (rc) return rc;
    rc = ldt_qkv_attention_bf16(p.batch, p.heads, Hd, p.ws_a, Hd, b.w_qkv_packed, Hd, b.b_qkv_packed, p.ws_att, stream);
    if (rc) return rc;
    g.N = Hd; g.K = Hd; g.A = p.ws_att; g.lda = Hd; g.W = b.w_o; g.ldw = Hd; g.bias = b.b_o; g.out = p.ws_h; g.ldo = Hd;
    g.epilogue = LDT_EPI_GATE_RESID_F32; g.resid = p.ws_h; g.gate = m + 2 * Hd;
    rc = ldt_gemm_bf16(&g, stream);
    if (rc) return rc;
    rc = ldt_layernorm_mod_bf16(M, Hd, p.ws_h, m + 3 * Hd, m + 4 * Hd, mod_stride, p.tokens, nullptr, nullptr, 1e-6f, p.ws_a, stream);
    if (rc) return rc;
    g.N = p.mlp_hidden; g.K = Hd; g.A = p.ws_a; g.lda = Hd; g.W = b.w_fc1; g.ldw = Hd; g.bias = b.b_fc1; g.out = p.ws_hid;
    g.ldo = p.mlp_hidden; g.epilogue = LDT_EPI_BIAS_GELU_BF16; g.resid = nullptr; g.gate = nullptr;
    rc = ldt_gemm_bf16(&g, stream);
    if (rc) return rc;
    g.N = Hd; g.K = p.mlp_hidden; g.A = p.ws_hid; g.lda = p.mlp_hidden; g.W = b.w_fc2; g.ldw = p.mlp_hidden; g.bias = b.b_fc2;
    g.out = p.ws_h; g.ldo = Hd; g.epilogue = LDT_EPI_GATE_RESID_F32; g.resid = p.ws_h; g.gate = m + 5 * Hd;
    rc = ldt_gemm_bf16(&g, stream);
    if (rc) return rc;
  }
  const float* m = mod + static_cast<long long>(p.num_blocks) * 6 * Hd;   // FinalLayer: (shift, scale)
  rc = ldt_layernorm_mod_bf16(M, Hd, p.ws_h, m, m + Hd, mod_stride, p.tokens, nullptr, nullptr, 1e-6f, p.ws_a, stream);
  if (rc) return rc;
  g.N = p.z_dim; g.K = Hd; g.A = p.ws_a; g.lda = Hd; g.W = p.w_out; g.ldw = Hd; g.bias = p.b_out; g.out = out; g.ldo = p.z_dim;
  g.epilogue = LDT_EPI_BIAS_F32; g.resid = nullptr; g.gate = nullptr;
  return ldt_gemm_bf16(&g, stream);
}

static int check_plan(const char* who, const ldt_score_plan* p) {
  LDT_REQUIRE(p != nullptr, LDT_ERR_INVALID, "%s: null plan", who);
  LDT_REQUIRE(p->batch > 0 && p->tokens == 32 && p->hidden > 0 && p->heads > 0 && p->num_blocks >= 0 && p->z_dim > 0, LDT_ERR_INVALID,
              "%s: bad shape batch=%d tokens=%d hidden=%d heads=%d blocks=%d z_dim=%d", who, p->batch, p->tokens, p->hidden, p->heads,
              p->num_blocks, p->z_dim);
  LDT_REQUIRE(p->hidden == 64 * p->heads, LDT_ERR_UNSUPPORTED, "%s: the fused projection + attention kernel needs head dim 64", who);
  LDT_REQUIRE(p->z_pad % 64 == 0 && p->z_pad >= p->z_dim && p->mlp_hidden % 64 == 0 && p->z_dim % 8 == 0, LDT_ERR_INVALID,
              "%s: z_pad=%d mlp_hidden=%d z_dim=%d: K extents must be multiples of 64, z_dim of 8", who, p->z_pad, p->mlp_hidden, p->z_dim);
  LDT_REQUIRE(p->w_in && p->w_out && p->ws_xa && p->ws_h && p->ws_a && p->ws_att && p->ws_hid && (p->num_blocks == 0 || p->blocks),
              LDT_ERR_INVALID, "%s: null weight / workspace pointer", who);
  return LDT_OK;
}

extern "C" int ldt_score_forward(const ldt_score_plan* plan, const float* x_tokens, const float* mod, long long mod_stride,
                                 float* out, void* stream) {
  int rc = check_plan("ldt_score_forward", plan);
  if (rc) return rc;
  LDT_REQUIRE(x_tokens && mod && out, LDT_ERR_INVALID, "ldt_score_forward: null pointer");
  LDT_REQUIRE(mod_stride % 4 == 0, LDT_ERR_INVALID, "ldt_score_forward: mod_stride must be a multiple of 4");
  return score_forward_impl(*plan, x_tokens, mod, mod_stride, out, stream);
}

static int one_step(const ldt_sample_args& a, void* stream) {
  int rc = ldt_select_row(a.mod_table, a.mod_len, a.step, a.mod_cur, stream);
  if (rc) return rc;
  rc = score_forward_impl(*a.score, a.x, a.mod_cur, 0, a.params, stream);
  if (rc) return rc;
  const long long numel = static_cast<long long>(a.score->batch) * a.score->tokens * a.score->z_dim;
  rc = ldt_sde_step(a.predictor, numel, a.x, a.params, nullptr, a.coef, a.step, 0ull, 0ull, a.offset_per_step, a.rng_state, a.rng_grid,
                    a.x, a.x_mean, stream);
  if (rc) return rc;
  return ldt_advance_step(a.step, stream);
}

extern "C" int ldt_sample_loop(const ldt_sample_args* args, void* stream) {
  LDT_REQUIRE(args != nullptr, LDT_ERR_INVALID, "ldt_sample_loop: null args");
  const ldt_sample_args& a = *args;
  int rc = check_plan("ldt_sample_loop", a.score);
  if (rc) return rc;
  LDT_REQUIRE(a.num_steps >= 0 && a.mod_table && a.mod_cur && a.coef && a.step && a.rng_state && a.x && a.x_mean && a.params,
              LDT_ERR_INVALID, "ldt_sample_loop: null pointer / negative step count");
  LDT_REQUIRE(a.mod_len == static_cast<long long>(a.score->num_blocks) * 6 * a.score->hidden + 2 * a.score->hidden, LDT_ERR_INVALID,
              "ldt_sample_loop: mod_len=%lld does not match the plan (6*hidden per block + 2*hidden)", a.mod_len);
  LDT_REQUIRE(a.predictor >= LDT_PRED_ANCESTRAL && a.predictor <= LDT_PRED_DDIM, LDT_ERR_INVALID, "ldt_sample_loop: unknown predictor %d",
              a.predictor);
  if (a.num_steps == 0) return LDT_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  const bool can_capture = (s != nullptr) && cudaStreamIsCapturing(s, &st) == cudaSuccess && st == cudaStreamCaptureStatusNone &&
                           a.use_graph != 0;
  if (!can_capture) {   // legacy default stream, a caller that is itself capturing, or use_graph == 0: plain launches
    (void)cudaGetLastError();
    for (int i = 0; i < a.num_steps; ++i) {
      rc = one_step(a, stream);
      if (rc) return rc;
    }
    return LDT_OK;
  }
  rc = one_step(a, stream);   // step 0 eagerly: first-use set-up (kernel attributes) must not happen under capture
  if (rc || a.num_steps == 1) return rc;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  LDT_CUDA_OK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  rc = one_step(a, stream);
  cudaError_t e = cudaStreamEndCapture(s, &graph);
  if (rc != 0 || e != cudaSuccess) {
    if (graph) cudaGraphDestroy(graph);
    if (rc == 0) rc = check_cuda(e, "cudaStreamEndCapture", __FILE__, __LINE__);
    return rc;
  }
  e = cudaGraphInstantiate(&exec, graph, 0);
  if (e == cudaSuccess)
    for (int i = 1; i < a.num_steps && e == cudaSuccess; ++i) e = cudaGraphLaunch(exec, s);
  if (exec) cudaGraphExecDestroy(exec);   // deferred by the runtime until the launches have run
  cudaGraphDestroy(graph);
  return check_cuda(e, "cudaGraphInstantiate / cudaGraphLaunch", __FILE__, __LINE__);
}

extern "C" int ldt_decoder_forward(const ldt_decoder_plan* plan, const float* eps, float* o, float* points8, void* stream) {
  LDT_REQUIRE(plan != nullptr, LDT_ERR_INVALID, "ldt_decoder_forward: null plan");
  const ldt_decoder_plan& p = *plan;
  LDT_REQUIRE(p.batch > 0 && p.num_points > 0 && p.hidden > 0 && p.heads > 0 && p.n_layers >= 0 && p.z_dim > 0 && p.mlp_hidden > 0,
              LDT_ERR_INVALID, "ldt_decoder_forward: bad shape batch=%d points=%d hidden=%d heads=%d layers=%d z_dim=%d", p.batch,
              p.num_points, p.hidden, p.heads, p.n_layers, p.z_dim);
  LDT_REQUIRE(p.hidden % 128 == 0 && p.z_pad % 64 == 0 && p.z_pad >= p.z_dim && p.mlp_hidden % 64 == 0, LDT_ERR_INVALID,
              "ldt_decoder_forward: hidden %% 128, z_pad %% 64, mlp_hidden %% 64 required");
  LDT_REQUIRE(eps && o && points8 && p.w_out && p.ws_e && p.ws_x && p.ws_kv && p.ws_a && p.ws_q && p.ws_att && p.ws_hid &&
                  (p.n_layers == 0 || p.layers),
              LDT_ERR_INVALID, "ldt_decoder_forward: null pointer");
  const int H = p.hidden, MQ = p.batch * p.num_points, MT = p.batch * 32;
  const int ld_eps = p.n_layers * p.z_dim;
  int rc;
  ldt_gemm_args g = {};
  g.rows_per_gate = 1;
  for (int idx = 0; idx < p.n_layers; ++idx) {
    const ldt_decoder_layer& L = p.layers[p.n_layers - 1 - idx];   // reversed(self.decoder), Network.py:263
    rc = ldt_cast_pad_bf16(MT, p.z_dim, eps + idx * p.z_dim, ld_eps, p.ws_e, p.z_pad, stream);   // torch.split(...)[idx], :262
    if (rc) return rc;
    g.resid = nullptr; g.gate = nullptr;
    g.M = MT; g.N = H; g.K = p.z_pad; g.A = p.ws_e; g.lda = p.z_pad; g.W = L.w_ln; g.ldw = p.z_pad; g.bias = L.b_ln; g.out = p.ws_x;
    g.ldo = H; g.epilogue = LDT_EPI_BIAS_BF16;                                     // x = self.ln(eps)            :81
    if ((rc = ldt_gemm_bf16(&g, stream))) return rc;
    g.N = 2 * H; g.K = H; g.A = p.ws_x; g.lda = H; g.W = L.w_kv; g.ldw = H; g.bias = L.b_kv; g.out = p.ws_kv; g.ldo = 2 * H;
    if ((rc = ldt_gemm_bf16(&g, stream))) return rc;                               // kv = fc_kv(x)       layers.py:187
    if ((rc = ldt_layernorm_mod_bf16(MQ, H, o, nullptr, nullptr, 0, 1, L.norm1_w, L.norm1_b, 1e-6f, p.ws_a, stream))) return rc;
    g.M = MQ; g.N = H; g.K = H; g.A = p.ws_a; g.lda = H; g.W = L.w_q; g.ldw = H; g.bias = L.b_q; g.out = p.ws_q; g.ldo = H;
    if ((rc = ldt_gemm_bf16(&g, stream))) return rc;                               // q = fc_q(norm1(o))
    rc = ldt_attention_nk32(p.batch, p.heads, p.num_points, H / p.heads, p.ws_q, H, p.ws_kv,
                            static_cast<const char*>(p.ws_kv) + 2 * static_cast<size_t>(H), 2 * H, p.ws_att, stream);
    if (rc) return rc;
    g.A = p.ws_att; g.W = L.w_o; g.bias = L.b_o; g.out = o; g.epilogue = LDT_EPI_GATE_RESID_F32; g.resid = o;
    if ((rc = ldt_gemm_bf16(&g, stream))) return rc;                               // o = o + fc_o(att)           :225
    if ((rc = ldt_layernorm_mod_bf16(MQ, H, o, nullptr, nullptr, 0, 1, L.norm2_w, L.norm2_b, 1e-6f, p.ws_a, stream))) return rc;
    g.N = p.mlp_hidden; g.A = p.ws_a; g.W = L.w_fc1; g.bias = L.b_fc1; g.out = p.ws_hid; g.ldo = p.mlp_hidden;
    g.epilogue = LDT_EPI_BIAS_GELU_BF16; g.resid = nullptr;
    if ((rc = ldt_gemm_bf16(&g, stream))) return rc;
    g.N = H; g.K = p.mlp_hidden; g.A = p.ws_hid; g.lda = p.mlp_hidden; g.W = L.w_fc2; g.ldw = p.mlp_hidden; g.bias = L.b_fc2;
    g.out = o; g.ldo = H; g.epilogue = LDT_EPI_GATE_RESID_F32; g.resid = o;
    if ((rc = ldt_gemm_bf16(&g, stream))) return rc;                               // o = o + mlp(norm2(o))       :226
  }
  if ((rc = ldt_cast_pad_bf16(MQ, H, o, H, p.ws_a, H, stream))) return rc;
  g.M = MQ; g.N = 8; g.K = H; g.A = p.ws_a; g.lda = H; g.W = p.w_out; g.ldw = H; g.bias = p.b_out; g.out = points8; g.ldo = 8;
  g.epilogue = LDT_EPI_BIAS_F32; g.resid = nullptr; g.gate = nullptr;
  return ldt_gemm_bf16(&g, stream);                                                 // self.output(o)              :266
}
