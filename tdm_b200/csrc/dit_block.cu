// One PixArt DiT block forward / backward as a single C call each (see include/tdm_b200.h).
// Pure sequencing of this library's kernels on the caller's stream: LN-modulate, tcgen05 GEMMs with
// fused epilogues (bias, GELU-tanh, gate*y + residual, LoRA concat-K), tcgen05 attention.
// Semantics: diffusers BasicTransformerBlock, norm_type="ada_norm_single" (SURVEY.md Appendix A).
#include <math.h>
#include <string.h>
#include "common.cuh"

namespace tdm {

static inline int64_t al256(int64_t b) { return (b + 255) & ~int64_t(255); }

struct BlockWs {  // carved from the caller's workspace; identical layout for save / scratch use
  float *mod, *mean1, *rstd1, *mean2, *rstd2, *lse1, *lse2, *x1, *x2;
  bf16 *h1, *qkv, *a1, *y1, *x1b, *q2, *kv2, *a2, *h2, *u, *g, *y3;
  bf16 *t_qkv, *t_o, *t_q2, *t_kv2, *t_o2, *t_ff1, *t_ff2;
  int64_t bytes;
};

static BlockWs carve_ws(const TdmDitShape* s, void* base) {
  BlockWs w;
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  int64_t off = 0;
  const int64_t M = (int64_t)s->B * s->N, D = s->D, BL = (int64_t)s->B * s->L, r = s->lora_r;
  auto f32 = [&](int64_t n) { float* q = reinterpret_cast<float*>(p + off); off += al256(n * 4); return q; };
  auto b16 = [&](int64_t n) { bf16* q = reinterpret_cast<bf16*>(p + off); off += al256(n * 2); return q; };
  w.mod = f32((int64_t)s->B * 6 * D);
  w.mean1 = f32(M); w.rstd1 = f32(M); w.mean2 = f32(M); w.rstd2 = f32(M);
  w.lse1 = f32((int64_t)s->B * s->H * s->N); w.lse2 = f32((int64_t)s->B * s->H * s->N);
  w.x1 = f32(M * D); w.x2 = f32(M * D);
  w.h1 = b16(M * D); w.qkv = b16(M * 3 * D); w.a1 = b16(M * D); w.y1 = b16(M * D); w.x1b = b16(M * D);
  w.q2 = b16(M * D); w.kv2 = b16(BL * 2 * D); w.a2 = b16(M * D); w.h2 = b16(M * D);
  w.u = b16(M * s->Dff); w.g = b16(M * s->Dff); w.y3 = b16(M * D);
  w.t_qkv = b16(M * 3 * r); w.t_o = b16(M * r); w.t_q2 = b16(M * r); w.t_kv2 = b16(BL * 2 * r);
  w.t_o2 = b16(M * r); w.t_ff1 = b16(M * r); w.t_ff2 = b16(M * r);
  w.bytes = off;
  return w;
}

struct BwdScratch {
  float *dmod, *dx2, *dq_accum, *delta;
  bf16 *dyA, *dyB, *dyC, *du, *dqkv, *dkv2, *dt;
  int64_t bytes;
};

static BwdScratch carve_scratch(const TdmDitShape* s, void* base) {
  BwdScratch w;
  uint8_t* p = reinterpret_cast<uint8_t*>(base);
  int64_t off = 0;
  const int64_t M = (int64_t)s->B * s->N, D = s->D, BL = (int64_t)s->B * s->L, r = s->lora_r;
  auto f32 = [&](int64_t n) { float* q = reinterpret_cast<float*>(p + off); off += al256(n * 4); return q; };
  auto b16 = [&](int64_t n) { bf16* q = reinterpret_cast<bf16*>(p + off); off += al256(n * 2); return q; };
  w.dmod = f32((int64_t)s->B * 6 * D);
  w.dx2 = f32(M * D); w.dq_accum = f32(M * D); w.delta = f32((int64_t)s->B * s->H * s->N);
  w.dyA = b16(M * D); w.dyB = b16(M * D); w.dyC = b16(M * D);
  w.du = b16(M * s->Dff); w.dqkv = b16(M * 3 * D); w.dkv2 = b16(BL * 2 * D);
  w.dt = b16((M > BL ? M : BL) * 3 * (r > 0 ? r : 8));
  w.bytes = off;
  return w;
}

// ---- GEMM descriptor helpers -------------------------------------------------------------------
static TdmGemm G(int M, int N, int K, const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn) {
  TdmGemm g;
  memset(&g, 0, sizeof(g));
  g.M = M; g.N = N; g.K = K; g.A = A; g.lda = lda; g.a_mn_major = a_mn; g.B = B; g.ldb = ldb; g.b_mn_major = b_mn;
  g.alpha = 1.f;
  return g;
}
static void pair2(TdmGemm& g, const void* A2, int64_t lda2, const void* B2, int64_t ldb2, int K2, int seg_n = 0,
                  int seg_k2 = 0) {
  g.A2 = A2; g.lda2 = lda2; g.B2 = B2; g.ldb2 = ldb2; g.K2 = K2; g.seg_n = seg_n; g.seg_k2 = seg_k2;
}
#define RUN(expr)            \
  do {                       \
    int _rc = (expr);        \
    if (_rc) return _rc;     \
  } while (0)

// y[M,N] (fp32 +=) = dy^T x : wgrad with both operands token-major (MN-major for the tensor core)
static int wgrad(const bf16* dy, int64_t ld_dy, int n_out, const bf16* x, int64_t ld_x, int k_in, int tokens, float* dw,
                 int64_t ld_dw, void* st) {
  TdmGemm g = G(n_out, k_in, tokens, dy, ld_dy, 1, x, ld_x, 1);
  g.res = dw; g.ld_res = ld_dw; g.out_f32 = dw; g.ld_out_f32 = ld_dw;
  return tdm_gemm_bf16(&g, st);
}

// LoRA backward pieces for one linear (or one q/k/v segment): dt = s * dy B ; dB += dy^T t ; (dA is done by the caller)
static int lora_dt_db(const TdmDitShape* s, const bf16* dy, int64_t ld_dy, int n_out, int tokens, const void* lb,
                      const bf16* t, int64_t ld_t, bf16* dt, int64_t ld_dt, float* g_lb, void* st) {
  const int r = s->lora_r;
  TdmGemm g = G(tokens, r, n_out, dy, ld_dy, 0, lb, r, 1);  // B[N_out, r] read as [K=N_out rows, N=r]
  g.alpha = s->lora_scale; g.out_bf16 = dt; g.ld_out_bf16 = ld_dt;
  RUN(tdm_gemm_bf16(&g, st));
  if (g_lb) RUN(wgrad(dy, ld_dy, n_out, t, ld_t, r, tokens, g_lb, r, st));  // t already carries the scale
  return TDM_OK;
}

}  // namespace tdm

using namespace tdm;

extern "C" int64_t tdm_dit_block_ws_bytes(const TdmDitShape* s, int save) {
  (void)save;
  if (!s) return 0;
  return carve_ws(s, nullptr).bytes;
}
extern "C" int64_t tdm_dit_block_bwd_scratch_bytes(const TdmDitShape* s) {
  if (!s) return 0;
  return carve_scratch(s, nullptr).bytes;
}

static int check_shape(const TdmDitShape* s, const char* who) {
  TDM_REQUIRE(s, "%s: null shape", who);
  TDM_REQUIRE(s->B > 0 && s->N > 0 && s->L > 0 && s->H > 0 && s->D % s->H == 0, "%s: bad shape", who);
  TDM_REQUIRE(s->D % 128 == 0 && s->Dff % 8 == 0, "%s: D must be a multiple of 128", who);
  TDM_REQUIRE(s->lora_r == 0 || s->lora_r == 64, "%s: lora_r must be 0 or 64 (got %d)", who, s->lora_r);
  return TDM_OK;
}

extern "C" int tdm_dit_block_fwd(const TdmDitShape* s, const TdmBlockWeights* w, const float* x, const float* tmod,
                                 const void* ctx, const float* ctx_bias, float* x_out, void* ws_, int save, void* st) {
  RUN(check_shape(s, "tdm_dit_block_fwd"));
  TDM_REQUIRE(w && x && tmod && ctx && x_out && ws_, "tdm_dit_block_fwd: null argument");
  TDM_REQUIRE(((uintptr_t)ws_ & 255) == 0, "tdm_dit_block_fwd: workspace must be 256-byte aligned");
  BlockWs a = carve_ws(s, ws_);
  const int B = s->B, N = s->N, D = s->D, H = s->H, d = D / H, L = s->L, Dff = s->Dff, r = s->lora_r;
  const int M = B * N, BL = B * L;
  const float ls = s->lora_scale;
  const int64_t ms = 6 * (int64_t)D;  // stride between samples in mod
  const bool keep_pre = save && s->train_base;

  RUN(tdm_modulation(w->table, tmod, 1, a.mod, B, 6, D, st));
  // ---- self attention ----
  RUN(tdm_ln_modulate_fwd(x, a.mod + 0 * D, a.mod + 1 * D, ms, N, a.h1, a.mean1, a.rstd1, M, D, s->eps, st));
  {
    TdmGemm g = G(M, 3 * D, D, a.h1, D, 0, w->w_qkv, D, 0);
    if (r && w->lora_a_qkv) {
      TdmGemm t = G(M, 3 * r, D, a.h1, D, 0, w->lora_a_qkv, D, 0);
      t.alpha = ls; t.out_bf16 = a.t_qkv; t.ld_out_bf16 = 3 * r;
      RUN(tdm_gemm_bf16(&t, st));
      pair2(g, a.t_qkv, 3 * r, w->lora_b_qkv, r, r, D, r);
    }
    g.bias = w->b_qkv; g.out_bf16 = a.qkv; g.ld_out_bf16 = 3 * D;
    RUN(tdm_gemm_bf16(&g, st));
  }
  {
    TdmAttn t;
    memset(&t, 0, sizeof(t));
    t.B = B; t.H = H; t.Sq = N; t.Sk = N; t.d = d; t.scale = 1.f / sqrtf((float)d);
    t.Q = a.qkv; t.K = a.qkv + D; t.V = a.qkv + 2 * D;
    t.q_stride_b = t.k_stride_b = t.v_stride_b = (int64_t)N * 3 * D;
    t.q_stride_s = t.k_stride_s = t.v_stride_s = 3 * D;
    t.q_stride_h = t.k_stride_h = t.v_stride_h = d;
    t.O = a.a1; t.o_stride_b = (int64_t)N * D; t.o_stride_s = D; t.o_stride_h = d;
    t.lse = a.lse1;
    RUN(tdm_attn_fwd(&t, st));
  }
  {
    TdmGemm g = G(M, D, D, a.a1, D, 0, w->w_o, D, 0);
    if (r && w->lora_a_o) {
      TdmGemm t = G(M, r, D, a.a1, D, 0, w->lora_a_o, D, 0);
      t.alpha = ls; t.out_bf16 = a.t_o; t.ld_out_bf16 = r;
      RUN(tdm_gemm_bf16(&t, st));
      pair2(g, a.t_o, r, w->lora_b_o, r, r);
    }
    g.bias = w->b_o; g.pre_out = keep_pre ? a.y1 : nullptr; g.ld_pre = D;
    g.gate = a.mod + 2 * D; g.rows_per_gate = N; g.gate_stride = ms;
    g.res = x; g.ld_res = D; g.out_f32 = a.x1; g.ld_out_f32 = D; g.out_bf16 = a.x1b; g.ld_out_bf16 = D;
    RUN(tdm_gemm_bf16(&g, st));
  }
  // ---- cross attention (on the un-normalised residual stream, no gate) ----
  {
    TdmGemm g = G(M, D, D, a.x1b, D, 0, w->w_q2, D, 0);
    if (r && w->lora_a_q2) {
      TdmGemm t = G(M, r, D, a.x1b, D, 0, w->lora_a_q2, D, 0);
      t.alpha = ls; t.out_bf16 = a.t_q2; t.ld_out_bf16 = r;
      RUN(tdm_gemm_bf16(&t, st));
      pair2(g, a.t_q2, r, w->lora_b_q2, r, r);
    }
    g.bias = w->b_q2; g.out_bf16 = a.q2; g.ld_out_bf16 = D;
    RUN(tdm_gemm_bf16(&g, st));
  }
  {
    TdmGemm g = G(BL, 2 * D, D, ctx, D, 0, w->w_kv2, D, 0);
    if (r && w->lora_a_kv2) {
      TdmGemm t = G(BL, 2 * r, D, ctx, D, 0, w->lora_a_kv2, D, 0);
      t.alpha = ls; t.out_bf16 = a.t_kv2; t.ld_out_bf16 = 2 * r;
      RUN(tdm_gemm_bf16(&t, st));
      pair2(g, a.t_kv2, 2 * r, w->lora_b_kv2, r, r, D, r);
    }
    g.bias = w->b_kv2; g.out_bf16 = a.kv2; g.ld_out_bf16 = 2 * D;
    RUN(tdm_gemm_bf16(&g, st));
  }
  {
    TdmAttn t;
    memset(&t, 0, sizeof(t));
    t.B = B; t.H = H; t.Sq = N; t.Sk = L; t.d = d; t.scale = 1.f / sqrtf((float)d);
    t.Q = a.q2; t.q_stride_b = (int64_t)N * D; t.q_stride_s = D; t.q_stride_h = d;
    t.K = a.kv2; t.V = a.kv2 + D;
    t.k_stride_b = t.v_stride_b = (int64_t)L * 2 * D; t.k_stride_s = t.v_stride_s = 2 * D; t.k_stride_h = t.v_stride_h = d;
    t.bias = ctx_bias;
    t.O = a.a2; t.o_stride_b = (int64_t)N * D; t.o_stride_s = D; t.o_stride_h = d;
    t.lse = a.lse2;
    RUN(tdm_attn_fwd(&t, st));
  }
  {
    TdmGemm g = G(M, D, D, a.a2, D, 0, w->w_o2, D, 0);
    if (r && w->lora_a_o2) {
      TdmGemm t = G(M, r, D, a.a2, D, 0, w->lora_a_o2, D, 0);
      t.alpha = ls; t.out_bf16 = a.t_o2; t.ld_out_bf16 = r;
      RUN(tdm_gemm_bf16(&t, st));
      pair2(g, a.t_o2, r, w->lora_b_o2, r, r);
    }
    g.bias = w->b_o2; g.res = a.x1; g.ld_res = D; g.out_f32 = a.x2; g.ld_out_f32 = D;
    RUN(tdm_gemm_bf16(&g, st));
  }
  // ---- feed forward ----
  RUN(tdm_ln_modulate_fwd(a.x2, a.mod + 3 * D, a.mod + 4 * D, ms, N, a.h2, a.mean2, a.rstd2, M, D, s->eps, st));
  {
    TdmGemm g = G(M, Dff, D, a.h2, D, 0, w->w_ff1, D, 0);
    if (r && w->lora_a_ff1) {
      TdmGemm t = G(M, r, D, a.h2, D, 0, w->lora_a_ff1, D, 0);
      t.alpha = ls; t.out_bf16 = a.t_ff1; t.ld_out_bf16 = r;
      RUN(tdm_gemm_bf16(&t, st));
      pair2(g, a.t_ff1, r, w->lora_b_ff1, r, r);
    }
    g.bias = w->b_ff1; g.act = TDM_ACT_GELU_TANH; g.pre_out = save ? a.u : nullptr; g.ld_pre = Dff;
    g.out_bf16 = a.g; g.ld_out_bf16 = Dff;
    RUN(tdm_gemm_bf16(&g, st));
  }
  {
    TdmGemm g = G(M, D, Dff, a.g, Dff, 0, w->w_ff2, Dff, 0);
    if (r && w->lora_a_ff2) {
      TdmGemm t = G(M, r, Dff, a.g, Dff, 0, w->lora_a_ff2, Dff, 0);
      t.alpha = ls; t.out_bf16 = a.t_ff2; t.ld_out_bf16 = r;
      RUN(tdm_gemm_bf16(&t, st));
      pair2(g, a.t_ff2, r, w->lora_b_ff2, r, r);
    }
    g.bias = w->b_ff2; g.pre_out = keep_pre ? a.y3 : nullptr; g.ld_pre = D;
    g.gate = a.mod + 5 * D; g.rows_per_gate = N; g.gate_stride = ms;
    g.res = a.x2; g.ld_res = D; g.out_f32 = x_out; g.ld_out_f32 = D;
    RUN(tdm_gemm_bf16(&g, st));
  }
  return TDM_OK;
}

// One linear's backward: optional base wgrad + bias grad, optional LoRA grads, then the input gradient
//   dx = dy W (+ dt A) with a caller-configured epilogue (g_dx comes pre-filled with outputs/epilogue).
struct LinBwd {
  const bf16* dy; int64_t ld_dy; int n_out;   // upstream gradient [tokens, n_out]
  const bf16* x; int64_t ld_x; int k_in;      // saved input [tokens, k_in]
  int tokens;
  const void* w;                              // [n_out, k_in] bf16
  float* g_w; float* g_b;                     // base grads (nullable)
  const void* la; const void* lb;             // LoRA A [r, k_in], B [n_out, r] (nullable)
  const bf16* t; int64_t ld_t;                // saved s * x A^T
  float* g_la; float* g_lb;
};

static int linear_bwd(const TdmDitShape* s, const LinBwd& l, bf16* dt_buf, TdmGemm* g_dx, void* st) {
  const int r = s->lora_r;
  if (l.g_w) RUN(wgrad(l.dy, l.ld_dy, l.n_out, l.x, l.ld_x, l.k_in, l.tokens, l.g_w, l.k_in, st));
  if (l.g_b) RUN(tdm_colsum_bf16(l.dy, l.ld_dy, l.tokens, l.n_out, l.g_b, st));
  const bool lora = r && l.la;
  if (lora) {
    RUN(lora_dt_db(s, l.dy, l.ld_dy, l.n_out, l.tokens, l.lb, l.t, l.ld_t, dt_buf, r, l.g_lb, st));
    if (l.g_la) RUN(wgrad(dt_buf, r, r, l.x, l.ld_x, l.k_in, l.tokens, l.g_la, l.k_in, st));
  }
  if (g_dx) {
    g_dx->M = l.tokens; g_dx->N = l.k_in; g_dx->K = l.n_out;
    g_dx->A = l.dy; g_dx->lda = l.ld_dy; g_dx->a_mn_major = 0;
    g_dx->B = l.w; g_dx->ldb = l.k_in; g_dx->b_mn_major = 1;
    if (lora) pair2(*g_dx, dt_buf, r, l.la, l.k_in, r);
    RUN(tdm_gemm_bf16(g_dx, st));
  }
  return TDM_OK;
}

// Stacked (q/k/v or k/v) linear: per-segment LoRA dt / dB, one stacked dA, one concat-K dgrad.
static int stacked_linear_bwd(const TdmDitShape* s, int nseg, const bf16* dy, int seg_out, const bf16* x, int k_in,
                              int tokens, const void* w, float* g_w, float* g_b, const void* la, const void* lb,
                              const bf16* t, float* g_la, float* g_lb, bf16* dt_buf, TdmGemm* g_dx, void* st) {
  const int r = s->lora_r;
  const int n_out = nseg * seg_out;
  if (g_w) RUN(wgrad(dy, n_out, n_out, x, k_in, k_in, tokens, g_w, k_in, st));
  if (g_b) RUN(tdm_colsum_bf16(dy, n_out, tokens, n_out, g_b, st));
  const bool lora = r && la;
  if (lora) {
    if (seg_out % 128 == 0) {
      // every segment in ONE launch each: dt[:, s] = ls * dy[:, s] B_s (column segments shift the contraction window)
      // and dB_s += dy[:, s]^T t[:, s] (row segments shift B's columns) -- the stacked [nseg * seg_out, r] matrices of
      // lora_B and its gradient are contiguous
      TdmGemm g = G(tokens, nseg * r, seg_out, dy, n_out, 0, lb, r, 1);
      g.seg_n = r; g.seg_k = seg_out; g.tile_n = 64; g.cta_pair = 1;
      g.alpha = s->lora_scale; g.out_bf16 = dt_buf; g.ld_out_bf16 = nseg * r;
      RUN(tdm_gemm_bf16(&g, st));
      if (g_lb) {
        TdmGemm w = G(n_out, r, tokens, dy, n_out, 1, t, nseg * r, 1);
        w.seg_m = seg_out; w.seg_nb = r; w.tile_n = 64; w.cta_pair = 1;
        w.res = g_lb; w.ld_res = r; w.out_f32 = g_lb; w.ld_out_f32 = r;
        RUN(tdm_gemm_bf16(&w, st));
      }
    } else {
      for (int j = 0; j < nseg; ++j) {
        const bf16* lbj = reinterpret_cast<const bf16*>(lb) + (int64_t)j * seg_out * r;
        RUN(lora_dt_db(s, dy + (int64_t)j * seg_out, n_out, seg_out, tokens, lbj, t + j * r, nseg * r, dt_buf + j * r,
                       nseg * r, g_lb ? g_lb + (int64_t)j * seg_out * r : nullptr, st));
      }
    }
    if (g_la) RUN(wgrad(dt_buf, nseg * r, nseg * r, x, k_in, k_in, tokens, g_la, k_in, st));
  }
  if (g_dx) {
    g_dx->M = tokens; g_dx->N = k_in; g_dx->K = n_out;
    g_dx->A = dy; g_dx->lda = n_out; g_dx->a_mn_major = 0;
    g_dx->B = w; g_dx->ldb = k_in; g_dx->b_mn_major = 1;
    if (lora) pair2(*g_dx, dt_buf, nseg * r, la, k_in, nseg * r);
    RUN(tdm_gemm_bf16(g_dx, st));
  }
  return TDM_OK;
}

static TdmGemm G0() {
  TdmGemm g;
  memset(&g, 0, sizeof(g));
  g.alpha = 1.f;
  return g;
}

extern "C" int tdm_dit_block_bwd(const TdmDitShape* s, const TdmBlockWeights* w, const TdmBlockGrads* gr,
                                 const float* x, const void* ctx, const float* ctx_bias, const float* dx,
                                 float* dx_out, const void* ws_, void* scratch_, void* st) {
  RUN(check_shape(s, "tdm_dit_block_bwd"));
  TDM_REQUIRE(w && gr && x && ctx && dx && dx_out && ws_ && scratch_, "tdm_dit_block_bwd: null argument");
  TDM_REQUIRE(((uintptr_t)ws_ & 255) == 0 && ((uintptr_t)scratch_ & 255) == 0,
              "tdm_dit_block_bwd: workspace/scratch must be 256-byte aligned");
  const BlockWs a = carve_ws(s, const_cast<void*>(ws_));
  const BwdScratch c = carve_scratch(s, scratch_);
  const int B = s->B, N = s->N, D = s->D, H = s->H, d = D / H, L = s->L, Dff = s->Dff;
  const int M = B * N, BL = B * L;
  const int64_t ms = 6 * (int64_t)D;
  const bool tb = s->train_base != 0;
  const bf16* ctxb = reinterpret_cast<const bf16*>(ctx);
  cudaStream_t cs = (cudaStream_t)st;
  if (tb) TDM_CHECK_CUDA(cudaMemsetAsync(c.dmod, 0, sizeof(float) * B * 6 * D, cs));

  // ---- feed forward ----
  RUN(tdm_gate_bwd(dx, tb ? a.y3 : nullptr, a.mod + 5 * D, ms, N, c.dyA, tb ? c.dmod + 5 * D : nullptr, ms, M, D, st));
  {
    LinBwd l = {c.dyA, D, D, a.g, Dff, Dff, M, w->w_ff2, tb ? gr->w_ff2 : nullptr, tb ? gr->b_ff2 : nullptr,
                w->lora_a_ff2, w->lora_b_ff2, a.t_ff2, s->lora_r, gr->lora_a_ff2, gr->lora_b_ff2};
    TdmGemm g = G0();
    g.act = TDM_ACT_GELU_TANH_BWD; g.act_in = a.u; g.ld_act = Dff; g.out_bf16 = c.du; g.ld_out_bf16 = Dff;
    RUN(linear_bwd(s, l, c.dt, &g, st));
  }
  {
    LinBwd l = {c.du, Dff, Dff, a.h2, D, D, M, w->w_ff1, tb ? gr->w_ff1 : nullptr, tb ? gr->b_ff1 : nullptr,
                w->lora_a_ff1, w->lora_b_ff1, a.t_ff1, s->lora_r, gr->lora_a_ff1, gr->lora_b_ff1};
    TdmGemm g = G0();
    g.out_bf16 = c.dyB; g.ld_out_bf16 = D;
    RUN(linear_bwd(s, l, c.dt, &g, st));
  }
  RUN(tdm_ln_modulate_bwd(c.dyB, a.x2, a.mean2, a.rstd2, a.mod + 4 * D, ms, N, dx, c.dx2, c.dyA,
                          tb ? c.dmod + 3 * D : nullptr, tb ? c.dmod + 4 * D : nullptr, ms, M, D, st));
  // ---- cross attention ----
  {
    LinBwd l = {c.dyA, D, D, a.a2, D, D, M, w->w_o2, tb ? gr->w_o2 : nullptr, tb ? gr->b_o2 : nullptr,
                w->lora_a_o2, w->lora_b_o2, a.t_o2, s->lora_r, gr->lora_a_o2, gr->lora_b_o2};
    TdmGemm g = G0();
    g.out_bf16 = c.dyB; g.ld_out_bf16 = D;  // d a2
    RUN(linear_bwd(s, l, c.dt, &g, st));
  }
  {
    TdmAttn t;
    memset(&t, 0, sizeof(t));
    t.B = B; t.H = H; t.Sq = N; t.Sk = L; t.d = d; t.scale = 1.f / sqrtf((float)d);
    t.Q = a.q2; t.q_stride_b = (int64_t)N * D; t.q_stride_s = D; t.q_stride_h = d;
    t.K = a.kv2; t.V = a.kv2 + D;
    t.k_stride_b = t.v_stride_b = (int64_t)L * 2 * D; t.k_stride_s = t.v_stride_s = 2 * D; t.k_stride_h = t.v_stride_h = d;
    t.bias = ctx_bias;
    t.O = a.a2; t.o_stride_b = (int64_t)N * D; t.o_stride_s = D; t.o_stride_h = d;
    t.lse = a.lse2; t.dO = c.dyB; t.delta = c.delta; t.dq_accum = c.dq_accum;
    t.dQ = c.dyC; t.dq_stride_b = (int64_t)N * D; t.dq_stride_s = D; t.dq_stride_h = d;
    t.dK = c.dkv2; t.dV = c.dkv2 + D;
    t.dk_stride_b = t.dv_stride_b = (int64_t)L * 2 * D; t.dk_stride_s = t.dv_stride_s = 2 * D; t.dk_stride_h = t.dv_stride_h = d;
    RUN(tdm_attn_bwd(&t, st));
  }
  {
    const bool need = tb || (s->lora_r && w->lora_a_kv2);
    if (need) {
      TdmGemm g = G0();
      TdmGemm* gp = nullptr;
      if (tb && gr->d_ctx) {
        g.res = gr->d_ctx; g.ld_res = D; g.out_f32 = gr->d_ctx; g.ld_out_f32 = D;
        gp = &g;
      }
      RUN(stacked_linear_bwd(s, 2, c.dkv2, D, ctxb, D, BL, w->w_kv2, tb ? gr->w_kv2 : nullptr, tb ? gr->b_kv2 : nullptr,
                             w->lora_a_kv2, w->lora_b_kv2, a.t_kv2, gr->lora_a_kv2, gr->lora_b_kv2, c.dt, gp, st));
    }
  }
  {
    LinBwd l = {c.dyC, D, D, a.x1b, D, D, M, w->w_q2, tb ? gr->w_q2 : nullptr, tb ? gr->b_q2 : nullptr,
                w->lora_a_q2, w->lora_b_q2, a.t_q2, s->lora_r, gr->lora_a_q2, gr->lora_b_q2};
    TdmGemm g = G0();
    g.res = c.dx2; g.ld_res = D; g.out_f32 = c.dx2; g.ld_out_f32 = D;  // dx1 = dx2 + dq2 Wq2, in place
    RUN(linear_bwd(s, l, c.dt, &g, st));
  }
  // ---- self attention ----
  RUN(tdm_gate_bwd(c.dx2, tb ? a.y1 : nullptr, a.mod + 2 * D, ms, N, c.dyA, tb ? c.dmod + 2 * D : nullptr, ms, M, D, st));
  {
    LinBwd l = {c.dyA, D, D, a.a1, D, D, M, w->w_o, tb ? gr->w_o : nullptr, tb ? gr->b_o : nullptr,
                w->lora_a_o, w->lora_b_o, a.t_o, s->lora_r, gr->lora_a_o, gr->lora_b_o};
    TdmGemm g = G0();
    g.out_bf16 = c.dyB; g.ld_out_bf16 = D;  // d a1
    RUN(linear_bwd(s, l, c.dt, &g, st));
  }
  {
    TdmAttn t;
    memset(&t, 0, sizeof(t));
    t.B = B; t.H = H; t.Sq = N; t.Sk = N; t.d = d; t.scale = 1.f / sqrtf((float)d);
    t.Q = a.qkv; t.K = a.qkv + D; t.V = a.qkv + 2 * D;
    t.q_stride_b = t.k_stride_b = t.v_stride_b = (int64_t)N * 3 * D;
    t.q_stride_s = t.k_stride_s = t.v_stride_s = 3 * D;
    t.q_stride_h = t.k_stride_h = t.v_stride_h = d;
    t.O = a.a1; t.o_stride_b = (int64_t)N * D; t.o_stride_s = D; t.o_stride_h = d;
    t.lse = a.lse1; t.dO = c.dyB; t.delta = c.delta; t.dq_accum = c.dq_accum;
    t.dQ = c.dqkv; t.dK = c.dqkv + D; t.dV = c.dqkv + 2 * D;
    t.dq_stride_b = t.dk_stride_b = t.dv_stride_b = (int64_t)N * 3 * D;
    t.dq_stride_s = t.dk_stride_s = t.dv_stride_s = 3 * D;
    t.dq_stride_h = t.dk_stride_h = t.dv_stride_h = d;
    RUN(tdm_attn_bwd(&t, st));
  }
  {
    TdmGemm g = G0();
    g.out_bf16 = c.dyC; g.ld_out_bf16 = D;  // d h1
    RUN(stacked_linear_bwd(s, 3, c.dqkv, D, a.h1, D, M, w->w_qkv, tb ? gr->w_qkv : nullptr, tb ? gr->b_qkv : nullptr,
                           w->lora_a_qkv, w->lora_b_qkv, a.t_qkv, gr->lora_a_qkv, gr->lora_b_qkv, c.dt, &g, st));
  }
  RUN(tdm_ln_modulate_bwd(c.dyC, x, a.mean1, a.rstd1, a.mod + 1 * D, ms, N, c.dx2, dx_out, nullptr,
                          tb ? c.dmod + 0 * D : nullptr, tb ? c.dmod + 1 * D : nullptr, ms, M, D, st));
  if (tb) RUN(tdm_modulation_bwd(c.dmod, gr->table, gr->d_tmod, 1, 1, B, 6, D, st));
  return TDM_OK;
}
