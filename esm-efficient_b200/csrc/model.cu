// Whole-model driver: the packed forward pass as a fixed sequence of esmk kernels
// on one stream (no host synchronisation, CUDA-graph capturable).  Replaces the
// Python layer loop of esme/esm.py:229-252 and the head of esme/head.py:25-27.
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "esmk_internal.h"

struct esmk_model {
  esmk_config cfg;
  esmk_weights w;
  std::vector<esmk_layer_weights> layers;
  // Quantised weights only: a side stream + events that expand weight n+2 into the second scratch buffer while the
  // GEMM of weight n runs (the expansion kernel needs no shared memory and co-resides with the persistent GEMM
  // CTAs), created lazily on the device of the first forward.  One forward in flight per quantised model handle.
  cudaStream_t aux = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_ready[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
  int aux_device = -1;
  ~esmk_model() {
    if (aux_device < 0) return;
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(aux_device);
    for (cudaEvent_t e : {ev_fork, ev_ready[0], ev_ready[1], ev_free[0], ev_free[1]})
      if (e) cudaEventDestroy(e);
    if (aux) cudaStreamDestroy(aux);
    cudaSetDevice(cur);
  }
};

namespace esmk {

namespace {

// ---------------------------------------------------------------------------
// Optional per-kernel-family device timing (bench.py's roofline block): CUDA event
// pairs recorded on the launching stream around each launch of esmk_forward.
// ---------------------------------------------------------------------------
struct Profiler {
  bool enabled = false;
  std::vector<cudaEvent_t> pool;
  std::vector<std::pair<int, int>> spans;   // (category, index of start event)
  size_t used = 0;
  cudaEvent_t get() {
    if (used == pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      pool.push_back(e);
    }
    return pool[used++];
  }
};
Profiler g_prof;

struct Span {
  int cat;
  cudaStream_t st;
  bool on;
  Span(int c, cudaStream_t s) : cat(c), st(s), on(g_prof.enabled) {
    if (on) {
      g_prof.spans.emplace_back(cat, (int)g_prof.used);
      cudaEventRecord(g_prof.get(), st);
    }
  }
  ~Span() {
    if (on) cudaEventRecord(g_prof.get(), st);
  }
};
#define PROF(cat, expr)       \
  do {                        \
    Span _span(cat, st);      \
    ESMK_TRY(expr);           \
  } while (0)

inline size_t align_up(size_t x, size_t a = 1024) { return (x + a - 1) / a * a; }

struct Workspace {
  uint8_t* base;
  size_t off = 0;
  explicit Workspace(void* p) : base(static_cast<uint8_t*>(p)) {}
  template <typename T>
  T* take(size_t count) {
    T* p = reinterpret_cast<T*>(base + off);
    off += align_up(count * sizeof(T));
    return p;
  }
};

struct Buffers {
  __nv_bfloat16 *x, *h, *qkv, *a, *u, *cosb, *sinb, *wq, *wq2, *qkvp, *ap;
  int32_t *pos, *tile_info;
  size_t bytes;
};

// elements of the largest quantised weight of the model (0 = nothing is quantised)
size_t quant_scratch_elems(const esmk_model* m) {
  const esmk_config& c = m->cfg;
  const size_t D = c.embed_dim, F = c.ffn_dim, F1 = c.family == 1 ? 2 * F : F;
  size_t n = 0;
  for (const esmk_layer_weights& l : m->layers) {
    if (l.q_wqkv.data) n = std::max(n, 3 * D * D);
    if (l.q_wo.data) n = std::max(n, D * D);
    if (l.q_w1.data) n = std::max(n, F1 * D);
    if (l.q_w2.data) n = std::max(n, D * F);
  }
  return n;
}

Buffers carve(const esmk_config& c, void* ws, int T, int B, int max_len, size_t wq_elems) {
  Workspace w(ws);
  Buffers b;
  const size_t D = c.embed_dim, F = c.ffn_dim, hd = c.embed_dim / c.attention_heads;
  b.x = w.take<__nv_bfloat16>((size_t)T * D);
  b.h = w.take<__nv_bfloat16>((size_t)T * D);
  b.qkv = w.take<__nv_bfloat16>((size_t)T * 3 * D);
  b.a = w.take<__nv_bfloat16>((size_t)T * D);
  b.u = w.take<__nv_bfloat16>((size_t)T * F);
  b.cosb = w.take<__nv_bfloat16>((size_t)max_len * hd);
  b.sinb = w.take<__nv_bfloat16>((size_t)max_len * hd);
  b.pos = w.take<int32_t>((size_t)T);
  b.tile_info = w.take<int32_t>((size_t)4 * tile_capacity(T, B));
  b.wq = wq_elems ? w.take<__nv_bfloat16>(wq_elems) : nullptr;     // two scratch weights: one being read by a GEMM,
  b.wq2 = wq_elems ? w.take<__nv_bfloat16>(wq_elems) : nullptr;    // one being expanded for the GEMM after next
  // other small head dims (ESM2-35M: 24): heads zero-padded to the next width the tcgen05 attention kernel is
  // instantiated for (32 or 64 columns, see pad_heads); hd 16 / 32 / 64 / 128 run it on the compact layout
  const bool pad = hd < 64 && hd != 16 && hd != 32;
  const int hp = hd < 32 ? 32 : 64;
  b.qkvp = pad ? w.take<__nv_bfloat16>((size_t)T * 3 * c.attention_heads * hp) : nullptr;
  b.ap = pad ? w.take<__nv_bfloat16>((size_t)T * c.attention_heads * hp) : nullptr;
  b.bytes = w.off;
  return b;
}

int linear(const void* A, int lda, const void* W, const void* bias, void* C, int ldc, int M, int N, int K, int epi,
           cudaStream_t st, const void* R = nullptr, int ldr = 0, float scale = 1.f) {
  esmk_gemm_args g{};
  g.A = A; g.lda = lda; g.W = W; g.bias = bias; g.C = C; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K; g.epilogue = epi; g.R = R; g.ldr = ldr; g.residue_scaling = scale;
  return gemm(g, st);
}

// ---------------------------------------------------------------------------
// Quantised weights -> bf16 scratch, two GEMMs ahead of their consumer on a side stream.
// ---------------------------------------------------------------------------
struct QItem { const esmk_qweight* q; int N, K; };

struct QPipe {
  esmk_model* m;
  std::vector<QItem> items;      // the quantised weights of one forward, in the order their GEMMs run
  __nv_bfloat16* scratch[2];
  cudaStream_t main;
  bool overlap;
  size_t next = 0;               // next item to hand to a GEMM

  int expand(size_t n, cudaStream_t st) {
    const QItem& it = items[n];
    return dequantize(it.q->data, it.q->scale, it.N, it.K, it.q->bits, scratch[n & 1], st);
  }
  int start() {
    if (items.empty()) return 0;
    if (!overlap) return 0;
    ESMK_CUDA(cudaEventRecord(m->ev_fork, main));                 // the side stream starts behind everything queued so far
    ESMK_CUDA(cudaStreamWaitEvent(m->aux, m->ev_fork, 0));
    for (size_t n = 0; n < 2 && n < items.size(); ++n) {
      ESMK_TRY(expand(n, m->aux));
      ESMK_CUDA(cudaEventRecord(m->ev_ready[n & 1], m->aux));
    }
    return 0;
  }
  // the bf16 weight the next GEMM should read (call in GEMM order; `w` = the caller's bf16 weight when not quantised)
  int weight(const void* w, const esmk_qweight& q, const void** out) {
    if (q.data == nullptr) { *out = w; return 0; }
    const size_t n = next++;
    *out = scratch[n & 1];
    if (!overlap) {
      Span span(ESMK_PROF_DEQUANT, main);
      return expand(n, main);
    }
    ESMK_CUDA(cudaStreamWaitEvent(main, m->ev_ready[n & 1], 0));
    return 0;
  }
  // after the GEMM that read item n has been enqueued: its scratch may be refilled with item n + 2
  int consumed() {
    if (!overlap || next == 0) return 0;
    const size_t n = next - 1;
    if (n + 2 >= items.size()) return 0;
    ESMK_CUDA(cudaEventRecord(m->ev_free[n & 1], main));
    ESMK_CUDA(cudaStreamWaitEvent(m->aux, m->ev_free[n & 1], 0));
    ESMK_TRY(expand(n + 2, m->aux));
    ESMK_CUDA(cudaEventRecord(m->ev_ready[n & 1], m->aux));
    return 0;
  }
};

int ensure_aux(esmk_model* m) {
  const int dev = current_device();
  if (m->aux_device == dev) return 0;
  ESMK_REQUIRE(m->aux_device < 0, "a quantised model handle is bound to the device of its first forward");
  ESMK_CUDA(cudaStreamCreateWithFlags(&m->aux, cudaStreamNonBlocking));
  for (cudaEvent_t* e : {&m->ev_fork, &m->ev_ready[0], &m->ev_ready[1], &m->ev_free[0], &m->ev_free[1]})
    ESMK_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  m->aux_device = dev;
  return 0;
}

int head_and_output(const esmk_model* m, const __nv_bfloat16* z, int T, __nv_bfloat16* t0, __nv_bfloat16* t1,
                    int kind, void* out, cudaStream_t st) {
  const esmk_config& c = m->cfg;
  const int D = c.embed_dim, V = c.vocab;
  // esme/head.py:25-27: final(LN(gelu(dense(x))))
  ESMK_TRY(linear(z, D, m->w.head_dense_w, m->w.head_dense_b, t0, D, T, D, D, ESMK_EPI_BIAS_GELU, st));
  ESMK_TRY(layernorm(t0, D, m->w.head_norm_w, m->w.head_norm_b, t1, D, T, D, 1e-5f, st));
  ESMK_TRY(linear(t1, D, m->w.head_final_w, m->w.head_final_b, out, V, T, V, D, ESMK_EPI_BIAS, st));
  if (kind == ESMK_OUT_LOG_PROB) ESMK_TRY(softmax(out, V, out, V, T, V, 1, st));
  if (kind == ESMK_OUT_PROB) ESMK_TRY(softmax(out, V, out, V, T, V, 0, st));
  return 0;
}

}  // namespace

void profile_enable(int on) {
  g_prof.enabled = on != 0;
  g_prof.spans.clear();
  g_prof.used = 0;
}

// Sums the elapsed time of every recorded span per category (ms) and the launch count; the caller
// must have synchronised the stream.  Resets the recording.
int profile_read(float* ms, int* launches, int n_categories) {
  for (int i = 0; i < n_categories; ++i) { ms[i] = 0.f; launches[i] = 0; }
  for (auto& sp : g_prof.spans) {
    float t = 0.f;
    ESMK_CUDA(cudaEventElapsedTime(&t, g_prof.pool[sp.second], g_prof.pool[sp.second + 1]));
    if (sp.first >= 0 && sp.first < n_categories) { ms[sp.first] += t; launches[sp.first] += 1; }
  }
  g_prof.spans.clear();
  g_prof.used = 0;
  return 0;
}

int model_create(const esmk_config* cfg, const esmk_weights* w, esmk_model** out) {
  ESMK_REQUIRE(cfg && w && out, "null argument");
  ESMK_REQUIRE(cfg->family == 0 || cfg->family == 1, "family must be 0 (ESM2) or 1 (ESMC)");
  ESMK_REQUIRE(cfg->num_layers >= 1 && cfg->embed_dim >= 8 && cfg->attention_heads >= 1, "bad model dims");
  ESMK_REQUIRE(cfg->embed_dim % cfg->attention_heads == 0, "embed_dim must be divisible by attention_heads");
  ESMK_REQUIRE(cfg->embed_dim % 8 == 0 && cfg->ffn_dim % 8 == 0, "embed_dim / ffn_dim must be multiples of 8");
  const int hd = cfg->embed_dim / cfg->attention_heads;
  ESMK_REQUIRE(hd % 8 == 0 && hd >= 8 && hd <= 128, "head_dim must be a multiple of 8, at most 128");
  if (cfg->family == 1) ESMK_REQUIRE(hd == 16 || hd == 32 || hd == 64 || hd == 128, "ESMC head_dim must be 16, 32, 64 or 128");
  ESMK_REQUIRE(cfg->vocab >= 1 && cfg->vocab <= 128 && cfg->embed_rows >= cfg->vocab - 0, "bad vocab");
  ESMK_REQUIRE(cfg->residue_scaling > 0.f, "residue_scaling must be positive");
  ESMK_REQUIRE((w->pos_embed != nullptr) == (cfg->pos_rows > 0), "pos_embed and pos_rows must be given together");
  ESMK_REQUIRE(cfg->family == 0 || (!cfg->no_rotary && w->pos_embed == nullptr), "learned positions are an ESM-1b/1v (family 0) feature");
  if (cfg->family == 1) ESMK_REQUIRE(cfg->ffn_dim % 32 == 0, "ESMC ffn_dim must be a multiple of 32");
  ESMK_REQUIRE(w->embed && w->layers && w->final_norm_w && w->head_dense_w && w->head_dense_b && w->head_norm_w &&
                   w->head_norm_b && w->head_final_w && w->head_final_b,
               "missing model-level weight");
  for (int i = 0; i < cfg->num_layers; ++i) {
    const esmk_layer_weights& l = w->layers[i];
    ESMK_REQUIRE(l.attn_norm_w && l.attn_norm_b && l.ffn_norm_w && l.ffn_norm_b, "missing layer weight");
    ESMK_REQUIRE((l.wqkv || l.q_wqkv.data) && (l.wo || l.q_wo.data) && (l.w1 || l.q_w1.data) && (l.w2 || l.q_w2.data),
                 "missing layer weight (neither bf16 nor quantised)");
    for (const esmk_qweight* q : {&l.q_wqkv, &l.q_wo, &l.q_w1, &l.q_w2})
      if (q->data) ESMK_REQUIRE(q->scale && (q->bits == 4 || q->bits == 8), "bad quantised weight descriptor");
    if (l.q_wqkv.data || l.q_wo.data || l.q_w2.data || l.q_w1.data)
      ESMK_REQUIRE(cfg->embed_dim % 64 == 0 && cfg->ffn_dim % 64 == 0, "quantised weights need dims that are multiples of 64");
    if (cfg->family == 0) ESMK_REQUIRE(l.bqkv && l.bo && l.b1 && l.b2, "ESM2 layers need biases");
    if (cfg->family == 1) ESMK_REQUIRE(l.qln_w && l.kln_w, "ESMC layers need QK-LayerNorm weights");
  }
  esmk_model* m = new esmk_model();
  m->cfg = *cfg;
  m->w = *w;
  m->layers.assign(w->layers, w->layers + cfg->num_layers);
  m->w.layers = m->layers.data();
  *out = m;
  return 0;
}

size_t workspace_bytes(const esmk_model* m, int T, int B, int max_len) {
  return carve(m->cfg, nullptr, T, B, max_len, quant_scratch_elems(m)).bytes + 1024;
}

int forward(esmk_model* m, const int64_t* tokens, const int32_t* cu_lens, int T, int B, int max_len,
            const uint8_t* zero_rows, void* workspace, size_t workspace_bytes_, int kind, void* out,
            void* const* layer_taps, cudaStream_t st) {
  ESMK_REQUIRE(m && tokens && cu_lens && workspace && out, "null argument");
  ESMK_REQUIRE(T >= 1 && B >= 1 && max_len >= 1, "empty batch");
  ESMK_REQUIRE(kind >= ESMK_OUT_LOGITS && kind <= ESMK_OUT_REPRESENTATION, "bad output kind");
  const esmk_config& c = m->cfg;
  void* ws = reinterpret_cast<void*>(align_up(reinterpret_cast<uintptr_t>(workspace)));
  Buffers b = carve(c, ws, T, B, max_len, quant_scratch_elems(m));
  ESMK_REQUIRE(b.bytes + (static_cast<uint8_t*>(ws) - static_cast<uint8_t*>(workspace)) <= workspace_bytes_,
               "workspace too small (see esmk_workspace_bytes)");
  const int D = c.embed_dim, H = c.attention_heads, hd = D / H, F = c.ffn_dim;
  const float s = c.residue_scaling;
  const bool fused_rope = (c.family == 0) && !c.no_rotary && (hd == 16 || hd == 32 || hd == 64) && ((2 * D) % 64 == 0);

  QPipe qp{m, {}, {b.wq, b.wq2}, st, false};
  {
    const int F1 = c.family == 0 ? F : 2 * F;
    for (const esmk_layer_weights& l : m->layers) {
      if (l.q_wqkv.data) qp.items.push_back({&l.q_wqkv, 3 * D, D});
      if (l.q_wo.data) qp.items.push_back({&l.q_wo, D, D});
      if (l.q_w1.data) qp.items.push_back({&l.q_w1, F1, D});
      if (l.q_w2.data) qp.items.push_back({&l.q_w2, D, F});
    }
    static const bool overlap_enabled = [] { const char* e = getenv("ESMK_QUANT_OVERLAP"); return e == nullptr || e[0] != '0'; }();
    if (!qp.items.empty() && overlap_enabled) {
      ESMK_TRY(ensure_aux(m));
      qp.overlap = true;
    }
    ESMK_TRY(qp.start());
  }
  PROF(ESMK_PROF_MISC, batch_meta(cu_lens, B, T, b.pos, b.tile_info, st));
  PROF(ESMK_PROF_MISC, rope_tables(b.cosb, b.sinb, max_len, hd, st));
  // esme/esm.py:188-189: ESM2 zeroes <mask>(32) rows; ESMC (esm.py:876) does not
  PROF(ESMK_PROF_MISC, embed(tokens, m->w.embed, b.x, T, D, c.embed_rows, c.family == 0 ? 32 : -1, zero_rows, st));
  if (m->w.pos_embed != nullptr) {   // ESM-1b / ESM-1v: learned positions (+ LayerNorm before the layers for 1b)
    PROF(ESMK_PROF_MISC, add_positions(b.x, m->w.pos_embed, b.pos, T, D, c.pos_rows, 2, st));
    if (m->w.pre_norm_w != nullptr)
      PROF(ESMK_PROF_LAYERNORM, layernorm(b.x, D, m->w.pre_norm_w, m->w.pre_norm_b, b.x, D, T, D, 1e-5f, st));
  }

  for (int i = 0; i < c.num_layers; ++i) {
    const esmk_layer_weights& l = m->layers[i];
    // ---- attention block: x = x + out(attn(rope(qkv(LN(x))))) / s   (esme/attention.py:126-139, 253-254)
    PROF(ESMK_PROF_LAYERNORM, layernorm(b.x, D, l.attn_norm_w, l.attn_norm_b, b.h, D, T, D, 1e-5f, st));
    const void *wqkv, *wo, *w1, *w2;
    ESMK_TRY(qp.weight(l.wqkv, l.q_wqkv, &wqkv));
    if (fused_rope) {
      esmk_gemm_args g{};
      g.A = b.h; g.lda = D; g.W = wqkv; g.bias = l.bqkv; g.C = b.qkv; g.ldc = 3 * D;
      g.M = T; g.N = 3 * D; g.K = D; g.epilogue = ESMK_EPI_QKV_ROPE;
      g.rope_cos = b.cosb; g.rope_sin = b.sinb; g.pos = b.pos; g.head_dim = hd; g.rope_cols = 2 * D;
      PROF(ESMK_PROF_GEMM_QKV, gemm(g, st));
      if (l.q_wqkv.data) ESMK_TRY(qp.consumed());
    } else {
      PROF(ESMK_PROF_GEMM_QKV, linear(b.h, D, wqkv, l.bqkv, b.qkv, 3 * D, T, 3 * D, D, ESMK_EPI_BIAS, st));
      if (l.q_wqkv.data) ESMK_TRY(qp.consumed());
      if (!c.no_rotary || l.qln_w != nullptr)
        PROF(ESMK_PROF_ROPE, qk_norm_rope(b.qkv, b.qkv + D, 3 * D, T, H, hd, l.qln_w, l.kln_w, c.no_rotary ? nullptr : b.cosb,
                                          c.no_rotary ? nullptr : b.sinb, b.pos, st));
    }
    if (hd < 64 && hd != 16 && hd != 32) {
      const int hp = hd < 32 ? 32 : 64;
      const int Dp = H * hp;
      PROF(ESMK_PROF_ATTENTION, pad_heads(b.qkv, 3 * D, b.qkvp, T, 3, H, hd, hp, st));
      PROF(ESMK_PROF_ATTENTION, attn_varlen(b.qkvp, b.qkvp + Dp, b.qkvp + 2 * Dp, 3 * Dp, b.ap, Dp, cu_lens, b.tile_info, B,
                                            T, H, hp, max_len, 0, st, hd));
      PROF(ESMK_PROF_ATTENTION, unpad_heads(b.ap, b.a, D, T, H, hd, hp, st));
    } else {
      PROF(ESMK_PROF_ATTENTION,
           attn_varlen(b.qkv, b.qkv + D, b.qkv + 2 * D, 3 * D, b.a, D, cu_lens, b.tile_info, B, T, H, hd, max_len, 0, st));
    }
    ESMK_TRY(qp.weight(l.wo, l.q_wo, &wo));
    PROF(ESMK_PROF_GEMM_OUT, linear(b.a, D, wo, l.bo, b.x, D, T, D, D, ESMK_EPI_RESIDUAL, st, b.x, D, s));
    if (l.q_wo.data) ESMK_TRY(qp.consumed());
    // ---- FFN block: x = x + final(x) / s   (esme/attention.py:217-236, 255)
    PROF(ESMK_PROF_LAYERNORM, layernorm(b.x, D, l.ffn_norm_w, l.ffn_norm_b, b.h, D, T, D, 1e-5f, st));
    ESMK_TRY(qp.weight(l.w1, l.q_w1, &w1));
    if (c.family == 0) {
      PROF(ESMK_PROF_GEMM_FFN_UP, linear(b.h, D, w1, l.b1, b.u, F, T, F, D, ESMK_EPI_BIAS_GELU, st));
    } else {
      PROF(ESMK_PROF_GEMM_FFN_UP, linear(b.h, D, w1, nullptr, b.u, F, T, 2 * F, D, ESMK_EPI_SWIGLU, st));
    }
    if (l.q_w1.data) ESMK_TRY(qp.consumed());
    ESMK_TRY(qp.weight(l.w2, l.q_w2, &w2));
    PROF(ESMK_PROF_GEMM_FFN_DOWN, linear(b.u, F, w2, l.b2, b.x, D, T, D, F, ESMK_EPI_RESIDUAL, st, b.x, D, s));
    if (l.q_w2.data) ESMK_TRY(qp.consumed());
    if (layer_taps != nullptr && layer_taps[i] != nullptr)
      ESMK_CUDA(cudaMemcpyAsync(layer_taps[i], b.x, (size_t)T * D * 2, cudaMemcpyDeviceToDevice, st));
  }
  // esme/esm.py:252
  if (kind == ESMK_OUT_REPRESENTATION)
    return layernorm(b.x, D, m->w.final_norm_w, m->w.final_norm_b, out, D, T, D, 1e-5f, st);
  PROF(ESMK_PROF_LAYERNORM, layernorm(b.x, D, m->w.final_norm_w, m->w.final_norm_b, b.h, D, T, D, 1e-5f, st));
  Span head_span(ESMK_PROF_HEAD, st);
  return head_and_output(m, b.h, T, b.a, b.x, kind, out, st);
}

int lm_head(esmk_model* m, const void* x, int T, void* workspace, size_t workspace_bytes_, int kind, void* out,
            cudaStream_t st) {
  ESMK_REQUIRE(m && x && workspace && out, "null argument");
  ESMK_REQUIRE(kind >= ESMK_OUT_LOGITS && kind <= ESMK_OUT_PROB, "bad output kind");
  if (T == 0) return 0;
  const size_t D = m->cfg.embed_dim;
  uint8_t* ws = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(workspace)));
  const size_t one = align_up((size_t)T * D * 2);
  ESMK_REQUIRE((size_t)(ws - static_cast<uint8_t*>(workspace)) + 2 * one <= workspace_bytes_,
               "workspace too small: need 2*T*D*2 + 3072 bytes");
  return head_and_output(m, static_cast<const __nv_bfloat16*>(x), T, reinterpret_cast<__nv_bfloat16*>(ws),
                         reinterpret_cast<__nv_bfloat16*>(ws + one), kind, out, st);
}

}  // namespace esmk
