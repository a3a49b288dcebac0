// extern "C" surface of libesmk.so (see include/esmk.h).  Thin forwarding layer:
// validates nothing itself, converts the opaque stream handle, never throws.
#include <atomic>

#include "common.cuh"
#include "esmk_internal.h"

struct esmk_model;
struct esmk_comm;

namespace esmk {
static std::atomic<uint64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
uint64_t launch_count() { return g_launches.load(std::memory_order_relaxed); }
}  // namespace esmk

#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define GUARD(expr)                                   \
  try {                                               \
    if (int _async = ::esmk::consume_async_error()) return _async; \
    return (expr);                                    \
  } catch (const std::exception& e) {                 \
    return ::esmk::fail("esmk", e.what());            \
  } catch (...) {                                     \
    return ::esmk::fail("esmk", "unknown exception"); \
  }

extern "C" {

const char* esmk_last_error(void) { return esmk::last_error().c_str(); }
int esmk_version(void) { return 100; }
uint64_t esmk_launch_count(void) { return esmk::launch_count(); }
int esmk_async_error(void) {
  uint32_t* w = esmk::async_error_word();
  if (w == nullptr) return 0;
  const uint32_t code = *reinterpret_cast<volatile uint32_t*>(w);
  *reinterpret_cast<volatile uint32_t*>(w) = 0;
  return (int)code;
}

int esmk_tile_capacity(int T, int B) { return (T < 0 || B < 0) ? 0 : esmk::tile_capacity(T, B); }
int esmk_batch_meta(const int32_t* cu_lens, int B, int T, int32_t* pos, int32_t* tile_info, esmk_stream_t s) {
  GUARD(esmk::batch_meta(cu_lens, B, T, pos, tile_info, ST(s)));
}
int esmk_rope_tables(void* cosb, void* sinb, int max_len, int head_dim, esmk_stream_t s) {
  GUARD(esmk::rope_tables(cosb, sinb, max_len, head_dim, ST(s)));
}
int esmk_embed(const int64_t* tokens, const void* table, void* out, int T, int D, int vocab, int zero_token,
               const uint8_t* zero_rows, esmk_stream_t s) {
  GUARD(esmk::embed(tokens, table, out, T, D, vocab, zero_token, zero_rows, ST(s)));
}
int esmk_unpad_tokens(const int64_t* tokens2d, int B, int S, int pad_token, int64_t* packed, int64_t* indices,
                      int32_t* cu_lens, int32_t* lens_scratch, int32_t* meta, esmk_stream_t s) {
  GUARD(esmk::unpad_tokens(tokens2d, B, S, pad_token, packed, indices, cu_lens, lens_scratch, meta, ST(s)));
}
int esmk_pad_rows(const void* x, int ldx, const int64_t* indices, int T, void* out, int rows, int D,
                  int32_t* inverse_scratch, esmk_stream_t s) {
  GUARD(esmk::pad_rows(x, ldx, indices, T, out, rows, D, inverse_scratch, ST(s)));
}
int esmk_add_positions(void* x, const void* table, const int32_t* pos, int T, int D, int rows, int offset, esmk_stream_t s) {
  GUARD(esmk::add_positions(x, table, pos, T, D, rows, offset, ST(s)));
}
int esmk_layernorm(const void* x, int ldx, const void* w, const void* b, void* y, int ldy, int T, int D, float eps,
                   esmk_stream_t s) {
  GUARD(esmk::layernorm(x, ldx, w, b, y, ldy, T, D, eps, ST(s)));
}
int esmk_qk_norm_rope(void* q, void* k, int ld, int T, int H, int head_dim, const void* lnq, const void* lnk,
                      const void* cosb, const void* sinb, const int32_t* pos, esmk_stream_t s) {
  GUARD(esmk::qk_norm_rope(q, k, ld, T, H, head_dim, lnq, lnk, cosb, sinb, pos, ST(s)));
}
int esmk_mean_pool(const void* x, int ldx, const int32_t* cu_lens, int B, int D, void* out, int ldo, esmk_stream_t s) {
  GUARD(esmk::mean_pool(x, ldx, cu_lens, B, D, out, ldo, ST(s)));
}
int esmk_residual_add(const void* x, const void* y, void* out, long n, float residue_scaling, esmk_stream_t s) {
  GUARD(esmk::residual_add(x, y, out, n, residue_scaling, ST(s)));
}
int esmk_softmax(const void* logits, int ld_in, void* out, int ld_out, int T, int V, int log, esmk_stream_t s) {
  GUARD(esmk::softmax(logits, ld_in, out, ld_out, T, V, log, ST(s)));
}
int esmk_gemm(const esmk_gemm_args* a, esmk_stream_t s) {
  if (a == nullptr) return esmk::fail("esmk_gemm", "null args");
  GUARD(esmk::gemm(*a, ST(s)));
}
int esmk_attn_varlen(const void* q, const void* k, const void* v, int ld, void* out, int ldo, const int32_t* cu_lens,
                     const int32_t* tile_info, int B, int T, int H, int head_dim, int max_len, int impl,
                     esmk_stream_t s) {
  GUARD(esmk::attn_varlen(q, k, v, ld, out, ldo, cu_lens, tile_info, B, T, H, head_dim, max_len, impl, ST(s)));
}
int esmk_attn_pool(const void* q, int ldq, const void* k, const void* v, int ld, void* out, const int32_t* cu_lens,
                   int B, int C, int H, int head_dim, esmk_stream_t s) {
  GUARD(esmk::attn_pool(q, ldq, k, v, ld, out, cu_lens, B, C, H, head_dim, ST(s)));
}
int esmk_quantize(const void* W, int N, int K, int bits, void* data, float* scale, esmk_stream_t s) {
  GUARD(esmk::quantize(W, N, K, bits, data, scale, ST(s)));
}
int esmk_dequantize(const void* data, const float* scale, int N, int K, int bits, void* W, esmk_stream_t s) {
  GUARD(esmk::dequantize(data, scale, N, K, bits, W, ST(s)));
}
int esmk_comm_unique_id(void* id128) { GUARD(esmk::comm_unique_id(id128)); }
int esmk_comm_create(esmk_comm_t** out, int world_size, int rank, const void* id128) {
  GUARD(esmk::comm_create(out, world_size, rank, id128));
}
void esmk_comm_destroy(esmk_comm_t* comm) { esmk::comm_destroy(comm); }
int esmk_allgather_logits(esmk_comm_t* comm, const void* local, int t_max, int V, const int64_t* perm, int T,
                          void* gathered, void* out, esmk_stream_t s) {
  GUARD(esmk::allgather_logits(comm, local, t_max, V, perm, T, gathered, out, ST(s)));
}
int esmk_comm_enable_peer(esmk_comm_t* comm, size_t buffer_bytes) { GUARD(esmk::comm_enable_peer(comm, buffer_bytes)); }
int esmk_comm_disable_peer(esmk_comm_t* comm) { GUARD(esmk::comm_disable_peer(comm)); }
int esmk_peer_allgather_logits(esmk_comm_t* comm, const void* local, int rows, int V, const int32_t* dest_rows, int T,
                               void* out, esmk_stream_t s) {
  GUARD(esmk::peer_allgather_logits(comm, local, rows, V, dest_rows, T, out, ST(s)));
}
void esmk_profile_enable(int on) { esmk::profile_enable(on); }
int esmk_profile_read(float* ms, int* launches, int n_categories) {
  if (ms == nullptr || launches == nullptr) return esmk::fail("esmk_profile_read", "null argument");
  GUARD(esmk::profile_read(ms, launches, n_categories));
}
int esmk_model_create(const esmk_config* cfg, const esmk_weights* w, esmk_model_t** out) {
  GUARD(esmk::model_create(cfg, w, out));
}
void esmk_model_destroy(esmk_model_t* m) { delete m; }
size_t esmk_workspace_bytes(const esmk_model_t* m, int T, int B, int max_len) {
  if (m == nullptr || T < 1 || B < 1 || max_len < 1) return 0;
  return esmk::workspace_bytes(m, T, B, max_len);
}
int esmk_forward(esmk_model_t* m, const int64_t* tokens, const int32_t* cu_lens, int T, int B, int max_len,
                 const uint8_t* zero_rows, void* workspace, size_t workspace_bytes, int output_kind, void* out,
                 void* const* layer_taps, esmk_stream_t s) {
  GUARD(esmk::forward(m, tokens, cu_lens, T, B, max_len, zero_rows, workspace, workspace_bytes, output_kind, out,
                      layer_taps, ST(s)));
}
int esmk_lm_head(esmk_model_t* m, const void* x, int T, void* workspace, size_t workspace_bytes, int output_kind,
                 void* out, esmk_stream_t s) {
  GUARD(esmk::lm_head(m, x, T, workspace, workspace_bytes, output_kind, out, ST(s)));
}

}  // extern "C"
