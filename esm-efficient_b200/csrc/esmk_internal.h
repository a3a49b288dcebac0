// Internal C++ interface between the kernel translation units and the C ABI (api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/esmk.h"

namespace esmk {

const std::string& last_error();
void count_launch(int n = 1);
uint64_t launch_count();

int tile_capacity(int T, int B);
int batch_meta(const int32_t* cu_lens, int B, int T, int32_t* pos, int32_t* tile_info, cudaStream_t st);
int rope_tables(void* cosb, void* sinb, int max_len, int hd, cudaStream_t st);
int embed(const int64_t* tokens, const void* table, void* out, int T, int D, int vocab, int zero_token,
          const uint8_t* zero_rows, cudaStream_t st);
int unpad_tokens(const int64_t* tokens2d, int B, int S, int pad_token, int64_t* packed, int64_t* indices,
                 int32_t* cu_lens, int32_t* lens_scratch, int32_t* meta, cudaStream_t st);
int pad_rows(const void* x, int ldx, const int64_t* indices, int T, void* out, int rows, int D, int32_t* inverse_scratch,
             cudaStream_t st);
int add_positions(void* x, const void* table, const int32_t* pos, int T, int D, int rows, int offset, cudaStream_t st);
int layernorm(const void* x, int ldx, const void* w, const void* b, void* y, int ldy, int T, int D, float eps,
              cudaStream_t st);
int qk_norm_rope(void* q, void* k, int ld, int T, int H, int hd, const void* lnq, const void* lnk, const void* cosb,
                 const void* sinb, const int32_t* pos, cudaStream_t st);
int mean_pool(const void* x, int ldx, const int32_t* cu_lens, int B, int D, void* out, int ldo, cudaStream_t st);
int residual_add(const void* x, const void* y, void* out, long n, float scale, cudaStream_t st);
int softmax(const void* x, int ldx, void* y, int ldy, int T, int V, int log_mode, cudaStream_t st);

int gemm(const esmk_gemm_args& a, cudaStream_t st);

int quantize(const void* w, int N, int K, int bits, void* data, float* scale, cudaStream_t st);
int dequantize(const void* data, const float* scale, int N, int K, int bits, void* w, cudaStream_t st);

int attn_varlen(const void* q, const void* k, const void* v, int ld, void* out, int ldo, const int32_t* cu_lens,
                const int32_t* tile_info, int B, int T, int H, int hd, int max_len, int impl, cudaStream_t st,
                int scale_hd = 0);
int pad_heads(const void* src, int ld_src, void* dst, int T, int parts, int H, int hd, int hp, cudaStream_t st);
int unpad_heads(const void* src, void* dst, int ld_dst, int T, int H, int hd, int hp, cudaStream_t st);

int attn_pool(const void* q, int ldq, const void* k, const void* v, int ld, void* out, const int32_t* cu_lens, int B,
              int C, int H, int hd, cudaStream_t st);

int comm_unique_id(void* id128);
int comm_create(esmk_comm** out, int world, int rank, const void* id128);
void comm_destroy(esmk_comm* c);
int allgather_logits(esmk_comm* c, const void* local, int t_max, int V, const int64_t* perm, int T, void* gathered,
                     void* out, cudaStream_t st);
int comm_enable_peer(esmk_comm* c, size_t buffer_bytes);
int comm_disable_peer(esmk_comm* c);
int peer_allgather_logits(esmk_comm* c, const void* local, int rows, int V, const int32_t* dest_rows, int T, void* out,
                          cudaStream_t st);

void profile_enable(int on);
int profile_read(float* ms, int* launches, int n_categories);

int model_create(const esmk_config* cfg, const esmk_weights* w, esmk_model** out);
size_t workspace_bytes(const esmk_model* m, int T, int B, int max_len);
int forward(esmk_model* m, const int64_t* tokens, const int32_t* cu_lens, int T, int B, int max_len,
            const uint8_t* zero_rows, void* workspace, size_t workspace_bytes, int kind, void* out,
            void* const* layer_taps, cudaStream_t st);
int lm_head(esmk_model* m, const void* x, int T, void* workspace, size_t workspace_bytes, int kind, void* out,
            cudaStream_t st);

}  // namespace esmk
