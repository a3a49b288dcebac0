/* esmk.h -- C ABI of libesmk.so, the sm_100a (B200) kernels behind the `esme`
 * drop-in for the ESM inference forward pass.
 *
 * Conventions
 *   - Every entry point returns 0 on success, non-zero on error; the message
 *     is available (thread-local) through esmk_last_error().  Nothing throws
 *     across the ABI.
 *   - All tensor pointers are DEVICE pointers to row-major bf16 unless stated;
 *     `ld*` are row pitches in ELEMENTS.  `stream` is a cudaStream_t.
 *   - The caller allocates every output and workspace; the library owns no
 *     tensor memory (esmk_model_t only keeps the pointers it was given plus
 *     pre-encoded TMA descriptors).  Calls are asynchronous on `stream`.
 *   - There is no CPU fallback: without a CUDA device every compute entry
 *     point fails with a CUDA error.
 *
 * Each function cites the reference (`/root/reference`, package `esme`)
 * operator it replaces.
 */
#ifndef ESMK_H
#define ESMK_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define ESMK_API __attribute__((visibility("default")))
#else
#define ESMK_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef void* esmk_stream_t; /* cudaStream_t */

/* ---- library ----------------------------------------------------------- */
ESMK_API const char* esmk_last_error(void);
ESMK_API int esmk_version(void);
/* Device-detected data errors.  Kernels cannot return a status, so invalid DATA met on the device (a token id outside
 * the embedding table -- the reference would trip a device-side assert in F.embedding, esme/esm.py:187 -- or a
 * sequence longer than the learned positional table, esme/embedding.py) sets a sticky flag that makes the NEXT entry
 * point called after the kernel ran fail with a message naming the cause (the flag is then cleared).
 * esmk_async_error() polls and clears it explicitly: 0 or a bit-or of the codes below (ESMK_ASYNC_PEER_TIMEOUT: a rank
 * never arrived in esmk_peer_allgather_logits). */
enum esmk_async_code { ESMK_ASYNC_BAD_TOKEN = 1, ESMK_ASYNC_BAD_POSITION = 2, ESMK_ASYNC_PEER_TIMEOUT = 4 };
ESMK_API int esmk_async_error(void);
/* number of kernels this library has launched in the calling process (for bench `gpu_launches`) */
ESMK_API uint64_t esmk_launch_count(void);

/* ---- batch metadata ---------------------------------------------------- */
/* Per-token position inside its own sequence and the attention work list, computed once per batch.
 * Replaces esme/rotary.py:5-14 `culen_indices` (which the reference recomputes, with host syncs,
 * twice per layer).
 *   cu_lens    int32[B+1]                            (device)
 *   pos        int32[T]   out (may be NULL): t - cu_lens[seq(t)]
 *   tile_info  int32[4 * esmk_tile_capacity(T, B)]  out (may be NULL), 16-byte aligned: one
 *              {sequence start row, sequence length, first query row of the tile, sequence id} record
 *              per 128-query tile, longest sequences first; unused records have length 0. */
ESMK_API int esmk_tile_capacity(int T, int B);
ESMK_API int esmk_batch_meta(const int32_t* cu_lens, int B, int T, int32_t* pos, int32_t* tile_info, esmk_stream_t stream);

/* cos/sin tables, esme/rotary.py:116-149: fp32 angle = p * 10000^(-2i/hd), table row = [f, f],
 * cast to bf16.  cos, sin: bf16 [max_len, hd]. */
ESMK_API int esmk_rope_tables(void* cos, void* sin, int max_len, int head_dim, esmk_stream_t stream);

/* ---- row-wise operators ------------------------------------------------ */
/* esme/esm.py:176-199 (ESM2.embedding) / esme/esm.py:876 (ESMC): out[t] = table[tokens[t]];
 * rows whose token == zero_token (ESM2: <mask>=32, ESMC: -1 = none) or whose
 * zero_rows[t] != 0 (optional uint8[T], e.g. pad rows of the 2-D entry) are zero. */
ESMK_API int esmk_embed(const int64_t* tokens, const void* table, void* out, int T, int D, int vocab, int zero_token,
               const uint8_t* zero_rows, esmk_stream_t stream);

/* Padded token grid <-> packed batch: the jobs of flash_attn.bert_padding.unpad_input / pad_input as the reference
 * calls them for 2-D token input (esme/esm.py:235-239, 254-261).
 *   esmk_unpad_tokens: tokens2d int64[B,S]; every token != pad_token is kept, in row-major order:
 *       packed  int64[B*S] (first T valid)       indices int64[B*S] (flat grid index b*S + s of each packed token)
 *       cu_lens int32[B+1]                        lens_scratch int32[B]
 *       meta    int32[2] = {T, max_len} (device; the caller reads it back once -- the only synchronisation of the entry)
 *   esmk_pad_rows: out[rows, D] (dense) = 0 except out[indices[t]] = x[t] for t < T; inverse_scratch int32[rows]. */
ESMK_API int esmk_unpad_tokens(const int64_t* tokens2d, int B, int S, int pad_token, int64_t* packed, int64_t* indices,
                               int32_t* cu_lens, int32_t* lens_scratch, int32_t* meta, esmk_stream_t stream);
ESMK_API int esmk_pad_rows(const void* x, int ldx, const int64_t* indices, int T, void* out, int rows, int D,
                           int32_t* inverse_scratch, esmk_stream_t stream);

/* ESM-1b / ESM-1v learned positional embedding, esme/esm.py:634-646 + esme/embedding.py:36-92, in place:
 * x[t] = bf(x[t] + table[pos[t] + offset]); pos from esmk_batch_meta, offset = padding_idx + 1 = 2, table bf16 [rows, D]. */
ESMK_API int esmk_add_positions(void* x, const void* table, const int32_t* pos, int T, int D, int rows, int offset,
                                esmk_stream_t stream);

/* torch.nn.LayerNorm over the last dim (eps, biased variance, fp32 statistics, bf16 in/out):
 * esme/attention.py:92, 222|230; esme/esm.py:252; esme/head.py:26.  bias may be NULL. */
ESMK_API int esmk_layernorm(const void* x, int ldx, const void* weight, const void* bias, void* y, int ldy, int T, int D,
                   float eps, esmk_stream_t stream);

/* esme/rotary.py:22-43 `apply_rotary` on packed [T,H,hd] q and k (in place allowed):
 * out = bf(bf(x*cos[pos]) + bf(rotate_half(x)*sin[pos])).  If ln_q_weight/ln_k_weight are
 * non-NULL the ESMC QK-LayerNorm over the full D = H*hd (weight only, eps 1e-5;
 * esme/attention.py:104-105) is applied first.  cos/sin may be NULL to skip the rotation. */
ESMK_API int esmk_qk_norm_rope(void* q, void* k, int ld, int T, int H, int head_dim, const void* ln_q_weight,
                      const void* ln_k_weight, const void* cos, const void* sin, const int32_t* pos,
                      esmk_stream_t stream);

/* esme/pooling.py:44-69 `partition_mean_pool`: out[s] = mean of the packed rows of sequence s
 * (all tokens incl. <cls>/<eos>, as the reference), fp32 accumulation, bf16 out [B, D]. */
ESMK_API int esmk_mean_pool(const void* x, int ldx, const int32_t* cu_lens, int B, int D, void* out, int ldo,
                            esmk_stream_t stream);

/* esme/attention.py:253-255 as stand-alone elementwise ops: out = bf(x + bf(y / residue_scaling)) over n contiguous
 * bf16 elements (n % 8 == 0; out may alias x).  The model path fuses this into the GEMM epilogue; this entry serves
 * the LoRA path, where adapters are added between the projection and the residual. */
ESMK_API int esmk_residual_add(const void* x, const void* y, void* out, long n, float residue_scaling, esmk_stream_t stream);

/* esme/esm.py:297,317: (log_)softmax over the last dim of bf16 logits [T,V], bf16 out. */
ESMK_API int esmk_softmax(const void* logits, int ld_in, void* out, int ld_out, int T, int V, int log, esmk_stream_t stream);

/* ---- GEMM family (tcgen05 + TMA) ---------------------------------------- */
enum esmk_epilogue {
  ESMK_EPI_BIAS = 0,      /* C = bf(A W^T + b)                       nn.Linear                         */
  ESMK_EPI_BIAS_GELU = 1, /* C = bf(gelu_erf(bf(A W^T + b)))          attention.py:231-233, head.py:26   */
  ESMK_EPI_RESIDUAL = 2,  /* C = bf(R + bf(bf(A W^T + b) / s))        attention.py:139,253-255           */
  ESMK_EPI_QKV_ROPE = 3,  /* C = [rope(q) | rope(k) | v], W = [Wq;Wk;Wv]  attention.py:102 + rotary.py:43 */
  ESMK_EPI_SWIGLU = 4     /* C = bf(bf(silu(bf(a))) * bf(f)), W rows interleaved in blocks of 32
                             (32 `activation` rows, then the 32 matching `fc` rows)  attention.py:281   */
};

typedef struct {
  const void* A;  /* [M,K] activations */
  int lda;
  const void* W;  /* [N,K] weight, nn.Linear layout (row pitch K) */
  const void* bias; /* [N] or NULL */
  void* C;        /* [M,N] (SWIGLU: [M,N/2]) */
  int ldc;
  int M, N, K;
  int epilogue;   /* enum esmk_epilogue */
  /* RESIDUAL */
  const void* R;  /* [M,N] residual (may alias C) */
  int ldr;
  float residue_scaling; /* s (divides, as the reference does) */
  /* QKV_ROPE */
  const void* rope_cos; /* bf16 [max_len, hd] */
  const void* rope_sin;
  const int32_t* pos;   /* int32 [M] */
  int head_dim;         /* 16, 32 or 64 (other head dims: ESMK_EPI_BIAS + esmk_qk_norm_rope) */
  int rope_cols;        /* columns [0, rope_cols) are rotated (= 2*D) */
} esmk_gemm_args;

ESMK_API int esmk_gemm(const esmk_gemm_args* args, esmk_stream_t stream);

/* ---- variable-length multi-head attention -------------------------------- */
/* Replaces flash_attn_varlen_func as called at esme/attention.py:115-123:
 * per sequence s and head h, O = softmax(Q K^T * hd^-0.5) V, non-causal, no dropout,
 * fp32 scores / accumulation, P rounded to bf16 before P V.
 *   q,k,v : bf16, token t / head h at  ptr + t*ld + h*hd   (e.g. three column
 *           blocks of one [T,3D] QKV GEMM output, ld = 3D)
 *   out   : bf16 [T, H*hd], pitch ldo
 *   tile_info from esmk_batch_meta (required for the tcgen05 kernel).
 * head_dim 16, 32, 64 and 128 run the tcgen05/TMA kernel (the whole-model entry runs head_dim 24 through it with
 * zero-padded heads); other head dims (<= 128, multiple of 8) run a CUDA-core kernel here.
 * impl: 0 = auto, 1 = force the CUDA-core kernel. */
ESMK_API int esmk_attn_varlen(const void* q, const void* k, const void* v, int ld, void* out, int ldo,
                     const int32_t* cu_lens, const int32_t* tile_info, int B, int T, int H, int head_dim,
                     int max_len, int impl, esmk_stream_t stream);

/* Attention pooling, esme/pooling.py:72-136 (`AttentionPool.forward`, which calls flash_attn_varlen_func with one
 * query per (class token, sequence), max_seqlen_q = 1): out[s, c, :] = softmax(q_c K_s^T * hd^-0.5) V_s per head.
 *   q   : bf16 [C, H*hd] class tokens, pitch ldq          k, v : bf16 [T, H*hd] (k already projected), pitch ld
 *   out : bf16 [B, C, H*hd] dense.  Same arithmetic as esmk_attn_varlen (fp32 softmax, P rounded to bf16). */
ESMK_API int esmk_attn_pool(const void* q, int ldq, const void* k, const void* v, int ld, void* out,
                            const int32_t* cu_lens, int B, int C, int H, int head_dim, esmk_stream_t stream);

/* ---- weight-only quantised storage --------------------------------------- */
/* Replaces the bitsandbytes modules the reference's quantised loaders install for q, k, v, out and the two
 * FFN linears (esme/esm.py:414-472 `_load_linear4bit` / `_load_linear8bit` / `_load_quantize`, ESMC :916-946).
 * bitsandbytes is not part of /root/reference; the formats are this library's (parity unpinned):
 *   bits = 4: `data` uint8[N*K/2], two 4-bit codes per byte, EVEN element in the high nibble; code = sign (8) |
 *             index into {0, 1/192, 2/3, 1, 1/3, 1/2, 1/6, 1/4} (bitsandbytes' fp4 codebook); `scale` fp32[N*K/64]
 *             = absmax of each block of 64 consecutive weights (K % 64 == 0; no double quantisation)
 *   bits = 8: `data` int8[N,K]; `scale` fp32[N] = row absmax / 127
 * W is the bf16 [N,K] nn.Linear weight.  As in bitsandbytes' batched path the GEMM itself runs in bf16 on the
 * dequantised weight. */
ESMK_API int esmk_quantize(const void* W, int N, int K, int bits, void* data, float* scale, esmk_stream_t stream);
ESMK_API int esmk_dequantize(const void* data, const float* scale, int N, int K, int bits, void* W, esmk_stream_t stream);

typedef struct {
  const void* data;   /* NULL = this weight is not quantised */
  const float* scale;
  int bits;           /* 4 or 8 */
} esmk_qweight;

/* ---- whole model --------------------------------------------------------- */
typedef struct {
  int family;          /* 0 = ESM2 (bias, GELU FFN, mask-row zeroing), 1 = ESMC (QK-LN, SwiGLU, residue scaling) */
  int num_layers, embed_dim, attention_heads, ffn_dim, vocab, embed_rows;
  float residue_scaling;
  int no_rotary;       /* 1 = no rotary embedding (ESM-1b / ESM-1v, esme/esm.py:628-629: rotary_embedding=False) */
  int pos_rows;        /* rows of esmk_weights.pos_embed (0 = none) */
} esmk_config;

typedef struct {
  const void *attn_norm_w, *attn_norm_b;
  const void *wqkv, *bqkv;       /* [3D,D], [3D] or NULL */
  const void *qln_w, *kln_w;     /* ESMC only, else NULL */
  const void *wo, *bo;           /* [D,D] */
  const void *ffn_norm_w, *ffn_norm_b;
  const void *w1, *b1;           /* ESM2 [F,D]; ESMC interleaved [2F,D] */
  const void *w2, *b2;           /* [D,F] */
  /* optional quantised storage of wqkv / wo / w1 / w2 (same shapes and row order as the bf16 weight it
   * replaces, whose pointer may then be NULL): esmk_forward expands it into workspace scratch right before
   * the GEMM that consumes it */
  esmk_qweight q_wqkv, q_wo, q_w1, q_w2;
} esmk_layer_weights;

typedef struct {
  const void* embed;                      /* [embed_rows, D] */
  const esmk_layer_weights* layers;       /* host array [num_layers] */
  const void *final_norm_w, *final_norm_b;
  const void *head_dense_w, *head_dense_b;
  const void *head_norm_w, *head_norm_b;
  const void *head_final_w, *head_final_b; /* [V,D], [V] */
  /* ESM-1b / ESM-1v (esme/esm.py:618-735): learned positional table [pos_rows, D] added to the token embedding
   * at row pos + 2; ESM-1b also applies emb_layer_norm_before.  NULL for ESM2 / ESMC. */
  const void* pos_embed;
  const void *pre_norm_w, *pre_norm_b;
} esmk_weights;

typedef struct esmk_model esmk_model_t;

ESMK_API int esmk_model_create(const esmk_config* cfg, const esmk_weights* w, esmk_model_t** out);
ESMK_API void esmk_model_destroy(esmk_model_t* m);
/* bytes of scratch esmk_forward needs for a T-token, B-sequence batch with the given max_len */
ESMK_API size_t esmk_workspace_bytes(const esmk_model_t* m, int T, int B, int max_len);

enum esmk_output {
  ESMK_OUT_LOGITS = 0,        /* [T,V]  ESM2.forward                esme/esm.py:268-282 */
  ESMK_OUT_LOG_PROB = 1,      /* [T,V]  predict_log_prob            esme/esm.py:284-298 */
  ESMK_OUT_PROB = 2,          /* [T,V]  predict_prob                esme/esm.py:300-317 */
  ESMK_OUT_REPRESENTATION = 3 /* [T,D]  forward_representation      esme/esm.py:201-266 */
};

/* The packed forward: embedding -> layers -> final LN -> (LM head -> (log_)softmax).
 * Replaces the loop at esme/esm.py:229-252 (+ head.py:25-27).
 *   tokens int64[T], cu_lens int32[B+1] (device), zero_rows optional uint8[T]
 *   out: bf16 [T,V] or [T,D] (dense, pitch V or D)
 *   layer_taps: optional host array of num_layers device pointers (or NULL entries);
 *               layer i's output x [T,D] is copied there (forward_representation(layers=[...])). */
ESMK_API int esmk_forward(esmk_model_t* m, const int64_t* tokens, const int32_t* cu_lens, int T, int B, int max_len,
                 const uint8_t* zero_rows, void* workspace, size_t workspace_bytes, int output_kind, void* out,
                 void* const* layer_taps, esmk_stream_t stream);

/* LM head alone on arbitrary rows (the padded entry runs it on pad rows too, esme/esm.py:281):
 * x [T,D] -> out [T,V]; workspace >= 2*T*D*2 bytes. */
ESMK_API int esmk_lm_head(esmk_model_t* m, const void* x, int T, void* workspace, size_t workspace_bytes, int output_kind,
                 void* out, esmk_stream_t stream);

/* ---- multi-GPU: the one collective of the path ------------------------------- */
/* Sequences are independent, so a packed batch is split by whole sequences over the ranks (host-side partition,
 * esme/parallel.py), every rank runs esmk_forward on its share with replicated weights, and ONE collective -- an
 * NCCL all-gather of the per-rank logits, padded to t_max rows -- followed by a row gather restores the original
 * packed order on every rank.  (The reference has no multi-GPU inference path)
 * NCCL is bound at run time (libnccl.so.2; override with ESMK_NCCL_LIB), one communicator per process / GPU.
 *   esmk_comm_unique_id : rank 0 fills 128 bytes, the caller distributes them (torch.distributed store, MPI, a file)
 *   esmk_comm_create    : collective over all ranks, on the calling thread's current device
 *   esmk_allgather_logits: local [t_max, V] bf16 (rows beyond the rank's own tokens are don't-care) ->
 *       gathered [world * t_max, V] (scratch) -> out [T, V] with out[t] = gathered[perm[t]]; perm int64[T] on the
 *       device, perm[t] = owner_rank(t) * t_max + row of token t inside the owner's share.  out / perm may be NULL
 *       to stop after the all-gather. */
typedef struct esmk_comm esmk_comm_t;
ESMK_API int esmk_comm_unique_id(void* id128);
ESMK_API int esmk_comm_create(esmk_comm_t** out, int world_size, int rank, const void* id128);
ESMK_API void esmk_comm_destroy(esmk_comm_t* comm);
ESMK_API int esmk_allgather_logits(esmk_comm_t* comm, const void* local, int t_max, int V, const int64_t* perm, int T,
                                   void* gathered, void* out, esmk_stream_t stream);
/* The same collective over NVLink / NVSwitch peer memory, without NCCL on the data path ("direct peer
 * stores ... into a symmetric buffer").
 *   esmk_comm_enable_peer     : collective over all ranks (one node); allocates this rank's window -- two buffers of
 *       buffer_bytes >= T*V*2 each -- exchanges CUDA IPC handles and maps every peer's window.  Non-zero where CUDA IPC
 *       or peer access is unavailable (same outcome on every rank): keep using esmk_allgather_logits then.
 *   esmk_peer_allgather_logits: local [rows, V] bf16 = this rank's rows, dest_rows int32[rows] (device) = packed row of
 *       each -> out [T, V] holding all ranks' rows in packed order.  One kernel stores the rows at their final position
 *       in EVERY rank's window and publishes a per-rank flag (system-scope release); a one-block kernel waits for all
 *       ranks' flags; the window is then copied to out.  All calls of one communicator must be issued on ONE stream, in
 *       the same order on every rank; a peer that never arrives sets ESMK_ASYNC_PEER_TIMEOUT after 30 s.
 *   esmk_comm_disable_peer    : collective; unmaps the peers' windows, meets all ranks, then frees this rank's window
 *       (CUDA leaves freeing exported memory that a peer still maps undefined).  Call it before esmk_comm_destroy;
 *       a communicator destroyed with its windows up leaves its own window to process exit. */
ESMK_API int esmk_comm_enable_peer(esmk_comm_t* comm, size_t buffer_bytes);
ESMK_API int esmk_comm_disable_peer(esmk_comm_t* comm);
ESMK_API int esmk_peer_allgather_logits(esmk_comm_t* comm, const void* local, int rows, int V, const int32_t* dest_rows,
                                        int T, void* out, esmk_stream_t stream);

/* ---- per-kernel-family device timing (measurement only) ------------------------- */
enum esmk_prof_category {
  ESMK_PROF_MISC = 0,          /* batch metadata, rope tables, embedding gather */
  ESMK_PROF_LAYERNORM = 1,
  ESMK_PROF_GEMM_QKV = 2,
  ESMK_PROF_ROPE = 3,          /* stand-alone QK-LayerNorm + RoPE kernel (ESMC) */
  ESMK_PROF_ATTENTION = 4,
  ESMK_PROF_GEMM_OUT = 5,
  ESMK_PROF_GEMM_FFN_UP = 6,
  ESMK_PROF_GEMM_FFN_DOWN = 7,
  ESMK_PROF_HEAD = 8,          /* LM head (2 GEMMs + LayerNorm + softmax) */
  ESMK_PROF_DEQUANT = 9,       /* quantised weight -> bf16 scratch */
  ESMK_PROF_COUNT = 10
};
/* When enabled, esmk_forward brackets every launch with CUDA events on its stream.
 * esmk_profile_read (after the caller synchronised the stream) returns the summed
 * milliseconds and launch counts per category since the last read, and resets. */
ESMK_API void esmk_profile_enable(int on);
ESMK_API int esmk_profile_read(float* ms, int* launches, int n_categories);

#ifdef __cplusplus
}
#endif
#endif /* ESMK_H */
