"""ctypes binding of libesmk.so (C ABI declared in include/esmk.h).

The library is the only compute backend of this package: if it cannot be loaded
the import of this module raises, and every operator raises when handed a
non-CUDA tensor.  There is no eager / CPU fallback.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('ESMK_LIB_PATH') or os.path.join(_HERE, '_lib', 'libesmk.so')   # override: debug builds (tracing)


class EsmkError(RuntimeError):
    pass


if not os.path.isfile(LIB_PATH):
    raise ImportError(
        f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
        f'(or `make -C esm-efficient_b200/csrc`). This package has no fallback backend.')

lib = C.CDLL(LIB_PATH)

c_void_p, c_int, c_float, c_size_t = C.c_void_p, C.c_int, C.c_float, C.c_size_t


class GemmArgs(C.Structure):
    _fields_ = [
        ('A', c_void_p), ('lda', c_int),
        ('W', c_void_p), ('bias', c_void_p),
        ('C', c_void_p), ('ldc', c_int),
        ('M', c_int), ('N', c_int), ('K', c_int),
        ('epilogue', c_int),
        ('R', c_void_p), ('ldr', c_int), ('residue_scaling', c_float),
        ('rope_cos', c_void_p), ('rope_sin', c_void_p), ('pos', c_void_p),
        ('head_dim', c_int), ('rope_cols', c_int),
    ]


class Config(C.Structure):
    _fields_ = [
        ('family', c_int), ('num_layers', c_int), ('embed_dim', c_int), ('attention_heads', c_int),
        ('ffn_dim', c_int), ('vocab', c_int), ('embed_rows', c_int), ('residue_scaling', c_float),
        ('no_rotary', c_int), ('pos_rows', c_int),
    ]


class QWeight(C.Structure):
    """esmk_qweight: optional quantised storage of one weight (data NULL = not quantised)."""
    _fields_ = [('data', c_void_p), ('scale', c_void_p), ('bits', c_int)]


class LayerWeights(C.Structure):
    _fields_ = [(n, c_void_p) for n in (
        'attn_norm_w', 'attn_norm_b', 'wqkv', 'bqkv', 'qln_w', 'kln_w', 'wo', 'bo',
        'ffn_norm_w', 'ffn_norm_b', 'w1', 'b1', 'w2', 'b2')] + \
        [(n, QWeight) for n in ('q_wqkv', 'q_wo', 'q_w1', 'q_w2')]


class Weights(C.Structure):
    _fields_ = [
        ('embed', c_void_p), ('layers', C.POINTER(LayerWeights)),
        ('final_norm_w', c_void_p), ('final_norm_b', c_void_p),
        ('head_dense_w', c_void_p), ('head_dense_b', c_void_p),
        ('head_norm_w', c_void_p), ('head_norm_b', c_void_p),
        ('head_final_w', c_void_p), ('head_final_b', c_void_p),
        ('pos_embed', c_void_p), ('pre_norm_w', c_void_p), ('pre_norm_b', c_void_p),
    ]


EPI_BIAS, EPI_BIAS_GELU, EPI_RESIDUAL, EPI_QKV_ROPE, EPI_SWIGLU = range(5)
OUT_LOGITS, OUT_LOG_PROB, OUT_PROB, OUT_REPRESENTATION = range(4)

# every symbol include/esmk.h declares: name -> (restype, argtypes)
SIGNATURES = {
    'esmk_last_error': (C.c_char_p, []),
    'esmk_version': (c_int, []),
    'esmk_launch_count': (C.c_uint64, []),
    'esmk_async_error': (c_int, []),
    'esmk_tile_capacity': (c_int, [c_int, c_int]),
    'esmk_batch_meta': (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'esmk_rope_tables': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    'esmk_embed': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'esmk_unpad_tokens': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'esmk_pad_rows': (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    'esmk_add_positions': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'esmk_layernorm': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p]),
    'esmk_qk_norm_rope': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p]),
    'esmk_mean_pool': (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    'esmk_residual_add': (c_int, [c_void_p, c_void_p, c_void_p, C.c_long, c_float, c_void_p]),
    'esmk_softmax': (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'esmk_gemm': (c_int, [C.POINTER(GemmArgs), c_void_p]),
    'esmk_attn_pool': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                             c_void_p]),
    'esmk_quantize': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'esmk_dequantize': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    'esmk_attn_varlen': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                 c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'esmk_model_create': (c_int, [C.POINTER(Config), C.POINTER(Weights), C.POINTER(c_void_p)]),
    'esmk_model_destroy': (None, [c_void_p]),
    'esmk_workspace_bytes': (c_size_t, [c_void_p, c_int, c_int, c_int]),
    'esmk_forward': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t,
                             c_int, c_void_p, c_void_p, c_void_p]),
    'esmk_comm_unique_id': (c_int, [c_void_p]),
    'esmk_comm_create': (c_int, [C.POINTER(c_void_p), c_int, c_int, c_void_p]),
    'esmk_comm_destroy': (None, [c_void_p]),
    'esmk_allgather_logits': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    'esmk_comm_enable_peer': (c_int, [c_void_p, c_size_t]),
    'esmk_comm_disable_peer': (c_int, [c_void_p]),
    'esmk_peer_allgather_logits': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    'esmk_profile_enable': (None, [c_int]),
    'esmk_profile_read': (c_int, [C.POINTER(c_float), C.POINTER(c_int), c_int]),
    'esmk_lm_head': (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_int, c_void_p, c_void_p]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here == ABI mismatch, fail at import
    _fn.restype = _res
    _fn.argtypes = _args


def check(rc: int, what: str):
    if rc != 0:
        msg = lib.esmk_last_error()
        raise EsmkError(f'{what} failed (rc={rc}): {msg.decode() if msg else "?"}')


PROF_CATEGORIES = ('misc', 'layernorm', 'gemm_qkv', 'rope', 'attention', 'gemm_out', 'gemm_ffn_up',
                   'gemm_ffn_down', 'head', 'dequant')


def profile_enable(on: bool):
    lib.esmk_profile_enable(int(on))


def profile_read() -> dict:
    """{category: (milliseconds, launches)} since the last read; synchronise the stream first."""
    n = len(PROF_CATEGORIES)
    ms, cnt = (c_float * n)(), (c_int * n)()
    check(lib.esmk_profile_read(ms, cnt, n), 'esmk_profile_read')
    return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(PROF_CATEGORIES)}


def launch_count() -> int:
    return int(lib.esmk_launch_count())
