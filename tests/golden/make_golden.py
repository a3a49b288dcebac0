"""Generate the committed golden fixtures by running the REAL reference.

Run once in the build container (it needs /root/reference, which does not exist
on the GPU box):

    python tests/golden/make_golden.py

What it does
  * imports `/root/reference/esme` unmodified, with two in-process shims:
      - a stub `accelerate` module (esme/esm.py:3 imports it at top level; it is
        only used by the loader, not by any arithmetic),
      - `esme.attention.flash_attn_varlen_func` replaced by a per-sequence
        torch SDPA (the reference has no CPU attention path, esme/attention.py:115),
  * runs the reference modules on CPU in bf16 and stores inputs + outputs as
    .npz (bf16 tensors are stored as raw uint16 bit patterns),
  * copies the two data fixtures of the reference test-suite (real ESM2-8M
    weights and test.fa) next to them.
Nothing under tests/ reads /root/reference at test time.
"""
import contextlib
import json
import os
import shutil
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'

# ---- shims -----------------------------------------------------------------
acc = types.ModuleType('accelerate')
acc.init_empty_weights = contextlib.nullcontext
acc.load_checkpoint_and_dispatch = None
sys.modules['accelerate'] = acc
tm = types.ModuleType('torchmetrics')                       # esme/variant.py:5 imports it at module level
tmt = types.ModuleType('torchmetrics.text')
tmt.Perplexity = object
tm.text = tmt
sys.modules['torchmetrics'] = tm
sys.modules['torchmetrics.text'] = tmt
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import esme.attention as ref_attention            # noqa: E402
from esme.esm import ESM2, ESMC                   # noqa: E402
from esme.alphabet import tokenize, tokenize_unpad, Alphabet, Alphabet3, padding_mask  # noqa: E402
from esme.rotary import RotaryEmbedding           # noqa: E402
from safetensors.torch import load_file           # noqa: E402

assert ref_attention.__file__.startswith(REF)


def sdpa_varlen(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k,
                dropout_p=0.0, softmax_scale=None, causal=False, **kw):
    assert not causal and dropout_p == 0.0 and softmax_scale is None
    out = torch.empty_like(q)
    cu = cu_seqlens_q.tolist()
    for a, b in zip(cu[:-1], cu[1:]):
        o = torch.nn.functional.scaled_dot_product_attention(
            q[a:b].transpose(0, 1)[None], k[a:b].transpose(0, 1)[None], v[a:b].transpose(0, 1)[None])
        out[a:b] = o[0].transpose(0, 1)
    return out


ref_attention.flash_attn_varlen_func = sdpa_varlen

from oracle import esm_oracle as O                # noqa: E402


def f32(t):
    """bf16 tensors are stored as their raw 16-bit patterns (dtype uint16) to keep the
    fixtures small; fp32 tensors are stored as-is.  tests/conftest.py::load_golden decodes."""
    t = t.detach()
    if t.dtype == torch.bfloat16:
        return t.contiguous().view(torch.int16).numpy().view(np.uint16)
    return t.to(torch.float32).numpy()


def build(cls, path):
    model = cls.create_model(path)
    missing, unexpected = model.load_state_dict(load_file(path), strict=True)
    assert not missing and not unexpected
    return model.eval()


def run_packed(model, tokens, cu, max_len, taps=False):
    out = {}
    hooks = []
    if taps:
        for i, layer in enumerate(model.layers):
            def tap_layer(m, a, o, i=i):
                out[f'layer{i}.x_out'] = f32(o)

            def tap_rot(m, a, o, i=i):
                out[f'layer{i}.q_rot'] = f32(o[0])
                out[f'layer{i}.k_rot'] = f32(o[1])

            hooks.append(layer.register_forward_hook(tap_layer))
            hooks.append(layer.self_attn.rot_emb.register_forward_hook(tap_rot))
    with torch.no_grad():
        logits = model(tokens, (cu, max_len))
        logp = model.predict_log_prob(tokens, (cu, max_len))
        rep = model.forward_representation(tokens, (cu, max_len))
    for h in hooks:
        h.remove()
    out.update(tokens=tokens.numpy(), cu_lens=cu.numpy(), max_len=np.int64(max_len),
               logits=f32(logits), log_prob=f32(logp), representation=f32(rep))
    return out


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)

    # ---- data fixtures of the reference test-suite -------------------------
    shutil.copyfile(f'{REF}/tests/data/8M.safetensors', f'{HERE}/esm2_8m.safetensors')
    shutil.copyfile(f'{REF}/tests/data/test.fa', f'{HERE}/test.fa')

    # ---- config 1: ESM2-8M, 2 x 64 residues, seed 0 -------------------------
    model = build(ESM2, f'{HERE}/esm2_8m.safetensors')
    g = torch.Generator().manual_seed(0)
    seqs = [''.join(Alphabet.amino_acids[int(i)] for i in torch.randint(0, 20, (64,), generator=g))
            for _ in range(2)]
    tokens, indices, cu, max_len = tokenize_unpad(seqs)
    d = run_packed(model, tokens, cu, max_len, taps=True)
    with torch.no_grad():
        d['padded_logits'] = f32(model(tokenize(seqs)))
    d['seqs'] = np.array(seqs)
    np.savez_compressed(f'{HERE}/esm2_8m_cfg1.npz', **d)

    # ---- ragged: test.fa (16 proteins) + a masked/unknown/padded case --------
    fa = [''.join(rec.split('\n')[1:]) for rec in open(f'{HERE}/test.fa').read().split('>') if rec.strip()]
    assert [len(x) for x in fa] == [256, 320, 458, 156, 438, 60, 217, 204, 352, 75, 128, 447, 347, 948, 85, 137]
    tokens, indices, cu, max_len = tokenize_unpad(fa)
    d = run_packed(model, tokens, cu, max_len)
    d['indices'] = indices.numpy()
    np.savez_compressed(f'{HERE}/esm2_8m_testfa.npz', **d)

    mixed = [fa[5][:40] + '<mask>' + fa[5][41:], 'MKT', fa[9], 'A', 'MXBZJ<mask>K']
    tok2d = tokenize(mixed)
    with torch.no_grad():
        pl = model(tok2d)
        plp = model.predict_log_prob(tok2d)
    np.savez_compressed(f'{HERE}/esm2_8m_padded.npz', tokens=tok2d.numpy(), logits=f32(pl),
                        log_prob=f32(plp), seqs=np.array(mixed))

    # ---- synthetic small models that exercise hd=64 kernels ------------------
    for family, cls, dims, seed, lens in [
        ('esm2', ESM2, (2, 256, 4), 11, [300, 131, 66, 2, 129, 257]),
        ('esmc', ESMC, (2, 192, 3), 12, [200, 70, 2, 130]),
    ]:
        cfg = O.OracleConfig(family, *dims)
        W = O.synthetic_weights(cfg, seed=seed, qk_gain=4.0)
        path = f'{HERE}/{family}_tiny.safetensors'
        O.save_checkpoint(path, cfg, W, tag='tiny')
        m = build(cls, path)
        tokens, cu, max_len = O.synthetic_batch(lens, seed=seed + 100)
        tokens[5] = 32  # one <mask> token
        d = run_packed(m, tokens, cu, max_len, taps=True)
        np.savez_compressed(f'{HERE}/{family}_tiny.npz', **d)

    # ---- RoPE module, fp32, the reference test's shape (tests/test_rotary.py:103-136)
    g = torch.Generator().manual_seed(7)
    q = torch.rand(280, 8, 64, generator=g)
    k = torch.rand(280, 8, 64, generator=g)
    cu = torch.tensor([0, 60, 100, 280], dtype=torch.int32)
    rot = RotaryEmbedding(dim=64)
    q_r, k_r = rot(q, k, cu, 180)
    qb, kb = rot(q.bfloat16(), k.bfloat16(), cu, 180)
    np.savez_compressed(f'{HERE}/rope.npz', q=q.numpy(), k=k.numpy(), cu_lens=cu.numpy(), max_len=np.int64(180),
                        q_rot=q_r.numpy(), k_rot=k_r.numpy(), q_rot_bf16=f32(qb), k_rot_bf16=f32(kb),
                        cos_bf16=f32(rot._cos_cached), sin_bf16=f32(rot._sin_cached))

    # ---- masked-marginal variant scores (esme/variant.py:110-165), 8M model, short + windowed ------
    from esme.variant import predict_mask_margin as ref_pmm, MaskMarginDataset as RefDS
    model8 = build(ESM2, f'{HERE}/esm2_8m.safetensors')
    vs = {}
    for name, seq, ml in (('short', 'MPEAAPPVAPAPAAPTW', None), ('windowed', fa[9], 40)):
        df = ref_pmm(model8, seq, batch_size=8, max_len=ml)
        vs[name] = dict(seq=seq, max_len=ml, variants=list(df.index), scores=[float(x) for x in df['score']])
    ds = RefDS(fa[9], max_len=40)
    vs['windows'] = [[int(ds[i]['local_pos']), int(ds[i]['pos']), int(ds[i]['wt_token']),
                      ds[i]['token'].tolist()] for i in (0, 1, 19, 20, 21, 40, len(fa[9]) - 1)]
    json.dump(vs, open(f'{HERE}/mask_margin_8m.json', 'w'))

    # ---- tokenizer known answers -------------------------------------------
    p53 = ('MEEPQSDPSVEPPLSQETFSDLWKLLPENNVLSPLPSQAMDDLMLSPDDIEQWFTEDPGPDEAPRMPEAAPPVAPAPAAPTPAAPAPAPSWPLSSSVPSQKTYQGSYGFRLGFLHSGTAKSVTCTYSPALNKMFCQLAKTCPVQLWVDSTPPPGTRVRAMAIYKQSQHMTEVVRRCPHHERCSDSDGLAPPQHLIRVEGNLRVEYLDDRNTFRHSVVVPYEPPEVGSDCTTIHYNYMCNSSCMGGMNRRPILTIITLEDSSGNLLGRNSFEVRVCACPGRDRRTEEENLRKKGEPHHELPPGSTKRALPNNTSSSPQPKKKPLDGEYFTLQIRGRERFEMFRELNEALELKDAQAGKEPGGSRAHSSHLKSKKGQSTSRHKKLMFKTEGPDSD')
    calm1 = 'MADQLTEEQIAEFKEAFSLFDKDGDGTITTKELGTVMRSLGQNPTEAELQDMINEVDADGNGTIDFPEFLTMMARKMKDTDSEEEIREAFRVFDKDGNGYISAAELRHVMTNLGEKLTDEEVDEMIREADIDGDGQVNYEEFVQMMTAK'
    cases = {}
    for name, seqs, alph in [
        ('p53', p53, Alphabet3),
        ('p53_list', [p53, p53 + p53, calm1], Alphabet3),
        ('mask_unknown', ['MK<mask>TJ?', 'A', 'ACDEFGHIKLMNPQRSTVWYXBUZO.-'], Alphabet3),
        ('esm2_alphabet', ['MK<mask>T<null_1>|', 'AC'], Alphabet),
    ]:
        tp = tokenize(seqs, alph)
        tu, iu, cu, ml = tokenize_unpad(seqs, alph)
        cases[name] = dict(seqs=seqs, alphabet='esm2' if alph is Alphabet else 'esmc',
                           padded=tp.tolist(), tokens=tu.tolist(), indices=iu.tolist(),
                           cu_lens=cu.tolist(), max_len=int(ml),
                           padding_mask=padding_mask(cu, ml).int().tolist())
    json.dump(cases, open(f'{HERE}/tokenizer.json', 'w'))
    print('golden fixtures written to', HERE)


if __name__ == '__main__':
    main()
