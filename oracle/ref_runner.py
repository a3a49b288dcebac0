"""Run the UNMODIFIED reference package (oracle/_ref/esme_ref.zip, see oracle/build_ref.py) in its own process.
TEST INFRASTRUCTURE ONLY -- used by tests/ (parity against the real reference + flash-attn on the GPU) and by
bench.py (the reference timed beside the product: `gpu_reference` block and the `--impl reference` CPU arm).

The product package is also called `esme`, so the reference can only be imported in a separate interpreter:

    python oracle/ref_runner.py job.json        (reads the job, writes job['out'], prints one JSON line)

job = {
  "device": "cuda" | "cpu",
  "family": "esm2" | "esmc", "num_layers": n, "embed_dim": D, "attention_heads": H,
  "weights": {"safetensors": path}  |  {"synthetic_seed": s},     # synthetic = esme/synthetic.py of the product,
                                                                   # loaded by file path (pure torch, no package import)
  "batch": path to .npz with tokens int64[T], cu_lens int32[B+1], max_len        (mode forward)
         | {"lens": [...], "seed": s}                                             (synthetic batch, both modes)
  "mode": "forward" -> writes logits / log_prob / representation (bf16 bit patterns as uint16) to job["out"] (.npz)
          "time"    -> {"ms_per_step": ..} for `steps` forwards after `warmup` (CUDA events on GPU, wall clock on CPU)
  "method": "forward" | "predict_log_prob",
  "steps", "warmup", "threads" (CPU),
  "sample": {"target_seconds": s, "min_tokens": a, "max_tokens": b}   (mode time, optional): time a PREFIX of the
          batch (whole sequences) sized by a ~256-token pilot so that one step costs about target_seconds
}

On "cpu" the single symbol esme.attention.flash_attn_varlen_func is replaced by a per-sequence torch SDPA: the
reference has no CPU attention path (esme/attention.py:115-123 calls the flash-attn CUDA op unconditionally).
Everything else -- modules, layer loop, rotary embedding with its host syncs, LM head -- is the reference's own code.
"""
import importlib.util
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _import_reference():
    ref_root = os.path.join(HERE, '_ref', 'esme_ref.zip')        # byte-compiled reference package (zipimport)
    if not os.path.isfile(ref_root):
        raise SystemExit('oracle/_ref/esme_ref.zip is absent: run `python oracle/build_ref.py` where /root/reference exists')
    # the reference first, the import stubs second; the product package must not be importable from here
    sys.path[:] = [ref_root, os.path.join(HERE, 'ref_shims')] + \
        [p for p in sys.path if os.path.abspath(p or '.') not in (ROOT, os.path.join(ROOT, 'esm-efficient_b200'))]
    import esme.attention as ref_attention
    from esme.esm import ESM2, ESMC
    assert os.path.abspath(ref_attention.__file__).startswith(ref_root), ref_attention.__file__
    return ref_attention, {'esm2': ESM2, 'esmc': ESMC}


def _synthetic_module():
    spec = importlib.util.spec_from_file_location(
        '_esmk_synthetic', os.path.join(ROOT, 'esm-efficient_b200', 'esme', 'synthetic.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _sdpa_varlen(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k,
                 dropout_p=0.0, softmax_scale=None, causal=False, **kw):
    import torch
    assert not causal and dropout_p == 0.0 and softmax_scale is None
    out = torch.empty_like(q)
    cu = cu_seqlens_q.tolist()
    for a, b in zip(cu[:-1], cu[1:]):
        o = torch.nn.functional.scaled_dot_product_attention(
            q[a:b].transpose(0, 1)[None], k[a:b].transpose(0, 1)[None], v[a:b].transpose(0, 1)[None])
        out[a:b] = o[0].transpose(0, 1)
    return out


def _bits(t):
    import numpy as np
    import torch
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy().view(np.uint16)
    return t.numpy()


def main():
    job = json.load(open(sys.argv[1]))
    import numpy as np
    import torch
    ref_attention, classes = _import_reference()
    dev = torch.device(job['device'])
    if dev.type == 'cpu':
        ref_attention.flash_attn_varlen_func = _sdpa_varlen
        torch.set_num_threads(int(job.get('threads') or os.cpu_count() or 1))
    family, layers, D, H = job['family'], job['num_layers'], job['embed_dim'], job['attention_heads']
    synth = None
    w = job['weights']
    if 'safetensors' in w:
        from safetensors.torch import load_file
        state = load_file(w['safetensors'])
    else:
        synth = _synthetic_module()
        state = synth.synthetic_state_dict(family, layers, D, seed=int(w['synthetic_seed']))
    model = classes[family](num_layers=layers, embed_dim=D, attention_heads=H)
    missing, unexpected = model.load_state_dict(state, strict=True)
    assert not missing and not unexpected
    model = model.to(dev).eval().requires_grad_(False)

    b = job['batch']
    if isinstance(b, str):
        z = np.load(b)
        tokens, cu, max_len = torch.from_numpy(z['tokens']), torch.from_numpy(z['cu_lens']), int(z['max_len'])
    else:
        synth = synth or _synthetic_module()
        tokens, cu, max_len = synth.synthetic_batch(b['lens'], seed=int(b['seed']))
    fn = getattr(model, job.get('method', 'forward'))

    def prefix(limit):
        c = cu.tolist()
        n = 1
        while n < len(c) - 1 and c[n + 1] <= limit:
            n += 1
        lens_ = [c[i + 1] - c[i] for i in range(n)]
        return tokens[:c[n]], cu[:n + 1].clone(), max(lens_), n

    sample_note = None
    if job['mode'] == 'time' and job.get('sample'):
        sm = job['sample']
        with torch.no_grad():
            t_, c_, m_, _ = prefix(256)
            t0 = time.perf_counter()
            fn(t_.to(dev), (c_.to(dev), m_))
            if dev.type == 'cuda':
                torch.cuda.synchronize()
            rate = t_.numel() / (time.perf_counter() - t0)
        limit = int(min(max(rate * float(sm['target_seconds']), sm.get('min_tokens', 512)), sm.get('max_tokens', 8192)))
        tokens, cu, max_len, nseq = prefix(limit)
        sample_note = f'first {nseq} sequences ({tokens.numel()} tokens) of the batch'
    tokens, cu = tokens.to(dev), cu.to(dev)
    res = {'device': str(dev), 'tokens': int(tokens.numel()), 'cores': os.cpu_count(), 'reference_file': ref_attention.__file__,
           'attention': 'flash_attn_varlen_func (flash-attn wheel)' if dev.type == 'cuda' else 'per-sequence torch SDPA'}

    with torch.no_grad():
        if job['mode'] == 'forward':
            out = {'logits': _bits(model(tokens, (cu, max_len))),
                   'log_prob': _bits(model.predict_log_prob(tokens, (cu, max_len)))}
            if family == 'esm2':       # (the reference's ESMC.forward_representation has an assertion bug, esm.py:873)
                out['representation'] = _bits(model.forward_representation(tokens, (cu, max_len)))
            if job.get('tokens2d') is not None:
                t2 = torch.from_numpy(np.load(job['tokens2d'])['tokens2d']).to(dev)
                out['logits_padded'] = _bits(model(t2))
            np.savez(job['out'], **out)
        else:
            steps, warmup = int(job.get('steps', 3)), int(job.get('warmup', 1))
            for _ in range(warmup):
                y = fn(tokens, (cu, max_len))
            if dev.type == 'cuda':
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    y = fn(tokens, (cu, max_len))
                e1.record()
                torch.cuda.synchronize()
                res['ms_per_step'] = e0.elapsed_time(e1) / steps
            else:
                t0 = time.perf_counter()
                for _ in range(steps):
                    y = fn(tokens, (cu, max_len))
                res['ms_per_step'] = (time.perf_counter() - t0) / steps * 1e3
                res['threads'] = torch.get_num_threads()
            res['steps'], res['warmup'], res['sample'] = steps, warmup, sample_note
            res['finite'] = bool(torch.isfinite(y.float()).all())
            if job.get('out'):
                np.savez(job['out'], out=_bits(y))
    print('@@REF@@' + json.dumps(res))


if __name__ == '__main__':
    main()
