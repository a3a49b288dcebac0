"""The reference's GPU op sequence restated call for call with the libraries it uses (torch bf16 ops / cuBLAS +
flash_attn_varlen_func).  TEST INFRASTRUCTURE ONLY: the timing fall-back of bench.py's `gpu_reference` block and of
tests/gpu_diag.py when the real package (oracle/_ref) is absent; ESM2 family only."""


def reference_gpu_forward(W, layers, D, H, tokens, cu_lens, max_len):
    """The reference's GPU op sequence restated call for call with the libraries it uses (torch bf16 ops /
    cuBLAS + flash_attn_varlen_func): esme/esm.py:176-282, attention.py:91-139,241-255, rotary.py:5-43, head.py:25-27.
    Context number only (the real package cannot travel to the GPU box)."""
    import torch
    import torch.nn.functional as F
    from flash_attn import flash_attn_varlen_func
    hd = D // H
    x = F.embedding(tokens, W['embed_tokens.weight'])
    x.masked_fill_((tokens == 32).unsqueeze(-1), 0.0)
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, hd, 2, device=x.device, dtype=torch.float32) / hd))
    freqs = torch.outer(torch.arange(max_len, device=x.device, dtype=torch.float32), inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    cos, sin = emb.cos().to(x.dtype), emb.sin().to(x.dtype)

    def culen_indices(cu):                      # esme/rotary.py:5-14 (host syncs included, as in the reference)
        lengths = cu[1:] - cu[:-1]
        starts = torch.cat([torch.tensor([0], device=cu.device), lengths.cumsum(0)[:-1]])
        ids = torch.repeat_interleave(torch.arange(len(lengths), device=cu.device), lengths)
        return torch.arange(cu[-1], device=cu.device) - starts[ids]

    def rot(t):
        idx = culen_indices(cu_lens)
        t1, t2 = t.chunk(2, dim=-1)
        return t * cos[idx].unsqueeze(1) + torch.cat((-t2, t1), dim=-1) * sin[idx].unsqueeze(1)

    T = tokens.numel()
    for i in range(layers):
        g = lambda k: W[f'layers.{i}.{k}']
        h = F.layer_norm(x, (D,), g('self_attn.norm.weight'), g('self_attn.norm.bias'))
        q = F.linear(h, g('self_attn.q.weight'), g('self_attn.q.bias')).view(T, H, hd)
        k = F.linear(h, g('self_attn.k.weight'), g('self_attn.k.bias')).view(T, H, hd)
        v = F.linear(h, g('self_attn.v.weight'), g('self_attn.v.bias')).view(T, H, hd)
        q, k = rot(q), rot(k)
        a = flash_attn_varlen_func(q, k, v, cu_lens, cu_lens, max_len, max_len, dropout_p=0.0, causal=False)
        x = x + F.linear(a.reshape(T, D), g('self_attn.out.weight'), g('self_attn.out.bias')) / 1.0
        f = F.layer_norm(x, (D,), g('final.0.weight'), g('final.0.bias'))
        f = F.linear(F.gelu(F.linear(f, g('final.1.weight'), g('final.1.bias'))), g('final.3.weight'), g('final.3.bias'))
        x = x + f / 1.0
    x = F.layer_norm(x, (D,), W['emb_layer_norm_after.weight'], W['emb_layer_norm_after.bias'])
    y = F.layer_norm(F.gelu(F.linear(x, W['lm_head.dense.weight'], W['lm_head.dense.bias'])), (D,),
                     W['lm_head.layer_norm.weight'], W['lm_head.layer_norm.bias'])
    return F.linear(y, W['lm_head.final.weight'], W['lm_head.final.bias'])
