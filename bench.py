#!/usr/bin/env python
"""Benchmark of the ESM forward hot path (BASELINE.json metric: residues/s, ESM2-650M forward).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--model esm2_650m]

One "step" = one forward pass (tokens -> logits) over one packed batch of synthetic
protein sequences (BASELINE config 2: <= 50,000 tokens per GPU, lognormal lengths,
seeded synthetic bf16 weights).  For N > 1 (launched by torch.distributed.run, one rank
per GPU) the global batch of N x 50k tokens is split by whole sequences across the ranks
(esme.parallel), every rank runs the forward on its share and the logits are all-gathered
(the one collective of the path): weak scaling.

Prints ONE JSON line on rank 0.  `value` is device-timed with the batch resident in HBM;
`e2e` repeats the measurement through the public API with pinned HOST buffers (H2D of the
tokens and D2H of the logits inside the timed region).  `--impl reference` times the
reference's algorithm on the host cores (the oracle port; the reference itself has no CPU
path and cannot travel to the GPU box) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'esm-efficient_b200'))

MODELS = {
    'esm2_8m': ('esm2', 6, 320, 20),
    'esm2_650m': ('esm2', 33, 1280, 20),
    'esm2_3b': ('esm2', 36, 2560, 40),
    'esmc_300m': ('esmc', 30, 960, 15),
}
METRIC = 'residues_per_sec_forward'
UNIT = 'residues/s'


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(path):
        d = json.load(open(path))
        return dict(bf16_tflops=d['bf16_tflops'], bf16_tflops_sustained=d['bf16_tflops_sustained'],
                    hbm_gbs=d['hbm_gbs'], source='MEASURED_PEAKS.json (of measured)')
    return dict(bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, hbm_gbs=6650.0,
                source='B200_PROFILING.md fallback (of fallback)')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ''
        sm, smax, reasons = [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['no samples'])
        busy = [c for c in sm if c >= 0.5 * max(sm)] or sm
        return dict(sm_mhz=statistics.median(busy), sm_max_mhz=max(smax), reasons=sorted(reasons), samples=len(sm))


def build_workload(args, rank, world):
    """Global batch = world x budget tokens; returns this rank's share (host tensors) + bookkeeping."""
    import torch
    from esme import parallel, synthetic
    family, layers, D, H = MODELS[args.model]
    dist_name = 'loguniform' if family == 'esmc' else 'lognormal'
    lens = []
    for r in range(world):                       # one 50k-token draw per GPU -> weak scaling
        lens += synthetic.synthetic_lengths(args.tokens, seed=2 + 10 * r, dist=dist_name)
    tokens, cu, max_len = synthetic.synthetic_batch(lens, seed=3)
    if world > 1:
        owned = parallel.partition_sequences(lens, world, D)
        shares = [parallel.take_sequences(tokens, cu, o) for o in owned]
        imb = parallel.imbalance(lens, owned, D)
    else:
        owned = [list(range(len(lens)))]
        shares = [(tokens, cu, max_len, torch.arange(tokens.numel()))]
        imb = 1.0
    return dict(lens=lens, tokens=tokens, cu=cu, max_len=max_len, shares=shares, owned=owned, imbalance=imb)


def bounded_cpu_sample(wl, sample_tokens):
    import torch
    cu = wl['cu'].tolist()
    n = 1
    while n < len(cu) - 1 and cu[n + 1] <= sample_tokens:
        n += 1
    T = cu[n]
    lens = [cu[i + 1] - cu[i] for i in range(n)]
    return wl['tokens'][:T], wl['cu'][:n + 1].clone(), max(lens), T, n


def run_cpu_reference(args, wl, W, steps, warmup, target_seconds=4.0):
    """The reference algorithm on the host cores: oracle port, bf16-faithful mode, all threads."""
    import torch
    from oracle import esm_oracle as O
    family, layers, D, H = MODELS[args.model]
    cfg = O.OracleConfig(family, layers, D, H)
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    # pilot on ~256 tokens to size the bounded sample (~target_seconds of CPU work per step)
    tok, cu, ml, T0, _ = bounded_cpu_sample(wl, 256)
    t0 = time.perf_counter()
    O.forward_packed(cfg, W, tok, cu, ml, 'bf16')
    pilot = time.perf_counter() - t0
    rate = T0 / pilot
    sample_tokens = int(min(max(rate * target_seconds, 512), 8192))
    tok, cu, ml, T, nseq = bounded_cpu_sample(wl, sample_tokens)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.forward_packed(cfg, W, tok, cu, ml, 'bf16')
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return dict(value=T / sec, unit=UNIT, cores=cores, kind='port',
                sample=f'first {nseq} sequences ({T} tokens) of the same batch, {steps} timed forward(s), '
                       f'oracle/esm_oracle.py bf16-faithful mode, torch CPU fp32 GEMMs on {cores} threads'), sec


def bench_mask_margin(args, family, layers, D, H):
    """BASELINE config 5 (bf16 part): masked-marginal sweep of one 1,024-residue protein, batch 32 ->
    32 forwards of 32 x 1,026 tokens; residues/s = 1024 * 1026 / sweep time (single GPU)."""
    import torch
    import esme
    from esme import _lib, synthetic
    from esme.alphabet import Alphabet
    from esme.variant import predict_mask_margin
    assert torch.cuda.is_available()
    dev = torch.device('cuda', 0)
    cls = esme.ESMC if family == 'esmc' else esme.ESM2
    model = cls(layers, D, H)
    model.load_state_dict(synthetic.synthetic_state_dict(family, layers, D, seed=1), strict=True)
    model = model.to(dev).eval().requires_grad_(False)
    g = torch.Generator().manual_seed(6)
    seq = ''.join(Alphabet.amino_acids[int(i)] for i in torch.randint(0, 20, (1024,), generator=g))
    quant_note, score_check = 'bf16 weights', None
    if args.quant != 'none':
        # BASELINE config 5: int4 weight-quantised FFN.  Scores of the bf16 model first, for the comparison.
        from esme.quantization import quantize_model_
        base = predict_mask_margin(model, seq, batch_size=32)
        which = ('ffn',) if args.quant.endswith('ffn') else ('q', 'k', 'v', 'out', 'ffn')
        quantize_model_(model, 4 if args.quant.startswith('4bit') else 8, which=which)
        quant_note = f'{args.quant} weight-only quantised linears (esme/quantization.py formats), GEMMs in bf16 on the dequantised weight'
    for _ in range(max(1, args.warmup // 3)):
        predict_mask_margin(model, seq, batch_size=32)
    torch.cuda.synchronize()
    launches0 = _lib.launch_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        df = predict_mask_margin(model, seq, batch_size=32)        # ends with the single D2H of the scores
    sec = (time.perf_counter() - t0) / args.steps
    if args.quant != 'none':
        a = torch.tensor(base['score'].to_numpy(), dtype=torch.float64)
        b = torch.tensor(df['score'].to_numpy(), dtype=torch.float64)
        ra, rb = a.argsort().argsort().double(), b.argsort().argsort().double()
        score_check = {'spearman_vs_bf16': float(torch.corrcoef(torch.stack((ra, rb)))[0, 1]),
                       'mean_abs_diff_vs_bf16': float((a - b).abs().mean()), 'max_abs_diff_vs_bf16': float((a - b).abs().max())}
    print(json.dumps({
        'metric': 'residues_per_sec_mask_margin_sweep', 'value': 1024 * 1026 / sec, 'unit': UNIT, 'n_gpus': 1,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
        'config': {'workload': f'{args.model} predict_mask_margin, one 1,024-residue protein (seed 6), batch_size 32: '
                               f'32 packed forwards of 32 x 1,026 tokens, LM head on the 1,024 masked rows only, '
                               f'one D2H of the [1024, 20] score matrix; wall-clock incl. host-side DataFrame',
                   'rows': int(df.shape[0]), 'weights': quant_note, 'scores_vs_bf16': score_check},
        'gpu_launches': _lib.launch_count() - launches0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--model', default='esm2_650m', choices=sorted(MODELS))
    ap.add_argument('--tokens', type=int, default=50000, help='token budget per GPU')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--quant', default='none', choices=['none', '4bit-ffn', '4bit', '8bit-ffn', '8bit'],
                    help='weight-only quantisation of the layer linears (mask_margin workload; BASELINE config 5 = 4bit-ffn)')
    ap.add_argument('--workload', default='forward', choices=['forward', 'mask_margin'],
                    help="'mask_margin' = BASELINE config 5: predict_mask_margin sweep over one 1,024-residue protein")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    family, layers, D, H = MODELS[args.model]

    import torch
    from esme import synthetic

    if args.impl == 'reference':
        if rank != 0:
            return
        wl = build_workload(args, 0, 1)
        W = synthetic.synthetic_state_dict(family, layers, D, seed=1)
        steps = max(1, min(args.steps, 5))
        base, sec = run_cpu_reference(args, wl, W, steps, 1)
        print(json.dumps({
            'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': steps, 'warmup': 1, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': f'{args.model} forward, packed batch of the {args.tokens}-token config; '
                                   f'bounded CPU sample: {base["sample"]}'},
            'cpu_baseline': base,
            'e2e': {'value': base['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}))
        return

    if args.workload == 'mask_margin':
        return bench_mask_margin(args, family, layers, D, H)

    # ------------------------------------------------------------------ B200 arm
    import torch.distributed as dist
    import esme
    from esme import _lib, parallel
    assert torch.cuda.is_available(), 'bench.py --impl b200 needs a CUDA device (no CPU fallback exists)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    wl = build_workload(args, rank, world)
    W = synthetic.synthetic_state_dict(family, layers, D, seed=1)
    cls = esme.ESMC if family == 'esmc' else esme.ESM2
    model = cls(layers, D, H)
    model.load_state_dict(W, strict=True)
    model = model.to(dev).eval().requires_grad_(False)
    V = model.lm_head.final.out_features

    tok_h, cu_h, max_len, token_index = wl['shares'][rank]
    T_local, T_global = tok_h.numel(), wl['tokens'].numel()
    tok_pin, cu_pin = tok_h.pin_memory(), cu_h.pin_memory()
    tok_d, cu_d = tok_pin.to(dev), cu_pin.to(dev)
    t_max = max(s[0].numel() for s in wl['shares'])
    gather_buf = torch.empty(world * t_max, V, dtype=torch.bfloat16, device=dev) if world > 1 else None
    local_buf = torch.zeros(t_max, V, dtype=torch.bfloat16, device=dev) if world > 1 else None
    out_pin = torch.empty((world * t_max if world > 1 else T_local, V), dtype=torch.bfloat16).pin_memory()

    # BASELINE config 3 (ESMC) is quoted on predict_log_prob; the ESM2 configs on forward (logits)
    forward = model.predict_log_prob if family == 'esmc' else model.__call__

    def step(tokens, cu):
        logits = forward(tokens, (cu, max_len))
        if world > 1:                       # the single collective of the path: all-gather of logits
            local_buf[:T_local] = logits
            dist.all_gather_into_tensor(gather_buf, local_buf)
            return gather_buf
        return logits

    def e2e_step():
        t = tok_pin.to(dev, non_blocking=True)
        c = cu_pin.to(dev, non_blocking=True)
        out = step(t, c)
        out_pin.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()     # the caller holds the logits on the host

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(args.warmup):
        step(tok_d, cu_d)
    sync_all()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = _lib.launch_count()
    _lib.profile_enable(True)                      # CUDA-event pairs around every launch of the timed steps
    ms_step = timed(lambda: step(tok_d, cu_d), args.steps)
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if sampler else None

    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (tensor bound), measured live over the timed steps
    peaks = measured_peaks()
    fl = synthetic.forward_flops(family, layers, D, [int(x) for x in (cu_h[1:] - cu_h[:-1]).tolist()])
    F = synthetic.ffn_dim(family, D)
    per_launch_flops = {
        'gemm_qkv': 2.0 * T_local * D * 3 * D, 'gemm_out': 2.0 * T_local * D * D,
        'gemm_ffn_up': 2.0 * T_local * D * (F if family == 'esm2' else 2 * F), 'gemm_ffn_down': 2.0 * T_local * F * D,
        'attention': fl['attention'] / layers,
    }
    kernels = {}
    for name, (ms, n) in prof.items():
        if n == 0:
            continue
        k = dict(ms_per_step=ms / args.steps, launches_per_step=n / args.steps, avg_launch_ms=ms / n)
        if name in per_launch_flops:
            k['tflops'] = per_launch_flops[name] / (ms / n) / 1e9
        kernels[name] = k
    gemm_names = [n for n in ('gemm_qkv', 'gemm_out', 'gemm_ffn_up', 'gemm_ffn_down') if n in kernels]
    dominant = max(gemm_names, key=lambda n: kernels[n]['ms_per_step'])
    peak = peaks['bf16_tflops_sustained']
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if os.path.isfile(tpath):
        traffic = json.load(open(tpath)).get(args.model, {}).get(dominant)
    roofline = dict(bound='tensor', kernel=f'esmk gemm_kernel ({dominant})', achieved=kernels[dominant]['tflops'],
                    peak=peak, unit='TFLOP/s', frac=kernels[dominant]['tflops'] / peak, traffic=traffic,
                    peak_source=peaks['source'] + ', bf16_tflops_sustained (kernel timed inside a long step)',
                    algorithmic_flops_per_launch=per_launch_flops[dominant],
                    avg_launch_ms=kernels[dominant]['avg_launch_ms'],
                    all_gemms_tflops=sum(per_launch_flops[n] * kernels[n]['launches_per_step'] for n in gemm_names)
                    / sum(kernels[n]['ms_per_step'] for n in gemm_names) / 1e9,
                    attention_tflops=kernels.get('attention', {}).get('tflops'),
                    attention_frac_of_burst_peak=(kernels['attention']['tflops'] / peaks['bf16_tflops']
                                                  if 'attention' in kernels else None))

    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_base, _ = run_cpu_reference(args, wl, W, steps=1, warmup=0, target_seconds=12.0)

    residues = sum(l - 2 for l in wl['lens'])
    result = {
        'metric': METRIC, 'value': T_global / ms_step * 1e3, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
        'config': {
            'workload': f'{args.model} {"predict_log_prob" if family == "esmc" else "forward (tokens -> logits)"}, packed batch <= {args.tokens} tokens per GPU, '
                        f'{"log-uniform 128-2048" if family == "esmc" else "lognormal(400, 0.75) clipped 30-3500"} '
                        f'residue lengths, seeded synthetic bf16 weights',
            'tokens_global': T_global, 'tokens_this_rank': T_local, 'sequences_global': len(wl['lens']),
            'max_len': wl['max_len'], 'residues_excl_cls_eos': residues,
            'residues_are': 'packed tokens incl. <cls>/<eos> (SURVEY.md 8d)',
            'parallelism': f'dp{world}: whole sequences per rank (LPT on a*L + b*L^2), replicated weights'
                           + (', NCCL all-gather of logits inside the step' if world > 1 else ''),
            'partition_imbalance_max_over_mean': wl['imbalance'],
            'l2_policy': 'inputs larger than L2: every step streams 1.3 GB of weights and ~1.3 GB of activations '
                         'per layer through a 126 MB L2; no explicit flush',
            'algorithmic_tflop_per_step_this_rank': fl['total'] / 1e12,
        },
        'clocks': clocks,
        'e2e': {'value': T_global / ms_e2e * 1e3, 'unit': UNIT, 'ms_per_step': ms_e2e,
                'h2d_bytes_per_step': tok_pin.numel() * 8 + cu_pin.numel() * 4,
                'd2h_bytes_per_step': out_pin.numel() * 2},
        'gpu_launches': launches,
        'roofline': roofline,
        'kernels': kernels,
        'model_tflops': fl['total'] / ms_step / 1e9,
        'cpu_baseline': cpu_base,
    }
    print(json.dumps(result))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
