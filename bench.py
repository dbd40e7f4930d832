#!/usr/bin/env python
"""bench.py -- sentences/s of the ICKA fusion + Viterbi hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W                  (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W (reference arm: CPU oracle port)

A "step" is one pass of the hot path over one batch of synthetic sentences per GPU: region relayout +
projection, text->image cross encoder, image->text cross encoders, gated fusion, and CRF Viterbi decode
of that batch's emission scores.  Workload = BASELINE.json configs[2] (Twitter-2017-shaped inference
sweep, batch-sharded, no communication): --batch sentences per GPU (default 1024, inside the 256-4096
sweep), S=128, R=49, H=768, 12 heads, I=3072, T=15, --layers cross layers per encoder (default 1 = the
reference constructor default, CMIM:888).

`value`  : whole-job sentences/s with the inputs already resident in HBM, CUDA-event timed, max over ranks.
`e2e`    : same metric through TaggingPipeline.infer_host: pinned HOST inputs -> H2D -> fusion -> BiLSTM + classifier ->
           Viterbi of those emissions -> chunk-F1 counters -> D2H of tags / lengths / counters, all inside the timed region.
`configs`: the other BASELINE.json configurations on the same GPUs (L=5, hi-res, B=256, training B=32 and B=128 per GPU).
`roofline`: the dominant kernel (tcgen05 bf16 GEMM): algorithmic 2*M*N*K FLOPs per launch / mean launch
           duration from CUDA events on the launch stream, against MEASURED_PEAKS.json (sustained bf16).
`cpu_baseline`: the oracle port (torch-CPU restatement of the reference modules + C Viterbi) timed on this
           box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line.  Native libraries write there too (NCCL prints its "NCCL version ..." banner to
# stdout at NCCL_DEBUG=VERSION and =WARN; device-side printf of a watchdog): file descriptor 1 is pointed at stderr for the
# whole run and the JSON line goes to the saved original descriptor (emit()).
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + '\n').encode())


METRIC = 'sentences/sec fusion+Viterbi'
UNIT = 'sentences/s'
FALLBACK_PEAKS = dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='icka', choices=['icka', 'reference'])
    ap.add_argument('--batch', type=int, default=None, help='sentences per GPU per step (default 1024; 128 for --mode train)')
    ap.add_argument('--mode', default='infer', choices=['infer', 'train'],
                    help="'infer' (default, the BASELINE metric) or 'train' (configs[1]/[4]: fwd + bwd + all-reduce + AdamW)")
    ap.add_argument('--layers', type=int, default=1, help='cross layers per encoder (layer_num1)')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--hires', action='store_true', help='S=256, R=196 variant (BASELINE configs[3])')
    ap.add_argument('--cpu-sample', type=int, default=256,
                    help='sentences per pass of the CPU baseline / reference arm (torch-CPU throughput still grows with the '
                         'batch up to a few hundred sentences: 32 under-reports the reference)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-widened', action='store_true', help='skip the extra fusion -> BiLSTM+classifier -> Viterbi -> chunk-F1 measurement')
    ap.add_argument('--inflight', type=int, default=2, choices=[1, 2],
                    help='captured steps in flight: 2 = two batches (own device buffers, own library workspace) replayed on '
                         'two streams, so the low-occupancy tail of one step overlaps the head of the next')
    ap.add_argument('--real-head', action='store_true',
                    help="--mode train: the reference's BiLSTM + classifier emission head (BPTT on per-step kernels) instead of the "
                         'nn.Linear(H,T) stand-in')
    ap.add_argument('--no-configs', action='store_true', help='skip the extra BASELINE.json configurations (`configs` block of the line)')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel from the host instead of replaying a CUDA graph')
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        d = json.load(open(p))
        d['_source'] = 'measured (MEASURED_PEAKS.json)'
        return d
    d = dict(FALLBACK_PEAKS)
    d['_source'] = 'fallback (B200_PROFILING.md)'
    return d


def measured_traffic(workload_tag):
    """DRAM bytes per launch of the dominant kernel from the newest committed `ncu --set full` capture
    (profiles/traffic_*.json, written by tools/summarize_profiles.py); None when no capture matches."""
    import glob
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, 'profiles', 'traffic_*.json'))):
        try:
            d = json.load(open(f))
        except Exception:
            continue
        if d.get('workload') == workload_tag:
            best = (d, os.path.basename(f))
    return best


def workload_name(args, shape):
    return (f'twitter2017_inference_B{args.batch}_per_gpu_S{shape.S}_R{shape.R}_H{shape.H}_nh{shape.heads}'
            f'_I{shape.inter}_T{shape.T}_L{shape.L}')


# --------------------------------------------------------------------------------------------------
# CPU baseline (oracle port) -- the only place bench.py executes oracle/
# --------------------------------------------------------------------------------------------------
class CpuBaseline:
    """The reference's modules restated on torch-CPU (oracle/fusion_ref.py; the reference itself is Python
    and /root/reference does not exist on the GPU box) + the C restatement of pytorch-crf's Viterbi."""

    def __init__(self, shape, sample, seed, threads=None):
        import torch
        from icka_b200 import synth
        from oracle import fusion_ref, viterbi_c
        self.torch, self.fusion_ref, self.viterbi_c = torch, fusion_ref, viterbi_c
        self.shape, self.sample = shape, sample
        self.cores = threads or os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.params = fusion_ref.make_params(shape.H, shape.heads, shape.inter, shape.L, seed=seed,
                                             distinct_layers=True, perturb_ln=False)
        self.inp = synth.fusion_inputs(sample, shape, seed=seed)
        self.crf = synth.crf_batch(sample, shape, seed=seed)
        self.cp = synth.crf_params(shape.T, seed)
        viterbi_c.lib()

    def step(self):
        torch, sh = self.torch, self.shape
        with torch.no_grad():
            self.fusion_ref.fusion_segment(
                self.inp['text_states'], self.inp['visual_embeds_att'], self.inp['clip_features'],
                self.inp['token_embedding'], self.inp['img_mask'], self.inp['text_mask'], self.params,
                num_layers=sh.L, num_heads=sh.heads, layer_norm_eps=sh.eps)
        self.viterbi_c.viterbi(self.crf['emissions'].numpy(), self.crf['mask'].numpy(),
                               self.cp['start_transitions'].numpy(), self.cp['end_transitions'].numpy(),
                               self.cp['transitions'].numpy())

    def run(self, steps, warmup):
        for _ in range(warmup):
            self.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step()
        dt = time.perf_counter() - t0
        return self.sample * steps / dt, dt / steps


def run_reference_arm(args, shape):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    base = CpuBaseline(shape, args.cpu_sample, seed=19260817)
    value, sec = base.run(args.steps, max(1, min(args.warmup, 3)))
    sample = (f'{args.cpu_sample} sentences per step (fusion fwd fp32 + Viterbi) of workload {workload_name(args, shape)}; '
              f'oracle port: torch-CPU restatement of CMIM:509-667,873-884,954-989,1029-1036 + C Viterbi')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args, shape), 'batch_per_gpu': args.batch, 'layers': shape.L,
                   'l2_policy': 'n/a (CPU)'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': base.cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------------------
# clocks sampler (NVML) -- runs during the timed region
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: 'gpu_idle', 0x2: 'applications_clocks_setting', 0x4: 'sw_power_cap', 0x8: 'hw_slowdown',
               0x10: 'sync_boost', 0x20: 'sw_thermal_slowdown', 0x40: 'hw_thermal_slowdown',
               0x80: 'hw_power_brake_slowdown', 0x100: 'display_clock_setting'}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                bits = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for b, name in self.REASONS.items():
                    if bits & b and name != 'gpu_idle':
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        return False

    def report(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': ['nvml unavailable']}
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2], 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
class Env:
    """Process-wide context of one bench run: rank / world, device, barrier, peaks."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get('RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit('launch with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...')
        torch.cuda.set_device(self.local_rank)
        self.dev = f'cuda:{self.local_rank}'
        if self.world > 1:
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            dist.init_process_group('nccl', device_id=torch.device(self.dev))
        self.peaks = load_peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        from icka_b200 import shard
        return shard.max_over_ranks(v, device=self.dev)

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def bench_inference(env, args, shape, batch, steps, seed, precision='bf16', inflight=2, use_graph=True, sample_clocks=False):
    """Device-resident timing of `steps` passes of the hot path over `batch` sentences per GPU + the per-kernel roofline pass.
    One step = ~40 kernel launches on two streams; by default it is captured once into a CUDA graph and the timed region
    replays it (one driver call per step), two captured batches in flight on two streams."""
    torch = env.torch
    from icka_b200 import _lib
    from icka_b200.pipeline import FusionViterbiPipeline
    from icka_b200.profiler import KernelTimer
    pipe = FusionViterbiPipeline(shape, env.dev, precision, seed=seed)
    host = pipe.make_host_batch(batch, shape, seed)
    d = pipe.to_device(host)
    torch.cuda.synchronize()
    inflight = inflight if use_graph else 1
    lanes = []
    if use_graph:
        graph, _outs = pipe.capture(d)
        per_step = pipe.graph_kernels
        run_step = graph.replay
        if inflight == 2:
            # a second batch with its own device buffers, graph and library handle slot (split-K workspace): steps i and
            # i + 1 run on two streams, so the single-query encoders' small GEMMs at the end of a step (a few dozen CTAs)
            # share the machine with the next step's region relayout / projections instead of leaving most SMs idle
            d2 = pipe.to_device(pipe.make_host_batch(batch, shape, seed + 500, pin=False))
            torch.cuda.synchronize()
            graph2, _outs2 = pipe.capture(d2, slot=1)
            lanes = [(torch.cuda.Stream(), graph), (torch.cuda.Stream(), graph2)]
            state = {'i': 0}

            def run_step():
                st, g = lanes[state['i'] & 1]
                state['i'] += 1
                with torch.cuda.stream(st):
                    g.replay()
    else:
        per_step = None
        run_step = lambda: pipe.step_device(d)

    def fork():
        for st, _ in lanes:
            st.wait_stream(torch.cuda.current_stream())

    def join():
        for st, _ in lanes:
            torch.cuda.current_stream().wait_stream(st)

    warm = max(args.warmup, 3)
    fork()
    for _ in range(warm):
        run_step()
    join()
    env.barrier()
    launches0 = _lib.launch_count(env.local_rank)
    s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clk = ClockSampler(env.local_rank) if sample_clocks else None
    if clk is not None:
        clk.__enter__()
    s_ev.record()
    fork()
    for _ in range(steps):
        run_step()
    join()
    e_ev.record()
    env.barrier()
    if clk is not None:
        clk.__exit__(None, None, None)
    ms_total = env.max_over_ranks(s_ev.elapsed_time(e_ev))
    launches = per_step * steps if use_graph else _lib.launch_count(env.local_rank) - launches0
    value = batch * env.world * steps / (ms_total * 1e-3)

    # ---- per-kernel roofline pass (same step, CUDA events around every C-ABI launch) ----
    peaks = env.peaks
    pipe.overlap_decode = False          # events on one stream: the decode must not be timed while it waits
    for _ in range(2):                   # eager warm-up: the caching allocator must not cudaMalloc inside the events
        pipe.step_device(d)
    torch.cuda.synchronize()
    n_prof = min(steps, 10)
    with KernelTimer() as kt:
        for _ in range(n_prof):
            pipe.step_device(d)
        kernels = kt.summary()
        gemm_shapes = kt.gemm_shapes()
    pipe.overlap_decode = True
    gemm = kernels.get('linear_bf16_tcgen05') or kernels.get('linear_fp32_ffma')
    peak_tf = peaks.get('bf16_tflops_sustained', FALLBACK_PEAKS['bf16_tflops_sustained'])
    roofline = {
        'kernel': 'gemm_bf16_tcgen05_kernel (icka_linear_fwd)' if 'linear_bf16_tcgen05' in kernels else 'sgemm_tn_kernel',
        'bound': 'tensor', 'achieved': gemm['tflops'], 'peak': peak_tf, 'unit': 'TFLOP/s',
        'frac': gemm['tflops'] / peak_tf, 'traffic': None,
        'peak_source': peaks['_source'] + ', bf16_tflops_sustained (kernel timed inside a long step)',
        'frac_of_burst_peak': gemm['tflops'] / peaks.get('bf16_tflops', FALLBACK_PEAKS['bf16_tflops']),
        'launches_per_step': gemm['launches'] // n_prof, 'ms_per_launch': gemm['ms_per_launch'],
        'flops_per_launch': gemm['flops_per_launch'],
        'gemm_ms_per_step': gemm['ms_total'] / n_prof,
    }
    if precision == 'bf16':
        cap = measured_traffic(f'B{batch}_L{shape.L}' + ('_hires' if shape.S != 128 else ''))
        if cap is not None:
            roofline['traffic'] = cap[0]['dram_bytes_per_launch_mean']
            roofline['traffic_source'] = (f'profiles/{cap[1]}: dram__bytes_read.sum + dram__bytes_write.sum, mean over the '
                                          f"{cap[0]['launches']} GEMM launches of one step (ncu --set full)")
            roofline['algorithmic_bytes_per_launch'] = gemm['bytes_per_launch']
    hbm = peaks.get('hbm_gbs', FALLBACK_PEAKS['hbm_gbs'])
    kernel_table = {}
    for name, k in kernels.items():
        kernel_table[name] = {'launches_per_step': k['launches'] // n_prof, 'ms_per_step': round(k['ms_total'] / n_prof, 5),
                              'tflops': round(k['tflops'], 2), 'gbs': round(k['gbs'], 1),
                              'frac_tensor': round(k['tflops'] / peak_tf, 4), 'frac_hbm': round(k['gbs'] / hbm, 4)}
    # whole-step fraction of the tensor peak: the reference formulation's FLOPs per sentence / step time
    S, R, H, I, L = shape.S, shape.R, shape.H, shape.inter, shape.L
    layer = lambda sq, skv: 2.0 * (2 * sq * H * H + 2 * skv * H * H + 2 * sq * H * I) + 4.0 * sq * skv * H
    ref_flops = 2.0 * R * shape.region_dim * H + L * layer(S, R) + 2 * L * layer(1, S)
    step_frac = ref_flops * batch * env.world * steps / (ms_total * 1e-3) / 1e12 / (peak_tf * env.world)
    return dict(value=value, ms_per_step=ms_total / steps, launches=int(launches), roofline=roofline, kernels=kernel_table,
                gemm_shapes={k: {'launches_per_step': v['launches'] // n_prof, 'ms_per_launch': round(v['ms_per_launch'], 4),
                                 'tflops': round(v['tflops'], 1)} for k, v in gemm_shapes.items()},
                clocks=clk.report() if clk is not None else None, pipe=pipe, host=host, d=d, inflight=inflight,
                step_frac_of_tensor_peak=step_frac, reference_formulation_gflop_per_sentence=ref_flops / 1e9)


def pcie_probe(env, nbytes=1 << 30, reps=4):
    """Raw pinned host -> device copy bandwidth of THIS rank while every rank copies at once: the ceiling of `e2e`."""
    torch = env.torch
    src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dst = torch.empty(nbytes, dtype=torch.uint8, device=env.dev)
    dst.copy_(src, non_blocking=True)
    env.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    b.record()
    env.barrier()
    ms = env.max_over_ranks(a.elapsed_time(b))
    return nbytes * reps / (ms * 1e-3) / 1e9


def bench_e2e(env, args, shape, batch, seed, steps, use_graph=True):
    """The same metric end to end through the public API with HOST buffers: TaggingPipeline.infer_host -- pinned host
    inputs -> H2D -> fusion -> BiLSTM + classifier -> Viterbi of THOSE emissions -> chunk-F1 counters -> D2H of tags,
    lengths and counters, all inside the timed region (copies of batch i+1 overlap the kernels of batch i)."""
    torch = env.torch
    from icka_b200.pipeline import TaggingPipeline
    pipe = TaggingPipeline(shape, env.dev, args.precision, seed=seed)
    n_e2e = max(4, min(steps, 10))
    out = None
    variants = [('fp32', False)] + ([('bf16', True)] if args.precision == 'bf16' else [])
    for tag, bf16_states in variants:
        hosts = [pipe.make_host_batch(batch, shape, seed + 1000 + i, bf16_states=bf16_states) for i in range(2)]
        seq = [hosts[i & 1] for i in range(n_e2e)]
        pipe.infer_host(seq[:2], use_graphs=use_graph)
        env.barrier()
        results, (s2, e2) = pipe.infer_host(seq, use_graphs=use_graph)
        env.barrier()
        ms = env.max_over_ranks(s2.elapsed_time(e2))
        h2d = pipe.h2d_bytes(hosts[0])
        d2h = sum(x.numel() * x.element_size() for x in results[0])
        entry = {'value': batch * env.world * n_e2e / (ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d,
                 'd2h_bytes_per_step': d2h, 'steps': n_e2e, 'h2d_gbs_per_rank': h2d * n_e2e / (ms * 1e-3) / 1e9}
        if tag == 'fp32':
            out = entry
            out['api'] = ('icka_b200.pipeline.TaggingPipeline.infer_host: pinned host fp32 inputs (text states, ResNet grid, CLIP '
                          'feature, token embedding, masks, gold labels) -> H2D -> fusion -> BiLSTM + classifier -> Viterbi of those '
                          'emissions -> chunk-F1 counters -> D2H of tags, lengths, counters; H2D of batch i+1 overlaps kernels of batch i')
            counters = [int(x) for x in results[-1][2].tolist()]
            out['chunk_f1_counters_last_batch'] = counters[:5]
        else:
            entry['inputs'] = ('text states + token embedding bf16, regions as bf16 K-major rows [B,R,2048] (producer-tail layout), '
                               'clip fp32: what a caller whose encoders run in bf16 holds')
            out['bf16_host_inputs'] = entry
        del hosts, seq, results
    out['pcie_h2d_gbs_per_rank_all_ranks_copying'] = round(pcie_probe(env), 2)
    out['bound'] = ('PCIe: h2d_gbs_per_rank vs the raw pinned-copy rate measured in the same run; aggregate H2D over N ranks is '
                    'limited by the host (one NUMA node, shared PCIe uplinks), not by the kernels')
    del pipe
    torch.cuda.empty_cache()
    return out


def run_gpu_arm(args, shape):
    env = Env(args)
    torch = env.torch
    rank, world, dev = env.rank, env.world, env.dev
    seed = 19260817 + rank
    use_graph = not args.no_graph
    head = bench_inference(env, args, shape, args.batch, args.steps, seed, precision=args.precision, inflight=args.inflight,
                           use_graph=use_graph, sample_clocks=True)
    value, roofline = head['value'], head['roofline']

    # ---- end to end from pinned host memory ----
    e2e = None
    if not args.no_e2e:
        if args.precision == 'bf16' and shape.H == 768 and shape.T == 15:
            e2e = bench_e2e(env, args, shape, args.batch, seed, args.steps, use_graph)
        else:
            e2e = bench_e2e_fusion_viterbi(env, args, head['pipe'], head['host'], shape, seed, use_graph)

    # ---- widened path (SURVEY 8f rows 1 + 2; extra, not part of `value`) ----
    widened = None
    if not args.no_widened and args.precision == 'bf16' and shape.H == 768:
        try:
            widened = run_widened(args, shape, dev, seed, head['d'], env.barrier)
        except RuntimeError as e:       # e.g. a profiler that cannot replay cooperative cluster launches; the headline is unaffected
            widened = None
            print(f'bench.py: widened measurement skipped: {e}', file=sys.stderr)
        if widened is not None and world > 1:
            widened['ms_per_step'] = env.max_over_ranks(widened['ms_per_step'])
        if widened is not None:
            widened['value'] = args.batch * world / (widened['ms_per_step'] * 1e-3)
    inflight = head['inflight']
    for k in ('pipe', 'host', 'd'):
        head.pop(k)
    torch.cuda.empty_cache()

    # ---- the other BASELINE.json configurations, each with its own roofline (same process, fewer steps) ----
    configs = None
    if not args.no_configs and args.precision == 'bf16' and not args.hires and shape.L == 1:
        configs = run_baseline_configs(env, args, seed)

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_block(args, shape)

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': head['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
            'config': {'workload': workload_name(args, shape), 'batch_per_gpu': args.batch, 'global_batch': args.batch * world,
                       'layers': shape.L, 'parallelism': f'batch-sharded x{world}, no collectives',
                       'precision': 'bf16 GEMM operands, fp32 accumulate/residual/LayerNorm/softmax' if args.precision == 'bf16' else 'fp32',
                       'weights': 'random init (nn.Linear default)',
                       'launch': ('CUDA graph replay of the captured step' + (', two batches in flight on two streams' if inflight == 2 else ''))
                                 if use_graph else 'eager host launches',
                       'l2_policy': f'inputs larger than L2 ({args.batch * 1.199e6 / 1e9:.1f} GB of inputs per step vs 126 MB L2); no flush needed'},
            'clocks': head['clocks'], 'e2e': e2e, 'gpu_launches': head['launches'], 'roofline': roofline,
            'step_frac_of_tensor_peak': round(head['step_frac_of_tensor_peak'], 4),
            'cpu_baseline': cpu, 'widened': widened, 'configs': configs, 'kernels': head['kernels'],
            'gemm_shapes': head['gemm_shapes'],
        }
        emit(line)
    env.close()


def bench_e2e_fusion_viterbi(env, args, pipe, host, shape, seed, use_graph):
    """Fallback e2e for shapes the emission head's persistent kernel is not built for: FusionViterbiPipeline.infer_host
    (emission scores are a host input, D2H of tags / lengths / gates)."""
    n_e2e = max(4, min(args.steps, 10))
    hosts = [host, pipe.make_host_batch(args.batch, shape, seed + 1000)]
    seq = [hosts[i & 1] for i in range(n_e2e)]
    pipe.infer_host(seq[:2], use_graphs=use_graph)
    env.barrier()
    results, (s2, e2) = pipe.infer_host(seq, use_graphs=use_graph)
    env.barrier()
    ms = env.max_over_ranks(s2.elapsed_time(e2))
    return {'value': args.batch * env.world * n_e2e / (ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': pipe.h2d_bytes(host),
            'd2h_bytes_per_step': sum(x.numel() * x.element_size() for x in results[0]), 'steps': n_e2e,
            'api': 'icka_b200.pipeline.FusionViterbiPipeline.infer_host (pinned host fp32 inputs incl. emission scores; D2H of tags, lengths, gates)'}


def cpu_baseline_block(args, shape):
    """The oracle port timed on this box's host cores: all threads on `--cpu-sample` sentences (the headline CPU number),
    plus BASELINE.md section 4.3's points: one thread, and all threads at B = 2 and B = 32."""
    import torch
    base = CpuBaseline(shape, args.cpu_sample, seed=19260817)
    reps = 8 if shape.L == 1 and not args.hires else 3      # a few seconds of CPU work on the box's cores
    v, sec = base.run(reps, 1)
    cpu = {'value': v, 'unit': UNIT, 'cores': base.cores, 'kind': 'port',
           'sample': f'{args.cpu_sample} sentences x {reps} passes (fusion fwd fp32 + Viterbi), oracle port on host CPU, '
                     f'{sec * 1e3:.0f} ms per pass'}
    if shape.L == 1 and not args.hires:
        points = {}
        for b in (2, 32):
            pb = CpuBaseline(shape, b, seed=19260817)
            pv, _ = pb.run(5 if b == 2 else 3, 1)
            points[f'B{b}_all_threads'] = round(pv, 1)
        one = CpuBaseline(shape, 32, seed=19260817, threads=1)
        ov, osec = one.run(2, 1)
        points['B32_one_thread'] = round(ov, 1)
        torch.set_num_threads(base.cores)
        cpu['points'] = points
        cpu['points_note'] = 'sentences/s of the same port at the batch sizes of BASELINE.md 4.3 (B = 2: configs[0]); one_thread = torch.set_num_threads(1)'
    return cpu


def run_baseline_configs(env, args, seed):
    """BASELINE.json configs[1], [3], [4] and the script-default depth of configs[2], measured in the same run so that the
    driver's BENCH / SCALE records carry them.  Fewer steps than the headline; each entry has its
    own roofline."""
    from icka_b200 import synth
    torch = env.torch
    out = {}
    steps = max(3, min(args.steps, 6))

    def slim(r, batch, shape, extra):
        e = {'value': r['value'], 'unit': UNIT, 'ms_per_step': r['ms_per_step'], 'steps': steps, 'batch_per_gpu': batch,
             'global_batch': batch * env.world, 'workload': f'S{shape.S}_R{shape.R}_H{shape.H}_L{shape.L}',
             'roofline': {k: r['roofline'][k] for k in ('kernel', 'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic',
                                                        'launches_per_step', 'gemm_ms_per_step')},
             'step_frac_of_tensor_peak': round(r['step_frac_of_tensor_peak'], 4), 'gpu_launches': r['launches']}
        e.update(extra)
        hbm_rows = {k: v for k, v in r['kernels'].items() if k in ('cross_attn_core', 'i2t_pool', 'layernorm', 'ln_gate_blend',
                                                                     'region_rows', 'viterbi', 'cast_bf16')}
        e['hbm_kernels'] = {k: {'ms_per_step': v['ms_per_step'], 'frac_hbm': v['frac_hbm']} for k, v in hbm_rows.items()}
        return e

    def infer(name, shape, batch, extra):
        try:
            r = bench_inference(env, args, shape, batch, steps, seed, precision='bf16', inflight=args.inflight,
                                use_graph=not args.no_graph)
            for k in ('pipe', 'host', 'd'):
                r.pop(k)
            out[name] = slim(r, batch, shape, extra)
        except RuntimeError as e:
            out[name] = {'error': str(e)[:300]}
        torch.cuda.empty_cache()

    infer('configs[2]_inference_L5', synth.Shape(L=5), args.batch,
          {'baseline_config': 'configs[2] at the training script\'s default depth layer_num1 = 5 (My_cross_attention.py:603)'})
    infer('configs[3]_hires_S256_R196', synth.Shape(S=256, R=196), 512,
          {'baseline_config': 'configs[3]: seq 256 x 196 regions fusion + Viterbi, batch-sharded over the N GPUs of this run (8 in the SCALE record)'})
    infer('configs[2]_inference_B256', synth.Shape(), 256,
          {'baseline_config': 'configs[2] at the low end of the 256-4096 sweep (north_star: >= 50 % of roofline per kernel at batch >= 256)'})
    for name, batch, real_head, note in (
            ('configs[1]_train_B32', 32, False, 'configs[1]: training step batch 32 bf16 on 1 B200 (per GPU when N > 1)'),
            ('configs[1]_train_B32_bilstm_head', 32, True, 'configs[1] with the reference\'s BiLSTM + classifier emission head (BPTT on per-step kernels)'),
            ('configs[4]_train_dp_B128_per_gpu', 128, False, 'configs[4]: data-parallel training, 128 sentences per GPU = global batch 1024 at N = 8, NCCL gradient all-reduce')):
        try:
            r = bench_training(env, args, synth.Shape(), batch, steps, real_head)
            r['baseline_config'] = note
            out[name] = r
        except RuntimeError as e:
            out[name] = {'error': str(e)[:300]}
        torch.cuda.empty_cache()
    return out


def run_widened(args, shape, dev, seed, d, barrier):
    """fusion -> emission head (BiLSTM + classifier) -> Viterbi of those emissions -> chunk-F1 counters, device-resident
    inputs, one CUDA-graph replay per step (the recurrent kernel is one cooperative launch per <= 2048 sentences);
    then each stage on its own, CUDA-event timed."""
    import torch
    from icka_b200 import _lib
    from icka_b200.pipeline import TaggingPipeline
    from icka_b200 import synth
    steps = max(3, min(args.steps, 10))
    pipe = TaggingPipeline(shape, dev, args.precision, seed=seed)
    labels = synth.crf_batch(args.batch, shape, seed=seed)['tags'].to(dev)
    with torch.no_grad():
        for _ in range(3):
            pipe.step_tagging(d, labels)
    barrier()
    idx = torch.device(dev).index or 0
    run, launch_mode = (lambda: pipe.step_tagging(d, labels)), 'eager launches'
    launches = None
    if not args.no_graph:                      # one graph replay per step, like the headline measurement
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count(idx)
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                pipe.step_tagging(d, labels)
        launches = _lib.launch_count(idx) - n0
        torch.cuda.current_stream().wait_stream(side)
        run, launch_mode = graph.replay, 'CUDA graph replay'
        pipe.f1.reset()
    for _ in range(2):
        run()
    barrier()
    n0 = _lib.launch_count(idx)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(steps):
        run()
    ev[1].record()
    barrier()
    ms = ev[0].elapsed_time(ev[1]) / steps
    if launches is None:
        launches = (_lib.launch_count(idx) - n0) // steps
    # stage split (same inputs, one stage at a time)
    with torch.no_grad():
        out = pipe.fusion(d['text_states'], d['visual_embeds_att'], d['clip_features'], d['token_embedding'],
                          d['img_mask'], d['text_mask'], return_dict=True, want_fused=False)
    result = out['result']

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / steps

    with torch.no_grad():
        ms_head = timed(lambda: pipe.head(result))
        em = pipe.head(result)
        ms_vit = timed(lambda: pipe.crf.decode_tensors(em, d['crf_mask']))
        tags, _ = pipe.crf.decode_tensors(em, d['crf_mask'])
        ms_f1 = timed(lambda: pipe.f1.update(tags, labels, d['crf_mask']))
        # the two remaining 8f rows, on their own (they feed the encoders, which are outside the path)
        from icka_b200 import FusionConfig, PromptMapping, ops
        prompt = PromptMapping(FusionConfig(hidden_size=shape.H)).to(dev).eval()
        vmean = d['visual_embeds_att'].mean(dim=(2, 3))
        ms_prompt = timed(lambda: prompt(out['clip'], vmean, d['text_mask']))
        ms_tail = timed(lambda: ops.region_tail(d['visual_embeds_att'], d['visual_embeds_att'].shape[-1],
                                                want_att=False, rows_dtype=torch.bfloat16))
    B, S, H = args.batch, shape.S, shape.H
    rec_flops = 2.0 * B * S * 8 * H * H
    return {'path': 'fusion -> BiLSTM + classifier (icka_lstm_rec_fwd, persistent tcgen05) -> Viterbi of those emissions -> '
                    'chunk-F1 counters (icka_ner_chunk_counts)', 'unit': UNIT, 'ms_per_step': ms, 'steps': steps,
            'launches_per_step': int(launches),
            'stage_ms': {'emission_head': round(ms_head, 4), 'viterbi': round(ms_vit, 4), 'chunk_f1': round(ms_f1, 4)},
            'other_rows_ms': {'prompt_mapping': round(ms_prompt, 4), 'region_tail_fc_rows': round(ms_tail, 4)},
            'launch': launch_mode,
            'emission_head_tflops': round(2 * rec_flops / (ms_head * 1e-3) / 1e12, 1),
            'note': 'extra measurement (SURVEY 8f rows), not included in `value`'}


def bench_training(env, args, shape, batch, steps, real_head, precision='bf16'):
    """One data-parallel training step of the hot path (BASELINE configs[1] / [4]), `steps` times:

    fusion forward (recording) -> emission head -> CRF negative log-likelihood (token_mean, CMIM:1047-1048) -> backward
    through the kernel-backed autograd nodes -> bucketed gradient all-reduce (NCCL, overlapped with backward) -> AdamW.
    The emission head is a torch nn.Linear(H, T) standing in for the reference's BiLSTM + classifier unless `real_head`
    (CMIM:1042-1043, SURVEY 8f "next" row) and AdamW is torch's (the reference uses transformers.AdamW): both are outside
    the hot path and are named in the result."""
    torch, dist = env.torch, env.dist
    from icka_b200 import CRF, CrossModalFusion, FusionConfig, _lib, precision as precision_ctx, shard, synth
    from icka_b200.profiler import KernelTimer
    rank, world, dev, local_rank = env.rank, env.world, env.dev, env.local_rank
    torch.manual_seed(19260817)
    cfg = FusionConfig(hidden_size=shape.H, num_attention_heads=shape.heads, intermediate_size=shape.inter,
                       layer_norm_eps=shape.eps)
    fusion = CrossModalFusion(cfg, layer_num1=shape.L, precision=precision).to(dev).train()      # dropout p = 0.1 on all three sites
    if real_head:
        from icka_b200 import EmissionHead
        head = EmissionHead(cfg, num_labels=shape.T).to(dev).train()
    else:
        head = torch.nn.Linear(shape.H, shape.T).to(dev)
    crf = CRF(shape.T, batch_first=True).to(dev)
    params = list(fusion.parameters()) + list(head.parameters()) + list(crf.parameters())
    for m in (fusion, head, crf):
        shard.broadcast_parameters(m)
    reducer = shard.GradientAllReducer(params)
    use_graph = not args.no_graph
    opt = torch.optim.AdamW(params, lr=1e-5, fused=True, capturable=use_graph)
    f = synth.fusion_inputs(batch, shape, seed=19260817 + rank)
    c = synth.crf_batch(batch, shape, seed=19260817 + rank)
    d = {k: f[k].to(dev) for k in ('text_states', 'visual_embeds_att', 'clip_features', 'token_embedding', 'img_mask',
                                   'text_mask')}
    tags, mask = c['tags'].to(dev), c['mask'].to(dev)

    def step():
        with precision_ctx(precision):
            opt.zero_grad(set_to_none=True)
            result, clip = fusion(d['text_states'], d['visual_embeds_att'], d['clip_features'], d['token_embedding'],
                                  d['img_mask'], d['text_mask'])
            loss = -crf(head(result), tags, mask, reduction='token_mean') + 1e-3 * clip.mean()
            loss.backward()
            reducer.finish()
            opt.step()
        return loss

    warm = max(args.warmup, 3)
    if not use_graph:
        for _ in range(warm):
            step()
    # the whole step (forward, backward, all-reduce, AdamW) as ONE CUDA graph: ~130 launches from Python make the eager
    # step host-bound at these batch sizes (icka_b200/graphs.py)
    captured = None
    run = step
    if use_graph:
        from icka_b200.graphs import CapturedStep
        captured = CapturedStep(step, dev, warmup=warm)          # (eager warm-up steps run on the capture stream)
        run = captured.replay
        for _ in range(2):
            run()
    env.barrier()
    l0 = _lib.launch_count(local_rank)
    s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        s_ev.record()
        for _ in range(steps):
            loss = run()
        e_ev.record()
        env.barrier()
    ms = env.max_over_ranks(s_ev.elapsed_time(e_ev))
    launches = captured.kernels * steps if captured is not None else _lib.launch_count(local_rank) - l0
    if captured is not None:
        captured.close()
    n_prof = min(steps, 5)
    early0 = reducer.launched_early
    # per-kernel CUDA-event pass of the same step (eager launches).  At 32-128 sentences the host issues kernels slower than
    # the GPU runs them, and an event pair around a launch that finds the GPU idle would measure launch latency, not the
    # kernel: park the GPU behind a spin kernel while the host queues the whole step (the one host sync of the step, the value
    # checks of CRF._validate on inputs known to be valid, is skipped for this pass only).
    crf._validate = lambda *a, **k: None
    with KernelTimer() as kt:
        for _ in range(n_prof):
            torch.cuda._sleep(int(60e6))      # ~30 ms at 1.9 GHz
            step()
        ksum = kt.summary()
    del crf._validate
    kernel_table = {name: {'launches_per_step': k['launches'] // n_prof, 'ms_per_step': round(k['ms_total'] / n_prof, 4),
                           'tflops': round(k['tflops'], 1), 'gbs': round(k['gbs'], 1)} for name, k in ksum.items()}
    # the one collective: time every bucket's all-reduce on its own (NCCL over NVLink), after the timed region
    allreduce = None
    if world > 1:
        per_bucket = []
        for b in reducer.buckets:
            env.barrier()
            a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                dist.all_reduce(b.flat, op=dist.ReduceOp.AVG)
            z.record()
            torch.cuda.synchronize()
            t = env.max_over_ranks(a.elapsed_time(z) / 3)
            per_bucket.append({'mb': round(b.numel * 4 / 1e6, 1), 'ms': round(t, 4),
                               'bus_gbs': round(2 * (world - 1) / world * b.numel * 4 / (t * 1e-3) / 1e9, 1)})
        allreduce = {'buckets': per_bucket, 'ms_total_if_serial': round(sum(x['ms'] for x in per_bucket), 4),
                     'launched_inside_backward_per_step': (reducer.launched_early - early0) // n_prof}
    reducer.remove_hooks()
    n_param = sum(p.numel() for p in params)
    # dense-GEMM FLOPs per sentence: forward (unfolded single-query encoders) x3 for forward + dgrad + wgrad
    S, R, H, I, L = shape.S, shape.R, shape.H, shape.inter, shape.L
    layer = lambda sq, skv: 2.0 * (sq * H * H + 2 * skv * H * H + sq * H * H + 2 * sq * H * I)
    fwd = 2.0 * R * shape.region_dim * H + L * layer(S, R) + 2 * L * layer(1, S)
    peak_tf = env.peaks.get('bf16_tflops_sustained', FALLBACK_PEAKS['bf16_tflops_sustained'])
    tf = 3.0 * fwd * batch * steps / (ms * 1e-3) / 1e12
    gemm_ms = sum(v['ms_per_step'] for k, v in kernel_table.items() if 'tcgen05' in k)
    gemm_fl = sum(ksum[k]['flops_per_launch'] * ksum[k]['launches'] for k in ksum if 'tcgen05' in k) / n_prof
    return {
        'metric': 'sentences/sec fusion+CRF training step', 'value': batch * world * steps / (ms * 1e-3),
        'unit': UNIT, 'n_gpus': world, 'steps': steps, 'warmup': warm, 'ms_per_step': ms / steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16' if precision == 'bf16' else 'f32',
        'data': 'synthetic', 'mode': 'train',
        'config': {'workload': f'twitter2015_training_B{batch}_per_gpu_S{S}_R{R}_H{H}_I{I}_T{shape.T}_L{L}',
                   'batch_per_gpu': batch, 'global_batch': batch * world,
                   'parallelism': f'data-parallel x{world}, bucketed NCCL gradient all-reduce',
                   'params_allreduced': n_param, 'buckets': len(reducer.buckets),
                   'outside_hot_path': ('emission head = icka_b200.EmissionHead (BiLSTM + classifier, BPTT on per-step kernels)' if real_head
                                        else 'emission head = torch nn.Linear(H,T) stand-in for BiLSTM+classifier') + '; optimizer = torch AdamW(fused)',
                   'dropout': 'hidden_dropout_prob = attention_probs_dropout_prob = 0.1 (Philox masks regenerated in backward)',
                   'launch': ('CUDA graph replay of the whole step (forward + backward + all-reduce + AdamW); dropout seed base advanced on the device'
                              if use_graph else 'eager host launches')},
        'clocks': clk.report(), 'gpu_launches': int(launches), 'loss': float(loss.detach()),
        'roofline': {'kernel': 'gemm_bf16_tcgen05_kernel (fwd + dgrad + wgrad)', 'bound': 'tensor',
                     'achieved': gemm_fl / (ms / steps * 1e-3) / 1e12, 'peak': peak_tf, 'unit': 'TFLOP/s',
                     'frac': gemm_fl / (ms / steps * 1e-3) / 1e12 / peak_tf, 'traffic': None,
                     'note': 'EXECUTED tcgen05 GEMM FLOPs of one step (forward, dgrad, wgrad) / WHOLE-step time of the timed '
                             'replays: a lower bound for the GEMM launches (the step also holds every non-GEMM kernel, the '
                             'all-reduce and the optimizer).  The per-kernel CUDA-event times in `kernels` come from an EAGER pass; '
                             'at 32-128 sentences the host cannot keep the GPU busy there, so they include launch gaps '
                             '(device-only durations: profiles/launches_*_train128.csv)',
                     'gemm_event_ms_per_step_eager': round(gemm_ms, 4)},
        'step_frac_of_tensor_peak': round(tf / peak_tf, 4),
        'step_frac_note': 'whole-step dense-GEMM FLOPs (3 x forward) / whole-step time: includes every non-GEMM kernel, the all-reduce and the optimizer',
        'allreduce': allreduce, 'kernels': kernel_table,
    }


def run_train_arm(args, shape):
    """`--mode train`: BASELINE configs[1] / [4] as the headline of the line (extra mode; the default line carries the same
    measurement inside `configs`)."""
    env = Env(args)
    line = bench_training(env, args, shape, args.batch, args.steps, args.real_head, precision=args.precision)
    if env.rank == 0:
        emit(line)
    env.close()


def main():
    args = parse()
    if args.batch is None:
        args.batch = 128 if args.mode == 'train' else 1024
    from icka_b200 import synth
    shape = synth.Shape(L=args.layers, S=256 if args.hires else 128, R=196 if args.hires else 49)
    if args.impl == 'reference':
        run_reference_arm(args, shape)
    elif args.mode == 'train':
        run_train_arm(args, shape)
    else:
        run_gpu_arm(args, shape)


if __name__ == '__main__':
    main()
