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
`e2e`    : same metric through FusionViterbiPipeline.infer_host: pinned HOST inputs -> H2D -> kernels ->
           D2H of tags/lengths/gates, all inside the timed region.
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

    def __init__(self, shape, sample, seed):
        import torch
        from icka_b200 import synth
        from oracle import fusion_ref, viterbi_c
        self.torch, self.fusion_ref, self.viterbi_c = torch, fusion_ref, viterbi_c
        self.shape, self.sample = shape, sample
        self.cores = os.cpu_count() or 1
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
def run_gpu_arm(args, shape):
    import torch
    import torch.distributed as dist
    from icka_b200 import _lib, shard
    from icka_b200.pipeline import FusionViterbiPipeline
    from icka_b200.profiler import KernelTimer

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('launch with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...')
    torch.cuda.set_device(local_rank)
    dev = f'cuda:{local_rank}'
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device(dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peaks = load_peaks()
    seed = 19260817 + rank
    pipe = FusionViterbiPipeline(shape, dev, args.precision, seed=seed)
    host = pipe.make_host_batch(args.batch, shape, seed)
    d = pipe.to_device(host)
    torch.cuda.synchronize()

    # ---- device-resident timing ----
    # One step = ~40 kernel launches on two streams; by default it is captured once into a CUDA graph and the timed
    # region replays it (one driver call per step), so a slow host cannot starve the GPU.  gpu_launches counts the
    # kernels the graph contains (icka_launch_count over the capture) times the replays.
    use_graph = not args.no_graph
    inflight = args.inflight if use_graph else 1
    if use_graph:
        graph, _outs = pipe.capture(d)
        per_step = pipe.graph_kernels
        run_step = graph.replay
        if inflight == 2:
            # a second batch with its own device buffers, graph and library handle slot (split-K workspace): steps i and
            # i + 1 run on two streams, so the single-query encoders' small GEMMs at the end of a step (a few dozen CTAs)
            # share the machine with the next step's region relayout / projections instead of leaving most SMs idle
            d2 = pipe.to_device(pipe.make_host_batch(args.batch, shape, seed + 500))
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
        if inflight == 2:
            for st, _ in lanes:
                st.wait_stream(torch.cuda.current_stream())

    def join():
        if inflight == 2:
            for st, _ in lanes:
                torch.cuda.current_stream().wait_stream(st)

    fork()
    for _ in range(max(args.warmup, 3)):
        run_step()
    join()
    barrier()
    launches0 = _lib.launch_count(local_rank)
    s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        s_ev.record()
        fork()
        for _ in range(args.steps):
            run_step()
        join()
        e_ev.record()
        barrier()
    ms_total = s_ev.elapsed_time(e_ev)
    launches = per_step * args.steps if use_graph else _lib.launch_count(local_rank) - launches0
    ms_total = shard.max_over_ranks(ms_total, device=dev)
    value = args.batch * world * args.steps / (ms_total * 1e-3)

    # ---- per-kernel roofline pass (same step, CUDA events around every C-ABI launch) ----
    pipe.overlap_decode = False          # events on one stream: the decode must not be timed while it waits
    for _ in range(2):                   # eager warm-up: the caching allocator must not cudaMalloc inside the events
        pipe.step_device(d)
    torch.cuda.synchronize()
    with KernelTimer() as kt:
        for _ in range(args.steps):
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
        'launches_per_step': gemm['launches'] // args.steps, 'ms_per_launch': gemm['ms_per_launch'],
        'flops_per_launch': gemm['flops_per_launch'],
    }
    if args.precision == 'bf16' and not args.hires:
        cap = measured_traffic(f'B{args.batch}_L{shape.L}')
        if cap is not None:
            roofline['traffic'] = cap[0]['dram_bytes_per_launch_mean']
            roofline['traffic_source'] = (f'profiles/{cap[1]}: dram__bytes_read.sum + dram__bytes_write.sum, mean over the '
                                          f"{cap[0]['launches']} GEMM launches of one step (ncu --set full)")
            roofline['algorithmic_bytes_per_launch'] = gemm['bytes_per_launch']
    hbm = peaks.get('hbm_gbs', FALLBACK_PEAKS['hbm_gbs'])
    kernel_table = {}
    for name, k in kernels.items():
        kernel_table[name] = {'launches_per_step': k['launches'] // args.steps, 'ms_per_step': k['ms_total'] / args.steps,
                              'tflops': round(k['tflops'], 2), 'gbs': round(k['gbs'], 1),
                              'frac_tensor': round(k['tflops'] / peak_tf, 4), 'frac_hbm': round(k['gbs'] / hbm, 4)}

    # ---- end to end from pinned host memory ----
    e2e = None
    if not args.no_e2e:
        hosts = [host, pipe.make_host_batch(args.batch, shape, seed + 1000)]
        n_e2e = max(4, min(args.steps, 10))
        seq = [hosts[i & 1] for i in range(n_e2e)]
        pipe.infer_host(seq[:2], use_graphs=use_graph)
        barrier()
        results, (s2, e2) = pipe.infer_host(seq, use_graphs=use_graph)
        barrier()
        ms_e2e = s2.elapsed_time(e2)
        ms_e2e = shard.max_over_ranks(ms_e2e, device=dev)
        d2h = sum(x.numel() * x.element_size() for x in results[0])
        e2e = {'value': args.batch * world * n_e2e / (ms_e2e * 1e-3), 'unit': UNIT,
               'h2d_bytes_per_step': pipe.h2d_bytes(host), 'd2h_bytes_per_step': d2h, 'steps': n_e2e,
               'api': 'icka_b200.pipeline.FusionViterbiPipeline.infer_host (pinned host fp32 inputs; H2D of batch i+1 '
                      'overlaps kernels of batch i; D2H of tags, lengths, gates)'}

        # the same call with the inputs a bf16 caller holds (half the PCIe bytes; extra, the headline `e2e` stays fp32)
        if args.precision == 'bf16':
            hosts16 = [pipe.make_host_batch(args.batch, shape, seed + 2000 + i, bf16_states=True) for i in range(2)]
            seq16 = [hosts16[i & 1] for i in range(n_e2e)]
            pipe.infer_host(seq16[:2], use_graphs=use_graph)
            barrier()
            _, (s3, e3) = pipe.infer_host(seq16, use_graphs=use_graph)
            barrier()
            ms16 = shard.max_over_ranks(s3.elapsed_time(e3), device=dev)
            e2e['bf16_host_inputs'] = {
                'value': args.batch * world * n_e2e / (ms16 * 1e-3), 'unit': UNIT,
                'h2d_bytes_per_step': pipe.h2d_bytes(hosts16[0]),
                'inputs': 'text states + token embedding bf16, regions as bf16 K-major rows [B,R,2048] '
                          '(producer-tail layout), clip fp32, emissions fp32'}
            del hosts16, seq16

    # ---- widened path (SURVEY 8f rows 1 + 2; extra, not part of `value`) ----
    widened = None
    if not args.no_widened and args.precision == 'bf16' and shape.H == 768:
        try:
            widened = run_widened(args, shape, dev, seed, d, barrier)
        except RuntimeError as e:       # e.g. a profiler that cannot replay cooperative cluster launches; the headline is unaffected
            widened = None
            print(f'bench.py: widened measurement skipped: {e}', file=sys.stderr)
        if widened is not None and world > 1:
            widened['ms_per_step'] = shard.max_over_ranks(widened['ms_per_step'], device=dev)
        if widened is not None:
            widened['value'] = args.batch * world / (widened['ms_per_step'] * 1e-3)

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        base = CpuBaseline(shape, args.cpu_sample, seed=19260817)
        reps = 8 if shape.L == 1 and not args.hires else 3      # a few seconds of CPU work on the box's cores
        v, sec = base.run(reps, 1)
        cpu = {'value': v, 'unit': UNIT, 'cores': base.cores, 'kind': 'port',
               'sample': f'{args.cpu_sample} sentences x {reps} passes (fusion fwd fp32 + Viterbi), oracle port on host CPU, '
                         f'{sec * 1e3:.0f} ms per pass'}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
            'config': {'workload': workload_name(args, shape), 'batch_per_gpu': args.batch, 'global_batch': args.batch * world,
                       'layers': shape.L, 'parallelism': f'batch-sharded x{world}, no collectives',
                       'precision': 'bf16 GEMM operands, fp32 accumulate/residual/LayerNorm/softmax' if args.precision == 'bf16' else 'fp32',
                       'weights': 'random init (nn.Linear default)',
                       'launch': ('CUDA graph replay of the captured step' + (', two batches in flight on two streams' if inflight == 2 else ''))
                                 if use_graph else 'eager host launches',
                       'l2_policy': f'inputs larger than L2 ({args.batch * 1.199e6 / 1e9:.1f} GB of inputs per step vs 126 MB L2); no flush needed'},
            'clocks': clk.report(), 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roofline,
            'cpu_baseline': cpu, 'widened': widened, 'kernels': kernel_table,
            'gemm_shapes': {k: {'launches_per_step': v['launches'] // args.steps, 'ms_per_launch': round(v['ms_per_launch'], 4),
                                'tflops': round(v['tflops'], 1)} for k, v in gemm_shapes.items()},
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


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


def run_train_arm(args, shape):
    """Extra (non-headline) mode for BASELINE configs[1] / [4]: one data-parallel training step of the hot path.

    step = fusion forward (recording) -> emission head -> CRF negative log-likelihood (token_mean, CMIM:1047-1048)
           -> backward through the kernel-backed autograd nodes -> bucketed gradient all-reduce (NCCL, overlapped
           with backward) -> AdamW.  The emission head is a torch nn.Linear(H, T) standing in for the reference's
           BiLSTM + classifier (CMIM:1042-1043, SURVEY 8f "next" row) and AdamW is torch's (the reference uses
           transformers.AdamW): both are outside the hot path and are named in `config`."""
    import torch
    import torch.distributed as dist
    from icka_b200 import CRF, CrossModalFusion, FusionConfig, _lib, set_precision, shard, synth

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = f'cuda:{local_rank}'
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device(dev))
    set_precision(args.precision)
    torch.manual_seed(19260817)
    cfg = FusionConfig(hidden_size=shape.H, num_attention_heads=shape.heads, intermediate_size=shape.inter,
                       layer_norm_eps=shape.eps)
    fusion = CrossModalFusion(cfg, layer_num1=shape.L).to(dev).train()      # dropout p = 0.1 on all three sites
    if args.real_head:
        from icka_b200 import EmissionHead
        head = EmissionHead(cfg, num_labels=shape.T).to(dev).train()
    else:
        head = torch.nn.Linear(shape.H, shape.T).to(dev)
    crf = CRF(shape.T, batch_first=True).to(dev)
    params = list(fusion.parameters()) + list(head.parameters()) + list(crf.parameters())
    for m in (fusion, head, crf):
        shard.broadcast_parameters(m)
    reducer = shard.GradientAllReducer(params)
    opt = torch.optim.AdamW(params, lr=1e-5, fused=True)
    f = synth.fusion_inputs(args.batch, shape, seed=19260817 + rank)
    c = synth.crf_batch(args.batch, shape, seed=19260817 + rank)
    d = {k: f[k].to(dev) for k in ('text_states', 'visual_embeds_att', 'clip_features', 'token_embedding', 'img_mask',
                                   'text_mask')}
    tags, mask = c['tags'].to(dev), c['mask'].to(dev)

    def step():
        opt.zero_grad(set_to_none=True)
        result, clip = fusion(d['text_states'], d['visual_embeds_att'], d['clip_features'], d['token_embedding'],
                              d['img_mask'], d['text_mask'])
        loss = -crf(head(result), tags, mask, reduction='token_mean') + 1e-3 * clip.mean()
        loss.backward()
        reducer.finish()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    l0 = _lib.launch_count(local_rank)
    s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        s_ev.record()
        for _ in range(args.steps):
            loss = step()
        e_ev.record()
        barrier()
    ms = shard.max_over_ranks(s_ev.elapsed_time(e_ev), device=dev)
    launches = _lib.launch_count(local_rank) - l0
    from icka_b200.profiler import KernelTimer
    n_prof = min(args.steps, 5)
    with KernelTimer() as kt:                 # per-kernel CUDA-event pass of the same step
        for _ in range(n_prof):
            step()
        ksum = kt.summary()
    hbm = load_peaks().get('hbm_gbs', FALLBACK_PEAKS['hbm_gbs'])
    kernel_table = {name: {'launches_per_step': k['launches'] // n_prof, 'ms_per_step': round(k['ms_total'] / n_prof, 4),
                           'tflops': round(k['tflops'], 1), 'gbs': round(k['gbs'], 1)} for name, k in ksum.items()}
    n_param = sum(p.numel() for p in params)
    # dense-GEMM FLOPs per sentence: forward (unfolded single-query encoders) x3 for forward + dgrad + wgrad
    S, R, H, I, L = shape.S, shape.R, shape.H, shape.inter, shape.L
    layer = lambda sq, skv: 2.0 * (sq * H * H + 2 * skv * H * H + sq * H * H + 2 * sq * H * I)
    fwd = 2.0 * R * shape.region_dim * H + L * layer(S, R) + 2 * L * layer(1, S)
    peaks = load_peaks()
    peak_tf = peaks.get('bf16_tflops_sustained', FALLBACK_PEAKS['bf16_tflops_sustained'])
    tf = 3.0 * fwd * args.batch * args.steps / (ms * 1e-3) / 1e12
    if rank == 0:
        emit({
            'metric': 'sentences/sec fusion+CRF training step', 'value': args.batch * world * args.steps / (ms * 1e-3),
            'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16' if args.precision == 'bf16' else 'f32',
            'data': 'synthetic', 'mode': 'train',
            'config': {'workload': f'twitter2015_training_B{args.batch}_per_gpu_S{S}_R{R}_H{H}_I{I}_T{shape.T}_L{L}',
                       'global_batch': args.batch * world, 'parallelism': f'data-parallel x{world}, bucketed NCCL gradient all-reduce',
                       'params_allreduced': n_param, 'buckets': len(reducer.buckets),
                       'buckets_launched_inside_backward_per_step': reducer.launched_early // (args.steps + max(args.warmup, 3)),
                       'outside_hot_path': ('emission head = icka_b200.EmissionHead (BiLSTM + classifier, BPTT on per-step kernels)' if args.real_head
                                            else 'emission head = torch nn.Linear(H,T) stand-in for BiLSTM+classifier') + '; optimizer = torch AdamW(fused)',
                       'dropout': 'hidden_dropout_prob = attention_probs_dropout_prob = 0.1 (Philox masks regenerated in backward)'},
            'clocks': clk.report(), 'gpu_launches': int(launches), 'loss': float(loss.detach()),
            'roofline': {'kernel': 'gemm_bf16_tcgen05_kernel (fwd + dgrad + wgrad)', 'bound': 'tensor', 'achieved': tf,
                         'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': tf / peak_tf, 'traffic': None,
                         'note': 'whole-step dense-GEMM FLOPs (3 x forward) / whole-step time: includes every non-GEMM kernel, '
                                 'the all-reduce and the optimizer'},
            'kernels': kernel_table,
        })
    if world > 1:
        dist.destroy_process_group()


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
