"""Fusion + Viterbi inference pipeline: the call a user of the drop-in makes for a batch of sentences.

Mirrors what ``MTCCMBertForMMTokenClassificationCRF.forward(mode='test')`` does on the hot path
(CMIM:954-989, 1029-1036, 1045, 1056): cross-modal fusion of the text states with the image regions,
then CRF Viterbi decode of the emission scores.  The encoders and the BiLSTM between the two stages are
outside the hot path; their outputs (`token_embedding`, `emissions`) are inputs here.

``infer_host`` is the end-to-end entry: pinned host buffers in, host tags out, with the host->device
copies of batch i+1 overlapped with the kernels of batch i on a second stream.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import _lib, synth
from .config import FusionConfig
from .crf import CRF
from .emission import EmissionHead
from .modules import CrossModalFusion
from .precision import precision as precision_ctx

FUSION_KEYS = ('text_states', 'visual_embeds_att', 'clip_features', 'token_embedding', 'img_mask', 'text_mask')
CRF_KEYS = ('emissions', 'crf_mask')


class FusionViterbiPipeline:
    def __init__(self, shape: synth.Shape = synth.STD, device: str = 'cuda:0', precision: str = 'bf16', seed: int = 0):
        self.shape = shape
        self.device = torch.device(device)
        self.precision = precision
        torch.manual_seed(seed)
        cfg = FusionConfig(hidden_size=shape.H, num_attention_heads=shape.heads, intermediate_size=shape.inter,
                           layer_norm_eps=shape.eps)
        self.fusion = CrossModalFusion(cfg, layer_num1=shape.L, region_dim=shape.region_dim,
                                       clip_dim=shape.clip_dim).to(self.device).eval()
        self.crf = CRF(shape.T, batch_first=True).to(self.device)
        self._copy_stream: Optional[torch.cuda.Stream] = None
        self._side_stream: Optional[torch.cuda.Stream] = None
        self.overlap_decode = True     # False: run the CRF decode on the main stream (per-kernel timing passes)
        self._capturing = False

    # ---- device-resident step -------------------------------------------------------------------
    @torch.no_grad()
    def step_device(self, d: Dict[str, torch.Tensor]):
        """One pass of the hot path over a device-resident batch; returns (result, clip, tags, lens, gate)."""
        with precision_ctx(self.precision):
            return self._step_device(d)

    def _step_device(self, d: Dict[str, torch.Tensor]):
        if not self.overlap_decode:
            tags, lens = self.crf.decode_tensors(d['emissions'], d['crf_mask'])
            out = self.fusion(d['text_states'], d['visual_embeds_att'], d['clip_features'], d['token_embedding'],
                              d['img_mask'], d['text_mask'], return_dict=True, want_fused=False)
            return out['result'], out['clip'], tags, lens, out['gate']
        main = torch.cuda.current_stream(self.device)
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(self.device)
        side = self._side_stream
        # The CRF decode depends only on the emissions: run it beside the fusion kernels.
        side.wait_stream(main)
        with torch.cuda.stream(side):
            tags, lens = self.crf.decode_tensors(d['emissions'], d['crf_mask'])
            if not self._capturing:      # (a captured graph owns its allocations; nothing to hand over)
                tags.record_stream(main)
                lens.record_stream(main)
        out = self.fusion(d['text_states'], d['visual_embeds_att'], d['clip_features'], d['token_embedding'],
                          d['img_mask'], d['text_mask'], return_dict=True, want_fused=False)
        main.wait_stream(side)
        return out['result'], out['clip'], tags, lens, out['gate']

    # ---- CUDA graph of the device-resident step -----------------------------------------------------
    def capture(self, d: Dict[str, torch.Tensor], slot: int = 0, step=None):
        """``step``: the bound step method to capture (default ``step_device``).  ``slot``: library handle slot (own split-K workspace) -- captures that will be replayed CONCURRENTLY on
        different streams must use different slots.

        Capture one ``step_device(d)`` -- ~40 kernel launches on two streams -- into a CUDA graph bound to the
        (static) device buffers ``d``.  Returns ``(graph, outputs)``; ``graph.replay()`` re-runs the step on new
        contents of ``d`` and refreshes ``outputs`` in place.  The step is launch-bound on a slow host (each
        launch goes Python -> ctypes -> cudaLaunchKernelEx); a replay is one driver call."""
        with _lib.use_slot(slot):
            return self._capture(d, step or self.step_device)

    def _capture(self, d: Dict[str, torch.Tensor], step):
        cur = torch.cuda.current_stream(self.device)
        warm = torch.cuda.Stream(self.device)
        warm.wait_stream(cur)
        with torch.cuda.stream(warm):          # operand / fold caches and kernel attributes settle before capture
            for _ in range(2):
                step(d)
        cur.wait_stream(warm)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        self._capturing = True
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        n0 = _lib.launch_count(idx)
        try:
            with torch.cuda.graph(graph):
                outputs = step(d)
        finally:
            self._capturing = False
        self.graph_kernels = _lib.launch_count(idx) - n0      # kernels one replay launches
        return graph, outputs

    # ---- host batches ---------------------------------------------------------------------------
    @staticmethod
    def make_host_batch(B: int, shape: synth.Shape, seed: int, pin: bool = True,
                        bf16_states: bool = False) -> Dict[str, torch.Tensor]:
        """Synthetic host batch.  ``bf16_states``: what a caller whose encoders run in bf16 holds -- text states and
        token embedding in bf16, the regions as the producer tail's bf16 K-major rows [B, R, 2048]
        (icka_region_tail_fwd / myResnet.forward_rows; CMIM:956 applied at the source) -- half the bytes per sentence."""
        f = synth.fusion_inputs(B, shape, seed=seed)
        c = synth.crf_batch(B, shape, seed=seed)
        host = {k: f[k] for k in FUSION_KEYS}
        if bf16_states:
            g = host['visual_embeds_att']
            host['visual_embeds_att'] = g.reshape(B, g.shape[1], -1).permute(0, 2, 1).to(torch.bfloat16).contiguous()
            host['text_states'] = host['text_states'].to(torch.bfloat16)
            host['token_embedding'] = host['token_embedding'].to(torch.bfloat16)
        host['emissions'] = c['emissions']
        host['crf_mask'] = c['mask'].to(torch.uint8)
        if pin:
            host = {k: v.contiguous().pin_memory() for k, v in host.items()}
        return host

    def to_device(self, host: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        return {k: v.to(self.device, non_blocking=True) for k, v in host.items()}

    @staticmethod
    def h2d_bytes(host: Dict[str, torch.Tensor]) -> int:
        return sum(v.numel() * v.element_size() for v in host.values())

    # what one host step hands back (device tensors to copy to pinned host memory), from the step's outputs
    def _host_step(self, d: Dict[str, torch.Tensor]):
        return self.step_device(d)

    def _host_results(self, outs):
        _, _, tags, lens, gate = outs
        return tags, lens, gate

    @torch.no_grad()
    def infer_host(self, batches: List[Dict[str, torch.Tensor]], use_graphs: bool = True):
        """End-to-end over pinned host batches: H2D copy of every input, the device step, D2H of its results (fusion +
        Viterbi pipeline: tags, lengths, gate values).  Copies of batch i+1 overlap the kernels of batch i (two device
        buffer sets).  Returns per-batch tuples of pinned host tensors and the (start, end) CUDA events that bracket all
        the work."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        cs = self._copy_stream
        main = torch.cuda.current_stream(self.device)
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if not hasattr(self, '_host_slots'):
            self._host_slots = {}
        key = tuple(sorted((k, tuple(v.shape), v.dtype) for k, v in batches[0].items())) if batches else None
        slots = self._host_slots.setdefault(key, {'bufs': [None, None], 'graphs': [None, None]})
        dev_bufs = slots['bufs']
        compute_done = [None, None]
        results = []
        cs.wait_stream(main)
        start.record(cs)
        for i, host in enumerate(batches):
            slot = i & 1
            with torch.cuda.stream(cs):
                if dev_bufs[slot] is None:
                    # Allocated from the COPY stream's pool: a block the main stream's pool hands out may still be read
                    # by the previous batch's in-flight kernels (eager launches free their temporaries in host order),
                    # and the copy stream would overwrite it without waiting for them.
                    dev_bufs[slot] = {k: torch.empty_like(v, device=self.device) for k, v in host.items()}
                    for t in dev_bufs[slot].values():
                        t.record_stream(main)
                if compute_done[slot] is not None:
                    cs.wait_event(compute_done[slot])          # buffer set is free again
                for k, v in host.items():
                    dev_bufs[slot][k].copy_(v, non_blocking=True)
                copied = torch.cuda.Event()
                copied.record(cs)
            main.wait_event(copied)
            if use_graphs:
                if slots['graphs'][slot] is None:      # first use of this buffer set: capture its step once
                    main.synchronize()
                    slots['graphs'][slot] = self.capture(dev_bufs[slot], step=self._host_step)
                graph, outs = slots['graphs'][slot]
                graph.replay()
            else:
                outs = self._host_step(dev_bufs[slot])
            out = tuple(torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in self._host_results(outs))
            for h, t in zip(out, self._host_results(outs)):
                h.copy_(t, non_blocking=True)
            compute_done[slot] = torch.cuda.Event()
            compute_done[slot].record(main)
            results.append(out)
        end.record(main)
        return results, (start, end)


class TaggingPipeline(FusionViterbiPipeline):
    """The hot path widened by the SURVEY 8f "next" rows 1 and 2: fusion -> BiLSTM + classifier (CMIM:1042-1043) ->
    CRF Viterbi decode of THOSE emissions (CMIM:1056) -> tag filtering + chunk-F1 counters (My_cross_attention.py:
    1052-1077, ner_evaluate.py) -- everything between the encoders' outputs and the F1 integers stays on the device."""

    def __init__(self, shape: synth.Shape = synth.STD, device: str = 'cuda:0', precision: str = 'bf16', seed: int = 0):
        super().__init__(shape, device, precision, seed)
        from . import ner
        cfg = FusionConfig(hidden_size=shape.H, num_attention_heads=shape.heads, intermediate_size=shape.inter,
                           layer_norm_eps=shape.eps)
        self.head = EmissionHead(cfg, num_labels=shape.T).to(self.device).eval()
        self.f1 = ner.ChunkF1(device=self.device) if shape.T == 15 else None

    @torch.no_grad()
    def step_tagging(self, d: Dict[str, torch.Tensor], label_ids: Optional[torch.Tensor] = None):
        """-> (emissions [B,S,T] fp32, tags [B,S] int32, lens [B] int32); updates ``self.f1`` when labels are given."""
        with precision_ctx(self.precision):
            return self._step_tagging(d, label_ids)

    def _step_tagging(self, d: Dict[str, torch.Tensor], label_ids: Optional[torch.Tensor] = None):
        out = self.fusion(d['text_states'], d['visual_embeds_att'], d['clip_features'], d['token_embedding'],
                          d['img_mask'], d['text_mask'], return_dict=True, want_fused=False)
        emissions = self.head(out['result'])
        tags, lens = self.crf.decode_tensors(emissions, d['crf_mask'])
        if label_ids is not None and self.f1 is not None:
            self.f1.update(tags, label_ids, d['crf_mask'])
        return emissions, tags, lens

    # ---- end to end from host buffers: the tags DEPEND on the fusion result, emissions never cross PCIe ----------------
    @staticmethod
    def make_host_batch(B: int, shape: synth.Shape, seed: int, pin: bool = True,
                        bf16_states: bool = False) -> Dict[str, torch.Tensor]:
        """The fusion inputs + the CRF mask + the gold label ids (for the chunk-F1 counters); no emissions -- the emission
        head produces them on the device from the fusion result (CMIM:1042-1043)."""
        host = FusionViterbiPipeline.make_host_batch(B, shape, seed, pin=False, bf16_states=bf16_states)
        del host['emissions']
        host['label_ids'] = synth.crf_batch(B, shape, seed=seed)['tags']
        if pin:
            host = {k: v.contiguous().pin_memory() for k, v in host.items()}
        return host

    def _host_step(self, d: Dict[str, torch.Tensor]):
        """fusion -> BiLSTM + classifier -> Viterbi of those emissions -> chunk-F1 counters of THIS batch."""
        if self.f1 is not None:
            self.f1.reset()
        _, tags, lens = self.step_tagging(d, d.get('label_ids'))
        return tags, lens, (self.f1.totals if self.f1 is not None else lens)

    def _host_results(self, outs):
        return outs                      # tags [B,S] int32, lens [B] int32, chunk-F1 counters [6] int64
