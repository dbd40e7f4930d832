"""Drop-in mirrors of the reference's cross-modal modules, running on libicka_b200.so.

Same class names, constructor arguments, ``forward`` signatures, parameter names (state_dict keys) and
error behaviour as /root/reference/Cross_Modal_Interaction_Module.py ("CMIM"):

    BertLayerNorm (509)  BertOutput (525)  BertIntermediate (539)  BertSelfOutput (554)
    BertCoAttention (568)  BertCrossAttention (627)  BertCrossAttentionLayer (639)
    BertCrossEncoder (653)  cls_layer_both (873)

plus ``CrossModalFusion``: the hot-path slice of ``MTCCMBertForMMTokenClassificationCRF.forward``
(CMIM:954-989, 1029-1036) under the reference's own attribute names (vismap2text, vismapping,
txt2img_attention, cls_layer_Y, cls_layer, aux_head), so a reference checkpoint loads with
``load_state_dict(..., strict=False)`` exactly as My_cross_attention.py:997-998 does.

Parameters are ordinary fp32 ``nn.Parameter``s.  Arithmetic never runs in PyTorch: every forward is a
sequence of C-ABI kernel launches (icka_b200.ops).  Two precision modes (``set_precision``):

  'bf16'  bf16 GEMM operands on tcgen05 tensor cores, fp32 accumulate; the residual stream, LayerNorm
          and softmax statistics stay fp32 (SURVEY 7.3 #3b) -- the fast path (max-abs 2e-2 gate)
  'fp32'  everything fp32 on CUDA cores -- the parity path (max-rel 1e-5 gate)

Training: when autograd is recording (``torch.is_grad_enabled()``), ``BertCrossAttentionLayer``,
``BertCrossEncoder``, ``CrossModalFusion`` and ``CRF.forward`` build their graph out of the nodes in
``icka_b200.autograd`` (kernel-backed forward and backward): one fused ``CrossLayerFn`` node per cross layer; the small
building blocks below them (``BertLayerNorm``, ``BertSelfOutput``, ``BertIntermediate``, ``BertOutput``,
``BertCoAttention``, ``BertCrossAttention``, ``cls_layer_both`` called on their own) compose the generic nodes
(``DenseActFn``, ``LayerNormFn``, ``AttnCoreFn``, ``DropoutFn``), so a stack assembled from them trains with the same
gradients.  Dropout (attention probabilities CMIM:616, dense outputs CMIM:563 / 534) is applied on the recording path when
the module is in training mode: a per-call seed expands into Philox keep-masks inside the kernels and backward
regenerates them.  The precision mode is never mutated inside a forward (``icka_b200.precision``).
"""
from __future__ import annotations

import copy
from typing import List, Optional

import torch
from torch import nn

from . import ops
from ._lib import ACT_GELU_ERF, ACT_NONE, ACT_RELU, ACT_SWISH
from .autograd import (AddFn, AttnCoreFn, CastFn, CrossLayerFn, DenseActFn, DenseFn, DropoutFn, GateBlendFn,
                       LayerNormFn)
from .precision import compute_dtype, get_precision, precision, set_precision  # noqa: F401  (re-exported)

_X3_SINGLE_QUERY = True   # bf16 inference: split-precision ("bf16 x 3", ops.split3) operands on the single-query (image->text) chain.
                          # One CLIP token per sentence walks 2 x layer_num1 layers; measured at 256 sentences, L = 5
                          # (tools/i2t_precision_probe*.py): plain bf16 operands end 2.5e-2 off the fp32 reference (gate 2e-2) --
                          # z0 = vismapping(clip) alone is 6e-3 off and the first LayerNorm amplifies it by 1 / std(z0) = 1.7.
                          # vismapping always runs split (1.5e-2 -> 7e-3 after 2 layers, +7 us per 1024 sentences); the FFN of
                          # the chain's layers does when the chain is >= _X3_FFN_MIN_LAYERS deep (10 layers: 1.9e-2 -> 1.5e-2,
                          # +3 % step time), below that the error budget does not need it
_X3_FFN_MIN_LAYERS = 4
import os as _os
_FUSE_LN = _os.environ.get('ICKA_FUSE_LN', '0') == '1'     # LayerNorm inside the out-proj / FFN-down GEMM epilogue (icka_linear_ln_fwd): correct, but on
                     # B200 the second (normalising) pass re-reads rows that have left L2 and is latency-bound:
                     # 511 us fused vs 227 + 152 us unfused at B=1024, so it stays off


def _cdt() -> torch.dtype:
    return compute_dtype()


_CACHE_EPOCH = [0]      # bumped by graphs.CapturedStep.replay(): a replayed optimizer step changes parameters without
                        # touching their Python-side version counters, so operand copies made before it are stale


def invalidate_operand_caches() -> None:
    _CACHE_EPOCH[0] += 1


class _OperandCache:
    """Compute-dtype copies of fp32 parameters, refreshed when the parameter is modified in place."""

    def __init__(self):
        self._store = {}

    def get(self, key: str, params, build):
        sig = (get_precision(), _CACHE_EPOCH[0]) + tuple((p.data_ptr(), p._version, p.device) for p in params)
        hit = self._store.get(key)
        if hit is None or hit[0] != sig:
            with torch.no_grad():
                hit = (sig, build())
            self._store[key] = hit
        return hit[1]


def _operand(cache: _OperandCache, key: str, w: torch.Tensor) -> torch.Tensor:
    if get_precision() == 'fp32':
        return w.detach()
    return cache.get(key, (w,), lambda: ops.cast_bf16(w.detach().contiguous()))


def _to_lp(x32: torch.Tensor) -> torch.Tensor:
    return ops.cast_bf16(x32) if get_precision() == 'bf16' else x32


def _widen(x: torch.Tensor) -> torch.Tensor:
    """-> fp32.  A bf16 CUDA tensor outside autograd goes through the library's exact widening kernel."""
    if x.dtype == torch.float32:
        return x
    if x.dtype == torch.bfloat16 and x.is_cuda and not (torch.is_grad_enabled() and x.requires_grad):
        return ops.cast_f32(x.contiguous())
    return x.float()


def _rows(x: torch.Tensor) -> torch.Tensor:
    """[B, S, H] (any float dtype / strides) -> contiguous fp32 [B*S, H] view."""
    x = _widen(x)
    return x.contiguous().view(-1, x.shape[-1])


def _rows_and_operand(x: torch.Tensor):
    """[B, S, H] -> (fp32 residual rows, GEMM operand rows in the compute dtype).  States that arrive in bf16 on the
    bf16 path (the caller's encoders ran in bf16) ARE the operand: no rounding pass, only the exact widening."""
    if (get_precision() == 'bf16' and x.dtype == torch.bfloat16 and x.is_cuda
            and not (torch.is_grad_enabled() and x.requires_grad)):
        x_lp = x.contiguous().view(-1, x.shape[-1])
        return ops.cast_f32(x_lp), x_lp
    x32 = _rows(x)
    return x32, _to_lp(x32.detach())


def _mask2d(mask: Optional[torch.Tensor], B: int, Skv: int) -> Optional[torch.Tensor]:
    """The reference passes an additive mask broadcastable as [B,1,1,Skv] (CMIM:962-965, 976-982)."""
    if mask is None:
        return None
    if mask.numel() != B * Skv:
        raise RuntimeError(f'attention mask of shape {tuple(mask.shape)} is not [B,1,1,Skv]=[{B},1,1,{Skv}]')
    return mask.reshape(B, Skv).float().contiguous()


def _check_inference(module: nn.Module, *ps: float, graph_capable: bool = True) -> None:
    """Dropout (CMIM:616, 563, 534) lives in the autograd-recording path only (Philox masks regenerated in backward): a
    module in training mode with p > 0 called with autograd disabled refuses instead of silently skipping dropout."""
    if module.training and any(p > 0 for p in ps) and not (graph_capable and torch.is_grad_enabled()):
        raise NotImplementedError('dropout is applied on the autograd-recording path only: call .eval() for '
                                  'forward-only use (or leave autograd enabled in training mode)')


def _seed() -> int:
    """One host-side draw per call (torch's CPU generator: reproducible under torch.manual_seed); the kernels expand it
    into per-element Philox masks and backward regenerates them from the same number."""
    return int(torch.randint(0, 2 ** 31 - 1, (1,)).item())


_ACT_CODES = {'gelu': ACT_GELU_ERF, 'relu': ACT_RELU, 'swish': ACT_SWISH}


def _recording(*tensors, module: Optional[nn.Module] = None) -> bool:
    """True when this call must build an autograd graph (kernel-backed nodes of icka_b200.autograd)."""
    if not torch.is_grad_enabled():
        return False
    if any(t is not None and t.requires_grad for t in tensors):
        return True
    return module is not None and any(p.requires_grad for p in module.parameters())


class BertLayerNorm(nn.Module):
    """CMIM:509-522."""

    def __init__(self, hidden_size, eps=1e-12):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.bias = nn.Parameter(torch.zeros(hidden_size))
        self.variance_epsilon = eps

    def forward(self, x):
        shape = x.shape
        if _recording(x, module=self):
            return LayerNormFn.apply(_rows(x), self.weight, self.bias, self.variance_epsilon).view(shape)
        y, _ = ops.layernorm(_rows(x), self.weight.detach(), self.bias.detach(), self.variance_epsilon)
        return y.view(shape)


class _DenseResidualNorm(nn.Module):
    """Shared body of BertSelfOutput (CMIM:554-565) and BertOutput (CMIM:525-536)."""

    def __init__(self, in_features, config):
        super().__init__()
        self.dense = nn.Linear(in_features, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self._cache = _OperandCache()

    def _run(self, h_lp: torch.Tensor, res32: torch.Tensor, defer_ln: bool = False):
        w = _operand(self._cache, 'w', self.dense.weight)
        if defer_ln:          # the caller fuses this LayerNorm into its consumer (icka_ln_gate_blend_fwd)
            return ops.linear(h_lp, w, self.dense.bias.detach(), residual=res32, out_dtype=torch.float32), None
        if h_lp.shape[0] < 2048 or (_FUSE_LN and h_lp.shape[1] <= 1024):
            # skinny problems (the single-query encoders): split-K partials + ONE reduce-and-normalise pass.  Large ones
            # (opt-in, ICKA_FUSE_LN=1): dense + residual + LayerNorm in one launch, normalisation in the GEMM epilogue
            return ops.linear_ln(h_lp, w, self.dense.bias.detach(), res32, self.LayerNorm.weight.detach(),
                                 self.LayerNorm.bias.detach(), self.LayerNorm.variance_epsilon,
                                 want_bf16=get_precision() == 'bf16')
        pre = ops.linear(h_lp, w, self.dense.bias.detach(), residual=res32, out_dtype=torch.float32)
        return ops.layernorm(pre, self.LayerNorm.weight.detach(), self.LayerNorm.bias.detach(),
                             self.LayerNorm.variance_epsilon, want_f32=True, want_bf16=get_precision() == 'bf16')

    def _run_x3(self, h32: torch.Tensor, res32: torch.Tensor, defer_ln: bool = False):
        """Same as ``_run`` with split-precision operands (fp32 product accuracy on the tensor cores): the single-query rows."""
        w3 = self._cache.get('w3', (self.dense.weight,), lambda: ops.split3(self.dense.weight.detach().contiguous(), True))
        pre = ops.linear(ops.split3(h32), w3, self.dense.bias.detach(), residual=res32, out_dtype=torch.float32,
                         alg_k=h32.shape[1])
        if defer_ln:
            return pre, None
        return ops.layernorm(pre, self.LayerNorm.weight.detach(), self.LayerNorm.bias.detach(),
                             self.LayerNorm.variance_epsilon, want_f32=True, want_bf16=True)

    def forward(self, hidden_states, input_tensor):
        _check_inference(self, self.dropout.p)
        shape = input_tensor.shape
        if _recording(hidden_states, input_tensor, module=self):
            bf = get_precision() == 'bf16'
            ln = self.LayerNorm
            if self.training and self.dropout.p > 0:      # LN(dropout(dense(h)) + input), CMIM:562-565 / 533-536
                dense = DenseActFn.apply(_rows(hidden_states), self.dense.weight, self.dense.bias, None, ACT_NONE, bf)
                pre = AddFn.apply(DropoutFn.apply(dense, self.dropout.p, _seed()), _rows(input_tensor))
            else:
                pre = DenseActFn.apply(_rows(hidden_states), self.dense.weight, self.dense.bias, _rows(input_tensor),
                                       ACT_NONE, bf)
            return LayerNormFn.apply(pre, ln.weight, ln.bias, ln.variance_epsilon).view(shape)
        y32, _ = self._run(_to_lp(_rows(hidden_states)), _rows(input_tensor))
        return y32.view(shape)


class BertSelfOutput(_DenseResidualNorm):
    def __init__(self, config):
        super().__init__(config.hidden_size, config)


class BertOutput(_DenseResidualNorm):
    def __init__(self, config):
        super().__init__(config.intermediate_size, config)


class BertIntermediate(nn.Module):
    """CMIM:539-551.  ``config.hidden_act`` names an entry of the reference's ACT2FN table (CMIM:43): 'gelu' (erf form,
    CMIM:31-37, the default), 'relu' or 'swish' (CMIM:38-39) -- each fused into the GEMM epilogue.  The reference also
    accepts a callable there (CMIM:543-546); an arbitrary Python function cannot be fused into a kernel, so only the
    three table entries (by name) are accepted."""

    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)
        if not isinstance(config.hidden_act, str):
            raise NotImplementedError('hidden_act must name an ACT2FN entry (gelu / relu / swish): a callable cannot be '
                                      'fused into the GEMM epilogue')
        if config.hidden_act not in _ACT_CODES:
            raise KeyError(config.hidden_act)           # ACT2FN[config.hidden_act] raises the same (CMIM:544)
        self.act = _ACT_CODES[config.hidden_act]
        self._cache = _OperandCache()

    def _run(self, x_lp: torch.Tensor) -> torch.Tensor:
        w = _operand(self._cache, 'w', self.dense.weight)
        return ops.linear(x_lp, w, self.dense.bias.detach(), act=self.act, out_dtype=_cdt())

    def _run_x3(self, x32: torch.Tensor) -> torch.Tensor:
        """fp32 in, fp32 out, split-precision operands (see ``_DenseResidualNorm._run_x3``)."""
        w3 = self._cache.get('w3', (self.dense.weight,), lambda: ops.split3(self.dense.weight.detach().contiguous(), True))
        return ops.linear(ops.split3(x32), w3, self.dense.bias.detach(), act=self.act, out_dtype=torch.float32,
                          alg_k=x32.shape[1])

    def forward(self, hidden_states):
        shape = hidden_states.shape[:-1]
        if _recording(hidden_states, module=self):
            return DenseActFn.apply(_rows(hidden_states), self.dense.weight, self.dense.bias, None, self.act,
                                    get_precision() == 'bf16').view(*shape, -1)
        return self._run(_to_lp(_rows(hidden_states))).float().view(*shape, -1)


class BertCoAttention(nn.Module):
    """CMIM:568-624."""

    def __init__(self, config):
        super().__init__()
        if config.hidden_size % config.num_attention_heads != 0:
            raise ValueError(
                "The hidden size (%d) is not a multiple of the number of attention "
                "heads (%d)" % (config.hidden_size, config.num_attention_heads))
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = int(config.hidden_size / config.num_attention_heads)
        self.all_head_size = self.num_attention_heads * self.attention_head_size
        self.query = nn.Linear(config.hidden_size, self.all_head_size)
        self.key = nn.Linear(config.hidden_size, self.all_head_size)
        self.value = nn.Linear(config.hidden_size, self.all_head_size)
        self.dropout = nn.Dropout(config.attention_probs_dropout_prob)
        self._cache = _OperandCache()

    def _kv_operands(self):
        """One [K|V] projection: weights [2H, H] in the compute dtype, bias [2H] fp32."""
        ps = (self.key.weight, self.value.weight, self.key.bias, self.value.bias)

        def build():
            w = torch.cat([self.key.weight.detach(), self.value.weight.detach()], dim=0).contiguous()
            b = torch.cat([self.key.bias.detach(), self.value.bias.detach()], dim=0).contiguous()
            return (ops.cast_bf16(w) if get_precision() == 'bf16' else w), b

        return self._cache.get('kv', ps, build)

    def _run(self, x_lp, y_lp, mask2d, B, Sq, Skv):
        H = self.all_head_size
        wq = _operand(self._cache, 'q', self.query.weight)
        wkv, bkv = self._kv_operands()
        q = ops.linear(x_lp, wq, self.query.bias.detach(), out_dtype=_cdt())
        kv = ops.linear(y_lp, wkv, bkv, out_dtype=_cdt())
        return ops.cross_attn_core(q, kv[:, :H], kv[:, H:], mask2d, B, Sq, Skv, self.num_attention_heads,
                                   self.attention_head_size)

    def forward(self, s1_hidden_states, s2_hidden_states, s2_attention_mask):
        _check_inference(self, self.dropout.p)
        B, Sq, H = s1_hidden_states.shape
        Skv = s2_hidden_states.shape[1]
        mask2d = _mask2d(s2_attention_mask, B, Skv)
        if _recording(s1_hidden_states, s2_hidden_states, module=self):
            bf = get_precision() == 'bf16'
            q32 = DenseActFn.apply(_rows(s1_hidden_states), self.query.weight, self.query.bias, None, ACT_NONE, bf)
            kv32 = DenseActFn.apply(_rows(s2_hidden_states), torch.cat([self.key.weight, self.value.weight], dim=0),
                                    torch.cat([self.key.bias, self.value.bias], dim=0), None, ACT_NONE, bf)
            p = self.dropout.p if self.training else 0.0
            ctx = AttnCoreFn.apply(CastFn.apply(q32, bf), CastFn.apply(kv32, bf), mask2d,
                                   (B, Sq, Skv, self.num_attention_heads, self.attention_head_size), float(p),
                                   _seed() if p > 0 else 0)
            return ctx.float().view(B, Sq, H)
        ctx = self._run(_to_lp(_rows(s1_hidden_states)), _to_lp(_rows(s2_hidden_states)), mask2d, B, Sq, Skv)
        return ctx.float().view(B, Sq, H)


class BertCrossAttention(nn.Module):
    """CMIM:627-636."""

    def __init__(self, config):
        super().__init__()
        self.self = BertCoAttention(config)
        self.output = BertSelfOutput(config)
        self._fold_cache = _OperandCache()

    def _run(self, x32, x_lp, y_lp, mask2d, B, Sq, Skv):
        if Sq == 1 and self._can_fold():
            return self._run_single_query(x32, x_lp, y_lp, mask2d, B, Skv)
        ctx = self.self._run(x_lp, y_lp, mask2d, B, Sq, Skv)
        return self.output._run(ctx, x32)

    # -- single-query fold (image->text encoders, CMIM:984-989) ------------------------
    def _can_fold(self) -> bool:
        att = self.self
        return (get_precision() == 'bf16' and att.attention_head_size == 64 and att.num_attention_heads <= 16
                and att.all_head_size in (768, 1024))

    def _folded_operands(self):
        """Weight products of the fold, built with the fp32 kernel and cached until a parameter changes:
             Wqk[(h,i), k] = sum_j Wk[hj, i] Wq[hj, k]      u0[(h,i)] = sum_j Wk[hj, i] bq[hj]
             Wov[n, (h,i)] = sum_j Wo[n, hj] Wv[hj, i]      bo'[n]    = sum_hj Wo[n, hj] bv[hj] + bo[n]
        so that  U = z Wqk^T + u0  and  dense(ctx) = xbar Wov^T + bo'  (see csrc/i2t_pool.cu)."""
        att, out = self.self, self.output
        ps = (att.query.weight, att.query.bias, att.key.weight, att.value.weight, att.value.bias,
              out.dense.weight, out.dense.bias)

        def build():
            H, nh, d = att.all_head_size, att.num_attention_heads, att.attention_head_size
            f32 = dict(dtype=torch.float32, device=att.query.weight.device)
            wq_t = att.query.weight.detach().t().contiguous()       # [H_in, H_out]
            wk_t = att.key.weight.detach().t().contiguous()
            wv_t = att.value.weight.detach().t().contiguous()
            wo = out.dense.weight.detach().contiguous()
            bq = att.query.bias.detach().contiguous()
            wqk = torch.empty(nh * H, H, **f32)
            u0 = torch.empty(nh * H, **f32)
            wov = torch.empty(H, nh * H, **f32)
            for h in range(nh):
                sl = slice(h * d, (h + 1) * d)
                ops.linear(wk_t[:, sl], wq_t[:, sl], None, out=wqk[h * H:(h + 1) * H])
                ops.linear(bq[sl].view(1, d), wk_t[:, sl], None, out=u0[h * H:(h + 1) * H].view(1, H))
                ops.linear(wo[:, sl], wv_t[:, sl], None, out=wov[:, h * H:(h + 1) * H])
            bo2 = ops.linear(att.value.bias.detach().view(1, H).contiguous(), wo, out.dense.bias.detach()).view(H)
            return ops.cast_bf16(wqk), u0, ops.cast_bf16(wov), bo2

        return self._fold_cache.get('fold', ps, build)

    def _run_single_query(self, z32, z_lp, y_lp, mask2d, B, Skv):
        att, out = self.self, self.output
        H, nh = att.all_head_size, att.num_attention_heads
        wqk, u0, wov, bo2 = self._folded_operands()
        u = ops.linear(z_lp, wqk, u0, out_dtype=torch.bfloat16)                       # [B, nh*H]
        xbar = ops.i2t_pool(u, y_lp, mask2d, B, Skv, H, nh)                           # [B, nh*H]
        return ops.linear_ln(xbar, wov, bo2, z32, out.LayerNorm.weight.detach(), out.LayerNorm.bias.detach(),
                             out.LayerNorm.variance_epsilon, want_bf16=True)        # LN(dense(ctx) + input)

    def forward(self, s1_input_tensor, s2_input_tensor, s2_attention_mask):
        _check_inference(self, self.self.dropout.p, self.output.dropout.p)
        B, Sq, H = s1_input_tensor.shape
        Skv = s2_input_tensor.shape[1]
        if _recording(s1_input_tensor, s2_input_tensor, module=self):      # CMIM:633-636, node by node
            return self.output(self.self(s1_input_tensor, s2_input_tensor, s2_attention_mask), s1_input_tensor)
        x32 = _rows(s1_input_tensor)
        y32, _ = self._run(x32, _to_lp(x32), _to_lp(_rows(s2_input_tensor)), _mask2d(s2_attention_mask, B, Skv),
                           B, Sq, Skv)
        return y32.view(B, Sq, H)


class BertCrossAttentionLayer(nn.Module):
    """CMIM:639-650: attention block, then FFN block."""

    def __init__(self, config):
        super().__init__()
        self.attention = BertCrossAttention(config)
        self.intermediate = BertIntermediate(config)
        self.output = BertOutput(config)

    def _run(self, x32, x_lp, y_lp, mask2d, B, Sq, Skv, y32=None, defer_ln=False, x3_ffn=False):
        if _recording(x32, y32, module=self):
            return self._run_recorded(x32, y32, x_lp, y_lp, mask2d, B, Sq, Skv)
        a32, a_lp = self.attention._run(x32, x_lp, y_lp, mask2d, B, Sq, Skv)
        if x3_ffn and Sq == 1 and _X3_SINGLE_QUERY and get_precision() == 'bf16' and a32.shape[1] % 4 == 0:
            o32, o_lp = self.output._run_x3(self.intermediate._run_x3(a32), a32, defer_ln=defer_ln)
            return o32, (o_lp if o_lp is not None else o32)
        f = self.intermediate._run(a_lp if a_lp is not None else a32)
        o32, o_lp = self.output._run(f, a32, defer_ln=defer_ln)      # deferred: o32 is the PRE-LayerNorm tensor
        return o32, (o_lp if o_lp is not None else o32)

    def _run_recorded(self, x32, y32, x_lp, y_lp, mask2d, B, Sq, Skv):
        """Autograd-recording pass: one CrossLayerFn node (kernel-backed forward and backward)."""
        att, so, out = self.attention.self, self.attention.output, self.output
        if so.LayerNorm.variance_epsilon != out.LayerNorm.variance_epsilon:
            raise RuntimeError('the two LayerNorms of a cross layer must share layer_norm_eps')
        if y32 is None:
            raise RuntimeError('recording pass needs the fp32 key/value states (y32)')
        wkv, bkv = att._kv_operands()
        p_attn = att.dropout.p if self.training else 0.0
        p_hid = so.dropout.p if self.training else 0.0
        if self.training and so.dropout.p != out.dropout.p:
            raise RuntimeError('the two hidden dropouts of a cross layer must share hidden_dropout_prob')
        seed = _seed() if (p_attn > 0 or p_hid > 0) else 0
        meta = (B, Sq, Skv, att.num_attention_heads, att.attention_head_size, so.LayerNorm.variance_epsilon,
                float(p_attn), float(p_hid), seed, self.intermediate.act)
        o32, o16 = CrossLayerFn.apply(
            x32, y32, x_lp.detach(), y_lp.detach(), mask2d, meta,
            att.query.weight, att.query.bias, att.key.weight, att.key.bias, att.value.weight, att.value.bias,
            so.dense.weight, so.dense.bias, so.LayerNorm.weight, so.LayerNorm.bias,
            self.intermediate.dense.weight, self.intermediate.dense.bias,
            out.dense.weight, out.dense.bias, out.LayerNorm.weight, out.LayerNorm.bias,
            _operand(att._cache, 'q', att.query.weight), wkv, bkv, _operand(so._cache, 'w', so.dense.weight),
            _operand(self.intermediate._cache, 'w', self.intermediate.dense.weight),
            _operand(out._cache, 'w', out.dense.weight))
        return o32, (o16 if o16 is not None else o32.detach())

    def _dropouts(self):
        return (self.attention.self.dropout.p, self.attention.output.dropout.p, self.output.dropout.p)

    def forward(self, s1_hidden_states, s2_hidden_states, s2_attention_mask):
        _check_inference(self, *self._dropouts(), graph_capable=True)
        B, Sq, H = s1_hidden_states.shape
        Skv = s2_hidden_states.shape[1]
        x32, y32 = _rows(s1_hidden_states), _rows(s2_hidden_states)
        o32, _ = self._run(x32, _to_lp(x32.detach()), _to_lp(y32.detach()), _mask2d(s2_attention_mask, B, Skv),
                           B, Sq, Skv, y32=y32)
        return o32.view(B, Sq, H)


class BertCrossEncoder(nn.Module):
    """CMIM:653-667.  All layers start as deep copies of one layer, as in the reference (656-657)."""

    def __init__(self, config, layer_num):
        super().__init__()
        layer = BertCrossAttentionLayer(config)
        self.layer = nn.ModuleList([copy.deepcopy(layer) for _ in range(layer_num)])

    def _run(self, x32, x_lp, y_lp, mask2d, B, Sq, Skv, keep_all=True, y32=None, defer_last_ln=False, x3_ffn=None):
        """``defer_last_ln`` (inference only): the last layer returns its PRE-LayerNorm tensor; the caller applies
        ``self.layer[-1].output.LayerNorm`` fused into the next kernel.  ``x3_ffn``: split-precision FFN for single-query
        rows (None: decide from this encoder's own depth)."""
        if x3_ffn is None:
            x3_ffn = len(self.layer) >= _X3_FFN_MIN_LAYERS
        outs = []
        for i, layer_module in enumerate(self.layer):
            x32, x_lp = layer_module._run(x32, x_lp, y_lp, mask2d, B, Sq, Skv, y32=y32, x3_ffn=x3_ffn,
                                          defer_ln=defer_last_ln and i == len(self.layer) - 1)
            if keep_all:
                outs.append(x32)
        if not keep_all:
            outs.append(x32)
        return outs, x_lp

    def forward(self, s1_hidden_states, s2_hidden_states, s2_attention_mask, output_all_encoded_layers=True):
        for l in self.layer:
            _check_inference(l, *l._dropouts(), graph_capable=True)
        B, Sq, H = s1_hidden_states.shape
        Skv = s2_hidden_states.shape[1]
        x32, y32 = _rows(s1_hidden_states), _rows(s2_hidden_states)
        outs, _ = self._run(x32, _to_lp(x32.detach()), _to_lp(y32.detach()), _mask2d(s2_attention_mask, B, Skv),
                            B, Sq, Skv, keep_all=output_all_encoded_layers, y32=y32)
        return [o.view(B, Sq, H) for o in outs]


class cls_layer_both(nn.Module):  # noqa: N801  (reference class name, CMIM:873)
    """CMIM:873-884; ``proj_norm`` and ``LayerNorm`` are the same module under two names (876)."""

    def __init__(self, input_dim, output_dim):
        super().__init__()
        self.proj_norm = self.LayerNorm = nn.LayerNorm(input_dim)
        self.proj = nn.Linear(input_dim, output_dim)

    def forward(self, lang_feat, img_feat):
        # a [B, H] x [H, H] product: always on the fp32 kernels (ops dispatch on the operand dtype; no mode switch)
        ln = self.proj_norm
        a, b = _widen(lang_feat).contiguous(), _widen(img_feat).contiguous()
        shape = a.shape[:-1]
        a, b = a.view(-1, a.shape[-1]), b.view(-1, b.shape[-1])
        if _recording(lang_feat, img_feat, module=self):
            n = LayerNormFn.apply(AddFn.apply(a, b), ln.weight, ln.bias, ln.eps)
            return DenseActFn.apply(n, self.proj.weight, self.proj.bias, None, ACT_NONE, False).view(*shape, -1)
        n, _ = ops.layernorm(ops.add_f32(a, b), ln.weight.detach(), ln.bias.detach(), ln.eps)
        return ops.linear(n, self.proj.weight.detach(), self.proj.bias.detach()).view(*shape, -1)


class CrossModalFusion(nn.Module):
    """Hot-path slice of MTCCMBertForMMTokenClassificationCRF (CMIM:887-1057).

    Constructor mirrors the reference's use of ``config`` and ``layer_num1`` (CMIM:888-901, 933-934);
    ``num_i2t_encoders`` is 2 in the live model and 5 in the ``_bert`` clone (CMIM:1075).  ``precision``: None = follow
    ``icka_b200.get_precision()``, or 'bf16' / 'fp32' for this module's forwards (thread-safe, see icka_b200.precision).
    """

    def __init__(self, config, layer_num1=1, region_dim=2048, clip_dim=512, num_i2t_encoders=2, precision=None):
        super().__init__()
        self.hidden_size = config.hidden_size
        self.precision = precision
        self.vismap2text = nn.Linear(region_dim, config.hidden_size)                 # CMIM:897
        self.vismapping = nn.Linear(clip_dim, config.hidden_size)                    # CMIM:899
        self.txt2img_attention = BertCrossEncoder(config, layer_num1)                # CMIM:900
        self.cls_layer_Y = nn.ModuleList([BertCrossEncoder(config, layer_num1)       # CMIM:901
                                          for _ in range(num_i2t_encoders)])
        self.cls_layer = cls_layer_both(config.hidden_size, config.hidden_size)      # CMIM:933
        self.aux_head = nn.Linear(config.hidden_size, 1)                             # CMIM:934
        self._cache = _OperandCache()

    # ---- pieces shared by forward / encode / blend ----------------------------------------------------------------
    def _check_layers(self):
        for enc in (self.txt2img_attention, *self.cls_layer_Y):
            for l in enc.layer:
                _check_inference(l, *l._dropouts(), graph_capable=True)

    def _region_operand(self, visual_embeds_att, B):
        """-> (rows [B*R, C] in the compute dtype, R): the K-major operand of the region projection (CMIM:956).  Region
        rows [B, R, C] straight from the producer tail (icka_b200.myResnet.forward_rows) need no relayout."""
        if visual_embeds_att.dim() == 3 and visual_embeds_att.shape[-1] == self.vismap2text.in_features:
            R = visual_embeds_att.shape[1]
            rows = visual_embeds_att.detach().contiguous().view(B * R, -1)
            if rows.dtype != _cdt():
                rows = _to_lp(rows.float()) if get_precision() == 'bf16' else rows.float()
            return rows, R
        grid = _widen(visual_embeds_att.detach()).contiguous()
        R = grid.numel() // (B * grid.shape[1])
        return ops.region_rows(grid, _cdt()), R

    def _regions(self, rows, rec):
        """Region projection, CMIM:956-958 (the ResNet grid carries no gradient: My_cross_attention.py:804-805)
        -> (regions32 | None, regions operand)."""
        w_vm2t = _operand(self._cache, 'vm2t', self.vismap2text.weight)
        if rec:
            regions32 = DenseFn.apply(rows, self.vismap2text.weight, self.vismap2text.bias, w_vm2t)
            return regions32, _to_lp(regions32.detach())
        return None, ops.linear(rows, w_vm2t, self.vismap2text.bias.detach(), out_dtype=_cdt())

    def _image_to_text(self, clip_features, fused32, fused_lp, txt_mask, B, S, rec):
        """CMIM:954, 981-989: the single CLIP token queries the fused text states through every encoder of
        ``cls_layer_Y`` -> z32 [B, H]."""
        clip32 = clip_features.detach().float().reshape(B, -1).contiguous()
        if not rec and _X3_SINGLE_QUERY and get_precision() == 'bf16' and clip32.shape[1] % 4 == 0:
            # z0 feeds a LayerNorm that amplifies its error by 1 / std(z0) ~ 1.7: split-precision operands (fp32 accuracy)
            w3 = self._cache.get('vmap3', (self.vismapping.weight,),
                                 lambda: ops.split3(self.vismapping.weight.detach().contiguous(), True))
            z32 = ops.linear(ops.split3(clip32), w3, self.vismapping.bias.detach(), out_dtype=torch.float32,
                             alg_k=clip32.shape[1])
        else:
            clip_in = _to_lp(clip32)
            w_vmap = _operand(self._cache, 'vmap', self.vismapping.weight)
            if rec:
                z32 = DenseFn.apply(clip_in, self.vismapping.weight, self.vismapping.bias, w_vmap)
            else:
                z32 = ops.linear(clip_in, w_vmap, self.vismapping.bias.detach(), out_dtype=torch.float32)
        z_lp = _to_lp(z32.detach())
        x3_ffn = sum(len(enc.layer) for enc in self.cls_layer_Y) >= _X3_FFN_MIN_LAYERS
        for enc in self.cls_layer_Y:
            zs, z_lp = enc._run(z32, z_lp, fused_lp, txt_mask, B, 1, S, keep_all=False, y32=fused32 if rec else None,
                                x3_ffn=x3_ffn)
            z32 = zs[-1]
        return z32

    def _gate_fold(self):
        gate_params = (self.cls_layer.proj.weight, self.cls_layer.proj.bias, self.aux_head.weight, self.aux_head.bias)
        return self._cache.get('gate_fold', gate_params, lambda: ops.gate_fold(
            self.cls_layer.proj.weight.detach(), self.cls_layer.proj.bias.detach(),
            self.aux_head.weight.detach().view(-1), self.aux_head.bias.detach()))

    def forward(self, sequence_output, visual_embeds_att, clip_features, token_embedding, added_attention_mask,
                ori_input_mask, return_dict=False, want_fused=True):
        """sequence_output [B,S,H] (CMIM:953), visual_embeds_att [B,2048,g,g] (or region rows [B,R,2048]), clip_features [B,1,512],
        token_embedding [B,S,H] (CMIM:1024), added_attention_mask [B,>=R], ori_input_mask [B,S].
        Returns (result [B,S,H], clip_features [B,1,H]) -- CMIM:1036 and the loop result of CMIM:984-989."""
        with precision(self.precision):
            return self._forward(sequence_output, visual_embeds_att, clip_features, token_embedding,
                                 added_attention_mask, ori_input_mask, return_dict, want_fused)

    def _forward(self, sequence_output, visual_embeds_att, clip_features, token_embedding, added_attention_mask,
                 ori_input_mask, return_dict, want_fused):
        self._check_layers()
        B, S, H = sequence_output.shape
        rec = _recording(sequence_output, token_embedding, module=self)
        rows, R = self._region_operand(visual_embeds_att, B)
        regions32, regions_lp = self._regions(rows, rec)
        # masks, CMIM:962-965 and 976-982
        img_mask = ops.mask_additive(added_attention_mask, R)
        txt_mask = ops.mask_additive(ori_input_mask, S)

        # text -> image, CMIM:968-969
        x32, x_lp = _rows_and_operand(sequence_output)
        tok32 = _rows(token_embedding).view(B, S, H)
        ln = self.cls_layer.proj_norm
        if rec:
            outs, fused_lp = self.txt2img_attention._run(x32, x_lp, regions_lp, img_mask, B, S, R,
                                                         keep_all=False, y32=regions32)
            fused32 = outs[-1]
        else:
            # inference: the encoder's last LayerNorm is fused with the gate + blend (CMIM:1029-1036), which also
            # emits the operand copy of `fused` the image->text encoders read
            outs, _ = self.txt2img_attention._run(x32, x_lp, regions_lp, img_mask, B, S, R, keep_all=False,
                                                  defer_last_ln=True)
            ln2 = self.txt2img_attention.layer[-1].output.LayerNorm
            w_fold, c_fold = self._gate_fold()
            bf = get_precision() == 'bf16'
            result, gate, fused32, fused16 = ops.ln_gate_blend(
                outs[-1].view(B, S, H), ln2.weight.detach(), ln2.bias.detach(), ln2.variance_epsilon, tok32,
                ln.weight.detach(), ln.bias.detach(), ln.eps, w_fold, c_fold,
                want_fused_f32=(return_dict and want_fused) or not bf, want_fused_bf16=bf)
            fused_lp = fused16 if bf else fused32.view(B * S, H)

        # image -> text, CMIM:954, 981-989 (single CLIP token as the query)
        z32 = self._image_to_text(clip_features, fused32, fused_lp, txt_mask, B, S, rec)

        # gated fusion, CMIM:1029-1036 (recording pass; the inference pass did it above)
        if rec:
            result, gate = GateBlendFn.apply(fused32.view(B, S, H), tok32, ln.weight, ln.bias,
                                             self.cls_layer.proj.weight, self.cls_layer.proj.bias,
                                             self.aux_head.weight, self.aux_head.bias, ln.eps)
        if return_dict:
            out = dict(regions=regions_lp.view(B, R, H), clip=z32.view(B, 1, H), result=result, gate=gate)
            if fused32 is not None:      # the inference path only materialises fp32 `fused` when asked to
                out['fused'] = fused32.view(B, S, H)
            return out
        return result, z32.view(B, 1, H)

    # ---- the same segment in the two phases the full model needs --------------------------------------------------
    # In MTCCMBertForMMTokenClassificationCRF.forward the gate's second operand, `token_embedding`, comes out of the
    # RoBERTa `last_encoder`, which is fed prompts computed FROM the image->text result (CMIM:995-1024): the encoders
    # (CMIM:954-989) and the gate + blend (CMIM:1029-1036) cannot be one call there.  Both phases build the autograd
    # graph when autograd is recording (mode='train', CMIM:1046-1048) and run the plain kernels otherwise.
    def encode(self, sequence_output, visual_embeds_att, clip_features, added_attention_mask, ori_input_mask):
        """CMIM:954-989 -> (cross_output_layer [B,S,H] fp32, clip_features [B,1,H] fp32)."""
        with precision(self.precision):
            self._check_layers()
            B, S, H = sequence_output.shape
            rec = _recording(sequence_output, module=self)
            with torch.set_grad_enabled(rec):
                rows, R = self._region_operand(visual_embeds_att, B)
                regions32, regions_lp = self._regions(rows, rec)
                img_mask = ops.mask_additive(added_attention_mask, R)
                txt_mask = ops.mask_additive(ori_input_mask, S)
                x32, x_lp = _rows_and_operand(sequence_output)
                outs, fused_lp = self.txt2img_attention._run(x32, x_lp, regions_lp, img_mask, B, S, R, keep_all=False,
                                                             y32=regions32)
                fused32 = outs[-1]
                z32 = self._image_to_text(clip_features, fused32, fused_lp, txt_mask, B, S, rec)
            return fused32.view(B, S, H), z32.view(B, 1, H)

    def blend(self, cross_output_layer, token_embedding):
        """CMIM:1029-1036 -> result [B,S,H] fp32 = g * token_embedding + (1 - g) * cross_output_layer."""
        with precision(self.precision):
            ln = self.cls_layer.proj_norm
            if _recording(cross_output_layer, token_embedding, module=self):
                result, _ = GateBlendFn.apply(_widen(cross_output_layer).contiguous(), _widen(token_embedding).contiguous(),
                                              ln.weight, ln.bias, self.cls_layer.proj.weight, self.cls_layer.proj.bias,
                                              self.aux_head.weight, self.aux_head.bias, ln.eps)
                return result
            with torch.no_grad():
                w_fold, c_fold = self._gate_fold()
                result, _ = ops.gate_blend(_widen(cross_output_layer).contiguous(), _widen(token_embedding).contiguous(),
                                           ln.weight.detach(), ln.bias.detach(), ln.eps, w_fold, c_fold)
            return result
