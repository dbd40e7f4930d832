"""Training path: ``torch.autograd.Function`` nodes whose forward AND backward are C-ABI kernel launches.

The reference trains the fusion stack through eager-PyTorch autograd (My_cross_attention.py:814-844:
``loss = model(...)``, ``loss.backward()``, clip, AdamW step).  Here autograd is only the bookkeeping: one
node per cross-attention layer (CMIM:639-650), one for a plain dense layer (vismap2text / vismapping,
CMIM:897-899), one for the gate + blend (CMIM:1029-1036) and one for the CRF log-likelihood (CMIM:1047-1048).
Each backward is a hand-scheduled sequence of the kernels declared in include/icka_b200.h:

    dgrad / wgrad   tcgen05 GEMMs reading weights and activations in place (MN-major operands, split-K)
    LayerNorm bwd   fused with the bias-gradient column sums
    attention bwd   probabilities recomputed from the saved Q and K|V
    GELU bwd        fused into the epilogue of the FFN-down dgrad GEMM

Precision follows ``modules.set_precision``: 'bf16' keeps fp32 for the residual stream, its gradient,
LayerNorm statistics and all parameter gradients, and bf16 for GEMM operands; 'fp32' runs every GEMM on
the FFMA kernels (the gradient-parity path).  Dropout (CMIM:616, 563, 534) uses counter-based Philox masks that
backward regenerates from a per-call seed, so no mask tensor is stored.
"""
from __future__ import annotations

import os

import torch

from . import _lib, ops
from ._lib import ACT_GELU_ERF, ACT_NONE, ACT_TANH

F32 = torch.float32


def _op(t32, t16):
    return t16 if t16 is not None else t32


class DenseFn(torch.autograd.Function):
    """out32[M,N] = x_op[M,K] . W^T + b for an input that needs no gradient (region rows, CLIP feature)."""

    @staticmethod
    def forward(ctx, x_op, weight, bias, w_op):
        out = ops.linear(x_op, w_op, bias, out_dtype=F32)
        ctx.save_for_backward(x_op)
        return out

    @staticmethod
    def backward(ctx, dout):
        (x_op,) = ctx.saved_tensors
        dout = dout.contiguous()
        d_op = ops.cast_bf16(dout) if x_op.dtype == torch.bfloat16 else dout
        return None, ops.linear_wgrad(d_op, x_op), ops.colsum(dout), None


class CrossLayerFn(torch.autograd.Function):
    """One BertCrossAttentionLayer (CMIM:639-650): attention block + FFN block, forward and backward."""

    @staticmethod
    def forward(ctx, x32, y32, x_op, y_op, mask2d, meta,
                wq, bq, wk, bk, wv, bv, wo, bo, g1, b1, wi, bi, wd, bd, g2, b2,
                wq_op, wkv_op, bkv, wo_op, wi_op, wd_op):
        B, Sq, Skv, nh, d, eps, p_attn, p_hid, seed, act = meta
        H = nh * d
        dt = x_op.dtype
        bf = dt == torch.bfloat16
        q = ops.linear(x_op, wq_op, bq, out_dtype=dt)
        kv = ops.linear(y_op, wkv_op, bkv, out_dtype=dt)
        att = ops.cross_attn_core(q, kv[:, :H], kv[:, H:], mask2d, B, Sq, Skv, nh, d, p_drop=p_attn, seed=3 * seed)
        if p_hid > 0:     # dropout(dense(ctx)) + input, CMIM:562-564 (Philox mask regenerated in backward)
            pre1 = ops.dropout(ops.linear(att, wo_op, bo, out_dtype=F32), p_hid, 3 * seed + 1, residual=x32)
        else:
            pre1 = ops.linear(att, wo_op, bo, residual=x32, out_dtype=F32)
        a32, a16 = ops.layernorm(pre1, g1, b1, eps, want_f32=True, want_bf16=bf)
        a_op = _op(a32, a16)
        u = torch.empty(a_op.shape[0], wi_op.shape[0], dtype=dt, device=a_op.device)
        f = ops.linear(a_op, wi_op, bi, act=act, out_dtype=dt, pre_act_out=u)
        if p_hid > 0:     # CMIM:533-535
            pre2 = ops.dropout(ops.linear(f, wd_op, bd, out_dtype=F32), p_hid, 3 * seed + 2, residual=a32)
        else:
            pre2 = ops.linear(f, wd_op, bd, residual=a32, out_dtype=F32)
        o32, o16 = ops.layernorm(pre2, g2, b2, eps, want_f32=True, want_bf16=bf)
        ctx.meta = meta
        ctx.has_mask = mask2d is not None
        ctx.save_for_backward(x_op, y_op, q, kv, att, pre1, a_op, u, f, pre2, g1, g2, wq_op, wkv_op, wo_op, wi_op,
                              wd_op, *([mask2d] if mask2d is not None else []))
        if o16 is not None:
            ctx.mark_non_differentiable(o16)
        return o32, o16          # o16: bf16 operand copy for the next GEMM (None in fp32 mode)

    @staticmethod
    def backward(ctx, do32, _unused):
        B, Sq, Skv, nh, d, eps, p_attn, p_hid, seed, act = ctx.meta
        H = nh * d
        saved = ctx.saved_tensors
        x_op, y_op, q, kv, att, pre1, a_op, u, f, pre2, g1, g2, wq_op, wkv_op, wo_op, wi_op, wd_op = saved[:17]
        mask2d = saved[17] if ctx.has_mask else None
        dt = x_op.dtype
        bf = dt == torch.bfloat16
        do32 = do32.contiguous()

        # ---- FFN block: X' = LN2(F Wd^T + bd + A1),  F = gelu(A1 Wi^T + bi) ----
        drop = p_hid > 0
        dpre2, dpre2_16, dg2, db2, dbd = ops.layernorm_bwd(do32, pre2, g2, eps, want_f32=True, want_bf16=bf and not drop,
                                                           want_dbias=not drop)
        if drop:          # gradient of the dense output = upstream gradient through the same Philox mask
            dpre2_op = ops.dropout(dpre2, p_hid, 3 * seed + 2, out_dtype=dt)
            dbd = ops.colsum(dpre2_op)
        else:
            dpre2_op = _op(dpre2, dpre2_16)
        dwd = ops.linear_wgrad(dpre2_op, f)
        if act == ACT_GELU_ERF:   # the default: gelu' fused into the dgrad epilogue
            dgl = ops.linear_dgrad(dpre2_op, wd_op, gelu_pre=u, out_dtype=dt)      # d(pre-activation) [M, I]
        else:                     # config.hidden_act = 'relu' / 'swish' (CMIM:43): one extra element-wise pass
            dgl = ops.linear_dgrad(dpre2_op, wd_op, out_dtype=dt)
            ops.act_bwd(dgl, u, act, out=dgl)
        dwi = ops.linear_wgrad(dgl, a_op)
        dbi = ops.colsum(dgl)
        da1 = ops.linear_dgrad(dgl, wi_op, residual=dpre2, out_dtype=F32)          # + the residual branch

        # ---- attention block: A1 = LN1(ctx Wo^T + bo + X) ----
        dpre1, dpre1_16, dg1, db1, dbo = ops.layernorm_bwd(da1, pre1, g1, eps, want_f32=True, want_bf16=bf and not drop,
                                                           want_dbias=not drop)
        if drop:
            dpre1_op = ops.dropout(dpre1, p_hid, 3 * seed + 1, out_dtype=dt)
            dbo = ops.colsum(dpre1_op)
        else:
            dpre1_op = _op(dpre1, dpre1_16)
        dwo = ops.linear_wgrad(dpre1_op, att)
        datt = ops.linear_dgrad(dpre1_op, wo_op, out_dtype=dt)
        dq, dkv = ops.cross_attn_core_bwd(q, kv[:, :H], kv[:, H:], mask2d, datt, B, Sq, Skv, nh, d, ctx=att,
                                          p_drop=p_attn, seed=3 * seed)
        dwq = ops.linear_wgrad(dq, x_op)
        dbq = ops.colsum(dq)
        dwkv = ops.linear_wgrad(dkv, y_op)                                         # [2H, H]: rows = key | value
        dbkv = ops.colsum(dkv)
        dx32 = ops.linear_dgrad(dq, wq_op, residual=dpre1, out_dtype=F32) if ctx.needs_input_grad[0] else None
        dy32 = ops.linear_dgrad(dkv, wkv_op, out_dtype=F32) if ctx.needs_input_grad[1] else None

        return (dx32, dy32, None, None, None, None,
                dwq, dbq, dwkv[:H], dbkv[:H], dwkv[H:], dbkv[H:], dwo, dbo, dg1, db1, dwi, dbi, dwd, dbd, dg2, db2,
                None, None, None, None, None, None)


class LayerNormFn(torch.autograd.Function):
    """BertLayerNorm / nn.LayerNorm over the last dimension of x32 [M, N] (CMIM:509-522, :876) as a node of its own --
    the building blocks (BertLayerNorm, BertSelfOutput, BertOutput, cls_layer_both) called outside a cross layer."""

    @staticmethod
    def forward(ctx, x32, gamma, beta, eps):
        x32 = x32.contiguous()
        g = gamma.detach().float().contiguous()
        y32, _ = ops.layernorm(x32, g, beta.detach().float().contiguous(), eps)
        ctx.eps = eps
        ctx.save_for_backward(x32, g)
        return y32

    @staticmethod
    def backward(ctx, dy):
        x32, g = ctx.saved_tensors
        dx32, _, dg, db, _ = ops.layernorm_bwd(dy.contiguous(), x32, g, ctx.eps, want_f32=True, want_bf16=False,
                                               want_dbias=False)
        return dx32, dg, db, None


class DropoutFn(torch.autograd.Function):
    """y = x * keep / (1 - p) with the Philox keep-mask of ``seed`` (regenerated, not stored, in backward):
    nn.Dropout of the prompt mapping networks (CMIM:915, :918, :923, :926)."""

    @staticmethod
    def forward(ctx, x32, p, seed):
        ctx.p, ctx.seed = p, seed
        return ops.dropout(x32.contiguous(), p, seed)

    @staticmethod
    def backward(ctx, dy):
        return ops.dropout(dy.contiguous(), ctx.p, ctx.seed), None, None


class DenseActFn(torch.autograd.Function):
    """out32[M,N] = act(x32[M,K] . W[N,K]^T + b) (+ residual32) with gradients for x, W, b and the residual: a dense layer
    as a node of its own (the prompt mapping networks CMIM:914-930, and BertIntermediate / BertSelfOutput / BertOutput /
    cls_layer_both.proj / the projections of BertCoAttention when they are called outside a cross layer).
    ``bf16``: tcgen05 operands (x and W rounded to bf16, as on the inference path); otherwise the fp32 FFMA kernels.
    bf16 needs N % 64 == 0 and K % 64 == 0 (the backward GEMMs read both operands MN-major)."""

    @staticmethod
    def forward(ctx, x32, weight, bias, residual, act, bf16):
        x32 = x32.contiguous()
        w = weight.detach().float().contiguous()
        b = bias.detach().float().contiguous() if bias is not None else None
        x_op, w_op = (ops.cast_bf16(x32), ops.cast_bf16(w)) if bf16 else (x32, w)
        u = None
        if act not in (ACT_NONE, ACT_TANH):
            u = torch.empty(x_op.shape[0], w_op.shape[0], dtype=x_op.dtype, device=x_op.device)
        res = residual.contiguous() if residual is not None else None
        if act == ACT_NONE or res is None:
            out = ops.linear(x_op, w_op, b, residual=res, act=act, out_dtype=F32, pre_act_out=u)
        else:
            raise RuntimeError('DenseActFn: an activation and a residual do not occur together on this path')
        ctx.act, ctx.has_bias, ctx.has_res = act, bias is not None, residual is not None
        ref = out if act == ACT_TANH else u
        ctx.save_for_backward(x_op, w_op, *([ref] if ref is not None else []))
        return out

    @staticmethod
    def backward(ctx, dout):
        saved = ctx.saved_tensors
        x_op, w_op = saved[:2]
        dout = dout.contiguous()
        bf = x_op.dtype == torch.bfloat16
        if ctx.act != ACT_NONE:
            ref = saved[2]
            if ctx.act == ACT_TANH:            # ref = the fp32 output
                dpre32 = ops.act_bwd(dout, ref, ctx.act)
                d_op = ops.cast_bf16(dpre32) if bf else dpre32
            else:                              # ref = the pre-activation in the operand dtype
                d_op = ops.act_bwd(ops.cast_bf16(dout) if bf else dout, ref, ctx.act)
        else:
            d_op = ops.cast_bf16(dout) if bf else dout
        dx = ops.linear_dgrad(d_op, w_op, out_dtype=F32) if ctx.needs_input_grad[0] else None
        dw = ops.linear_wgrad(d_op, x_op) if ctx.needs_input_grad[1] else None
        db = _colsum_any(d_op) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        dres = dout if (ctx.has_res and ctx.needs_input_grad[3]) else None
        return dx, dw, db, dres, None, None


def _colsum_any(x):
    """Column sums for any column count (the kernel reads column pairs: an odd count gets a zero column)."""
    N = x.shape[1]
    if N % 2 == 0:
        return ops.colsum(x)
    padded = torch.zeros(x.shape[0], N + 1, dtype=x.dtype, device=x.device)
    padded[:, :N] = x
    return ops.colsum(padded)[:N]


class AttnCoreFn(torch.autograd.Function):
    """ctx = softmax(Q K^T / sqrt(d) + mask) V per head (CMIM:605-623) as a node of its own: BertCoAttention called
    outside a cross layer.  q [B*Sq, H], kv [B*Skv, 2H] in the operand dtype; the core's output in the same dtype."""

    @staticmethod
    def forward(ctx, q, kv, mask2d, dims, p_drop, seed):
        B, Sq, Skv, nh, d = dims
        H = nh * d
        out = ops.cross_attn_core(q, kv[:, :H], kv[:, H:], mask2d, B, Sq, Skv, nh, d, p_drop=p_drop, seed=seed)
        ctx.dims, ctx.p_drop, ctx.seed, ctx.has_mask = dims, p_drop, seed, mask2d is not None
        ctx.save_for_backward(q, kv, out, *([mask2d] if mask2d is not None else []))
        return out

    @staticmethod
    def backward(ctx, dout):
        saved = ctx.saved_tensors
        q, kv, out = saved[:3]
        mask2d = saved[3] if ctx.has_mask else None
        B, Sq, Skv, nh, d = ctx.dims
        H = nh * d
        dq, dkv = ops.cross_attn_core_bwd(q, kv[:, :H], kv[:, H:], mask2d, dout.contiguous().to(q.dtype), B, Sq, Skv, nh, d,
                                          ctx=out, p_drop=ctx.p_drop, seed=ctx.seed)
        return dq, dkv, None, None, None, None


class AddFn(torch.autograd.Function):
    """a + b (fp32, same shape) on the library's kernel."""

    @staticmethod
    def forward(ctx, a, b):
        return ops.add_f32(a.contiguous(), b.contiguous())

    @staticmethod
    def backward(ctx, d):
        return d, d


class CastFn(torch.autograd.Function):
    """fp32 -> operand dtype (bf16 rounding / identity) whose backward widens the gradient again."""

    @staticmethod
    def forward(ctx, x32, bf16):
        return ops.cast_bf16(x32.contiguous()) if bf16 else x32

    @staticmethod
    def backward(ctx, d):
        return (ops.cast_f32(d.contiguous()) if d.dtype == torch.bfloat16 else d), None


class GateBlendFn(torch.autograd.Function):
    """result = g tok + (1 - g) fused with the sentence gate of CMIM:1029-1036."""

    @staticmethod
    def forward(ctx, fused, tok, ln_w, ln_b, wp, bp, wa, ba, ln_eps):
        wa_v = wa.reshape(-1).contiguous()
        w_fold, c_fold = ops.gate_fold(wp, bp, wa_v, ba)
        out, gate = ops.gate_blend(fused, tok, ln_w, ln_b, ln_eps, w_fold, c_fold)
        ctx.ln_eps = ln_eps
        ctx.save_for_backward(fused, tok, gate, ln_w, ln_b, w_fold, wp, bp, wa_v)
        ctx.mark_non_differentiable(gate)
        return out, gate

    @staticmethod
    def backward(ctx, dout, _dgate):
        fused, tok, gate, ln_w, ln_b, w_fold, wp, bp, wa_v = ctx.saved_tensors
        dfused, dtok, d_ln_w, d_ln_b, d_wf, d_cf = ops.gate_blend_bwd(
            dout.contiguous(), fused, tok, gate, ln_w, ln_b, ctx.ln_eps, w_fold, want_dtok=ctx.needs_input_grad[1])
        dwp, dbp, dwa, dba = ops.gate_fold_bwd(wp, bp, wa_v, d_wf, d_cf)
        return dfused, dtok, d_ln_w, d_ln_b, dwp, dbp, dwa.view(1, -1), dba, None


class CrfLlhFn(torch.autograd.Function):
    """Per-sentence CRF log-likelihood (pytorch-crf forward, reduction='none')."""

    @staticmethod
    def forward(ctx, emissions, start, end, trans, tags, mask_u8):
        llh = ops.crf_llh(emissions, tags, mask_u8, start, end, trans)
        ctx.has_mask = mask_u8 is not None
        ctx.save_for_backward(emissions, start, end, trans, tags, *([mask_u8] if mask_u8 is not None else []))
        return llh

    @staticmethod
    def backward(ctx, dllh):
        saved = ctx.saved_tensors
        emissions, start, end, trans, tags = saved[:5]
        mask_u8 = saved[5] if ctx.has_mask else None
        de, ds, dend, dtr = ops.crf_llh_bwd(emissions, tags, mask_u8, start, end, trans, dllh.float().contiguous())
        return de, ds, dend, dtr, None, None


class LinearFn(torch.autograd.Function):
    """out32[M,N] = x[M,K] . W^T + b with gradients for x, W and b (fp32 FFMA kernels): the 2H -> num_labels classifier
    (CMIM:910, 1043), too narrow for a tensor-core tile."""

    @staticmethod
    def forward(ctx, x32, weight, bias):
        x32 = x32.contiguous()
        w = weight.detach().float().contiguous()
        ctx.save_for_backward(x32, w)
        return ops.linear(x32, w, bias.detach().float().contiguous(), out_dtype=F32)

    @staticmethod
    def backward(ctx, dout):
        x32, w = ctx.saved_tensors
        dout = dout.contiguous()
        N = dout.shape[1]
        if N % 2:                              # the column-sum kernel reads pairs: pad the odd label count with a zero column
            padded = torch.zeros(dout.shape[0], N + 1, dtype=dout.dtype, device=dout.device)
            padded[:, :N] = dout
            db = ops.colsum(padded)[:N]
        else:
            db = ops.colsum(dout)
        return ops.linear_dgrad(dout, w, out_dtype=F32), ops.linear_wgrad(dout, x32), db


class ClassifierFn(torch.autograd.Function):
    """emissions[M,T] = y[M,2H] . W^T + b for the 2H -> num_labels classifier (CMIM:910, 1043), T <= 16: the skinny
    streaming kernels (``icka_emission_head_fwd`` / ``_bwd``: one pass over the states each) instead of GEMM tiles."""

    @staticmethod
    def forward(ctx, y32, weight, bias, time_major_S=0):
        """``time_major_S`` = S: the rows of y32 are time-major (t*B + b, as the fused BiLSTM training kernels write them);
        the emissions come out batch-major (b*S + t) either way."""
        y32 = y32.contiguous()
        w = weight.detach().float().contiguous()
        ctx.save_for_backward(y32, w)
        ctx.tm_S = int(time_major_S)
        return ops.emission_head(y32, w, bias.detach().float().contiguous(), time_major_S=ctx.tm_S)

    @staticmethod
    def backward(ctx, dout):
        y32, w = ctx.saved_tensors
        dout = dout.contiguous()
        dy, dw = ops.emission_head_bwd(dout, y32, w, want_dx=ctx.needs_input_grad[0], want_dw=ctx.needs_input_grad[1],
                                       time_major_S=ctx.tm_S)
        return dy, dw, (_colsum_any(dout) if ctx.needs_input_grad[2] else None), None


_side_streams = {}


def _two_streams(dev, run_direction):
    """The two directions of the BiLSTM are independent chains of small kernels: direction 1 runs on a side stream with
    its own library handle slot (workspace), direction 0 on the current stream."""
    main = torch.cuda.current_stream(dev)
    side = _side_streams.get(dev)
    if side is None:
        side = _side_streams[dev] = torch.cuda.Stream(dev)
    side.wait_stream(main)
    with torch.cuda.stream(side), _lib.use_slot(1):
        run_direction(1)
    run_direction(0)
    main.wait_stream(side)


_FUSED_STEP_H = 768                                   # hidden size csrc/lstm_train.cu is built for
_FUSED_STEPS = os.environ.get('ICKA_LSTM_FUSED_STEPS', '1') != '0'
_FUSED_STEP_MAX_B = int(os.environ.get('ICKA_LSTM_FUSED_STEPS_MAX_B', '64'))


class BiLstmFn(torch.autograd.Function):
    """Bidirectional single-layer LSTM (CMIM:905-908, 1042) with backpropagation through time.  bf16, H = 768, up to
    ``_FUSED_STEP_MAX_B`` sentences: one fused launch per time step for both directions (csrc/lstm_train.cu); otherwise the
    per-step kernels described below.

    forward   Gx = x . [W_ih; W_ih_r]^T + (b_ih + b_hh) in one GEMM; per direction and step: gates = h_{t-1} . W_hh^T
              (GEMM) and ``icka_lstm_cell_fwd_save`` (keeps the gate activations and the cell state, fp32)
    backward  per direction, steps in reverse: ``icka_lstm_cell_bwd`` (gate pre-activation gradients straight into a
              [B, S, 8H] buffer in position order) and dh_{t-1} = dpre_t . W_hh (dgrad GEMM); afterwards the weight,
              bias and input gradients as FOUR big GEMMs over all B*S rows (dW_hh per direction against the shifted
              output sequence, dW_ih, dx) and one column sum.
    The persistent tcgen05 kernel serves inference.
    """

    @staticmethod
    def forward(ctx, x, w_ih, w_hh, b_ih, b_hh, w_ih_r, w_hh_r, b_ih_r, b_hh_r, bf16):
        B, S, I = x.shape
        H = w_hh.shape[1]
        cdt = torch.bfloat16 if bf16 else F32
        dev = x.device
        op = (lambda t: ops.cast_bf16(t.detach().float().contiguous())) if bf16 else (lambda t: t.detach().float().contiguous())
        # above 64 sentences the fused step kernel walks its 32-row blocks one after the other and loses to the GEMM launches
        ctx.fused_steps = bool(bf16 and H == _FUSED_STEP_H and _FUSED_STEPS and B <= _FUSED_STEP_MAX_B)
        if ctx.fused_steps:
            return BiLstmFn._forward_fused(ctx, x, op(torch.cat([w_ih, w_ih_r])), (op(w_hh), op(w_hh_r)),
                                           (b_ih, b_ih_r, b_hh, b_hh_r), B, S, I, H)
        x_op = op(x.reshape(B * S, I))
        wi_op = op(torch.cat([w_ih, w_ih_r]))
        wh_ops = (op(w_hh), op(w_hh_r))
        bsum = ops.add_f32(torch.cat([b_ih, b_ih_r]).detach().float().contiguous(),
                           torch.cat([b_hh, b_hh_r]).detach().float().contiguous())
        gx = ops.linear(x_op, wi_op, bsum, out_dtype=cdt).view(B, S, 8 * H)
        y32 = torch.empty(B, S, 2 * H, dtype=F32, device=dev)
        y_op = torch.empty(B, S, 2 * H, dtype=cdt, device=dev)
        acts = torch.empty(2, S, B, 4 * H, dtype=F32, device=dev)
        c_all = torch.empty(2, S, B, H, dtype=F32, device=dev)
        lib = _lib.load()
        esz_dt = _lib.BF16 if bf16 else _lib.F32

        def run_direction(d):
            h_dev = _lib.handle(dev.index if dev.index is not None else torch.cuda.current_device())
            gates = torch.empty(B, 4 * H, dtype=F32, device=dev)
            h_scr = torch.empty(2, B, H, dtype=cdt, device=dev)
            st = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.icka_lstm_dir_fwd_save(
                h_dev, gx[:, :, d * 4 * H:].data_ptr(), S * 8 * H, 8 * H, wh_ops[d].data_ptr(), acts[d].data_ptr(),
                c_all[d].data_ptr(), y_op[:, :, d * H:].data_ptr(), y32[:, :, d * H:].data_ptr(), S * 2 * H, 2 * H,
                gates.data_ptr(), h_scr.data_ptr(), esz_dt, B, S, H, d, st), 'icka_lstm_dir_fwd_save')
            gates.record_stream(torch.cuda.current_stream(dev))
            h_scr.record_stream(torch.cuda.current_stream(dev))

        _two_streams(dev, run_direction)
        ctx.save_for_backward(x_op, wi_op, wh_ops[0], wh_ops[1], y_op, acts, c_all)
        ctx.dims = (B, S, I, H, bf16)
        return y32

    @staticmethod
    def _forward_fused(ctx, x, wi_op, wh_ops, biases, B, S, I, H):
        """One launch per time step for both directions (csrc/lstm_train.cu).  Every sequence tensor is TIME-MAJOR
        ([S, B, .]): a step touches one contiguous block per tensor instead of B rows S x 12 KB apart.  The result is
        returned as the batch-first VIEW of the time-major tensor."""
        dev = x.device
        lib = _lib.load()
        b_ih, b_ih_r, b_hh, b_hh_r = biases
        x_tm = ops.cast_bf16_time_major(x.detach().float().contiguous())             # [S, B, I] bf16
        bsum = ops.add_f32(torch.cat([b_ih, b_ih_r]).detach().float().contiguous(),
                           torch.cat([b_hh, b_hh_r]).detach().float().contiguous())
        gx = ops.linear(x_tm.view(S * B, I), wi_op, bsum, out_dtype=torch.bfloat16)  # [S*B, 8H]
        y32 = torch.empty(S, B, 2 * H, dtype=F32, device=dev)
        y_op = torch.empty(S, B, 2 * H, dtype=torch.bfloat16, device=dev)
        acts = torch.empty(2, S, B, 4 * H, dtype=F32, device=dev)
        c_all = torch.empty(2, S, B, H, dtype=F32, device=dev)
        h_dev = _lib.handle(dev.index if dev.index is not None else torch.cuda.current_device())
        _lib.check(lib.icka_lstm_bidir_fwd_save(h_dev, gx.data_ptr(), wh_ops[0].data_ptr(), wh_ops[1].data_ptr(),
                                                y_op.data_ptr(), y32.data_ptr(), acts.data_ptr(), c_all.data_ptr(),
                                                B, S, H, torch.cuda.current_stream(dev).cuda_stream),
                   'icka_lstm_bidir_fwd_save')
        ctx.save_for_backward(x_tm.view(S * B, I), wi_op, wh_ops[0], wh_ops[1], y_op, acts, c_all)
        ctx.dims = (B, S, I, H, True)
        return y32.transpose(0, 1)

    @staticmethod
    def _backward_fused(ctx, dy):
        x_op, wi_op, wh0, wh1, y_op, acts, c_all = ctx.saved_tensors
        B, S, I, H, _ = ctx.dims
        dev = dy.device
        lib = _lib.load()
        dy_tm = dy.transpose(0, 1).contiguous()                      # no copy when dy is the view of a time-major gradient
        dg = torch.empty(S, B, 8 * H, dtype=torch.bfloat16, device=dev)
        dc = torch.empty(2, B, H, dtype=F32, device=dev)
        wt0, wt1 = wh0.t().contiguous(), wh1.t().contiguous()        # W_hh^T [H, 4H]: rows = the units a CTA owns
        h_dev = _lib.handle(dev.index if dev.index is not None else torch.cuda.current_device())
        _lib.check(lib.icka_lstm_bidir_bwd(h_dev, dy_tm.data_ptr(), wt0.data_ptr(), wt1.data_ptr(), acts.data_ptr(),
                                           c_all.data_ptr(), dg.data_ptr(), dc.data_ptr(), B, S, H,
                                           torch.cuda.current_stream(dev).cuda_stream), 'icka_lstm_bidir_bwd')
        dg2 = dg.view(S * B, 8 * H)
        # h_{t-1} of every step = the output sequence shifted by one position (zeros at the sequence ends)
        hprev = torch.zeros(S, B, 2 * H, dtype=torch.bfloat16, device=dev)
        hprev[1:, :, :H] = y_op[:-1, :, :H]
        hprev[:-1, :, H:] = y_op[1:, :, H:]
        hp2 = hprev.view(S * B, 2 * H)
        d_whh = ops.linear_wgrad(dg2[:, :4 * H], hp2[:, :H])
        d_whh_r = ops.linear_wgrad(dg2[:, 4 * H:], hp2[:, H:])
        d_wih = ops.linear_wgrad(dg2, x_op)
        db = ops.colsum(dg2)
        dx = ops.linear_dgrad(dg2, wi_op, out_dtype=F32).view(S, B, I).transpose(0, 1)
        return (dx, d_wih[:4 * H], d_whh, db[:4 * H], db[:4 * H], d_wih[4 * H:], d_whh_r, db[4 * H:], db[4 * H:], None)

    @staticmethod
    def backward(ctx, dy):
        if ctx.fused_steps:
            return BiLstmFn._backward_fused(ctx, dy)
        x_op, wi_op, wh0, wh1, y_op, acts, c_all = ctx.saved_tensors
        B, S, I, H, bf16 = ctx.dims
        cdt = torch.bfloat16 if bf16 else F32
        dev = dy.device
        dy = dy.contiguous()
        dg = torch.empty(B, S, 8 * H, dtype=cdt, device=dev)           # gate pre-activation gradients, position order
        lib = _lib.load()
        esz_dt = _lib.BF16 if bf16 else _lib.F32

        def run_direction(d):
            wh = wh1 if d else wh0
            h_dev = _lib.handle(dev.index if dev.index is not None else torch.cuda.current_device())
            dc = torch.empty(B, H, dtype=F32, device=dev)
            dh = torch.empty(B, H, dtype=F32, device=dev)
            st = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.icka_lstm_dir_bwd(
                h_dev, dy[:, :, d * H:].data_ptr(), S * 2 * H, 2 * H, wh.data_ptr(), acts[d].data_ptr(), c_all[d].data_ptr(),
                dg[:, :, d * 4 * H:].data_ptr(), S * 8 * H, 8 * H, dc.data_ptr(), dh.data_ptr(), esz_dt, B, S, H, d, st),
                'icka_lstm_dir_bwd')
            dc.record_stream(torch.cuda.current_stream(dev))
            dh.record_stream(torch.cuda.current_stream(dev))

        _two_streams(dev, run_direction)
        return BiLstmFn._weight_grads(ctx, dg, x_op, wi_op, y_op, B, S, I, H, cdt, dev)

    @staticmethod
    def _weight_grads(ctx, dg, x_op, wi_op, y_op, B, S, I, H, cdt, dev):
        """Weight, bias and input gradients from the gate pre-activation gradients of all steps: four big GEMMs."""
        dg2 = dg.view(B * S, 8 * H)
        # h_{t-1} of every step = the output sequence shifted by one position (zeros at the sequence ends)
        hprev = torch.zeros(B, S, 2 * H, dtype=cdt, device=dev)
        hprev[:, 1:, :H] = y_op[:, :-1, :H]
        hprev[:, :-1, H:] = y_op[:, 1:, H:]
        hp2 = hprev.view(B * S, 2 * H)
        d_whh = ops.linear_wgrad(dg2[:, :4 * H], hp2[:, :H])
        d_whh_r = ops.linear_wgrad(dg2[:, 4 * H:], hp2[:, H:])
        d_wih = ops.linear_wgrad(dg2, x_op)
        db = ops.colsum(dg2)
        dx = ops.linear_dgrad(dg2, wi_op, out_dtype=F32).view(B, S, I)
        return (dx, d_wih[:4 * H], d_whh, db[:4 * H], db[:4 * H], d_wih[4 * H:], d_whh_r, db[4 * H:], db[4 * H:], None)
