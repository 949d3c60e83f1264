"""Training path: every Unet3D block as one torch.autograd.Function whose forward AND backward are fixed
sequences of `vmm_*` kernel launches (blocks.py holds the forward sequences).

Conventions
  * differentiable tensor inputs of a Function are activations (16-bit channels-last) and the small fp32
    conditioning tensors (scale/shift, cond keys|values, position bias); autograd only routes those.
  * parameter gradients are NOT returned to autograd: the backward kernels accumulate straight into
    `param.grad` (views of one flat fp32 arena, see `GradArena`), in the parameter's own layout.  That is what
    makes a single fused Adam / all-reduce over the arena possible and avoids one add kernel per parameter.
  * the 31 parameters the shipped configuration never uses simply keep a zero gradient on
    every rank - no unused-parameter discovery is needed.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import blocks, ops
from .blocks import _flat

Tensor = torch.Tensor


# ------------------------------------------------------------------------------------------------
# flat parameter / gradient arenas
# ------------------------------------------------------------------------------------------------
class GradArena:
    """All trainable parameters (and their gradients) of a module as views of two flat fp32 buffers."""

    def __init__(self, module: torch.nn.Module):
        params, seen = [], set()
        for p in module.parameters():
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                params.append(p)
        self.params = params
        n = sum(p.numel() for p in params)
        dev = params[0].device
        with torch.no_grad():
            # ONE concatenation kernel (not one copy per parameter: model load must not bury the vmm kernels under ~400 tiny launches)
            self.flat_param = torch.cat([p.data.reshape(-1).to(torch.float32) for p in params])
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        o = 0
        for p in params:
            k = p.numel()
            p.data = self.flat_param[o:o + k].view(p.shape)
            p.grad = self.flat_grad[o:o + k].view(p.shape)
            o += k
        self.numel = n

    def zero_grad(self) -> None:
        self.flat_grad.zero_()

    def reattach(self) -> None:
        """Make sure .grad still points into the arena (optimizers / users may have set it to None)."""
        o = 0
        for p in self.params:
            k = p.numel()
            if p.grad is None or p.grad.data_ptr() != self.flat_grad[o:o + k].data_ptr():
                p.grad = self.flat_grad[o:o + k].view(p.shape)
            o += k


def get_arena(model) -> GradArena:
    arena = getattr(model, "_vmm_arena", None)
    if arena is None or arena.flat_param.device != next(model.parameters()).device:
        arena = GradArena(model)
        model._vmm_arena = arena
    arena.reattach()
    return arena


class _Env:
    """What the Functions need besides tensors."""

    def __init__(self, model):
        self.model = model
        self.P = model.packed()
        self.sd = blocks.param_dict(model)
        self.mode = model.padding_mode
        self.groups = model.groups
        self.heads = model.heads
        self.cond = None                 # blocks.CondState when the conditioning path runs on its kernels


# ------------------------------------------------------------------------------------------------
# ResnetBlock
# ------------------------------------------------------------------------------------------------
class ResnetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, env: _Env, pre: str, ss: Optional[Tensor], *xs: Tensor):
        xs = tuple(x.contiguous() for x in xs)
        out, saved = blocks.resnet_fwd(env.P, env.sd, pre, xs, ss, env.groups, env.mode)
        ctx.env, ctx.pre, ctx.n_src, ctx.has_ss = env, pre, len(xs), ss is not None
        ctx.save_for_backward(*xs, *saved, *( (ss,) if ss is not None else ()))
        return out

    @staticmethod
    def backward(ctx, dout: Tensor):
        env, pre, P, sd, g = ctx.env, ctx.pre, ctx.env.P, ctx.env.sd, ctx.env.groups
        t = ctx.saved_tensors
        xs = t[:ctx.n_src]
        h1, st1, a1, h2, st2 = t[ctx.n_src:ctx.n_src + 5]
        ss = t[ctx.n_src + 5] if ctx.has_ss else None
        dout = dout.contiguous()
        B, Fr, H, W, cout = dout.shape
        pix = Fr * H * W
        cins = [x.shape[-1] for x in xs]
        cin = sum(cins)
        dxs = [torch.empty_like(x) for x in xs]
        d2 = _flat(dout)
        has_res = (pre + "res.w") in P
        if has_res:
            # 1x1 res_conv: data gradient first (block1's data gradient is added on top of it below)
            ops.linear_rows([d2], P[pre + "res.wd"], cin, _flat(dxs[0]), out2=_flat(dxs[1]) if len(xs) > 1 else None, nsplit=cins[0])
            ops.wgrad_linear(d2, [_flat(x) for x in xs], sd[pre + "res_conv.weight"].grad)
            ops.colsum(d2, sd[pre + "res_conv.bias"].grad)
        # block2: GroupNorm + SiLU backward, then conv gradients
        dh2 = torch.empty_like(h2)
        ops.gn_silu_bwd(h2, dout, dh2, st2, sd[pre + "block2.norm.weight"], sd[pre + "block2.norm.bias"], None, B, pix, cout, g,
                        sd[pre + "block2.norm.weight"].grad, sd[pre + "block2.norm.bias"].grad, None,
                        dx_colsum=sd[pre + "block2.proj.bias"].grad)       # conv bias gradient = column sums of dh2
        ops.wgrad_conv3x3(ops.as_bfhwc(dh2), [ops.as_bfhwc(a1)], sd[pre + "block2.proj.weight"].grad, mode=env.mode)
        da1 = torch.empty_like(a1)
        ops.conv3x3([ops.as_bfhwc(dh2)], P[pre + "block2.wd"], cout, da1, mode=env.mode)      # the adjoint of a wrap-mode conv is one too
        # block1
        dh1 = dh2   # reuse the buffer
        if ss is not None and env.cond is not None:
            dss = env.cond.grad_views()[0][pre]          # accumulated in place; vmm_cond_bwd consumes the whole buffer after the pass
        else:
            dss = torch.zeros_like(ss) if ss is not None else None
        ops.gn_silu_bwd(h1, da1, dh1, st1, sd[pre + "block1.norm.weight"], sd[pre + "block1.norm.bias"], ss, B, pix, cout, g,
                        sd[pre + "block1.norm.weight"].grad, sd[pre + "block1.norm.bias"].grad, dss,
                        dx_colsum=sd[pre + "block1.proj.bias"].grad)
        ops.wgrad_conv3x3(ops.as_bfhwc(dh1), [ops.as_bfhwc(x) for x in xs], sd[pre + "block1.proj.weight"].grad, mode=env.mode)
        if has_res:
            res, res2 = dxs[0], (dxs[1] if len(xs) > 1 else None)     # in-place accumulate on the res_conv gradient
        else:
            res, res2 = dout, None                                      # identity skip
        ops.conv3x3([ops.as_bfhwc(dh1)], P[pre + "block1.wd"], cin, dxs[0], mode=env.mode,
                    out2=dxs[1] if len(xs) > 1 else None, nsplit=cins[0], res=res, res2=res2)
        return (None, None, None if env.cond is not None else dss, *dxs)


# ------------------------------------------------------------------------------------------------
# Residual(PreNorm(attention))
# ------------------------------------------------------------------------------------------------
class AttnBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, env: _Env, pre: str, kind: str, x: Tensor, ekv: Optional[Tensor], bias: Optional[Tensor], rot: Optional[Tensor]):
        x = x.contiguous()
        out, (xn, qkv, ao, extra) = blocks.attn_block_fwd(env.P, env.sd, pre, kind, x, ekv, bias, rot, env.heads)
        ctx.env, ctx.pre, ctx.kind = env, pre, kind
        ctx.has_ekv, ctx.has_bias = ekv is not None, bias is not None
        ctx.n_extra = len(extra) if extra is not None else 0
        tensors = [x, xn, qkv, ao] + list(extra or ())
        if ekv is not None:
            tensors.append(ekv)
        if bias is not None:
            tensors += [bias, rot]
        ctx.save_for_backward(*tensors)
        return out

    @staticmethod
    def backward(ctx, dout: Tensor):
        env, pre, kind, P, sd, heads = ctx.env, ctx.pre, ctx.kind, ctx.env.P, ctx.env.sd, ctx.env.heads
        t = list(ctx.saved_tensors)
        x, xn, qkv, ao = t[:4]
        extra = t[4:4 + ctx.n_extra]
        k = 4 + ctx.n_extra
        ekv = None
        if ctx.has_ekv:
            ekv = t[k]
            k += 1
        bias = rot = None
        if ctx.has_bias:
            bias, rot = t[k], t[k + 1]
        dout = dout.contiguous()
        B, Fr, H, W, Cc = x.shape
        hd = heads * 32
        d2 = _flat(dout)
        # to_out
        dao = torch.empty_like(ao)
        ops.linear_rows([d2], P[pre + "out.wd"], hd, dao)
        ops.wgrad_linear(d2, [ao], sd[pre + "to_out.weight"].grad)
        if (pre + "to_out.bias") in sd:
            ops.colsum(d2, sd[pre + "to_out.bias"].grad)
        # attention core
        dqkv = torch.empty_like(qkv)
        dbias = None
        if env.cond is not None:
            _, d_ekv, d_bias = env.cond.grad_views()
            dekv = d_ekv[pre] if ekv is not None else None
        else:
            dekv = torch.zeros_like(ekv) if ekv is not None else None
        if kind == "temporal":
            dbias = d_bias if env.cond is not None else torch.zeros_like(bias)
            ops.tattn_bwd(qkv, ekv, bias, rot[1], dao, dqkv, dekv, dbias, B, Fr, H * W, heads, pre_rotated=blocks.ROTARY_IN_EPILOGUE)
        elif kind == "linear":
            ctxm, kstat = extra
            dctx = torch.empty_like(ctxm)
            ops.lattn_bwd(qkv, ekv, ekv.shape[1], dao, ctxm, kstat, dctx, dqkv, dekv, B * Fr, Fr, H * W, heads)
        else:
            (lse,) = extra
            ops.sattn_bwd(qkv, ekv, ao, dao, lse, dqkv, dekv, B * Fr, H * W, heads)
        # to_qkv
        norm_pre = pre[: pre.index("fn.fn.") + 3]
        gamma = sd[norm_pre + "norm.gamma"]
        dx = torch.empty_like(x)
        if ops.qkv_bwd_eligible(_flat(dqkv), _flat(xn)) and ops.FUSED_QKV_LN_BWD:
            # 64-channel levels: ONE pass over the 768-wide gradient rows for the data gradient and the weight gradient of to_qkv, with the
            # PreNorm backward + the Residual skip in the epilogue (csrc/qkvbwd.cu): dxn never reaches HBM
            ops.qkv_ln_bwd(_flat(dqkv), _flat(xn), P[pre + "qkv.wd"], _flat(x), d2, gamma.reshape(-1), _flat(dx), sd[pre + "to_qkv.weight"].grad,
                           gamma.grad.reshape(-1))
        else:
            dxn = torch.empty_like(xn)
            if ops.qkv_bwd_eligible(_flat(dqkv), _flat(xn)):
                ops.qkv_bwd(_flat(dqkv), _flat(xn), P[pre + "qkv.wd"], _flat(dxn), sd[pre + "to_qkv.weight"].grad)
            else:
                ops.linear_rows([dqkv], P[pre + "qkv.wd"], Cc, dxn)
                ops.wgrad_linear(dqkv, [xn], sd[pre + "to_qkv.weight"].grad)
            # PreNorm + the Residual skip
            ops.ln_bwd(_flat(x), dxn, d2, _flat(dx), gamma.reshape(-1), gamma.grad.reshape(-1))
        if env.cond is not None:
            return None, None, None, dx, None, None, None
        return None, None, None, dx, dekv, dbias, None


class DownFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, env: _Env, pre: str, x: Tensor):
        x = x.contiguous()
        ctx.env, ctx.pre = env, pre
        ctx.save_for_backward(x)
        return blocks.down_fwd(env.P, env.sd, pre, x, env.mode)

    @staticmethod
    def backward(ctx, dout: Tensor):
        env, pre = ctx.env, ctx.pre
        (x,) = ctx.saved_tensors
        dout = dout.contiguous()
        Cc = x.shape[-1]
        dx = torch.empty_like(x)
        ops.conv_up(ops.as_bfhwc(dout), env.P[pre + "wd"], Cc, dx, mode=env.mode)
        ops.wgrad_down(ops.as_bfhwc(dout), ops.as_bfhwc(x), env.sd[pre + "weight"].grad, mode=env.mode)
        ops.colsum(_flat(dout), env.sd[pre + "bias"].grad)
        return None, None, dx


class UpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, env: _Env, pre: str, x: Tensor):
        x = x.contiguous()
        ctx.env, ctx.pre = env, pre
        ctx.save_for_backward(x)
        return blocks.up_fwd(env.P, env.sd, pre, x, env.mode)

    @staticmethod
    def backward(ctx, dout: Tensor):
        env, pre = ctx.env, ctx.pre
        (x,) = ctx.saved_tensors
        dout = dout.contiguous()
        Cc = x.shape[-1]
        dx = torch.empty_like(x)
        ops.conv_down(ops.as_bfhwc(dout), env.P[pre + "wd"], Cc, dx, mode=env.mode)
        ops.wgrad_up(ops.as_bfhwc(dout), ops.as_bfhwc(x), env.sd[pre + "weight"].grad, mode=env.mode)
        ops.colsum(_flat(dout), env.sd[pre + "bias"].grad)
        return None, None, dx


class InitFn(torch.autograd.Function):
    """prep_input (q_sample fused) + init_conv.  `anchor` is a dummy differentiable input that makes autograd
    call backward (the data itself needs no gradient)."""

    @staticmethod
    def forward(ctx, env: _Env, anchor: Tensor, x: Tensor, noise: Optional[Tensor], a, c, s):
        out, xin = blocks.init_fwd(env.P, env.sd, env.model, x, noise, (a, c, s))
        ctx.env = env
        ctx.shape = x.shape
        ctx.save_for_backward(xin)
        return out

    @staticmethod
    def backward(ctx, dout: Tensor):
        env = ctx.env
        (xin,) = ctx.saved_tensors
        dout = dout.contiguous()
        B, Cc, Fr, H, W = ctx.shape
        ops.wgrad_init_conv(ops.as_bfhwc(dout), xin, env.sd["init_conv.weight"].grad, Cc, mode=env.mode)
        ops.colsum(_flat(dout), env.sd["init_conv.bias"].grad)
        return None, torch.zeros(1, device=dout.device), None, None, None, None, None


def _queue_cond_backward(env: _Env) -> None:
    """The conditioning kernels' backward runs once, after every block has accumulated its d(scale|shift) / d(ek|ev) / d(bias)
    into the CondState buffer: queued on the autograd engine from the FIRST backward node of the pass (loss / final conv)."""
    if env.cond is not None:
        st = env.cond
        torch.autograd.Variable._execution_engine.queue_callback(st.backward)


class LossFn(torch.autograd.Function):
    """final 1x1x1 conv (VDDP:708) + F.l1_loss / F.mse_loss against the noise (VDDP:1053-1056)."""

    @staticmethod
    def forward(ctx, env: _Env, h: Tensor, target: Tensor, l2: bool):
        h = h.contiguous()
        model = env.model
        pred = blocks.final_fwd(env.P, env.sd, model, h)                   # fp32 (b, f, h, w, c)
        B, Fr, H, W, Cc = pred.shape
        loss = torch.zeros(1, dtype=torch.float32, device=h.device)
        dpred = torch.empty(B * Fr * H * W, 8, dtype=h.dtype, device=h.device)
        ops.loss_fwd_bwd(pred, target, loss, dpred, B, Cc, Fr, H, W, l2=l2)
        ctx.env = env
        ctx.save_for_backward(h, dpred)
        return loss[0]

    @staticmethod
    def backward(ctx, gout: Tensor):
        env = ctx.env
        h, dpred = ctx.saved_tensors
        model = env.model
        _queue_cond_backward(env)
        dpred = dpred * gout.to(dpred.dtype)                                # loss scale (1 in bf16 training)
        dh = torch.empty_like(h)
        ops.linear_rows([dpred], env.P["final.wd"], h.shape[-1], _flat(dh))
        ops.wgrad_linear(dpred, [_flat(h)], env.sd["final_conv.1.weight"].grad)
        db8 = torch.zeros(8, dtype=torch.float32, device=h.device)
        ops.colsum(dpred, db8)
        env.sd["final_conv.1.bias"].grad.add_(db8[: model.channels])
        return None, dh, None, None


class FinalFn(torch.autograd.Function):
    """final 1x1x1 conv (VDDP:708) as a differentiable output: fp32 channels-last prediction (b, f, h, w, c).  Used when the network
    itself is called under autograd (Unet3D.forward with gradients enabled); the training step fuses this conv with the loss (LossFn)."""

    @staticmethod
    def forward(ctx, env: _Env, h: Tensor):
        h = h.contiguous()
        ctx.env = env
        ctx.save_for_backward(h)
        return blocks.final_fwd(env.P, env.sd, env.model, h)

    @staticmethod
    def backward(ctx, dpred: Tensor):
        env = ctx.env
        (h,) = ctx.saved_tensors
        model = env.model
        _queue_cond_backward(env)
        Cc = model.channels
        dp8 = torch.zeros(h.numel() // h.shape[-1], 8, dtype=h.dtype, device=h.device)      # rows carry 8 (zero padded) channels
        dp8[:, :Cc] = dpred.reshape(-1, Cc).to(h.dtype)
        dh = torch.empty_like(h)
        ops.linear_rows([dp8], env.P["final.wd"], h.shape[-1], _flat(dh))
        ops.wgrad_linear(dp8, [_flat(h)], env.sd["final_conv.1.weight"].grad)
        db8 = torch.zeros(8, dtype=torch.float32, device=h.device)
        ops.colsum(dp8, db8)
        env.sd["final_conv.1.bias"].grad.add_(db8[:Cc])
        return None, dh


# ------------------------------------------------------------------------------------------------
# whole network, training form
# ------------------------------------------------------------------------------------------------
def _trunk(model, x0: Tensor, noise: Optional[Tensor], qcoef, t: Tensor, cond: Tensor, null_mask: Tensor):
    """Everything of Unet3D.forward (VDDP:730-820) up to the input of the final 1x1x1 conv, as autograd Functions whose
    backward kernels accumulate parameter gradients into the arena.  Returns (env, hidden activations)."""
    if not x0.is_cuda:
        raise RuntimeError("videometamaterials_b200 has no CPU path: move the model and inputs to a CUDA device")
    get_arena(model)
    env = _Env(model)
    L = len(model.dim_mults)
    frames = x0.shape[2]
    ss, ekv, bias, rot, env.cond = blocks.conditioning_state(model, t, cond, null_mask, frames)
    blocks.begin_stats_pool(2 * len(blocks.resnet_names(model)), x0.shape[0], model.groups, x0.device)
    tkv = (lambda q: ekv[q]) if model.use_temporal_attention_cond else (lambda q: None)      # VDDP:792-795
    anchor = torch.zeros(1, device=x0.device, requires_grad=True)
    a, c, s = qcoef if qcoef is not None else (None, None, None)
    h = InitFn.apply(env, anchor, x0, noise, a, c, s)
    h = AttnBlockFn.apply(env, "init_temporal_attn.fn.fn.fn.", "temporal", h, None, bias, rot)
    r = h
    skips = []
    for i in range(L):
        p = f"downs.{i}."
        h = ResnetFn.apply(env, p + "0.", ss[p + "0."], h)
        h = ResnetFn.apply(env, p + "1.", ss[p + "1."], h)
        h = AttnBlockFn.apply(env, p + "2.fn.fn.", "linear", h, ekv[p + "2.fn.fn."], None, None)
        h = AttnBlockFn.apply(env, p + "3.fn.fn.fn.", "temporal", h, tkv(p + "3.fn.fn.fn."), bias, rot)
        skips.append(h)
        if i < L - 1:
            h = DownFn.apply(env, p + "4.", h)
    h = ResnetFn.apply(env, "mid_block1.", ss["mid_block1."], h)
    h = AttnBlockFn.apply(env, "mid_spatial_attn.fn.fn.fn.", "spatial", h, ekv["mid_spatial_attn.fn.fn.fn."], None, None)
    h = AttnBlockFn.apply(env, "mid_temporal_attn.fn.fn.fn.", "temporal", h, tkv("mid_temporal_attn.fn.fn.fn."), bias, rot)
    h = ResnetFn.apply(env, "mid_block2.", ss["mid_block2."], h)
    for i in range(L):
        p = f"ups.{i}."
        h = ResnetFn.apply(env, p + "0.", ss[p + "0."], h, skips.pop())
        h = ResnetFn.apply(env, p + "1.", ss[p + "1."], h)
        h = AttnBlockFn.apply(env, p + "2.fn.fn.", "linear", h, ekv[p + "2.fn.fn."], None, None)
        h = AttnBlockFn.apply(env, p + "3.fn.fn.fn.", "temporal", h, tkv(p + "3.fn.fn.fn."), bias, rot)
        if i < L - 1:
            h = UpFn.apply(env, p + "4.", h)
    h = ResnetFn.apply(env, "final_conv.0.", None, h, r)
    return env, h


def training_loss(model, x0: Tensor, noise: Tensor, qcoef, t: Tensor, cond: Tensor, null_mask: Tensor, l2: bool = False) -> Tensor:
    """loss(noise, Unet3D(q_sample(x0, t, noise), t, cond)) with gradients accumulated into the parameter arena.
    x0, noise: fp32 (b, c, f, h, w).  qcoef = (a[b], c[b] or None, s[b]): x_t = a x0 + c + s noise."""
    env, h = _trunk(model, x0, noise, qcoef, t, cond, null_mask)
    return LossFn.apply(env, h, noise, l2)


def unet_forward_autograd(model, x: Tensor, t: Tensor, cond: Tensor, null_mask: Tensor) -> Tensor:
    """Unet3D.forward under autograd: fp32 channels-last prediction (b, f, h, w, c) whose backward accumulates every parameter
    gradient into `param.grad` (views of the arena), like nn.Module parameters under torch autograd.  The input `x` itself
    receives no gradient (nothing in the reference's training or sampling paths asks for one)."""
    env, h = _trunk(model, x.contiguous().float(), None, None, t, cond, null_mask)
    return FinalFn.apply(env, h)
