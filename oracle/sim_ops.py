"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

`SimOps` implements the `interactron_b200.ops.CudaOps` interface with plain torch
ops on the CPU (fp32, or fp64 for tight checks).  It lets the CPU test-suite run the
hand-derived forward/backward orchestration (detr_t / fusion_a / fusion_b / episode)
against the reference's autograd without a GPU.  TF32 rounding flags are accepted
and ignored (the simulation is exact fp32/fp64 arithmetic).
"""
import torch
import torch.nn.functional as F


class SimOps:
    name = "sim"
    _clean = False      # exact arithmetic: no TF32-clean operand twins
    precision = "sim"

    def __init__(self, dtype=torch.float32):
        self.dtype = dtype
        self.device = torch.device("cpu")
        self.n_tf32 = 0
        self.n_simt = 0
        self.calls = 0

    def empty(self, *shape):
        return torch.zeros(*shape, dtype=self.dtype)

    def zeros(self, *shape):
        return torch.zeros(*shape, dtype=self.dtype)

    def launch_count(self):
        return self.calls

    def matmul(self, a, b, *, bias=None, act=None, residual=None, out=None, out_pre=None,
               alpha=1.0, accumulate=False, epi=None, aux=None, rnd=False, act_after_residual=False,
               out_pad=False):
        self.calls += 1
        v = alpha * torch.matmul(a, b)
        if bias is not None:
            v = v + bias.unsqueeze(-2)
        if out_pre is not None:
            out_pre.copy_(v.reshape(out_pre.shape) if v.numel() == out_pre.numel() else v)
        late_act = act if act_after_residual else None
        if act_after_residual:
            act = None
        if act == "relu":
            v = torch.relu(v)
        elif act == "gelu":
            v = F.gelu(v)
        if epi == "relu_mask":
            v = v * (aux > 0).to(v.dtype)
        elif epi == "gelu_grad":
            x = aux
            cdf = 0.5 * (1 + torch.erf(x * 0.7071067811865476))
            pdf = torch.exp(-0.5 * x * x) * 0.3989422804014327
            v = v * (cdf + x * pdf)
        if residual is not None:
            v = v + residual
        if accumulate:
            v = v + out.reshape(v.shape)
        if late_act == "relu":
            v = torch.relu(v)
        elif late_act == "gelu":
            v = F.gelu(v)
        if out is None:
            return v.contiguous()
        out.copy_(v.reshape(out.shape) if v.shape != out.shape else v)
        return out

    def layernorm_fwd(self, x, gamma, beta, eps=1e-5):
        self.calls += 1
        rows, cols = x.shape
        g2, b2 = gamma.reshape(-1, cols), beta.reshape(-1, cols)
        G = g2.shape[0]
        mean = x.mean(-1)
        var = x.var(-1, unbiased=False)
        rstd = (var + eps).rsqrt()
        xh = (x - mean[:, None]) * rstd[:, None]
        y = (xh.view(G, -1, cols) * g2[:, None] + b2[:, None]).reshape(rows, cols)
        return y, y, mean, rstd

    def layernorm_bwd(self, dy, x, mean, rstd, gamma, dgamma=None, dbeta=None):
        self.calls += 1
        rows, cols = x.shape
        g2 = gamma.reshape(-1, cols)
        G = g2.shape[0]
        xh = (x - mean[:, None]) * rstd[:, None]
        dg_ = (dy.view(G, -1, cols) * g2[:, None]).reshape(rows, cols)
        m1 = dg_.mean(-1, keepdim=True)
        m2 = (dg_ * xh).mean(-1, keepdim=True)
        dx = rstd[:, None] * (dg_ - m1 - xh * m2)
        if dgamma is not None:
            Gd = dgamma.shape[0]
            dgamma.copy_((dy * xh).view(Gd, -1, cols).sum(1))
            dbeta.copy_(dy.view(Gd, -1, cols).sum(1))
        return dx, dx

    def softmax_(self, s, cols, scale, key_mask=None, rows_per_mask=1):
        self.calls += 1
        ld = s.shape[-1]
        v = s.reshape(-1, ld)[:, :cols] * scale
        if key_mask is not None:
            m = key_mask.reshape(-1, cols).bool().repeat_interleave(rows_per_mask, dim=0)
            v = v.masked_fill(m, float("-inf"))
        s.reshape(-1, ld)[:, :cols] = torch.softmax(v, -1)
        return s

    def softmax_bwd_(self, p, dp, cols, scale):
        self.calls += 1
        ld = p.shape[-1]
        pv = p.reshape(-1, ld)[:, :cols]
        dv = dp.reshape(-1, ld)[:, :cols]
        dp.reshape(-1, ld)[:, :cols] = scale * pv * (dv - (pv * dv).sum(-1, keepdim=True))
        return dp

    def dropout(self, x, key, residual=None, out=None):
        """Same mask function as itn_dropout (oracle/philox.py); key = (p, seed int tensor[1], site)."""
        from . import philox
        self.calls += 1
        p_, seed, site = key
        x2 = x if x.dim() == 2 else x.reshape(-1, x.shape[-1])
        rows, cols = x2.shape
        keep = torch.from_numpy(philox.keep_mask(int(seed.item()) & 0xFFFFFFFFFFFFFFFF, int(site), rows, cols, float(p_)))
        y = x2 * keep.to(x.dtype) * (1.0 / (1.0 - float(p_)))
        if residual is not None:
            y = residual.reshape(rows, cols) + y
        if out is None:
            return y.reshape(x.shape).contiguous()
        out.copy_(y.reshape(out.shape) if out.numel() == y.numel() and out.is_contiguous() else y)
        return out

    def ckpt_accumulate_(self, acc, x, w, first):
        self.calls += 1
        wx = torch.as_tensor(w, dtype=x.dtype) * x
        acc.copy_(wx if first else acc + wx)
        return acc

    def colsum(self, x, out=None):
        self.calls += 1
        if x.dim() == 2:
            x = x.unsqueeze(0)
        if out is None:
            return x.sum(1)
        out.copy_(x.sum(1))
        return out

    def add(self, a, b, rnd=False):
        self.calls += 1
        n = a.numel()
        if b.dim() >= 2 and b.shape[0] == a.shape[0] and b.shape[0] > 1 and n != b.numel():
            G = a.shape[0]
            return (a.reshape(G, -1, b.numel() // G) + b.reshape(G, 1, -1)).reshape(a.shape)
        return (a.reshape(-1, b.numel()) + b.reshape(1, -1)).reshape(a.shape)

    def copy2d_(self, dst, src, rnd=False):
        self.calls += 1
        dst.copy_(src)
        return dst

    def transpose_(self, dst, src):
        self.calls += 1
        dst.copy_(src.transpose(-1, -2))
        return dst

    def round_tf32(self, x, out=None):
        self.calls += 1
        if out is None:
            return x.clone()
        if out.data_ptr() != x.data_ptr():
            out.copy_(x)
        return out

    def sigmoid(self, x):
        self.calls += 1
        return torch.sigmoid(x)

    def sigmoid_bwd(self, dy, y):
        self.calls += 1
        return dy * y * (1 - y)

    def l2norm_fwd_bwd(self, x):
        self.calls += 1
        nrm = x.norm(dim=1)
        return nrm, x / nrm[:, None]

    # ---- tangent (forward-mode) rules used by interactron_b200.dual.DualOps --------------------
    @staticmethod
    def _z(t, like):
        return torch.zeros_like(like) if t is None else t

    def mask_mul_(self, y, ref):
        self.calls += 1
        y.mul_((ref > 0).to(y.dtype).reshape(y.shape))
        return y

    def gelu_grad_dual(self, raw, raw_dot, aux, aux_dot, out=None):
        """y = raw * gelu'(aux);  y_dot = raw_dot * gelu'(aux) + raw * gelu''(aux) * aux_dot."""
        self.calls += 1
        x = aux.reshape(raw.shape)
        cdf = 0.5 * (1 + torch.erf(x * 0.7071067811865476))
        pdf = torch.exp(-0.5 * x * x) * 0.3989422804014327
        g1 = cdf + x * pdf
        y = raw * g1
        if out is not None:
            out.copy_(y)
            y = out
        if raw_dot is None and aux_dot is None:
            return y, None
        yd = torch.zeros_like(raw)
        if raw_dot is not None:
            yd = yd + raw_dot * g1
        if aux_dot is not None:
            yd = yd + raw * pdf * (2 - x * x) * aux_dot.reshape(raw.shape)
        return y, yd

    def layernorm_fwd_jvp(self, x, x_dot, mean, rstd, gamma, gamma_dot, beta_dot):
        self.calls += 1
        rows, cols = x.shape
        xh = (x - mean[:, None]) * rstd[:, None]
        g2 = gamma.reshape(-1, cols)
        G = g2.shape[0]
        yd = torch.zeros_like(x)
        if x_dot is not None:
            xc = x_dot - x_dot.mean(-1, keepdim=True)
            m = (xh * xc).mean(-1, keepdim=True)
            xhd = rstd[:, None] * (xc - xh * m)
            yd = yd + (xhd.view(G, -1, cols) * g2[:, None]).reshape(rows, cols)
        if gamma_dot is not None:
            gd = gamma_dot.reshape(-1, cols)
            yd = yd + (xh.view(gd.shape[0], -1, cols) * gd[:, None]).reshape(rows, cols)
        if beta_dot is not None:
            bd = beta_dot.reshape(-1, cols)
            yd = (yd.view(bd.shape[0], -1, cols) + bd[:, None]).reshape(rows, cols)
        return yd

    def layernorm_bwd_jvp(self, dy, dy_dot, x, x_dot, mean, rstd, gamma, gamma_dot, dgamma_dot=None, dbeta_dot=None):
        self.calls += 1
        rows, cols = x.shape
        r = rstd[:, None]
        xh = (x - mean[:, None]) * r
        g2 = gamma.reshape(-1, cols)
        G = g2.shape[0]
        dy_dot = self._z(dy_dot, dy)
        u = (dy.view(G, -1, cols) * g2[:, None]).reshape(rows, cols)
        ud = (dy_dot.view(G, -1, cols) * g2[:, None]).reshape(rows, cols)
        if gamma_dot is not None:
            gd = gamma_dot.reshape(-1, cols)
            ud = ud + (dy.view(gd.shape[0], -1, cols) * gd[:, None]).reshape(rows, cols)
        a = u.mean(-1, keepdim=True)
        b = (u * xh).mean(-1, keepdim=True)
        dx = r * (u - a - xh * b)
        if x_dot is not None:
            xc = x_dot - x_dot.mean(-1, keepdim=True)
            m = (xh * xc).mean(-1, keepdim=True)
            xhd = r * (xc - xh * m)
        else:
            m = torch.zeros_like(a)
            xhd = torch.zeros_like(x)
        ad = ud.mean(-1, keepdim=True)
        bd = (ud * xh + u * xhd).mean(-1, keepdim=True)
        dxd = -r * m * dx + r * (ud - ad - xhd * b - xh * bd)
        if dgamma_dot is not None:
            Gd = dgamma_dot.shape[0]
            dgamma_dot.copy_((dy_dot * xh + dy * xhd).view(Gd, -1, cols).sum(1))
            dbeta_dot.copy_(dy_dot.view(Gd, -1, cols).sum(1))
        return dxd

    def softmax_bwd_jvp_(self, p, p_dot, dp, dp_dot, cols, scale):
        """dp <- dS = scale*p*(dp - sum p dp);  dp_dot <- d/d(eps) of the same."""
        self.calls += 1
        ld = p.shape[-1]
        pv = p.reshape(-1, ld)[:, :cols]
        dv = dp.reshape(-1, ld)[:, :cols].clone()
        dd = dp_dot.reshape(-1, ld)[:, :cols].clone()
        r = (pv * dv).sum(-1, keepdim=True)
        rd = (pv * dd).sum(-1, keepdim=True)
        out_d = pv * (dd - rd)
        if p_dot is not None:
            pd = p_dot.reshape(-1, ld)[:, :cols]
            rd2 = (pd * dv).sum(-1, keepdim=True)
            out_d = out_d + pd * (dv - r) - pv * rd2
        dp.reshape(-1, ld)[:, :cols] = scale * pv * (dv - r)
        dp_dot.reshape(-1, ld)[:, :cols] = scale * out_d
        return dp

    def sigmoid_bwd_jvp(self, dy, dy_dot, y, y_dot):
        self.calls += 1
        out = torch.zeros_like(y)
        if dy_dot is not None:
            out = out + dy_dot * y * (1 - y)
        if y_dot is not None:
            out = out + dy * (1 - 2 * y) * y_dot
        return out

    def l2norm_jvp(self, x_dot, nrm, d):
        self.calls += 1
        nd = (d * x_dot).sum(1)
        return nd, (x_dot - d * nd[:, None]) / nrm[:, None]

    def mul_mask_u8(self, x, mask, scale=1.0):
        self.calls += 1
        return x * mask.to(x.dtype) * scale

    def matcher_cost(self, logits, boxes, tgt_boxes, tgt_labels, tgt_off, w_class, w_bbox, w_giou):
        from . import port
        self.calls += 1
        off = tgt_off.tolist()
        parts = []
        for f in range(logits.shape[0]):
            lo, hi = off[f], off[f + 1]
            if hi > lo:
                parts.append(port.hungarian_cost(logits[f], boxes[f], tgt_labels[lo:hi], tgt_boxes[lo:hi],
                                                 w_class, w_bbox, w_giou).reshape(-1))
        return torch.cat(parts) if parts else torch.zeros(1)

    def criterion(self, logits, boxes, tgt_boxes, tgt_labels, tgt_off, match_row, match_tgt, match_off,
                  groups, background_c, weights=(1.0, 1.0, 1.0), want_grad=False):
        from . import port
        self.calls += 1
        Fn, Q, Cn = logits.shape
        per = Fn // groups
        off, moff = tgt_off.tolist(), match_off.tolist()
        lg = logits.detach().clone().requires_grad_(want_grad)
        bx = boxes.detach().clone().requires_grad_(want_grad)
        rows_all, tg_all = match_row.tolist(), match_tgt.tolist()
        losses = []
        total = 0
        for gi in range(groups):
            fr = range(gi * per, (gi + 1) * per)
            targets = [{"labels": tgt_labels[off[f]:off[f + 1]], "boxes": tgt_boxes[off[f]:off[f + 1]]} for f in fr]
            idx = [([], []) for _ in fr]
            for r, t in zip(rows_all[moff[gi]:moff[gi + 1]], tg_all[moff[gi]:moff[gi + 1]]):
                f = r // Q
                idx[f - gi * per][0].append(r - f * Q)
                idx[f - gi * per][1].append(t - off[f])
            idx = [(torch.tensor(i, dtype=torch.int64), torch.tensor(j, dtype=torch.int64)) for i, j in idx]
            out = port.set_criterion(lg[gi * per:(gi + 1) * per], bx[gi * per:(gi + 1) * per], targets, idx,
                                     background_c)
            losses.append(torch.stack([out[k].detach().to(logits.dtype) for k in
                                       ("loss_ce", "class_error", "cardinality_error", "loss_bbox", "loss_giou")]))
            total = total + weights[0] * out["loss_ce"] + weights[1] * out["loss_bbox"] + weights[2] * out["loss_giou"]
        losses = torch.stack(losses)
        if not want_grad:
            return losses
        dl, db = torch.autograd.grad(total, (lg, bx))
        return losses, dl, db

    def sgd_clip_update(self, theta, g, lr, clip=0.01, want_mask=False):
        self.calls += 1
        th = theta if theta.dim() == 2 else theta[None]
        step = lr * g
        out = th - torch.clip(step, min=-clip, max=clip)
        if want_mask:
            return out, out, (step.abs() <= clip).to(torch.uint8)
        return out, out

    # trainer step: torch.nn.utils.clip_grad_norm_ + torch.optim.Adam restated on flat buffers
    # (reference engine/interactron_trainer.py:106-110; formulas of torch/optim/adam.py _single_tensor_adam)
    def sumsq_partials(self, g):
        self.calls += 1
        return (torch.nan_to_num(g.reshape(-1), nan=0.0) ** 2).sum().reshape(1)      # NaN marks "no gradient"

    def clip_adam_step_(self, w, g, m, v, partials, max_norm, lr, betas, eps, step, zero_grad=False, norm_out=None):
        self.calls += 1
        skip = torch.isnan(g)                       # the simulator treats every NaN as the no-gradient marker
        w0, m0, v0 = w.clone(), m.clone(), v.clone()
        g.copy_(torch.nan_to_num(g, nan=0.0))
        if partials is not None:
            total = partials.sum().sqrt()
            if norm_out is not None:
                norm_out.copy_(total)
            if max_norm > 0:
                g.mul_(torch.clamp(max_norm / (total + 1e-6), max=1.0))
        b1, b2 = betas
        m.lerp_(g, 1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
        denom = (v.sqrt() / (bc2 ** 0.5)).add_(eps)
        w.addcdiv_(m, denom, value=-(lr / bc1))
        for t, t0 in ((w, w0), (m, m0), (v, v0)):
            t.copy_(torch.where(skip, t0, t))
        if zero_grad:
            g.zero_()
        else:
            g.masked_fill_(skip, float("nan"))

    def im2col_nhwc(self, x, kh, kw, stride, pad, dil):
        self.calls += 1
        N, H, W, Cc = x.shape
        Ho = (H + 2 * pad - dil * (kh - 1) - 1) // stride + 1
        Wo = (W + 2 * pad - dil * (kw - 1) - 1) // stride + 1
        cols = F.unfold(x.permute(0, 3, 1, 2), (kh, kw), dilation=dil, padding=pad, stride=stride)  # [N, C*kh*kw, L]
        cols = cols.view(N, Cc, kh * kw, Ho * Wo).permute(0, 3, 2, 1).reshape(N * Ho * Wo, kh * kw * Cc)
        ld = (kh * kw * Cc + 3) // 4 * 4
        out = torch.zeros(N * Ho * Wo, ld, dtype=x.dtype)
        out[:, :kh * kw * Cc] = cols
        return out, Ho, Wo

    def maxpool3x3s2_nhwc(self, x):
        self.calls += 1
        return F.max_pool2d(x.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1).contiguous()

    def pos_embed_sine(self, mask, feats=128):
        self.calls += 1
        not_mask = ~mask.bool()
        y_embed = not_mask.cumsum(1, dtype=torch.float32)
        x_embed = not_mask.cumsum(2, dtype=torch.float32)
        eps, scale = 1e-6, 2 * 3.141592653589793
        y_embed = y_embed / (y_embed[:, -1:, :] + eps) * scale
        x_embed = x_embed / (x_embed[:, :, -1:] + eps) * scale
        dim_t = torch.arange(feats, dtype=torch.float32)
        dim_t = 10000 ** (2 * (dim_t // 2) / feats)
        px = x_embed[:, :, :, None] / dim_t
        py = y_embed[:, :, :, None] / dim_t
        px = torch.stack((px[:, :, :, 0::2].sin(), px[:, :, :, 1::2].cos()), dim=4).flatten(3)
        py = torch.stack((py[:, :, :, 0::2].sin(), py[:, :, :, 1::2].cos()), dim=4).flatten(3)
        F_, h, w = mask.shape
        return torch.cat((py, px), dim=3).reshape(F_, h * w, 2 * feats).to(self.dtype)
