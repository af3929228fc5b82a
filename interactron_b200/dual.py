"""Forward-mode (dual-number) execution of the kernel interface: the second-order half of the
meta-training step (reference models/interactron.py:98-123).

The reference obtains the supervisor's meta-gradient by `supervisor_loss.backward()` through the
graph of `autograd.grad(learned_loss, theta, create_graph=True)` - a reverse-over-reverse pass.
With  v = 1[|lr*g| <= clip] * dL_sup/dtheta'  held constant, that gradient is, for every parameter
w the inner gradient g(theta, w) depends on (fusion phi, in_proj psi):

    dL_sup/dw (second-order part) = -lr * d/dw <g(theta, w), v> = -lr * d/d(eps) [ dl/dw (theta + eps*v, w) ]

i.e. the directional derivative, along v in theta, of the ordinary first-order gradient dl/dw of
the learned loss l (identical to the reference's double backward, checked in
float64 in tests/test_dual_cpu.py).  So the whole thing is the *existing* forward + backward
orchestration (detr_t / fusion / layers) run once more on dual numbers (value, d/d(eps) value):
`DualOps` wraps a kernel backend and applies the tangent rule of every op, `Dual` carries the pair
through the views / slices the orchestration takes, `DualWeights` serves theta with tangent v.

Tangent `None` means structurally zero (inputs, psi, phi) and costs nothing.
"""
import torch


class Dual:
    """(primal, tangent) pair that mimics the few torch.Tensor view methods the orchestration uses.
    The tangent may have a different leading (group) extent than the primal: theta is shared by
    the episodes ([1, ...]) while its tangent v is per episode ([E, ...])."""

    __slots__ = ("p", "t")

    def __init__(self, p, t=None):
        self.p, self.t = p, t

    # -- metadata (of the primal)
    @property
    def shape(self):
        return self.p.shape

    def dim(self):
        return self.p.dim()

    def numel(self):
        return self.p.numel()

    def is_contiguous(self):
        return self.p.is_contiguous()

    def data_ptr(self):
        return self.p.data_ptr()

    # -- views
    def _lead(self, args):
        """view/reshape arguments for the tangent when its group extent differs from the primal's."""
        if len(args) == 1 and isinstance(args[0], (tuple, list, torch.Size)):
            args = tuple(args[0])
        if self.t is None or self.t.shape[0] == self.p.shape[0] or not args or args[0] != self.p.shape[0]:
            return args
        return (self.t.shape[0],) + tuple(args[1:])

    def view(self, *args):
        return Dual(self.p.view(*args), None if self.t is None else self.t.view(*self._lead(args)))

    def reshape(self, *args):
        return Dual(self.p.reshape(*args), None if self.t is None else self.t.reshape(*self._lead(args)))

    def permute(self, *dims):
        return Dual(self.p.permute(*dims), None if self.t is None else self.t.permute(*dims))

    def transpose(self, a, b):
        return Dual(self.p.transpose(a, b), None if self.t is None else self.t.transpose(a, b))

    def unsqueeze(self, d):
        return Dual(self.p.unsqueeze(d), None if self.t is None else self.t.unsqueeze(d))

    def expand(self, *sizes):
        return Dual(self.p.expand(*sizes), None if self.t is None else self.t.expand(*sizes))

    def contiguous(self):
        return Dual(self.p.contiguous(), None if self.t is None else self.t.contiguous())

    def __getitem__(self, idx):
        return Dual(self.p[idx], None if self.t is None else self.t[idx])


def _p(x):
    return x.p if isinstance(x, Dual) else x


def _t(x):
    return x.t if isinstance(x, Dual) else None


class DualWeights:
    """`params.Weights` interface over dual flat buffers: every tuple is
    (pack, flat, flat_t | None, flat_dot | None, flat_dot_t | None) - the weights, their transposed
    twins, the tangent direction and its transposed twins (same layout as the weights)."""

    def __init__(self, *tuples):
        self.tuples = tuples

    def _find(self, name):
        for t in self.tuples:
            if name in t[0]:
                return t
        raise KeyError(name)

    def p(self, name):
        pack, flat, _, dot, _ = self._find(name)
        return Dual(pack.view(flat, name), None if dot is None else pack.view(dot, name))

    w = p

    @staticmethod
    def _bwd(pack, flat, flat_t, name, lo, hi):
        if flat_t is not None:
            wt = pack.view_t(flat_t, name)
            if lo is not None:
                wt = wt[..., lo:hi]
            return wt.transpose(-1, -2)
        N, K = pack.matrix_shape(name)
        w = pack.view(flat, name).reshape(flat.shape[0], N, K)
        return w[..., lo:hi, :] if lo is not None else w

    def bwd(self, name, lo=None, hi=None):
        pack, flat, flat_t, dot, dot_t = self._find(name)
        return Dual(self._bwd(pack, flat, flat_t, name, lo, hi),
                    None if dot is None else self._bwd(pack, dot, dot_t, name, lo, hi))


class DualOps:
    """Kernel interface on `Dual`s.  `base` is the real backend (ops.CudaOps on the B200; the
    float64 torch simulation in the CPU tests).  Only what detr_t / fusion / layers call."""

    name = "dual"
    _clean = False

    def __init__(self, base):
        if base._clean:
            raise ValueError("the second-order pass runs in tf32x3 (fp32-accurate) mode only")
        self.o = base
        self.device = base.device
        self.precision = base.precision

    # ------------------------------------------------------------------ allocation
    def empty(self, *shape):
        return Dual(self.o.empty(*shape), self.o.zeros(*shape))

    def zeros(self, *shape):
        return Dual(self.o.zeros(*shape), self.o.zeros(*shape))

    def launch_count(self):
        return self.o.launch_count()

    # ------------------------------------------------------------------ GEMM
    def matmul(self, a, b, *, bias=None, act=None, residual=None, out=None, out_pre=None, alpha=1.0,
               accumulate=False, epi=None, aux=None, rnd=False, act_after_residual=False, out_pad=False):
        o = self.o
        ap, at, bp, bt = _p(a), _t(a), _p(b), _t(b)
        if act_after_residual:
            raise NotImplementedError("act_after_residual is a backbone-only epilogue")
        terms = [(x, y) for x, y in ((at, bp), (ap, bt)) if x is not None and y is not None]
        nonlinear = act is not None or epi is not None
        if nonlinear and (residual is not None or accumulate):
            raise NotImplementedError("dual GEMM: activation together with residual/accumulate")

        if epi == "gelu_grad":
            # out = (a b) * gelu'(aux): the raw product is needed for the aux-tangent term
            assert bias is None and out is None and out_pre is None
            raw = o.matmul(ap, bp, alpha=alpha)
            raw_t = self._tangent_gemm(terms, None, None, None, False, alpha, raw.shape)
            y, y_t = o.gelu_grad_dual(raw, raw_t, _p(aux), _t(aux))
            return Dual(y, y_t)

        yp = o.matmul(ap, bp, bias=_p(bias), act=act, residual=_p(residual), out=_p(out), out_pre=_p(out_pre),
                      alpha=alpha, accumulate=accumulate, epi=epi, aux=_p(aux), rnd=rnd, out_pad=out_pad)
        bias_t, res_t = _t(bias), _t(residual)
        out_t = _t(out) if out is not None else None
        pre_t = _t(out_pre) if out_pre is not None else None
        if out is not None and out_t is None:
            raise ValueError("dual GEMM: `out` must be a Dual with an allocated tangent")
        if not terms and bias_t is None and res_t is None:
            # structurally zero tangent: freshly allocated tangents are zero-initialised
            return Dual(yp, out_t) if out is not None else Dual(yp, None)
        if not nonlinear:
            yt = self._tangent_gemm(terms, bias_t, res_t, out_t, accumulate, alpha, yp.shape, out_pad)
            if pre_t is not None:
                # out_pre holds alpha*a@b + bias (before the residual): only used without residual
                assert residual is None and not accumulate
                o.copy2d_(pre_t.reshape(-1, pre_t.shape[-1]), yt.reshape(-1, yt.shape[-1]))
            return Dual(yp, yt)
        # relu / gelu forward activations and the relu-mask backward epilogue
        if act == "relu" or epi == "relu_mask":
            ref = yp if act == "relu" else _p(aux)          # relu(z) > 0  <=>  z > 0
            if len(terms) == 1 and pre_t is None:
                yt = o.matmul(terms[0][0], terms[0][1], bias=bias_t, out=out_t, alpha=alpha, epi="relu_mask",
                              aux=ref, out_pad=out_pad)
            else:
                yt = self._tangent_gemm(terms, bias_t, None, out_t, False, alpha, yp.shape, out_pad)
                if pre_t is not None:
                    o.copy2d_(pre_t.reshape(-1, pre_t.shape[-1]), yt.reshape(-1, yt.shape[-1]))
                o.mask_mul_(yt, ref)
            return Dual(yp, yt)
        if act == "gelu":
            if out_pre is None or len(terms) != 1:
                raise NotImplementedError("dual gelu GEMM needs out_pre and a single tangent term")
            z_t = o.matmul(terms[0][0], terms[0][1], bias=bias_t, out=pre_t, alpha=alpha)
            yt = o.gelu_grad_dual(z_t, None, _p(out_pre), None, out=out_t)[0]
            return Dual(yp, yt)
        raise NotImplementedError(f"dual GEMM epilogue act={act} epi={epi}")

    def _tangent_gemm(self, terms, bias_t, res_t, out_t, accumulate, alpha, shape, out_pad=False):
        """sum of the tangent products (+ bias/residual tangents, + existing out if accumulate)."""
        o = self.o
        if not terms:
            if bias_t is None and res_t is None and not accumulate:
                return None
            raise NotImplementedError("dual GEMM with only bias/residual tangents")
        y = None
        for i, (x, w) in enumerate(terms):
            first = i == 0
            y = o.matmul(x, w, bias=bias_t if first else None, residual=res_t if first else None,
                         out=out_t if first else y, alpha=alpha, accumulate=accumulate if first else True,
                         out_pad=out_pad)
        return y

    # ------------------------------------------------------------------ row-wise
    def layernorm_fwd(self, x, gamma, beta, eps=1e-5):
        o = self.o
        y, y_r, mean, rstd = o.layernorm_fwd(_p(x), _p(gamma), _p(beta), eps)
        if _t(x) is None and _t(gamma) is None and _t(beta) is None:
            return Dual(y), Dual(y_r), Dual(mean), Dual(rstd)
        yt = o.layernorm_fwd_jvp(_p(x), _t(x), mean, rstd, _p(gamma), _t(gamma), _t(beta))
        return Dual(y, yt), Dual(y_r, yt), Dual(mean), Dual(rstd)

    def layernorm_bwd(self, dy, x, mean, rstd, gamma, dgamma=None, dbeta=None):
        o = self.o
        dx, dx_r = o.layernorm_bwd(_p(dy), _p(x), _p(mean), _p(rstd), _p(gamma), _p(dgamma), _p(dbeta))
        dxt = o.layernorm_bwd_jvp(_p(dy), _t(dy), _p(x), _t(x), _p(mean), _p(rstd), _p(gamma), _t(gamma),
                                  _t(dgamma), _t(dbeta))
        return Dual(dx, dxt), Dual(dx_r, dxt)

    def softmax_(self, s, cols, scale, key_mask=None, rows_per_mask=1):
        o = self.o
        o.softmax_(s.p, cols, scale, key_mask, rows_per_mask)
        if s.t is not None:
            # the softmax Jacobian is symmetric: its JVP is the kernel of its VJP
            o.softmax_bwd_(s.p, s.t, cols, scale)
        return s

    def softmax_bwd_(self, p, dp, cols, scale):
        self.o.softmax_bwd_jvp_(_p(p), _t(p), dp.p, dp.t, cols, scale)
        return dp

    def colsum(self, x, out=None):
        o = self.o
        if out is None:
            return Dual(o.colsum(_p(x)), None if _t(x) is None else o.colsum(_t(x)))
        o.colsum(_p(x), out=out.p)
        if _t(x) is not None:
            o.colsum(_t(x), out=out.t)
        return out

    # ------------------------------------------------------------------ element-wise
    def add(self, a, b, rnd=False):
        o = self.o
        y = o.add(_p(a), _p(b), rnd=rnd)
        at, bt = _t(a), _t(b)
        if at is not None and bt is not None:
            yt = o.add(at, bt)
        elif at is not None:
            yt = at
        elif bt is not None:
            yt = o.add(o.zeros(*y.shape), bt)
        else:
            yt = None
        return Dual(y, yt)

    def dropout(self, x, key, residual=None, out=None):
        """Dropout is linear in x: the same mask on the value and on the tangent."""
        o = self.o
        yp = o.dropout(_p(x), key, residual=_p(residual), out=_p(out))
        xt, rt = _t(x), _t(residual)
        ot = _t(out) if out is not None else None
        if xt is None and rt is None:
            return Dual(yp, ot) if out is not None else Dual(yp, None)
        if xt is None:
            yt = rt if ot is None else ot.copy_(rt)
        else:
            yt = o.dropout(xt, key, residual=rt, out=ot)
        return Dual(yp, yt)

    def copy2d_(self, dst, src, rnd=False):
        o = self.o
        o.copy2d_(dst.p, _p(src), rnd=rnd)
        if _t(src) is not None:
            o.copy2d_(dst.t, _t(src))
        return dst

    def round_tf32(self, x, out=None):
        return x                                      # tf32x3 mode: operands are used at full precision

    def sigmoid(self, x):
        o = self.o
        y = o.sigmoid(_p(x))
        return Dual(y, None if _t(x) is None else o.sigmoid_bwd(_t(x), y))

    def sigmoid_bwd(self, dy, y):
        o = self.o
        dx = o.sigmoid_bwd(_p(dy), _p(y))
        return Dual(dx, o.sigmoid_bwd_jvp(_p(dy), _t(dy), _p(y), _t(y)))

    def l2norm_fwd_bwd(self, x):
        o = self.o
        nrm, d = o.l2norm_fwd_bwd(_p(x))
        if _t(x) is None:
            return Dual(nrm), Dual(d)
        nt, dt = o.l2norm_jvp(_t(x), nrm, d)
        return Dual(nrm, nt), Dual(d, dt)

    def pos_embed_sine(self, mask, feats=128):
        return Dual(self.o.pos_embed_sine(mask, feats))
