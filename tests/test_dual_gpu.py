"""GPU parity of the tangent (forward-mode) kernels of the meta-training step (csrc/itn_dual.cu,
through the C ABI) against their float64 torch restatement (oracle/sim_ops.py), and of a whole
dual-number decoder layer (forward + backward) against the same orchestration on the simulation."""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 2e-5


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def ops():
    from interactron_b200.ops import CudaOps
    return CudaOps()


@pytest.fixture(scope="module")
def sim():
    from oracle.sim_ops import SimOps
    return SimOps(torch.float64)


def rnd(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float64) * scale


def cu(t):
    return None if t is None else t.float().cuda().contiguous()


@pytest.mark.parametrize("cols,rows,G,Gd", [(256, 500, 1, 2), (512, 410, 1, 1), (256, 96, 2, 2)])
def test_layernorm_jvp(ops, sim, cols, rows, G, Gd):
    x, xd = rnd(rows, cols, seed=1), rnd(rows, cols, seed=2)
    gamma, beta = rnd(G, cols, seed=3) + 1, rnd(G, cols, seed=4)
    gd, bd = rnd(Gd, cols, seed=5), rnd(Gd, cols, seed=6)
    y, _, mean, rstd = sim.layernorm_fwd(x, gamma, beta)
    yd = sim.layernorm_fwd_jvp(x, xd, mean, rstd, gamma, gd, bd)
    _, _, m_c, r_c = ops.layernorm_fwd(cu(x), cu(gamma), cu(beta))
    yd_c = ops.layernorm_fwd_jvp(cu(x), cu(xd), m_c, r_c, cu(gamma), cu(gd), cu(bd))
    assert rel(yd_c, yd) < TOL
    assert rel(ops.layernorm_fwd_jvp(cu(x), None, m_c, r_c, cu(gamma), cu(gd), None),
               sim.layernorm_fwd_jvp(x, None, mean, rstd, gamma, gd, None)) < TOL
    dy, dyd = rnd(rows, cols, seed=7), rnd(rows, cols, seed=8)
    dg_d, db_d = torch.zeros(Gd, cols, dtype=torch.float64), torch.zeros(Gd, cols, dtype=torch.float64)
    dxd = sim.layernorm_bwd_jvp(dy, dyd, x, xd, mean, rstd, gamma, gd, dg_d, db_d)
    dg_c, db_c = torch.zeros(Gd, cols, device="cuda"), torch.zeros(Gd, cols, device="cuda")
    dxd_c = ops.layernorm_bwd_jvp(cu(dy), cu(dyd), cu(x), cu(xd), m_c, r_c, cu(gamma), cu(gd), dg_c, db_c)
    assert rel(dxd_c, dxd) < TOL
    assert rel(dg_c, dg_d) < 5 * TOL and rel(db_c, db_d) < TOL
    assert rel(ops.layernorm_bwd_jvp(cu(dy), cu(dyd), cu(x), cu(xd), m_c, r_c, cu(gamma), None),
               sim.layernorm_bwd_jvp(dy, dyd, x, xd, mean, rstd, gamma, None)) < TOL


@pytest.mark.parametrize("rows,cols", [(64, 361), (40, 50), (24, 1805), (16, 2060), (8, 255)])
def test_softmax_bwd_jvp(ops, sim, rows, cols):
    ld = (cols + 3) // 4 * 4
    s = rnd(rows, ld, seed=1)
    p = s.clone()
    sim.softmax_(p, cols, 0.3)
    pd, dp, dpd = rnd(rows, ld, seed=2, scale=0.1) * p, rnd(rows, ld, seed=3), rnd(rows, ld, seed=4)
    for with_pd in (True, False):
        a, b = dp.clone(), dpd.clone()
        sim.softmax_bwd_jvp_(p, pd if with_pd else None, a, b, cols, 0.3)
        ac, bc = cu(dp), cu(dpd)
        ops.softmax_bwd_jvp_(cu(p), cu(pd) if with_pd else None, ac, bc, cols, 0.3)
        assert rel(ac[:, :cols], a[:, :cols]) < TOL
        assert rel(bc[:, :cols], b[:, :cols]) < TOL
    # forward tangent = the backward kernel (symmetric Jacobian)
    t = rnd(rows, ld, seed=5)
    pr = (p * 1).requires_grad_(False)
    sraw = (s[:, :cols] * 1).requires_grad_(True)
    jv = torch.autograd.functional.jvp(lambda z: torch.softmax(0.3 * z, -1), sraw, t[:, :cols])[1]
    tc = cu(t)
    ops.softmax_bwd_(cu(pr), tc, cols, 0.3)
    assert rel(tc[:, :cols], jv) < TOL


def test_elementwise_tangents(ops, sim):
    n = 4 * 1237
    raw, rawd, aux, auxd = (rnd(n, seed=i) for i in range(4))
    y, yd = sim.gelu_grad_dual(raw, rawd, aux, auxd)
    yc, ydc = ops.gelu_grad_dual(cu(raw), cu(rawd), cu(aux), cu(auxd))
    assert rel(yc, y) < TOL and rel(ydc, yd) < TOL
    yc2, ydc2 = ops.gelu_grad_dual(cu(raw), None, cu(aux), None)
    assert rel(yc2, y) < TOL and ydc2 is None
    # gelu'' against autograd
    a = aux.clone().requires_grad_(True)
    g1 = torch.autograd.grad(torch.nn.functional.gelu(a).sum(), a, create_graph=True)[0]
    g2 = torch.autograd.grad(g1.sum(), a)[0]
    _, only_aux = ops.gelu_grad_dual(cu(torch.ones(n, dtype=torch.float64)), None, cu(aux), cu(torch.ones(n, dtype=torch.float64)))
    assert rel(only_aux, g2) < TOL
    ref = rnd(n, seed=9)
    v = rnd(n, seed=10)
    vc = cu(v)
    ops.mask_mul_(vc, cu(ref))
    assert torch.equal(vc.cpu(), (v * (ref > 0)).float())
    mask = (rnd(n, seed=11) > 0).to(torch.uint8)
    assert torch.equal(ops.mul_mask_u8(cu(v), mask.cuda(), -1e-3).cpu(), (v.float() * mask.float() * -1e-3))
    sy = torch.sigmoid(rnd(n, seed=12))
    dy, dyd, syd = rnd(n, seed=13), rnd(n, seed=14), rnd(n, seed=15)
    assert rel(ops.sigmoid_bwd_jvp(cu(dy), cu(dyd), cu(sy), cu(syd)), sim.sigmoid_bwd_jvp(dy, dyd, sy, syd)) < TOL
    x, xd = rnd(3, 250, seed=16), rnd(3, 250, seed=17)
    nrm, d = sim.l2norm_fwd_bwd(x)
    nd, dd = sim.l2norm_jvp(xd, nrm, d)
    nc, dc = ops.l2norm_fwd_bwd(cu(x))
    ndc, ddc = ops.l2norm_jvp(cu(xd), nc, dc)
    assert rel(ndc, nd) < TOL and rel(ddc, dd) < TOL


def test_dual_decoder_layer_matches_simulation(ops, sim):
    """One DETR decoder layer, forward + backward on dual numbers with per-episode theta tangents:
    CUDA kernels vs the float64 simulation of the same orchestration."""
    from interactron_b200 import layers
    from interactron_b200.dual import Dual, DualOps, DualWeights
    from interactron_b200.layers import GradSink
    from interactron_b200.params import ParamPack
    E, Fe, Lq, Lk, D, nh = 2, 2, 50, 361, 256, 8
    shapes = {"self_attn.in_proj_weight": (3 * D, D), "self_attn.in_proj_bias": (3 * D,),
              "self_attn.out_proj.weight": (D, D), "self_attn.out_proj.bias": (D,),
              "multihead_attn.in_proj_weight": (3 * D, D), "multihead_attn.in_proj_bias": (3 * D,),
              "multihead_attn.out_proj.weight": (D, D), "multihead_attn.out_proj.bias": (D,),
              "linear1.weight": (2048, D), "linear1.bias": (2048,), "linear2.weight": (D, 2048), "linear2.bias": (D,),
              "norm1.weight": (D,), "norm1.bias": (D,), "norm2.weight": (D,), "norm2.bias": (D,),
              "norm3.weight": (D,), "norm3.bias": (D,)}
    psi_names = [n for n in shapes if "in_proj" in n]
    th_names = [n for n in shapes if "in_proj" not in n]
    mk = lambda names, seed: [(("l." + n), rnd(*shapes[n], seed=seed + i, scale=0.06) + (1.0 if "norm" in n and "weight" in n else 0.0))
                              for i, n in enumerate(names)]
    th_items, ps_items = mk(th_names, 100), mk(psi_names, 200)
    tp, pp = ParamPack(th_items), ParamPack(ps_items)
    theta = tp.pack([t for _, t in th_items], dtype=torch.float64).unsqueeze(0)
    psi = pp.pack([t for _, t in ps_items], dtype=torch.float64).unsqueeze(0)
    v = rnd(E, tp.numel, seed=7, scale=0.02)
    B, Q, R = E * Fe, Fe * Lq, Fe * Lk
    tgt, tgt_d = rnd(E, Q, D, seed=8), rnd(E, Q, D, seed=9)
    qpos = rnd(1, Lq, D, seed=10)
    memp, memp_d = rnd(1, E * R, D, seed=11), rnd(1, E * R, D, seed=12)
    mem, mem_d = rnd(1, E * R, D, seed=13), rnd(1, E * R, D, seed=14)
    dt, dt_d = rnd(E * Q, D, seed=15), rnd(E * Q, D, seed=16)
    kmask = torch.zeros(B, Lk, dtype=torch.uint8)
    kmask[1, 300:] = 1

    def run(base, conv):
        dops = DualOps(base)
        th, ps, vv = conv(theta), conv(psi), conv(v)
        th_t = tp.transpose_into(base, th, base.zeros(1, tp.numel))
        v_t = tp.transpose_into(base, vv, base.zeros(E, tp.numel))
        ps_t = pp.transpose_into(base, ps, base.zeros(1, pp.numel))
        W = DualWeights((tp, th, th_t, vv, v_t), (pp, ps, ps_t, None, None))
        dm = layers.DecDims(E, B, Lq, Lk, D, nh)
        t_in = Dual(conv(tgt), conv(tgt_d))
        out, out_r, cache = layers.decoder_layer_fwd(dops, W, "l.", dm, t_in, t_in, Dual(conv(qpos)),
                                                     Dual(conv(memp), conv(memp_d)), Dual(conv(mem), conv(mem_d)),
                                                     conv(kmask))
        gpsi = dops.zeros(1, pp.numel)
        dqpos, dmp, dmem = dops.zeros(E, Q, D), dops.zeros(1, E * R, D), dops.zeros(1, E * R, D)
        din = layers.decoder_layer_bwd(dops, W, "l.", dm, cache, Dual(conv(dt), conv(dt_d)),
                                       GradSink(dops, pp, gpsi, shared=True), dqpos, dmp, dmem)
        return out, din, gpsi, dmp, dmem, dqpos

    ref = run(sim, lambda t: t.clone())
    got = run(ops, lambda t: (t.float() if t.dtype == torch.float64 else t).cuda().contiguous())
    for name, a, b in zip(("out", "din", "gpsi", "dmp", "dmem", "dqpos"), got, ref):
        # forward primal: tight.  Everything downstream of the ReLU mask (backward, tangents): one ReLU unit whose pre-activation is within fp32 round-off of 0 switches the tangent of that
        # unit on/off: a single such flip among the 4e5 hidden units is already 1.5e-3 relative L2
        assert rel(a.p, b.p) < (1e-4 if name == "out" else 5e-3), name + ".p"
        assert rel(a.t, b.t) < 5e-3, name + ".t"
