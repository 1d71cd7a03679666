"""CPU: the oracle's layer math against an INDEPENDENT implementation (PyTorch fp64 CPU ops + autograd).

The reference's own tests do not pin the numerics of conv3x3 / BatchNorm(train) / dense / activations inside an ODE block
("parity unpinned": smoke tests only, DL4J 1.0.0-beta3 sources absent), so the oracle DEFINES them
(oracle/node_oracle.cpp).  These tests tie that definition to the textbook operators a second code base computes:

* inner-graph forward f(z) and the vector-Jacobian products g^T df/dz, g^T df/dtheta (the pieces of
  BackpropagateAdjoint.java:87-143) against torch.nn.functional + autograd, including the parameter layouts (conv W
  [Cout,Cin,3,3] c-order; dense W f-order [nIn,nOut] then bias, MultiStepAdjointTest.java:216; BatchNorm gamma, beta, mean, var with
  mean/var outside the real gradients, GradientViewSelectionFromBlacklisted.java:33-37);
* the whole forward solve + adjoint backward of a conv/BatchNorm block against backpropagation through an independent
  fixed-step RK4 integration of the same ODE in torch (continuous adjoint == discrete gradient up to the solver tolerance).

torch is used as a checker only; nothing under neuralode4j_b200/ imports it for compute.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import helpers as H
from neuralode4j_b200 import workloads as W
from oracle import oracle as O

torch.set_num_threads(4)

_ACT = {"identity": lambda x: x, "relu": torch.relu, "elu": F.elu, "tanh": torch.tanh}


def _specs(w):
    """Layer descriptions with nIn = 0 ("take it from the incoming activations") resolved."""
    feat, out = w.shape[1], []
    for l in w.conf.layers:
        s = dict(l.spec())
        if s["kind"] in ("dense", "conv3x3"):
            s["n_in"] = s["n_in"] or feat
            feat = s["n_out"]
        elif s["kind"] == "batchnorm":
            s["n_in"] = s["n_out"] = s["n_out"] or feat
        out.append(s)
    return out


def _split_params(layers, params):
    """Flat DL4J-ordered parameter vector -> per-layer torch leaf tensors (+ the slices that are real gradients, in order)."""
    p = torch.tensor(np.asarray(params, np.float64))
    out, real, off = [], [], 0
    for s in layers:
        if s["kind"] == "dense":
            nin, nout = s["n_in"], s["n_out"]
            w = p[off:off + nin * nout].clone().requires_grad_(True)          # f-order [nIn,nOut]: W(i,o) = w[i + o*nIn]
            off += nin * nout
            real.append(w)
            b = None
            if s["has_bias"]:
                b = p[off:off + nout].clone().requires_grad_(True)
                off += nout
                real.append(b)
            out.append((w, b))
        elif s["kind"] == "conv3x3":
            n = s["n_out"] * s["n_in"] * 9
            w = p[off:off + n].clone().requires_grad_(True)                   # [Cout,Cin,3,3] c-order
            off += n
            real.append(w)
            out.append((w,))
        elif s["kind"] == "batchnorm":
            c = s["n_out"]
            g = p[off:off + c].clone().requires_grad_(True)
            b = p[off + c:off + 2 * c].clone().requires_grad_(True)
            off += 4 * c                                                      # gamma, beta, running mean, running var
            real += [g, b]
            out.append((g, b))
        else:
            out.append(())
    assert off == p.numel()
    return out, real


def _f(layers, tensors, z):
    """The inner graph, training mode (SingleStep.java:37): BatchNorm uses batch mean and biased batch variance."""
    x = z
    for s, t in zip(layers, tensors):
        if s["kind"] == "dense":
            w, b = t
            x = x @ w.reshape(s["n_out"], s["n_in"]).T
            if b is not None:
                x = x + b
        elif s["kind"] == "conv3x3":
            x = F.conv2d(x, t[0].reshape(s["n_out"], s["n_in"], 3, 3), padding=1)
        elif s["kind"] == "batchnorm":
            x = F.batch_norm(x, None, None, t[0], t[1], training=True, eps=s["eps"])
        x = _ACT[s["activation"]](x)
    return x


def _check_rhs_vjp(w, seed, rtol):
    rng = np.random.default_rng(seed)
    params = w.params.astype(np.float64) + 0.1 * rng.standard_normal(w.params.size)
    z = w.y0.astype(np.float64)
    g = rng.standard_normal(z.shape)
    orc = H.oracle_for(w, np.float64)
    fz, dz, dth = orc.rhs_vjp(params, z, g)
    tensors, real = _split_params(_specs(w), params)
    zt = torch.tensor(z, requires_grad=True)
    ft = _f(_specs(w), tensors, zt)
    grads = torch.autograd.grad((ft * torch.tensor(g)).sum(), [zt] + real)
    np.testing.assert_allclose(fz, ft.detach().numpy(), rtol=rtol, atol=rtol)
    np.testing.assert_allclose(dz, grads[0].numpy(), rtol=rtol, atol=rtol)
    tg = np.concatenate([x.numpy().ravel() for x in grads[1:]])
    assert tg.size == orc.n_real == dth.size
    np.testing.assert_allclose(dth, tg, rtol=rtol, atol=rtol * max(1.0, np.abs(tg).max()))


@pytest.mark.parametrize("act", ["relu", "tanh", "elu"])
def test_conv_batchnorm_block_matches_torch_autograd(act):
    w = W.mnist_block(batch=5, channels=8, hw=5)
    for l in w.conf.layers:
        if l.kind == 3:
            l.activation = act
    _check_rhs_vjp(w, 3, 1e-9)


def test_dense_blocks_match_torch_autograd():
    _check_rhs_vjp(W.spiral_block(batch=7), 4, 1e-10)                        # fc 4-20-20-4, ELU, ELU, identity, biases
    _check_rhs_vjp(W.mlp_block(batch=6, dim=24, act="tanh"), 5, 1e-10)
    _check_rhs_vjp(W.timeseries_block(batch=4, dim=12, nt=3), 6, 1e-10)


def test_adjoint_of_conv_batchnorm_block_matches_backprop_through_rk4():
    """Forward solve + adjoint backward (SingleStepAdjoint.java:37-88) vs autograd through 200 RK4 steps of the same ODE."""
    w = W.mnist_block(batch=4, channels=4, hw=4, abs_tol=1e-10, rel_tol=1e-10)
    for l in w.conf.layers:
        if l.kind == 3:
            l.activation = "tanh"                                            # smooth: the two gradients converge to each other
    rng = np.random.default_rng(9)
    params = w.params.astype(np.float64) + 0.05 * rng.standard_normal(w.params.size)
    z0 = w.y0.astype(np.float64)
    eps = rng.standard_normal(z0.shape)
    orc = H.oracle_for(w, np.float64)
    t = np.array([0.0, 1.0])
    z1 = orc.forward(params, z0, t)
    dy0, grad, _ = orc.backward(params, z1, eps, t)

    tensors, real = _split_params(_specs(w), params)
    zt = torch.tensor(z0, requires_grad=True)
    x, n = zt, 200
    h = 1.0 / n
    for _ in range(n):
        k1 = _f(_specs(w), tensors, x)
        k2 = _f(_specs(w), tensors, x + 0.5 * h * k1)
        k3 = _f(_specs(w), tensors, x + 0.5 * h * k2)
        k4 = _f(_specs(w), tensors, x + h * k3)
        x = x + (h / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
    grads = torch.autograd.grad((x * torch.tensor(eps)).sum(), [zt] + real)
    np.testing.assert_allclose(z1, x.detach().numpy(), rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(dy0, grads[0].numpy(), rtol=1e-6, atol=1e-7)
    # oracle gradient view: DL4J flat order with zeros in the BatchNorm mean / var slots
    C = 4
    flat = []
    it = iter(grads[1:])
    for l in w.conf.layers:
        if l.kind == 3:
            flat += [next(it).numpy(), next(it).numpy(), np.zeros(2 * C)]
        else:
            flat += [next(it).numpy().ravel()]
    np.testing.assert_allclose(grad, np.concatenate(flat), rtol=1e-6, atol=1e-7)
