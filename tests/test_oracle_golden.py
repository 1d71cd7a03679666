"""Pins of the CPU oracle (oracle/node_oracle.cpp) against everything the reference's own tests hold for this path.
Runs without a GPU.

  MultiStepAdjointTest.java:174-296   torchdiffeq golden vectors (ys, dL/dy0, dL/dtheta, dL/dt), both time directions
  InterpolationTest.java:20-49        torchdiffeq golden tensor for the quartic dense output
  DormandPrince54MseTest.java:22-62   error norm vs commons-math3 estimateError (published formula restated below)
  AdaptiveRungeKuttaStepPolicyTest.java:26-142   initial step and step filter vs commons-math3 (restated below)
  DormandPrince54SolverTest.java:51-95           CircleODE: solution, number of accepted steps and step times vs
                                                 commons-math3 DormandPrince54Integrator (restated below), double precision
  SingleStepTest.java:30-46, MultiStepTest.java:28-47, InputStepTest.java:27-68   exp(kt) closed forms
  InterpolatingMultiStepSolverTest.java:26-37, SingleSteppingMultiStepSolverTest.java:25-39   equivalences

commons-math3 3.6.1 is a dependency of the reference (pom.xml:18) that is not vendored; `_commons_dp54` restates its
EmbeddedRungeKuttaIntegrator.integrate / DormandPrince54Integrator.estimateError / AdaptiveStepsizeIntegrator.initializeStep
algorithm (the textbook Hairer-Norsett-Wanner controller) so that the equal-step-count pin can be evaluated here.
"""
import ctypes
import math

import numpy as np
import pytest

from oracle import oracle as O
from neuralode4j_b200 import workloads as W

import helpers as H


# ---------------------------------------------------------------------------------------------------------------------
# commons-math3 restatement (test-side, independent of the oracle's loop structure)
# ---------------------------------------------------------------------------------------------------------------------
def _commons_dp54(f, t0, t1, y0, min_step, max_step, abs_tol, rel_tol):
    A, B, E, Cc, _ = O.tableau()
    safety, min_red, max_growth, order = 0.9, 0.2, 10.0, 5
    exp = -1.0 / order
    y = np.array(y0, dtype=np.float64)
    n = y.size
    forward = t1 > t0
    K = np.zeros((7, n))
    step_start = t0
    first = True
    times = []
    h_new = 0.0
    nfe = 0

    def filter_step(h, fwd, accept_small):
        if abs(h) < min_step:
            h = min_step if fwd else -min_step   # acceptSmall path not needed for these problems
        if h > max_step:
            h = max_step
        elif h < -max_step:
            h = -max_step
        return h

    def estimate_error(K, y0_, y1_, h):
        err_sum = E @ K
        tol = abs_tol + rel_tol * np.maximum(np.abs(y0_), np.abs(y1_))
        ratio = h * err_sum / tol
        return math.sqrt(np.sum(ratio * ratio) / n)

    def initialize_step(fwd, scale, t, y_, yd0):
        r = y_ / scale
        y_on = float(np.sum(r * r))
        r = yd0 / scale
        yd_on = float(np.sum(r * r))
        h = 1e-6 if (y_on < 1e-10 or yd_on < 1e-10) else 0.01 * math.sqrt(y_on / yd_on)
        if not fwd:
            h = -h
        y1_ = y_ + h * yd0
        yd1 = f(t + h, y1_)
        r = (yd1 - yd0) / scale
        ydd_on = math.sqrt(float(np.sum(r * r))) / h
        max_inv2 = max(math.sqrt(yd_on), ydd_on)
        h1 = max(1e-6, 0.001 * abs(h)) if max_inv2 < 1e-15 else (0.01 / max_inv2) ** (1.0 / order)
        h = min(100.0 * abs(h), h1)
        h = max(h, 1e-12 * abs(t))
        h = max(h, min_step)
        h = min(h, max_step)
        if not fwd:
            h = -h
        return h, yd1

    last = False
    while not last:
        error = 10.0
        while error >= 1.0:
            if first:
                K[0] = f(step_start, y); nfe += 1
                scale = abs_tol + rel_tol * np.abs(y)
                h_new, K[1] = initialize_step(forward, scale, step_start, y, K[0]); nfe += 1
                first = False
            step = h_new
            if forward:
                if step_start + step >= t1:
                    step = t1 - step_start
            else:
                if step_start + step <= t1:
                    step = t1 - step_start
            for k in range(1, 7):
                ytmp = y + step * (A[k - 1, :k] @ K[:k])
                K[k] = f(step_start + Cc[k - 1] * step, ytmp); nfe += 1
            ytmp = y + step * (B @ K)
            error = estimate_error(K, y, ytmp, step)
            if error >= 1.0:
                factor = min(max_growth, max(min_red, safety * error ** exp))
                h_new = filter_step(step * factor, forward, False)
        step_start = step_start + step
        y = ytmp
        times.append(step_start)
        K[0] = K[6]
        last = (step_start >= t1) if forward else (step_start <= t1)
        if not last:
            factor = min(max_growth, max(min_red, safety * error ** exp))
            scaled = step * factor
            next_t = step_start + scaled
            next_last = (next_t >= t1) if forward else (next_t <= t1)
            h_new = filter_step(scaled, forward, next_last)
            filt_t = step_start + h_new
            if (filt_t >= t1) if forward else (filt_t <= t1):
                h_new = t1 - step_start
    return y, times, nfe


def _circle_oracle(omega, c, cfg, dtype=np.float64):
    # CircleODE (T/ode/solve/CircleODE.java:29-38): yDot = omega * [c1 - y1, y0 - c0]  ==  y*W + b as a dense layer
    layers = [dict(kind="dense", n_in=2, n_out=2, activation="identity", has_bias=True)]
    # W(i,o) f-order: W(1,0) = -omega, W(0,1) = +omega
    params = np.array([0.0, -omega, omega, 0.0, omega * c[1], -omega * c[0]])
    return O.Oracle(layers, (1, 2), cfg, dtype=dtype), params


# ---------------------------------------------------------------------------------------------------------------------
# golden vectors
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("sign,ys,dy0,dth,dt", [
    (1, H.GOLDEN_YS_FWD, H.GOLDEN_DY0_FWD, H.GOLDEN_DTHETA_FWD, H.GOLDEN_DT_FWD),
    (-1, H.GOLDEN_YS_BWD, H.GOLDEN_DY0_BWD, H.GOLDEN_DTHETA_BWD, H.GOLDEN_DT_BWD)])
def test_multistep_adjoint_golden(dtype, sign, ys, dy0, dth, dt):
    w = W.golden_linear(sign)
    o = H.run_oracle(w, W.golden_lossgrad(), dtype)
    assert np.abs(o["out"] - ys).max() < 1e-3                     # MultiStepAdjointTest.java:199-203
    assert np.abs(o["dy0"] - dy0).max() < 1e-3                    # :214-218
    assert np.abs(o["grad"] / dth - 1).max() < 1e-4               # :223-227 (flat f-order view)
    assert np.abs(o["dt"] / dt - 1).max() < 1e-4                  # :233-236
    if dtype == np.float64:
        assert np.abs(o["dy0"] - dy0).max() < 1e-6 * np.abs(dy0).max() + 2e-5
        assert np.abs(o["dt"] / dt - 1).max() < 5e-6
        # regression values of the step sequence
        if sign == 1:
            assert o["fwd_stats"] == (93, 15, 0)


def test_interpolation_golden():
    # InterpolationTest.java:20-49
    n = 30
    y0 = np.linspace(-10, 10, n); ymid = np.linspace(-7, 13, n); y1 = np.linspace(-12, 7, n)
    f0 = np.linspace(2, 5, n); f1 = np.linspace(-3, -1, n)
    expected = np.array([-10.8218, -10.1690, -9.5162, -8.8633, -8.2105, -7.5577, -6.9049, -6.2521, -5.5993, -4.9465,
                         -4.2936, -3.6408, -2.9880, -2.3352, -1.6824, -1.0296, -0.3768, 0.2760, 0.9288, 1.5817,
                         2.2345, 2.8873, 3.5401, 4.1929, 4.8457, 5.4985, 6.1513, 6.8042, 7.4570, 8.1098])
    for dtype in (np.float32, np.float64):
        out = O.interpolate(y0, y1, ymid, f0, f1, 1.23, -2, 3, 2.34, dtype)
        assert np.abs(out - expected).max() < 6e-5                # printed to 4 decimals in the reference


# ---------------------------------------------------------------------------------------------------------------------
# solver internals vs the commons-math3 algorithm
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("h", [1.23, -1.23])
def test_error_norm_vs_commons(h):
    # DormandPrince54MseTest.java:36-61
    dim = 100
    K = np.linspace(-10, 10, 7 * dim).reshape(7, dim)
    y0 = np.linspace(-2.34, 3.45, dim)
    y1 = np.linspace(-3.45, 4.56, dim)[::-1]
    atol, rtol = 2.34e-3, 3.45e-5
    _, _, E, _, _ = O.tableau()
    ratio = h * (E @ K) / (atol + rtol * np.maximum(np.abs(y0), np.abs(y1)))
    expected = math.sqrt(np.sum(ratio ** 2) / dim)
    assert abs(O.error_norm(K, y0, y1, h, atol, rtol, np.float64) - expected) < 1e-9 * expected
    assert abs(O.error_norm(K, y0, y1, h, atol, rtol, np.float32) - expected) < 1e-4   # the reference's tolerance


def test_initial_step_vs_commons():
    # AdaptiveRungeKuttaStepPolicyTest.java:26-63: CircleODE(c=[1.23,4.56], omega=0.666), t=[-1.2,7.4], y0=[3,5]
    atol, rtol = 1.23e-3, 2.34e-5
    c, omega = (1.23, 4.56), 0.666
    cfg = O.solver_cfg(atol, rtol, 1e-20, 1e20)
    orc, params = _circle_oracle(omega, c, cfg)
    y0 = np.array([[3.0, 5.0]])
    h, k0, k1 = orc.init_step(params, y0, -1.2, 7.4)

    f = lambda t, y: omega * np.array([c[1] - y[1], y[0] - c[0]])
    yd0 = f(-1.2, y0[0])
    scale = atol + rtol * np.abs(y0[0])
    # commons initializeStep
    r = y0[0] / scale; y_on = np.sum(r * r); r = yd0 / scale; yd_on = np.sum(r * r)
    h0 = 0.01 * math.sqrt(y_on / yd_on)
    yd1 = f(-1.2 + h0, y0[0] + h0 * yd0)
    r = (yd1 - yd0) / scale
    ydd = math.sqrt(np.sum(r * r)) / h0
    h1 = (0.01 / max(math.sqrt(yd_on), ydd)) ** 0.2
    expected = max(min(100 * h0, h1), 1e-12 * 1.2)
    assert abs(h - expected) < 1e-6
    np.testing.assert_allclose(k0[0], yd0, atol=1e-6)
    np.testing.assert_allclose(k1[0], yd1, atol=1e-6)


@pytest.mark.parametrize("step", [1.23, -1.23, 12345, -12345, 0.000001234, -0.000001234])
def test_step_filter_vs_commons(step):
    # AdaptiveRungeKuttaStepPolicyTest.java:69-142: error 0.666, bounds [1e-2, 1e2]
    err = 0.666
    cfg = O.solver_cfg(1e-20, 1e-20, 1e-2, 1e2)
    factor = min(10.0, max(0.2, 0.9 * err ** (-1.0 / 5)))
    h = step * factor
    expected = math.copysign(min(1e2, max(1e-2, abs(h))), step)
    assert abs(O.next_step(step, err, cfg, np.float64) - expected) < 1e-6 * abs(expected)
    assert abs(O.next_step(step, err, cfg, np.float32) - expected) < 1e-6 * max(1.0, abs(expected))


@pytest.mark.parametrize("ts", [(-0.023, 0.0456), (0.023, -0.0456)])
def test_circle_vs_commons_steps_and_closed_form(ts):
    # DormandPrince54SolverTest.java:51-95 (double precision, tolerances 1e-10)
    c, omega = (1.23, 4.56), 20.666
    cfg = O.solver_cfg(1e-10, 1e-10, 1e-10, 100)
    orc, params = _circle_oracle(omega, c, cfg)
    y0 = np.array([[3.0, -5.0]])
    y = orc.forward(params, y0, np.array(ts))
    log = orc.step_log()
    f = lambda t, yy: omega * np.array([c[1] - yy[1], yy[0] - c[0]])
    y_ref, times_ref, _ = _commons_dp54(f, ts[0], ts[1], y0[0], 1e-10, 100, 1e-10, 1e-10)
    np.testing.assert_allclose(y[0], y_ref, rtol=1e-9)
    acc_t = log.t[log.accepted == 1]
    assert len(acc_t) == len(times_ref)                                  # "Incorrect number of steps"
    np.testing.assert_allclose(acc_t, times_ref, atol=1e-8)              # "Incorrect time at step"
    th = omega * (ts[1] - ts[0])
    d = y0[0] - np.array(c)
    exact = np.array(c) + np.array([[math.cos(th), -math.sin(th)], [math.sin(th), math.cos(th)]]) @ d
    np.testing.assert_allclose(y[0], exact, atol=1e-7)
    assert orc.stats.nfe == 3 + 6 * (orc.stats.accepted + orc.stats.rejected)


# ---------------------------------------------------------------------------------------------------------------------
# closed forms and equivalences (vertex-level tests of the reference)
# ---------------------------------------------------------------------------------------------------------------------
def _exp_oracle(k, n, dtype=np.float32, cfg=None):
    layers = [dict(kind="dense", n_in=n, n_out=n, activation="identity", has_bias=False)]
    params = (np.eye(n) * k).ravel(order="F")
    return O.Oracle(layers, (1, n), cfg or O.solver_cfg(1e-3, 1e-3, 1e-10, 100), dtype=dtype), params


def test_single_step_exponential():
    # SingleStepTest.java:30-46: dy/dt = k*y  ->  y0*exp(k*t), tolerance 1e-3
    orc, p = _exp_oracle(2.0, 3, cfg=O.solver_cfg(1e-6, 1e-6, 1e-10, 100))
    y0 = np.array([[-3.0, 5.0, 7.0]])
    y = orc.forward(p, y0, np.array([0.0, 1.0]))
    np.testing.assert_allclose(y, y0 * math.exp(2.0), rtol=1e-3)


@pytest.mark.parametrize("interp", [True, False])
def test_multi_step_exponential(interp):
    # MultiStepTest.java:28-47 / InputStepTest.java:27-68
    orc, p = _exp_oracle(1.3, 3, cfg=O.solver_cfg(1e-6, 1e-6, 1e-10, 100))
    y0 = np.array([[1.0, -2.0, 3.0]])
    t = np.linspace(0, 2, 7)
    ys = orc.forward(p, y0, t, interpolate=interp)             # [B, H, T]
    assert ys.shape == (1, 3, 7)
    np.testing.assert_allclose(ys, y0[:, :, None] * np.exp(1.3 * t)[None, None, :], rtol=1e-3)


def test_interpolating_equals_single_stepping():
    # InterpolatingMultiStepSolverTest.java:26-37 (1e-4)
    w = W.golden_linear(1)
    orc = H.oracle_for(w, np.float64)
    a = orc.forward(w.params, w.y0, w.time, interpolate=True)
    b = orc.forward(w.params, w.y0, w.time, interpolate=False)
    np.testing.assert_allclose(a, b, rtol=1e-4)
    # SingleSteppingMultiStepSolverTest.java:25-39: multi-step == one-shot at the last time
    c = orc.forward(w.params, w.y0, w.time[[0, -1]])
    np.testing.assert_allclose(b[:, :, -1], c, rtol=1e-5)


def test_conv_layout_rank5_output():
    w = W.mnist_block(batch=2, channels=4, hw=3)
    orc = H.oracle_for(w)
    ys = orc.forward(w.params, w.y0, np.array([0.0, 0.5, 1.0]), interpolate=False)
    assert ys.shape == (2, 3, 4, 3, 3)                          # [B,T,C,H,W] (MultiStep.java:55)
    np.testing.assert_array_equal(ys[:, 0], w.y0)


def test_rhs_vjp_matches_finite_differences():
    # the conv3x3 + BatchNorm(train) + ReLU numerics are unpinned by the reference; pin the oracle's vjp to its own forward
    w = W.mnist_block(batch=3, channels=4, hw=4)
    for l in w.conf.layers:
        if l.kind == 3:
            l.activation = "tanh"                               # smooth, so central differences are meaningful
    orc = H.oracle_for(w, np.float64)
    rng = np.random.default_rng(0)
    params = w.params.astype(np.float64) + 0.1 * rng.standard_normal(w.params.size)
    z = w.y0.astype(np.float64)
    g = rng.standard_normal(z.shape)
    _, dz, dth = orc.rhs_vjp(params, z, g)
    eps = 1e-6
    for idx in [(0, 0, 0, 0), (1, 2, 3, 1), (2, 3, 1, 2)]:
        zp, zm = z.copy(), z.copy(); zp[idx] += eps; zm[idx] -= eps
        fd = np.sum(g * (orc.rhs(params, zp) - orc.rhs(params, zm))) / (2 * eps)
        assert abs(fd - dz[idx]) < 1e-6 * max(1.0, abs(fd))
    # parameter gradients: conv weight, gamma, beta (real-gradient order: bn1[g,b], conv1, bn2[g,b], conv2, bn3[g,b])
    C = 4
    real_to_flat = np.concatenate([np.arange(0, 2 * C), 4 * C + np.arange(C * C * 9), 4 * C + C * C * 9 + np.arange(2 * C)])
    for r_idx in (1, 2 * C + 5, 2 * C + C * C * 9 + 3):
        f_idx = real_to_flat[r_idx]
        pp, pm = params.copy(), params.copy(); pp[f_idx] += eps; pm[f_idx] -= eps
        fd = np.sum(g * (orc.rhs(pp, z) - orc.rhs(pm, z))) / (2 * eps)
        assert abs(fd - dth[r_idx]) < 1e-6 * max(1.0, abs(fd))


def test_relu_bn_adjoint_sensitivity():
    """Documents why gradient parity of the BatchNorm+ReLU block is tested in parity mode: perturbing the ORACLE's own
    contractions by 1e-6 (relative) moves its gradients by ~1e-3, while 1e-7 leaves them at ~1e-7."""
    L = O.lib()
    L.orc_set_probe_noise.argtypes = [ctypes.c_double]
    w = W.mnist_block(batch=16, channels=64)
    eps = w.epsilon(w.shape)
    try:
        L.orc_set_probe_noise(0.0)
        a = H.run_oracle(w, eps)
        L.orc_set_probe_noise(1e-6)
        b = H.run_oracle(w, eps)
    finally:
        L.orc_set_probe_noise(0.0)
    rel = lambda k: np.linalg.norm(a[k] - b[k]) / np.linalg.norm(a[k])
    assert rel("out") < 1e-5
    assert 1e-5 < rel("dy0") < 5e-2


def test_time_must_be_sorted():
    w = W.golden_linear(1)
    orc = H.oracle_for(w)
    t = np.array([1.0, 3.0, 2.0])
    ys = orc.forward(w.params, w.y0, t)
    with pytest.raises(ValueError):                             # MultiStepAdjoint.java:41-46
        orc.backward(w.params, ys, np.zeros_like(ys), t, time_grad=O.TG_CALC)
    with pytest.raises(ValueError):                             # SolverConfig.java:22-24
        O.Oracle(w.conf.graph_spec(), w.shape, O.solver_cfg(1e-3, 1e-3, 1.0, 0.5))


# ---------------------------------------------------------------------------------------------------------------------
# time as input to the inner graph: TimeInput + ShapeMatchVertex(MergeVertex)
# ---------------------------------------------------------------------------------------------------------------------
def test_time_input_forward_pass_reference_case():
    """ForwardPassTest.calculateDerivativeWithTimeDependency (ForwardPassTest.java:60-91): all-ones weights, so every output is
    sum(input) + t."""
    layers = [dict(kind="time_concat", n_out=1), dict(kind="dense", n_out=5, has_bias=False, activation="identity")]
    o = O.Oracle(layers, (1, 5), dtype=np.float64)
    inp = np.arange(5, dtype=np.float64).reshape(1, 5)
    fz = o.rhs(np.ones(o.n_params), inp, t=7.89)
    np.testing.assert_allclose(fz, np.full((1, 5), inp.sum() + 7.89), rtol=0, atol=1e-12)


@pytest.mark.parametrize("case", ["dense", "conv"])
def test_time_input_vjp_matches_finite_differences(case):
    """g^T df/dt (DuplicateScalarToShape.backprop = sum of the time part of the merged epsilon) against central differences, fp64."""
    rng = np.random.default_rng(0)
    if case == "dense":
        layers = [dict(kind="time_concat", n_out=2), dict(kind="dense", n_out=7, activation="tanh"), dict(kind="dense", n_out=4)]
        shape = (3, 4)
    else:
        layers = [dict(kind="time_concat", n_out=3), dict(kind="conv3x3", n_out=5, has_bias=False, activation="tanh"),
                  dict(kind="batchnorm", activation="relu"), dict(kind="conv3x3", n_out=4, has_bias=False)]
        shape = (2, 4, 5, 5)
    o = O.Oracle(layers, shape, dtype=np.float64)
    th = rng.standard_normal(o.n_params) * 0.4
    z, g = rng.standard_normal(shape), rng.standard_normal(shape)
    _, dz, _ = o.rhs_vjp(th, z, g, t=0.3)
    dt = o.last_dt()
    e = 1e-6
    fd = (np.sum(g * o.rhs(th, z, t=0.3 + e)) - np.sum(g * o.rhs(th, z, t=0.3 - e))) / (2 * e)
    assert abs(dt - fd) < 1e-6 * max(1.0, abs(fd))
    idx = tuple(s // 2 for s in shape)
    zp, zm = z.copy(), z.copy()
    zp[idx] += e
    zm[idx] -= e
    fdz = (np.sum(g * o.rhs(th, zp, t=0.3)) - np.sum(g * o.rhs(th, zm, t=0.3))) / (2 * e)
    assert abs(dz[idx] - fdz) < 1e-6 * max(1.0, abs(fdz))


def test_time_input_adjoint_time_gradients():
    """With a time-dependent f the adjoint's time slot integrates -a^T df/dt (TimeInput.update).  Single step: both entries of
    dL/dt equal central differences of the loss.  Multi step: interior and last entries do; the FIRST entry is the reference's
    `lastTime` accumulator, -sum(dL/dt[1:]) (CalcMultiStepTimeGrad.java:58-63 writes lastTime, not the integrated slot), which is
    only the true derivative for autonomous f — restated as it is."""
    rng = np.random.default_rng(1)
    layers = [dict(kind="time_concat", n_out=1), dict(kind="dense", n_out=6, activation="tanh"), dict(kind="dense", n_out=3)]
    o = O.Oracle(layers, (2, 3), O.solver_cfg(1e-11, 1e-11, 1e-20, 100), dtype=np.float64)
    th = rng.standard_normal(o.n_params) * 0.7
    y0, w = rng.standard_normal((2, 3)), rng.standard_normal((2, 3))
    e = 1e-5

    t = np.array([0.2, 1.1])
    out = o.forward(th, y0, t, interpolate=False)
    _, _, dt = o.backward(th, out, w, t, time_grad=O.TG_CALC)
    loss = lambda t0, t1: float(np.sum(w * o.forward(th, y0, np.array([t0, t1]), interpolate=False)))
    fd0 = (loss(t[0] + e, t[1]) - loss(t[0] - e, t[1])) / (2 * e)
    fd1 = (loss(t[0], t[1] + e) - loss(t[0], t[1] - e)) / (2 * e)
    np.testing.assert_allclose(dt, [fd0, fd1], rtol=2e-5)

    tt = np.linspace(0.1, 1.3, 5)
    out = o.forward(th, y0, tt, interpolate=True)
    wm = rng.standard_normal(out.shape)
    _, _, dtm = o.backward(th, out, wm, tt, time_grad=O.TG_CALC)
    lm = lambda ts: float(np.sum(wm * o.forward(th, y0, ts, interpolate=True)))
    for k in range(1, 5):
        a, b = tt.copy(), tt.copy()
        a[k] += e
        b[k] -= e
        assert abs(dtm[k] - (lm(a) - lm(b)) / (2 * e)) < 3e-4 * max(1.0, abs(dtm[k]))   # dense-output (4th order) error included
    assert abs(dtm[0] + dtm[1:].sum()) < 1e-9 * np.abs(dtm).sum()


def test_stage_time_quirk_is_observable_only_with_time_input():
    """Quirk A.3 (every stage evaluated at t+h, FirstOrderEquationWithState.java:45,82): irrelevant for autonomous f, visible with
    a time input — and both variants still converge to the same solution as the tolerance shrinks."""
    w = W.anode_time_block(batch=8)
    a = H.oracle_for(w).forward(w.params, w.y0, w.time, interpolate=True)
    b = H.oracle_for(w, exact_stage_times=True).forward(w.params, w.y0, w.time, interpolate=True)
    assert np.abs(a - b).max() > 1e-5
    s = W.spiral_block(batch=8, nt=4)
    a2 = H.oracle_for(s).forward(s.params, s.y0, s.time, interpolate=True)
    b2 = H.oracle_for(s, exact_stage_times=True).forward(s.params, s.y0, s.time, interpolate=True)
    np.testing.assert_array_equal(a2, b2)
    wt = W.anode_time_block(batch=8, abs_tol=1e-9, rel_tol=1e-9)
    c = H.oracle_for(wt, np.float64).forward(wt.params, wt.y0, wt.time, interpolate=True)
    d = H.oracle_for(wt, np.float64, exact_stage_times=True).forward(wt.params, wt.y0, wt.time, interpolate=True)
    assert np.abs(c - d).max() < 1e-3          # the quirk makes the scheme first-order consistent in t only: slow convergence


def test_time_concat_must_be_first():
    with pytest.raises(Exception):
        O.Oracle([dict(kind="dense", n_out=4), dict(kind="time_concat", n_out=1), dict(kind="dense", n_out=4)], (2, 4))


# ---------------------------------------------------------------------------------------------------------------------
# committed fixtures (tests/golden/)
# ---------------------------------------------------------------------------------------------------------------------
def _golden():
    import importlib.util
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, np.load(os.path.join(here, "oracle_vectors.npz")), os.path.join(here, "reference_vectors.json")


def test_reference_vector_fixture_is_what_the_oracle_is_pinned_against():
    import json
    _, _, path = _golden()
    ref = json.load(open(path))
    np.testing.assert_array_equal(np.array(ref["ys_fwd"]), H.GOLDEN_YS_FWD)
    np.testing.assert_array_equal(np.array(ref["dtheta_bwd"]), H.GOLDEN_DTHETA_BWD)
    np.testing.assert_array_equal(np.array(ref["dt_fwd"]), H.GOLDEN_DT_FWD)


def test_oracle_reproduces_committed_fixtures():
    """tests/golden/oracle_vectors.npz was written by tests/golden/make_golden.py from this oracle; a change of its arithmetic shows
    up here (same compiler flags, no contraction: reproducible to the last bits; 1e-6 leaves room for a different libm)."""
    mod, fx, _ = _golden()
    for name, w in mod.cases().items():
        interp = w.conf.odeForwardConf.interpolateForwardIfMultiStep
        out = H.oracle_for(w).forward(w.params, w.y0, w.time, interpolate=interp)
        o = H.run_oracle(w, w.epsilon(out.shape))
        assert list(o["fwd_stats"]) + list(o["bwd_stats"]) == fx[name + "/counts"].tolist(), name
        np.testing.assert_allclose(o["out"], fx[name + "/out"], rtol=1e-6, atol=1e-7, err_msg=name)
        np.testing.assert_allclose(o["dy0"], fx[name + "/dy0"], rtol=1e-5, atol=1e-7 * np.abs(fx[name + "/dy0"]).max(), err_msg=name)
        g = np.asarray(o["grad"], np.float32)
        gs = fx[name + "/grad_sample"]
        np.testing.assert_allclose(g[mod.grad_sample_index(g.size)], gs, rtol=1e-5, atol=1e-6 * np.abs(gs).max(), err_msg=name)
