"""Shared helpers for the tests: building the oracle for a workload and running the CUDA path through the
reference-shaped Python API (which itself only calls the C ABI)."""
from __future__ import annotations

import numpy as np

from oracle import oracle as O
from neuralode4j_b200 import _lib
from neuralode4j_b200.workloads import Workload

GOLDEN_YS_FWD = np.array(
    [[[-1.23, 1.2753194351313588, 4.694932254437282, 9.36250433337755, 15.73345986285669, 24.429438907671752, 36.29894018833409, 52.50011823313958, 74.61376567400269, 104.79757535834692],
      [-0.040000000000000036, 2.4653194351313603, 5.884932254437281, 10.552504333377543, 16.92345986285671, 25.61943890767168, 37.488940188334105, 53.69011823313958, 75.8037656740027, 105.98757535834685]],
     [[1.15, 4.523878831433274, 9.129023844467975, 15.414778231911933, 23.994455416185012, 35.70520942482531, 51.689701681157864, 73.50760277718506, 103.28774415967294, 143.93586077027197],
      [2.34, 5.7138788314332745, 10.319023844467976, 16.604778231911936, 25.184455416185003, 36.89520942482521, 52.879701681157854, 74.69760277718505, 104.47774415967294, 145.12586077027197]]])
GOLDEN_DY0_FWD = np.array([[330.3413671858052, 350.8541876986256], [641.5288548171546, 667.1698804581804]])
GOLDEN_DTHETA_FWD = np.array([26399.28182908193, 28805.83177942673, 29982.48959443847, 32597.88284962647, 2022.3108828170273, 2197.8094583156044])
GOLDEN_DT_FWD = np.array([-6617.017543489908, 63.564528808863464, 185.79311916695133, 262.00021152296233, 369.0851188922797, 519.4357040639363,
                          730.3690674996153, 1026.0795986642863, 1440.3518309918456, 2020.3383638791681])
GOLDEN_YS_BWD = np.array(
    [[[-1.23, -3.0654783048539187, -4.410209436983333, -5.395402672421891, -6.117186729574334, -6.645989569834943, -7.033407758920375, -7.317242594011386, -7.525189149078903, -7.677538128370835],
      [-0.040000000000000036, -1.875478304853917, -3.2202094369833336, -4.205402672421891, -4.927186729574334, -5.4559895698349425, -5.843407758920375, -6.127242594011386, -6.335189149078852, -6.487538128370829]],
     [[1.15, -1.321813099544714, -3.132743808435675, -4.459489833436336, -5.4315063823619125, -6.143637811088716, -6.665368496900053, -7.047604920850001, -7.327643653785053, -7.532809904849017],
      [2.34, -0.13181309954471376, -1.9427438084356758, -3.2694898334363294, -4.241506382361924, -4.953637811088717, -5.475368496900053, -5.857604920850002, -6.137643653785053, -6.342809904849003]]])
GOLDEN_DY0_BWD = np.array([[0.4791608727137171, 20.99198138553419], [22.890949753489018, 48.53197539451459]])
GOLDEN_DTHETA_BWD = np.array([902.9542328576254, 696.559656699922, 1683.4912664748028, 1268.25338547379, -173.44082030059053, -348.93939579916616])
GOLDEN_DT_BWD = np.array([-229.96656912353825, 34.11827941396485, 53.52715859322072, 40.515245326331545, 30.634856397062812, 23.14160607429379,
                          17.465314677923114, 13.17005445759693, 9.923109065345606, 7.470945117798939])


def oracle_for(w: Workload, dtype=np.float32, exact_stage_times=False) -> O.Oracle:
    c = w.conf.odeForwardConf.solver.config
    cfg = O.solver_cfg(c.absTol, c.relTol, c.minStep, c.maxStep, exact_stage_times)
    return O.Oracle(w.conf.graph_spec(), w.shape, cfg, dtype=dtype)


def time_grad_mode(w: Workload) -> int:
    h = w.conf.odeBackwardConf
    if h.time_input_index is None:
        return O.TG_NONE
    return O.TG_CALC if h.needTimeGradient else O.TG_ZERO


def run_cuda(w: Workload, flags: int = 0, eps=None, device_tensors: bool = False):
    """Forward + backward through the reference-shaped API.  Returns dict(out, dy0, grad, dt, fwd_stats, bwd_stats, logs)."""
    v = w.conf.instantiate(w.shape, device=0, flags=flags)
    try:
        params, y0, t = w.params.copy(), w.y0.copy(), w.time.copy()
        if device_tensors:
            import torch
            params, y0 = torch.from_numpy(params).cuda(), torch.from_numpy(y0).cuda()
        v.setParams(params)
        if w.time_is_input:
            v.setInputs(y0, t)
        else:
            v.setInputs(y0)
        out = v.doForward(True)
        fwd_log = v.step_log()
        fs = (v.fwd_stats.nfe, v.fwd_stats.accepted, v.fwd_stats.rejected)
        fwd_launches = v.fwd_stats.gpu_launches
        out_np = out.cpu().numpy() if device_tensors else out
        e = w.epsilon(out_np.shape) if eps is None else np.ascontiguousarray(eps, dtype=np.float32)
        grad = np.zeros(v.numParams(), dtype=np.float32)
        if device_tensors:
            import torch
            grad = torch.from_numpy(grad).cuda()
            e_in = torch.from_numpy(e).cuda()
        else:
            e_in = e
        v.setBackpropGradientsViewArray(grad)
        v.setEpsilon(e_in)
        g, epsilons = v.doBackward()
        bwd_log = v.step_log()
        bs = (v.bwd_stats.nfe, v.bwd_stats.accepted, v.bwd_stats.rejected)
        if w.time_is_input:
            idx = w.conf.odeBackwardConf.time_input_index
            dt = np.asarray(epsilons[idx]).ravel()
            dy0 = epsilons[(idx + 1) % 2]
        else:
            dt, dy0 = None, epsilons[0]
        if device_tensors:
            g, dy0 = g.cpu().numpy(), dy0.cpu().numpy()
        return dict(out=out_np, dy0=np.asarray(dy0), grad=np.asarray(g), dt=dt, fwd_stats=fs, bwd_stats=bs, eps=e,
                    fwd_log=fwd_log, bwd_log=bwd_log, fwd_launches=fwd_launches, bwd_launches=v.bwd_stats.gpu_launches,
                    fwd_ms=v.fwd_stats.gpu_ms, bwd_ms=v.bwd_stats.gpu_ms)
    finally:
        v.close()


def run_oracle(w: Workload, eps, dtype=np.float32):
    orc = oracle_for(w, dtype)
    interp = w.conf.odeForwardConf.interpolateForwardIfMultiStep
    out = orc.forward(w.params, w.y0, w.time, interpolate=interp)
    fs = (orc.stats.nfe, orc.stats.accepted, orc.stats.rejected)
    fwd_log = orc.step_log()
    dy0, grad, dt = orc.backward(w.params, out, eps, w.time, time_grad=time_grad_mode(w))
    bs = (orc.stats.nfe, orc.stats.accepted, orc.stats.rejected)
    return dict(out=out, dy0=dy0, grad=grad, dt=dt, fwd_stats=fs, bwd_stats=bs, fwd_log=fwd_log, bwd_log=orc.step_log())


def assert_close(a, b, rtol, atol, what=""):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    bad = err > tol
    assert not bad.any(), f"{what}: {bad.sum()} of {a.size} outside rtol={rtol} atol={atol}; worst abs {err.max():.3e}, worst rel {(err / (np.abs(b) + 1e-30)).max():.3e}"
