"""Worker for the multi-GPU parity test: launched by torchrun with one process per GPU.

Global minibatch = world * local rows.  Every rank runs forward + adjoint backward on its row shard through the C ABI with
the sharded controller (node4j_comm_init); rank 0 gathers the shards and compares them with the CPU oracle evaluated on
the GLOBAL minibatch.  Exit code 0 = parity.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers as H                                    # noqa: E402
from neuralode4j_b200 import _lib, workloads as W      # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "mnist"
    flags = int(sys.argv[2]) if len(sys.argv) > 2 else _lib.FLAG_EXACT_CONTRACTIONS
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("gloo")
    local = 8
    default_mode = which == "mnist_default"      # the production kernels: tcgen05 convolutions with fused prologue / epilogue exchanges
    if default_mode:
        which, flags = "mnist", _lib.FLAG_FORCE_TENSOR
    if which == "mnist":
        wg = W.mnist_block(batch=local * world, channels=64)
    elif which == "mlp":
        wg = W.mlp_block(batch=local * world, dim=96)
    elif which == "anode_time":      # time-dependent f with time gradients: the summed time epsilon crosses the ranks as well
        wg = W.anode_time_block(batch=local * world, time_input=True, nt=3)
    else:
        wg = W.golden_linear(1)
        local = 1
    rows = slice(rank * local, (rank + 1) * local)
    shape = (local,) + wg.shape[1:]
    v = wg.conf.instantiate(shape, device=local_rank, flags=flags)
    v.init_comm()
    v.setParams(wg.params.copy())
    y0 = np.ascontiguousarray(wg.y0[rows])
    if wg.time_is_input:
        v.setInputs(y0, wg.time.copy())
    else:
        v.setInputs(y0)
    out = v.doForward(True)
    fs = (v.fwd_stats.nfe, v.fwd_stats.accepted, v.fwd_stats.rejected)
    eps_g = wg.epsilon(wg.shape if out.ndim == len(wg.shape) else (wg.shape[0],) + out.shape[1:]) if which != "golden" else W.golden_lossgrad()
    grad = np.zeros(v.numParams(), np.float32)
    v.setBackpropGradientsViewArray(grad)
    v.setEpsilon(np.ascontiguousarray(eps_g[rows]))
    g, epsilons = v.doBackward()
    bs = (v.bwd_stats.nfe, v.bwd_stats.accepted, v.bwd_stats.rejected)
    if wg.time_is_input:
        idx = wg.conf.odeBackwardConf.time_input_index
        dy0, dt = epsilons[(idx + 1) % 2], np.asarray(epsilons[idx]).ravel()
    else:
        dy0, dt = epsilons[0], None
    v.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, dict(out=out, dy0=dy0, grad=np.asarray(g), dt=dt, fs=fs, bs=bs))
    ok = True
    if rank == 0:
        o = H.run_oracle(wg, eps_g)
        outs = np.concatenate([r["out"] for r in gathered], axis=0)
        dy0s = np.concatenate([r["dy0"] for r in gathered], axis=0)
        try:
            for r in gathered:
                # atol 1e-12 / rtol 1e-6 (golden problem) puts the error estimate at fp32 rounding level: the accept/reject
                # sequence is then sensitive to the last bit of the summed parameter gradients, so counts are compared for
                # the tolerance-1e-3 configurations only
                if which != "golden":
                    assert r["fs"] == o["fwd_stats"] and r["bs"] == o["bwd_stats"], (r["fs"], o["fwd_stats"], r["bs"], o["bwd_stats"])
                assert r["fs"] == gathered[0]["fs"] and r["bs"] == gathered[0]["bs"], "ranks disagree on the step sequence"
                assert np.array_equal(r["grad"], gathered[0]["grad"]), "parameter gradients differ between ranks"
                if dt is not None:
                    assert np.array_equal(r["dt"], gathered[0]["dt"])
            H.assert_close(outs, o["out"], 1e-4, 1e-5 * np.abs(o["out"]).max(), "out")
            if default_mode:
                # fp32-accumulating contractions on a ReLU + BatchNorm block: gradients agree in relative L2 only
                for k, got in (("dy0", dy0s), ("grad", gathered[0]["grad"])):
                    rel = np.linalg.norm(got - o[k]) / np.linalg.norm(o[k])
                    assert rel < 1e-2, (k, rel)
            elif which == "mnist" and world > 2:
                # Parity mode keeps every contraction exact per rank, but the parameter-gradient partials of the ranks are summed in
                # fp32 (rank order, identical on all ranks): with more than two ranks that sum is no longer the once-rounded double sum
                # of the oracle, and on a BatchNorm+ReLU block such a last-bit difference is amplified by ReLU mask flips (DESIGN.md
                # section 2: the ORACLE moves by 5e-4 .. 5e-3 in relative L2 when its own contractions are perturbed by 1e-12 .. 1e-8).
                # Bounded against that yardstick: counts exact (above), forward state rtol 1e-4 (above), gradients within a small
                # multiple of the oracle's own sensitivity; the fraction of entries outside rtol 1e-4 is reported.
                import ctypes
                from oracle import oracle as O
                L = O.lib()
                L.orc_set_probe_noise.argtypes = [ctypes.c_double]
                own = {"dy0": 0.0, "grad": 0.0}
                try:
                    for noise in (1e-10, 1e-8, 1e-6):
                        L.orc_set_probe_noise(noise)
                        p = H.run_oracle(wg, eps_g)
                        for k in own:
                            own[k] = max(own[k], float(np.linalg.norm(p[k] - o[k]) / np.linalg.norm(o[k])))
                finally:
                    L.orc_set_probe_noise(0.0)
                for k, got in (("dy0", dy0s), ("grad", gathered[0]["grad"])):
                    rel = float(np.linalg.norm(got - o[k]) / np.linalg.norm(o[k]))
                    err = np.abs(got.astype(np.float64) - o[k])
                    frac = float((err > 1e-4 * np.abs(o[k]) + 1e-5 * np.abs(o[k]).max()).mean())
                    print(f"  {k}: rel-L2 {rel:.2e} vs the oracle (outside rtol 1e-4: {frac:.2e} of {err.size}); oracle's own sensitivity up to {own[k]:.2e}")
                    assert rel < 1e-2 and rel <= 5 * max(own[k], 1e-5), (k, rel, own[k])
            else:
                H.assert_close(dy0s, o["dy0"], 1e-4, 1e-5 * np.abs(o["dy0"]).max(), "dy0")
                H.assert_close(gathered[0]["grad"], o["grad"], 1e-4, 1e-5 * np.abs(o["grad"]).max(), "grad")
            if dt is not None:
                H.assert_close(gathered[0]["dt"], o["dt"], 1e-4, 1e-5 * np.abs(o["dt"]).max(), "dt")
            print(f"MGPU PARITY OK ({which}, world={world}): fwd {gathered[0]['fs']} bwd {gathered[0]['bs']}")
        except AssertionError as e:
            print("MGPU PARITY FAILED:", e)
            ok = False
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
