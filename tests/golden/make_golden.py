#!/usr/bin/env python
"""Generates tests/golden/*.npz|json.

  reference_vectors.json   the golden vectors the REFERENCE's own tests hold for this path, transcribed with file:line
                           (they are also what tests/helpers.py pins the oracle against)
  oracle_vectors.npz       outputs of the CPU oracle (fp32) on seeded small workloads: forward states, dL/dy0, dL/dt, a
                           fixed sample of dL/dtheta, and the solver's step counts.  The GPU parity tests compare the CUDA path
                           against THESE committed numbers as well as against the oracle built on the box, and a CPU test checks
                           that the oracle still reproduces them.

Run from the repository root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers as H                                    # noqa: E402
from neuralode4j_b200 import workloads as W            # noqa: E402


def cases():
    return {
        "mnist_b4_hw5": W.mnist_block(batch=4, channels=64, hw=5),
        "mnist_c8_b3_hw4": W.mnist_block(batch=3, channels=8, hw=4),
        "spiral_b8_nt5": W.spiral_block(batch=8, nt=5),
        "anode_time_input_b16_nt3": W.anode_time_block(batch=16, time_input=True, nt=3),
        "anode_time_fixed_b16": W.anode_time_block(batch=16),
        "conv_time_b2_c8_hw4_k3": W.conv_time_block(batch=2, channels=8, hw=4, time_channels=3),
    }


def grad_sample_index(n):
    return np.unique(np.linspace(0, n - 1, min(n, 512)).astype(np.int64))


def main():
    out = {}
    for name, w in cases().items():
        interp = w.conf.odeForwardConf.interpolateForwardIfMultiStep
        o0 = H.oracle_for(w).forward(w.params, w.y0, w.time, interpolate=interp)
        eps = w.epsilon(o0.shape)
        o = H.run_oracle(w, eps)
        out[name + "/out"] = o["out"].astype(np.float32)
        out[name + "/dy0"] = np.asarray(o["dy0"], np.float32)
        g = np.asarray(o["grad"], np.float32)
        out[name + "/grad_sample"] = g[grad_sample_index(g.size)]
        out[name + "/grad_sum_abs"] = np.array([np.abs(g.astype(np.float64)).sum()])
        if o["dt"] is not None:
            out[name + "/dt"] = np.asarray(o["dt"], np.float32)
        out[name + "/counts"] = np.array(list(o["fwd_stats"]) + list(o["bwd_stats"]), np.int64)
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **out)
    ref = {
        "_source": "MultiStepAdjointTest.java:174-296 (torchdiffeq), tolerances 1e-3 abs (ys, dL/dy0) and 1e-4 rel (dL/dtheta, dL/dt)",
        "ys_fwd": H.GOLDEN_YS_FWD.tolist(), "dy0_fwd": H.GOLDEN_DY0_FWD.tolist(), "dtheta_fwd": H.GOLDEN_DTHETA_FWD.tolist(),
        "dt_fwd": H.GOLDEN_DT_FWD.tolist(),
        "ys_bwd": H.GOLDEN_YS_BWD.tolist(), "dy0_bwd": H.GOLDEN_DY0_BWD.tolist(), "dtheta_bwd": H.GOLDEN_DTHETA_BWD.tolist(),
        "dt_bwd": H.GOLDEN_DT_BWD.tolist(),
    }
    with open(os.path.join(HERE, "reference_vectors.json"), "w") as f:
        json.dump(ref, f, indent=1)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
