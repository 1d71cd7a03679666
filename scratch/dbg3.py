import sys, numpy as np, ctypes
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import helpers as H
from oracle import oracle as O
from neuralode4j_b200 import workloads as W
L = O.lib(); L.orc_set_probe_noise.argtypes=[ctypes.c_double]
w = W.mnist_block(batch=16, channels=64)
eps = w.epsilon(w.shape)
L.orc_set_probe_noise(0.0)
a = H.run_oracle(w, eps, np.float32)
for noise in (1e-7, 1e-6, 1e-5):
    L.orc_set_probe_noise(noise)
    b = H.run_oracle(w, eps, np.float32)
    print("noise", noise, b["bwd_stats"], "err", b["bwd_log"].err[-3:])
    for k in ("out", "dy0", "grad"):
        print("  ", k, "rel-l2:", np.linalg.norm(a[k]-b[k])/np.linalg.norm(a[k]), "maxdiff", np.abs(a[k]-b[k]).max())
