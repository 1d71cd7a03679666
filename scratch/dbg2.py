import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import helpers as H
from neuralode4j_b200 import workloads as W
for act in ("relu", "tanh"):
    w = W.mnist_block(batch=16, channels=64)
    for l in w.conf.layers:
        if l.kind == 3: l.activation = act
    eps = w.epsilon(w.shape)
    a = H.run_oracle(w, eps, np.float32)
    b = H.run_oracle(w, eps, np.float64)
    print(act, "fwd", a["fwd_stats"], b["fwd_stats"], "bwd", a["bwd_stats"], b["bwd_stats"])
    for k in ("out", "dy0", "grad"):
        print("  ", k, "rel-l2 f32 vs f64:", np.linalg.norm(a[k]-b[k])/np.linalg.norm(b[k]), "maxdiff", np.abs(a[k]-b[k]).max(), "max", np.abs(b[k]).max())
