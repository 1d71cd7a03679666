import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import helpers as H
from neuralode4j_b200 import workloads as W, _lib
w = W.mnist_block(batch=16, channels=64)
for flags in (0, _lib.FLAG_NO_GRAPH):
    r = H.run_cuda(w, flags=flags)
    o = H.run_oracle(w, r["eps"])
    print("flags", flags, "fwd", r["fwd_stats"], o["fwd_stats"], "bwd", r["bwd_stats"], o["bwd_stats"])
    print(" fwd log gpu h", r["fwd_log"]["h"], "err", r["fwd_log"]["err"])
    print(" fwd log orc h", o["fwd_log"].h, "err", o["fwd_log"].err)
    print(" bwd log gpu h", r["bwd_log"]["h"], "err", r["bwd_log"]["err"], "t", r["bwd_log"]["t"])
    print(" bwd log orc h", o["bwd_log"].h, "err", o["bwd_log"].err, "t", o["bwd_log"].t)
    for k in ("out", "dy0", "grad"):
        d = np.abs(r[k] - o[k]); print(" ", k, "max", np.abs(o[k]).max(), "maxdiff", d.max(), "rel-l2", np.linalg.norm(r[k]-o[k])/np.linalg.norm(o[k]))
    print(" launches", r["fwd_launches"], r["bwd_launches"], "ms", r["fwd_ms"], r["bwd_ms"])
