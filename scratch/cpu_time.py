import sys, time, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import helpers as H
from oracle import oracle as O
from neuralode4j_b200 import workloads as W
print("threads", O.num_threads())
for b in (16, 128):
    w = W.mnist_block(batch=b)
    eps = w.epsilon(w.shape)
    t0 = time.time(); o = H.run_oracle(w, eps); t1 = time.time()
    print(b, "oracle fwd+bwd s", t1 - t0, o["fwd_stats"], o["bwd_stats"])
