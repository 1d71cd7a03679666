"""Multi-GPU (row-sharded minibatch) tests.

* test_sharded_parity_*: needs >= 2 GPUs on one box; spawns one process per GPU with torchrun (tests/mgpu_worker.py) and checks
  the sharded run against the CPU oracle on the global minibatch.
* test_comm_plumbing_gloo: CPU, world_size 2 over gloo — the host-side handle exchange and row sharding used by
  OdeVertexImpl.init_comm, with the native calls stubbed (the C ABI itself needs a GPU).
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        from neuralode4j_b200 import _lib
        return _lib.load().node4j_device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["mnist", "mnist_default", "mlp", "golden", "anode_time"])
def test_sharded_parity(which):
    if _ngpu() < 2:
        pytest.skip("needs two GPUs on one box")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "mgpu_worker.py"), which]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0 and "MGPU PARITY OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from neuralode4j_b200.vertex import OdeVertexImpl

    calls = {}

    class FakeLib:
        def node4j_comm_export(self, h, buf):
            C.memmove(buf, bytes([rank + 1]) * 64, 64)
            return 0

        def node4j_comm_init(self, h, r, w, allh, global_batch):
            calls["init"] = (r, w, bytes(allh.raw), global_batch)
            return 0

    v = OdeVertexImpl.__new__(OdeVertexImpl)
    v._L, v._h, v.shape = FakeLib(), C.c_void_p(1), (16, 4)
    out = v.init_comm()
    v._h = None
    q.put((rank, out, calls["init"]))
    dist.destroy_process_group()


def test_comm_plumbing_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, 29533, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out, (r, w, handles, gb) in res:
        assert out == (rank, 2) and (r, w) == (rank, 2)
        assert handles == bytes([1]) * 64 + bytes([2]) * 64          # rank order
        assert gb == 32                                              # global batch = world x local rows
