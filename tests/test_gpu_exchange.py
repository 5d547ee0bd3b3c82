"""The multi-GPU exchange over peer memory (hx_ipc_*, hector_b200.sharding.PeerExchange), with two
processes sharing whatever GPUs the box has (both on cuda:0 when there is only one: CUDA IPC works
between processes on the same device too).  Each rank runs its own members; afterwards every rank
must hold both ranks' trajectories, equal to what each rank fetched for itself."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np
import torch
import torch.distributed as dist
import hector_b200 as hb
from hector_b200.sharding import PeerExchange, PushExchange, shard_range
from tests import util

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
dev = rank %% torch.cuda.device_count()
torch.cuda.set_device(dev)
M = 300                                    # per rank; pads to 384 on the device
X = util.lhs(M * world, seed=3)[rank * M:(rank + 1) * M]
variables = ["CO2_concentration", "global_tas"]
ens = hb.Ensemble(M, util.scenarios()["ssp245"], device=dev, outputs=variables)
for j, n in enumerate(["S", "q10_rh", "beta", "diff"]):
    ens.setvar(n, X[:, j])
ens.prepare()
MODE = %(mode)r
if MODE == "pull":
    ex = PeerExchange(ens, variables, dist.group.WORLD, segments=3)
    for rep in range(2):                       # twice: buffers and events are reused
        ens.reset()
        blocks = ex.run()
else:
    ex = PushExchange(ens, dist.group.WORLD)
    for rep in range(2):
        ens.reset()
        blk = ex.run()                         # [world, n_out, years, stride]
    torch.cuda.synchronize()
    blocks = {v: blk[:, k] for k, v in enumerate(variables)}
years = np.arange(1746, 2301, dtype=np.float64)
mine = {v: ens.fetch(v, years) for v in variables}           # [M, years]
out = {}
for v in variables:
    b = blocks[v].cpu().numpy()                               # [world, years, stride]
    assert b.shape[0] == world and b.shape[1] == 555
    assert np.array_equal(b[rank, :, :M].T, mine[v])
    out[v] = b[:, :, :M]
np.savez(%(out)r %% rank, **out, **{"mine_" + v: mine[v] for v in variables})
dist.barrier()
ens.close()
dist.destroy_process_group()
"""


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("mode", ["pull", "push"])
def test_two_process_peer_exchange(tmp_path, mode):
    """pull: PeerExchange (run segments pulled by the receivers); push: PushExchange (one launch,
    every finished slab written into the peers' gather blocks: hx_xchg_* / hx_run_exchange)"""
    out = str(tmp_path / "rank%d.npz")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "out": out, "mode": mode})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    procs = []
    for rank in range(2):
        e = dict(env, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=e))
    for p in procs:
        assert p.wait(timeout=600) == 0
    r0, r1 = np.load(out % 0), np.load(out % 1)
    for v in ("CO2_concentration", "global_tas"):
        # what rank 0 pulled from rank 1 is what rank 1 computed, and vice versa
        assert np.array_equal(r0[v][1].T, r1["mine_" + v])
        assert np.array_equal(r1[v][0].T, r0["mine_" + v])
        assert np.array_equal(r0[v], r1[v])


def test_push_exchange_from_cpp():
    """the same exchange without Python: tests/cpp/test_xchg.cpp forks two processes that use only
    the C ABI (hx_xchg_create / hx_xchg_open / hx_run_exchange / hx_xchg_block) and a socketpair"""
    sys.path.insert(0, os.path.join(ROOT, "tests", "cpp"))
    import build_compat
    exe = os.path.join(build_compat.OUT, "xchg_cpp")
    if not os.path.exists(exe):
        try:
            exe = build_compat.build_xchg()
        except Exception as ex:   # no compiler / CUDA headers on this box
            pytest.skip("cannot build tests/cpp/test_xchg.cpp here: %r" % (ex,))
    ini = None
    for d in ("/root/reference/inst/input", os.path.join(ROOT, "oracle", "_ref", "input")):
        if os.path.exists(os.path.join(d, "hector_ssp245.ini")):
            ini = os.path.join(d, "hector_ssp245.ini")
    if ini is None:
        pytest.skip("reference input data not available")
    r = subprocess.run([exe, ini], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "XCHG_OK" in r.stdout, r.stdout + r.stderr
