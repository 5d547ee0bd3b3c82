"""Host-side logic of the multi-GPU path, exercised with gloo on CPU (world_size 2): member
sharding, the LHS slice every rank draws, and the single all-gather of summary outputs.  The
compute itself needs a GPU (tests/test_gpu_parity.py); here each rank substitutes the CPU oracle
for its engine so the gathered result can be checked end to end."""
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np
import torch
import torch.distributed as dist
from hector_b200.sharding import shard_range, gather_summary
from bench import lhs
from oracle import port
from tests import util

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
M = 6
lo, hi = shard_range(M, rank, world)
X = lhs(M)[lo:hi]
raw = util.scenarios()["ssp245"]
# [2 vars][years][members of this rank]  (layout of the engine's output block)
block = np.empty((2, 555, hi - lo))
for i, x in enumerate(X):
    st, _, out, _, _ = port.run_member(raw, S=x[0], q10_rh=x[1], beta=x[2], diff=x[3])
    block[0, :, i] = out[0]
    block[1, :, i] = out[1]
full = gather_summary(torch.from_numpy(block), M, world)
if rank == 0:
    np.save(%(out)r, full.numpy())
dist.destroy_process_group()
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_ranges_cover_members():
    from hector_b200.sharding import shard_range
    for M in (1, 7, 8, 65536, 262144):
        for world in (1, 2, 3, 8):
            r = [shard_range(M, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == M
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gather_matches_single_process(tmp_path):
    from bench import lhs
    from oracle import port
    from tests import util
    out = str(tmp_path / "gathered.npy")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "out": out})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    procs = []
    for rank in range(2):
        e = dict(env, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=e))
    for p in procs:
        assert p.wait(timeout=300) == 0
    full = np.load(out)
    assert full.shape == (2, 555, 6)
    raw = util.scenarios()["ssp245"]
    X = lhs(6)
    for i in (0, 2, 3, 5):
        st, _, o, _, _ = port.run_member(raw, S=X[i, 0], q10_rh=X[i, 1], beta=X[i, 2], diff=X[i, 3])
        assert np.array_equal(full[0, :, i], o[0]) and np.array_equal(full[1, :, i], o[1])


def test_scenario_sorted_shards_unpermute():
    """BASELINE.json configs[3] partition: members interleaved over 8 scenarios in
    API order are sorted by scenario and cut into contiguous per-rank ranges; the gathered
    [rank][column] block is un-permuted back to API order through the inverse of the sort."""
    from hector_b200.sharding import scenario_sorted_shards
    M = 8 * 24
    ms = np.arange(M) % 8
    payload = np.arange(M) * 10.0 + ms            # something that identifies the API member
    for world in (1, 2, 4, 8):
        order, bounds = scenario_sorted_shards(ms, world)
        assert sorted(order.tolist()) == list(range(M))
        per = M // world
        gathered = np.empty((world, per))
        for r, (lo, hi) in enumerate(bounds):
            mine = order[lo:hi]
            assert hi - lo == per
            assert (np.diff(ms[mine]) >= 0).all()              # scenario-sorted inside the shard
            assert len(set(ms[mine].tolist())) == max(1, 8 // world)
            gathered[r] = payload[mine]
        inv = np.argsort(order)
        back = gathered[inv // per, inv % per]
        assert np.array_equal(back, payload)
