"""Multi-GPU plumbing: ensembles shard by member (no exchange during the run); the single
exchange of the job is an all-gather of the per-member summary outputs.

  gather_summary   torch.distributed all-gather (NCCL on GPUs, gloo in the CPU tests)
  PushExchange     the same result over peer memory, one kernel launch per rank: a rank's finished
                   16-year slabs are copied into every peer's gather block by the copy engines
                   while its kernel computes the later slabs (hx_xchg_* / hx_run_exchange)
  PeerExchange     the round-1 form: pulls of whole run segments: every rank opens its peers' output blocks
                   through CUDA IPC and pulls finished run segments with copy-engine transfers
                   over NVLink while its own kernel computes the next segment.  A collective
                   kernel cannot do that: it finds no room next to the persistent run kernel.
"""
import os
import subprocess

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_members, rank, world):
    """contiguous, balanced member range [lo, hi) of `rank`"""
    base, extra = divmod(n_members, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_summary(block, n_members, world):
    """block: [..., members_of_this_rank] tensor (last dim = members, possibly padded beyond the
    rank's share).  Returns [..., n_members] on every rank with members in global order."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_range(n_members, rank, world)
    mine = block[..., : hi - lo].contiguous()
    if world == 1:
        return mine
    width = shard_range(n_members, 0, world)[1]  # the largest share
    pad = torch.zeros(mine.shape[:-1] + (width,), dtype=mine.dtype, device=mine.device)
    pad[..., : hi - lo] = mine
    flat = torch.empty(world * pad.numel(), dtype=pad.dtype, device=pad.device)
    dist.all_gather_into_tensor(flat, pad.reshape(-1))
    out = flat.view((world,) + tuple(pad.shape))
    parts = []
    for k in range(world):
        a, b = shard_range(n_members, k, world)
        parts.append(out[k][..., : b - a])
    return torch.cat(parts, dim=-1)


class PeerExchange:
    """All ranks (one engine each, same outputs / years / padded member count) end up with every
    rank's trajectories of `variables`: self.blocks[v] is a device tensor [world, n_years,
    member_stride], rank k's block at index k.

        ex = PeerExchange(ens, ["CO2_concentration", "global_tas"], gloo_group)
        ens.reset(); ex.run()          # run to the end date, exchanging as segments finish

    Ordering across ranks is done on the host: a rank pulls segment k only after its own event
    for k has completed and a barrier on `group` (a gloo group: an NCCL barrier is a kernel and
    would queue behind the run kernel) says everybody's has."""

    def __init__(self, ens, variables, group, segments=4):
        import numpy as np
        self.ens, self.variables, self.group = ens, list(variables), group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        # every rank must reach both collectives even if its own CUDA IPC call fails
        try:
            mine, err = ens.ipc_export(), None
        except Exception as ex:
            mine, err = None, ex
        handles = [None] * self.world
        dist.all_gather_object(handles, mine, group=group)
        if err is None and all(h is not None for h in handles):
            try:
                ens.ipc_open(handles, self.rank)
            except Exception as ex:
                err = ex
        elif err is None:
            err = RuntimeError("a peer could not export its output block")
        oks = [None] * self.world
        dist.all_gather_object(oks, err is None, group=group)
        if not all(oks):
            raise err if err is not None else RuntimeError("a peer could not open the blocks")
        _, stride, ny = ens.output_device(self.variables[0])
        self.n_years, self.stride = ny, stride
        self.blocks = {v: torch.empty((self.world, ny, stride), dtype=torch.float64,
                                      device="cuda") for v in self.variables}
        first = int(ens.start_year) + 1
        nslab = (ny + 15) // 16
        segments = max(1, min(int(segments), nslab))
        cuts = sorted({min(ny, ((nslab * (k + 1)) // segments) * 16) for k in range(segments)}
                      | {ny})
        cuts = [c for c in cuts if c > 0]
        self.segments = [(first + a, first + b - 1) for a, b in zip([0] + cuts[:-1], cuts)]
        assert len(self.segments) <= 16
        del np

    def run(self):
        ens = self.ens
        for k, (ya, yb) in enumerate(self.segments):   # every segment is queued right away
            ens.run(yb)
            ens.event_record(k)
        for k, (ya, yb) in enumerate(self.segments):
            ens.event_synchronize(k)                   # my segment k is done ...
            dist.barrier(group=self.group)             # ... and so is everybody's
            for v in self.variables:
                ens.ipc_pull(v, ya, yb, self.blocks[v].data_ptr())
        ens.ipc_wait()
        # nobody may overwrite its outputs (the next reset / run) while a peer still reads them
        dist.barrier(group=self.group)
        return self.blocks


class _DevView:
    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f8", "data": (ptr, False),
                                         "version": 2, "strides": None}


class PushExchange:
    """Every rank (one engine each, same outputs / years / padded member count) ends up with every
    rank's recorded outputs: self.block is a device tensor [world, n_outputs, n_years, stride]
    (outputs in the engine's selection order), rank k's at index k.

        ex = PushExchange(ens, group)
        ens.reset(); block = ex.run()

    ONE launch of the persistent run kernel per rank; each finished slab is pushed to all peers
    over NVLink while later slabs compute; a host barrier on `group` (gloo: an NCCL barrier is a
    kernel and would queue behind work on the device) completes the exchange."""

    def __init__(self, ens, group):
        self.ens, self.group = ens, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        try:
            mine, err = ens.xchg_create(self.world, self.rank), None
        except Exception as ex:
            mine, err = None, ex
        handles = [None] * self.world
        dist.all_gather_object(handles, mine, group=group)
        if err is None and all(h is not None for h in handles):
            try:
                ens.xchg_open(handles)
            except Exception as ex:
                err = ex
        elif err is None:
            err = RuntimeError("a peer could not create its gather block")
        oks = [None] * self.world
        dist.all_gather_object(oks, err is None, group=group)
        if not all(oks):
            raise err if err is not None else RuntimeError("a peer could not open the blocks")
        _, stride, ny = ens.output_device(ens.outputs[0])
        ptr, per_rank = ens.xchg_block()
        n_out = per_rank // (ny * stride)
        self.n_years, self.stride, self.n_out = ny, stride, n_out
        self.block = torch.as_tensor(_DevView(ptr, (self.world, n_out, ny, stride)), device="cuda")

    def run(self, to_date=-1):
        dist.barrier(group=self.group)      # nobody still reads the blocks this run overwrites
        self.ens.run_exchange(to_date)
        dist.barrier(group=self.group)      # everybody's pushes have landed
        return self.block


def scenario_sorted_shards(member_scenario, world):
    """members are sorted by scenario (stable: API order inside a
    scenario) and the sorted list is cut into `world` contiguous balanced ranges, so that a GPU
    holds few scenarios' tables and whole tiles of one scenario.
    -> (order, bounds): order[p] = API index of the member at sorted position p; bounds[r] =
    [lo, hi) of rank r's sorted positions.  inverse: np.argsort(order)."""
    ms = np.asarray(member_scenario)
    order = np.argsort(ms, kind="stable")
    return order, [shard_range(len(order), r, world) for r in range(world)]


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_numa(device_index):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE it allocates
    pinned host buffers (first touch then places them on that node): eight ranks streaming
    582 MB each per step into NUMA-remote memory is what made the 8-GPU end-to-end step 37 %
    slower than the 1-GPU one in round 1.  -> dict describing what was done (for the bench line)."""
    try:
        bus = None
        try:
            p = torch.cuda.get_device_properties(device_index)
            bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        except Exception:
            out = subprocess.run(["nvidia-smi", "-i", str(device_index), "--query-gpu=pci.bus_id",
                                  "--format=csv,noheader"], capture_output=True, text=True).stdout
            bus = out.strip().lower()
            if len(bus.split(":")[0]) == 8:      # nvidia-smi prints an 8-digit domain
                bus = bus[4:]
        base = "/sys/bus/pci/devices/" + bus
        node = int(open(base + "/numa_node").read())
        cpus = _parse_cpulist(open(base + "/local_cpulist").read())
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if node < 0 or not use or use == allowed:
            return {"numa_node": node, "bound": False, "cpus": len(allowed)}
        os.sched_setaffinity(0, use)
        return {"numa_node": node, "bound": True, "cpus": len(use)}
    except Exception as ex:  # binding is an optimisation, never a requirement
        return {"bound": False, "error": repr(ex)}
