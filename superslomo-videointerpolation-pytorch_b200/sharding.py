"""Partition of the (frame pair, timestep) work items over ranks.

Every output pixel of every (pair, timestep) depends only on that pair's two frames and its
flows, so the path shards with no data-path collective.  All N timesteps of a pair stay on one
rank whenever there are at least as many pairs as ranks (stage 1 then runs once per pair and
both kernels get their timestep batching); otherwise the timesteps of each pair are split.
"""
import ctypes
import os


def _block(n_items, n_parts, part):
    """[start, end) of block `part` when n_items are split into n_parts nearly equal blocks."""
    base, rem = divmod(n_items, n_parts)
    start = part * base + min(part, rem)
    return start, start + base + (1 if part < rem else 0)


def shard_work(n_pairs, n_timesteps, rank, world_size):
    """Work of `rank`: a list of (pair, t_start, t_end) with t_end exclusive; deterministic and
    exhaustive over ranks, no overlaps."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank %r / world_size %r" % (rank, world_size))
    if n_pairs >= world_size:
        p0, p1 = _block(n_pairs, world_size, rank)
        return [(p, 0, n_timesteps) for p in range(p0, p1)]
    # fewer pairs than ranks: flatten (pair, timestep) and block-distribute
    lo, hi = _block(n_pairs * n_timesteps, world_size, rank)
    work = []
    for p in range(n_pairs):
        a, b = max(lo, p * n_timesteps), min(hi, (p + 1) * n_timesteps)
        if a < b:
            work.append((p, a - p * n_timesteps, b - p * n_timesteps))
    return work


def frames_of(work):
    return sum(t1 - t0 for _, t0, t1 in work)


# ---------------------------------------------------------------------------------------------
# Host placement for the host-buffer entry point (ssm_synthesize_host is PCIe-bound: pinned buffers
# that live on the other socket's memory cross the inter-socket link on every copy).
def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


class numa_local_to_gpu:
    """Context manager: while active, the calling thread runs on the CPUs of the NUMA node the GPU's
    PCIe root hangs off and prefers that node's memory, so that pinned host buffers allocated inside
    land next to the GPU.  One rank per GPU: each rank binds to its own GPU's node.  Best effort --
    without sysfs topology, or when the process's CPU set does not reach that node, it does nothing.
    `info` says what was done."""

    def __init__(self, device_index, sysfs="/sys/bus/pci/devices"):
        self.device_index, self.sysfs = device_index, sysfs
        self.info = {"numa_node": None, "cpus_bound": 0, "mempolicy": False}
        self._affinity = None

    def _pci_dir(self):
        import torch
        p = torch.cuda.get_device_properties(self.device_index)
        return os.path.join(self.sysfs, "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id))

    @staticmethod
    def _set_mempolicy(mode, node):
        libc = ctypes.CDLL(None, use_errno=True)
        if node is None:
            return libc.syscall(238, 0, None, 0) == 0                 # MPOL_DEFAULT
        mask = (ctypes.c_ulong * 16)()
        mask[node // 64] = 1 << (node % 64)
        return libc.syscall(238, mode, mask, 16 * 64 + 1) == 0        # x86-64 set_mempolicy

    def __enter__(self):
        try:
            d = self._pci_dir()
            with open(os.path.join(d, "numa_node")) as f:
                node = int(f.read())
            with open(os.path.join(d, "local_cpulist")) as f:
                local = _parse_cpulist(f.read())
        except Exception:
            return self
        if node < 0:
            return self
        self.info["numa_node"] = node
        try:
            current = os.sched_getaffinity(0)
            both = current & local
            if both and both != current:
                os.sched_setaffinity(0, both)
                self._affinity = current
            self.info["cpus_bound"] = len(both)
        except Exception:
            pass
        try:
            self.info["mempolicy"] = bool(self._set_mempolicy(1, node))  # MPOL_PREFERRED
        except Exception:
            pass
        return self

    def __exit__(self, *exc):
        if self.info["mempolicy"]:
            try:
                self._set_mempolicy(0, None)
            except Exception:
                pass
        if self._affinity is not None:
            try:
                os.sched_setaffinity(0, self._affinity)
            except Exception:
                pass
        return False
