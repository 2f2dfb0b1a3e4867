"""Multi-GPU host logic: one process per GPU (torchrun), replay sharded by actor stream, no data-path collective.

The reference keeps one replay shard per env instance and trains on them round-robin in one process
(franQ/Replay/__init__.py:13-16, franQ/Agent/deepQlearning.py:106).  Here actor stream i belongs to rank i % world_size;
each rank samples, relabels and builds targets from its own HBM arena, and only gradients are all-reduced (Learner)."""
import os

import torch


def shards_of_rank(num_instances, rank, world_size):
    """actor streams owned by `rank`: i with i % world_size == rank"""
    return [i for i in range(int(num_instances)) if i % int(world_size) == int(rank)]


def init_from_env(backend=None):
    """torchrun contract: RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT from the environment."""
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if world > 1 and not dist.is_initialized():
        use_cuda = torch.cuda.is_available()
        if use_cuda:
            torch.cuda.set_device(local)
        dist.init_process_group(backend or ("nccl" if use_cuda else "gloo"),
                                **({"device_id": torch.device("cuda", local)} if use_cuda else {}))
    return rank, world, local


def make_local_replays(conf, Replay, **kwargs):
    """Replay.make for the shards of this rank only."""
    import copy
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    mine = shards_of_rank(conf.num_instances, rank, world)
    local = copy.copy(conf)
    local.num_instances = len(mine)
    read, write = Replay.make(local, **kwargs)
    return mine, read, write


def bind_to_gpu_numa_node(local_rank):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs: /sys/bus/pci/devices/<bdf>/local_cpulist), so that
    pinned staging buffers are first-touched next to the GPU and host<->device copies of different ranks do not cross sockets.
    Returns the CPU set, or None when the topology cannot be read (nothing is changed then)."""
    try:
        p = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None
