"""world_size-2 gloo runs of the multi-GPU host logic (CPU): shard -> rank assignment and the flat-bucket gradient
all-reduce of the learner.  Replay data never crosses ranks; only gradients do."""
import os
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fastdeepqlearning_b200 import Agent
        from fastdeepqlearning_b200.parallel import shards_of_rank
        torch.manual_seed(0)  # identical replicas
        conf = Agent.LearnerConf(training_device="cpu", obs_space={"obs_1d": 6, "achieved_goal": 2, "desired_goal": 2},
                                 action_space=types.SimpleNamespace(shape=(3,)), num_critics=2, num_q_predictions=5,
                                 pi_hidden_dims=(16,), critic_hidden_dims=(16, 16))
        learner = Agent.Learner(conf)
        g = torch.Generator().manual_seed(100 + rank)  # different local gradients
        local = []
        for p in learner.params:
            p.grad = torch.randn(p.shape, generator=g)
            local.append(p.grad.clone())
        learner._allreduce_grads()
        # expected: mean over ranks, recomputed from the other rank's generator
        exp = []
        for i, p in enumerate(learner.params):
            acc = torch.zeros_like(p)
            exp.append(acc)
        gens = [torch.Generator().manual_seed(100 + r) for r in range(world)]
        for i, p in enumerate(learner.params):
            exp[i] = sum(torch.randn(p.shape, generator=gens[r]) for r in range(world)) / world
        ok = all(torch.allclose(p.grad, e, rtol=1e-6, atol=1e-7) for p, e in zip(learner.params, exp))
        # after the all-reduce one optimizer step keeps replicas bit-identical
        learner.optimizer.step()
        flat = torch.cat([p.detach().reshape(-1) for p in learner.params])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        same = all(torch.equal(gathered[0], x) for x in gathered)
        mine = shards_of_rank(5, rank, world)
        ret[rank] = (ok, same, mine)
    finally:
        dist.destroy_process_group()


def test_grad_allreduce_and_sharding_world2():
    world = 2
    port = 29500 + os.getpid() % 1000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0][0] and ret[1][0], "all-reduced gradients differ from the mean over ranks"
    assert ret[0][1] and ret[1][1], "replicas diverged after the step"
    assert sorted(ret[0][2] + ret[1][2]) == [0, 1, 2, 3, 4] and ret[0][2] == [0, 2, 4] and ret[1][2] == [1, 3]


def test_shard_assignment_properties():
    from fastdeepqlearning_b200.parallel import shards_of_rank
    for n in (1, 3, 8, 13):
        for world in (1, 2, 4, 8):
            allv = sorted(sum((shards_of_rank(n, r, world) for r in range(world)), []))
            assert allv == list(range(n))
            sizes = [len(shards_of_rank(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
