"""CPU, world_size 2, gloo: the multi-process host logic of bench.py -- batch sharding by rank (distinct seeds), the
max-over-ranks reduction of the step time, and the rule that only rank 0 reports (and that the reference arm runs on
rank 0 alone).  No GPU, no product kernels."""
import json
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    import bench
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    x = bench.make_images(2, 32, 32, seed=rank)                      # each rank draws ITS shard of the global batch
    got = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(got, x)
    assert not torch.equal(got[0], got[1]), "ranks must hold different shards"
    t = torch.tensor([10.0 + 5.0 * rank], dtype=torch.float64)       # pretend step times: rank 1 is the slow one
    assert bench.max_over_ranks(t, world) == 15.0
    total = bench.whole_job_images(batch_per_gpu=2, world=world, steps=3)
    assert total == 12
    if rank == 0:
        print(json.dumps({"ok": True, "world": world}))
    dist.destroy_process_group()
""") % ROOT


GRAD_WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    from autofocusformermod_b200.aff import FlatGradAllReduce
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
    frozen = torch.nn.Parameter(torch.ones(4), requires_grad=False)
    unused = torch.nn.Parameter(torch.ones(2))                        # never reaches the loss: its gradient stays None
    sync = FlatGradAllReduce(list(net.parameters()) + [frozen, unused])
    x = torch.randn(6, 5, generator=torch.Generator().manual_seed(10 + rank))     # each rank its own shard
    net(x).square().mean().backward()
    local = [p.grad.clone() for p in net.parameters()]
    sync()
    for p, g in zip(net.parameters(), local):                         # every gradient = the mean over ranks, and lives in the flat buffer
        both = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(both, g)
        assert torch.allclose(p.grad, sum(both) / world, atol=1e-7), "gradient is not the rank mean"
        assert p.grad.untyped_storage().data_ptr() == sync.flat.untyped_storage().data_ptr()
    assert unused.grad is not None and float(unused.grad.abs().sum()) == 0.0 and frozen.grad is None
    # second step: the previous views are still param.grad when backward accumulates; zero_grad(set_to_none) as bench.py does
    for p in list(net.parameters()) + [unused]:
        p.grad = None
    net(x).square().mean().backward()
    sync()
    for p, g in zip(net.parameters(), local):
        both = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(both, g)
        assert torch.allclose(p.grad, sum(both) / world, atol=1e-7)
    if rank == 0:
        print(json.dumps({"ok": True, "bytes": sync.nbytes}))
    dist.destroy_process_group()
""") % ROOT


def _torchrun(args, env_extra=None):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    env.update(env_extra or {})
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                           "--master-addr", "127.0.0.1", "--master-port", "29533"] + args,
                          cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)


def test_sharding_and_max_over_ranks_gloo(tmp_path):
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    r = _torchrun([str(w)])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines == ['{"ok": true, "world": 2}']


def test_reference_arm_runs_on_rank0_only():
    r = _torchrun(["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                   "--workload", "aff_test_fwd_b2_128"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1 and lines[0]["impl"] == "reference" and lines[0]["n_gpus"] == 2
    assert lines[0]["cpu_baseline"]["kind"] == "port" and lines[0]["value"] > 0


def test_flat_gradient_all_reduce_gloo(tmp_path):
    """aff.FlatGradAllReduce (the collective of the CUDA-graph training step under data parallelism): rank-mean gradients, frozen and
    unused parameters, param.grad aliased to the flat buffer."""
    w = tmp_path / "grad_worker.py"
    w.write_text(GRAD_WORKER)
    r = _torchrun([str(w)])
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines and lines[0]["ok"] and lines[0]["bytes"] == 4 * (5 * 7 + 7 + 7 * 3 + 3 + 2)
