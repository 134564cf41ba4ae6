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
