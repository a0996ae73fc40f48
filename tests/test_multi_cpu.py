"""N>1 host logic on CPU: world-size-2 `gloo` runs of the data-parallel path.

The ranks use the product's own sharding / all-reduce helpers (mfnerf_b200/dist.py); the per-shard gradients come from the CPU
oracle (oracle/cpu_step.py) because the CUDA kernels cannot run here.  Property checked: sum-over-ranks of the shard gradients,
scaled by dist.grad_scale, equals the gradient of the single concatenated batch -- i.e. ray sharding + one all-reduce is the
same optimisation step as the reference's single-GPU step on the union of the rays."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "mf-nerf_b200")); sys.path.insert(0, os.path.join({root!r}, "tests"))
from mfnerf_b200 import dist as mdist
from mfnerf_b200 import synthetic as syn
from oracle.cpu_step import CpuTrainer
rank, local, world = mdist.env_world()
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=world)
R = 96
def grads_of(o, d, tgt):
    tr = CpuTrainer(scale=0.5, log2_T=12, threads=1, seed=mdist.model_seed(7))
    tr.set_density_grid(syn.lego_density_grid(0.5, 1))
    class FixedJitter:                                          # same jitter for every ray so that shards and union march identically
        def rand(self, n):
            return np.full(n, 0.5, np.float64)
    tr.rng = FixedJitter()
    tr.opt.step = lambda: None                                  # gradients only
    loss, n = tr.train_step(o, d, tgt)
    return torch.cat([p.grad.reshape(-1) for p in tr.model.parameters()]).double(), n
# every rank draws its own rays ...
o, d, _, _ = syn.random_rays(R, seed=mdist.shard_seed(11, rank))
tgt = syn.analytic_render(o, d).numpy().astype(np.float32)
g, n = grads_of(o, d, tgt)
LOSS_SCALE = 128.0
flat = (g * LOSS_SCALE).clone()
mdist.allreduce_gradients(flat, world)
flat *= mdist.grad_scale(LOSS_SCALE, world)
# ... and rank 0 recomputes the union batch in one piece
all_o = [None] * world; dist.all_gather_object(all_o, (o, d, tgt))
ms = mdist.max_over_ranks(10.0 + rank, "cpu", world); tot = mdist.sum_over_ranks(n, "cpu", world)
if rank == 0:
    O = np.concatenate([x[0] for x in all_o]); D = np.concatenate([x[1] for x in all_o]); T = np.concatenate([x[2] for x in all_o])
    gu, nu = grads_of(O, D, T)
    err = float((flat - gu).abs().max()); ref = float(gu.abs().max())
    print(json.dumps(dict(err=err, ref=ref, n_union=nu, n_sum=tot, ms=ms, tiles=[mdist.tile_rows(800, r, 3) for r in range(3)])))
dist.barrier(); dist.destroy_process_group()
'''


def _free_port():
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def test_ray_sharded_allreduce_equals_union_batch(tmp_path):
    port = _free_port()
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=port))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=600) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-2000:]
    res = json.loads(outs[0][0].strip().splitlines()[-1])
    assert res["n_union"] == res["n_sum"] > 0                      # shards march exactly the rays of the union
    assert res["err"] <= 1e-6 * max(res["ref"], 1e-12) + 1e-12, res  # fp64 comparison of fp32 gradients: summation order only
    assert res["ms"] == 11.0                                       # max over ranks
    assert res["tiles"] == [[0, 267], [267, 534], [534, 800]]


def test_tile_rows_cover_image_exactly():
    sys.path.insert(0, os.path.join(ROOT, "mf-nerf_b200"))
    from mfnerf_b200 import dist as mdist
    for H in (800, 756, 7):
        for w in (1, 2, 4, 8):
            rows = [mdist.tile_rows(H, r, w) for r in range(w)]
            assert rows[0][0] == 0 and rows[-1][1] == H
            assert all(rows[i][1] == rows[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in rows) - min(b - a for a, b in rows) <= 1


def test_reference_arm_runs_on_rank0_only():
    """bench.py --impl reference under a 2-rank launch: rank 0 prints the one JSON line, rank 1 exits 0 silently"""
    outs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MFN_REF_BUDGET_S="2")
        p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                           env=env, capture_output=True, text=True, timeout=900)
        assert p.returncode == 0, p.stderr[-2000:]
        outs.append(p.stdout.strip())
    assert outs[1] == ""
    line = json.loads(outs[0].splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "train_rays_per_sec" and line["n_gpus"] == 2
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] == "port" and line["value"] > 0


WARMUP_WORKER = r'''
import os, sys, json, time
import torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "mf-nerf_b200"))
from mfnerf_b200 import dist as mdist
rank, local, world = mdist.env_world()
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=world)
# a stub of bench.py's loop structure: every "training step" enqueues the collectives of engine._step_dp (reduce-scatter, flag
# all-reduce, all-gather); the ranks' clocks run at very different speeds (rank 1's is 5x faster and it is a slower stepper)
calls = dict(n=0)
g = torch.ones(8); shard = torch.zeros(8 // world); flag = torch.zeros(1, dtype=torch.int32); full = torch.zeros(8)
def step():
    calls["n"] += 1
    dist.all_reduce(g); dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    dist.all_gather_into_tensor(full, shard)
    g.fill_(1.0)
    if rank == 1:
        time.sleep(0.0005)
t0 = time.perf_counter()
clock = (lambda: time.perf_counter()) if rank == 0 else (lambda: t0 + 5.0 * (time.perf_counter() - t0))
n = mdist.agreed_warmup(step, world, "cpu", seconds=0.15, chunk=16, max_chunks=1000, clock=clock)
dist.barrier()      # would mis-pair with a straggler's all_reduce (and hang) if the counts differed
print(json.dumps(dict(rank=rank, n=n, calls=calls["n"])))
dist.destroy_process_group()
'''


def test_untimed_warmup_runs_the_same_number_of_steps_on_every_rank(tmp_path):
    """round 1's SCALE failure: a per-rank wall-clock warm-up loop let the ranks issue different numbers of collectives"""
    port = _free_port()
    script = tmp_path / "warm.py"
    script.write_text(WARMUP_WORKER.format(root=ROOT, port=port))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-2000:]
    res = [json.loads(o[0].strip().splitlines()[-1]) for o in outs]
    assert res[0]["n"] == res[1]["n"] == res[0]["calls"] == res[1]["calls"] > 0
    assert res[0]["n"] % 16 == 0


def test_bench_has_no_rank_local_wall_clock_loop_around_collectives():
    """static guard: the only `while ... perf_counter()` loops in bench.py are in the CPU reference arm (no collectives there)"""
    import re
    src = open(os.path.join(ROOT, "bench.py")).read()
    ours = src[src.index("def run_ours"):src.index("def main")]
    assert not re.search(r"while[^\n]*perf_counter", ours)
    assert "agreed_warmup" in ours
