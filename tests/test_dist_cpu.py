"""CPU, world_size 2, gloo: the N>1 host logic of bench.py's clip-parallel mode -- each rank owns a
disjoint batch, fixed-shape detections are all-gathered once per batch, rank order is preserved and the
max-over-ranks timing reduction works.  (No GPU kernels are called here.)"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from step_b200 import synth
    B, cap = 2, 5
    # rank r owns clips [r*B, (r+1)*B): per-rank seed as in bench.py; what is gathered has the layout of bench.py's
    # step(): per clip the compact detections [cap, 8] of step_detect_f32 flattened, then the count (stand-in values here:
    # the kernels need a GPU; tests/test_gpu_multi.py checks the real thing on 2 GPUs)
    clips = synth.make_clips(B, 4, 8, 8, seed=1234 + rank)
    det = torch.full((B, cap * 8 + 1), float(rank)) + clips.mean()
    gather = torch.empty((world, B, cap * 8 + 1))
    dist.all_gather_into_tensor(gather.view(-1, cap * 8 + 1), det)
    ms = torch.tensor([10.0 + rank])
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        torch.save({"gather": gather, "ms": ms}, out)
    dist.destroy_process_group()


def test_clip_parallel_gather_two_ranks(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    from step_b200 import synth
    for rank in range(2):
        exp = float(rank) + synth.make_clips(2, 4, 8, 8, seed=1234 + rank).mean()
        assert torch.allclose(r["gather"][rank], torch.full((2, 41), float(exp)))
    assert float(r["ms"]) == 11.0


def test_reference_arm_under_torchrun_prints_one_line():
    """`bench.py --impl reference` launched the way the driver launches N > 1 (torchrun, one process per GPU):
    rank 0 alone runs the reference arithmetic on the host cores and prints ONE JSON line, the other ranks exit 0."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "0"]
    r = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0 and d["unit"] == "clips/s"
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["e2e"]["h2d_bytes_per_step"] == 0
