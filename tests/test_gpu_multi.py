"""GPU, more than one device (skipped on a 1-GPU box; run with `gpurun --gpus 2`):
  * clip-parallel sharding over 2 ranks + one NCCL gather reproduces the 1-GPU detections bit for bit (SURVEY.md
    section 4 last bullet / section 8e);
  * the reference drivers' head placement, det_net i on cuda:(i+1) % gpu_count (test.py:85-87), works: every launch
    runs on the device and stream that own its tensors."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

from step_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

needs2 = pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")


@needs2
def test_two_rank_sharding_equals_single_gpu_bit_for_bit():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _shard_worker as w
    B, T_in, HW, N = 4, 16, 112, 5
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "gathered.npy")
        env = dict(os.environ, NCCL_DEBUG="WARN")
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
               "--master-port", "29731", os.path.join(ROOT, "tests", "_shard_worker.py"), out, str(B), str(T_in), str(HW), str(N)]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        gathered = np.load(out)
    dev = torch.device("cuda", 0)
    clips = synth.make_clips(2 * B, T_in, HW, HW, seed=77)
    det = dict(conf_thresh=0.01, nms_thresh=0.4, topk=50)
    single = torch.stack([w.run_shard(dev, clips[r * B:(r + 1) * B].contiguous(), B, T_in, HW, N, det) for r in range(2)])
    assert gathered.shape == tuple(single.shape)
    assert np.array_equal(gathered, single.cpu().numpy())
    assert gathered[..., -1].min() > 0            # every clip produced detections


@needs2
def test_head_on_another_gpu_like_the_reference_driver():
    """test.py:85-87: nets['det_net%d' % i].to('cuda:%d' % ((i+1) % gpu_count)) + set_device."""
    import step_b200
    cfg = synth.make_cfg(fp16=True, T=4, max_iter=2, NUM_CHUNKS={1: 1, 2: 1}, image_size=(112, 112))

    def build(place):
        nets = {"base_net": step_b200.BaseNet(cfg), "roi_net": step_b200.ROINet(cfg.pool_mode, cfg.pool_size)}
        nets["base_net"].load_state_dict(synth.base_net_state_dict())
        for i in range(cfg.max_iter):
            h = step_b200.TwoBranchNet(cfg)
            h.load_state_dict(synth.head_state_dict(100 + i, cfg))
            nets["det_net%d" % i] = h
        for k in nets:
            nets[k] = nets[k].cuda().eval()
        gpu_count = torch.cuda.device_count()
        for i in range(cfg.max_iter):
            d = "cuda:%d" % (((i + 1) % gpu_count) if place else 0)
            nets["det_net%d" % i].to(d)
            nets["det_net%d" % i].set_device(d)
        return nets

    x = synth.make_clips(2, 16, 112, 112).cuda()
    tubes = synth.make_proposals(2, 3, 4, 112, 112)
    outs = []
    for place in (False, True):
        nets = build(place)
        with torch.no_grad():
            cf = nets["base_net"](x)
            hist, _ = step_b200.inference(cfg, cf, None, nets, cfg.max_iter, tubes)
        torch.cuda.synchronize()
        outs.append([(h["pred_prob"][:, 0].cpu(), h["pred_loc"].cpu()) for h in hist])
    for (p0, l0), (p1, l1) in zip(*outs):
        assert p1.device.type == "cpu" and torch.equal(p0, p1) and torch.equal(l0, l1)
    # the module-level call of the drivers (two_branch.py:225-229): features on cuda:0, head on cuda:1
    nets = build(True)
    with torch.no_grad():
        pooled = torch.randn(3, 4, 832, 7, 7, device="cuda:0")
        prob = nets["det_net0"](pooled)[0]
    assert prob.device == torch.device(nets["det_net0"].device)


@needs2
def test_two_rank_training_step_all_reduce_equals_averaged_single_process():
    """train_step.sh's data-parallel update: two ranks, each with its own clip, gradients all-reduced over NCCL inside
    training.sgd_step -- equals one process that computes both clips' gradients, averages them and applies the same SGD."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _shard_worker as w
    from step_b200 import training
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "params.npz")
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
               "--master-port", "29733", os.path.join(ROOT, "tests", "_shard_worker.py"), "train", out]
        r = subprocess.run(cmd, env=dict(os.environ, NCCL_DEBUG="WARN"), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        got = dict(np.load(out))
    dev = torch.device("cuda", 0)
    cfg, nets = w.train_nets(dev)
    grads = []
    for rank in range(2):
        x, tubes, tg = w.train_inputs(rank)
        grads.append(training.train_step(cfg, nets, x.to(dev), [tubes.to(dev)], [tg.to(dev)], lr=None)["grads"])
    mean = {p: (grads[0][p] + grads[1][p]) / 2 for p in grads[0]}
    before = w.watched(nets)
    training.sgd_step(mean, lr=0.05, momentum=0.9, weight_decay=1e-4, world_size=1)
    after = w.watched(nets)
    for k in w.WATCH:
        assert np.abs(after[k] - before[k]).max() > 0                      # the step moved the weights
        assert np.allclose(got[k], after[k], rtol=1e-5, atol=1e-7), k
