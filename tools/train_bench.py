"""Time one training step (step_b200.training.train_step: train.py:263-348 for already selected samples) at the C4 shape:
B clips of T=32 x 224 x 224, 11 tubes per clip, 3 refinement steps.  Correctness is covered by tests/test_gpu_train.py; this
only reports where the (not yet optimised) step stands.   python tools/train_bench.py [B]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import step_b200
from step_b200 import synth, training
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = 11
cfg = synth.make_cfg(fp16=True, T=8, max_iter=3, NUM_CHUNKS={1: 1, 2: 1, 3: 1}, image_size=(224, 224))
nets = {"base_net": step_b200.BaseNet(cfg), "roi_net": step_b200.ROINet("align", 7)}
nets["base_net"].load_state_dict(synth.base_net_state_dict())
for i in range(3):
    h = step_b200.TwoBranchNet(cfg); h.load_state_dict(synth.head_state_dict(100 + i, cfg)); nets["det_net%d" % i] = h
for k in nets:
    nets[k] = nets[k].cuda().eval()
    if hasattr(nets[k], "set_device"):
        nets[k].set_device("cuda:0")
x = synth.make_clips(B, 32, 224, 224).cuda()
props = synth.make_proposals(B, N, cfg.T, 224, 224)
flat, _ = step_b200.tube_utils.flatten_tubes(props, batch_idx=True)
tubes = torch.from_numpy(flat).cuda()
gen = torch.Generator().manual_seed(0)
tg = torch.zeros(B * N, 3, 66)
tg[:, :, :4] = tubes[:, 4:5, 1:].cpu() + torch.rand(B * N, 3, 4, generator=gen) * 6
tg[:, :, 4:6] = (torch.rand(B * N, 3, 2, generator=gen) > 0.3).float(); tg[0, :, 4:6] = 1
tg[:, :, 6:] = (torch.rand(B * N, 3, 60, generator=gen) > 0.9).float()
tg = tg.cuda()
# timing of the full step incl. an SGD update.  The update is layer-wise normalised (every tensor moves by 3e-4 of its own norm):
# the synthetic nets pair regressor weights of std 5e-5 with convolution weights of O(0.05), one global rate cannot suit both
for it in range(4):
    training.TIMING = {} if it == 3 else None
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = training.train_step(cfg, nets, x, [tubes] * 3, [tg] * 3, lr=None)
    for p, g in r["grads"].items():
        pn, gn = float(p.detach().norm()), float(g.norm())
        if pn > 0 and gn > 0:
            training.sgd_step({p: g}, lr=3e-4 * pn / gn, momentum=0.0)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(json.dumps({"iter": it, "B": B, "train_step_ms": round(dt * 1e3, 1), "loss": round(float(r["loss"]), 5),
                      "clips_per_s": round(B / dt, 1), "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2)}), flush=True)
print(json.dumps({"device_ms_by_phase_of_the_backward_tape": training.timing_summary()}))
