"""`StepRunner` -- one CUDA graph for the whole batch step.

The reference runs ~600 PyTorch kernels per batch from Python with B*max_iter host round trips
(SURVEY.md section 3.1).  Our step is ~155 launches with no host dependency, so the entire
trunk -> (ROIAlign -> head -> tube update) x max_iter chain is captured ONCE per (B, T, H, W, N) shape into
a CUDA graph (TMA descriptors and buffer addresses are baked at capture) and replayed with a single
`cudaGraphLaunch`: launch latency and Python overhead disappear from the timed path.

    runner = StepRunner(cfg, nets, B, T_in, H, W, tubes)     # captures
    hist = runner(clips)          # clips: [B,T_in,3,H,W] fp32, CUDA or pinned host (copied in asynchronously)
`hist` is the same structure `inference()` returns (tensors are static graph outputs: copy them out
before the next call if they must survive it).
"""
import torch

from . import _lib as L
from . import engine as E
from .inference import inference_device, stage_tubes
from .postprocess import Detector


class StepRunner:
    def __init__(self, cfg, nets, B, T_in, H, W, tubes, device=None, context=False, use_graph=True, warmup=2,
                 detect=None):
        """detect: None, or dict(conf_thresh, nms_thresh, topk[, steps]) -- the reference drivers' detection
        post-processing (test.py:156-218: confidence threshold, valid_tubes, per-class NMS, top-k) appended to the
        captured step for the listed refinement steps (default: the last one); results in `self.detections`."""
        self.cfg, self.nets = cfg, nets
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.context = context and not cfg.no_context
        self.x = torch.zeros((B, T_in, 3, H, W), dtype=torch.float32, device=self.device)
        self.flat, self.clip_of_tube, self.tubes_nums = stage_tubes(tubes, self.device)
        self.graph = None
        self.history = None
        self.detections = None
        self.detectors = {}
        if detect is not None:
            for i in detect.get("steps", [cfg.max_iter - 1]):
                self.detectors[i] = Detector(self.tubes_nums, cfg.num_classes, self.device, detect["conf_thresh"],
                                             detect["nms_thresh"], W, H, topk=detect.get("topk", 0))
        if not use_graph:
            return
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):  # packs weights, sets kernel attributes, warms the allocator
                self._body()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.history = self._body()
        self.graph = g

    def _body(self):
        with torch.no_grad():
            feat = self.nets["base_net"].forward_act(self.x)
            ctx_all = self.nets["context_net"].forward_act(feat) if self.context else None
            hist, _ = inference_device(self.cfg, feat, ctx_all, self.nets, self.cfg.max_iter, self.flat,
                                       self.clip_of_tube, self.tubes_nums)
            if self.detectors:
                self.detections = {i: d.run(hist[i]["pred_prob"], hist[i]["pred_loc"]) for i, d in self.detectors.items()}
        return hist

    def __call__(self, clips=None):
        if clips is not None and clips.data_ptr() != self.x.data_ptr():
            self.x.copy_(clips, non_blocking=True)
        if self.graph is None:
            return self._body()
        self.graph.replay()
        return self.history
