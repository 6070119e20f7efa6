"""Flip-test inference as one replayable CUDA graph (BASELINE configs[4]).

What the reference does per validation batch (lib/core/function.py:212-240, lib/core/inference.py:18-46): forward,
`np.flip(input.cpu().numpy(), 3)` + H2D, second forward, D2H of the flipped heat-maps, `flip_back` in numpy, H2D, shift by
one pixel, average, D2H of the merged heat-maps, numpy arg-max. Here, per batch:

    graph replay { image -> NHWC and W-mirrored NHWC (one pass each)
                   forward (eval BN, cached 3xFP16 weights)      | second stream: forward of the mirrored image
                   fused flip_back + shift + average + arg-max -> idx[B,J], maxval[B,J] (+ merged heat-maps)
                   [optional box NMS of the detector boxes that produced the crops: lib/nms, csrc/nms.cu] }

Only [B,J] indices and maxima need to cross PCIe. The module must be one of fpd_b200.lib.models (hourglass / pose_hrnet).
"""
import torch

from . import ops


class FlipTestInference:
    def __init__(self, net, flip_pairs, shift_heatmap=True, flip_test=True, use_graph=True, want_avg=True):
        self.net = net
        self.flip_test = flip_test
        self.shift = bool(shift_heatmap)
        self.use_graph = use_graph
        self.want_avg = want_avg
        self.pairs = [tuple(p) for p in flip_pairs]
        self.graph = None
        self.static_x = None
        self.static_boxes = None
        self.result = None
        self._shape = None
        self._gen = -1
        self._side = torch.cuda.Stream() if flip_test else None
        self._keep = None
        self.nms_thresh = None
        self._stage_x = self._copy_stream = self._stage_ev = self._stage_free = None
        self._staged_for = None
        net.eval()

    def _perm(self, J, device):
        p = list(range(J))
        for a, b in self.pairs:
            p[a], p[b] = b, a
        return torch.tensor(p, dtype=torch.int32, device=device)

    def _body(self, x, boxes):
        eng = self.net.engine()
        main = torch.cuda.current_stream()
        nhwc = ops.nchw_to_nhwc(x)
        ctx_f = None
        if self.flip_test:
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                nhwc_f = ops.nchw_to_nhwc_flipw(x)
                ctx_f = eng.forward(x, False, record_tape=False, shared_stem={"img": x, "nhwc": nhwc_f})
        ctx = eng.forward(x, False, record_tape=False, shared_stem={"img": x, "nhwc": nhwc})
        hm = ctx.outs[-1].data
        hm_f = None
        if ctx_f is not None:
            main.wait_stream(self._side)
            hm_f = ctx_f.outs[-1].data
        J = hm.shape[-1]
        if getattr(self, "_perm_t", None) is None or self._perm_t.numel() != J:
            self._perm_t = self._perm(J, x.device)
        avg, idx, maxval = ops.flip_merge_argmax(hm, hm_f, self._perm_t if hm_f is not None else None, shift=self.shift,
                                                 want_avg=self.want_avg)
        keep = num = None
        if boxes is not None:
            keep, num = ops.nms_device(boxes, self.nms_thresh)
        self._keep = (ctx, ctx_f, nhwc)      # every intermediate stays referenced: the graph pool never recycles them
        return {"idx": idx, "maxval": maxval, "avg_nhwc": avg, "nms_keep": keep, "nms_num": num}

    def _stage(self, x):
        main = torch.cuda.current_stream()
        if self._staged_for is not None and self._staged_for is x:
            main.wait_event(self._stage_ev)
            self.static_x.copy_(self._stage_x, non_blocking=True)
            self._stage_free.record(main)
        else:
            self.static_x.copy_(x, non_blocking=True)
        self._staged_for = None

    def _prefetch(self, next_x):
        """H2D of the next batch (pinned host tensor) on a copy stream, under this batch's graph replay."""
        if next_x is None or next_x.is_cuda or not next_x.is_pinned() or tuple(next_x.shape) != tuple(self.static_x.shape):
            return
        if self._stage_x is None or self._stage_x.shape != self.static_x.shape:
            self._stage_x = torch.empty_like(self.static_x)
            self._copy_stream = torch.cuda.Stream()
            self._stage_ev, self._stage_free = torch.cuda.Event(), torch.cuda.Event()
            self._stage_free.record(torch.cuda.current_stream())
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._stage_free)
            self._stage_x.copy_(next_x, non_blocking=True)
            self._stage_ev.record(self._copy_stream)
        self._staged_for = next_x

    def __call__(self, x, boxes_sorted=None, nms_thresh=0.6, next_x=None):
        """x: [B,3,H,W] fp32, CUDA or pinned host. boxes_sorted: optional device [n,5] detector boxes sorted by score.
        next_x: optional pinned host tensor = the NEXT batch (its H2D copy overlaps this batch; pass it as `x` next time).
        Returns a dict of DEVICE tensors (static across calls in graph mode): idx int32 [B,J] (flat arg-max, first maximum),
        maxval [B,J], avg_nhwc [B,h,w,J] merged heat-maps, nms_keep / nms_num."""
        self.nms_thresh = float(nms_thresh)
        if not self.use_graph:
            xd = x.cuda(non_blocking=True).float().contiguous()
            return self._body(xd, boxes_sorted)
        shape = (tuple(x.shape), None if boxes_sorted is None else tuple(boxes_sorted.shape), self.nms_thresh)
        gen = self.net.engine().generation
        if self.graph is None or shape != self._shape or gen != self._gen:
            xd = x.cuda(non_blocking=True).float().contiguous()
            self.static_x = torch.empty_like(xd)
            self.static_x.copy_(xd)
            self.static_boxes = None if boxes_sorted is None else boxes_sorted.clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                with torch.no_grad():
                    self._body(self.static_x, self.static_boxes)      # warm-up: lazy init, workspaces, weight cache
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            from . import _native as N
            n0 = N.lib().fpd_launch_count()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.result = self._body(self.static_x, self.static_boxes)
            self.launches = int(N.lib().fpd_launch_count() - n0)
            self._shape, self._gen = shape, self.net.engine().generation
        self._stage(x)
        if boxes_sorted is not None:
            self.static_boxes.copy_(boxes_sorted, non_blocking=True)
        self.graph.replay()
        self._prefetch(next_x)
        return self.result
