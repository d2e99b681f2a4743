#!/usr/bin/env python
"""Warm, graph-replayed time of the heads tail only (label_proposals + WSDDN MIL + S x (pgt, label, stage))
at the bench workload's shapes (R=4000, K=20, S=3, G=2).  GPU box only."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from drn_wsod_pytorch_b200 import ops, synth

R, K, S = 4000, 20, 3
dev = "cuda:0"
inp = synth.make_inputs(600, 1000, R, seed=0)
boxes = inp["boxes"].to(dev)
gtb, gtc = inp["gt_boxes"].to(dev), inp["gt_classes"].to(dev)
gt_int = torch.unique(gtc)
gt_oh = torch.zeros(K, device=dev); gt_oh[gt_int] = 1
logits = torch.randn(R, 128, device=dev) * 2
counter = torch.zeros(1, dtype=torch.int32, device=dev)


def tail():
    loss = torch.zeros(1, 4, device=dev)
    lab0, midx0, cnt0 = ops.label_proposals(boxes, gtb, gtc, K, [0.5], [0, 1])
    scores, img = ops.wsddn_mil(logits, K, 0, K, gt_oh, True, 1.0, loss[0, 0:1])
    prev = scores
    for k in range(S):
        pi, ps, pb, pw = ops.oicr_pgt(prev, boxes, gt_int, img, k > 0, None, 0, False, (10.0, 10.0, 5.0, 5.0))
        lab, midx, cnt = ops.label_proposals(boxes, pb, gt_int, K, [0.5], [0, 1])
        probs, stats, w = ops.oicr_stage(logits, 2 * K + k * (K + 1), K, lab, midx, pw, 1.0, loss[0, k + 1:k + 2], counter)
        prev = probs
    return loss


def tail_fused():
    loss = torch.zeros(1, 4, device=dev)
    scores, img, pgt = ops.wsddn_mil_pgt(logits, K, 0, K, gt_oh, True, 1.0, loss[0, 0:1], boxes, gt_int, counter)
    for k in range(S):
        nxt = None if k == S - 1 else dict(img_score=img, deltas=None, ld_deltas=0, cls_agnostic=False, bbox_w=(10.0, 10.0, 5.0, 5.0))
        o = ops.oicr_stage_fused(logits, 2 * K + k * (K + 1), K, boxes, gt_int, pgt[2], pgt[3], [0.5], [0, 1], 1.0,
                                 loss[0, k + 1:k + 2], counter, first_gt=(gtb, gtc) if k == 0 else None, nxt=nxt)
        pgt = o["next"]
    return loss


for name, fn in (("per-function kernels", tail), ("fused", tail_fused)):
    for _ in range(3):
        fn()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    print(f"heads tail, {name} (warm, graph): {e0.elapsed_time(e1) * 100:.1f} us per image")
