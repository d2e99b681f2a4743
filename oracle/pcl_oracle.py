"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the PCL head.  Only `tests/` may import this module.

Restates, with numpy / torch-CPU (paths relative to the reference root; WSL = projects/WSL/wsl):
  * `PCLROIHeads._forward_box`, train branch           WSL/modeling/roi_heads/roi_heads_pcl.py:291-352
  * `OICROutputLayers.losses_pcl` / `PCLOutputs.pcl_loss`  WSL/modeling/roi_heads/fast_rcnn.py:1417-1442, :1725-1744
  * `PCL`, `_get_graph_centers`, `_get_top_ranking_propoals`, `_build_graph`, `_get_proposal_clusters`
                                                       WSL/modeling/roi_heads/third_party/pcl.py:27-200
  * `pcl_loss_forward_cpu` / `pcl_loss_backward_cpu`   WSL/layers/csrc/pcl_loss/pcl_loss_cpu.cpp:8-62, :64-115
    (the reference's CUDA kernel is dead code behind `&& false`, pcl_loss.h:64,101) and the scaling of
    WSL/layers/pcl_loss.py:52,119 (sum over classes / number of proposals).
  * eval: `OICROutputLayers.inference(..., pcl_bg=True)`  fast_rcnn.py:1444-1474 (background is column 0 of the refinement
    heads in PCL; it is rotated to the last column before fast_rcnn_inference).
Third-party arithmetic not under /root/reference: scikit-learn `KMeans(n_clusters<=3, random_state=3)` on the (R, 1) score column
(third_party/pcl.py:62-73; scikit-learn is an un-pinned dependency of the reference, 1.x in this image) -- called, not restated:
its k-means++ seeding consumes a numpy RandomState stream, so the oracle and the product call the same library.
Pinned against golden vectors of the UNMODIFIED reference PCL model run under oracle/refstub.py with the reference's own
pcl_loss op compiled by oracle/build_ref.py (tests/golden/pcl_r18_small*.npz, tests/golden/make_golden_pcl.py).
"""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import wsl_oracle as O

KMEANS_CLUSTERS, KMEANS_SEED = 3, 3          # third_party/pcl.py:11-12
GRAPH_IOU, MAX_PC_NUM = 0.4, 5               # :13-14
FG_THRESH, BG_THRESH = 0.5, 0.1              # :15-16


def top_ranking(probs):
    """third_party/pcl.py:62-73: rows of the k-means cluster with the highest centre."""
    from sklearn.cluster import KMeans

    n = min(KMEANS_CLUSTERS, probs.shape[0])
    km = KMeans(n_clusters=n, random_state=KMEANS_SEED).fit(probs)
    idx = np.where(km.labels_ == np.argmax(km.cluster_centers_))[0]
    return idx if len(idx) else np.array([np.argmax(probs)])


def graph_centers(boxes, cls_prob, im_labels):
    """third_party/pcl.py:88-145.  boxes R x 4, cls_prob R x K (already clipped), im_labels 1 x K.
    Returns (centre boxes P x 4, centre classes P (1-based), centre scores P)."""
    boxes, cls_prob = boxes.copy(), cls_prob.copy()
    cb, cc, cs = np.zeros((0, 4), np.float32), np.zeros((0,), np.int32), np.zeros((0,), np.float32)
    for c in range(im_labels.shape[1]):
        if im_labels[0, c] != 1:
            continue
        col = cls_prob[:, c].copy()
        pool = np.where(col >= 0)[0]
        pool = pool[top_ranking(col[pool].reshape(-1, 1))]
        pb, ps = boxes[pool].copy(), col[pool]
        iou = O.pairwise_iou(torch.from_numpy(pb), torch.from_numpy(pb)).numpy()
        graph = (iou > GRAPH_IOU).astype(np.float32)
        keep, kscore, left = [], [], ps.size
        while True:
            best = np.sum(graph, axis=1).argsort()[::-1][0]
            keep.append(best)
            nb = np.where(graph[best, :] > 0)[0]
            kscore.append(np.max(ps[nb]))
            graph[:, nb] = 0
            graph[nb, :] = 0
            left -= len(nb)
            if left <= 5:
                break
        kscore = np.array(kscore)
        order = np.argsort(kscore)[-1:(-1 - min(len(kscore), MAX_PC_NUM)):-1]
        chosen = np.array(keep)[order]
        cb = np.vstack((cb, pb[chosen]))
        cs = np.concatenate((cs, kscore[order].astype(np.float32)))
        cc = np.concatenate((cc, np.full(len(order), c + 1, np.int32)))
        # a chosen centre leaves the candidate pool of the later classes (:139-141)
        cls_prob = np.delete(cls_prob, pool[chosen], axis=0)
        boxes = np.delete(boxes, pool[chosen], axis=0)
    return cb, cc, cs


def proposal_clusters(rois, cb, cc, cs, cls_prob_new):
    """third_party/pcl.py:148-200.  Returns dict(labels R, cls_loss_weights R, gt_assignment R, pc_labels P, pc_probs P,
    pc_count P, img_cls_loss_weights P)."""
    iou = O.pairwise_iou(torch.from_numpy(rois), torch.from_numpy(cb)).numpy()
    assign = iou.argmax(axis=1)
    mx = iou.max(axis=1)
    labels = cc[assign].copy()
    w = cs[assign].copy()
    w[mx < BG_THRESH] = 0.0
    bg = mx < FG_THRESH
    labels[bg] = 0
    assign = assign.copy()
    assign[bg] = -1
    P = cb.shape[0]
    imgw, pcp, pcl, pcn = np.zeros(P, np.float32), np.zeros(P, np.float32), np.zeros(P, np.int32), np.zeros(P, np.int32)
    for i in range(P):
        members = np.where(assign == i)[0]
        imgw[i] = np.sum(w[members])
        pcl[i] = cc[i]
        pcn[i] = len(members)
        pcp[i] = np.average(cls_prob_new[members, pcl[i]])
    return dict(labels=labels, cls_loss_weights=w, gt_assignment=assign, pc_labels=pcl, pc_probs=pcp, pc_count=pcn,
                img_cls_loss_weights=imgw)


def pcl_mine(boxes, last_scores, im_labels, probs_new):
    """third_party/pcl.py:27-59 `PCL`: clip both score matrices to [1e-9, 1 - 1e-9], drop the background column of the previous
    scores when they have one, mine the cluster centres, assign the proposals."""
    cp = last_scores.detach().numpy().copy()
    cn = probs_new.detach().numpy().copy()
    if cp.shape[1] != im_labels.shape[1]:
        cp = cp[:, 1:]
    eps = 1e-9
    cp = np.clip(cp, eps, 1 - eps)
    cn = np.clip(cn, eps, 1 - eps)
    cb, cc, cs = graph_centers(boxes, cp, im_labels)
    out = proposal_clusters(boxes, cb, cc, cs, cn)
    out.update(center_boxes=cb, center_classes=cc, center_scores=cs)
    return out


def _fmaxf(x, lo):
    """C fmaxf(x, lo): a NaN operand is DROPPED (an empty cluster's mean probability is NaN; the reference's loop then
    uses eps, pcl_loss_cpu.cpp:43,51) -- torch.clamp would propagate it."""
    return torch.where(torch.isnan(x), torch.full_like(x, lo), torch.clamp(x, min=lo))


class _PCLLoss(torch.autograd.Function):
    """pcl_loss_cpu.cpp forward (:8-62) and backward (:64-115) with the scaling of wsl/layers/pcl_loss.py:52,119."""

    @staticmethod
    def forward(ctx, probs, m, im_labels_real):
        R, C = probs.shape
        p = probs.detach()
        lab = torch.from_numpy(m["labels"].astype(np.int64))
        w = torch.from_numpy(m["cls_loss_weights"].astype(np.float32))
        out = torch.zeros(C)
        bg = lab == 0
        out[0] = -(w[bg] * torch.log(_fmaxf(p[bg, 0], 1e-6))).sum()  # im_labels_real[0] == 1 always
        pcl = torch.from_numpy(m["pc_labels"].astype(np.int64))
        pcp = torch.from_numpy(m["pc_probs"].astype(np.float32))
        imgw = torch.from_numpy(m["img_cls_loss_weights"].astype(np.float32))
        for c in range(1, C):
            if im_labels_real[c] != 0:
                sel = pcl == c
                out[c] = -(imgw[sel] * torch.log(_fmaxf(pcp[sel], 1e-6))).sum()
        ctx.save_for_backward(p, lab, w, torch.from_numpy(m["gt_assignment"].astype(np.int64)), pcl, pcp,
                              torch.from_numpy(m["pc_count"].astype(np.float32)), imgw, torch.as_tensor(im_labels_real))
        return out.sum() / R

    @staticmethod
    def backward(ctx, g):
        p, lab, w, assign, pcl, pcp, pcn, imgw, iml = ctx.saved_tensors
        R, C = p.shape
        grad = torch.zeros(R, C)
        bg = lab == 0
        grad[bg, 0] = -w[bg] / _fmaxf(p[bg, 0], 1e-5)
        fg = (lab > 0) & (iml[lab.clamp(min=0)] != 0)
        rows = torch.nonzero(fg).flatten()
        a = assign[rows]
        grad[rows, lab[rows]] = -imgw[a] / _fmaxf(pcn[a] * pcp[a], 1e-5)
        return grad * (g / R), None, None


def forward_train(batched, state, spec):
    """GeneralizedRCNNWSL.forward in train mode with PCLROIHeads (one image per call: third_party/pcl.py asserts batch size 1),
    box_head dropout off.  Returns (losses, trace); trace["stages"][k] holds the mined clusters and the stage's softmax."""
    assert len(batched) == 1, "the PCL head mines clusters for one image at a time (third_party/pcl.py:92,155)"
    b = batched[0]
    K = spec.num_classes
    t = O.forward_features(b["image"], b["boxes"], b["objectness"], state, spec)
    t["scores"] = O.wsddn_scores(t["feat"], state)
    t["img_score"] = O.image_scores(t["scores"])
    gt = torch.unique(b["gt_classes"], sorted=True).to(torch.int64)
    oh = torch.zeros(1, K).scatter_(1, gt[None], 1)
    losses = {"loss_cls": F.binary_cross_entropy(t["img_score"], oh, reduction="mean" if spec.mean_loss else "sum")}
    prev = t["scores"].detach()
    real = np.concatenate(([1.0], oh.numpy()[0])).astype(np.float32)
    t["stages"] = []
    for k in range(spec.refine_num):
        pre = f"roi_heads.box_refinery_{k}."
        logits = F.linear(t["feat"], O._qw(state, pre + "cls_score.weight"), state[pre + "cls_score.bias"])
        probs = F.softmax(logits, dim=-1)
        m = pcl_mine(b["boxes"].numpy(), prev, oh.numpy(), probs)
        losses[f"loss_cls_r{k}"] = _PCLLoss.apply(probs, m, real)
        t["stages"].append(dict(m, probs=probs.detach(), logits=logits.detach()))
        prev = probs.detach()
    return losses, t


def forward_eval_scores(b, state, spec):
    """Eval branch up to (all_scores, all_boxes): mean of the refinement softmaxes with the background column rotated from
    the front to the back (fast_rcnn.py:1463-1465, :1577-1594), boxes = apply_deltas(zeros) (:1534-1559)."""
    t = O.forward_features(b["image"], b["boxes"], b["objectness"], state, spec)
    K = spec.num_classes
    probs = torch.zeros(len(b["boxes"]), K + 1)
    for k in range(spec.refine_num):
        pre = f"roi_heads.box_refinery_{k}."
        probs += F.softmax(F.linear(t["feat"], O._qw(state, pre + "cls_score.weight"), state[pre + "cls_score.bias"]), -1)
    probs = probs / spec.refine_num
    probs = torch.cat((probs[:, 1:], probs[:, :1]), 1)
    boxes = O.apply_deltas(torch.zeros(len(b["boxes"]), 4 * K), b["boxes"], spec.bbox_reg_weights)
    return {"all_scores": probs, "all_boxes": boxes}
