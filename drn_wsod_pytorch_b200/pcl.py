"""Host side of the PCL head: mining of the proposal-cluster centres.

In the reference this step (projects/WSL/wsl/modeling/roi_heads/third_party/pcl.py:62-145 -- scikit-learn k-means over one class's
scores, then a greedy cover of the IoU graph of the top-scoring proposals) runs in numpy on the host once per ground-truth class
and refinement stage; its sizes are data dependent, its output is a handful of boxes, and the k-means seeding consumes a numpy
RandomState stream, so it stays a host step here as well and calls the same library.  Everything that is per-proposal
arithmetic -- assignment of the R proposals to the centres, the cluster statistics, the loss and its gradient -- runs on the
B200 (csrc/drn_pcl.cu).  Inputs arrive as small host arrays: the G score columns of the image's classes (R x G floats) and
the proposal boxes.
"""
import numpy as np

NUM_KMEANS_CLUSTER = 3      # third_party/pcl.py:11
RNG_SEED = 3                # :12
GRAPH_IOU_THRESHOLD = 0.4   # :13
MAX_PC_NUM = 5              # :14


def _iou_matrix(b):
    """detectron2 pairwise_iou (structures/boxes.py:329-361) of a box set with itself, in float32 like the reference."""
    b = np.ascontiguousarray(b, dtype=np.float32)
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    wh = np.minimum(b[:, None, 2:], b[None, :, 2:]) - np.maximum(b[:, None, :2], b[None, :, :2])
    wh = np.maximum(wh, np.float32(0))
    inter = wh[:, :, 0] * wh[:, :, 1]
    union = area[:, None] + area[None, :] - inter
    return np.where(inter > 0, inter / np.where(inter > 0, union, np.float32(1)), np.float32(0)).astype(np.float32)


def _high_score_rows(scores):
    """Rows of the k-means cluster (k <= 3, seed 3) whose centre is the largest; the arg-max row if that cluster is empty."""
    from sklearn.cluster import KMeans

    col = scores.reshape(-1, 1)
    km = KMeans(n_clusters=min(NUM_KMEANS_CLUSTER, col.shape[0]), random_state=RNG_SEED).fit(col)
    rows = np.flatnonzero(km.labels_ == int(np.argmax(km.cluster_centers_)))
    return rows if rows.size else np.array([int(np.argmax(col))])


def _cover(adj, scores):
    """Greedy cover of the IoU graph: repeatedly take the node with the most live neighbours, score it with the best score
    in its neighbourhood, retire the neighbourhood; stop when at most five nodes are left (third_party/pcl.py:116-128)."""
    adj = adj.copy()
    picked, picked_score, alive = [], [], scores.size
    while True:
        node = int(adj.sum(axis=1).argsort()[::-1][0])
        hood = np.flatnonzero(adj[node] > 0)
        picked.append(node)
        picked_score.append(scores[hood].max())
        adj[:, hood] = 0
        adj[hood, :] = 0
        alive -= hood.size
        if alive <= 5:
            return np.array(picked), np.array(picked_score)


def mine_cluster_centres(boxes, class_scores, classes):
    """boxes [R, 4] float32; class_scores [R, G] float32 = the previous stage's (clipped) scores of the image's G classes;
    classes [G] int = their 0-based class ids, ascending.  Returns (centre boxes [P, 4] float32, centre classes [P] int32 --
    1-based, i.e. the column of the refinement head, 0 being background --, centre scores [P] float32); at most five centres
    per class, best first.  A proposal that became a centre leaves the pool of the classes after it (:139-141)."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    scores = np.ascontiguousarray(class_scores, dtype=np.float32)
    live = np.arange(boxes.shape[0])  # rows still in the pool, in order
    out_b, out_c, out_s = [], [], []
    for g, cls in enumerate(classes):
        col = scores[live, g]
        top = _high_score_rows(col)
        sub_boxes, sub_scores = boxes[live[top]], col[top]
        nodes, node_scores = _cover((_iou_matrix(sub_boxes) > GRAPH_IOU_THRESHOLD).astype(np.float32), sub_scores)
        best_first = np.argsort(node_scores)[-1:(-1 - min(node_scores.size, MAX_PC_NUM)):-1]
        chosen = nodes[best_first]
        out_b.append(sub_boxes[chosen])
        out_s.append(node_scores[best_first].astype(np.float32))
        out_c.append(np.full(chosen.size, int(cls) + 1, dtype=np.int32))
        live = np.delete(live, top[chosen])
    return np.concatenate(out_b).astype(np.float32), np.concatenate(out_c), np.concatenate(out_s)
