"""CPU restatement of the detection decode / NMS path (TEST INFRASTRUCTURE).

Follows /root/reference/src/models/yolo_head.py (decode :220-294,
post_process :571-676, class-agnostic NMS :678-731, IoU :733-755) and
/root/reference/src/inference/postprocessing.py (class-aware NMS :505-607,
IoU :772-802, centre->corner :540-549).

IoU arithmetic is done in IEEE fp32 with the reference's operation order and
no fused multiply-add (numpy float32 scalars/arrays), because the keep
indices are a bit-exact target.  Tie-break for equal scores: lower input index
first (``torch.sort`` on CPU is not stable under ties, SURVEY.md Appendix B,
so the reference does not define it).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

F32 = np.float32

# anchor sizes in pixels of a 416-pixel image, small -> large objects
# (yolo_head.py:27-31); normalised by 416 (:50-51).
DEFAULT_ANCHORS = (
    ((10, 13), (16, 30), (33, 23)),
    ((30, 61), (62, 45), (59, 119)),
    ((116, 90), (156, 198), (373, 326)),
)


def anchors_wh(scale_idx: int, anchors=DEFAULT_ANCHORS) -> torch.Tensor:
    """[A,2] fp32 anchor (w,h)/416 of scale ``scale_idx`` (:48-51).  Scale 0 is
    the stride-8 map and takes the small-object anchors (repair R5:
    configs/base.yaml:69-73 pairs small<->finest grid)."""
    a = torch.tensor(anchors[scale_idx], dtype=torch.float32)
    return a / 416.0


def yolo_decode(pred: torch.Tensor, anchor_wh: torch.Tensor) -> Dict[str, torch.Tensor]:
    """YOLODecoder.forward (:220-294) with repair R7 (box_x/box_y indexed on the
    last dim so boxes are [B,A,H,W,4], not [B,A,H,W,W,4]).

    pred [B,A,H,W,5+C] (any strides), anchor_wh [A,2] (w,h normalised by 416).
    """
    b, a, h, w, _ = pred.shape
    p = pred.to(torch.float32)
    sx = torch.sigmoid(p[..., 0])
    sy = torch.sigmoid(p[..., 1])
    obj = torch.sigmoid(p[..., 4:5])
    cls = torch.sigmoid(p[..., 5:])
    gy, gx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    bx = (gx.view(1, 1, h, w) + sx) / w                    # :262
    by = (gy.view(1, 1, h, w) + sy) / h                    # :263
    aw = anchor_wh[:, 0].view(1, a, 1, 1)
    ah = anchor_wh[:, 1].view(1, a, 1, 1)
    bw = aw * torch.exp(p[..., 2])                         # :269
    bh = ah * torch.exp(p[..., 3])                         # :270
    boxes = torch.stack([bx - bw / 2, by - bh / 2, bx + bw / 2, by + bh / 2], dim=-1)
    scores = obj * cls                                     # :282
    class_scores, class_indices = torch.max(scores, dim=-1)  # :285
    return {"boxes": boxes, "scores": scores, "class_scores": class_scores,
            "class_indices": class_indices, "objectness": obj}


def _max_np(a, b):
    # torch.max(a, b) propagates NaN; np.maximum does too.
    return np.maximum(a, b)


def iou_one_to_many(box: np.ndarray, others: np.ndarray) -> np.ndarray:
    """compute_iou (yolo_head.py:733-755) == _compute_iou (postprocessing.py:772-802):
    ``inter / (((a1 + a2) - inter) + 1e-6)`` in fp32, xyxy boxes."""
    box = box.astype(F32)
    others = others.astype(F32)
    ix1 = np.maximum(box[0], others[:, 0])
    iy1 = np.maximum(box[1], others[:, 1])
    ix2 = np.minimum(box[2], others[:, 2])
    iy2 = np.minimum(box[3], others[:, 3])
    iw = ix2 - ix1
    ih = iy2 - iy1
    iw = np.where(iw < 0, F32(0), iw)          # clamp(min=0): NaN stays NaN
    ih = np.where(ih < 0, F32(0), ih)
    inter = iw * ih
    a1 = (box[2] - box[0]) * (box[3] - box[1])
    a2 = (others[:, 2] - others[:, 0]) * (others[:, 3] - others[:, 1])
    union = ((a1 + a2) - inter) + F32(1e-6)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (inter / union).astype(F32)


def _order_desc(scores: np.ndarray) -> np.ndarray:
    # descending score, ties -> lower index first
    return np.lexsort((np.arange(len(scores)), -scores.astype(np.float64)))


def nms_agnostic(boxes, scores, iou_threshold: float = 0.5,
                 max_detections: int = 100) -> np.ndarray:
    """YOLODetectionHead.non_max_suppression (:678-731).  Greedy, class-agnostic:
    take the best remaining box, stop once ``max_detections`` are kept (:710),
    keep a remaining box only if ``iou < thr`` (:727; a NaN IoU is suppressed).
    Returns int64 indices into the input, descending score."""
    boxes = np.asarray(boxes, dtype=F32).reshape(-1, 4)
    scores = np.asarray(scores, dtype=F32).reshape(-1)
    if len(boxes) == 0:
        return np.zeros(0, np.int64)
    thr = F32(iou_threshold)
    order = _order_desc(scores)
    keep: List[int] = []
    while len(order) > 0:
        cur = order[0]
        keep.append(int(cur))
        if len(keep) >= max_detections:
            break
        order = order[1:]
        if len(order) == 0:
            break
        iou = iou_one_to_many(boxes[cur], boxes[order])
        order = order[iou < thr]
    return np.asarray(keep, np.int64)


def center_to_corner(boxes: np.ndarray) -> np.ndarray:
    """NMSFilter._center_to_corner (postprocessing.py:540-549), fp32."""
    b = np.asarray(boxes, dtype=F32).reshape(-1, 4)
    cx, cy, w, h = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    two = F32(2)
    return np.stack([cx - w / two, cy - h / two, cx + w / two, cy + h / two], axis=-1).astype(F32)


def nms_class_aware(boxes_cxcywh, scores, class_ids, iou_threshold: float = 0.45,
                    max_detections: int = 100, boxes_are_corners: bool = False) -> np.ndarray:
    """NMSFilter.apply -> _standard_nms (postprocessing.py:505-607) with repair
    R9 (``ious`` squeezed so the same-class mask indexes it).  A kept box
    suppresses later boxes of the SAME class with ``iou > thr`` (:594; a NaN
    IoU is kept); the result is truncated to ``max_detections`` (:604-605)."""
    b = np.asarray(boxes_cxcywh, dtype=F32).reshape(-1, 4)
    if len(b) == 0:
        return np.zeros(0, np.int64)
    boxes = b if boxes_are_corners else center_to_corner(b)
    scores = np.asarray(scores, dtype=F32).reshape(-1)
    cls = np.asarray(class_ids).reshape(-1)
    thr = F32(iou_threshold)
    order = _order_desc(scores)
    sb, sc = boxes[order], cls[order]
    n = len(order)
    alive = np.ones(n, bool)
    for i in range(n):
        if not alive[i]:
            continue
        later = alive.copy()
        later[: i + 1] = False
        later &= sc == sc[i]
        idx = np.nonzero(later)[0]
        if len(idx) == 0:
            continue
        iou = iou_one_to_many(sb[i], sb[idx])
        alive[idx[iou > thr]] = False
    keep = order[alive]
    return keep[:max_detections].astype(np.int64)


def post_process(decoded: Sequence[Dict[str, torch.Tensor]],
                 confidence_threshold: float = 0.5, iou_threshold: float = 0.5,
                 max_detections: int = 100) -> List[Dict[str, np.ndarray]]:
    """YOLODetectionHead.post_process (:571-676): per scale, flatten row-major
    over (A,H,W) (:600-602), keep ``score > conf`` (:605), NMS capped at
    ``max_detections`` (:625-629); concatenate the per-scale survivors in scale
    order (:646-654) and run a second NMS over them (:658-662).

    ``decoded`` is the list of per-scale decode dicts.  Returns per image
    dict(boxes [K,4], scores [K], labels [K], and the bookkeeping the GPU path
    is compared on: per-scale keep indices into the compacted candidate lists
    and the final keep indices into the concatenation)."""
    nb = decoded[0]["class_scores"].shape[0]
    out = []
    for b in range(nb):
        per_scale = []
        cat_boxes, cat_scores, cat_labels = [], [], []
        for d in decoded:
            boxes = d["boxes"][b].reshape(-1, 4).numpy().astype(F32)
            scores = d["class_scores"][b].reshape(-1).numpy().astype(F32)
            labels = d["class_indices"][b].reshape(-1).numpy().astype(np.int64)
            mask = scores > F32(confidence_threshold)
            cb, cs, cl = boxes[mask], scores[mask], labels[mask]
            keep = nms_agnostic(cb, cs, iou_threshold, max_detections)
            per_scale.append(keep)
            cat_boxes.append(cb[keep]); cat_scores.append(cs[keep]); cat_labels.append(cl[keep])
        ab = np.concatenate(cat_boxes).reshape(-1, 4)
        asc = np.concatenate(cat_scores)
        al = np.concatenate(cat_labels)
        keep2 = nms_agnostic(ab, asc, iou_threshold, max_detections)
        out.append({"boxes": ab[keep2], "scores": asc[keep2], "labels": al[keep2],
                    "scale_keep": per_scale, "final_keep": keep2})
    return out
