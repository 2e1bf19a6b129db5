"""Detection decode / NMS modules with the reference's signatures.

* ``YOLODecoder``        src/models/yolo_head.py:206-294   (forward(predictions, anchors, grid_size) -> dict)
* ``YOLODetectionHead``  src/models/yolo_head.py:468-755   (forward / post_process / non_max_suppression /
                         compute_iou; same sub-module names and state_dict keys)
* ``NMSFilter``          src/inference/postprocessing.py:498-607 (apply(boxes, scores, class_ids))

Repairs baked in: anchors are kept per scale as [S,A,1,1,4] and paired
small<->finest grid (D2/D3), the decoder indexes the last dim so boxes are [B,A,H,W,4] (D4), the
class-aware NMS compares 1-D IoUs (D9).  Decode and NMS run in libhvs_b200.so.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import ops
from .mhc import ManifoldHyperConnection

DEFAULT_ANCHORS = [[(10, 13), (16, 30), (33, 23)], [(30, 61), (62, 45), (59, 119)], [(116, 90), (156, 198), (373, 326)]]


class YOLOAnchorGenerator(nn.Module):
    """yolo_head.py:11-90; buffer ``anchors`` [S,A,1,1,4] with (w,h)/416 in slots 2:4."""

    def __init__(self, anchor_sizes=None, grid_sizes=(13, 26, 52)):
        super().__init__()
        self.anchor_sizes = anchor_sizes if anchor_sizes is not None else DEFAULT_ANCHORS
        self.grid_sizes = list(grid_sizes)
        self.num_scales = len(self.anchor_sizes)
        self.num_anchors = len(self.anchor_sizes[0])
        wh = torch.tensor(self.anchor_sizes, dtype=torch.float32) / 416.0
        anchors = torch.zeros(self.num_scales, self.num_anchors, 1, 1, 4)
        anchors[..., 0, 0, 2:4] = wh
        self.register_buffer("anchors", anchors)

    def forward(self, scale_idx: int) -> torch.Tensor:
        return self.anchors[scale_idx]

    def get_num_anchors(self) -> int:
        return self.num_anchors


class YOLODecoder(nn.Module):
    """YOLODecoder(image_size=416).forward(predictions [B,A,H,W,5+C], anchors [A,*,*,4], grid_size)."""

    def __init__(self, image_size: int = 416):
        super().__init__()
        self.image_size = image_size

    def forward(self, predictions: torch.Tensor, anchors: torch.Tensor, grid_size: Tuple[int, int] = None,
                want_scores: bool = True) -> Dict[str, torch.Tensor]:
        a = predictions.shape[1]
        anchor_wh = anchors.reshape(a, -1, 4)[:, 0, 2:4]
        out = ops.yolo_decode(predictions, anchor_wh, want_scores=want_scores, want_objectness=True)
        out["raw_predictions"] = predictions
        return out


class YOLOPredictionHead(nn.Module):
    """yolo_head.py:93-203: conv-BN-LeakyReLU x2 -> mHC over [B*H*W, C] -> 1x1 conv -> [B,A,H,W,5+C] view."""

    def __init__(self, in_channels: int, num_classes: int = 80, num_anchors: int = 3, use_mhc: bool = True):
        super().__init__()
        self.in_channels, self.num_classes, self.num_anchors = in_channels, num_classes, num_anchors
        self.output_dim = num_anchors * (5 + num_classes)
        self.conv_layers = nn.Sequential(
            nn.Conv2d(in_channels, in_channels * 2, 3, padding=1), nn.BatchNorm2d(in_channels * 2), nn.LeakyReLU(0.1),
            nn.Conv2d(in_channels * 2, in_channels, 3, padding=1), nn.BatchNorm2d(in_channels), nn.LeakyReLU(0.1))
        self.mhc_enhance = ManifoldHyperConnection(input_dim=in_channels, expansion_rate=2) if use_mhc else nn.Identity()
        self.pred_conv = nn.Conv2d(in_channels, self.output_dim, kernel_size=1)
        for m in self.conv_layers:
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="leaky_relu")
                nn.init.zeros_(m.bias)
        nn.init.normal_(self.pred_conv.weight, std=0.01)
        nn.init.zeros_(self.pred_conv.bias)
        with torch.no_grad():                              # :165-168
            bias = self.pred_conv.bias.view(num_anchors, -1)
            bias[:, 4] = -4.0
            bias[:, 5:] = -math.log((1 - 0.01) / 0.01) / num_classes

    def _pred_operands(self):
        """pred_conv as a GEMM B operand: [256, C_in] bf16 (255 real rows + one zero row) and the padded fp32 bias, cached
        until the parameters change."""
        wt, bs = self.pred_conv.weight, self.pred_conv.bias
        key = (wt.data_ptr(), wt._version, bs.data_ptr(), bs._version)
        if getattr(self, "_pred_key", None) != key:
            w256 = torch.zeros((256, self.in_channels), dtype=torch.bfloat16, device=wt.device)
            w256[: self.output_dim] = wt.detach().reshape(self.output_dim, self.in_channels).to(torch.bfloat16)
            b256 = torch.zeros(256, dtype=torch.float32, device=wt.device)
            b256[: self.output_dim] = bs.detach().float()
            self._pred_w256, self._pred_b256, self._pred_key = w256, b256, key
        return self._pred_w256, self._pred_b256

    def forward_decoded(self, x: torch.Tensor, anchor_wh: torch.Tensor, want_objectness: bool = False) -> Dict[str, torch.Tensor]:
        """Inference tail without the raw prediction tensor: conv stack -> mHC over the pixel tokens -> ONE kernel doing the
        1x1 prediction convolution and the decode (hvs_head_decode_fused).  3 anchors x 85."""
        x = self.conv_layers(x)
        b, c, h, w = x.shape
        tok = x.permute(0, 2, 3, 1).reshape(-1, c)
        if not isinstance(self.mhc_enhance, nn.Identity):
            tok = self.mhc_enhance(tok)
        w256, b256 = self._pred_operands()
        return ops.head_decode_fused(tok.to(torch.bfloat16).contiguous(), w256, b256, anchor_wh, b, h, w, want_objectness)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.conv_layers(x)
        b, c, h, w = x.shape
        if not isinstance(self.mhc_enhance, nn.Identity):
            # [B*H*W, C] token view: free (no copy) when x is channels_last, as in hybrid_vision.mhc_over_pixels
            x = self.mhc_enhance(x.permute(0, 2, 3, 1).reshape(-1, c)).reshape(b, h, w, c).permute(0, 3, 1, 2)
        pred = self.pred_conv(x)
        if not pred.is_contiguous():
            # channels_last conv output: one 0.1 ms copy to NCHW so decode takes its vectorised plane-strided mapping
            # (the channel-contiguous mapping works on the view in place but is 3x slower than copy + decode)
            pred = pred.contiguous()
        # [B, A*(5+C), H, W] -> [B,A,H,W,5+C] as a VIEW (NCHW or channels_last alike): the decode kernel reads it through its strides
        return pred.view(b, self.num_anchors, 5 + self.num_classes, h, w).permute(0, 1, 3, 4, 2)


class YOLOLoss(nn.Module):
    """YOLOLoss(num_classes, anchors, image_size, lambda_coord=5, lambda_noobj=.5, lambda_obj=1, lambda_cls=1)
    (yolo_head.py:297-465).  predictions {scale_i: [B,A,H,W,5+C]}, targets [per scale [B,A,H,W,5+C]] -> dict with
    'coord_loss', 'obj_loss', 'noobj_loss', 'cls_loss', 'total_loss'.

    Same arithmetic -- squared error on the 4 box slots and BCE-with-logits on objectness / classes, summed over the
    cells with target objectness > 0.5 (no-object term over the cells < 0.5), every term divided by the scale's object
    count, a scale without objects contributing nothing (:411-413) -- written with masks instead of boolean indexing,
    so the loss has NO host synchronisation (the reference does five .item() per scale) and can be captured in a CUDA
    graph.  The four component entries are detached 0-dim tensors (the reference returns Python floats)."""

    def __init__(self, num_classes: int = 80, anchors=None, image_size: int = 416, lambda_coord: float = 5.0,
                 lambda_noobj: float = 0.5, lambda_obj: float = 1.0, lambda_cls: float = 1.0):
        super().__init__()
        self.num_classes, self.image_size = num_classes, image_size
        self.lambda_coord, self.lambda_noobj, self.lambda_obj, self.lambda_cls = lambda_coord, lambda_noobj, lambda_obj, lambda_cls
        self.anchors = anchors if anchors is not None else DEFAULT_ANCHORS
        self.num_scales = len(self.anchors)

    @staticmethod
    def compute_iou(box1: torch.Tensor, box2: torch.Tensor) -> torch.Tensor:      # :341-372
        ix1, iy1 = torch.max(box1[..., 0], box2[..., 0]), torch.max(box1[..., 1], box2[..., 1])
        ix2, iy2 = torch.min(box1[..., 2], box2[..., 2]), torch.min(box1[..., 3], box2[..., 3])
        inter = torch.clamp(ix2 - ix1, min=0) * torch.clamp(iy2 - iy1, min=0)
        a1 = (box1[..., 2] - box1[..., 0]) * (box1[..., 3] - box1[..., 1])
        a2 = (box2[..., 2] - box2[..., 0]) * (box2[..., 3] - box2[..., 1])
        return inter / (a1 + a2 - inter + 1e-6)

    def forward(self, predictions: Dict[str, torch.Tensor], targets: Sequence[torch.Tensor]) -> Dict[str, torch.Tensor]:
        bce = nn.functional.binary_cross_entropy_with_logits
        total = None
        parts = {"coord_loss": 0.0, "obj_loss": 0.0, "noobj_loss": 0.0, "cls_loss": 0.0}
        for i in range(self.num_scales):
            key = f"scale_{i}"
            if key not in predictions:
                continue
            pred = predictions[key].float()
            tgt = targets[i].to(pred.dtype)
            t_obj = tgt[..., 4]
            obj = (t_obj > 0.5).to(pred.dtype)
            noobj = (t_obj < 0.5).to(pred.dtype)
            n = obj.sum()
            has = (n > 0).to(pred.dtype)
            coord = (((pred[..., :4] - tgt[..., :4]) ** 2).sum(-1) * obj).sum()
            e_obj = bce(pred[..., 4], t_obj, reduction="none")
            l_obj, l_noobj = (e_obj * obj).sum(), (e_obj * noobj).sum()
            l_cls = (bce(pred[..., 5:], tgt[..., 5:], reduction="none").sum(-1) * obj).sum()
            term = (self.lambda_coord * coord + self.lambda_obj * l_obj + self.lambda_noobj * l_noobj + self.lambda_cls * l_cls)
            term = has * term / n.clamp_min(1.0)
            total = term if total is None else total + term
            for k, v in (("coord_loss", coord), ("obj_loss", l_obj), ("noobj_loss", l_noobj), ("cls_loss", l_cls)):
                parts[k] = parts[k] + (has * v).detach()
        out: Dict[str, Any] = dict(parts)
        out["total_loss"] = total if total is not None else 0.0
        return out


def dense_targets_from_boxes(boxes: Sequence[torch.Tensor], labels: Sequence[torch.Tensor], grid_sizes: Sequence[Tuple[int, int]],
                             num_classes: int = 80, anchors=None, device=None) -> List[torch.Tensor]:
    """Padded-box annotations -> the dense per-scale targets [B,A,H,W,5+C] YOLOLoss consumes.  The reference has no
    assigner anywhere; this is the rule SURVEY section 8(d) cfg 4 fixes: a box (cx,cy,w,h
    normalised) goes to the cell floor(cx*W), floor(cy*H) of EVERY scale, on the anchor of that scale with the best
    width/height IoU against anchors/416; slots 0:4 = (cx,cy,w,h), 4 = 1, 5+class = 1."""
    anchors = anchors if anchors is not None else DEFAULT_ANCHORS
    b = len(boxes)
    out = []
    for s, (h, w) in enumerate(grid_sizes):
        a_wh = torch.tensor(anchors[s], dtype=torch.float32) / 416.0
        t = torch.zeros(b, len(anchors[s]), h, w, 5 + num_classes)
        for i in range(b):
            bx = boxes[i].detach().float().cpu().reshape(-1, 4)
            lb = labels[i].detach().cpu().reshape(-1).long()
            for (cx, cy, bw, bh), c in zip(bx.tolist(), lb.tolist()):
                inter = torch.minimum(a_wh[:, 0], torch.tensor(bw)) * torch.minimum(a_wh[:, 1], torch.tensor(bh))
                a = int(torch.argmax(inter / (a_wh[:, 0] * a_wh[:, 1] + bw * bh - inter)))
                gx, gy = min(int(cx * w), w - 1), min(int(cy * h), h - 1)
                t[i, a, gy, gx, 0:4] = torch.tensor([cx, cy, bw, bh])
                t[i, a, gy, gx, 4] = 1.0
                t[i, a, gy, gx, 5 + c] = 1.0
        out.append(t.to(device) if device is not None else t)
    return out


class YOLODetectionHead(nn.Module):
    """YOLODetectionHead(in_channels_list, num_classes=80, anchors=None, use_mhc=True)  (yolo_head.py:468-755)."""

    def __init__(self, in_channels_list: List[int], num_classes: int = 80, anchors=None, use_mhc: bool = True):
        super().__init__()
        self.in_channels_list = in_channels_list
        self.num_classes = num_classes
        self.num_scales = len(in_channels_list)
        self.anchor_generator = YOLOAnchorGenerator(anchors)
        self.num_anchors = self.anchor_generator.get_num_anchors()
        self.pred_heads = nn.ModuleList([YOLOPredictionHead(c, num_classes, self.num_anchors, use_mhc) for c in in_channels_list])
        self.decoder = YOLODecoder(image_size=416)
        self.loss_fn = YOLOLoss(num_classes=num_classes, anchors=anchors)        # :505-508
        self.grid_sizes = [(13, 13), (26, 26), (52, 52)]                          # :511 (informational: the decoder uses the actual H, W)
        self.want_scores = True      # full [B,A,H,W,C] obj*cls tensor (the reference's 'scores'); post_process does not need it
        self.fuse_pred_decode = False  # inference only: prediction conv + decode in one kernel, 'predictions' left empty

    def forward(self, features: Dict[str, torch.Tensor], targets=None, compute_loss: bool = False,
                want_scores: Optional[bool] = None) -> Dict[str, Any]:
        """yolo_head.py:515-569.  With compute_loss and targets the result carries 'loss' (:556-563)."""
        want_scores = self.want_scores if want_scores is None else want_scores
        predictions, decoded = {}, {}
        fused = (self.fuse_pred_decode and not compute_loss and not torch.is_grad_enabled() and not self.training
                 and self.num_anchors == 3 and self.num_classes == 80)
        for i in range(self.num_scales):
            key = ["scale_small", "scale_medium", "scale_large"][i]
            if key not in features:
                continue
            if fused and features[key].is_cuda:
                decoded[f"scale_{i}"] = self.pred_heads[i].forward_decoded(features[key], self.anchor_generator(i).reshape(3, -1, 4)[:, 0, 2:4])
                continue
            pred = self.pred_heads[i](features[key])
            predictions[f"scale_{i}"] = pred
        pending = [k for k in predictions if k not in decoded]
        with torch.no_grad():                           # decode is an inference product; the loss works on raw predictions
            if (not want_scores and len(pending) > 1 and all(predictions[k].is_cuda for k in pending)
                    and len({(predictions[k].dtype, predictions[k].shape[0], predictions[k].shape[4]) for k in pending}) == 1):
                # the decode loop of yolo_head.py:536-555 as one launch (the coarse grids are a fraction of a wave each)
                preds = [predictions[k].detach() for k in pending]
                awhs = [self.anchor_generator(int(k[6:])).reshape(p.shape[1], -1, 4)[:, 0, 2:4] for k, p in zip(pending, preds)]
                for k, p, d in zip(pending, preds, ops.yolo_decode_scales(preds, awhs, want_objectness=True)):
                    d["raw_predictions"] = p
                    decoded[k] = d
            else:
                for k in pending:
                    pred = predictions[k]
                    decoded[k] = self.decoder(pred.detach(), self.anchor_generator(int(k[6:])), pred.shape[2:4], want_scores=want_scores)
        decoded = {k: decoded[k] for k in sorted(decoded)}
        out = {"predictions": predictions, "decoded": decoded}
        if compute_loss and targets is not None:
            out["loss"] = self.loss_fn(predictions, targets)
        return out

    def post_process(self, decoded_outputs: Dict[str, Dict[str, torch.Tensor]], confidence_threshold: float = 0.5,
                     iou_threshold: float = 0.5, max_detections: int = 100) -> List[Dict[str, torch.Tensor]]:
        """yolo_head.py:571-676: per-scale thresholded NMS, concatenation, second NMS -- one batched call."""
        scales = list(decoded_outputs.values())
        db, ds, dl, dc = ops.post_process(scales, confidence_threshold, iou_threshold, max_detections)
        counts = dc.cpu().tolist()                        # the only host sync: the result sizes
        return [{"boxes": db[b, :k], "scores": ds[b, :k], "labels": dl[b, :k]} for b, k in enumerate(counts)]

    def non_max_suppression(self, boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float = 0.5,
                            max_detections: int = 100) -> torch.Tensor:
        """yolo_head.py:678-731; returns int64 indices into the input, descending score."""
        if boxes.numel() == 0:
            return torch.tensor([], dtype=torch.long, device=boxes.device)
        _, keep_src, cnt = ops.nms(boxes, scores, None, iou_threshold, max_detections)
        return keep_src[0, :int(cnt[0])]              # indices into the INPUT (differs from the compacted index when a score is NaN / -inf)

    def compute_iou(self, box1: torch.Tensor, box2: torch.Tensor) -> torch.Tensor:
        """yolo_head.py:733-755 (elementwise torch ops; not on the hot path)."""
        ix1, iy1 = torch.max(box1[..., 0], box2[..., 0]), torch.max(box1[..., 1], box2[..., 1])
        ix2, iy2 = torch.min(box1[..., 2], box2[..., 2]), torch.min(box1[..., 3], box2[..., 3])
        inter = torch.clamp(ix2 - ix1, min=0) * torch.clamp(iy2 - iy1, min=0)
        a1 = (box1[..., 2] - box1[..., 0]) * (box1[..., 3] - box1[..., 1])
        a2 = (box2[..., 2] - box2[..., 0]) * (box2[..., 3] - box2[..., 1])
        return inter / (a1 + a2 - inter + 1e-6)


@dataclass
class PostprocessingConfig:
    """The NMS fields of src/inference/postprocessing.py:31-67."""
    nms_iou_threshold: float = 0.45
    nms_score_threshold: float = 0.25
    nms_max_detections: int = 100
    nms_method: str = "standard"


class NMSFilter:
    """NMSFilter(config).apply(boxes [N,4] cx,cy,w,h, scores [N], class_ids [N]) -> keep indices
    (postprocessing.py:498-607, "standard" method)."""

    def __init__(self, config: Optional[PostprocessingConfig] = None):
        self.config = config or PostprocessingConfig()
        if self.config.nms_method != "standard":
            raise ValueError(f"Unknown / unsupported NMS method: {self.config.nms_method}")

    def apply(self, boxes: torch.Tensor, scores: torch.Tensor, class_ids: torch.Tensor) -> torch.Tensor:
        if len(boxes) == 0:
            return torch.empty(0, dtype=torch.long, device=boxes.device)
        _, keep_src, cnt = ops.nms(boxes, scores, class_ids, self.config.nms_iou_threshold,
                                   self.config.nms_max_detections, class_aware=True, boxes_xyxy=False)
        return keep_src[0, :int(cnt[0])]
