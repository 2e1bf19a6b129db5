"""Measurement harnesses for the whole-model configurations of BASELINE.json:

  config 3  batched inference, global batch 64 at 640 x 640, sharded 64/G per GPU, decode + two-stage NMS, no collective
  config 4  bf16 training, 16 images per GPU, DistributedDataParallel gradient all-reduce over NCCL, synthetic
            COCO-shaped boxes written as dense targets, YOLOLoss
  config 5  streaming batch-1 inference under a CUDA graph: p50 / p99 latency, bitwise-identical repeats
  (config 2, K2 reading)  ManifoldHyperConnection(512, expansion_rate=4) on [2^20, 512]: TFLOP/s of the fused token path
  detection tail           decode + two-stage NMS alone at batch 64 (worst case D18 and objectness bias -4)

They re-create, in a few lines each, the role of the reference's InferenceEngine.infer_batch (src/inference/engine.py:
319-387), scripts/train.py:160-222 and scripts/benchmark.py:124-176 (none of which can run as shipped, SURVEY D12/D13/
D15).  Every function times on the device with CUDA events and returns plain dicts; bench.py assembles them.
"""
from __future__ import annotations

import time
from typing import Any, Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib, ops
from .detection import dense_targets_from_boxes
from .hybrid_vision import HybridVisionSystem, to_channels_last
from .mhc import ManifoldHyperConnection

FWD_GFLOP_PER_IMAGE_640 = 698.0          # SURVEY.md section 3.1 (80.5 % inside mHC)


def build_model(device, seed: int = 0, channels_last: bool = True, bf16_activations: bool = True) -> HybridVisionSystem:
    """Random-init HybridVisionSystem (identical on every rank for a given seed)."""
    torch.manual_seed(seed)
    model = HybridVisionSystem({"num_classes": 80, "image_size": 640}).to(device)
    if channels_last:
        to_channels_last(model)
    if bf16_activations:
        for m in model.modules():
            if isinstance(m, ManifoldHyperConnection):
                m.output_dtype = torch.bfloat16
    return model


class FoldedBiasAct(nn.Module):
    """The bias of a BatchNorm-folded convolution and the activation after it as one pass (hvs_bias_act_bf16) -- ATen runs a
    bias-carrying cuDNN convolution as convolution + a broadcast add pass, then the activation as a third kernel.  Built by
    fold_batchnorm_for_inference in place of the activation of a conv -> BatchNorm -> activation stack."""

    def __init__(self, bias: torch.Tensor, act: nn.Module, kind: str):
        super().__init__()
        self.register_buffer("bias", bias.detach().float().clone())
        self.act, self.kind = act, kind

    def forward(self, y: torch.Tensor) -> torch.Tensor:
        if (not torch.is_grad_enabled() and y.is_cuda and y.dtype == torch.bfloat16 and y.dim() == 4 and y.shape[1] % 8 == 0
                and y.is_contiguous(memory_format=torch.channels_last)):
            return ops.bias_act(y, self.bias, self.kind)
        return self.act(y + self.bias.to(y.dtype).view(1, -1, 1, 1))


def _act_kind(m: nn.Module) -> Optional[str]:
    if isinstance(m, nn.ReLU):
        return "relu"
    if isinstance(m, nn.SiLU):
        return "silu"
    if isinstance(m, nn.LeakyReLU) and abs(m.negative_slope - 0.1) < 1e-12:
        return "leaky_relu_0.1"
    return None


def fold_batchnorm_for_inference(model: nn.Module) -> int:
    """Eval-mode conv + BatchNorm pairs of the host model (ConvMHCLayer.conv/.bn, the FPN and head conv stacks) folded into
    one convolution with bias: y = conv(x) * g / sqrt(var + eps) + (b - mean * g / sqrt(var + eps)).  A caller-side
    inference optimisation (the standard torch.nn.utils.fusion recipe), numerically the same function; the BatchNorm
    module is replaced by Identity, so use it on a model that will only run inference.  Returns the number of pairs."""
    from torch.nn.utils.fusion import fuse_conv_bn_eval
    n = 0
    for mod in list(model.modules()):
        if hasattr(mod, "conv") and hasattr(mod, "bn") and isinstance(mod.conv, nn.Conv2d) and isinstance(mod.bn, nn.BatchNorm2d):
            fused = fuse_conv_bn_eval(mod.conv.eval(), mod.bn.eval())
            if hasattr(mod, "folded_bias") and mod.conv.bias is None:
                # ConvMHCLayer: the folded bias stays OUT of the convolution (ATen would add it in a broadcast pass of its own)
                # and is applied together with the activation by hvs_bias_act_bf16
                with torch.no_grad():
                    mod.conv.weight.copy_(fused.weight)
                mod.folded_bias = fused.bias.detach().float().clone()
            else:
                mod.conv = fused
            mod.bn = nn.Identity()
            n += 1
        if isinstance(mod, nn.Sequential):
            kids = list(mod.named_children())
            for i, ((na, a), (nb, b)) in enumerate(zip(kids, kids[1:])):
                if isinstance(a, nn.Conv2d) and isinstance(b, nn.BatchNorm2d):
                    fused = fuse_conv_bn_eval(a.eval(), b.eval())
                    kind = _act_kind(kids[i + 2][1]) if i + 2 < len(kids) else None
                    if kind is not None and fused.bias is not None and fused.out_channels % 8 == 0:
                        # conv -> BatchNorm -> activation: the folded bias leaves the convolution and joins the activation
                        setattr(mod, kids[i + 2][0], FoldedBiasAct(fused.bias, kids[i + 2][1], kind))
                        fused.bias = None
                    setattr(mod, na, fused)
                    setattr(mod, nb, nn.Identity())
                    n += 1
    return n


def cast_weights_for_bf16_inference(model: nn.Module) -> int:
    """Store the weights and biases of the host model's convolutions and linear layers in bf16.  Under bf16 autocast those
    parameters are cast to bf16 on EVERY call (autocast caches a cast only for parameters that require grad inside one
    autocast region, and a CUDA graph replays the cast kernels anyway): 230 small `bfloat16_copy` launches per forward,
    0.8 of the 7.4 ms of a batch-1 frame.  Casting once gives the same bf16 operands, so the outputs do not change.
    The mHC modules are left alone (fp32 master parameters; this library keeps its own bf16 copies of them).  Inference
    only: use it on a model that will not be trained.  Returns the number of tensors cast."""
    from .mhc import ManifoldHyperConnection
    skip = set()
    for mod in model.modules():
        if isinstance(mod, ManifoldHyperConnection):
            skip.update(id(m) for m in mod.modules())
    n = 0
    with torch.no_grad():
        for mod in model.modules():
            if id(mod) in skip or not isinstance(mod, (nn.Conv2d, nn.Linear, nn.ConvTranspose2d, nn.Conv1d)):
                continue
            for name in ("weight", "bias"):
                t = getattr(mod, name, None)
                if isinstance(t, nn.Parameter) and t.dtype == torch.float32:
                    t.data = t.data.to(torch.bfloat16)
                    t.requires_grad_(False)
                    n += 1
    return n


def _time_steps(fn, steps: int, warmup: int) -> float:
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


# ----------------------------------------------------------------------------------------------- config 3
def inference_sharded(model: HybridVisionSystem, device, world: int = 1, rank: int = 0, global_batch: int = 64,
                      image: int = 640, steps: int = 5, warmup: int = 3, objectness_bias: Optional[float] = None,
                      host_input: bool = False, fuse_head: bool = True, use_graph: bool = True) -> Dict[str, Any]:
    """Each rank takes global_batch / world images (strong scaling, no collective): forward under bf16 autocast,
    decode, two-stage NMS (conf 0.25, iou 0.45, max 100).  host_input=True also copies the shard from pinned host memory
    and reads the detections back inside the timed region (the e2e reading)."""
    model.eval()
    per = global_batch // world
    g = torch.Generator(device="cpu").manual_seed(1000 + rank)
    x_host = torch.randn(per, 3, image, image, generator=g).to(torch.bfloat16).pin_memory()
    x_dev = x_host.to(device).contiguous(memory_format=torch.channels_last)
    head = model.detection_head
    saved_bias = None
    if objectness_bias is not None:
        saved_bias = [h.pred_conv.bias.detach().clone() for h in head.pred_heads]
        with torch.no_grad():
            for h in head.pred_heads:
                h.pred_conv.bias.view(head.num_anchors, -1)[:, 4] = objectness_bias
    head.want_scores = False
    head.fuse_pred_decode = fuse_head
    result = {}
    static_x = x_dev.clone()

    def forward_and_nms():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            out = model(static_x)
            return ops.post_process(list(out["decoded"].values()), 0.25, 0.45, 100)

    try:
        launches0 = _lib.launch_count()
        graph = None
        if use_graph:
            # the step is ~1100 kernel launches; at 8 images per GPU (64 / 8) the eager loop is bound by the host's launch
            # rate (~20 us per launch from Python), not by the GPU: capture forward + decode + NMS once, replay per step
            side = torch.cuda.Stream(device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                for _ in range(2):
                    forward_and_nms()
            torch.cuda.current_stream(device).wait_stream(side)
            launches0 = _lib.launch_count()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                dets = forward_and_nms()
            launches = _lib.launch_count() - launches0
            host_out = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in dets]

            def step():
                if host_input:
                    static_x.copy_(x_host.to(device, non_blocking=True).contiguous(memory_format=torch.channels_last))
                graph.replay()
                if host_input:
                    for h, d in zip(host_out, dets):
                        h.copy_(d, non_blocking=True)
                    torch.cuda.current_stream(device).synchronize()
                    result["dets"] = host_out
                else:
                    result["count"] = dets[3]
            ms = _time_steps(step, steps, warmup)
        else:
            def step():
                if host_input:
                    static_x.copy_(x_host.to(device, non_blocking=True).contiguous(memory_format=torch.channels_last))
                boxes, scores, labels, count = forward_and_nms()
                if host_input:
                    result["dets"] = (boxes.cpu(), scores.cpu(), labels.cpu(), count.cpu())
                else:
                    result["count"] = count
            ms = _time_steps(step, steps, warmup)
            launches = (_lib.launch_count() - launches0) // (steps + warmup)
    finally:
        head.want_scores = True
        head.fuse_pred_decode = False
        if saved_bias is not None:
            with torch.no_grad():
                for h, b in zip(head.pred_heads, saved_bias):
                    h.pred_conv.bias.copy_(b)
        graph = None
    kept = float(result["dets"][3].float().mean()) if host_input else float(result["count"].float().mean())
    return {"ms_per_step": ms, "images_per_rank": per, "hvs_launches_per_step": int(launches), "mean_detections": kept, "cuda_graph": bool(use_graph),
            "h2d_bytes_per_step": x_host.numel() * 2 if host_input else 0,
            "d2h_bytes_per_step": per * 100 * (16 + 4 + 8) + per * 4 if host_input else 0}


# ----------------------------------------------------------------------------------------------- config 4
def synthetic_targets(batch: int, image: int, rank: int, device, num_classes: int = 80, boxes_per_image: int = 8) -> List[torch.Tensor]:
    """8 boxes per image, cx, cy ~ U(0.1, 0.9), w, h ~ U(0.05, 0.4), class ~ U{0..79}, seed 1234 + rank (SURVEY 8(d) cfg 4)."""
    g = torch.Generator().manual_seed(1234 + rank)
    boxes, labels = [], []
    for _ in range(batch):
        c = 0.1 + 0.8 * torch.rand(boxes_per_image, 2, generator=g)
        wh = 0.05 + 0.35 * torch.rand(boxes_per_image, 2, generator=g)
        boxes.append(torch.cat([c, wh], 1))
        labels.append(torch.randint(0, num_classes, (boxes_per_image,), generator=g))
    grids = [(image // 8,) * 2, (image // 16,) * 2, (image // 32,) * 2]
    return dense_targets_from_boxes(boxes, labels, grids, num_classes, device=device)


def training_ddp(model: HybridVisionSystem, device, world: int = 1, rank: int = 0, batch_per_gpu: int = 16, image: int = 640,
                 steps: int = 3, warmup: int = 2, use_graph: Optional[bool] = None) -> Dict[str, Any]:
    """One optimisation step = forward (bf16 autocast) + YOLOLoss + backward (DDP all-reduces the 353.8 M fp32 gradients
    in 25 MB buckets, overlapped with the backward) + AdamW.  Weak scaling: batch_per_gpu images on every rank.
    use_graph: the whole step (forward, loss, backward, optimizer; ~5000 launches) captured once into a CUDA graph and
    replayed -- the step is otherwise bound by the host's launch rate.  Default: on; the DDP step
    follows the recipe for DistributedDataParallel under capture (DDP built on the side stream, 11 eager warm-up steps so its
    buckets are final, NCCL's watchdog-side error handling off: bench.py sets TORCH_NCCL_ASYNC_ERROR_HANDLING=0)."""
    import torch.distributed as dist
    from .mhc import dropout_step_counter
    model.train()
    net: nn.Module = model
    if use_graph is None:
        use_graph = True
    side = torch.cuda.Stream(device)
    if world > 1:
        # (for a captured step DDP has to be built on the side stream the warm-up and the capture run on)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            net = nn.parallel.DistributedDataParallel(model, device_ids=[device.index], gradient_as_bucket_view=True)
        torch.cuda.current_stream(device).wait_stream(side)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-5, fused=True, capturable=bool(use_graph))
    used = batch_per_gpu
    while True:
        try:
            g = torch.Generator(device="cpu").manual_seed(2000 + rank)
            x = torch.randn(used, 3, image, image, generator=g).to(device).contiguous(memory_format=torch.channels_last)
            targets = synthetic_targets(used, image, rank, device)
            losses = []
            step_counter = dropout_step_counter(device)

            def eager_step():
                step_counter.add_(1)                   # new dropout masks every step (also when the step is a graph replay)
                opt.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    out = net(x, targets=targets, compute_loss=True)
                    # final_fusion / output_projection get no gradient from the detection loss: a zero-weight
                    # term keeps every parameter in the graph so DDP needs no find_unused_parameters pass
                    loss = out["loss"]["total_loss"] + 0.0 * out["final_features"].float().sum()
                loss.backward()
                opt.step()
                return loss.detach()

            graph = None
            if use_graph:
                side.wait_stream(torch.cuda.current_stream(device))
                with torch.cuda.stream(side):
                    for _ in range(11 if world > 1 else 3):   # allocator, cuDNN plans, optimizer state, DDP buckets: settled before the capture
                        eager_step()
                torch.cuda.current_stream(device).wait_stream(side)
                torch.cuda.synchronize(device)
                launches0 = _lib.launch_count()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):       # the stream DDP and the warm-up ran on (AccumulateGrad nodes live there)
                    static_loss = eager_step()
                launches = _lib.launch_count() - launches0

                def step():
                    graph.replay()
                    losses.append(static_loss.clone())
                ms = _time_steps(step, steps, warmup)
            else:
                def step():
                    losses.append(eager_step())
                launches0 = _lib.launch_count()
                ms = _time_steps(step, steps, warmup)
                launches = (_lib.launch_count() - launches0) // (steps + warmup)
            break
        except torch.OutOfMemoryError:
            del x
            torch.cuda.empty_cache()
            if used <= 1:
                raise
            used //= 2
            if world > 1:
                raise                                   # ranks must agree on the batch: do not shrink unilaterally
    vals = torch.stack(losses).float().cpu()
    grad_bytes = sum(p.numel() for p in model.parameters()) * 4
    model.eval()
    return {"ms_per_step": ms, "batch_per_gpu": used, "loss_first": float(vals[0]), "loss_last": float(vals[-1]),
            "finite": bool(torch.isfinite(vals).all()), "grad_allreduce_bytes": grad_bytes if world > 1 else 0,
            "hvs_launches_per_step": int(launches), "peak_mem_gb": torch.cuda.max_memory_allocated(device) / 2 ** 30,
            "cuda_graph": bool(use_graph)}


# ----------------------------------------------------------------------------------------------- config 5
def streaming_latency(model: HybridVisionSystem, device, frames: int = 300, image: int = 640) -> Dict[str, Any]:
    """Batch-1 frames through ONE CUDA graph (forward + decode + two-stage NMS): per-frame latency from the frame being
    resident in HBM to the detections being on the host (H2D of the frame excluded, stated), p50 / p99, and a bitwise
    check that repeated frames give identical detections."""
    model.eval()
    head = model.detection_head
    head.want_scores = False
    head.fuse_pred_decode = True
    try:
        g = torch.Generator(device="cpu").manual_seed(3000)
        pool = [torch.randn(1, 3, image, image, generator=g).to(torch.bfloat16).to(device).contiguous(memory_format=torch.channels_last)
                for _ in range(8)]
        static_x = pool[0].clone()

        def run():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                out = model(static_x)
                return ops.post_process(list(out["decoded"].values()), 0.25, 0.45, 100)

        side = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(3):
                run()
        torch.cuda.current_stream(device).wait_stream(side)
        launches0 = _lib.launch_count()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            dets = run()
        launches = _lib.launch_count() - launches0
        host = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in dets]
        lat = []
        first: Dict[int, List[torch.Tensor]] = {}
        identical = True
        for f in range(frames + 10):
            static_x.copy_(pool[f % len(pool)])
            torch.cuda.synchronize(device)
            t0 = time.perf_counter()
            graph.replay()
            for h, d in zip(host, dets):
                h.copy_(d, non_blocking=True)
            torch.cuda.synchronize(device)
            dt = (time.perf_counter() - t0) * 1e3
            if f >= 10:
                lat.append(dt)
            snap = [h.clone() for h in host]
            k = f % len(pool)
            if k in first:
                identical = identical and all(torch.equal(a, b) for a, b in zip(first[k], snap))
            else:
                first[k] = snap
        lat.sort()
        return {"frames": frames, "p50_ms": lat[len(lat) // 2], "p99_ms": lat[min(len(lat) - 1, int(len(lat) * 0.99))],
                "mean_ms": sum(lat) / len(lat), "bitwise_identical_repeats": bool(identical), "hvs_kernels_in_graph": int(launches),
                "latency_scope": "frame resident in HBM -> detections on host (CUDA graph replay + D2H of [100,4]+[100]+[100]+count); H2D of the frame excluded"}
    finally:
        head.want_scores = True
        head.fuse_pred_decode = False


# ----------------------------------------------------------------------------------------------- K2 microbenchmark
def k2_microbench(device, tokens: int = 1 << 20, dim: int = 512, expansion: int = 4, steps: int = 5, warmup: int = 3) -> Dict[str, Any]:
    """ManifoldHyperConnection(512, expansion_rate=4) on [2^20, 512] bf16, eval: the fused kernel path against the same
    module's library path (torch ops under bf16 autocast = the reference's own CUDA execution)."""
    torch.manual_seed(0)
    mod = ManifoldHyperConnection(dim, expansion_rate=expansion).to(device).eval()
    with torch.no_grad():                               # trained-like coefficients: at the xavier(0.1) init the module is so
        for p in (mod.H_pre_raw, mod.H_post_raw, mod.H_res_raw):   # ill-conditioned under bf16 that two correct paths differ by O(1)
            p.normal_(0, 1.0)                           # (tests/test_gpu_k2.py); timing does not depend on the values
    mod.output_dtype = torch.bfloat16
    h = mod.hidden_dim
    flop_per_token = 2.0 * (2 * dim * h + 4 * h * h + dim * dim)
    x = torch.randn(tokens, dim, device=device, dtype=torch.bfloat16)
    lib = _lib.load()

    def fused():
        with torch.no_grad():
            mod(x)

    def eager():
        with torch.no_grad():
            mod.forward_library(x)

    fused()
    lib.hvs_mhc_stream_profile(1)
    ms_fused = _time_steps(fused, steps, warmup)
    import ctypes
    buf = (ctypes.c_float * 8)()
    lib.hvs_profile_kernel_ms(buf)
    lib.hvs_mhc_stream_profile(0)
    ms_eager = _time_steps(eager, max(2, steps // 2), 2)
    with torch.no_grad():
        a, b = mod(x[:4096]).float(), mod.forward_library(x[:4096]).float()
    return {"workload": f"ManifoldHyperConnection({dim}, expansion_rate={expansion}) eval forward, x [{tokens}, {dim}] bf16",
            "flop_per_token": flop_per_token, "ms_fused": ms_fused, "ms_library_path": ms_eager,
            "tflops_fused": flop_per_token * tokens / (ms_fused * 1e-3) / 1e12,
            "tflops_library_path": flop_per_token * tokens / (ms_eager * 1e-3) / 1e12,
            "tokens_per_s_fused": tokens / (ms_fused * 1e-3), "gemm_kernel_mean_ms": float(buf[4]),
            "max_abs_diff_vs_library_path_4096_tokens": float((a - b).abs().max())}


# ----------------------------------------------------------------------------------------------- detection tail
def detect_tail(device, batch: int = 64, objectness_bias: float = 0.0, steps: int = 10, warmup: int = 3) -> Dict[str, Any]:
    """decode (reading the head's permuted NCHW view in place) + two-stage NMS at 640 x 640 grids, bf16 predictions."""
    g = torch.Generator(device=device).manual_seed(0)
    preds, awh = [], []
    anchors = torch.tensor([[(10, 13), (16, 30), (33, 23)], [(30, 61), (62, 45), (59, 119)], [(116, 90), (156, 198), (373, 326)]],
                           dtype=torch.float32, device=device) / 416.0
    for s, hw in enumerate((80, 40, 20)):
        nchw = (torch.randn(batch, 3 * 85, hw, hw, generator=g, device=device) * 0.5)
        v = nchw.view(batch, 3, 85, hw, hw)
        v[:, :, 4] += objectness_bias
        preds.append(v.to(torch.bfloat16).permute(0, 1, 3, 4, 2))
        awh.append(anchors[s])
    out = {}

    def step():
        dec = ops.yolo_decode_scales(preds, awh, want_objectness=False)        # the head's decode loop: one launch
        out["r"] = ops.post_process(dec, 0.25, 0.45, 100)
        out["dec"] = dec

    def decode_only():
        out["dec"] = ops.yolo_decode_scales(preds, awh, want_objectness=False)

    ms = _time_steps(step, steps, warmup)
    ms_dec = _time_steps(decode_only, steps, warmup)

    def graphed(fn):                                       # the same launches replayed from one CUDA graph (no host launch gaps)
        side = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream(device).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            fn()
        return _time_steps(graph.replay, steps, warmup)

    ms_graph = graphed(step)
    ms_dec_graph = graphed(decode_only)
    cand = sum(int((d["class_scores"] > 0.25).sum()) for d in out["dec"]) / batch
    in_bytes = sum(p.numel() * 2 for p in preds)
    out_bytes = sum(d["boxes"].numel() * 4 + d["class_scores"].numel() * 4 + d["class_indices"].numel() * 8 for d in out["dec"])
    return {"workload": f"decode + two-stage NMS, batch {batch}, grids 80/40/20, 80 classes, bf16 predictions, conf 0.25 iou 0.45 max 100, objectness bias {objectness_bias}",
            "ms_per_batch": ms, "img_per_s": batch / ms * 1e3, "decode_ms": ms_dec, "nms_ms": ms - ms_dec,
            "graph_replay": {"ms_per_batch": ms_graph, "img_per_s": batch / ms_graph * 1e3, "decode_ms": ms_dec_graph,
                             "nms_ms": ms_graph - ms_dec_graph,
                             "note": "the same launches as one CUDA graph replay (how hybrid_vision runs them); the eager "
                                     "figures above include the host's launch gaps between the decode launch and the 4 NMS launches"},
            "decode_GBps": (in_bytes + out_bytes) / (ms_dec_graph * 1e-3) / 1e9, "decode_bytes": in_bytes + out_bytes,
            "candidates_over_threshold_per_image": cand, "mean_detections": float(out["r"][3].float().mean())}
