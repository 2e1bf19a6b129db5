"""Host model the hot-path modules drop into: the reference's ``HybridVisionSystem`` (detection task) rebuilt around
``hvs_b200.ManifoldHyperConnection`` / ``RMSNorm`` / ``YOLODetectionHead``.

This is the CALLER of the path (SURVEY.md section 8 rows a10 / "next"), not part of it: convolutions, batch norm,
attention matmuls stay library calls (cuDNN / cuBLAS through torch).  It exists because BASELINE configs 1, 3, 4, 5 time
the whole model and ``/root/reference`` does not travel to the GPU box.  Module / parameter names follow the reference
so that its checkpoints load (state_dict keys are pinned by tests/golden/hybrid_vision_keys.json, generated from the
reference itself), and the forward follows

    src/models/hybrid_vision.py:222-402       HybridVisionSystem.forward / _extract_final_features / detect :404-437
    src/models/vision_backbone.py:99-134, :329-397
    src/models/vit_encoder_decoder.py:57-76 (patch embedding), :174-210 (block), :470-520 (hybrid encoder)
    src/models/manifold_layers.py:386-434     MultiHeadManifoldAttention
    src/models/feature_fusion.py:82-157       FeaturePyramidNetwork

with the repair set of SURVEY Appendix A built in (the reference cannot run without it): R3 mHC is fed channels-last
tokens in the FPN and the hybrid encoder, R4 the ViT position embedding is interpolated to the token count, R8 the output
projection skips its pool / flatten on the already pooled vector; R2/R5/R7 live in ``detection.py``.

Layout fold (row a10): every mHC hop is ``x.permute(0, 2, 3, 1).reshape(-1, C)`` -> mHC -> back.  With the model in
``torch.channels_last`` that permute is a free view in both directions, so the reference's two full-tensor copies per
call disappear without the mHC kernels having to address NCHW.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .detection import YOLODetectionHead
from .mhc import (ManifoldHyperConnection, RMSNorm, batched_training_coefficients, clear_training_coefficients,
                  refresh_static_coefficients)


def mhc_over_pixels(mhc: nn.Module, x: torch.Tensor) -> torch.Tensor:
    """NCHW feature map -> [B*H*W, C] tokens -> mHC -> NCHW (vision_backbone.py:117-123).  Free views when x is
    channels_last; the result is again channels_last."""
    b, c, h, w = x.shape
    t = x.permute(0, 2, 3, 1)
    y = mhc(t.reshape(-1, c))
    return y.reshape(b, h, w, c).permute(0, 3, 1, 2)


def to_channels_last(model: nn.Module) -> nn.Module:
    """Put every convolution weight in channels_last (nn.Module.to(memory_format=...) would also try the 5-D anchor
    buffer).  Activations follow the input's format through cuDNN, BatchNorm and the elementwise ops."""
    for m in model.modules():
        if isinstance(m, nn.Conv2d):
            m.weight.data = m.weight.data.contiguous(memory_format=torch.channels_last)
    return model


def interpolate_positions(owner: nn.Module, pos: torch.Tensor, size: int) -> torch.Tensor:
    """Position embeddings [1, P, E] linearly resampled to `size` tokens (repair R4).  The result depends on the parameter and
    the token count only, so inference keeps it (keyed on the parameter's version counter: an optimizer step or a checkpoint
    load invalidates it) instead of launching the resampling on every frame."""
    if pos.shape[1] == size:
        return pos
    if torch.is_grad_enabled() and pos.requires_grad:
        return F.interpolate(pos.transpose(1, 2), size=(size,), mode="linear").transpose(1, 2)
    key = (size, pos._version, pos.dtype, pos.device, pos.data_ptr())
    cache = owner.__dict__.setdefault("_pos_cache", {})
    if cache.get("key") != key:
        with torch.no_grad():
            cache["value"] = F.interpolate(pos.detach().transpose(1, 2), size=(size,), mode="linear").transpose(1, 2).contiguous()
        cache["key"] = key
    return cache["value"]


class ConvMHCLayer(nn.Module):
    """conv -> BN -> act -> mHC over pixels -> squeeze-excite gate -> (+ identity)   (vision_backbone.py:10-134)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 3, stride: int = 1, padding: Optional[int] = None,
                 groups: int = 1, expansion_rate: int = 4, use_mhc: bool = True, activation: str = "silu"):
        super().__init__()
        pad = kernel_size // 2 if padding is None else padding
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, pad, groups=groups, bias=False)
        self.bn = nn.BatchNorm2d(out_channels)
        self.activation = {"silu": nn.SiLU, "relu": nn.ReLU, "gelu": nn.GELU}[activation]()
        self.mhc = ManifoldHyperConnection(out_channels, expansion_rate=expansion_rate) if use_mhc else None
        self.use_residual = in_channels == out_channels and stride == 1
        self.activation_name = activation
        self.folded_bias: Optional[torch.Tensor] = None      # set by harness.fold_batchnorm_for_inference (bn becomes Identity)
        self.fuse_se_gate = True                              # inference under bf16: the squeeze-excite gate as one launch
        self._se_ws: Optional[torch.Tensor] = None
        self.channel_attention = None
        if use_mhc and out_channels >= 32:
            self.channel_attention = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(out_channels, out_channels // 4, 1),
                                                   self.activation, nn.Conv2d(out_channels // 4, out_channels, 1), nn.Sigmoid())
        nn.init.kaiming_normal_(self.conv.weight, mode="fan_out", nonlinearity="relu")

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.folded_bias is not None:
            y = self.conv(x)                            # BatchNorm folded into the weights; its bias and the activation in one pass
            if (not torch.is_grad_enabled() and y.is_cuda and y.dtype == torch.bfloat16 and y.shape[1] % 8 == 0
                    and y.is_contiguous(memory_format=torch.channels_last) and self.activation_name in ops.ACTIVATIONS):
                y = ops.bias_act(y, self.folded_bias, self.activation_name)
            else:
                y = self.activation(y + self.folded_bias.to(y.dtype).view(1, -1, 1, 1))
        else:
            y = self.activation(self.bn(self.conv(x)))
        if self.mhc is not None:
            y = mhc_over_pixels(self.mhc, y)
            if self.channel_attention is not None:
                cl = torch.channels_last
                if (self.fuse_se_gate and not torch.is_grad_enabled() and y.is_cuda and y.dtype == torch.bfloat16 and y.shape[1] % 32 == 0
                        and y.shape[1] <= 2048 and y.is_contiguous(memory_format=cl) and self.activation_name in ops.ACTIVATIONS):
                    # inference: pool + both 1x1 convolutions + activation + sigmoid in one launch (hvs_se_gate_bf16)
                    ca = self.channel_attention
                    gate, self._se_ws = ops.se_gate(y, ca[1].weight, ca[1].bias, ca[3].weight, ca[3].bias, self.activation_name, self._se_ws)
                else:
                    gate = self.channel_attention(y)
                if (not torch.is_grad_enabled() and y.is_cuda and y.dtype == torch.bfloat16 and gate.dtype == torch.bfloat16
                        and y.shape[1] % 8 == 0 and y.is_contiguous(memory_format=cl)
                        and (not self.use_residual or (x.dtype == torch.bfloat16 and x.is_contiguous(memory_format=cl)))):
                    # inference: gate multiply and residual add in one pass (hvs_gate_residual_bf16)
                    return ops.gate_residual(y, gate, x if self.use_residual else None)
                y = y * gate
        return y + x if self.use_residual else y


class ResidualMHCLayer(nn.Module):
    """1x1 squeeze -> 3x3 expand -> 1x1 projection, each a ConvMHCLayer, plus identity (vision_backbone.py:137-197)."""

    def __init__(self, channels: int, num_blocks: int = 2, expansion_rate: int = 4, bottleneck: bool = True):
        super().__init__()
        if bottleneck and channels >= 64:
            self.blocks = nn.Sequential(ConvMHCLayer(channels, channels // 2, 1, expansion_rate=expansion_rate),
                                        ConvMHCLayer(channels // 2, channels, 3, expansion_rate=expansion_rate))
            self.projection = ConvMHCLayer(channels, channels, 1, expansion_rate=expansion_rate)
        else:
            self.blocks = nn.Sequential(*[ConvMHCLayer(channels, channels, 3, expansion_rate=expansion_rate) for _ in range(num_blocks)])
            self.projection = nn.Identity()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.projection(self.blocks(x)) + x


class HybridVisionBackbone(nn.Module):
    """Stem (3 conv-mHC layers + pool), four stages, three mHC "enhance" hops (vision_backbone.py:200-397)."""

    def __init__(self, input_channels: int = 3, base_channels: int = 32, num_blocks=(2, 3, 4, 2), use_mhc: bool = True,
                 activation: str = "silu", dropout_rate: float = 0.1):
        super().__init__()
        c = base_channels
        self.use_mhc = use_mhc
        self.stem = nn.Sequential(ConvMHCLayer(input_channels, c, 3, 2, 1, use_mhc=use_mhc, activation=activation),
                                  ConvMHCLayer(c, c, 3, 1, 1, use_mhc=use_mhc, activation=activation),
                                  ConvMHCLayer(c, 2 * c, 3, 1, 1, use_mhc=use_mhc, activation=activation),
                                  nn.MaxPool2d(2, 2))
        widths = [2 * c, 4 * c, 8 * c, 16 * c]
        self.stages = nn.ModuleList()
        cur = 2 * c
        for i, (depth, width) in enumerate(zip(num_blocks, widths)):
            layers: List[nn.Module] = [ConvMHCLayer(cur, width, 3, 2 if i > 0 else 1, use_mhc=use_mhc, activation=activation)]
            layers += [ResidualMHCLayer(width, 2, 4, True) for _ in range(1, depth)]
            self.stages.append(nn.Sequential(*layers))
            cur = width
        mk = (lambda d: ManifoldHyperConnection(d, expansion_rate=4)) if use_mhc else (lambda d: nn.Identity())
        self.enhance_large, self.enhance_medium, self.enhance_small = mk(widths[3]), mk(widths[2]), mk(widths[1])
        self.dropout = nn.Dropout2d(dropout_rate) if dropout_rate > 0 else nn.Identity()
        self.output_channels = {"stem": 2 * c, "stage_1": widths[0], "stage_2": widths[1], "stage_3": widths[2], "stage_4": widths[3]}

    def get_output_channels(self) -> Dict[str, int]:
        oc = self.output_channels
        return {"scale_small": oc["stage_2"], "scale_medium": oc["stage_3"], "scale_large": oc["stage_4"]}

    def forward(self, x: torch.Tensor) -> Dict[str, Any]:
        raw = {"stem": self.stem(x)}
        y = raw["stem"]
        for i, stage in enumerate(self.stages):
            y = stage(y)
            raw[f"stage_{i + 1}"] = y
        out: Dict[str, Any] = {}
        for name, key, enh in (("scale_small", "stage_2", self.enhance_small), ("scale_medium", "stage_3", self.enhance_medium),
                               ("scale_large", "stage_4", self.enhance_large)):
            f = raw[key]
            if self.use_mhc:
                f = mhc_over_pixels(enh, f)
            out[name] = self.dropout(f)
        out["raw_features"] = raw
        return out


class MultiHeadManifoldAttention(nn.Module):
    """softmax(q k^T / sqrt(d)) v with q / k / v / out projections that are mHC modules (manifold_layers.py:349-434)."""

    def __init__(self, embed_dim: int, num_heads: int = 8, dropout: float = 0.1, use_mhc: bool = True):
        super().__init__()
        assert embed_dim % num_heads == 0
        self.embed_dim, self.num_heads, self.head_dim = embed_dim, num_heads, embed_dim // num_heads
        mk = (lambda: ManifoldHyperConnection(embed_dim, expansion_rate=2)) if use_mhc else (lambda: nn.Linear(embed_dim, embed_dim))
        self.q_proj, self.k_proj, self.v_proj, self.out_proj = mk(), mk(), mk(), mk()
        self.dropout = nn.Dropout(dropout)
        self.scaling = self.head_dim ** -0.5
        self.use_sdpa = True       # bf16 autocast inference: the library's fused attention instead of the five-op form below

    def forward(self, query, key, value, key_padding_mask=None, need_weights: bool = False):
        b, n, e = query.shape
        split = lambda t: t.reshape(b, -1, self.num_heads, self.head_dim).transpose(1, 2)
        q, k, v = split(self.q_proj(query)), split(self.k_proj(key)), split(self.v_proj(value))
        if (self.use_sdpa and not need_weights and not torch.is_grad_enabled() and not self.training and q.is_cuda
                and torch.is_autocast_enabled()):
            # the same softmax(q k^T / sqrt(d)) v (a caller of the path, manifold_layers.py:410-424) without the [B, heads, N, N]
            # score tensor and its fp32 <-> bf16 round trips in HBM: at batch 64 those were 5 ms of a 77 ms forward
            dt = torch.get_autocast_dtype("cuda")
            mask = None if key_padding_mask is None else ~key_padding_mask[:, None, None, :]
            o = F.scaled_dot_product_attention(q.to(dt), k.to(dt), v.to(dt), attn_mask=mask, scale=self.scaling)
            o = o.transpose(1, 2).reshape(b, n, e)
            return self.out_proj(o), None
        w = torch.matmul(q, k.transpose(-2, -1)) * self.scaling
        if key_padding_mask is not None:
            w = w.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
        w = self.dropout(F.softmax(w, dim=-1))
        o = torch.matmul(w, v).transpose(1, 2).reshape(b, n, e)
        o = self.out_proj(o)
        return (o, w) if need_weights else (o, None)


class PatchEmbedding(nn.Module):
    """Patch projection -> mHC -> [cls | patches] + position embedding -> norm (vit_encoder_decoder.py:11-76; R4:
    the position table is resampled linearly to the actual token count)."""

    def __init__(self, image_size: int = 224, patch_size: int = 16, in_channels: int = 3, embed_dim: int = 768, use_mhc: bool = True):
        super().__init__()
        self.num_patches = (image_size // patch_size) ** 2
        self.projection = nn.Conv2d(in_channels, embed_dim, patch_size, patch_size)
        self.mhc_enhance = ManifoldHyperConnection(embed_dim, expansion_rate=2) if use_mhc else nn.Identity()
        self.position_embeddings = nn.Parameter(torch.zeros(1, self.num_patches + 1, embed_dim))
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.norm = RMSNorm(embed_dim) if use_mhc else nn.LayerNorm(embed_dim)
        nn.init.trunc_normal_(self.position_embeddings, std=0.02)
        nn.init.trunc_normal_(self.cls_token, std=0.02)
        nn.init.xavier_uniform_(self.projection.weight)
        nn.init.zeros_(self.projection.bias)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        tok = self.projection(x).flatten(2).transpose(1, 2)
        tok = self.mhc_enhance(tok)
        tok = torch.cat([self.cls_token.expand(tok.shape[0], -1, -1).to(tok.dtype), tok], dim=1)
        pos = interpolate_positions(self, self.position_embeddings, tok.shape[1])
        return self.norm(tok + pos)


class TransformerEncoderBlock(nn.Module):
    """x + mHC(attn(norm1 x));  x + mHC(mlp(norm2 x))   (vit_encoder_decoder.py:79-210)."""

    def __init__(self, embed_dim: int = 768, num_heads: int = 8, mlp_ratio: float = 4.0, dropout: float = 0.1, use_mhc: bool = True):
        super().__init__()
        self.attention = MultiHeadManifoldAttention(embed_dim, num_heads, dropout, use_mhc)
        norm = (lambda: RMSNorm(embed_dim)) if use_mhc else (lambda: nn.LayerNorm(embed_dim))
        self.norm1 = norm()
        hidden = int(embed_dim * mlp_ratio)
        self.mlp = nn.Sequential(nn.Linear(embed_dim, hidden), nn.GELU(), nn.Dropout(dropout), nn.Linear(hidden, embed_dim), nn.Dropout(dropout))
        self.norm2 = norm()
        mk = (lambda: ManifoldHyperConnection(embed_dim, expansion_rate=2)) if use_mhc else (lambda: nn.Identity())
        self.residual_mhc1, self.residual_mhc2 = mk(), mk()
        self.dropout = nn.Dropout(dropout)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        h = self.norm1(x)
        a, _ = self.attention(h, h, h)
        x = x + self.dropout(self.residual_mhc1(a))
        m = self.mlp(self.norm2(x))
        return x + self.dropout(self.residual_mhc2(m))


class VisionTransformerEncoder(nn.Module):
    """vit_encoder_decoder.py:213-315."""

    def __init__(self, image_size: int = 224, patch_size: int = 16, in_channels: int = 3, embed_dim: int = 768, depth: int = 12,
                 num_heads: int = 12, mlp_ratio: float = 4.0, dropout: float = 0.1, use_mhc: bool = True, num_classes: int = 1000):
        super().__init__()
        self.patch_embed = PatchEmbedding(image_size, patch_size, in_channels, embed_dim, use_mhc)
        self.blocks = nn.ModuleList([TransformerEncoderBlock(embed_dim, num_heads, mlp_ratio, dropout, use_mhc) for _ in range(depth)])
        self.norm = RMSNorm(embed_dim) if use_mhc else nn.LayerNorm(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        t = self.patch_embed(x)
        for blk in self.blocks:
            t = blk(t)
        return self.head(self.norm(t)[:, 0])


class HybridVisionEncoder(nn.Module):
    """1x1 conv to the ViT width, + position table, ViT, class token broadcast back over the map, 1x1 conv, + input,
    mHC over pixels (vit_encoder_decoder.py:400-520; R3)."""

    def __init__(self, cnn_channels: int = 512, vit_embed_dim: int = 256, vit_depth: int = 6, vit_num_heads: int = 8, use_mhc: bool = True):
        super().__init__()
        self.cnn_to_vit = nn.Conv2d(cnn_channels, vit_embed_dim, 1)
        self.pos_embed = nn.Parameter(torch.zeros(1, 256, vit_embed_dim))
        self.vit_encoder = VisionTransformerEncoder(16, 1, vit_embed_dim, vit_embed_dim, vit_depth, vit_num_heads, 4.0, 0.1, use_mhc, 0)
        self.vit_to_cnn = nn.Conv2d(vit_embed_dim, cnn_channels, 1)
        self.fusion_mhc = ManifoldHyperConnection(cnn_channels, expansion_rate=2) if use_mhc else nn.Identity()
        nn.init.trunc_normal_(self.pos_embed, std=0.02)

    def forward(self, feat: torch.Tensor) -> torch.Tensor:
        b, c, h, w = feat.shape
        tok = self.cnn_to_vit(feat).flatten(2).transpose(1, 2)
        pos = interpolate_positions(self, self.pos_embed, h * w)
        tok = tok + pos.to(tok.dtype)
        grid = tok.reshape(b, h, w, -1).permute(0, 3, 1, 2)
        cls = self.vit_encoder(grid)
        back = self.vit_to_cnn(cls[:, :, None, None].expand(-1, -1, h, w).to(feat.dtype))
        fused = feat + back
        return mhc_over_pixels(self.fusion_mhc, fused) if isinstance(self.fusion_mhc, ManifoldHyperConnection) else fused


class FeaturePyramidNetwork(nn.Module):
    """Top-down FPN: lateral 1x1 -> (+ upsampled coarser level) -> two 3x3 conv-BN-ReLU -> mHC over pixels -> 1x1 to
    256 / 512 / 1024 channels (feature_fusion.py:11-157; R3)."""

    def __init__(self, channels: List[int], use_mhc: bool = True, fusion_method: str = "add"):
        super().__init__()
        if fusion_method != "add":
            raise ValueError("only fusion_method='add' is built (the reference's default)")
        n = len(channels)
        self.lateral_convs = nn.ModuleList([nn.Conv2d(c, 256, 1) for c in channels])
        self.refinement_convs = nn.ModuleList([nn.Sequential(nn.Conv2d(256, 256, 3, padding=1), nn.BatchNorm2d(256), nn.ReLU(inplace=True),
                                                             nn.Conv2d(256, 256, 3, padding=1), nn.BatchNorm2d(256), nn.ReLU(inplace=True))
                                               for _ in range(n)])
        self.mhc_fusions = nn.ModuleList([ManifoldHyperConnection(256, expansion_rate=2) if use_mhc else nn.Identity() for _ in range(n)])
        self.output_convs = nn.ModuleList([nn.Conv2d(256, oc, 1) for oc in (256, 512, 1024)[:n]])

    def _refine(self, i: int, p: torch.Tensor) -> torch.Tensor:
        p = self.refinement_convs[i](p)
        return mhc_over_pixels(self.mhc_fusions[i], p) if isinstance(self.mhc_fusions[i], ManifoldHyperConnection) else p

    def forward(self, features: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        small, medium, large = (features.get(k) for k in ("scale_small", "scale_medium", "scale_large"))
        out: Dict[str, torch.Tensor] = {}
        if large is None:
            return out
        p_l = self._refine(2, self.lateral_convs[2](large))
        out["fused_large"] = self.output_convs[2](p_l)
        if medium is not None:
            lat = self.lateral_convs[1](medium)
            p_m = self._refine(1, lat + F.interpolate(p_l, size=lat.shape[2:], mode="nearest"))
            out["fused_medium"] = self.output_convs[1](p_m)
            if small is not None:
                lat = self.lateral_convs[0](small)
                p_s = self._refine(0, lat + F.interpolate(p_m, size=lat.shape[2:], mode="nearest"))
                out["fused_small"] = self.output_convs[0](p_s)
        return out


class HybridVisionSystem(nn.Module):
    """HybridVisionSystem(config: dict) -- detection / features tasks of the reference composition root
    (hybrid_vision.py:17-485).  RAG, segmentation and depth heads (off by default there) are not built."""

    def __init__(self, config: Dict[str, Any]):
        super().__init__()
        self.config = config
        self.image_size = config.get("image_size", 416)
        self.num_classes = config.get("num_classes", 80)
        self.use_mhc = config.get("use_mhc", True)
        self.use_vit = config.get("use_vit", True)
        for off in ("use_rag", "has_segmentation", "has_depth"):
            if config.get(off, False):
                raise NotImplementedError(f"{off}: outside the hot path this package rebuilds")
        if not config.get("use_fpn", True):
            raise NotImplementedError("use_fpn=False (AdaptiveFeatureFusion) does not run in the reference either (D7)")
        self.backbone = HybridVisionBackbone(3, 32, [2, 3, 4, 2], self.use_mhc, "silu", 0.1)
        ch = self.backbone.get_output_channels()
        if self.use_vit:
            self.vit_encoder = HybridVisionEncoder(ch["scale_large"], 256, 6, 8, self.use_mhc)
        self.feature_fusion = FeaturePyramidNetwork([ch["scale_small"], ch["scale_medium"], ch["scale_large"]], self.use_mhc, "add")
        fused = [256, 512, 1024]
        self.detection_head = YOLODetectionHead(fused, self.num_classes, config.get("anchors", None), self.use_mhc)
        self.final_fusion = ManifoldHyperConnection(sum(fused), expansion_rate=2) if self.use_mhc else nn.Identity()
        self.output_projection = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(sum(fused), 512), nn.ReLU(), nn.Linear(512, 256))
        self._initialize_weights()

    def _initialize_weights(self):                          # hybrid_vision.py:183-197 (this is what zeroes the head biases, D18)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0, 0.01)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def refresh_coefficients(self) -> int:
        """All 76 layers' constrained matrices in one launch (they depend on parameters only)."""
        return refresh_static_coefficients(self)

    def forward(self, x: torch.Tensor, targets=None, text_query=None, task: str = "detection", compute_loss: bool = False) -> Dict[str, Any]:
        if x.is_cuda and torch.is_grad_enabled():
            # training: the coefficients of all layers hang on ONE autograd node (one forward launch, one backward launch)
            batched_training_coefficients(self)
            try:
                return self._forward(x, targets, task, compute_loss)
            finally:
                clear_training_coefficients(self)
        if x.is_cuda:
            self.refresh_coefficients()
        return self._forward(x, targets, task, compute_loss)

    def _forward(self, x: torch.Tensor, targets, task: str, compute_loss: bool) -> Dict[str, Any]:
        out: Dict[str, Any] = {}
        feats = self.backbone(x)
        out["backbone_features"] = feats
        if self.use_vit:
            vit = self.vit_encoder(feats["scale_large"])
            feats["scale_large"] = (feats["scale_large"] + vit) / 2
            out["vit_features"] = vit
        fused = self.feature_fusion(feats)
        out["fused_features"] = fused
        if task == "detection":
            det_in = {"scale_small": fused.get("fused_small", feats["scale_small"]),
                      "scale_medium": fused.get("fused_medium", feats["scale_medium"]),
                      "scale_large": fused.get("fused_large", feats["scale_large"])}
            out.update(self.detection_head(det_in, targets=targets, compute_loss=compute_loss))
        elif task == "features":
            out["all_features"] = {"backbone": feats, "fused": fused, "final": self._extract_final_features(fused)}
        elif task in ("segmentation", "depth"):
            raise NotImplementedError(f"task={task!r}: head not built (off by default in the reference)")
        if "final_features" not in out:
            out["final_features"] = self._extract_final_features(fused)
        return out

    def _extract_final_features(self, fused: Dict[str, torch.Tensor]) -> torch.Tensor:       # :369-402 with R8
        pooled = [F.adaptive_avg_pool2d(fused[k], (1, 1)).flatten(1) for k in ("fused_small", "fused_medium", "fused_large") if k in fused]
        if not pooled:
            return torch.tensor([], device=next(self.parameters()).device)
        v = self.final_fusion(torch.cat(pooled, dim=1))
        proj = self.output_projection[2:]
        return proj(v)

    def detect(self, x: torch.Tensor, confidence_threshold: float = 0.5, iou_threshold: float = 0.5, max_detections: int = 100,
               text_query=None) -> List[Dict[str, torch.Tensor]]:
        out = self.forward(x, task="detection")
        if "decoded" not in out:
            return []
        return self.detection_head.post_process(out["decoded"], confidence_threshold, iou_threshold, max_detections)

    def get_stability_metrics(self) -> Dict[str, Any]:
        metrics: Dict[str, Any] = {}
        for name, m in self.named_modules():
            if m is not self and hasattr(m, "get_stability_metrics"):
                for k, v in m.get_stability_metrics().items():
                    metrics[f"{name}.{k}"] = v
        return metrics

    def get_parameter_count(self) -> Dict[str, int]:
        counts = {n: sum(p.numel() for p in m.parameters()) for n, m in self.named_children()}
        counts["total"] = sum(p.numel() for p in self.parameters())
        counts["trainable"] = sum(p.numel() for p in self.parameters() if p.requires_grad)
        return counts
