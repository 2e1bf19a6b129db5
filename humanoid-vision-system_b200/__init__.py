"""hvs_b200 -- B200 (sm_100a) implementation of the humanoid-vision-system hot path.

* ``ops``        functional wrappers over the C ABI (include/hvs_b200.h)
* ``mhc``        nn.Modules: StreamMHC (K1), SinkhornKnoppProjection, RMSNorm,
                 ManifoldHyperConnection (K2, the reference's signature and state_dict)
* ``detection``  YOLODecoder, YOLODetectionHead (forward / post_process / NMS / loss), NMSFilter
* ``hybrid_vision``  the host model (the reference's HybridVisionSystem) the modules above drop into

Import as ``import hvs_b200`` (see the shim in ``hvs_b200/__init__.py``).
"""
from . import _lib, build, ops  # noqa: F401
from ._lib import HvsError, load as load_library  # noqa: F401
from .mhc import (ManifoldHyperConnection, RMSNorm, SinkhornKnoppProjection, StreamMHC,  # noqa: F401
                  refresh_static_coefficients, stream_mhc_fwd_bwd_host)
from .detection import (NMSFilter, PostprocessingConfig, YOLOAnchorGenerator, YOLODecoder,  # noqa: F401
                        YOLODetectionHead, YOLOLoss, YOLOPredictionHead, dense_targets_from_boxes)
from .hybrid_vision import HybridVisionSystem  # noqa: F401

__all__ = ["ops", "build", "load_library", "HvsError", "StreamMHC", "SinkhornKnoppProjection", "RMSNorm",
           "ManifoldHyperConnection", "stream_mhc_fwd_bwd_host", "YOLODecoder", "YOLODetectionHead",
           "YOLOAnchorGenerator", "YOLOPredictionHead", "NMSFilter", "PostprocessingConfig", "YOLOLoss",
           "dense_targets_from_boxes", "HybridVisionSystem", "refresh_static_coefficients"]
