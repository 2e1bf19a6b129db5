"""hvs_b200 -- B200 (sm_100a) implementation of the humanoid-vision-system hot path.

* ``ops``        functional wrappers over the C ABI (include/hvs_b200.h)
* ``mhc``        nn.Modules: StreamMHC (K1), SinkhornKnoppProjection, RMSNorm,
                 ManifoldHyperConnection (K2, the reference's signature and state_dict)
* ``detection``  YOLODecoder, YOLODetectionHead (forward / post_process / NMS), NMSFilter

Import as ``import hvs_b200`` (see the shim in ``hvs_b200/__init__.py``).
"""
from . import _lib, build, ops  # noqa: F401
from ._lib import HvsError, load as load_library  # noqa: F401
from .mhc import (ManifoldHyperConnection, RMSNorm, SinkhornKnoppProjection, StreamMHC,  # noqa: F401
                  stream_mhc_fwd_bwd_host)
from .detection import (NMSFilter, PostprocessingConfig, YOLOAnchorGenerator, YOLODecoder,  # noqa: F401
                        YOLODetectionHead, YOLOPredictionHead)

__all__ = ["ops", "build", "load_library", "HvsError", "StreamMHC", "SinkhornKnoppProjection", "RMSNorm",
           "ManifoldHyperConnection", "stream_mhc_fwd_bwd_host", "YOLODecoder", "YOLODetectionHead",
           "YOLOAnchorGenerator", "YOLOPredictionHead", "NMSFilter", "PostprocessingConfig"]
