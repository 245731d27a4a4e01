"""practical-collab-perception_b200 (import alias: ``pcp_b200``) - B200-native point->BEV front end.

Drop-in replacements for the reference's ``DynamicPillarVFE`` / ``PFNLayerV2`` /
``PointPillarScatter`` modules and its MoDAR exchange arithmetic, executing in hand-written sm_100a
CUDA kernels behind a C ABI (``include/pcp_b200.h``, ``csrc/``).  There is no CPU or torch fallback:
every entry point raises if ``libpcp_b200.so`` is missing or no CUDA device is present.
"""
__version__ = "0.1.0"

from .config import CfgDict  # noqa: F401


def __getattr__(name):
    # heavy members are resolved lazily so that ``import pcp_b200`` works on a CPU-only build host
    if name in ("DynamicPillarVFE", "PFNLayerV2", "PointPillarScatter", "DynamicMeanVFE", "DynamicPillarVFESimple2D"):
        from . import modules
        return getattr(modules, name)
    if name in ("modar_exchange", "ModarExchange"):
        from . import modar
        return getattr(modar, name)
    if name in ("bev_scatter", "interpolate_points_feat_from_bev_img"):
        from . import hunter_toolbox
        return getattr(hunter_toolbox, name)
    if name in ("class_agnostic_nms", "multi_classes_nms", "nms_gpu", "nms_normal_gpu", "boxes_iou_bev"):
        from . import nms
        return getattr(nms, name)
    if name in ("pack_exchange", "unpack_exchange", "read_exchange", "write_exchange", "ExchangeMessage", "select_foreground",
                "exchange_payloads"):
        from . import exchange
        return getattr(exchange, name)
    if name in ("fuse_agent_points",):
        from . import early_fusion
        return getattr(early_fusion, name)
    if name in ("collate_points", "load_points_to_gpu", "PackedPoints", "PointsPrefetcher"):
        from . import loader
        return getattr(loader, name)
    if name in ("FrontEnd",):
        from . import frontend
        return getattr(frontend, name)
    raise AttributeError(name)
