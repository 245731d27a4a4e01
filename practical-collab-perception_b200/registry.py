"""Plug the B200 modules into a pcdet installation's registries.

The reference builds its modules by name from two dicts
(pcdet/models/backbones_3d/vfe/__init__.py:8-16, pcdet/models/backbones_2d/map_to_bev/__init__.py:5-9;
looked up in detector3d_template.py:96,129 and bev_layers/bev_maker.py:57-76), so replacing two entries
is the whole integration.
"""
from __future__ import annotations


def patch_pcdet(vfe_registry: dict = None, map_to_bev_registry: dict = None) -> None:
    """``patch_pcdet()`` imports pcdet and patches it in place; or pass the two ``__all__`` dicts."""
    from .modules import DynamicMeanVFE, DynamicPillarVFE, DynamicPillarVFESimple2D, PointPillarScatter
    if vfe_registry is None or map_to_bev_registry is None:
        from pcdet.models.backbones_3d import vfe                     # noqa: WPS433 (optional dependency)
        from pcdet.models.backbones_2d import map_to_bev
        vfe_registry = vfe.__all__ if vfe_registry is None else vfe_registry
        map_to_bev_registry = map_to_bev.__all__ if map_to_bev_registry is None else map_to_bev_registry
    vfe_registry["DynPillarVFE"] = DynamicPillarVFE
    vfe_registry["DynMeanVFE"] = DynamicMeanVFE                                 # vfe/__init__.py:13
    vfe_registry["DynamicPillarVFESimple2D"] = DynamicPillarVFESimple2D         # vfe/__init__.py:15
    map_to_bev_registry["PointPillarScatter"] = PointPillarScatter
