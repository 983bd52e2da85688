"""`mmdet.datasets`-side helpers on the OBB hot path: the DOTA scene merge (mmdet/datasets/dota.py:310-336)."""
from .dota_merge import (DOTA_CLASSES, format_dota_results, merge_txt, mergebypoly, mergebypoly_mp, mergebyrec,
                         mergebyrec_mp, parse_tile_name)

__all__ = ['DOTA_CLASSES', 'format_dota_results', 'merge_txt', 'mergebypoly', 'mergebypoly_mp', 'mergebyrec',
           'mergebyrec_mp', 'parse_tile_name']
