from .rbbox_target import rbbox_target, rbbox_target_single
from .transforms import (delta2hobb, delta2pointobb, delta2thetaobb, hobb2delta, hobb2pointobb, hobb_flip,
                         hobb_mapping, hobb_mapping_back, hobb_rescale, pointobb2bbox, pointobb2delta,
                         pointobb2thetaobb, pointobb_best_point_sort, pointobb_extreme_sort, pointobb_flip,
                         pointobb_mapping, pointobb_mapping_back, pointobb_rescale, rbbox2result, rbbox2roi,
                         thetaobb2delta, thetaobb2hobb, thetaobb2pointobb, thetaobb_flip, thetaobb_mapping,
                         thetaobb_mapping_back, thetaobb_rescale)

__all__ = ['rbbox_target', 'rbbox_target_single', 'delta2hobb', 'delta2pointobb', 'delta2thetaobb', 'hobb2delta',
           'hobb2pointobb', 'hobb_flip', 'hobb_mapping', 'hobb_mapping_back', 'hobb_rescale', 'pointobb2bbox',
           'pointobb2delta', 'pointobb2thetaobb', 'pointobb_best_point_sort', 'pointobb_extreme_sort',
           'pointobb_flip', 'pointobb_mapping', 'pointobb_mapping_back', 'pointobb_rescale', 'rbbox2result',
           'rbbox2roi', 'thetaobb2delta', 'thetaobb2hobb', 'thetaobb2pointobb', 'thetaobb_flip', 'thetaobb_mapping',
           'thetaobb_mapping_back', 'thetaobb_rescale']
