from .rbbox_nms import (get_det_rbboxes, multiclass_nms, multiclass_nms_with_index, multiclass_thetaobb_nms,
                        thetaobb_nms_by_bbox_nms)

__all__ = ['get_det_rbboxes', 'multiclass_nms', 'multiclass_nms_with_index', 'multiclass_thetaobb_nms',
           'thetaobb_nms_by_bbox_nms']
