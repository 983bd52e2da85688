from .transforms import (delta2pointobb, delta2thetaobb, pointobb2bbox, pointobb2delta, pointobb_rescale, rbbox2result,
                         thetaobb2delta, thetaobb2pointobb, thetaobb_rescale)

__all__ = ['delta2pointobb', 'delta2thetaobb', 'pointobb2bbox', 'pointobb2delta', 'pointobb_rescale', 'rbbox2result',
           'thetaobb2delta', 'thetaobb2pointobb', 'thetaobb_rescale']
