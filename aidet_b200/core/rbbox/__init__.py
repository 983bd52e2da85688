from .transforms import (delta2hobb, delta2pointobb, delta2thetaobb, hobb2delta, hobb2pointobb, hobb_rescale, pointobb2bbox,
                         pointobb2delta, pointobb_rescale, rbbox2result, thetaobb2delta, thetaobb2pointobb,
                         thetaobb_rescale)

__all__ = ['delta2hobb', 'delta2pointobb', 'delta2thetaobb', 'hobb2delta', 'hobb2pointobb', 'hobb_rescale',
           'pointobb2bbox', 'pointobb2delta', 'pointobb_rescale', 'rbbox2result', 'thetaobb2delta', 'thetaobb2pointobb',
           'thetaobb_rescale']
