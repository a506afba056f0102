"""
Stand-in for the third-party `ndsplines` package (unpinned in the reference's
requirements.txt:3; source not vendored).  Restates its published behaviour for
the two call shapes the reference uses (calculator.py:540,568; 207,242,286,337):
a tensor-product B-spline  S(x) = sum_{i..} C[i..] prod_d B_{i_d,k}(x_d)  with
`extrapolate=False` (NaN outside the knot span), evaluated with optional
per-axis derivative orders `nus`.  Built on scipy.interpolate.BSpline.
TEST INFRASTRUCTURE ONLY.
"""
import numpy as np
from scipy.interpolate import BSpline


class NDSpline:
    def __init__(self, knots, coefficients, degrees, periodic=False,
                 extrapolate=True):
        self.knots = [np.asarray(k, dtype=float) for k in knots]
        self.coefficients = np.asarray(coefficients, dtype=float)
        self.xdim = len(self.knots)
        if np.isscalar(degrees):
            degrees = [int(degrees)] * self.xdim
        self.degrees = list(degrees)
        self.extrapolate = extrapolate

    def _design(self, axis, x, nu):
        t = self.knots[axis]
        k = self.degrees[axis]
        n = len(t) - k - 1
        out = np.zeros((len(x), n))
        eye = np.eye(n)
        for i in range(n):
            vals = BSpline(t, eye[i], k, extrapolate=False)(x, nu=nu)
            out[:, i] = vals
        return out

    def __call__(self, x, nus=0):
        x = np.asarray(x, dtype=float)
        if self.xdim == 1:
            x = x.reshape(-1, 1)
        x = x.reshape(-1, self.xdim)
        if np.isscalar(nus):
            nus = [int(nus)] * self.xdim
        nus = [int(v) for v in np.asarray(nus).ravel()]
        if len(x) == 0:
            return np.zeros(0)
        mats = [self._design(d, x[:, d], nus[d]) for d in range(self.xdim)]
        if self.xdim == 1:
            return mats[0] @ self.coefficients.reshape(-1)
        if self.xdim == 3:
            return np.einsum('pl,pm,pn,lmn->p', mats[0], mats[1], mats[2],
                             self.coefficients, optimize=True)
        raise NotImplementedError
