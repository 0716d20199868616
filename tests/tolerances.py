"""Parity tolerances, written once.

FV features : |a-b| <= 1e-4*|b| + 2e-6.  The absolute term covers the reference's discontinuity
              sign(x)*sqrt(max(|x|,1e-12)) (utils/dpdist_util.py:118-121): an entry whose true value
              underflows is 0 or +-1e-6 before the L2 normalisation depending on where exp()
              flushes, i.e. <= 1e-6/norm after it.
distances   : |a-b| <= 1e-4*|b| + 1e-5 on outputs in [0,2], i.e. the absolute term is 5e-6 of the
              output range: fp32 rounding noise of three K~1000-2500 dot-product layers.  Measured on
              B200 against the fp64 twin: fp32 oracle 1.3e-6, SIMT fp32 path 2.5e-6, tensor-core
              path 4.2e-6; torch's own fp32 matmul is 6-8e-6 off fp64 on ONE such layer.
indices     : bit-exact.
"""
import numpy as np

FV_RTOL, FV_ATOL = 1e-4, 2e-6
OUT_RTOL, OUT_ATOL = 1e-4, 1e-5


def _np(x):
    try:
        import torch
        if torch.is_tensor(x):
            return x.detach().cpu().double().numpy()
    except ImportError:
        pass
    return np.asarray(x, dtype=np.float64)


def assert_close(got, want, rtol, atol, what=""):
    g, w = _np(got), _np(want)
    assert g.shape == w.shape, "%s: shape %s vs %s" % (what, g.shape, w.shape)
    assert np.isfinite(g).all(), "%s: non-finite values" % what
    err = np.abs(g - w)
    tol = rtol * np.abs(w) + atol
    bad = err > tol
    if bad.any():
        i = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError("%s: %d/%d entries out of tolerance; worst at %s got %.9g want %.9g (err %.3g, tol %.3g)"
                             % (what, bad.sum(), bad.size, i, g[i], w[i], err[i], tol[i]))
    return float(err.max())


def assert_fv_close(got, want, what="fv"):
    return assert_close(got, want, FV_RTOL, FV_ATOL, what)


def assert_out_close(got, want, what="out"):
    return assert_close(got, want, OUT_RTOL, OUT_ATOL, what)


# gradients w.r.t. the point clouds (DPDist as a loss): the signed square root has derivative 0.5/sqrt(|x|), and
# the mean channels are sums with cancellation, so wherever a mean statistic nearly cancels (|x| ~ 1e-9 from terms of
# 1e-4) its derivative is large AND carries the full fp32 rounding noise of the sum.  This is inherent to the reference's
# formula in fp32: the fp32 CPU oracle itself is up to 0.3 % of max|g| away from its fp64 twin on isolated entries
# (median relative error 2e-6).  Hence two kinds of check:
#   conditioned   : the upstream gradient is zeroed where |fv| < 1e-2 -> every entry must agree tightly
#   unconditioned : >= `frac` of the entries agree tightly, none is grossly off, the median relative error is ~1e-6
GRAD_RTOL, GRAD_RMS_TOL = 1e-3, 1e-3


def assert_grad_close(got, want, what="grad", rtol=GRAD_RTOL, rms_tol=GRAD_RMS_TOL, frac=1.0, gross=0.05):
    g, w = _np(got), _np(want)
    assert g.shape == w.shape, "%s: shape %s vs %s" % (what, g.shape, w.shape)
    assert np.isfinite(g).all(), "%s: non-finite values" % what
    rms = float(np.sqrt((w * w).mean()))
    assert rms > 0, "%s: reference gradient is identically zero" % what
    err = np.abs(g - w)
    ok = err <= rtol * np.abs(w) + rms_tol * rms
    assert ok.mean() >= frac, "%s: only %.4f of the entries within tolerance (need %.4f); max err %.3g, rms %.3g" % (
        what, ok.mean(), frac, err.max(), rms)
    assert err.max() <= gross * np.abs(w).max(), "%s: max err %.3g vs max|g| %.3g" % (what, err.max(), np.abs(w).max())
    nz = np.abs(w) > 1e-3 * rms
    med = float(np.median(err[nz] / np.abs(w)[nz]))
    assert med <= 1e-4, "%s: median relative error %.3g" % (what, med)
    return float(err.max())
