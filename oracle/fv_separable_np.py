"""CPU ORACLE cross-check -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Independent (separable, numpy, fp64 by default) formulation of the 3DmFV encoding
of utils/dpdist_util.py:22-141.  The reference GMM is an axis-aligned grid with one
isotropic sigma and equal weights, so the responsibility factorises:

    Q[n,(i0,i1,i2)] = qy[n,i0] * qx[n,i1] * qz[n,i2],   q_a[n,i] = e_a[n,i] / sum_i e_a[n,i],
    e_a[n,i] = exp(-0.5*((p_a - l[i])/sigma)^2)

(flat g = i0*G^2 + i1*G + i2 has mu_g = (x=l[i1], y=l[i0], z=l[i2]) because
np.meshgrid defaults to 'xy' indexing, :47-48).  Used to validate the literal
restatement in oracle/dpdist_oracle.py (two formulations must agree) and as the
arithmetic the CUDA kernel is designed after.  PARITY UNPINNED (see dpdist_oracle.py).
"""
import numpy as np


def fv_separable(points, n_gaussians=512, sigma=0.125, full_fv=True, dtype=np.float64):
    pts = np.asarray(points, dtype=dtype)
    B, N, _ = pts.shape
    G = int(np.ceil(np.power(n_gaussians, 1 / 3)))
    assert G ** 3 == n_gaussians
    l = (np.linspace(-1, 1, G, False) + 1 / G).astype(np.float32).astype(dtype)
    w = dtype(1.0) / n_gaussians
    z = (pts[:, :, :, None] - l[None, None, None, :]) / dtype(sigma)      # [B,N,3,G]
    e = np.exp(-0.5 * z * z)
    q = e / e.sum(-1, keepdims=True)
    m = q * z
    s = q * (z * z - 1)
    X, Y, Z = 0, 1, 2

    def outer(ay, ax, az):   # [B,N,G] x3 -> [B,N,G,G,G] indexed (i0=y, i1=x, i2=z)
        return ay[:, :, :, None, None] * ax[:, :, None, :, None] * az[:, :, None, None, :]

    Q = outer(q[:, :, Y], q[:, :, X], q[:, :, Z])
    dmu = [outer(q[:, :, Y], m[:, :, X], q[:, :, Z]), outer(m[:, :, Y], q[:, :, X], q[:, :, Z]),
           outer(q[:, :, Y], q[:, :, X], m[:, :, Z])]
    dsg = [outer(q[:, :, Y], s[:, :, X], q[:, :, Z]), outer(s[:, :, Y], q[:, :, X], q[:, :, Z]),
           outer(q[:, :, Y], q[:, :, X], s[:, :, Z])]
    V = n_gaussians

    def red(t, op):
        return op(t.reshape(B, N, V), axis=1)

    cpi = 1.0 / (np.sqrt(w) * N)
    d_pi = [(red(Q, np.mean) - w) * cpi]
    if full_fv:
        d_pi.append((red(Q, np.max) - w) * cpi)
    ops = [np.mean, np.max, np.min] if full_fv else [np.mean]
    d_mu = [red(t, op) / np.sqrt(w) for op in ops for t in dmu]
    d_sg = [red(t, op) / np.sqrt(2 * w) for op in ops for t in dsg]

    def norm(chs):
        t = np.stack(chs, -1)                                              # [B,V,c]
        t = np.sign(t) * np.sqrt(np.maximum(np.abs(t), 1e-12))
        ss = (t * t).sum(1, keepdims=True)
        return t / np.sqrt(np.maximum(ss, 1e-12))

    return np.concatenate([norm(d_pi), norm(d_mu), norm(d_sg)], -1)        # [B,V,20]
