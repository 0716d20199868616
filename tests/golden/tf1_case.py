"""Inputs and variables of the TF1 golden cases, generated with numpy's legacy MT19937 stream (stable across every numpy
release TF 1.x and this repo can run on), so that tools/make_tf_golden.py (TensorFlow 1.15 side, no torch) and
tests/test_tf_golden.py (this side) build bit-identical fp32 arrays without shipping 18.7 MB of weights.

Pure numpy on purpose: no torch, no TensorFlow, no import from the package."""
import numpy as np

PREFIX = "pc_compare/dpdist_local/"
CASES = {
    # name: (pairs, num_point, Embedding_Size, k, sigma3dmfv, H, bn)
    "anchor": (1, 64, 512, 5, 0.125, 1024, 0),        # BASELINE configs[0]
    "batch4": (4, 64, 512, 5, 0.125, 1024, 0),
    "g5k3": (3, 32, 125, 3, 0.2, 256, 0),             # another grid / patch edge, a narrower head
    "batch4_bn": (4, 64, 512, 5, 0.125, 1024, 1),     # --BN 1, is_training True: batch statistics
}


def inputs(name, seed=0):
    pairs, n, _, _, _, _, _ = CASES[name]
    rs = np.random.RandomState(1000 + seed + 17 * sorted(CASES).index(name))
    shift = rs.uniform(-0.1, 0.1, size=(pairs, 1, 3))
    pcA = (rs.uniform(-0.8, 0.8, size=(pairs, n, 3)) + shift).astype(np.float32)
    pcB = (rs.uniform(-0.8, 0.8, size=(pairs, n, 3)) + shift).astype(np.float32)
    pcB[:, -2:, :] += np.float32(0.6)                 # a few queries outside the unit cube: exercises the mask
    labels = rs.uniform(0.0, 0.3, size=(pairs, n)).astype(np.float32)
    return pcA, pcB, labels


def variables(name, seed=0):
    """{tf variable name: fp32 array} in the reference's HWIO layouts (utils/tf_util.py:199-218).  Scales are chosen so
    that the outputs spread over (0, 2) instead of saturating relu6 or dying in the ReLUs."""
    _, _, _, k, _, H, bn = CASES[name]
    rs = np.random.RandomState(2000 + seed + 17 * sorted(CASES).index(name))
    E = 20 * k ** 3
    shapes = [(1, E + 3, 1, H), (1, 1, H, H), (1, 1, H, H), (1, 1, H, 3)]
    gains = (600.0, 2.0, 2.0, 1.0)
    out = {}
    for i, (shp, gain) in enumerate(zip(shapes, gains), 1):
        fan_in, fan_out = shp[0] * shp[1] * shp[2], shp[0] * shp[1] * shp[3]
        limit = gain * np.sqrt(6.0 / (fan_in + fan_out))
        out[PREFIX + "mapper_conv%d/weights" % i] = rs.uniform(-limit, limit, size=shp).astype(np.float32)
        b = rs.normal(0.0, 0.05, size=(shp[3],))
        if i == 4:
            b = b + 1.0
        out[PREFIX + "mapper_conv%d/biases" % i] = b.astype(np.float32)
        if bn:
            out[PREFIX + "mapper_conv%d/bn/beta" % i] = rs.normal(0.0, 0.1, size=(shp[3],)).astype(np.float32)
            out[PREFIX + "mapper_conv%d/bn/gamma" % i] = rs.uniform(0.5, 1.5, size=(shp[3],)).astype(np.float32)
            out[PREFIX + "mapper_conv%d/bn/moving_mean" % i] = np.zeros((shp[3],), np.float32)
            out[PREFIX + "mapper_conv%d/bn/moving_variance" % i] = np.ones((shp[3],), np.float32)
    return out
