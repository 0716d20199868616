"""Variable scopes and initialisers with the reference's names (utils/tf_util.py).

The reference creates its trainable state through `tf.get_variable` inside nested
`tf.variable_scope`s (`pc_compare/dpdist_local/mapper_conv{1..4}/{weights,biases}`;
models/dpdist_and_aue.py:36, utils/dpdist_util.py:514, utils/tf_util.py:199-218) and consumers
bind by those names (checkpoints, import_meta_graph).  This module keeps that contract for
torch tensors: a VariableStore keyed by the TF names, HWIO shapes, Xavier-uniform weights and
zero biases.  Only what the hot path uses is here (`conv2d`'s variables; BN is off at the
reference defaults and is not implemented).
"""
import contextlib
import math

import numpy as np
import torch


class VariableStore:
    """name -> torch.nn.Parameter, with TF1 get_variable / reuse semantics."""

    def __init__(self, device=None, seed=None):
        self.vars = {}
        self.device = device
        self.gen = torch.Generator().manual_seed(seed) if seed is not None else None

    def get_variable(self, name, shape, initializer, reuse=None, trainable=True):
        if name in self.vars:
            v = self.vars[name]
            if tuple(v.shape) != tuple(shape):
                raise ValueError("variable %s exists with shape %s, requested %s" % (name, tuple(v.shape), tuple(shape)))
            if not trainable and v.requires_grad:      # e.g. moving statistics loaded from a checkpoint before first use
                v.requires_grad_(False)
            return v
        if reuse is True:
            raise ValueError("Variable %s does not exist, or was not created with tf.get_variable()" % name)
        t = initializer(shape, self.gen).to(torch.float32)
        dev = self.device if self.device is not None else ("cuda" if torch.cuda.is_available() else "cpu")
        v = torch.nn.Parameter(t.to(dev), requires_grad=trainable)
        self.vars[name] = v
        return v

    def trainable_variables(self, scope=None):
        return [v for n, v in self.vars.items() if v.requires_grad and (scope is None or n.startswith(scope))]

    def names(self):
        return list(self.vars.keys())

    def state_dict(self):
        return {n: v.detach().cpu().clone() for n, v in self.vars.items()}

    def load_state_dict(self, sd, strict=True):
        """Load a {tf_variable_name: array} mapping (e.g. read from a TF1 checkpoint)."""
        for n, a in sd.items():
            t = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a, dtype=torch.float32)
            if n in self.vars:
                if tuple(self.vars[n].shape) != tuple(t.shape):
                    raise ValueError("shape mismatch for %s" % n)
                with torch.no_grad():
                    self.vars[n].copy_(t)
            else:
                dev = self.device if self.device is not None else ("cuda" if torch.cuda.is_available() else "cpu")
                self.vars[n] = torch.nn.Parameter(t.to(dev))
        if strict:
            missing = [n for n in self.vars if n not in sd]
            if missing:
                raise KeyError("missing variables: %s" % missing)


_DEFAULT_STORE = VariableStore()
_SCOPE = []          # stack of (name, reuse)
_STORE_STACK = []


def default_store():
    return _STORE_STACK[-1] if _STORE_STACK else _DEFAULT_STORE


def reset_default_store(device=None, seed=None):
    global _DEFAULT_STORE
    _DEFAULT_STORE = VariableStore(device=device, seed=seed)
    return _DEFAULT_STORE


@contextlib.contextmanager
def use_store(store):
    _STORE_STACK.append(store)
    try:
        yield store
    finally:
        _STORE_STACK.pop()


@contextlib.contextmanager
def variable_scope(name, reuse=None):
    _SCOPE.append((name, reuse))
    try:
        yield "/".join(n for n, _ in _SCOPE)
    finally:
        _SCOPE.pop()


def _scoped(name):
    return "/".join([n for n, _ in _SCOPE] + [name])


def _reuse():
    for _, r in reversed(_SCOPE):
        if r is not None:
            return r
    return None


def xavier_initializer():
    """TF-semantics tf.contrib.layers.xavier_initializer(uniform=True): limit = sqrt(6/(fan_in+fan_out)),
    fan_in = prod(shape[:-2])*shape[-2], fan_out = prod(shape[:-2])*shape[-1]."""
    def init(shape, gen):
        rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
        fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
        limit = math.sqrt(6.0 / (fan_in + fan_out))
        return (torch.rand(tuple(shape), generator=gen, dtype=torch.float64) * 2 - 1).mul_(limit)
    return init


def constant_initializer(value):
    def init(shape, gen):
        return torch.full(tuple(shape), float(value), dtype=torch.float64)
    return init


def _variable_on_cpu(name, shape, initializer, use_fp16=False):
    """utils/tf_util.py:57-71.  The reference pins variables to /cpu:0 and ships them over PCIe every
    step; here they live in HBM on the rank's GPU."""
    return default_store().get_variable(_scoped(name), shape, initializer, reuse=_reuse())


def _variable_with_weight_decay(name, shape, stddev, wd, use_xavier=True):
    """utils/tf_util.py:73-98.  wd=0.0 registers a zero-valued weight_loss in the reference (:95-97),
    which has no effect on the gradients; it is not materialised here."""
    if not use_xavier:
        raise NotImplementedError("only the xavier initialiser is used on the DPDist path")
    return _variable_on_cpu(name, shape, xavier_initializer())


def conv2d_variables(num_in_channels, num_output_channels, kernel_size, scope, reuse=None):
    """The variables `tf_util.conv2d` creates (utils/tf_util.py:199-218): HWIO `weights`, zero `biases`."""
    with variable_scope(scope, reuse=reuse):
        kernel_h, kernel_w = kernel_size
        kernel = _variable_with_weight_decay("weights", [kernel_h, kernel_w, num_in_channels, num_output_channels],
                                             stddev=1e-3, wd=0.0, use_xavier=True)
        biases = _variable_on_cpu("biases", [num_output_channels], constant_initializer(0.0))
    return kernel, biases


def conv3d_variables(num_in_channels, num_output_channels, kernel_size, scope, reuse=None):
    """The variables `tf_util.conv3d` creates (utils/tf_util.py:344-359): DHWIO `weights`, zero `biases`."""
    with variable_scope(scope, reuse=reuse):
        kd, kh, kw = kernel_size
        kernel = _variable_with_weight_decay("weights", [kd, kh, kw, num_in_channels, num_output_channels],
                                             stddev=1e-3, wd=0.0, use_xavier=True)
        biases = _variable_on_cpu("biases", [num_output_channels], constant_initializer(0.0))
    return kernel, biases


BN_EPSILON = 0.001     # TF-semantics: tf.contrib.layers.batch_norm default epsilon (utils/tf_util.py:573-577 passes none)


def batch_norm_variables(num_channels, reuse=None):
    """The four variables tf.contrib.layers.batch_norm(center=True, scale=True, scope='bn') creates inside a conv
    scope (utils/tf_util.py:221-224, 558-577): bn/beta (0), bn/gamma (1) trainable; bn/moving_mean (0),
    bn/moving_variance (1) not trainable."""
    store = default_store()
    with variable_scope('bn', reuse=reuse):
        beta = store.get_variable(_scoped('beta'), [num_channels], constant_initializer(0.0), reuse=_reuse())
        gamma = store.get_variable(_scoped('gamma'), [num_channels], constant_initializer(1.0), reuse=_reuse())
        mean = store.get_variable(_scoped('moving_mean'), [num_channels], constant_initializer(0.0), reuse=_reuse(), trainable=False)
        var = store.get_variable(_scoped('moving_variance'), [num_channels], constant_initializer(1.0), reuse=_reuse(), trainable=False)
    return beta, gamma, mean, var


# ---- graph collections (tf.add_to_collection / tf.get_collection), used for the two losses ----
_COLLECTIONS = {}


def add_to_collection(name, value):
    _COLLECTIONS.setdefault(name, []).append(value)


def get_collection(name, clear=False):
    vals = list(_COLLECTIONS.get(name, []))
    if clear:
        _COLLECTIONS.pop(name, None)
    return vals


def clear_collections():
    _COLLECTIONS.clear()
