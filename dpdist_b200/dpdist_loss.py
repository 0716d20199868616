"""DPDist as a drop-in loss module: the reference's *name-addressed serialized graph* contract.

Consumers of a trained DPDist never call its Python functions; they `tf.train.import_meta_graph('model.ckpt.meta',
input_map={'input1': ..., 'input2': ..., 'Placeholder': is_training[, 'add_noise': ...]})`, fetch
`pc_compare/output1:0` / `pc_compare/output2:0`, and differentiate
    loss = (mean(output1[:, :, :, 0]) + mean(output2[:, :, :, 0])) / 2
into whatever produced `input1` (pcrnet-registration/iterative_PCRNet_ours.py:229-257;
train_multi_gpu_pc_compare_dist.py:427-463).  The tensors they bind are those of the first, full-batch
`get_model` call of the trainer (train...py:224-236) with the flags of log/.../log_trainours.txt:1.

`DPDistLoss` keeps that contract over torch tensors: the same input / output names, variables under their TF names
(`pc_compare/dpdist_local/mapper_conv{1..4}/{weights,biases}`, HWIO), loadable from a TF V2 checkpoint
(dpdist_b200/tf_checkpoint.py) or an .npz keyed by those names.  The variables are frozen (the consumers only
train their own scope); gradients flow into input1 / input2 / add_noise through dpd_head_backward_inputs and
dpd_fv_backward.
"""
import numpy as np
import torch

from . import dpdist_and_aue as MODEL
from . import tf_checkpoint, tf_util

INPUT_NAMES = ("input1", "input2", "labels12", "labels21", "Placeholder", "add_noise")
OUTPUT_NAMES = ("pc_compare/output1", "pc_compare/output2")


def _strip(name):
    name = name[:-2] if name.endswith(":0") else name
    for scope in ("g1/",):                       # import_scope the consumers use
        if name.startswith(scope):
            name = name[len(scope):]
    return name


class DPDistLoss(torch.nn.Module):
    def __init__(self, num_point=64, Embedding_Size=512, k=5, sigma3dmfv=0.125, localSNmlp=(1024, 1024, 1024),
                 full_fv=True, device=None, seed=None, train_variables=False):
        super().__init__()
        self.num_point, self.Embedding_Size, self.k = int(num_point), int(Embedding_Size), int(k)
        self.sigma3dmfv, self.localSNmlp, self.full_fv = float(sigma3dmfv), list(localSNmlp), bool(full_fv)
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.store = tf_util.VariableStore(device=dev, seed=seed)
        self.train_variables = bool(train_variables)
        # create the variables the way the graph does (scope pc_compare/dpdist_local, utils/dpdist_util.py:514-545)
        C = 20 if self.full_fv else 7
        from . import dpdist_util
        with tf_util.use_store(self.store), tf_util.variable_scope("pc_compare"):
            dpdist_util._head_variables(C * self.k ** 3, 3, self.localSNmlp, None)
        self._params = torch.nn.ParameterList(list(self.store.vars.values()))      # so .parameters() / .to() see them
        for p in self._params:
            p.requires_grad_(self.train_variables)

    # ---- variables under their TF names -------------------------------------------------------------
    def tf_variables(self):
        return dict(self.store.vars)

    def tf_state_dict(self):
        return {n: v.detach().cpu().numpy().copy() for n, v in self.store.vars.items()}

    def load_tf_state_dict(self, sd, strict=True):
        sd = {_strip(n): a for n, a in sd.items()}
        unknown = [n for n in sd if n not in self.store.vars]
        if strict and unknown:
            raise KeyError("unexpected variables: %s" % unknown)
        self.store.load_state_dict({n: a for n, a in sd.items() if n in self.store.vars}, strict=strict)
        return self

    def restore(self, path):
        """saver.restore(sess, path): a TF V2 checkpoint prefix ('.../model.ckpt') or an .npz of TF-named arrays.
        Optimizer slots and other scopes in the file are ignored."""
        if path.endswith(".npz"):
            with np.load(path) as z:
                sd = {n: z[n] for n in z.files}
        else:
            sd = tf_checkpoint.load_checkpoint(path, names=set(self.store.vars))
        return self.load_tf_state_dict({n: a for n, a in sd.items() if _strip(n) in self.store.vars})

    def save(self, path):
        """saver.save(sess, path): TF V2 checkpoint files (or .npz if the path says so)."""
        if path.endswith(".npz"):
            np.savez(path, **self.tf_state_dict())
            return path
        return tf_checkpoint.save_checkpoint(path, self.tf_state_dict())

    # ---- the graph -----------------------------------------------------------------------------------
    def forward(self, input1, input2, add_noise=None, is_training=False):
        """-> (output1, output2) = ('pc_compare/output1:0', 'pc_compare/output2:0'), each [B, NP, 1, 3]."""
        if input1.shape[1:] != (self.num_point, 3) or input2.shape != input1.shape:
            raise ValueError("input1 / input2 must be [B, %d, 3] (the serialized graph has static shapes; got %s, %s)"
                             % (self.num_point, tuple(input1.shape), tuple(input2.shape)))
        noise = 0 if add_noise is None else add_noise
        with tf_util.use_store(self.store):
            pred, _, _ = MODEL.get_model(input1, input2, bool(is_training) and self.train_variables, bn=0,
                                         Embedding_Size=self.Embedding_Size, k=self.k, sigma3dmfv=self.sigma3dmfv,
                                         localSNmlp=self.localSNmlp, full_fv=self.full_fv, add_noise=noise, reuse=True)
        return pred["pred_listAB"], pred["pred_listBA"]

    def loss(self, input1, input2, add_noise=None):
        """(mean(output1[...,0]) + mean(output2[...,0])) / 2  (iterative_PCRNet_ours.py:253-254; train...py:456-457)."""
        o1, o2 = self.forward(input1, input2, add_noise)
        return (o1[:, :, :, 0].mean() + o2[:, :, :, 0].mean()) / 2.0

    def run(self, fetches, feed_dict):
        """sess.run by tensor name: feed 'input1:0', 'input2:0' (optionally 'add_noise:0', 'Placeholder:0'); fetch
        'pc_compare/output1:0' / 'pc_compare/output2:0' (an import scope such as 'g1/' is accepted)."""
        feeds = {_strip(k): v for k, v in feed_dict.items()}
        unknown = [k for k in feeds if k not in INPUT_NAMES]
        if unknown:
            raise KeyError("unknown placeholders %s (the graph has %s)" % (unknown, list(INPUT_NAMES)))
        dev = self._params[0].device

        def dev_tensor(x):
            return x.to(dev) if torch.is_tensor(x) else torch.as_tensor(np.asarray(x, dtype=np.float32), device=dev)
        o1, o2 = self.forward(dev_tensor(feeds["input1"]), dev_tensor(feeds["input2"]),
                              dev_tensor(feeds["add_noise"]) if "add_noise" in feeds else None,
                              bool(feeds.get("Placeholder", False)))
        outs = {OUTPUT_NAMES[0]: o1, OUTPUT_NAMES[1]: o2}
        single = isinstance(fetches, str)
        res = []
        for f in ([fetches] if single else fetches):
            n = _strip(f)
            if n not in outs:
                raise KeyError("unknown tensor %r (the graph exposes %s)" % (f, list(OUTPUT_NAMES)))
            res.append(outs[n])
        return res[0] if single else res
