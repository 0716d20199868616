"""Host-side mirror of the reference's model API `models/dpdist_and_aue.py` (hot path only).

Same names and signatures: placeholder_inputs (:23-28), get_model (:31-86), get_loss (:203-204).
The two toy auto-encoders in the reference file (:88-200) are consumers of the path, not part of
it, and are out of scope.
"""
import torch

from . import dpdist_util as dpdist
from . import tf_util


FUSED_INFERENCE = True   # module switch: False forces the staged path (get_3dmfv_tf -> local_z -> DPDist) everywhere


def placeholder_inputs(batch_size, num_point, NUM_DIMS=2, device=None):
    """models/dpdist_and_aue.py:23-28.  TF placeholders become zero-filled device buffers with the
    graph names the consumers bind to: input1, input2, labels12, labels21."""
    dev = device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu")
    pcA_pl = torch.zeros((batch_size, num_point, NUM_DIMS), dtype=torch.float32, device=dev)   # 'input1'
    pcB_pl = torch.zeros((batch_size, num_point, NUM_DIMS), dtype=torch.float32, device=dev)   # 'input2'
    labels_AB = torch.zeros((batch_size, num_point), dtype=torch.float32, device=dev)          # 'labels12'
    labels_BA = torch.zeros((batch_size, num_point), dtype=torch.float32, device=dev)          # 'labels21'
    return pcA_pl, pcB_pl, labels_AB, labels_BA


def get_model(pcA, pcB,
              is_training, bn_decay=None, wd=0.0, bn=True,
              Embedding_Size=512, pn='pn', sig=True,
              k=0, overlap=False,
              localSNmlp=[1024, 1024, 1024], full_fv=True, sigma3dmfv=0.0625 * 2, conv_version=1, add_noise=0,
              reuse=None, materialize_embeddings=False):
    """models/dpdist_and_aue.py:31-86 -> (pred_set, end_points, embedding_set).

    pred_set      {'pred_listAB','pred_listBA'}: [B,NP,1,3] each ('pc_compare/output1', '.../output2')
    embedding_set {'embedding_A','embedding_B'}: what local_z returned.  The reference holds the dense
                  [B,V,k^3*20] tensors here; they are LocalPatches handles unless
                  materialize_embeddings=True (5.12 MB per cloud at the defaults).
    `reuse` and `materialize_embeddings` are additions; everything else is the reference signature.
    `bn` defaults to True as in the reference, whose trainer passes bn=int('0') (train...py:98,225).  With bn truthy the
    head runs batch norm after every conv: is_training True = batch statistics, moving-average updates with `bn_decay`
    (layer-by-layer fp32 path); is_training False = moving statistics folded into the fused layers."""
    with tf_util.variable_scope('pc_compare', reuse=reuse):                               # :36
        NUM_DIMS = pcA.shape[-1]
        n_gaussians = Embedding_Size
        if k > 0:
            flatten = False
        else:
            flatten = True
        if pn == 'pointnet':
            raise NotImplementedError("the PointNet encoder variant is out of scope (SURVEY.md 2.1 row 15)")
        if k <= 0:
            raise NotImplementedError("k == 0 (global FV + MLP) is not the DPDist hot path")
        pcA_noise = pcA + add_noise                                                       # :45
        B = pcA.shape[0]
        # Inference (no gradient can be asked of the result): the whole graph below is one library call,
        # dpd_model_forward.  Same ops, same results; the training path keeps the three reference stages.
        wants_grad = torch.is_grad_enabled() and (bool(is_training) if isinstance(is_training, (bool, int)) else True)
        # ... unless the graph is differentiated through into the point clouds (DPDist as a loss, :203 consumers)
        wants_grad = wants_grad or (torch.is_grad_enabled() and any(torch.is_tensor(t) and t.requires_grad
                                                                    for t in (pcA, pcB, add_noise)))
        if FUSED_INFERENCE and not wants_grad and NUM_DIMS == 3 and conv_version == 1 and not bn:
            fv, out, C = dpdist.model_forward(torch.cat([pcA_noise, pcB], 0), torch.cat([pcB, pcA], 0), n_gaussians,
                                              sigma3dmfv, full_fv, k, localSNmlp, reuse=reuse)
            out = out.view(2, B, pcA.shape[1], 1, 3)
            embedding_A, embedding_B = dpdist.LocalPatches(fv[:B], k), dpdist.LocalPatches(fv[B:], k)
            if materialize_embeddings:
                embedding_A, embedding_B = embedding_A.materialize(), embedding_B.materialize()
            return ({'pred_listAB': out[0], 'pred_listBA': out[1]}, {},
                    {'embedding_A': embedding_A, 'embedding_B': embedding_B})
        inputs_need = torch.is_grad_enabled() and any(torch.is_tensor(t) and t.requires_grad for t in (pcA, pcB, add_noise))
        if FUSED_INFERENCE and not inputs_need and NUM_DIMS == 3 and conv_version == 1 and not bn:
            # training step of DPDist itself (the clouds are data): still one library call for 3DmFV + head, with an
            # autograd node for the 8 variables
            fv, out, C = dpdist.model_forward_train(torch.cat([pcA_noise, pcB], 0), torch.cat([pcB, pcA], 0), n_gaussians,
                                                    sigma3dmfv, full_fv, k, localSNmlp, reuse=reuse)
            out = out.view(2, B, pcA.shape[1], 1, 3)
            embedding_A, embedding_B = dpdist.LocalPatches(fv[:B], k), dpdist.LocalPatches(fv[B:], k)
            if materialize_embeddings:
                embedding_A, embedding_B = embedding_A.materialize(), embedding_B.materialize()
            return ({'pred_listAB': out[0], 'pred_listBA': out[1]}, {},
                    {'embedding_A': embedding_A, 'embedding_B': embedding_B})
        # one launch encodes both clouds of every pair: rows [A | B]
        emb = dpdist.get_3dmfv_tf(torch.cat([pcA_noise, pcB], 0), n_gaussians=n_gaussians,
                                  flatten=flatten, full_fv=full_fv,
                                  normalize=True, sigma=sigma3dmfv)                       # :56-61
        embedding_A, embedding_B = emb[:B], emb[B:]
        embedding_A, C = dpdist.local_z(embedding_A, is_training, reuse=None, NUM_DIMS=NUM_DIMS, k=k, overlap=overlap)   # :64
        embedding_B, _ = dpdist.local_z(embedding_B, is_training, reuse=True, NUM_DIMS=NUM_DIMS, k=k, overlap=overlap)   # :65
        net = dpdist.DPDist(pcA, pcB, embedding_A,
                            embedding_B, C, is_training, bn_decay=bn_decay,
                            reuse=reuse,
                            bn=bn, wd=wd,
                            sig=sig, Embedding_Size=Embedding_Size,
                            NUM_DIMS=NUM_DIMS, mlp=localSNmlp, k=k, output_act='relu',
                            conv_version=conv_version)                                    # :69-75
        pred_listAB = net[0]                                                              # 'output1' :78
        pred_listBA = net[1]                                                              # 'output2' :79
        pred_set = {'pred_listAB': pred_listAB,
                    'pred_listBA': pred_listBA}
        if materialize_embeddings:
            embedding_A, embedding_B = embedding_A.materialize(), embedding_B.materialize()
        embedding_set = {'embedding_A': embedding_A,
                         'embedding_B': embedding_B}
        end_points = {}
        return pred_set, end_points, embedding_set


def get_loss(pred_set, end_points, labels, loss_type='l1_dist'):
    """models/dpdist_and_aue.py:203-204."""
    return dpdist.get_loss(pred_set, end_points, labels, loss_type=loss_type)
