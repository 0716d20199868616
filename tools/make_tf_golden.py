#!/usr/bin/env python
"""Pins the CPU oracle (and through it the CUDA path) against the REAL reference: run this wherever TensorFlow 1.14 /
1.15 exists, against an UNMODIFIED checkout of dahliau/DPDist, and commit what it writes.

    python tools/make_tf_golden.py --reference /path/to/DPDist [--out tests/golden] [--cases anchor batch4 g5k3 batch4_bn]

For every case of tests/golden/tf1_case.py it builds the reference graph exactly as the trainer does
(train_multi_gpu_pc_compare_dist.py:192-228: placeholder_inputs, is_training / add_noise placeholders, MODEL.get_model,
MODEL.get_loss), assigns the case's variables, runs one forward + gradient evaluation on the CPU and writes
    tests/golden/tf1_<case>.npz      inputs, output1 / output2, both 3DmFV embeddings' FV records, loss_samples, loss_pred,
                                     d loss_samples / d variables, d ((mean out1 + mean out2)/2) / d input1, input2,
                                     and for --BN 1 the moving statistics after one training-mode evaluation
    tests/golden/tf1_ckpt/model.ckpt*   a tf.train.Saver checkpoint of the anchor case's variables (TF V2 bundle)
tests/test_tf_golden.py picks these files up when present: oracle vs TF1 (CPU), tf_checkpoint reader vs the TF-written
bundle (CPU), CUDA path vs TF1 (-m gpu).  Until they exist every parity claim of this repo reads "matches the CPU
restatement", not "matches TF1" (DESIGN.md section 2).

This script imports nothing from dpdist_b200 and needs only numpy + tensorflow 1.x."""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import tf1_case  # noqa: E402


def run_case(tf, MODEL, name, out_dir, save_ckpt):
    pairs, n, emb, k, sigma, H, bn = tf1_case.CASES[name]
    pcA, pcB, labels = tf1_case.inputs(name)
    var = tf1_case.variables(name)
    g = tf.Graph()
    with g.as_default():
        with tf.device('/cpu:0'):
            pcA_pl, pcB_pl, labels_AB_pl, labels_BA_pl = MODEL.placeholder_inputs(pairs, n, NUM_DIMS=3)
            is_training_pl = tf.placeholder(tf.bool, shape=())
            noise_pl = tf.placeholder(tf.float32, shape=(pairs, n, 3), name='add_noise')
            bn_decay = tf.constant(0.5, tf.float32)                     # get_bn_decay(0) (:992-1000)
            pred, end_points, emb_set = MODEL.get_model(pcA_pl, pcB_pl, is_training_pl, bn_decay=bn_decay, wd=0.0,
                                                        bn=bn, sig=False, Embedding_Size=emb, pn='3dmfv', k=k,
                                                        localSNmlp=[H, H, H], overlap=True, full_fv=True, conv_version=1,
                                                        sigma3dmfv=sigma, add_noise=noise_pl)
            MODEL.get_loss(pred, end_points, labels_AB_pl, loss_type='l1_dist')
            loss_samples = tf.add_n(tf.get_collection('loss_samples'))
            loss_pred = tf.add_n(tf.get_collection('loss_pred'))
            tvars = tf.get_collection(tf.GraphKeys.TRAINABLE_VARIABLES, scope='pc_compare')
            grads = tf.gradients(loss_samples, tvars)
            consumer_loss = (tf.reduce_mean(pred['pred_listAB'][:, :, :, 0]) + tf.reduce_mean(pred['pred_listBA'][:, :, :, 0])) / 2
            g_in = tf.gradients(consumer_loss, [pcA_pl, pcB_pl])
            all_vars = {v.op.name: v for v in tf.global_variables()}
            missing = [n_ for n_ in var if n_ not in all_vars]
            if missing:
                raise SystemExit("graph has no variable(s) %s; it has %s" % (missing, sorted(all_vars)))
            assign = [tf.assign(all_vars[n_], var[n_]) for n_ in var]
            saver = tf.train.Saver()
        config = tf.ConfigProto(allow_soft_placement=True, device_count={'GPU': 0})
        with tf.Session(config=config) as sess:
            sess.run(tf.global_variables_initializer())
            sess.run(assign)
            feed = {pcA_pl: pcA, pcB_pl: pcB, labels_AB_pl: labels, labels_BA_pl: -np.ones_like(labels),
                    is_training_pl: bool(bn), noise_pl: np.zeros_like(pcA)}
            # the dense patch tensors [B, V, k^3*20] are not stored (5 MB per cloud); the FV record of voxel v is the
            # centre tap of its own patch
            c = ((k - 1) // 2 * k + (k - 1) // 2) * k + (k - 1) // 2
            fetch = [pred['pred_listAB'], pred['pred_listBA'], emb_set['embedding_A'][:, :, c * 20:(c + 1) * 20],
                     emb_set['embedding_B'][:, :, c * 20:(c + 1) * 20], loss_samples, loss_pred, grads,
                     [x if x is not None else tf.zeros_like(pcA_pl) for x in g_in]]
            out1, out2, fvA, fvB, ls, lp, gv, gi = sess.run(fetch, feed)
            rec = dict(pcA=pcA, pcB=pcB, labels=labels, output1=out1, output2=out2, fvA=fvA, fvB=fvB,
                       loss_samples=np.float32(ls), loss_pred=np.float32(lp), grad_input1=gi[0], grad_input2=gi[1],
                       tf_version=np.array(tf.__version__))
            for v_, g_ in zip(tvars, gv):
                rec["grad/" + v_.op.name] = g_
            if bn:
                for n_, v_ in all_vars.items():
                    if "/bn/moving_" in n_:
                        rec["after/" + n_] = sess.run(v_)
                feed[is_training_pl] = False
                rec["eval_output1"], rec["eval_output2"] = sess.run([pred['pred_listAB'], pred['pred_listBA']], feed)
            np.savez_compressed(os.path.join(out_dir, "tf1_%s.npz" % name), **rec)
            print("wrote tf1_%s.npz  loss_samples %.6f loss_pred %.6f" % (name, ls, lp))
            if save_ckpt:
                os.makedirs(os.path.join(out_dir, "tf1_ckpt"), exist_ok=True)
                print("checkpoint:", saver.save(sess, os.path.join(out_dir, "tf1_ckpt", "model.ckpt")))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True, help="unmodified checkout of dahliau/DPDist")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--cases", nargs="*", default=sorted(tf1_case.CASES))
    args = ap.parse_args()
    for sub in ("", "models", "utils"):
        sys.path.insert(0, os.path.join(args.reference, sub))
    import tensorflow as tf
    if not tf.__version__.startswith("1."):
        raise SystemExit("the reference needs TensorFlow 1.14 / 1.15 (tf.contrib); found %s" % tf.__version__)
    import dpdist_and_aue as MODEL
    for name in args.cases:
        run_case(tf, MODEL, name, args.out, save_ckpt=(name == "anchor"))


if __name__ == "__main__":
    main()
