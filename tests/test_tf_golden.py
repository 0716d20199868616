"""Slots for vectors produced by the REAL reference under TensorFlow 1.x (tools/make_tf_golden.py, run off-box against an
unmodified checkout of dahliau/DPDist).  When tests/golden/tf1_<case>.npz / tests/golden/tf1_ckpt/ exist these tests
pin the CPU oracle, the checkpoint reader and (with -m gpu) the CUDA path against TF1 itself; while they are absent the
tests SKIP with the reason "parity unpinned", which is the honest status of every parity claim in this repository.
The case generator (numpy legacy RandomState, shared with the TF-side script) is pinned here unconditionally."""
import os
import sys

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLDEN)
import tf1_case  # noqa: E402

from oracle import dpdist_oracle as O  # noqa: E402

UNPINNED = "parity unpinned: no TF1 vectors under tests/golden (run tools/make_tf_golden.py where TensorFlow 1.15 exists)"
CHECKSUMS = {   # sum |inputs|, sum |variables| in float64, number of variables
    "anchor": (173.19715353939682, 1289525.6357412045, 8),
    "batch4": (655.6460178337584, 1289638.8630086218, 8),
    "batch4_bn": (657.1286211311817, 1296025.2048518918, 24),
    "g5k3": (248.752915489953, 287662.67679995426, 8),
}


def _oracle(name):
    pairs, n, emb, k, sigma, H, bn = tf1_case.CASES[name]
    pcA, pcB, labels = tf1_case.inputs(name)
    var = {k_: torch.tensor(v) for k_, v in tf1_case.variables(name).items()}
    with O.tf_cpu_numerics():
        pred, aux, _ = O.get_model(torch.tensor(pcA), torch.tensor(pcB), var, Embedding_Size=emb, k=k, sigma3dmfv=sigma,
                                   bn=("train" if bn else False))
        loss, loss_pred = O.get_loss(pred, {}, torch.tensor(labels))
    return pred, aux, float(loss), float(loss_pred)


@pytest.mark.parametrize("name", sorted(tf1_case.CASES))
def test_case_generator_is_pinned(name):
    a, b, l = tf1_case.inputs(name)
    v = tf1_case.variables(name)
    s_in = float(np.abs(a).astype(np.float64).sum() + np.abs(b).astype(np.float64).sum() + l.astype(np.float64).sum())
    s_var = float(sum(np.abs(x).astype(np.float64).sum() for x in v.values()))
    want = CHECKSUMS[name]
    assert abs(s_in - want[0]) <= 1e-9 * want[0] and abs(s_var - want[1]) <= 1e-9 * want[1] and len(v) == want[2]
    assert all(x.dtype == np.float32 for x in list(v.values()) + [a, b, l])


@pytest.mark.parametrize("name", ["anchor", "g5k3"])
def test_oracle_runs_on_the_cases_and_uses_the_output_range(name):
    pred, _, loss, _ = _oracle(name)
    out = torch.cat([pred["pred_listAB"], pred["pred_listBA"]]).numpy()
    assert np.isfinite(out).all() and out.min() >= 0.0 and out.max() <= 2.0
    inside = out[..., 0][out[..., 0] > 0]
    assert inside.size > 0 and inside.std() > 0.05 and np.isfinite(loss)      # not saturated, not dead


@pytest.mark.parametrize("name", sorted(tf1_case.CASES))
def test_oracle_matches_tf1_vectors(name):
    path = os.path.join(GOLDEN, "tf1_%s.npz" % name)
    if not os.path.exists(path):
        pytest.skip(UNPINNED)
    if tf1_case.CASES[name][6] and not getattr(O, "BN_TRAINING", False):
        pytest.skip("oracle has no training-mode batch norm")
    z = np.load(path)
    a, b, l = tf1_case.inputs(name)
    assert np.array_equal(z["pcA"], a) and np.array_equal(z["pcB"], b) and np.array_equal(z["labels"], l)
    pred, aux, loss, loss_pred = _oracle(name)
    for key, got in (("output1", pred["pred_listAB"]), ("output2", pred["pred_listBA"])):
        want = z[key]
        assert (np.abs(got.numpy() - want) <= 1e-4 * np.abs(want) + 1e-5).all(), key
    for key, got in (("fvA", aux["fvA"]), ("fvB", aux["fvB"])):
        want = z[key]
        assert (np.abs(got.numpy() - want) <= 1e-4 * np.abs(want) + 2e-6).all(), key
    assert abs(loss - float(z["loss_samples"])) <= 1e-5 and abs(loss_pred - float(z["loss_pred"])) <= 1e-5


def test_checkpoint_reader_reads_a_tf_written_bundle():
    prefix = os.path.join(GOLDEN, "tf1_ckpt", "model.ckpt")
    if not os.path.exists(prefix + ".index"):
        pytest.skip(UNPINNED.replace("vectors", "checkpoint"))
    from dpdist_b200 import tf_checkpoint
    got = tf_checkpoint.load_checkpoint(prefix)
    for name, want in tf1_case.variables("anchor").items():
        assert name in got and np.array_equal(np.asarray(got[name]), want), name


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["anchor", "batch4", "g5k3"])
def test_cuda_path_matches_tf1_vectors(name):
    path = os.path.join(GOLDEN, "tf1_%s.npz" % name)
    if not os.path.exists(path):
        pytest.skip(UNPINNED)
    from dpdist_b200 import dpdist_and_aue as MODEL, tf_util
    z = np.load(path)
    pairs, n, emb, k, sigma, H, bn = tf1_case.CASES[name]
    store = tf_util.VariableStore(device="cuda:0")
    store.load_state_dict(tf1_case.variables(name), strict=False)
    a, b, _ = tf1_case.inputs(name)
    with tf_util.use_store(store):
        pred, _, emb_set = MODEL.get_model(torch.tensor(a, device="cuda:0"), torch.tensor(b, device="cuda:0"), False, bn=0,
                                           Embedding_Size=emb, k=k, sigma3dmfv=sigma, localSNmlp=[H, H, H], reuse=True)
    for key, got in (("output1", pred["pred_listAB"]), ("output2", pred["pred_listBA"])):
        want = z[key]
        assert (np.abs(got.cpu().numpy() - want) <= 1e-4 * np.abs(want) + 1e-5).all(), key
    want = z["fvA"]
    assert (np.abs(emb_set["embedding_A"].fv.cpu().numpy() - want) <= 1e-4 * np.abs(want) + 2e-6).all()
