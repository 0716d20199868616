"""DPDist training driver on synthetic ModelNet-shaped data (the `--train_comp dpdist` branch of the
reference's train_multi_gpu_pc_compare_dist.py, :186-357, with the same flag names).

    python -m dpdist_b200.train_multi_gpu_pc_compare_dist --batch_size 16 --num_point 64 --steps 100
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        -m dpdist_b200.train_multi_gpu_pc_compare_dist --num_gpus N --batch_size 16

One process per GPU: rank i takes the slice [i*DEVICE_BATCH_SIZE, (i+1)*DEVICE_BATCH_SIZE) of every global
batch (:241-251), gradients are averaged with NCCL (:936-974), Adam runs on every rank.  The reference's
dataset files do not ship with it, so batches come from dpdist_b200.synthetic.dataset_batch (same shapes as
ModelNetDataset.next_batch, modelnet_dataset.py:170-187); the AUE / PCRNet consumers are out of scope.
"""
import argparse
import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from . import synthetic, train


def train_on_dataset(FLAGS, tr, dev, rank, world):
    """The reference's epoch loop on the real files (train_multi_gpu_pc_compare_dist.py:181-188, 332-357, 732-873):
    ModelNetDataset(npoints = 2 * NUM_POINT, class_choice = [category]) with augmentation, train_one_epoch_3d, an
    evaluation epoch on the test split every 10 epochs and at the end, `model.ckpt` saved next to the log."""
    from . import modelnet_dataset, tf_checkpoint
    # every rank reads the same global batch (same seed) and keeps its own slice, like the towers' tf.slice (:241-251)
    mk = lambda split: modelnet_dataset.ModelNetDataset(root=FLAGS.data_root, npoints=FLAGS.num_point * 2, split=split,
                                                        normal_channel=False, batch_size=FLAGS.batch_size,
                                                        class_choice=[FLAGS.category] if FLAGS.category else None,
                                                        device=dev, seed=0)
    TRAIN_DATASET, TEST_DATASET = mk('train'), mk('test')

    def batches(ds, augment):
        # "Make sure batch data is of same size" (:735-739, :815-817): the reference feeds fixed BATCH_SIZE buffers and
        # overwrites only the first `bsize` rows, so a short last batch is completed by the rows the previous batch left
        # there (zeros in the very first batch).  Same here, which also keeps every batch divisible by the towers.
        cur = None
        while ds.has_next_batch():
            pcA, pcB, lab = ds.next_batch_device(FLAGS.num_point, augment=augment)
            bsize = pcA.shape[0]
            if cur is None:
                cur = [torch.zeros((FLAGS.batch_size,) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype) for t in (pcA, pcB, lab)]
            for buf, t in zip(cur, (pcA, pcB, lab)):
                buf[:bsize] = t
            yield [train.shard(x, rank, world).contiguous() for x in cur]
        ds.reset()

    def eval_one_epoch():
        from . import dpdist_and_aue as MODEL, tf_util
        tot, cnt = torch.zeros((), device=dev), 0
        with torch.no_grad():
            for a, b, l in batches(TEST_DATASET, False):
                with tf_util.use_store(tr.store):
                    pred, _, _ = MODEL.get_model(a, b, False, **tr.kw)
                tot += (pred['pred_listAB'][:, :, 0, 0] - l).abs().mean()                  # :840-852 loss_samples
                cnt += 1
        t = torch.stack([tot, torch.tensor(float(cnt), device=dev)])
        if world > 1:       # the towers' losses are averaged (:297), so the figure does not depend on the number of GPUs
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t[0] / t[1].clamp_min(1.0))

    for epoch in range(FLAGS.max_epoch):
        losses = []
        for a, b, l in batches(TRAIN_DATASET, True):
            noise = torch.randn_like(a) * FLAGS.add_noise if FLAGS.add_noise > 0 else 0
            losses.append(tr.step(a, b, l, add_noise=noise))
        mean_loss = float(torch.stack(losses).mean()) if losses else float('nan')
        if world > 1:
            t = torch.tensor([mean_loss], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.AVG)
            mean_loss = float(t.item())
        if rank == 0:
            print(' ---- epoch: %03d ---- mean loss: %f' % (epoch + 1, mean_loss), flush=True)
        if epoch % 10 == 0 or epoch == FLAGS.max_epoch - 1:                              # :354-357
            ev = eval_one_epoch()
            if rank == 0:
                os.makedirs(FLAGS.log_dir, exist_ok=True)
                path = tf_checkpoint.save_checkpoint(os.path.join(FLAGS.log_dir, 'model.ckpt'), tr.store.state_dict())
                print('eval mean loss: %f   Model saved in file: %s' % (ev, path), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument('--num_gpus', type=int, default=1)
    p.add_argument('--num_point', type=int, default=64)
    p.add_argument('--batch_size', type=int, default=16)
    p.add_argument('--learning_rate_dpdist', type=float, default=0.0001)
    p.add_argument('--decay_step', type=int, default=300 * 512)
    p.add_argument('--decay_rate', type=float, default=0.5)
    p.add_argument('--embedding_size', type=int, default=8 ** 3)
    p.add_argument('--K', default='5')
    p.add_argument('--sigma3dmfv', type=float, default=2.0)
    p.add_argument('--add_noise', type=float, default=0.0)
    p.add_argument('--implicit_net_type', default='1', help='1: shared-MLP head, 3: 3-D CNN head (reference flag, :65,112)')
    p.add_argument('--BN', default='0', help='1: batch norm after every conv of the head (reference flag, :61,105)')
    p.add_argument('--steps', type=int, default=50, help='training steps to run (synthetic data has no epochs)')
    p.add_argument('--warmup', type=int, default=5)
    p.add_argument('--distinct_batches', type=int, default=4)
    p.add_argument('--log_every', type=int, default=10)
    # real data (reference flags: --category chair, --max_epoch, --log_dir; train_multi_gpu_pc_compare_dist.py:41-69)
    p.add_argument('--data_root', default='', help='ModelNet root with ground-truth files (dpdist_b200.dataset_sample_with_gt); '
                                                   'empty = synthetic batches')
    p.add_argument('--category', default='chair')
    p.add_argument('--max_epoch', type=int, default=1)
    p.add_argument('--log_dir', default='log/dpdist_b200')
    p.add_argument('--cuda_graph', type=int, default=0, help='replay the captured training step (single GPU)')
    FLAGS = p.parse_args(argv)

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert FLAGS.batch_size % world == 0                                    # :125
    sigma = FLAGS.sigma3dmfv * 0.0625                                       # :103
    tr = train.DPDistTrainer(dev, base_lr=FLAGS.learning_rate_dpdist, decay_step=FLAGS.decay_step,
                             decay_rate=FLAGS.decay_rate, Embedding_Size=FLAGS.embedding_size, k=int(FLAGS.K),
                             sigma3dmfv=sigma, seed=1, cuda_graph=bool(FLAGS.cuda_graph), bn=int(FLAGS.BN),
                             conv_version=int(FLAGS.implicit_net_type))
    if FLAGS.data_root:
        return train_on_dataset(FLAGS, tr, dev, rank, world)
    batches = []
    for i in range(FLAGS.distinct_batches):
        if FLAGS.batch_size <= 64:
            pts, lab = synthetic.dataset_batch(100 + i, FLAGS.batch_size, FLAGS.num_point)
            pcA, pcB, lab_ab = train.assemble_batch(pts, lab, FLAGS.num_point)   # :749-766
        else:   # large synthetic batches: the chair sampler is a python loop, use the vectorised generator
            pcA, pcB, lab_ab = synthetic.uniform_batch(100 + i, FLAGS.batch_size, FLAGS.num_point)
        mine = [torch.from_numpy(np.ascontiguousarray(train.shard(x, rank, world))).to(dev) for x in (pcA, pcB, lab_ab)]
        batches.append(mine)

    def one(i):
        a, b, l = batches[i % len(batches)]
        noise = torch.randn_like(a) * FLAGS.add_noise if FLAGS.add_noise > 0 else 0      # :768-771
        return tr.step(a, b, l, add_noise=noise)

    for i in range(FLAGS.warmup):
        one(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    losses = []
    e0.record()
    for i in range(FLAGS.steps):
        losses.append(one(FLAGS.warmup + i))
        if rank == 0 and FLAGS.log_every and (i + 1) % FLAGS.log_every == 0:
            print(' ---- batch: %03d ---- mean loss: %f' % (i + 1, float(torch.stack(losses[-FLAGS.log_every:]).mean())), flush=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        print(json.dumps({"mode": "train", "n_gpus": world, "global_batch": FLAGS.batch_size, "num_point": FLAGS.num_point,
                          "steps": FLAGS.steps, "ms_per_step": ms / FLAGS.steps,
                          "pairs_per_s": FLAGS.batch_size * FLAGS.steps / (ms * 1e-3),
                          "first_loss": float(losses[0]), "last_loss": float(losses[-1])}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
