"""ConvolutionalModel (train / predict / save / restore) on the GPU against the oracle pipeline.

End-to-end gate of BASELINE.json: predicted road masks in >= 99.5 % pixel agreement with the
reference path and patch-level F1 within 0.002, on identical weights and synthetic inputs.  The
weights come from a short training run of the engine itself on a learnable synthetic task (blob
masks), so that probabilities are confident like a real model's rather than ~0.5 everywhere.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import images_oracle as IO
from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu
FLAGSHIP_TRAIN_STEPS = 60


def blobs(rs, n, size, cells=6):
    low = rs.rand(n, 1, cells, cells).astype(np.float32)
    up = F.interpolate(torch.tensor(low), size=(size, size), mode="bilinear", align_corners=False)
    return (up.numpy()[:, 0] > 0.55).astype(np.float32)


def make_data(rs, n, size):
    lab = blobs(rs, n, size)
    img = np.stack([0.25 + 0.5 * lab + 0.1 * rs.randn(n, size, size) for _ in range(3)], -1)
    img[..., 1] = 0.5 + 0.1 * rs.randn(n, size, size)
    return np.clip(img, 0, 1).astype(np.float32), lab


def make_model(tmp_path, **kw):
    from road_segmentation_unet_b200 import tf_aerial_images as tfa
    opts = tfa.Options()
    opts.num_layers, opts.root_size, opts.dilated_layers = 3, 64, True
    opts.patch_size, opts.batch_size, opts.stride = 36, 4, 12
    opts.dropout, opts.lr, opts.momentum = 1.0, 0.02, 0.9
    opts.ensemble_prediction = True
    opts.save_path = str(tmp_path)
    opts.logdir = str(tmp_path / "logdir")
    opts.eval_every = opts.train_score_every = 10 ** 9
    for k, v in kw.items():
        setattr(opts, k, v)
    return tfa.ConvolutionalModel(opts, None), opts


def oracle_predict(imgs, params, opts, S):
    """The reference's predict() (tf_aerial_images.py:271-328) restated with the CPU oracle."""
    x = IO.image_augmentation_ensemble(imgs) if opts.ensemble_prediction else imgs
    n_img = x.shape[0]
    off = (S - opts.patch_size) // 2
    patches = IO.extract_patches(IO.mirror_border(x, off), S, stride=opts.stride,
                                 predict_patch_size=opts.patch_size)
    tp = O.to_torch(params)
    preds = []
    with torch.no_grad():
        step = 16 if S <= 256 else 3  # 764^2 windows: ~2 GB of fp32 activations per patch
        for k in range(0, patches.shape[0], step):
            logits = O.forward(torch.tensor(patches[k:k + step], dtype=torch.float32), tp,
                               opts.num_layers, opts.root_size, opts.dilated_layers)
            preds.append(torch.softmax(logits, dim=3)[..., 1].numpy())
    preds = np.concatenate(preds).astype(np.float64)
    per_img = patches.shape[0] // n_img
    masks = IO.images_from_patches(preds.reshape(n_img, per_img, opts.patch_size, opts.patch_size, 1),
                                   stride=opts.stride)
    if opts.ensemble_prediction:
        masks = IO.invert_image_augmentation_ensemble(masks)
    return masks


def test_train_then_predict_parity(tmp_path):
    model, opts = make_model(tmp_path)
    S = model.input_size
    assert S == 76
    off = (S - opts.patch_size) // 2
    rs = np.random.RandomState(0)
    losses = []
    for step in range(30):
        img, lab = make_data(rs, opts.batch_size, S)
        loss, probs = model.train_batch(img, lab[:, off:off + 36, off:off + 36])
        losses.append(loss)
        assert probs.shape == (4, 36, 36)
    assert model.global_step == 30
    assert np.mean(losses[-5:]) < 0.35 < losses[0], losses  # the engine learns the task

    params = model.net.state_dict()
    imgs, labels = make_data(np.random.RandomState(1), 2, 84)  # (84 + 40 - 76) % 12 == 0 -> 5x5 patches
    masks = model.predict(imgs)
    assert masks.shape == (2, 84, 84, 1) and masks.dtype == np.float64
    ref = oracle_predict(imgs, params, opts, S)
    assert ref.shape == masks.shape
    agree = float(((masks > 0.5) == (ref > 0.5)).mean())
    f1_dev, f1_ref = IO.patch_f1(masks, labels[..., None]), IO.patch_f1(ref, labels[..., None])
    print("pixel agreement %.5f  max|dp| %.4f  F1 device %.4f  oracle %.4f"
          % (agree, np.abs(masks - ref).max(), f1_dev, f1_ref))
    assert agree >= 0.995
    assert abs(f1_dev - f1_ref) <= 0.002
    assert f1_ref > 0.5  # the synthetic task was actually learnt
    assert np.abs(masks - ref).mean() < 5e-3

    # without the ensemble and through predict_batchwise
    opts.ensemble_prediction = False
    m2 = model.predict_batchwise(imgs, 1)
    r2 = oracle_predict(imgs, params, opts, S)
    assert m2.shape == (2, 84, 84, 1)
    assert float(((m2 > 0.5) == (r2 > 0.5)).mean()) >= 0.995

    # checkpoint round trip: reference file naming, momentum slots and global_step included
    path = model.save(3)
    assert path.endswith("model-epoch-003.chkpt")
    model2, _ = make_model(tmp_path, ensemble_prediction=False)
    model2.restore(file=path)
    assert model2.global_step == 30
    assert np.array_equal(model2.predict(imgs), m2)
    for k, v in model.net.state_dict("momentum").items():
        assert np.array_equal(model2.net.state_dict("momentum")[k], v), k
    model3, _ = make_model(tmp_path, ensemble_prediction=False)
    model3.experiment_name = "zzz-other"
    model3.restore()  # latest directory / latest epoch resolution (tf_aerial_images.py:351-379)
    assert np.array_equal(model3.predict(imgs), m2)


def test_train_epoch_semantics(tmp_path):
    """train(): labels binarised at 0.5, shuffled indices, range(0, n - B, B) drops the last batch
    even when n is divisible (tf_aerial_images.py:221-232); augmentation keeps images and masks
    aligned (both transformed by the same dihedral element)."""
    model, opts = make_model(tmp_path, image_augmentation=True, dropout=0.8)
    S = model.input_size
    off = (S - 36) // 2
    rs = np.random.RandomState(2)
    img, lab = make_data(rs, 12, S)
    model.train(img.astype(np.float64), lab[:, off:off + 36, off:off + 36] * 0.9, img, lab)
    assert model.global_step == len(range(0, 12 - 4, 4)) == 2
    assert len(model.scalars) == 2 and all(np.isfinite(s[1]) for s in model.scalars)
    x = torch.tensor(img[:4]).cuda()
    y = torch.tensor((lab[:4, off:off + 36, off:off + 36] > 0.5).astype(np.uint8)).cuda()
    xa, ya = model.stochastic_images_augmentation(x, y)
    # recover the op from the image, then the mask must have moved identically
    for i in range(4):
        ops_i = [op for op in range(8) if np.array_equal(IO.d4(x[i].cpu().numpy(), op), xa[i].cpu().numpy())]
        assert len(ops_i) == 1
        assert np.array_equal(IO.d4(y[i].cpu().numpy(), ops_i[0]), ya[i].cpu().numpy())


def test_prefetch_is_transparent(tmp_path):
    """Staging the next batch with prefetch() (what the epoch loop and bench.py's end-to-end leg do)
    must not change the training trajectory: two models, same seed, same batches, one fed through
    prefetch + train_batch, one through train_batch alone.  (Equality up to the summation order of
    the split-K fp32 atomics, which differs from run to run.)"""
    rs = np.random.RandomState(5)
    a, opts = make_model(tmp_path)
    b, _ = make_model(tmp_path)
    S, P, B = a.input_size, opts.patch_size, opts.batch_size
    batches = []
    for _ in range(4):
        x = rs.rand(B, S, S, 3).astype(np.float32)
        y = (rs.rand(B, P, P) < 0.3).astype(np.float64)
        batches.append((x, y))
    pinned = [(torch.from_numpy(x).pin_memory(), torch.from_numpy(y.astype(np.uint8)).pin_memory())
              for x, y in batches]
    la, lb = [], []
    a.prefetch(*pinned[0])
    for i in range(4):
        if i + 1 < 4:
            assert a.prefetch(*pinned[i + 1])
        la.append(a.train_batch(*pinned[i])[0])
    for x, y in batches:  # plain NumPy batches, no prefetch
        lb.append(b.train_batch(x, y)[0])
    assert np.allclose(la, lb, rtol=1e-4, atol=0)
    pa, pb = a.net.state_dict(), b.net.state_dict()
    for k in pa:
        den = max(np.linalg.norm(pb[k]), 1e-12)
        assert np.linalg.norm(pa[k] - pb[k]) / den < 1e-3, k
    # a third prefetch while two batches are staged is refused, and training still works
    a.prefetch(*pinned[0])
    a.prefetch(*pinned[1])
    assert a.prefetch(*pinned[2]) is False
    a.train_batch(*pinned[2])  # not staged: drops the stale entries and stages synchronously
    assert a.train_batch(*pinned[0])[0] > 0


def test_main_train_eval_submission(tmp_path):
    """The command-line flow of tf_aerial_images.main (tf_aerial_images.py:382-461) on PNG files:
    train one epoch, dump the training-set evaluation images, then a submission run that writes
    overlays, submission.csv and the model copy."""
    from PIL import Image
    from road_segmentation_unet_b200 import images, tf_aerial_images as tfa
    rs = np.random.RandomState(11)
    imgs, labs = make_data(rs, 4, 48)
    for sub in ("train/images", "train/groundtruth", "test"):
        (tmp_path / sub).mkdir(parents=True)
    for i in range(4):
        Image.fromarray(images.img_float_to_uint8(imgs[i]), "RGB").save(tmp_path / "train/images" / ("s_%02d.png" % i))
        Image.fromarray(images.img_float_to_uint8(labs[i]), "L").save(tmp_path / "train/groundtruth" / ("s_%02d.png" % i))
        Image.fromarray(images.img_float_to_uint8(imgs[i]), "RGB").save(tmp_path / "test" / ("t_%02d.png" % i))
    common = ["--num_layers=3", "--dilated_layers", "--patch_size=36", "--batch_size=4", "--stride=12",
              "--dropout=1.0", "--gpu=0", "--seed=3", "--pred_batch_size=2",
              "--save_path=" + str(tmp_path / "runs"), "--train_data_dir=" + str(tmp_path / "train")]
    model = tfa.main(common + ["--num_epoch=1", "--rotation_angles=0,90", "--eval_train",
                               "--eval_data_dir=" + str(tmp_path / "dump")])
    names = sorted(p.name for p in (tmp_path / "dump").iterdir())
    assert len(names) == 20 and names[0] == "eval_binary_pred_001.png" and "eval_orror_004.png" in names
    conf = images.load(str(tmp_path / "dump"))  # all five dumps are RGBA files of the image size
    assert conf.shape == (20, 48, 48, 4)
    run_dir = tmp_path / "runs" / model.experiment_name
    # a TensorFlow V2 bundle under the reference's naming, plus the .meta marker restore() globs for
    for ext in (".index", ".data-00000-of-00001", ".meta"):
        assert (run_dir / ("model-epoch-000.chkpt" + ext)).exists()

    model2 = tfa.main(common + ["--num_epoch=0", "--restore_model", "--eval_data_dir=" + str(tmp_path / "test")])
    out_dir = tmp_path / "runs" / model2.experiment_name
    rows = (out_dir / "submission.csv").read_text().splitlines()
    assert rows[0] == "id,prediction" and len(rows) == 1 + 4 * 9
    assert rows[1].startswith("001_0_0,") and rows[4].startswith("001_16_0,") and rows[-1].startswith("004_32_32,")
    assert sorted(p.name for p in out_dir.glob("images_*.png")) == ["images_%03d.png" % i for i in range(1, 5)]
    assert (tmp_path / "runs" / (model2.experiment_name + "-model.chkpt.index")).exists()
    # the restored model predicts what the trained one does, and the csv carries the 16 x 16 vote
    test_imgs = images.load(str(tmp_path / "test"))
    masks = model2.predict_batchwise(test_imgs, 2)
    assert np.abs(masks - model.predict_batchwise(test_imgs, 2)).max() == 0.0
    lab = images.patch_labels(images.quantize_mask(masks, 0.25, 16), 16)
    assert [int(r.split(",")[1]) for r in rows[1:]] == lab.reshape(-1).tolist()


@pytest.mark.parametrize("layers,patch,size,cap", [(3, 36, 72, 1400), (4, 36, 84, 1400), (4, 36, 84, 150)])
def test_shared_windows_equal_window_loop(tmp_path, layers, patch, size, cap):
    """predict() with aligned windows evaluated once (shared_window_plan) against the
    reference's loop of one forward pass per window, same weights.  A window output inside an
    enlarged window is the same dot products over the same receptive field; layer sizes differ,
    so a layer may run in the other convolution kernel (another fp32 summation order) and a few
    bf16 activations round the other way: observed <= 2e-5 on the probabilities."""
    model, opts = make_model(tmp_path, num_layers=layers, patch_size=patch, ensemble_prediction=True)
    opts.shared_window_max_input = cap
    rs = np.random.RandomState(5)
    imgs, _ = make_data(rs, 3, size)
    assert model.input_size + 0 == {3: 76, 4: 124}[layers]
    opts.shared_windows = True
    shared = model.predict(imgs)
    assert getattr(model, "_shared_nets", None), "the shared-window path did not run"
    opts.shared_windows = False
    loop = model.predict(imgs)
    assert shared.shape == loop.shape == (3, size, size, 1)
    assert np.abs(shared - loop).max() <= 2e-4 and np.abs(shared - loop).mean() <= 1e-5


def test_cuda_graph_steps_match_eager(tmp_path):
    """train_batch replayed from CUDA graphs (small batches) against the same steps launched
    eagerly: same losses, same weights (up to the summation order of the fp32 atomics in the
    weight gradients), same step counter -- across a change of the staircase learning rate,
    which re-captures the graphs."""
    eager, o_e = make_model(tmp_path / "e")
    graph, o_g = make_model(tmp_path / "g")
    o_e.cuda_graphs, o_g.cuda_graphs = "0", "1"
    B, S, P = o_e.batch_size, eager.input_size, o_e.patch_size
    rs = np.random.RandomState(21)
    for m in (eager, graph):
        m.net.global_step = 997
    lrs = []
    for step in range(6):
        x = rs.rand(B, S, S, 3).astype(np.float32)
        lab = (rs.rand(B, P, P) > 0.7).astype(np.float32)
        le, pe = eager.train_batch(x, lab)
        lg, pg = graph.train_batch(x, lab)
        assert abs(le - lg) <= 5e-3 * abs(le), (step, le, lg)
        assert np.abs(pe - pg).max() <= 2e-2
        lrs.append(graph.scalars[-1][2])
    assert graph._graphs is not None and getattr(eager, "_graphs", None) is None
    assert eager.global_step == graph.global_step == 1003
    assert lrs[:3] == [o_g.lr] * 3 and all(abs(v - o_g.lr * 0.95) < 1e-12 for v in lrs[3:])
    for name in eager.net.live_variables():
        a, b = eager.net.var(name).cpu().numpy(), graph.net.var(name).cpu().numpy()
        assert np.linalg.norm(a - b) <= 1e-3 * np.linalg.norm(a) + 1e-5, name


def test_stochastic_images_augmentation(tmp_path):
    """tf_aerial_images.py:173-210: per sample three fair coins each applying flip_up_down (:188),
    then rot90 by floor(4U) -- the same dihedral element on the image and on its mask, bit-exact
    per element, all 8 elements reachable, roughly uniform."""
    model, opts = make_model(tmp_path, batch_size=8)
    S, P = model.input_size, opts.patch_size
    rs = np.random.RandomState(13)
    x = rs.rand(8, S, S, 3).astype(np.float32)
    m = (rs.rand(8, P, P) > 0.5).astype(np.uint8)
    xd, md = torch.tensor(x).cuda(), torch.tensor(m).cuda()
    seen = {}
    for _ in range(40):
        state = model._aug_rng.get_state()
        xa, ma = model.stochastic_images_augmentation(xd, md)
        replay = np.random.RandomState()
        replay.set_state(state)  # the draws the call consumed: 3 x B coins, then B rotations
        coins = replay.random_sample((3, 8)) > 0.5
        flip = coins[0] ^ coins[1] ^ coins[2]
        k = np.floor(replay.random_sample(8) * 4).astype(int)
        xa, ma = xa.cpu().numpy(), ma.cpu().numpy()
        for b in range(8):
            assert np.array_equal(xa[b], O.d4_apply(x[b], flip[b], k[b]))
            assert np.array_equal(ma[b], O.d4_apply(m[b], flip[b], k[b]))
            seen[(bool(flip[b]), int(k[b]))] = seen.get((bool(flip[b]), int(k[b])), 0) + 1
    assert len(seen) == 8 and min(seen.values()) >= 15, seen  # 320 draws, 40 expected per element


def test_flagship_predict_parity(tmp_path):
    """BASELINE.json configs[2]'s model (run.py:122-132: 6 layers, dilated, 388^2 patches, 764^2
    windows) through ConvolutionalModel.predict against the oracle pipeline (the reference's
    predict(), tf_aerial_images.py:271-328, restated on the CPU): >= 99.5 % of the thresholded
    pixels equal, patch-level F1 within 0.002.
      * a 400^2 image at stride 12 with the 6-way ensemble -- the reference's training-image case,
        4 windows x 6 variants, one forward pass per window;
      * a 452^2 image at stride 32 (3 x 3 windows whose origins differ by the pooling period 32):
        the shared-window path (one 828^2 pass covers all nine) AND the window loop, both against
        the oracle's nine separate passes."""
    model, opts = make_model(tmp_path, num_layers=6, patch_size=388, batch_size=2, stride=12, lr=0.01)
    S = model.input_size
    assert S == 764
    off = (S - 388) // 2
    rs = np.random.RandomState(0)
    losses = []
    for step in range(FLAGSHIP_TRAIN_STEPS):
        lab = blobs(rs, 2, S, cells=12)
        img = np.stack([0.25 + 0.5 * lab + 0.1 * rs.randn(2, S, S) for _ in range(3)], -1)
        img[..., 1] = 0.5 + 0.1 * rs.randn(2, S, S)
        img = np.clip(img, 0, 1).astype(np.float32)
        losses.append(model.train_batch(img, lab[:, off:off + 388, off:off + 388])[0])
    print("flagship training loss: first %.3f last %.3f" % (losses[0], np.mean(losses[-5:])))
    assert np.mean(losses[-5:]) < 0.4 < losses[0], losses
    params = model.net.state_dict()

    def data(seed, size):
        r = np.random.RandomState(seed)
        lab = blobs(r, 1, size, cells=8)
        img = np.stack([0.25 + 0.5 * lab + 0.1 * r.randn(1, size, size) for _ in range(3)], -1)
        img[..., 1] = 0.5 + 0.1 * r.randn(1, size, size)
        return np.clip(img, 0, 1).astype(np.float32), lab

    def gate(tag, masks, ref, labels):
        agree = float(((masks > 0.5) == (ref > 0.5)).mean())
        f1_dev, f1_ref = IO.patch_f1(masks, labels[..., None]), IO.patch_f1(ref, labels[..., None])
        print("%s: pixel agreement %.5f  max|dp| %.4f  mean|dp| %.5f  F1 device %.4f  oracle %.4f"
              % (tag, agree, np.abs(masks - ref).max(), np.abs(masks - ref).mean(), f1_dev, f1_ref))
        assert masks.shape == ref.shape and masks.dtype == np.float64
        assert agree >= 0.995, (tag, agree)
        assert abs(f1_dev - f1_ref) <= 0.002, (tag, f1_dev, f1_ref)
        assert f1_ref > 0.5, (tag, f1_ref)  # a confident model, not p ~ 0.5 everywhere

    # (1) 400^2, stride 12, ensemble: 2 x 2 windows per variant, nothing to share
    imgs, labels = data(1, 400)
    opts.ensemble_prediction, opts.stride, opts.shared_windows = True, 12, True
    gate("400^2 stride 12 ensemble", model.predict(imgs), oracle_predict(imgs, params, opts, S), labels)

    # (2) 452^2, stride 32, no ensemble: shared windows on and off against the same oracle masks
    imgs, labels = data(2, 452)
    opts.ensemble_prediction, opts.stride = False, 32
    ref = oracle_predict(imgs, params, opts, S)
    model.__dict__.pop("_shared_nets", None)
    opts.shared_windows = True
    shared = model.predict(imgs)
    assert model.__dict__.get("_shared_nets"), "the shared-window path did not run"
    gate("452^2 stride 32 shared windows", shared, ref, labels)
    opts.shared_windows = False
    gate("452^2 stride 32 window loop", model.predict(imgs), ref, labels)


def test_eval_every_fires_streaming_metrics(tmp_path):
    """SURVEY 8(f) row 3 (tf_aerial_images.py:247-267, :428; summary.py:104-147): the periodic
    evaluation inside train() predicts num_eval_images images and feeds STREAMING patch metrics
    (counts accumulate over evaluations, reset once per epoch); train_score_every does the same
    on the whole training set; loss / learning rate / misclassification are logged every step."""
    from road_segmentation_unet_b200 import images
    from road_segmentation_unet_b200.summary import StreamingMetrics
    model, opts = make_model(tmp_path, eval_every=2, train_score_every=3, num_eval_images=2,
                             ensemble_prediction=False, logdir=str(tmp_path / "logs"))
    S = model.input_size
    off = (S - 36) // 2
    rs = np.random.RandomState(4)
    imgs, labs = make_data(rs, 3, 48)                       # (48 + 40 - 76) % 12 == 0; 16 | 48
    pat, plab = make_data(rs, 24, S)
    model._summary.reset()
    model.train(pat.astype(np.float64), plab[:, off:off + 36, off:off + 36], imgs, labs)
    steps = len(range(0, 24 - 4, 4))
    assert model.global_step == steps == 5
    tags = [t for t, _, _ in model._summary.scalars]
    assert tags.count("loss") == tags.count("learning_rate") == tags.count("misclassification_rate") == steps
    assert tags.count("eval f1_score") == 2 and tags.count("train f1_score") == 1   # steps 2, 4 / step 3
    # the streamed values equal tf.metrics semantics replayed on the masks the evaluations produced
    replay = StreamingMetrics()
    pred = images.patch_labels(model.last_eval_masks, 16).reshape(-1)
    true = images.patch_labels((labs[:2] >= 0.5) * 1., 16).reshape(-1)
    got = [v for t, v, s in model._summary.scalars if t == "eval accuracy"]
    assert len(got) == 2 and 0.0 <= got[-1] <= 1.0
    # (second evaluation: counts of both evaluations; replaying the last one twice bounds it)
    one = replay.update(true, pred, padded_zeros=pred.size * 255)
    assert model._summary.eval_metrics.total == 2 * replay.total
    assert abs(model.last_eval_scores[0] - got[-1]) < 1e-12
    # per-epoch reset (main does it before every epoch)
    model._summary.reset()
    assert model._summary.eval_metrics.total == 0 and model._summary.train_metrics.total == 0
    # the log file of rank 0 carries every scalar, image dumps exist
    import json
    rows = [json.loads(l) for l in open(tmp_path / "logs" / model.experiment_name / "scalars.jsonl")]
    assert len(rows) == len(model._summary.scalars)
    assert {r["tag"] for r in rows} >= {"loss", "learning_rate", "misclassification_rate", "eval accuracy",
                                        "eval recall", "eval precision", "eval f1_score", "train f1_score"}
    dumped = sorted(p.name for p in (tmp_path / "logs" / model.experiment_name / "images").iterdir())
    assert any(n.startswith("eval_masks_000002") for n in dumped)
    assert any(n.startswith("groundtruth_vs_prediction_000004") for n in dumped)


def test_sharded_train_prep_and_strict_restore(tmp_path):
    """(1) prepare_train_patches: the shards of all ranks, in rank order, are the reference's
    train-prep block (expand_and_rotate + extract_patches of images and ground truth) exactly.
    (2) restore() refuses checkpoints that do not match the architecture, and reads both the
    TensorFlow bundle save() writes and the earlier .npz payload."""
    from road_segmentation_unet_b200 import images, tf_aerial_images as tfa, tf_checkpoint
    model, opts = make_model(tmp_path)
    opts.rotation_angles = [0, 30, 90]
    rs = np.random.RandomState(8)
    imgs, labs = make_data(rs, 3, 48)
    S = model.input_size
    off = (S - 36) // 2
    full_p = images.extract_patches(images.expand_and_rotate(imgs, opts.rotation_angles, off),
                                    patch_size=S, predict_patch_size=36, stride=12)
    full_l = images.extract_patches(images.expand_and_rotate(labs, opts.rotation_angles, 0),
                                    patch_size=36, stride=12)
    p1, l1 = tfa.prepare_train_patches(imgs, labs, opts)
    assert np.array_equal(p1, full_p) and np.array_equal(l1, full_l)
    for world in (2, 4):
        parts = [tfa.prepare_train_patches(imgs, labs, opts, r, world) for r in range(world)]
        assert np.array_equal(np.concatenate([p for p, _ in parts]), full_p)
        assert np.array_equal(np.concatenate([l for _, l in parts]), full_l)
        assert max(p.shape[0] for p, _ in parts) < full_p.shape[0]

    path = model.save(1)
    assert tf_checkpoint.is_bundle(path)
    other, o2 = make_model(tmp_path, num_layers=4)
    with pytest.raises(ValueError, match="does not match the model"):
        other.restore(file=path)
    t = tf_checkpoint.read_bundle(path)
    assert int(t["global_step"]) == 0 and "conv_0/conv1/kernel/Momentum" in t
    t["conv_0/conv2/bias"] = np.zeros(32, np.float32)
    tf_checkpoint.write_bundle(str(tmp_path / "bad.chkpt"), t)
    with pytest.raises(ValueError, match="conv_0/conv2/bias"):
        model.restore(file=str(tmp_path / "bad.chkpt"))
    with pytest.raises(FileNotFoundError, match="index"):
        model.restore(file=str(tmp_path / "nothing.chkpt"))
    # the round-1 .npz payload still restores
    t = tf_checkpoint.read_bundle(path)
    with open(tmp_path / "old.chkpt", "wb") as f:
        np.savez(f, **t)
    model.restore(file=str(tmp_path / "old.chkpt"))


def test_reference_submission_files_reproduced(tmp_path):
    """Last steps of the reference's `main` (tf_aerial_images.py:453-458) on the reference's own
    committed outputs (tests/golden/make_submission_golden.py): the quantised masks located by its
    overlay PNGs go through quantize_mask (a fixed point), overlays and save_submission_csv here
    and give back its overlay pixels and its submission.csv files byte for byte."""
    import hashlib
    import os
    from road_segmentation_unet_b200 import images
    from road_segmentation_unet_b200.constants import FOREGROUND_THRESHOLD, IMG_PATCH_SIZE
    S = np.load(os.path.join(os.path.dirname(__file__), "golden", "submission_golden.npz"))
    side = S["crop_rgb"].shape[1]
    for r in range(S["crop_overlay"].shape[0]):
        masks = np.kron(S["labels"][r].transpose(0, 2, 1), np.ones((16, 16), np.uint8)).astype(np.float64)
        masks = masks[..., None]
        q = images.quantize_mask(masks, patch_size=IMG_PATCH_SIZE, threshold=FOREGROUND_THRESHOLD)
        assert q.shape == masks.shape and np.array_equal(q, masks)
        out = tmp_path / ("run%d" % r)
        images.save_submission_csv(q, str(out), IMG_PATCH_SIZE)
        with open(out / "submission.csv", "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == str(S["csv_sha256"][r]), str(S["runs"][r])
        for j, (k, top, left) in enumerate(S["crops"]):
            img = S["crop_rgb"][j:j + 1].astype(np.float32) / np.float32(255)
            m = q[k:k + 1, top:top + side, left:left + side]
            assert np.array_equal(images.overlays(img, m, fade=0.4)[0], S["crop_overlay"][r, j])
