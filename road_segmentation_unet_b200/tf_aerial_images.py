"""Train / predict loop -- host-side mirror of the reference's src/tf_aerial_images.py.

Keeps the reference's flag names and defaults (tf_aerial_images.py:15-46), `Options` (:51-84) and
`ConvolutionalModel(options, session)` with `.train`, `.predict`, `.predict_batchwise`, `.save`,
`.restore`, `.input_size`, `.experiment_name` (:87-379).  `session` has no TensorFlow meaning any
more and is accepted and ignored.  One process drives one GPU; under torchrun (WORLD_SIZE > 1,
`--num_gpu` = world size) training is data parallel with one bucketed NCCL all-reduce of the flat
gradient per step, overlapped with the remaining backward kernels, and prediction shards the
patch list across ranks with a single final reduction of the per-rank partial mask sums.
"""
import argparse
import glob
import math
import os
import time
from datetime import datetime

import numpy as np
import torch

from . import images
from . import unet
from .summary import Summary
from .constants import NUM_CHANNELS, IMG_PATCH_SIZE, FOREGROUND_THRESHOLD


# ------------------------------------------------------------------ flags (names/defaults verbatim)
def _bool(v):
    if isinstance(v, bool):
        return v
    return str(v).lower() in ("1", "true", "t", "yes", "y")


FLAG_DEFS = [
    ("batch_size", int, 25, "Batch size of training instances"),
    ("dilated_layers", _bool, False, "Add dilated CNN layers"),
    ("dropout", float, 0.8, "Probability to keep an input"),
    ("ensemble_prediction", _bool, False, "Ensemble Prediction"),
    ("eval_data_dir", str, None, "Directory containing eval images"),
    ("eval_every", int, 500, "Number of steps between evaluations"),
    ("eval_train", _bool, False, "Evaluate training data"),
    ("gpu", int, -1, "GPU to run the model on"),
    ("image_augmentation", _bool, False, "Augment training set of images with transformations"),
    ("interactive", _bool, False, "Spawn interactive Tensorflow session"),
    ("logdir", str, os.path.abspath("./logdir"), "Directory where to write logfiles"),
    ("lr", float, 0.01, "Initial learning rate"),
    ("model_path", str, None, "Restore exact model path"),
    ("momentum", float, 0.9, "Momentum"),
    ("num_epoch", int, 5, "Number of pass on the dataset during training"),
    ("num_eval_images", int, 4, "Number of images to predict for an evaluation"),
    ("num_gpu", int, 1, "Number of available GPUs to run the model on"),
    ("num_layers", int, 5, "Number of layers of the U-Net"),
    ("patch_size", int, 128, "Size of the prediction image"),
    ("pred_batch_size", int, 2, "Batch size of batchwise prediction"),
    ("restore_date", str, None, "Restore the model from specific date"),
    ("restore_epoch", int, None, "Restore the model from specific epoch"),
    ("restore_model", _bool, False, "Restore the model from previous checkpoint"),
    ("root_size", int, 64, "Number of filters of the first U-Net layer"),
    ("rotation_angles", str, None, "Rotation angles"),
    ("save_path", str, os.path.abspath("./runs"),
     "Directory where to write checkpoints, overlays and submissions"),
    ("seed", int, 2017, "Random seed for reproducibility"),
    ("stride", int, 16, "Sliding delta for patches"),
    ("train_data_dir", str, os.path.abspath("./data/training"),
     "Directory containing training images/ groundtruth/"),
    ("train_score_every", int, 1000, "Compute training score after the given number of iterations"),
]


def make_parser():
    p = argparse.ArgumentParser(description="B200 U-Net road segmentation (reference flag surface)")
    for name, typ, default, helptext in FLAG_DEFS:
        if typ is _bool:
            p.add_argument("--" + name, type=_bool, nargs="?", const=True, default=default, help=helptext)
        else:
            p.add_argument("--" + name, type=typ, default=default, help=helptext)
    return p


FLAGS = make_parser().parse_args([])


class Options(object):
    """Options used by our model (tf_aerial_images.py:51-84)."""

    def __init__(self, flags=None):
        f = flags if flags is not None else FLAGS
        for name, _, _, _ in FLAG_DEFS:
            setattr(self, name, getattr(f, name))
        self.rotation_angles = None if not f.rotation_angles else \
            [int(i) for i in str(f.rotation_angles).split(",")]
        # not a reference flag: predict() evaluates aligned sliding windows once (see
        # shared_window_plan); RSU_SHARED_WINDOWS=0 runs every window as its own forward pass
        self.shared_windows = os.environ.get("RSU_SHARED_WINDOWS", "1") != "0"
        self.shared_window_max_input = int(os.environ.get("RSU_SHARED_WINDOW_MAX_INPUT", "1400"))
        # "auto": training steps of batch <= 8 are replayed from CUDA graphs; "0" / "1" force it
        self.cuda_graphs = os.environ.get("RSU_CUDA_GRAPHS", "auto")


# ------------------------------------------------------------------ distributed plumbing
class _Dist:
    """torch.distributed (NCCL) as plumbing: rank / world from the torchrun environment."""

    def __init__(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.active = self.world > 1
        if self.active:
            import torch.distributed as dist
            if not dist.is_initialized():
                backend = "nccl" if torch.cuda.is_available() else "gloo"
                dist.init_process_group(backend=backend)
            self.dist = dist
        self.comm_stream = None

    def barrier(self):
        if self.active:
            self.dist.barrier()


def shard_range(n_items, rank, world):
    """Contiguous slice [k0, k1) of a work list owned by `rank` (prediction patches)."""
    return n_items * rank // world, n_items * (rank + 1) // world


def shared_window_plan(side, stride, input_size, num_layers, max_input=1400):
    """Sliding windows of a valid-padding U-Net whose origins differ by a multiple of the pooling
    period 2^(L-1) compute identical features wherever they overlap.  Along one axis the `side`
    patch positions k*stride fall into period/gcd(stride, period) alignment classes; the members
    of a class lie q = lcm(stride, period) apart, so n of them are covered by ONE window of input
    size input_size + q*(n-1) whose output holds the n patch outputs q apart.
    Returns None when nothing can be shared, else (n, q, windows): windows = lists of the patch
    positions (along the axis) that one window covers; its origin is stride * window[0]."""
    period = 2 ** (num_layers - 1)
    classes = period // math.gcd(stride, period)
    q = stride * classes
    n = min(-(-side // classes), 1 + max(0, (max_input - input_size) // q))
    if n <= 1:
        return None
    windows = []
    for a in range(min(classes, side)):
        members = list(range(a, side, classes))
        windows.extend(members[b:b + n] for b in range(0, len(members), n))
    return n, q, windows


def shared_window_jobs(num_images, windows, rank=0, world=1):
    """The enlarged-window jobs (image, window along x, window along y) of one rank -- a contiguous
    slice of the image-major job list -- grouped by the number of window positions per side the
    job needs (the longer of its two windows), i.e. by the engine size that runs it."""
    jobs = [(img, wx, wy) for img in range(num_images) for wx in windows for wy in windows]
    j0, j1 = shard_range(len(jobs), rank, world)
    by_size = {}
    for job in jobs[j0:j1]:
        by_size.setdefault(max(len(job[1]), len(job[2])), []).append(job)
    return by_size


def shard_pairs(num_images, angles, rank=0, world=1):
    """The (angle, image) pairs of the training-set preparation (main, tf_aerial_images.py:403-423)
    owned by `rank`: a contiguous slice of expand_and_rotate's angle-major output order, so that
    concatenating the ranks' shards in rank order reproduces the unsharded tensor.  Returns
    [(angle, [image indices])] with the images of one angle grouped."""
    pairs = [(a, i) for a in angles for i in range(num_images)]
    k0, k1 = shard_range(len(pairs), rank, world)
    groups = []
    for a, i in pairs[k0:k1]:
        if groups and groups[-1][0] == a:
            groups[-1][1].append(i)
        else:
            groups.append((a, [i]))
    return groups


def prepare_train_patches(train_images, train_groundtruth, opts, rank=0, world=1):
    """The train-prep block of the reference's main (tf_aerial_images.py:403-423) for the
    (angle, image) pairs one rank owns: expand_and_rotate + extract_patches of the images (input
    size, offset border) and of the ground truth (patch size).  world = 1 is the reference's block
    verbatim; with more ranks every rank prepares only the patches it will train on (SURVEY.md
    8(e) row 3) and the union over ranks is the unsharded set, in the same order."""
    input_size = unet.input_size_needed(opts.patch_size, opts.num_layers)
    offset = int((input_size - opts.patch_size) / 2)
    pats, labs = [], []
    for angle, idx in shard_pairs(train_images.shape[0], opts.rotation_angles, rank, world):
        extended_images = images.expand_and_rotate(train_images[idx], [angle], offset)
        pats.append(images.extract_patches(extended_images, patch_size=input_size,
                                           predict_patch_size=opts.patch_size, stride=opts.stride))
        train_groundtruth_exp = images.expand_and_rotate(train_groundtruth[idx], [angle], 0)
        labs.append(images.extract_patches(train_groundtruth_exp, patch_size=opts.patch_size,
                                           stride=opts.stride))
    if not pats:
        return (np.zeros((0, input_size, input_size) + train_images.shape[3:]),
                np.zeros((0, opts.patch_size, opts.patch_size)))
    return np.concatenate(pats, axis=0), np.concatenate(labs, axis=0)


def rank_batch_indices(indices, offset, rank, batch_size):
    """The slice of the (identically shuffled) epoch permutation that `rank` trains on at the
    global step starting at `offset`: ranks take consecutive batch_size-sized pieces."""
    lo = offset + rank * batch_size
    return indices[lo:lo + batch_size]


class GradientAllReducer:
    """Bucketed all-reduce (sum) of the flat gradient, overlapped with the backward pass: as soon
    as a block's weight gradients are complete its slice is reduced on a side stream.  finish()
    waits for all buckets and returns the 1/world factor the optimizer applies (mean gradient =
    gradient of the global-batch mean loss, tf_aerial_images.py:108)."""

    def __init__(self, dist_module, world, flat_grads):
        self.dist = dist_module
        self.world = world
        self.g = flat_grads
        self.cuda = flat_grads.is_cuda
        self.stream = torch.cuda.Stream() if self.cuda else None
        self.pending = []
        self.reduced = []

    def bucket_ready(self, start, end):
        self.reduced.append((start, end))
        if not self.cuda:
            self.pending.append(self.dist.all_reduce(self.g[start:end], async_op=True))
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ev)
            self.pending.append(self.dist.all_reduce(self.g[start:end], async_op=True))

    def finish(self):
        for w in self.pending:
            w.wait()
        self.pending = []
        self.reduced = []
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.stream)
        return 1.0 / self.world


# ------------------------------------------------------------------ the model
class ConvolutionalModel:
    def __init__(self, options, session=None):
        self._options = options
        self._session = session  # accepted for signature parity, unused

        np.random.seed(options.seed)
        self.input_size = unet.input_size_needed(options.patch_size, options.num_layers)

        self.experiment_name = datetime.now().strftime("%Y-%m-%dT%Hh%Mm%Ss")
        self._dist = _Dist()
        if torch.cuda.is_available():
            torch.cuda.set_device(self._dist.local_rank if self._dist.active else
                                  (options.gpu if options.gpu >= 0 else 0))
        self._aug_rng = np.random.RandomState(options.seed + 7919 * self._dist.rank)
        self.scalars = []  # (step, loss, lr): what the reference sends to TensorBoard
        # rank 0 writes <logdir>/<experiment_name>/scalars.jsonl (+ image dumps); every rank keeps
        # the values in memory (tf_aerial_images.py:98-100)
        summary_path = os.path.join(options.logdir, self.experiment_name)
        self._summary = Summary(options, session, summary_path, write=self._dist.rank == 0)
        # True when train() is handed only this rank's shard of the patches (main does that under
        # torchrun): batches are then drawn from the local shard instead of a slice of the global
        # permutation
        self.sharded_data = False
        self.build_graph()

    # -- "graph": the planned engine --------------------------------------------------------
    def _make_peer_optimizer(self):
        """Data-parallel ranks on NVLink-connected GPUs exchange gradients and weights inside the
        optimizer kernel (dp.PeerOptimizer).  RSU_DP_MODE = peer | nccl | auto (default: peer when
        symmetric memory can be set up, else the bucketed NCCL all-reduce)."""
        mode = os.environ.get("RSU_DP_MODE", "auto")
        if not self._dist.active or not torch.cuda.is_available() or mode == "nccl":
            return None
        opts = self._options
        try:
            from .dp import PeerOptimizer
            _, n_flat = unet.flat_layout(opts.num_layers, opts.root_size, opts.dilated_layers)
            return PeerOptimizer(self._dist.dist, self._dist.world, self._dist.rank, n_flat)
        except Exception as e:  # no peer access / symmetric memory unavailable
            if mode == "peer":
                raise
            print("peer-memory optimizer unavailable ({}); using the NCCL all-reduce".format(e))
            return None

    def build_graph(self):
        opts = self._options
        self._peer = self._make_peer_optimizer()
        flat = None if self._peer is None else {"params": self._peer.params, "grads": self._peer.grads}
        self._net = unet.UNet(opts.num_layers, opts.root_size, opts.dilated_layers, opts.batch_size,
                              self.input_size, seed=opts.seed, training=True, flat_buffers=flat)
        B, S, P = opts.batch_size, self.input_size, opts.patch_size
        assert self._net.P == P
        # two input slots: while a step computes from one, the next batch is staged (pinned host
        # copy + host->device transfer on a copy stream) into the other -- see prefetch()
        self._h_slots = [(torch.empty(B, S, S, NUM_CHANNELS, dtype=torch.float32).pin_memory(),
                          torch.empty(B, P, P, dtype=torch.uint8).pin_memory()) for _ in range(2)]
        self._d_slots = [(torch.empty(B, S, S, NUM_CHANNELS, dtype=torch.float32, device="cuda"),
                          torch.empty(B, P, P, dtype=torch.uint8, device="cuda")) for _ in range(2)]
        self._slot = 0
        self._staged = {}  # id(patches object) -> (patches, labels, slot, ready event)
        self._slot_free = [None, None]  # event: the last step that read the slot has consumed it
        self._copy_stream = torch.cuda.Stream()
        self._h_probs = torch.empty(B, P, P, dtype=torch.float32).pin_memory()
        self._h_loss = torch.zeros(1, dtype=torch.float32).pin_memory()
        self._reducer = GradientAllReducer(self._dist.dist, self._dist.world, self._net.grads) \
            if self._dist.active and self._peer is None else None
        if self._reducer is not None:
            self._net.on_bucket_ready = self._reducer.bucket_ready
        if self._peer is not None:
            self._peer._acc = self._net.momentum
            self._net.on_bucket_ready = self._peer.bucket_ready
            self._net.on_backward_begin = lambda scale=1.0: self._peer.arm(
                self._net.learning_rate(opts.lr), opts.momentum, scale)

    @property
    def net(self):
        return self._net

    @property
    def global_step(self):
        return self._net.global_step

    # -- in-graph augmentation (tf_aerial_images.py:173-210) ---------------------------------
    def stochastic_images_augmentation(self, imgs, masks):
        """Per sample: flip_up_down applied by each of three fair coins (the reference hard-codes
        flip_up_down at :188), then rot90 by floor(4U): a uniform element of the dihedral group.
        imgs [B,S,S,3] fp32 / masks [B,P,P] uint8 device tensors."""
        B = imgs.shape[0]
        coins = self._aug_rng.random_sample((3, B)) > 0.5
        flip = coins[0] ^ coins[1] ^ coins[2]
        k = np.floor(self._aug_rng.random_sample(B) * 4).astype(np.int64)
        ops_np = (flip.astype(np.uint8) << 2) | k.astype(np.uint8)
        ops_t = torch.from_numpy(ops_np).cuda()
        return images.d4_transform_dev(imgs, ops_t), images.d4_transform_dev(masks, ops_t)

    # -- one training step on host batches ---------------------------------------------------
    def _stage(self, patches_batch, labels_batch, slot, stream):
        """host batch -> device slot `slot` on `stream` (asynchronous for pinned sources)."""
        dx, dy = self._d_slots[slot]
        if torch.is_tensor(patches_batch) and patches_batch.is_pinned():
            # caller already staged the batch in pinned host memory (fp32 patches, uint8 labels)
            hx, hy = patches_batch, labels_batch
        else:
            hx, hy = self._h_slots[slot]
            hx.copy_(torch.from_numpy(np.ascontiguousarray(patches_batch, dtype=np.float32)))
            hy.copy_(torch.from_numpy(np.ascontiguousarray(labels_batch).astype(np.uint8)))
        with torch.cuda.stream(stream):
            if self._slot_free[slot] is not None:
                stream.wait_event(self._slot_free[slot])
            dx.copy_(hx, non_blocking=True)
            dy.copy_(hy, non_blocking=True)

    def _free_slot(self):
        taken = {v[2] for v in self._staged.values()}
        for slot in (1 - self._slot, self._slot):
            if slot not in taken:
                return slot
        return None

    def prefetch(self, patches_batch, labels_batch):
        """Start moving a FUTURE batch to the GPU (pinned host copy + transfer on a copy stream)
        while the current step computes; the train_batch() call with the same patches object picks
        it up.  Returns False when both input slots are already spoken for.  (The epoch loop of
        train() and bench.py's end-to-end leg use it; train_batch() alone stays correct.)"""
        slot = self._free_slot()
        if slot is None:
            return False
        self._stage(patches_batch, labels_batch, slot, self._copy_stream)
        ev = torch.cuda.Event()
        ev.record(self._copy_stream)
        self._staged[id(patches_batch)] = (patches_batch, labels_batch, slot, ev)
        return True

    def _step_graphs(self, lr):
        """CUDA graphs of one training step, or None when the step runs eagerly.

        A step is a fixed sequence of ~140 launches over preallocated buffers; at small batch
        sizes (the README recipe trains at batch 1) issuing them from Python takes longer than
        the GPU needs to run them, so the sequence is captured once -- forward pass in one graph,
        backward pass + momentum update + weight repack in a second (the fetch of loss and
        probabilities starts between the two) -- and replayed.  The learning rate is a kernel
        argument: the graphs are re-captured when exponential_decay changes it (every 1,000
        steps).  Eager: the first step of a process (one-time kernel attribute set-up), dropout
        < 1 (per-step seeds), data-parallel runs (NCCL buckets on a side stream), profiling."""
        opts = self._options
        mode = getattr(opts, "cuda_graphs", "auto")
        want = mode == "1" or (mode == "auto" and opts.batch_size <= 8)
        from . import ops as _ops
        if (not want or opts.dropout != 1.0 or self._dist.active or _ops._PROFILE is not None
                or getattr(self, "_eager_steps", 0) < 1):
            return None
        cached = getattr(self, "_graphs", None)
        if cached is not None and cached[0] == lr:
            return cached
        net = self._net
        B, S, P = opts.batch_size, self.input_size, opts.patch_size
        if cached is None:
            gx = torch.empty(B, S, S, NUM_CHANNELS, dtype=torch.float32, device="cuda")
            gy = torch.empty(B, P, P, dtype=torch.uint8, device="cuda")
        else:
            gx, gy = cached[1], cached[2]
        self._graphs = None
        g_fwd, g_bwd = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_fwd):
            net.zero_grads()
            net.forward(gx, gy, keep=1.0)
        with torch.cuda.graph(g_bwd, pool=g_fwd.pool()):
            net.backward()
            net.apply_gradients(opts.lr, opts.momentum, 1.0)
        net.global_step -= 1  # capture ran the host side of apply_gradients, not its kernels
        self._graphs = (lr, gx, gy, g_fwd, g_bwd)
        return self._graphs

    def train_batch(self, patches_batch, labels_batch, copy_probs=True):
        """patches_batch [B,S,S,3], labels_batch [B,P,P] (host arrays).  Returns (loss, probs)."""
        opts = self._options
        net = self._net
        cur = torch.cuda.current_stream()
        st = self._staged.pop(id(patches_batch), None)
        if st is not None and st[0] is patches_batch:
            self._slot = st[2]
            cur.wait_event(st[3])
        else:
            slot = self._free_slot()
            if slot is None:  # both slots hold batches staged for later: drop the older one
                self._staged.clear()
                slot = self._slot
            self._slot = slot
            self._stage(patches_batch, labels_batch, slot, cur)
        x, y = self._d_slots[self._slot]
        if opts.image_augmentation:
            x, y = self.stochastic_images_augmentation(x, y)
        lr = net.learning_rate(opts.lr)
        graphs = self._step_graphs(lr)
        if graphs is None:
            net.zero_grads()
            net.forward(x, y, keep=opts.dropout)
        else:  # replay the captured forward pass on the step's inputs
            graphs[1].copy_(x)
            graphs[2].copy_(y)
            graphs[3].replay()
        # loss and probabilities are final once the forward pass is: their device->host copies run
        # on the copy stream underneath the backward pass
        fwd_done = torch.cuda.Event()
        fwd_done.record(cur)
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(fwd_done)
            self._h_probs.copy_(net.probs, non_blocking=True)
            self._h_loss.copy_(net.loss, non_blocking=True)
            fetched = torch.cuda.Event()
            fetched.record(self._copy_stream)
        if graphs is None:
            net.backward()
        else:  # backward pass + momentum update + weight repack, captured together
            graphs[4].replay()
            net.global_step += 1
        # (the inputs are read by the forward pass and by the first-layer weight gradient at the
        # very end of the backward pass: only now may a prefetch overwrite this slot)
        free = torch.cuda.Event()
        free.record(cur)
        self._slot_free[self._slot] = free
        if graphs is None:
            self.apply_update()
            self._eager_steps = getattr(self, "_eager_steps", 0) + 1
        # the step's fetches (loss, probabilities) are what the caller waits for; the update that
        # follows them on the stream is ordered before everything the next call enqueues
        fetched.synchronize()
        # (data parallel: the loss of this rank's slice of the global batch -- the step has ONE
        # collective, the gradient all-reduce; mean_loss() averages the logged values on demand)
        loss = float(self._h_loss[0])
        self.scalars.append((net.global_step, loss, lr))
        # a fresh array per step, like session.run returns (the pinned fetch buffer is overwritten
        # by the next step's copy); copy_probs=False hands out the buffer itself
        probs = self._h_probs.numpy()
        return loss, (probs.copy() if copy_probs else probs)

    def apply_update(self):
        """The optimizer step that follows a backward pass: gradient exchange of the data-parallel
        ranks + momentum update + global_step + operand repack."""
        opts = self._options
        if self._peer is not None:
            self._net.apply_gradients(opts.lr, opts.momentum, peer=self._peer)
        else:
            scale = self._reducer.finish() if self._reducer is not None else 1.0
            self._net.apply_gradients(opts.lr, opts.momentum, scale)

    def mean_loss(self, last=1):
        """Mean over ranks of the mean of the last `last` logged losses (one small all-reduce, on
        demand -- not inside the training step)."""
        v = float(np.mean([s[1] for s in self.scalars[-last:]])) if self.scalars else 0.0
        if self._dist.active:
            t = torch.tensor([v], dtype=torch.float64, device="cuda" if torch.cuda.is_available() else "cpu")
            self._dist.dist.all_reduce(t)
            v = float(t.item()) / self._dist.world
        return v

    def train(self, patches, labels_patches, imgs, labels):
        """Train the model for one epoch (tf_aerial_images.py:212-269)."""
        opts = self._options
        world, rank = self._dist.world, self._dist.rank

        labels_patches = (labels_patches >= 0.5) * 1.
        labels = (labels >= 0.5) * 1.

        num_train_patches = patches.shape[0]

        indices = np.arange(0, num_train_patches)
        np.random.shuffle(indices)

        num_errors = 0
        total = 0
        if self.sharded_data and world > 1:
            # `patches` is this rank's shard (prepare_train_patches): every rank walks its own
            # shuffled shard; the step count follows the reference's rule on the smallest shard so
            # that all ranks run the same number of collective steps
            t = torch.tensor([num_train_patches], device="cuda" if torch.cuda.is_available() else "cpu")
            self._dist.dist.all_reduce(t, op=self._dist.dist.ReduceOp.MIN)
            offsets = list(range(0, int(t.item()) - opts.batch_size, opts.batch_size))
            pick_rank = 0
        else:
            gb = opts.batch_size * world  # global batch: every rank takes its own slice of it
            offsets = list(range(0, num_train_patches - gb, gb))
            pick_rank = rank

        def host_batch(offset):
            idx = rank_batch_indices(indices, offset, pick_rank, opts.batch_size)
            return idx, patches[idx, :, :, :], labels_patches[idx]

        nxt = host_batch(offsets[0]) if offsets else None
        for batch_i, offset in enumerate(offsets):
            batch_indices, pb, lb = nxt
            # stage batch i+1 (host gather + pinned copy + transfer) while step i computes
            nxt = host_batch(offsets[batch_i + 1]) if batch_i + 1 < len(offsets) else None
            if nxt is not None:
                self.prefetch(nxt[1], nxt[2])
            l, predictions = self.train_batch(pb, lb, copy_probs=False)
            step = self._net.global_step
            print("Batch {} Step {}".format(batch_i, step), end="\r")
            self._summary.add({"loss": l, "learning_rate": self.scalars[-1][2]}, global_step=step)

            num_errors += np.abs(labels_patches[batch_indices] - predictions).sum()
            total += opts.batch_size
            self.misclassification = (num_errors, total)
            self._summary.add_to_pixel_missclassification_summary(num_errors, total, step)

            # from time to time do full prediction on some images
            if step > 0 and step % opts.eval_every == 0:
                print()

                images_to_predict = imgs[:opts.num_eval_images, :, :, :]
                masks = self.predict(images_to_predict)
                self.last_eval_masks = masks
                overlays = images.overlays(images_to_predict, masks)
                pred_masks = ((masks > 0.5) * 1).squeeze(-1)
                true_masks = labels[:opts.num_eval_images, :, :]

                self.last_eval_scores = self._summary.add_to_eval_summary(masks, overlays, labels, step)
                self._summary.add_to_overlap_summary(true_masks, pred_masks, step)

            if step > 0 and step % opts.train_score_every == 0:
                self.last_train_masks = self.predict(imgs)
                self.last_train_scores = self._summary.add_to_training_summary(self.last_train_masks, labels, step)

        self._summary.flush()

    # -- sliding-window prediction (tf_aerial_images.py:271-328) -----------------------------
    def predict(self, imgs):
        """Run inference on `imgs` and return predicted masks

        imgs: [num_images, image_height, image_width, num_channel]
        returns: masks [num_images, images_height, image_width, 1] with road probabilities
        """
        opts = self._options
        net = self._net
        world, rank = self._dist.world, self._dist.rank

        num_images = imgs.shape[0]
        print("Running prediction on {} images... ".format(num_images), end="")

        x = torch.from_numpy(np.ascontiguousarray(imgs, dtype=np.float32)).cuda()
        if opts.ensemble_prediction:
            x = images.image_augmentation_ensemble_dev(x)
            num_images = x.shape[0]

        S, P, B = self.input_size, opts.patch_size, opts.batch_size
        offset = int((S - P) / 2)
        x = images.mirror_border_dev(x, offset)
        H = x.shape[1]
        assert x.shape[1] == x.shape[2], "Assume square images"
        assert (H - S) % opts.stride == 0, "Stride sliding should cover the whole image"
        side = (H - S) // opts.stride + 1
        num_patches = num_images * side * side

        plan = shared_window_plan(side, opts.stride, S, opts.num_layers, opts.shared_window_max_input) \
            if getattr(opts, "shared_windows", False) else None
        if plan is not None:
            masks = self._predict_shared(x, num_images, side, plan)
        else:
            masks = self._predict_windows(x, num_images, side)

        if opts.ensemble_prediction:
            masks = images.invert_image_augmentation_ensemble_dev(masks[..., 0].contiguous())[..., None]

        print("Prediction Done")
        return masks.cpu().numpy().astype(np.float64)

    def _predict_windows(self, x, num_images, side):
        """every sliding window is its own forward pass (the reference's loop, :296-315); ranks
        take contiguous slices of the patch list"""
        opts, net = self._options, self._net
        world, rank = self._dist.world, self._dist.rank
        S, P, B = self.input_size, opts.patch_size, opts.batch_size
        num_patches = num_images * side * side
        k0, k1 = shard_range(num_patches, rank, world)
        preds = torch.empty(max(k1 - k0, 1), P, P, 1, dtype=torch.float32, device="cuda")
        if getattr(self, "_d_predict", None) is None:
            self._d_predict = torch.empty(B, S, S, NUM_CHANNELS, dtype=torch.float32, device="cuda")
        batch = self._d_predict
        for k in range(k0, k1, B):
            cnt = min(B, k1 - k)
            # a short tail batch keeps stale patches in the remaining slots (the reference's
            # intent at :298-301 is zero patches whose outputs are dropped at :315)
            images.extract_patches_dev(x, S, opts.stride, k, cnt, out=batch)
            net.forward(batch, keep=1.0)
            preds[k - k0:k - k0 + cnt, :, :, 0].copy_(net.probs[:cnt])

        # construct masks: overlap average in gather form
        if world == 1:
            return images.images_from_patches_dev(preds, num_images, side, opts.stride)
        part = images.images_from_patches_dev(preds, num_images, side, opts.stride, k0, k1 - k0,
                                              normalize=False)
        self._dist.dist.all_reduce(part)  # the single exchange of a sharded prediction
        return images.divide_by_hits_dev(part, side, P, opts.stride)

    def _shared_net(self, big_input, big_batch):
        """Forward-only engine for enlarged windows.  It aliases the training engine's master and
        packed weights (no second copy, no random init, no repack) and is cached by window size
        alone: the first request fixes its batch, later requests run on it whatever their batch
        (predict() pads a short chunk with stale windows whose outputs are dropped)."""
        opts = self._options
        nets = self.__dict__.setdefault("_shared_nets", {})
        if big_input not in nets:
            while len(nets) >= 4:  # activations are the big allocations: keep at most four sizes
                nets.pop(next(iter(nets)))
            nets[big_input] = unet.UNet(opts.num_layers, opts.root_size, opts.dilated_layers, big_batch,
                                        big_input, seed=opts.seed, training=False, weights_from=self._net)
        return nets[big_input]

    def _predict_shared(self, x, num_images, side, plan):
        """Aligned windows evaluated once (shared_window_plan): each forward pass covers up to
        n x n windows of the reference's loop; their outputs are cut out of the enlarged output
        and stored at their positions in the patch list, then overlap-averaged as usual.  A job
        (image, window along x, window along y) runs at the size of its longer window, so the
        classes with fewer members use a smaller engine.  Ranks take contiguous slices of the
        job list."""
        opts = self._options
        world, rank = self._dist.world, self._dist.rank
        S, P, B, stride = self.input_size, opts.patch_size, opts.batch_size, opts.stride
        n, q, wins = plan
        by_size = shared_window_jobs(num_images, wins, rank, world)
        num_patches = num_images * side * side
        preds = torch.empty(num_patches, P, P, 1, dtype=torch.float32, device="cuda")
        if world > 1:  # foreign patches stay zero in this rank's partial sums
            from . import ops as _ops
            _ops.fill_zero(preds)
        for m, group in sorted(by_size.items()):
            win_in, win_out = S + q * (m - 1), P + q * (m - 1)
            if m == 1:
                net, win_b = self._net, B
            else:
                # as many windows per pass as the training batch has pixels, evened out over the
                # passes so that the last one is not mostly padding
                win_b = max(1, int(B * S * S / (win_in * win_in)))
                win_b = -(-len(group) // -(-len(group) // win_b))
                net = self._shared_net(win_in, win_b)
                win_b = net.B
            assert net.P == win_out
            # device job tables of the whole group, uploaded once: (image, top row, left column,
            # destination) of every enlarged input window -- pixels beyond the padded image read
            # as zero and only reach outputs that are never used -- and of every patch output that
            # is cut out of the enlarged probability maps into its slot of the patch list
            t_in, t_out, out_ofs = [], [], [0]
            for jj, (img, wx, wy) in enumerate(group):
                slot = jj % win_b
                t_in.append((img, stride * wy[0], stride * wx[0], slot))
                for mx, kx in enumerate(wx):
                    for my, ky in enumerate(wy):
                        t_out.append((slot, my * q, mx * q, (img * side + kx) * side + ky))
                out_ofs.append(len(t_out))
            jobs_in = torch.tensor(t_in, dtype=torch.int32).cuda()
            jobs_out = torch.tensor(t_out, dtype=torch.int32).cuda()
            batch = self.__dict__.setdefault("_window_batches", {}).get((win_b, win_in))
            if batch is None:
                batch = torch.empty(win_b, win_in, win_in, NUM_CHANNELS, dtype=torch.float32, device="cuda")
                self._window_batches[(win_b, win_in)] = batch
            for j in range(0, len(group), win_b):
                cnt = min(win_b, len(group) - j)
                images.copy_windows_dev(x, win_in, jobs_in[j:j + cnt], batch)
                net.forward(batch, keep=1.0)
                images.copy_windows_dev(net.probs.view(net.B, win_out, win_out, 1), P,
                                        jobs_out[out_ofs[j]:out_ofs[j + cnt]], preds)
        if world == 1:
            return images.images_from_patches_dev(preds, num_images, side, stride)
        part = images.images_from_patches_dev(preds, num_images, side, stride, normalize=False)
        self._dist.dist.all_reduce(part)  # the single exchange of a sharded prediction
        return images.divide_by_hits_dev(part, side, P, stride)

    def predict_batchwise(self, imgs, pred_batch_size):
        masks = []
        for i in range(int(np.ceil(imgs.shape[0] / pred_batch_size))):
            start = i * pred_batch_size
            end = start + pred_batch_size
            masks.append(self.predict(imgs[start:end]))

        if len(masks) > 1:
            masks = np.concatenate(masks, axis=0)
            return masks
        else:
            return masks[0]

    # -- checkpoints (tf_aerial_images.py:343-379) -------------------------------------------
    def save(self, epoch=0):
        """Writes <save_path>/<experiment_name>/model-epoch-NNN.chkpt (.npz payload + the .meta
        marker the reference's restore() globs for); all variables incl. momentum slots,
        global_step and the dead conv_dilut_{L-1} weights, keyed by TensorFlow variable names."""
        opts = self._options
        return self.save_to(os.path.abspath(
            os.path.join(opts.save_path, self.experiment_name, 'model-epoch-{:03d}.chkpt'.format(epoch))))

    def save_to(self, model_data_dir):
        """The checkpoint write behind save(); also used for the '<save_dir>-model.chkpt' copy
        of a submission run (tf_aerial_images.py:461).  Written as a TensorFlow V2 checkpoint
        bundle (<path>.index + <path>.data-00000-of-00001, tf_checkpoint.write_bundle) -- the
        files tf.train.Saver writes for the reference and can restore -- plus the <path>.meta
        marker the reference's restore() globs for."""
        from . import tf_checkpoint
        opts = self._options
        momentum = self._net.momentum
        if self._peer is not None:  # every rank holds the momentum of its own slices only
            momentum = self._peer.full_momentum(self._net)
        if self._dist.rank == 0:
            os.makedirs(os.path.dirname(model_data_dir), exist_ok=True)
            payload = {}
            for k, v in self._net.state_dict("params").items():
                payload[k] = v
            keep, self._net.momentum = self._net.momentum, momentum
            for k, v in self._net.state_dict("momentum").items():
                payload[k + "/Momentum"] = v
            self._net.momentum = keep
            payload["global_step"] = np.array(self._net.global_step, dtype=np.int32)
            tf_checkpoint.write_bundle(model_data_dir, payload)
            with open(model_data_dir + ".meta", "w") as f:
                f.write("rsu_b200 checkpoint; num_layers={} root_size={} dilated_layers={}\n".format(
                    opts.num_layers, opts.root_size, opts.dilated_layers))
            print("Model saved in file: {}".format(model_data_dir))
        self._dist.barrier()
        return model_data_dir

    def restore(self, date=None, epoch=None, file=None):
        """Restores model from saved checkpoint

        date: which model should be restored (most recent if None)
        epoch: at which epoch model should be restored (most recent if None)
        file: provide directly the checkpoint file te restore
        """
        opts = self._options

        if file is not None:
            model_data_dir = file
        else:
            # get experiment name to restore from
            if date is None:
                dates = [date for date in glob.glob(os.path.join(opts.save_path, "*")) if os.path.isdir(date)]
                model_data_dir = sorted(dates)[-1]
            else:
                model_data_dir = os.path.abspath(os.path.join(opts.save_path, date))

            # get epoch construct final path
            if epoch is None:
                model_data_dir = os.path.abspath(os.path.join(model_data_dir, 'model-epoch-*.chkpt.meta'))
                model_data_dir = sorted(glob.glob(model_data_dir))[-1][:-5]
            else:
                model_data_dir = os.path.abspath(
                    os.path.join(model_data_dir, 'model-epoch-{:03d}.chkpt'.format(epoch)))

        from . import tf_checkpoint
        if tf_checkpoint.is_bundle(model_data_dir):
            # a TensorFlow V2 checkpoint: this engine's own save() or the reference's tf.train.Saver
            # (same variable names and layouts, momentum slots under <name>/Momentum)
            tensors = tf_checkpoint.read_bundle(model_data_dir)
        elif os.path.isfile(model_data_dir):
            with np.load(model_data_dir) as z:  # round-1 .npz payload
                tensors = {k: z[k] for k in z.files}
        else:
            raise FileNotFoundError(
                "no checkpoint at {0}: expected {0}.index + {0}.data-00000-of-00001 (TensorFlow V2 "
                "bundle, written by save() and by the reference's Saver)".format(model_data_dir))
        known = ("global_step", "beta1_power", "beta2_power")
        params = {k: v for k, v in tensors.items() if not k.endswith("/Momentum") and k not in known}
        mom = {k[:-len("/Momentum")]: v for k, v in tensors.items() if k.endswith("/Momentum")}
        step = int(tensors["global_step"]) if "global_step" in tensors else 0
        # strict: every variable of the architecture must be in the file with its exact shape and
        # nothing unknown may be (a mismatched checkpoint must not leave random-init layers behind)
        self._net.load_state(params, mom if mom else None, strict=True)
        self._net.global_step = step
        self._net.pack_weights()
        print("Model restored from from file: {}".format(model_data_dir))


def main(argv=None):
    flags = make_parser().parse_args(argv)
    opts = Options(flags)
    device = '/device:CPU:0' if opts.gpu == -1 else '/device:GPU:{}'.format(opts.gpu)
    print("Running on device {}".format(device if opts.gpu >= 0 else "cuda:0 (no CPU path exists)"))
    model = ConvolutionalModel(opts, None)

    if opts.restore_model:
        if opts.model_path is not None:
            model.restore(file=opts.model_path)
            print("Restore model: {}".format(opts.model_path))
        else:
            print("Restore date: {}".format(opts.restore_date))
            model.restore(date=opts.restore_date, epoch=opts.restore_epoch)

    if opts.num_epoch > 0:
        train_images, train_groundtruth = images.load_train_data(opts.train_data_dir)

        # every rank prepares only the (angle, image) pairs it will train on; one process: the
        # reference's block (expand_and_rotate + extract_patches of images and ground truth)
        dist_ = model._dist
        patches, labels_patches = prepare_train_patches(train_images, train_groundtruth, opts,
                                                        dist_.rank, dist_.world)
        model.sharded_data = dist_.world > 1

        print("Train on {} patches of size {}x{}".format(patches.shape[0], patches.shape[1], patches.shape[2]))

        print("Train on {} groundtruth patches of size {}x{}".format(
            labels_patches.shape[0], labels_patches.shape[1], labels_patches.shape[2]))

        model._summary.add_to_eval_patch_summary(train_groundtruth)
        for i in range(opts.num_epoch):
            print("==== Train epoch: {} ====".format(i))
            model._summary.reset()  # Reset scores (tf.local_variables_initializer, :428)
            model.train(patches, labels_patches, train_images, train_groundtruth)  # Process one epoch
            model.save(i)  # Save model to disk

    if opts.eval_train:
        # dump predictions on the training set (tf_aerial_images.py:432-446)
        print("Evaluate Test")
        eval_images, eval_groundtruth = images.load_train_data(opts.train_data_dir)
        pred_masks = model.predict_batchwise(eval_images, opts.pred_batch_size)
        pred_labels = ((pred_masks > 0.5) * 1).squeeze(-1)
        dumps = (
            (pred_labels, "eval_binary_pred_{:03d}.png", True),
            (pred_masks, "eval_probability_pred_{:03d}.png", True),
            (images.overlays(eval_images, pred_masks, fade=0.5), "eval_overlays_pred_{:03d}.png", False),
            (images.overlap_pred_true(pred_labels, eval_groundtruth), "eval_confusion_{:03d}.png", False),
            (images.overlapp_error(pred_labels, eval_groundtruth), "eval_orror_{:03d}.png", True),
        )
        for arrays, name, grey in dumps:
            images.save_all(arrays, opts.eval_data_dir, name, greyscale=grey)

    if opts.eval_data_dir and not opts.eval_train:
        # submission run (tf_aerial_images.py:448-461): masks -> 16 x 16 vote -> overlays + csv
        print("Running inference on eval data {}".format(opts.eval_data_dir))
        eval_images = images.load(opts.eval_data_dir)
        start = time.time()
        masks = model.predict_batchwise(eval_images, opts.pred_batch_size)
        stop = time.time()
        print("Prediction time:{} mins".format((stop - start) / 60))
        masks = images.quantize_mask(masks, patch_size=IMG_PATCH_SIZE, threshold=FOREGROUND_THRESHOLD)
        save_dir = os.path.abspath(os.path.join(opts.save_path, model.experiment_name))
        images.save_all(images.overlays(eval_images, masks, fade=0.4), save_dir)
        images.save_submission_csv(masks, save_dir, IMG_PATCH_SIZE)
        model.save_to(save_dir + "-model.chkpt")  # the model used for the submission
    return model


load_images = images.load                    # kept under their earlier names
load_train_data = images.load_train_data


if __name__ == '__main__':
    main()
