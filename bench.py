#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 U-Net hot path.

    python bench.py --gpus N --steps K --warmup W            (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): training of the 6-layer dilated U-Net (root 64), batch 32
of 388^2 output patches (764^2 inputs) per GPU, bf16 tensor-core math with fp32 accumulate and
fp32 master weights, momentum SGD.  One step = forward + backward + (all-reduce) + update.
`value` is whole-job patches/s with the batch resident in HBM, `e2e` the same metric through the
public API (ConvolutionalModel.train_batch) with pinned host buffers copied in every step and
the loss + probabilities read back.  Data are synthetic (seed 2017), weights random glorot init.
Inputs per step (19 GB of activations) are far larger than the 126 MB L2, so no explicit flush.

`--impl reference` times the CPU oracle restatement of the reference's TensorFlow path
(oracle/unet_oracle.py; TF 1.4 itself cannot be installed here) on the host cores, one patch per
step of the same model.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# NCCL prints its version banner to stdout at NCCL_DEBUG=VERSION; stdout carries one JSON line
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", ""):
    os.environ["NCCL_DEBUG"] = "WARN"

METRIC = "U-Net train patches/s (388^2)"
UNIT = "patches/s"
CFG = dict(num_layers=6, root_size=64, dilated_layers=True, patch_size=388, batch_size=32)


def workload_config(B, world):
    """`config` of the JSON line -- identical for the GPU arm and the reference arm."""
    return {"workload": "U-Net num_layers=6 root_size=64 --dilated_layers, batch %d/GPU of "
                        "764^2->388^2 patches, momentum SGD lr 0.01 mu 0.9, dropout 1.0 "
                        "(BASELINE.json configs[1])" % B,
            "global_batch": B * world, "parallelism": "dp%d" % world,
            "l2": "inputs larger than L2 (>= 19 GB of activations per step)",
            "timing": "CUDA events on the launch stream, max over ranks (reference arm: host clock)"}


def host_threads():
    """Cores this process may use (torchrun exports OMP_NUM_THREADS=1 for its workers: the CPU
    legs set the thread count themselves)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), \
            d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.proc = None
        self.lines = []
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                smax.append(float(p[2]))
                power.append(float(p[3]))
            except ValueError:
                continue
            for name, val in zip(names, p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load" = samples drawing more than half of the maximum observed power
        pmax = max(power)
        load = [s for s, w in zip(sm, power) if w >= 0.5 * pmax] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(smax)),
                "power_w_max": pmax, "samples": len(sm), "reasons": sorted(reasons)}


def synthetic_batch(B, S, P, seed):
    rs = np.random.RandomState(seed)
    x = rs.rand(B, S, S, 3).astype(np.float32)
    lab = (rs.rand(B, P, P) < 0.25)
    # smooth-ish road-like field: 9x9 box filter then threshold
    k = 9
    f = lab.astype(np.float32)
    c = np.cumsum(np.cumsum(np.pad(f, ((0, 0), (k // 2 + 1, k // 2), (k // 2 + 1, k // 2))), 1), 2)
    box = c[:, k:, k:] - c[:, :-k, k:] - c[:, k:, :-k] + c[:, :-k, :-k]
    return x, (box / (k * k) >= 0.25).astype(np.uint8)


# ------------------------------------------------------------------ CPU oracle timing
def time_oracle(steps, warmup, threads=None):
    """The oracle (port of the reference's TF graph) on the host: fwd + bwd + momentum update of
    the SAME model, one 764^2 -> 388^2 patch per step, fp32."""
    import torch
    from oracle import unet_oracle as O
    torch.set_num_threads(threads or host_threads())
    L, root, dil, P = CFG["num_layers"], CFG["root_size"], CFG["dilated_layers"], CFG["patch_size"]
    S = O.input_size_needed(P, L)
    params = O.init_params(L, root, dil, seed=2017)
    tp = O.to_torch(params, requires_grad=True)
    accs = {k: torch.zeros_like(v) for k, v in tp.items()}
    x, lab = synthetic_batch(1, S, P, 2017)
    xt, lt = torch.tensor(x), torch.tensor(lab.astype(np.int64))
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        for v in tp.values():
            v.grad = None
        logits = O.forward(xt, tp, L, root, dil)
        loss, _ = O.loss_and_probs(logits, lt)
        loss.backward()
        with torch.no_grad():
            for k, v in tp.items():
                if v.grad is None:
                    continue
                accs[k].mul_(0.9).add_(v.grad)
                v.sub_(0.01 * accs[k])
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return 1.0 / float(np.mean(times)), float(np.mean(times)), torch.get_num_threads()


def time_oracle_predict(n_forwards=4):
    """CPU leg of the prediction half of the metric (SURVEY 8(d) "reference CPU path timed beside
    it"): the oracle pipeline of ConvolutionalModel.predict on one synthetic 604^2 image, stride
    12, 6-way ensemble.  The NumPy helpers run in full at config-3 sizes; of the 2,166 forward
    passes of the L=6 dilated model a bounded sample is timed and extrapolated."""
    import torch
    from oracle import images_oracle as IO
    from oracle import unet_oracle as O
    torch.set_num_threads(host_threads())
    L, root, dil, P = CFG["num_layers"], CFG["root_size"], CFG["dilated_layers"], CFG["patch_size"]
    S = O.input_size_needed(P, L)
    off = (S - P) // 2
    img = np.random.RandomState(2017).rand(1, 604, 604, 3).astype(np.float32)
    helpers = {}

    def clock(name, nbytes, fn):
        t0 = time.perf_counter()
        out = fn()
        dt = time.perf_counter() - t0
        helpers[name] = {"seconds": dt, "gbs": nbytes / dt / 1e9, "alg_bytes": int(nbytes)}
        return out

    ens = clock("image_augmentation_ensemble", 7 * img.size * 8, lambda: IO.image_augmentation_ensemble(img))
    pad = clock("mirror_border", (ens.size + 6 * 980 * 980 * 3) * 8, lambda: IO.mirror_border(ens, off))
    # one variant's patch tensor (361 x 764^2 x 3 float64 = 5 GB); the six variants cost six times that
    pat = clock("extract_patches", (980 * 980 * 3 + 361 * S * S * 3) * 8,
                lambda: IO.extract_patches(pad[:1], S, stride=12, predict_patch_size=P))
    tp = O.to_torch(O.init_params(L, root, dil, seed=2017))
    x = torch.tensor(pat[:n_forwards + 1], dtype=torch.float32)
    del pat
    with torch.no_grad():
        O.forward(x[:1], tp, L, root, dil)  # warm-up
        t0 = time.perf_counter()
        for k in range(n_forwards):
            O.forward(x[k + 1:k + 2], tp, L, root, dil)
        fwd_s = (time.perf_counter() - t0) / n_forwards
    preds = np.random.RandomState(1).rand(6, 361, P, P, 1)
    masks = clock("images_from_patches", (preds.size + 6 * 604 * 604) * 8,
                  lambda: IO.images_from_patches(preds, stride=12))
    clock("invert_image_augmentation_ensemble", 7 * 604 * 604 * 8,
          lambda: IO.invert_image_augmentation_ensemble(masks))
    n_fwd = 6 * 19 * 19
    total = (helpers["image_augmentation_ensemble"]["seconds"] + helpers["mirror_border"]["seconds"]
             + 6 * helpers["extract_patches"]["seconds"] + n_fwd * fwd_s
             + helpers["images_from_patches"]["seconds"] + helpers["invert_image_augmentation_ensemble"]["seconds"])
    return {"value": 604 * 604 / 1e6 / total, "unit": "Mpix/s", "seconds_per_image": total,
            "patch_forwards_per_s": 1.0 / fwd_s, "cores": torch.get_num_threads(), "kind": "port",
            "sample": "helpers in full (extract_patches: one of the 6 variants, x 6); %d of the %d forward passes "
                      "of the L=6 dilated 764^2->388^2 model timed (fp32 torch-CPU oracle, %.2f s each), "
                      "extrapolated" % (n_forwards, n_fwd, fwd_s),
            "helpers": helpers}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    warmup = max(0, args.warmup)
    v, sec, threads = time_oracle(steps, warmup)
    B = args.batch or CFG["batch_size"]
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(B, max(1, args.gpus)),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d steps of batch 1 of the same model (patches/s does not depend on the "
                                   "batch size on the CPU), fwd+bwd+momentum update, fp32 torch-CPU oracle on "
                                   "%d host threads; rank 0 only" % (steps, threads)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ BASELINE.json configs[4]
def activation_bytes(L, root, dilated, P, training=True):
    """bf16 bytes of the engine's activation (+ gradient) buffers per patch: what bounds the
    per-GPU micro-batch of the large-tile configurations."""
    from road_segmentation_unet_b200 import unet
    S = unet.input_size_needed(P, L)
    f = [root * 2 ** i for i in range(L)]
    in_size, s = [], S
    for i in range(L):
        in_size.append(s)
        s = (s - 4) // 2
    up, net = [], in_size[L - 1] - 4
    for j in range(L - 1):
        up.append(2 * net)
        net = 2 * net - 4
    fwd = bwd = 0
    for i in range(L):
        s = in_size[i]
        fwd += ((s - 2) ** 2 + (s - 4) ** 2 + (((s - 4) // 2) ** 2 if i < L - 1 else 0)) * f[i]
        bwd += ((s - 2) ** 2 + (s - 4) ** 2) * f[i] + (s * s * f[i - 1] if i > 0 else 0)
        if dilated and i < L - 1:
            t = up[L - 2 - i]
            fwd += ((t + 4) ** 2 + t * t) * f[i]
            bwd += (t + 4) ** 2 * f[i]
    for j in range(L - 1):
        fo, t = f[L - 2 - j], up[j]
        fwd += (t * t + (t - 2) ** 2 + (t - 4) ** 2) * fo
        bwd += (t * t * (3 if dilated else 2) + (t - 2) ** 2 + (t - 4) ** 2) * fo
    return 2 * (fwd + (bwd if training else 0)) + S * S * 3 * 4


def run_large_tiles(args, torch, dist, tfa, unet, world, rank, barrier):
    """BASELINE.json configs[4]: root_size 64 / 128 U-Net (L=6, dilated) on 1024^2 / 2048^2 tiles,
    global batch 64 over 8 GPUs = 8 patches per GPU, bf16 train + predict.  1024 and 2048 are not
    valid patch sizes (unet.input_size_needed: P must be a multiple of 32 plus 4), so
      train:   tiles as patches, P = 1028 (1404^2 inputs) and P = 2052 (2428^2 inputs); 8 patches per
               GPU and optimizer step, as micro-batches that fit in HBM (gradient accumulation);
      predict: tiles as images, one synthetic 1024^2 image at stride 12 and one 2048^2 image at
               stride 20 per call, 6-way ensemble, 388^2 windows, work sharded over the ranks."""
    out = {"train": [], "predict": []}
    peak, peak_sus, _, _ = measured_peaks()
    free_b = torch.cuda.mem_get_info()[0]
    for root, P in ((64, 1028), (128, 1028), (64, 2052), (128, 2052)):
        per_patch = activation_bytes(6, root, True, P)
        n_params = sum(int(np.prod(sh)) for sh in unet.variable_shapes(6, root, True).values())
        budget = int(0.8 * (free_b - n_params * 4 * 5 - (6 << 30)))
        micro = max(1, min(8, int(budget // per_patch)))
        while 8 % micro:
            micro -= 1
        opts = tfa.Options()
        opts.batch_size, opts.num_layers, opts.root_size, opts.dilated_layers = micro, 6, root, True
        opts.patch_size, opts.dropout, opts.lr, opts.momentum = P, 1.0, 0.01, 0.9
        model = tfa.ConvolutionalModel(opts, None)
        net, S = model.net, model.input_size
        g = torch.Generator(device="cuda").manual_seed(2017 + rank)
        mb = [(torch.rand(micro, S, S, 3, device="cuda", generator=g),
               (torch.rand(micro, P, P, device="cuda", generator=g) < 0.3).to(torch.uint8))
              for _ in range(8 // micro)]
        finish = (lambda: model._reducer.finish()) if model._reducer is not None else None

        def step():
            net.accumulate_step(mb, opts.lr, opts.momentum, peer=model._peer, finish=finish)

        step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K = 3
        for _ in range(K):
            step()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / K
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        f_alg = 3 * sum(unet.plan_flops(6, root, True, P).values())
        pps = world * 8 / (ms / 1e3)
        out["train"].append({"root_size": root, "patch_size": P, "input_size": S, "global_batch": 8 * world,
                             "micro_batch": micro, "ms_per_step": ms, "patches_per_s": pps,
                             "tflops_per_gpu": pps / world * f_alg / 1e12,
                             "frac_of_sustained_peak": pps / world * f_alg / 1e12 / peak_sus,
                             "loss": float(net.loss.item())})
        del model, net, mb
        torch.cuda.empty_cache()
    # predict: the flagship model (388^2 windows) over large images
    for root in (64, 128):
        opts = tfa.Options()
        opts.batch_size, opts.num_layers, opts.root_size, opts.dilated_layers = 32 if root == 64 else 16, 6, root, True
        opts.patch_size, opts.dropout, opts.ensemble_prediction = 388, 1.0, True
        model = tfa.ConvolutionalModel(opts, None)
        for size, stride in ((1024, 12), (2048, 20)):
            opts.stride = stride
            img = np.random.RandomState(2017).rand(1, size, size, 3).astype(np.float32)
            with contextlib.redirect_stdout(io.StringIO()):
                model.predict(img)  # warm-up: allocates the enlarged-window engines of this plan
                barrier()
                t0 = time.perf_counter()
                mask = model.predict(img)
                torch.cuda.synchronize()
                sec = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([sec], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                sec = float(t.item())
            side = (size + 376 - 764) // stride + 1
            out["predict"].append({"root_size": root, "image": size, "stride": stride, "windows": 6 * side * side,
                                   "seconds": sec, "mpix_per_s": size * size / 1e6 / sec,
                                   "window_equivalents_per_s": 6 * side * side / sec,
                                   "mask_mean": float(mask.mean())})
        del model
        torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from road_segmentation_unet_b200 import ops, unet
    from road_segmentation_unet_b200 import tf_aerial_images as tfa

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group(backend="nccl")

    B = args.batch or CFG["batch_size"]
    opts = tfa.Options()
    opts.batch_size, opts.num_layers, opts.root_size = B, CFG["num_layers"], CFG["root_size"]
    opts.dilated_layers, opts.patch_size = CFG["dilated_layers"], CFG["patch_size"]
    opts.dropout, opts.lr, opts.momentum, opts.image_augmentation = 1.0, 0.01, 0.9, False
    opts.num_gpu = world
    model = tfa.ConvolutionalModel(opts, None)
    net = model.net
    S, P = model.input_size, opts.patch_size
    xh, lh = synthetic_batch(B, S, P, 2017 + rank)
    x = torch.tensor(xh).cuda()
    lab = torch.tensor(lh).cuda()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t_start = time.perf_counter()

    def note(msg):  # progress on stderr (stdout carries the one JSON line): where a stuck run stopped
        sys.stderr.write("[bench rank %d +%.1fs] %s\n" % (rank, time.perf_counter() - t_start, msg))
        sys.stderr.flush()

    def device_step():
        net.zero_grads()
        net.forward(x, lab, keep=1.0)
        net.backward()
        model.apply_update()

    W, K = max(args.warmup, 3), max(args.steps, 1)
    note("model built (%s), %d warm-up steps" % (
        "single GPU" if model._peer is None and world == 1 else
        "NCCL all-reduce" if model._peer is None else
        "peer optimizer, overlap=%s, multicast mode %d" % (model._peer.overlap, model._peer.mc_mode), W))
    for _ in range(W):
        device_step()
    barrier()
    note("warm-up done, timing %d steps" % K)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    lib = ops._lib.load()
    lib.rsu_reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K):
        device_step()
    e1.record()
    barrier()
    launches = int(lib.rsu_launch_count())
    launch_hist = {k: v / K for k, v in ops._lib.launch_histogram().items()}  # per step
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    loss_val = float(net.loss.item())
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / K
    value = world * B * K / (ms / 1e3)
    note("device-timed region done: %.2f ms per step" % ms_per_step)

    if args.large_tiles:  # BASELINE.json configs[4] only (its own line; not the headline config)
        del x, lab
        model._net = None
        del model, net
        torch.cuda.empty_cache()
        res = run_large_tiles(args, torch, dist, tfa, unet, world, rank, barrier)
        if rank == 0:
            print(json.dumps({"metric": "large-tile sweep (BASELINE.json configs[4])", "n_gpus": world,
                              "headline_value": value, "large_tiles": res}))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    if args.quick:  # profiling runs (ncu): only the device-timed region
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
                              "warmup": W, "ms_per_step": ms_per_step, "gpu_launches": launches,
                              "quick": True}))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- end to end through the public API: pinned host batch in, loss + probabilities out
    # (like the epoch loop of ConvolutionalModel.train: the transfer of batch i+1 is started before
    # step i is run, so it overlaps that step's kernels; every step's copy is inside the timed region)
    xp = [torch.from_numpy(xh).pin_memory(), torch.from_numpy(xh.copy()).pin_memory()]
    lp = [torch.from_numpy(lh).pin_memory(), torch.from_numpy(lh.copy()).pin_memory()]
    for _ in range(2):
        model.train_batch(xp[0], lp[0])
    barrier()
    t0 = time.perf_counter()
    ke = max(2, min(K, 10))
    model.prefetch(xp[0], lp[0])
    for i in range(ke):
        if i + 1 < ke:
            model.prefetch(xp[(i + 1) % 2], lp[(i + 1) % 2])
        model.train_batch(xp[i % 2], lp[i % 2])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * B * ke / e2e_s
    note("end-to-end leg done")
    h2d = B * S * S * 3 * 4 + B * P * P
    d2h = B * P * P * 4 + 4
    # the same loop with --image_augmentation (BASELINE.json configs[3]): every step additionally
    # applies a random dihedral transform to the batch and its masks on the device (rsu_d4_transform)
    opts.image_augmentation = True
    model.train_batch(xp[0], lp[0])
    barrier()
    t0 = time.perf_counter()
    model.prefetch(xp[0], lp[0])
    for i in range(ke):
        if i + 1 < ke:
            model.prefetch(xp[(i + 1) % 2], lp[(i + 1) % 2])
        model.train_batch(xp[i % 2], lp[i % 2])
    torch.cuda.synchronize()
    aug_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([aug_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        aug_s = float(t.item())
    e2e_aug_value = world * B * ke / aug_s
    note("end-to-end leg with image augmentation done")
    opts.image_augmentation = False

    # ---- sliding-window ensemble prediction (BASELINE.json configs[2]) through the public API:
    # one synthetic 604^2 image, stride 12, 6-way flip/rot90 ensemble = 2,166 patch forwards of the
    # same model, patches sharded over the ranks, host image in / host mask out
    predict = None
    if not args.no_predict:
        import contextlib
        import io
        opts.ensemble_prediction, opts.stride = True, 12
        imgs = np.random.RandomState(2017).rand(args.predict_images, 604, 604, 3).astype(np.float32)

        def timed_predict(shared, warm):
            opts.shared_windows = shared
            with contextlib.redirect_stdout(io.StringIO()):
                model.predict(warm)
                barrier()
                t0 = time.perf_counter()
                out = model.predict(imgs)
                torch.cuda.synchronize()
                sec = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([sec], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                sec = float(t.item())
            return out, sec

        def flops(patch):
            return sum(unet.plan_flops(CFG["num_layers"], CFG["root_size"], CFG["dilated_layers"], patch).values())

        mpix = args.predict_images * 604 * 604 / 1e6
        # (a) the reference's loop: one forward pass per sliding window (2,166 per image)
        masks_loop, loop_s = timed_predict(False, imgs[:, :400, :400])  # warm-up: 4 patches per variant
        n_fwd = args.predict_images * 6 * 19 * 19
        # (b) the default path: windows whose origins differ by a multiple of the pooling period
        # are evaluated once inside an enlarged window (tf_aerial_images.shared_window_plan)
        masks, pred_s = timed_predict(True, imgs)  # warm-up allocates the enlarged engine
        n_big, q_big, wins = tfa.shared_window_plan(19, 12, S, CFG["num_layers"], opts.shared_window_max_input)
        sizes = {}
        for wx in wins:
            for wy in wins:
                m = max(len(wx), len(wy))
                sizes[m] = sizes.get(m, 0) + args.predict_images * 6
        n_win = sum(sizes.values())
        done_flops = sum(c * flops(P + q_big * (m - 1)) for m, c in sizes.items())
        predict = {"value": mpix / pred_s, "unit": "Mpix/s", "seconds": pred_s,
                   "mode": "shared windows: %s forward passes, each covering up to %dx%d of the %d sliding "
                           "windows" % (" + ".join("%d of %d^2 -> %d^2" % (c, S + q_big * (m - 1), P + q_big * (m - 1))
                                                   for m, c in sorted(sizes.items())), n_big, n_big, n_fwd),
                   "executed_tflops": done_flops / pred_s / 1e12,
                   "window_equivalents_per_s": n_fwd / pred_s,
                   "window_loop": {"value": mpix / loop_s, "unit": "Mpix/s", "seconds": loop_s,
                                   "patch_forwards_per_s": n_fwd / loop_s,
                                   "tflops": n_fwd * flops(P) / loop_s / 1e12,
                                   "mode": "one forward pass per sliding window (the reference's loop)"},
                   "agreement": {"max_abs_diff": float(np.abs(masks - masks_loop).max()),
                                 "pixels_equal_at_0.5": float(np.mean((masks > 0.5) == (masks_loop > 0.5)))},
                   "config": "%d synthetic 604^2 image(s), stride 12, 6-way ensemble, work sharded over %d "
                             "GPU(s), ConvolutionalModel.predict (host in / host out)"
                             % (args.predict_images, world),
                   "mask_mean": float(masks.mean())}

    # ---- roofline of the dominant kernel: CUDA events around every tcgen05 launch (same steps,
    # instrumented pass so that the headline region above stays free of event records)
    roof = None
    extra = {}
    # (every rank runs the instrumented steps -- they contain the gradient all-reduce -- but only
    # rank 0 brackets its kernels with events)
    note("prediction leg done" if predict is not None else "prediction leg skipped")
    kp = min(K, 3)
    if rank == 0:
        ops.profile_start()
    for _ in range(kp):
        device_step()
    barrier()
    note("instrumented pass done")
    if rank == 0:
        peak, peak_sus, hbm, src = measured_peaks()
        rec = ops.profile_stop()
        if args.dump_layers:
            per = {}
            for kind, layer, fl, t_ms in rec:
                a = per.setdefault("%s|%s" % (kind, layer), [0.0, 0.0, 0])
                a[0] += fl / kp
                a[1] += t_ms / kp
                a[2] += 1
            with open(args.dump_layers, "w") as f:
                json.dump({k: {"alg_gflop": v[0] / 1e9, "ms": v[1], "launches": v[2] / kp,
                               "tflops": v[0] / max(v[1], 1e-9) / 1e9} for k, v in per.items()}, f, indent=1)
        by_kind = {}
        for kind, layer, fl, t_ms in rec:
            a = by_kind.setdefault(kind, [0.0, 0.0, 0])
            a[0] += fl
            a[1] += t_ms
            a[2] += 1
        conv = by_kind.get("conv_gemm", [0.0, 1e-9, 1])
        wg = by_kind.get("wgrad_gemm", [0.0, 1e-9, 1])
        conv_tf = conv[0] / (conv[1] * 1e-3) / 1e12
        wg_tf = wg[0] / (wg[1] * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": conv_tf, "peak": peak_sus, "unit": "TFLOP/s",
                "frac": conv_tf / peak_sus, "traffic": None,
                "kernel": "conv_gemm2_kernel / conv_halo2_kernel / conv_halo_kernel / conv_gemm_kernel (tcgen05 implicit GEMM: forward + data gradients)",
                "launches_per_step": conv[2] // kp, "ms_per_step": conv[1] / kp,
                "alg_flops_per_step": conv[0] / kp,
                "peak_source": "%s sustained cuBLAS bf16 (MEASURED_PEAKS.json), kernel timed inside a long step" % src}
        f_alg = 3 * sum(unet.plan_flops(CFG["num_layers"], CFG["root_size"], CFG["dilated_layers"],
                                        CFG["patch_size"]).values())
        net_tf = value / world * f_alg / 1e12
        extra = {
            "roofline_wgrad": {"bound": "tensor", "achieved": wg_tf, "peak": peak_sus, "unit": "TFLOP/s",
                               "frac": wg_tf / peak_sus, "launches_per_step": wg[2] // kp,
                               "ms_per_step": wg[1] / kp, "kernel": "wgrad_gemm2_kernel / wgrad_halo_kernel / wgrad_gemm_kernel"},
            "roofline_network": {"bound": "tensor", "achieved": net_tf, "peak": peak, "unit": "TFLOP/s",
                                 "frac": net_tf / peak, "frac_of_sustained": net_tf / peak_sus,
                                 "alg_flops_per_patch": f_alg,
                                 "note": "per-GPU patches/s x 3 x F_min / measured burst bf16 peak"},
            "other_kernels_ms_per_step": ms_per_step - (conv[1] + wg[1]) / kp,
        }

    # ---- HBM-bound kernels of the path (geometry helpers, head, pool, SGD): achieved GB/s against
    # the measured copy bandwidth (north_star item 3); single GPU, rank 0 only, no collectives
    hbm = None
    if rank == 0 and not args.no_hbm:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_hbm
            res = bench_hbm.run()
            hbm = {"peak_gbs": res["peak_gbs"], "peak_source": res["peak_source"],
                   "kernels": {r["kernel"]: {"gbs": round(r["gbs"], 1), "frac": round(r["frac"], 3),
                                             "ms": round(r["ms"], 4), "alg_bytes": r["alg_bytes"]}
                               for r in res["kernels"]}}
        except Exception as e:  # the headline numbers above must survive a failure here
            hbm = {"error": repr(e)}
    # DRAM traffic of the dominant kernel class per step, from the committed ncu pass of the same
    # step -- accepted only if that capture saw the kernels this run launched (same names, same
    # launches per step); a capture of another kernel mix is reported as stale, not as a number
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r2_step_traffic.json")
    if rank == 0 and roof is not None:
        conv_names = ("conv_gemm_kernel", "conv_gemm2_kernel", "conv_halo_kernel", "conv_halo2_kernel", "first_conv_kernel")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f)
            seen = {k: v["launches"] for k, v in traffic.get("by_kernel", {}).items() if k in conv_names}
            ran = {k: int(round(v)) for k, v in launch_hist.items() if k in conv_names}
            if seen == ran:
                roof["traffic"] = traffic.get("conv_class_dram_bytes_per_step")
                roof["traffic_source"] = traffic.get("source")
            else:
                roof["traffic"] = None
                roof["traffic_source"] = "profiles/r2_step_traffic.json is stale: it holds %s, this run launched %s" \
                    % (seen, ran)
        else:
            roof["traffic_source"] = "no ncu capture committed for this kernel mix"

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, sec, threads = time_oracle(2, 1)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "2 timed steps (after 1 warm-up) of batch 1 of the same L=6 dilated 764^2->388^2 "
                         "model, fwd+bwd+momentum update, fp32 torch-CPU oracle (%.1f s/step)" % sec}
        if not args.no_predict:
            try:  # prediction half of the metric + the NumPy helpers at config-3 sizes
                cpu["predict"] = time_oracle_predict()
            except Exception as e:
                cpu["predict"] = {"error": repr(e)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(B, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": ke, "api": "tf_aerial_images.ConvolutionalModel.prefetch + train_batch, as its epoch loop does "
                                        "(pinned host batch in, next batch's copy overlapped with the step, "
                                        "loss + probabilities read back every step)",
                    # the second half of the metric ("... & sliding-window predict Mpix/s") through
                    # ConvolutionalModel.predict: host image in, host mask out, all ranks
                    "predict": None if predict is None else {
                        "value": predict["value"], "unit": "Mpix/s", "seconds": predict["seconds"],
                        "window_loop_value": predict["window_loop"]["value"],
                        "patch_forwards_per_s": predict["window_loop"]["patch_forwards_per_s"],
                        "config": predict["config"]},
                    # BASELINE.json configs[3]: the same loop with --image_augmentation
                    "with_image_augmentation": {"value": e2e_aug_value, "unit": UNIT}},
            "gpu_launches": launches, "launches_per_step": launch_hist, "loss": loss_val, "clocks": clocks,
            "roofline": roof, "cpu_baseline": cpu, "predict": predict, "hbm_kernels": hbm,
        }
        line.update(extra)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default 32 = the named config)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--quick", action="store_true", help="device-timed region only (for ncu runs)")
    ap.add_argument("--no-predict", action="store_true", help="skip the sliding-window prediction leg")
    ap.add_argument("--no-hbm", action="store_true", help="skip the HBM-bound kernel table")
    ap.add_argument("--predict-images", type=int, default=1, help="604^2 images in the prediction leg")
    ap.add_argument("--large-tiles", action="store_true",
                    help="run the BASELINE.json configs[4] sweep (P = 1028 / 2052, root 64 / 128) instead of "
                         "the prediction / roofline / CPU legs")
    ap.add_argument("--dump-layers", default="", help="write the per-layer tcgen05 kernel timing table here")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON record: everything else that libraries write to
    # file descriptor 1 (NCCL prints its version banner there) is routed to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            if args.impl == "reference":
                run_reference(args)
            else:
                run_gpu(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
        lines = [ln for ln in buf.getvalue().splitlines() if ln.strip()]
        for ln in lines[:-1]:
            print(ln, file=sys.stderr)
        if lines:
            print(lines[-1], flush=True)


if __name__ == "__main__":
    main()
