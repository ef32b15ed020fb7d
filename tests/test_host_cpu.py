"""Host-side logic that needs no GPU: the C-ABI library loads and exports every symbol the header
declares, the flag surface matches the reference, FLOP accounting matches SURVEY.md, and the
data-parallel plumbing (gradient buckets + rank sharding) works at world_size 2 over gloo."""
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_header_symbols():
    from road_segmentation_unet_b200 import _lib
    lib = _lib.load()  # no compute calls: loading must work without a GPU
    header = open(os.path.join(ROOT, "include", "rsu_b200.h")).read()
    declared = set(re.findall(r"^(?:unsigned int|int|void|long long|const char\*)\s+(rsu_[a-z0-9_]+)\s*\(", header, re.M))
    assert len(declared) >= 25
    for name in sorted(declared):
        assert hasattr(lib, name), "librsu_b200.so does not export %s" % name
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert lib.rsu_version() >= 100
    assert lib.rsu_last_error() is not None


def test_no_cpu_fallback_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from road_segmentation_unet_b200 import unet
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        unet.UNet(3, 64, False, 1, 60)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "road_segmentation_unet_b200")
    for name in os.listdir(pkg):
        if name.endswith(".py"):
            src = open(os.path.join(pkg, name)).read()
            assert "oracle" not in src.replace("the oracle", "").replace("CPU oracle", ""), name


def test_flags_match_reference_defaults():
    """tf_aerial_images.py:15-46 -- names and defaults verbatim."""
    from road_segmentation_unet_b200 import tf_aerial_images as tfa
    o = tfa.Options()
    expect = dict(batch_size=25, dilated_layers=False, dropout=0.8, ensemble_prediction=False,
                  eval_data_dir=None, eval_every=500, eval_train=False, gpu=-1, image_augmentation=False,
                  interactive=False, lr=0.01, model_path=None, momentum=0.9, num_epoch=5,
                  num_eval_images=4, num_gpu=1, num_layers=5, patch_size=128, pred_batch_size=2,
                  restore_date=None, restore_epoch=None, restore_model=False, root_size=64,
                  rotation_angles=None, seed=2017, stride=16, train_score_every=1000)
    for k, v in expect.items():
        assert getattr(o, k) == v, k
    assert len(tfa.FLAG_DEFS) == 30
    f = tfa.make_parser().parse_args(["--rotation_angles", "15,30", "--dilated_layers", "--num_layers", "6"])
    o = tfa.Options(f)
    assert o.rotation_angles == [15, 30] and o.dilated_layers is True and o.num_layers == 6


def test_constants():
    from road_segmentation_unet_b200 import constants as c
    assert (c.FOREGROUND_THRESHOLD, c.IMG_PATCH_SIZE, c.NUM_CHANNELS, c.NUM_LABELS, c.PIXEL_DEPTH) == \
        (.25, 16, 3, 2, 255)


def test_input_size_and_flops():
    from road_segmentation_unet_b200 import unet
    assert [unet.input_size_needed(388, L) for L in (4, 5, 6)] == [476, 572, 764]
    with pytest.raises(AssertionError):
        unet.input_size_needed(390, 4)
    fl = unet.plan_flops(6, 64, True, 388)
    assert abs(sum(fl.values()) / 1e9 - 680.82) < 0.01          # F_min, SURVEY.md 8(d)
    assert "conv_dilut_5/atrous_conv1" not in fl                   # dead pair counts zero
    assert abs(sum(unet.plan_flops(4, 64, False, 388).values()) / 1e9 - 195.59) < 0.01
    shapes = unet.variable_shapes(6, 64, True)
    assert sum(int(np.prod(s)) for s in shapes.values()) == 212403278
    a = unet.glorot_init(3, 64, True, 2017)
    b = unet.glorot_init(3, 64, True, 2017)
    assert all(np.array_equal(a[k], b[k]) for k in a)


def test_host_visual_and_io_helpers(tmp_path):
    """The host-side helpers around the path against goldens made by
    the reference's own images.py: overlays, confusion / error images, predictions_to_patches,
    patch scores, PNG save -> load round trip.  (The device scoring rules are GPU tests.)"""
    from road_segmentation_unet_b200 import images
    from oracle import images_oracle as IO
    G = np.load(os.path.join(ROOT, "tests", "golden", "images_golden.npz"))
    assert np.array_equal(images.overlays(G["overlay_img"], G["overlay_mask"]), G["overlay_out_095"])
    assert np.array_equal(images.overlays(G["overlay_img"], G["overlay_mask"], fade=0.4), G["overlay_out_040"])
    assert np.array_equal(images.overlap_pred_true(G["confusion_pred"], G["confusion_true"]), G["confusion_out"])
    assert np.array_equal(images.overlapp_error(G["confusion_pred"], G["confusion_true"]), G["error_out"])
    assert np.array_equal(images.predictions_to_patches(G["pred_to_patches_in"], 4), G["pred_to_patches_out"])
    assert images.predictions_to_patches(np.array([0, 1]), 2).shape == (2, 2, 2, 1)
    assert IO.patch_f1(G["quant_in"], G["quant_in"]) == 1.0
    # accuracy / recall / precision / F1 of summary.py:141-147
    acc, rec, prec, f1 = images.patch_scores([1, 1, 0, 0, 1], [1, 0, 0, 1, 1])
    assert (acc, rec, prec) == (0.6, 2 / 3, 2 / 3) and abs(f1 - 2 / 3) < 1e-12
    assert images.patch_scores([0, 0], [0, 1])[3] == 0.0
    # PNG round trip: RGB float, RGBA overlays, greyscale masks (min/max normalised like imsave)
    rgb = np.random.RandomState(0).randint(0, 256, (2, 6, 6, 3)).astype(np.uint8)
    images.save_all(rgb.astype(np.float32) / 255, str(tmp_path / "rgb"))
    back = images.load(str(tmp_path / "rgb"))
    assert back.dtype == np.float32 and back.shape == (2, 6, 6, 4)
    assert np.array_equal(images.img_float_to_uint8(back[..., :3]), rgb) and np.all(back[..., 3] == 1)
    assert sorted(os.listdir(tmp_path / "rgb")) == ["images_001.png", "images_002.png"]
    images.save_all(G["overlay_out_095"], str(tmp_path / "ov"), "o_{:03d}.png")
    assert np.array_equal(images.img_float_to_uint8(images.load(str(tmp_path / "ov"))), G["overlay_out_095"])
    binary = (np.random.RandomState(1).rand(1, 8, 8, 1) > 0.5) * 1
    images.save_all(binary, str(tmp_path / "bin"), "b_{:03d}.png", greyscale=True)
    grey = images.load(str(tmp_path / "bin"))
    assert np.array_equal(grey[0, :, :, 0], binary[0, :, :, 0].astype(np.float32))
    with pytest.raises(NotImplementedError):
        images.save_all(binary, str(tmp_path / "bin"))
    # greyscale groundtruth-style file -> 2-D float32 in [0, 1]
    from PIL import Image
    os.makedirs(tmp_path / "gt" / "groundtruth")
    os.makedirs(tmp_path / "gt" / "images")
    Image.fromarray(rgb[0, :, :, 0], "L").save(tmp_path / "gt" / "groundtruth" / "a.png")
    Image.fromarray(rgb[0], "RGB").save(tmp_path / "gt" / "images" / "a.png")
    im, gt = images.load_train_data(str(tmp_path / "gt"))
    assert im.shape == (1, 6, 6, 3) and gt.shape == (1, 6, 6)
    assert np.array_equal(gt[0], rgb[0, :, :, 0].astype(np.float32) / np.float32(255))


def test_shared_window_plan():
    """every sliding-window position is covered exactly once, members of a window are one
    alignment period apart, and the enlarged input respects the cap"""
    from road_segmentation_unet_b200 import tf_aerial_images as tfa
    for side, stride, S, L, cap in ((19, 12, 764, 6, 1400), (54, 12, 764, 6, 1400), (5, 12, 124, 4, 1400),
                                    (2, 12, 76, 3, 1400), (9, 110, 764, 6, 1400), (30, 16, 316, 5, 500),
                                    (7, 32, 764, 6, 1400)):
        plan = tfa.shared_window_plan(side, stride, S, L, cap)
        period = 2 ** (L - 1)
        if plan is None:
            continue
        n, q, wins = plan
        assert q % period == 0 and q % stride == 0 and n >= 2
        assert S + q * (n - 1) <= max(cap, S)
        assert sorted(k for w in wins for k in w) == list(range(side))
        for w in wins:
            assert 1 <= len(w) <= n and all((b - a) * stride == q for a, b in zip(w, w[1:]))
    assert tfa.shared_window_plan(19, 12, 764, 6)[0] == 3          # BASELINE configs[2]
    assert tfa.shared_window_plan(19, 12, 764, 6, max_input=800) is None
    assert tfa.shared_window_plan(1, 12, 764, 6) is None
    assert tfa.shared_window_plan(7, 32, 764, 6)[1] == 32          # stride = period: one class
    # the jobs of all ranks together evaluate every (image, x position, y position) exactly once,
    # and a job runs at the size of its longer window
    n, q, wins = tfa.shared_window_plan(19, 12, 764, 6)
    for world in (1, 2, 3, 8):
        seen = []
        for rank in range(world):
            for m, group in tfa.shared_window_jobs(6, wins, rank, world).items():
                for img, wx, wy in group:
                    assert max(len(wx), len(wy)) == m <= n
                    seen.extend((img, kx, ky) for kx in wx for ky in wy)
        assert sorted(seen) == [(i, kx, ky) for i in range(6) for kx in range(19) for ky in range(19)]
    sizes = {m: len(g) for m, g in tfa.shared_window_jobs(6, wins).items()}
    assert sizes == {2: 150, 3: 234}                               # bench.py's predict.mode line


def test_shard_helpers():
    from road_segmentation_unet_b200 import tf_aerial_images as tfa
    for n in (1, 7, 2166, 17328):
        for world in (1, 2, 4, 8):
            cover = []
            for r in range(world):
                k0, k1 = tfa.shard_range(n, r, world)
                cover.extend(range(k0, k1))
            assert cover == list(range(n))
    idx = np.arange(100)
    got = [tfa.rank_batch_indices(idx, 32, r, 8).tolist() for r in range(4)]
    assert sum(got, []) == list(range(32, 64))


def test_shard_pairs_cover_the_training_set():
    """every rank prepares only its (angle, image) pairs; the shards of all
    ranks, concatenated in rank order, are expand_and_rotate's angle-major order exactly once."""
    from road_segmentation_unet_b200 import tf_aerial_images as tfa
    for n_img, angles in ((100, [15, 30, 45, 60, 75]), (3, [0, 90]), (10, [0]), (1, [15, 30, 45])):
        full = [(a, i) for a in angles for i in range(n_img)]
        for world in (1, 2, 3, 4, 8):
            got = []
            for r in range(world):
                for angle, idx in tfa.shard_pairs(n_img, angles, r, world):
                    assert idx == sorted(idx) and len(idx) >= 1
                    got.extend((angle, i) for i in idx)
            assert got == full, (n_img, angles, world)
            sizes = [sum(len(idx) for _, idx in tfa.shard_pairs(n_img, angles, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_dp_rank_slices_partition_the_live_parameters():
    """dp.rank_slices: the slices of all ranks tile every live range of the flat parameter vector
    exactly once, at 16-byte boundaries, whatever the world size (the fused reduce + momentum +
    broadcast kernel of each rank works on its own slices)."""
    from road_segmentation_unet_b200 import unet
    from road_segmentation_unet_b200.dp import rank_slices
    offsets, n_flat = unet.flat_layout(6, 64, True)
    d0, d1 = offsets["conv_dilut_5/atrous_conv1/kernel"], offsets["conv_5/conv1/kernel"]
    live = [(0, d0), (d1, n_flat)]
    assert n_flat - (d1 - d0) >= 155776078          # live parameters of the flagship model (+ padding)
    for world in (1, 2, 3, 4, 8):
        cover = []
        for r in range(world):
            for lo, hi in rank_slices(live, r, world):
                assert lo % 4 == 0 and hi % 4 == 0 and hi > lo
                cover.append((lo, hi))
        cover.sort()
        merged = [list(cover[0])]
        for lo, hi in cover[1:]:
            if lo == merged[-1][1]:
                merged[-1][1] = hi
            else:
                assert lo > merged[-1][1], "overlapping slices"
                merged.append([lo, hi])
        assert [tuple(m) for m in merged] == live
        sizes = [sum(hi - lo for lo, hi in rank_slices(live, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 8


def test_dp_exchange_buckets_match_the_backward_pass():
    """dp.PeerOptimizer exchanges the gradient bucket by bucket while the backward pass runs: the
    pieces UNet._ready hands to the hook (live parts of the decoder / encoder buckets, in backward
    order) must be exactly PeerOptimizer.buckets(), they must tile the live ranges, and the
    per-bucket ownership of all ranks must partition them (no GPU needed: layout logic only)."""
    import types
    from road_segmentation_unet_b200 import unet
    from road_segmentation_unet_b200.dp import PeerOptimizer, rank_slices
    for L, dil in ((6, True), (4, False), (3, True)):
        n = types.SimpleNamespace(L=L, dilated=dil, root=64, _live_ranges=None)
        n.offsets, n.n_flat = unet.flat_layout(L, 64, dil)
        n._first_offset = lambda prefix, n=n: next(o for name, o in n.offsets.items() if name.startswith(prefix))
        for meth in ("live_ranges", "_bucket_bounds", "_ready"):
            setattr(n, meth, types.MethodType(getattr(unet.UNet, meth), n))
        got = []
        n.on_bucket_ready = lambda a, b: got.append((a, b))
        enc, dec = n._bucket_bounds()
        for b in dec[::-1] + enc[::-1]:      # the order backward() finishes them
            n._ready(b)
        po = PeerOptimizer.__new__(PeerOptimizer)
        po.rank, po.world = 0, 8
        assert sorted(got) == po.buckets(n)
        covered = sum(b - a for a, b in got)
        assert covered == sum(b - a for a, b in n.live_ranges())
        for world in (2, 8):
            owned = []
            for r in range(world):
                po.rank, po.world = r, world
                owned += po.owned(n)
            owned.sort()
            assert all(a2 >= b1 for (_, b1), (a2, _) in zip(owned, owned[1:]))   # disjoint
            assert sum(b - a for a, b in owned) == covered


def test_streaming_metrics_match_tf_metrics_semantics():
    """summary.py:141-147 / tf_aerial_images.py:428: tf.metrics.* accumulate counts over calls and
    are zeroed once per epoch; F1 = 2 / (1/recall + 1/precision); the zero entries the reference
    appends through ndarray.resize inflate the accuracy only."""
    from road_segmentation_unet_b200.summary import StreamingMetrics
    m = StreamingMetrics()
    acc, rec, prec, f1 = m.update([1, 1, 0, 0, 1], [1, 0, 0, 1, 1])
    assert (acc, rec, prec) == (0.6, 2 / 3, 2 / 3) and abs(f1 - 2 / 3) < 1e-12
    acc, rec, prec, f1 = m.update([1, 0, 0], [1, 0, 1])      # running totals: 8 labels, tp 3, fn 1, fp 2
    assert acc == 5 / 8 and rec == 3 / 4 and prec == 3 / 5
    assert abs(f1 - 2 / (4 / 3 + 5 / 3)) < 1e-12
    m.reset()
    assert m.update([0, 0], [0, 1]) == (0.5, 0.0, 0.0, 0.0)   # 1/0 -> inf -> F1 0 like TensorFlow
    m.reset()
    acc, rec, prec, f1 = m.update([1, 0], [1, 1], padded_zeros=2 * 255)
    assert acc == (1 + 510) / 512 and rec == 1.0 and prec == 0.5


def test_tf_checkpoint_bundle_round_trip(tmp_path):
    """TensorFlow V2 checkpoint bundle writer / importer (tf.train.Saver's files,
    tf_aerial_images.py:343-379): table magic, masked CRC-32C known answers, multi-block index,
    scalars, every dtype the model stores; corruption is detected."""
    from road_segmentation_unet_b200 import tf_checkpoint as T
    assert T.crc32c(b"123456789") == 0xE3069283                     # CRC-32C check value
    assert T.crc32c(b"") == 0 and T.crc32c(b"6789", T.crc32c(b"12345")) == 0xE3069283
    assert T.crc32c(bytes(32)) == 0x8A9136AA                         # RFC 3720 B.4 test vector
    assert T.mask_crc(0) == 0xA282EAD8
    rs = np.random.RandomState(0)
    t = {"conv_0/conv1/kernel": rs.rand(3, 3, 3, 64).astype(np.float32),
         "conv_0/conv1/kernel/Momentum": rs.rand(3, 3, 3, 64).astype(np.float32),
         "global_step": np.array(1234, np.int32), "d": rs.rand(5), "i": np.arange(7, dtype=np.int64)}
    for i in range(6000):  # > 256 KiB of index entries: several data blocks + restart points
        t["scope_%05d/bias" % i] = rs.rand(3).astype(np.float32)
    prefix = str(tmp_path / "model-epoch-003.chkpt")
    T.write_bundle(prefix, t)
    assert T.is_bundle(prefix) and os.path.exists(prefix + ".data-00000-of-00001")
    raw = open(prefix + ".index", "rb").read()
    assert raw[-8:] == bytes.fromhex("57fb808b247547db")            # LevelDB table magic, little endian
    assert len(T.read_table(prefix + ".index")) == len(t) + 1      # + the header entry under ""
    back = T.read_bundle(prefix)
    assert set(back) == set(t)
    for k in t:
        assert back[k].dtype == t[k].dtype and back[k].shape == np.asarray(t[k]).shape, k
        assert np.array_equal(back[k], t[k]), k
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[100] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(data)
    with pytest.raises(ValueError, match="checksum"):
        T.read_bundle(prefix)
    idx = bytearray(raw)
    idx[50] ^= 1
    open(prefix + ".index", "wb").write(idx)
    with pytest.raises(ValueError, match="checksum"):
        T.read_table(prefix + ".index")
    with pytest.raises(ValueError, match="magic"):
        open(prefix + ".index", "wb").write(b"\0" * 64)
        T.read_table(prefix + ".index")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _dp_worker(rank, world, port, bounds, n_flat, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from road_segmentation_unet_b200 import tf_aerial_images as tfa
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.arange(n_flat, dtype=torch.float32) * (rank + 1)
    red = tfa.GradientAllReducer(dist, world, g)
    for (a, b) in bounds:  # buckets become ready in backward order
        red.bucket_ready(a, b)
    scale = red.finish()
    np.save(os.path.join(out_dir, "g%d.npy" % rank), (g * scale).numpy())
    dist.destroy_process_group()


def test_dp_gradient_buckets_gloo(tmp_path):
    """world_size 2 on CPU: bucketed all-reduce == mean over ranks, every element exactly once."""
    import torch.multiprocessing as mp
    n_flat = 1000
    bounds = [(700, 1000), (300, 700), (0, 300)]
    port = _free_port()
    mp.spawn(_dp_worker, args=(2, port, bounds, n_flat, str(tmp_path)), nprocs=2, join=True)
    expect = np.arange(n_flat, dtype=np.float32) * 1.5  # mean of x*1 and x*2
    for r in range(2):
        assert np.allclose(np.load(str(tmp_path / ("g%d.npy" % r))), expect)


def test_overlays_reproduce_the_reference_submission_images():
    """images.overlays (host side, PIL) on the real test-image windows + the masks of the reference's
    own submissions == the overlay PNGs the reference committed (tests/golden/make_submission_golden.py)."""
    from road_segmentation_unet_b200 import images
    S = np.load(os.path.join(ROOT, "tests", "golden", "submission_golden.npz"))
    side = S["crop_rgb"].shape[1]
    for r in range(S["crop_overlay"].shape[0]):
        masks = np.kron(S["labels"][r].transpose(0, 2, 1), np.ones((16, 16), np.uint8)).astype(np.float64)
        for j, (k, top, left) in enumerate(S["crops"]):
            img = S["crop_rgb"][j:j + 1].astype(np.float32) / np.float32(255)
            m = masks[k:k + 1, top:top + side, left:left + side, None]
            assert np.array_equal(images.overlays(img, m, fade=0.4)[0], S["crop_overlay"][r, j])


def _header_prototypes():
    """{name: (return type, [parameter types])} parsed from include/rsu_b200.h (comments stripped)."""
    text = open(os.path.join(ROOT, "include", "rsu_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    protos = {}
    for ret, name, params in re.findall(
            r"^\s*((?:const\s+)?(?:unsigned\s+int|long\s+long|int|void|char)\s*\*?)\s*(rsu_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;",
            text, flags=re.M | re.S):
        plist = [] if params.strip() in ("", "void") else [" ".join(p.split()) for p in params.split(",")]
        protos[name] = (" ".join(ret.split()), plist)
    return protos


def _ctype_class(param):
    """C parameter declaration -> the ctypes class the binding must use."""
    import ctypes as C
    from road_segmentation_unet_b200 import _lib
    structs = {"rsu_view": _lib.View, "rsu_conv_gemm_desc": _lib.ConvGemmDesc, "rsu_wgrad_desc": _lib.WgradDesc,
               "rsu_pack_job": _lib.PackJob, "rsu_dp_peers": _lib.DpPeers}
    decl = re.sub(r"\b[A-Za-z_][A-Za-z0-9_]*$", "", param).strip() if not param.endswith("*") else param
    decl = decl.replace("const", "").strip()
    if "*" in decl:
        base = decl.replace("*", "").strip()
        if base in structs:
            return C.POINTER(structs[base])
        if base == "int":
            return (C.POINTER(C.c_int), C.c_void_p)     # host int* or device int*: either binding
        if base == "double":
            return (C.POINTER(C.c_double), C.c_void_p)
        if base == "char":
            return C.c_char_p
        return C.c_void_p
    return {"int": C.c_int, "long long": C.c_longlong, "unsigned long long": C.c_ulonglong,
            "unsigned int": C.c_uint, "float": C.c_float, "double": C.c_double}[decl]


def test_ctypes_signatures_match_header_prototypes():
    """ctypes checks nothing at call time: a float bound as c_double, a missing argument or a
    long long bound as c_int would corrupt a call silently.  Every prototype of the header is
    compared with the binding's argtypes / restype, parameter by parameter."""
    import ctypes as C
    from road_segmentation_unet_b200 import _lib
    protos = _header_prototypes()
    assert set(protos) == set(_lib._SIGNATURES), set(protos) ^ set(_lib._SIGNATURES)
    ret_map = {"int": C.c_int, "long long": C.c_longlong, "void": None, "unsigned int": C.c_uint,
               "const char*": C.c_char_p, "const char *": C.c_char_p}
    for name, (ret, params) in sorted(protos.items()):
        res, args = _lib._SIGNATURES[name]
        assert res is ret_map[ret], (name, ret, res)
        assert len(args) == len(params), (name, len(args), params)
        for k, (p, a) in enumerate(zip(params, args)):
            want = _ctype_class(p)
            ok = a in want if isinstance(want, tuple) else a is want
            assert ok, "%s argument %d (%s): bound as %s" % (name, k, p, a)


def test_ctypes_struct_layouts_match_header(tmp_path):
    """sizeof / offsetof of every struct of the header, as gcc lays them out, against the ctypes
    Structure classes of the binding (also proves the header compiles as plain C)."""
    import subprocess
    from road_segmentation_unet_b200 import _lib
    structs = {"rsu_view": _lib.View, "rsu_conv_gemm_desc": _lib.ConvGemmDesc, "rsu_wgrad_desc": _lib.WgradDesc,
               "rsu_pack_job": _lib.PackJob, "rsu_dp_peers": _lib.DpPeers}
    text = re.sub(r"/\*.*?\*/", " ", open(os.path.join(ROOT, "include", "rsu_b200.h")).read(), flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "rsu_b200.h"', 'int main(void) {']
    fields = {}
    for body, name in re.findall(r"typedef\s+struct\s*\{(.*?)\}\s*(rsu_[a-z0-9_]+)\s*;", text, flags=re.S):
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):  # "int H, W;" declares several fields
                m = re.search(r"([A-Za-z_][A-Za-z0-9_]*)\s*(?:\[[^\]]*\])?\s*$", part.strip())
                names.append(m.group(1))
        fields[name] = names
        lines.append('  printf("%s sizeof %%zu\\n", sizeof(%s));' % (name, name))
        for f in names:
            lines.append('  printf("%s %s %%zu\\n", offsetof(%s, %s));' % (name, f, name, f))
    lines += ['  return 0;', '}']
    assert set(fields) == set(structs), set(fields) ^ set(structs)
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split("\n")
    import ctypes as C
    seen = 0
    for line in out:
        if not line.strip():
            continue
        s, f, v = line.split()
        cls = structs[s]
        if f == "sizeof":
            assert C.sizeof(cls) == int(v), (s, C.sizeof(cls), v)
        else:  # fields are matched by position (`in` of rsu_pack_job is `inp` in Python)
            py_names = [n for n, _ in cls._fields_]
            assert len(py_names) == len(fields[s]), (s, py_names, fields[s])
            py = py_names[fields[s].index(f)]
            assert py == f or (py, f) == ("inp", "in"), (s, py, f)
            assert getattr(cls, py).offset == int(v), (s, f, getattr(cls, py).offset, v)
        seen += 1
    assert seen >= 60


def _bundle_proto_classes():
    """BundleHeaderProto / BundleEntryProto of tensorflow/core/protobuf/tensor_bundle.proto (with
    TensorShapeProto and VersionDef), declared to the protobuf runtime from their published field
    numbers -- an independent implementation of the wire format to hold tf_checkpoint's
    hand-written encoder / parser against."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name="rsu_test_tensor_bundle.proto", package="rsu_test", syntax="proto3")

    def msg(name, fields, nested=()):
        m = descriptor_pb2.DescriptorProto(name=name)
        for fname, num, ftype, label, type_name in fields:
            f = m.field.add(name=fname, number=num, type=ftype, label=label)
            if type_name:
                f.type_name = type_name
        for n in nested:
            m.nested_type.add().CopyFrom(n)
        return m

    opt, rep = F.LABEL_OPTIONAL, F.LABEL_REPEATED
    dim = msg("Dim", [("size", 1, F.TYPE_INT64, opt, None), ("name", 2, F.TYPE_STRING, opt, None)])
    shape = msg("TensorShapeProto", [("dim", 2, F.TYPE_MESSAGE, rep, ".rsu_test.TensorShapeProto.Dim"),
                                     ("unknown_rank", 3, F.TYPE_BOOL, opt, None)], nested=[dim])
    version = msg("VersionDef", [("producer", 1, F.TYPE_INT32, opt, None), ("min_consumer", 2, F.TYPE_INT32, opt, None),
                                 ("bad_consumers", 3, F.TYPE_INT32, rep, None)])
    header = msg("BundleHeaderProto", [("num_shards", 1, F.TYPE_INT32, opt, None), ("endianness", 2, F.TYPE_INT32, opt, None),
                                       ("version", 3, F.TYPE_MESSAGE, opt, ".rsu_test.VersionDef")])
    entry = msg("BundleEntryProto", [("dtype", 1, F.TYPE_INT32, opt, None),
                                     ("shape", 2, F.TYPE_MESSAGE, opt, ".rsu_test.TensorShapeProto"),
                                     ("shard_id", 3, F.TYPE_INT32, opt, None), ("offset", 4, F.TYPE_INT64, opt, None),
                                     ("size", 5, F.TYPE_INT64, opt, None), ("crc32c", 6, F.TYPE_FIXED32, opt, None)])
    for m in (shape, version, header, entry):
        fd.message_type.add().CopyFrom(m)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("rsu_test." + n))
    return get("BundleHeaderProto"), get("BundleEntryProto")


def test_tf_checkpoint_wire_formats_against_protobuf_and_snappy():
    """The hand-written protobuf encoder / parser and the snappy decoder of tf_checkpoint.py against
    independent implementations: the protobuf runtime and pyarrow's snappy codec."""
    from road_segmentation_unet_b200 import tf_checkpoint as T
    Header, Entry = _bundle_proto_classes()
    h = Header.FromString(T._encode_header())
    assert (h.num_shards, h.endianness, h.version.producer, h.version.min_consumer) == (1, 0, 1, 0)
    assert h.SerializeToString() == T._encode_header()            # canonical proto3 bytes
    cases = [(T.DT_FLOAT, (3, 3, 64, 128), 0, 3 * 3 * 64 * 128 * 4, 0xDEADBEEF),
             (T.DT_FLOAT, (2, 2, 1024, 2048), 5_000_000_000, 2 * 2 * 1024 * 2048 * 4, 1),
             (T.DT_INT32, (), 849_613_112, 4, 0x80000000), (T.DT_INT64, (7,), 12, 56, 0)]
    for dtype, shape, offset, size, crc in cases:
        raw = T._encode_entry(dtype, shape, offset, size, crc)
        e = Entry.FromString(raw)
        assert (e.dtype, [d.size for d in e.shape.dim], e.shard_id, e.offset, e.size, e.crc32c) == \
            (dtype, list(shape), 0, offset, size, crc)
        assert e.HasField("shape")                                # scalars carry an empty shape message
        if crc:  # (proto3 omits a zero fixed32; the encoder always writes the checksum)
            assert e.SerializeToString() == raw
        # ... and the parser reads what the protobuf runtime writes (entries TensorFlow would write)
        e2 = Entry(dtype=dtype, shard_id=0, offset=offset, size=size, crc32c=crc)
        e2.shape.SetInParent()
        for s in shape:
            e2.shape.dim.add(size=s)
        p = T._parse_entry(e2.SerializeToString())
        assert (p["dtype"], p["shape"], p["shard_id"], p["offset"], p["size"], p["sliced"]) == \
            (dtype, list(shape), 0, offset, size, False)
        assert p["crc32c"] == (crc if crc else None)
    # a zero-sized dimension is omitted on the wire by proto3 writers: the parser must keep the dim
    e3 = Entry(dtype=T.DT_FLOAT, size=0, crc32c=5)
    e3.shape.dim.add(size=0)
    e3.shape.dim.add(size=4)
    assert T._parse_entry(e3.SerializeToString())["shape"] == [0, 4]
    # snappy-compressed index blocks (TensorFlow's table builder may use them)
    pa = pytest.importorskip("pyarrow")
    rs = np.random.RandomState(0)
    for data in (b"", b"a", b"conv_0/conv1/kernel/Momentum" * 500,
                 bytes(rs.randint(0, 256, 70000, dtype=np.uint8)),
                 b"ab" * 40000 + bytes(rs.randint(0, 4, 9000, dtype=np.uint8)) + b"\0" * 200000):
        assert T._snappy_uncompress(pa.compress(data, codec="snappy", asbytes=True)) == data


def test_tf_checkpoint_reads_snappy_compressed_tables(tmp_path):
    """An index whose blocks are snappy-compressed (block type 1, as LevelDB-format writers may
    produce) reads like the uncompressed one; the compressor is pyarrow's, not ours."""
    import struct
    pa = pytest.importorskip("pyarrow")
    from road_segmentation_unet_b200 import tf_checkpoint as T
    items = [(b"", T._encode_header())] + [
        (("scope_%04d/kernel" % i).encode(), T._encode_entry(T.DT_FLOAT, (3, 3, 64, 64), i * 147456, 147456, i + 1))
        for i in range(300)]
    out = bytearray()

    def emit(block):
        comp = pa.compress(block, codec="snappy", asbytes=True)
        off = len(out)
        out.extend(comp)
        out.extend(b"\x01" + struct.pack("<I", T.mask_crc(T.crc32c(comp + b"\x01"))))
        return T._put_varint(off) + T._put_varint(len(comp))

    index = T._BlockBuilder()
    for lo in range(0, len(items), 100):          # several data blocks
        blk = T._BlockBuilder()
        for k, v in items[lo:lo + 100]:
            blk.add(k, v)
        index.add(items[min(lo + 99, len(items) - 1)][0], emit(blk.finish()))
    footer = emit(T._BlockBuilder().finish()) + emit(index.finish())
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", T.TABLE_MAGIC))
    path = tmp_path / "snappy.index"
    path.write_bytes(bytes(out))
    assert T.read_table(str(path)) == items
    plain = tmp_path / "plain.index"
    T.write_table(str(plain), items)
    assert T.read_table(str(plain)) == items


def test_dp_peer_optimizer_protocol_simulated(monkeypatch):
    """Host logic of dp.PeerOptimizer without a GPU: streams, events, the symmetric-memory handles and
    the kernel entry point are replaced by recorders, then `world` ranks run the calls of a training
    step (arm -> bucket hooks in backward order -> step) with and without the overlapped exchange.
    Every rank must issue the SAME sequence of barriers (a rank-dependent sequence deadlocks), the
    launches of all ranks must update every live parameter exactly once with the same scalars, an
    overlapped bucket must be preceded by its own barrier on the side stream, and the weights-published
    barrier must come last."""
    import contextlib
    import ctypes as C
    import types
    from road_segmentation_unet_b200 import _lib, dp, unet

    class Stream:
        def __init__(self, tag):
            self.tag = tag

        def wait_event(self, ev):
            log.append(("wait_event", self.tag, ev.stream))

        def wait_stream(self, other):
            log.append(("wait_stream", self.tag, other.tag))

    class Event:
        def record(self, stream):
            self.stream = stream.tag

    state = {"cur": None}
    main, log = Stream("main"), []
    state["cur"] = main

    @contextlib.contextmanager
    def use_stream(s):
        prev, state["cur"] = state["cur"], s
        try:
            yield
        finally:
            state["cur"] = prev

    monkeypatch.setattr(dp.torch.cuda, "Event", Event)
    monkeypatch.setattr(dp.torch.cuda, "current_stream", lambda: state["cur"])
    monkeypatch.setattr(dp.torch.cuda, "stream", use_stream)

    class Handle:
        def __init__(self, name):
            self.name = name

        def barrier(self, channel=0, timeout_ms=0):
            assert timeout_ms > 0
            log.append(("barrier", self.name, channel, state["cur"].tag))

    class Lib:
        @staticmethod
        def rsu_dp_momentum_sgd(peers, acc, lo, hi, lr, momentum, scale, stream):
            log.append(("sgd", lo, hi, lr, momentum, scale, state["cur"].tag))
            return 0

    monkeypatch.setattr(_lib, "load", lambda: Lib)
    monkeypatch.setattr(_lib, "stream_ptr", lambda: None)

    L, dil = 6, True
    net = types.SimpleNamespace(L=L, dilated=dil, root=64, _live_ranges=None,
                                momentum=types.SimpleNamespace(data_ptr=lambda: 4096))
    net.offsets, net.n_flat = unet.flat_layout(L, 64, dil)
    net._first_offset = lambda prefix: next(o for name, o in net.offsets.items() if name.startswith(prefix))
    for meth in ("live_ranges", "_bucket_bounds", "_ready"):
        setattr(net, meth, types.MethodType(getattr(unet.UNet, meth), net))
    enc, dec = net._bucket_bounds()
    live = net.live_ranges()

    def run_rank(rank, world, overlap, extra_scale):
        del log[:]
        po = dp.PeerOptimizer.__new__(dp.PeerOptimizer)
        po.rank, po.world, po.overlap, po.barrier_timeout_ms = rank, world, overlap, 600000
        po._hg, po._hp, po._side = Handle("grads"), Handle("params"), Stream("side")
        po._peers, po._acc = _lib.DpPeers(), net.momentum
        po._armed, po._done = None, []
        net.on_bucket_ready = po.bucket_ready
        # a micro-batch whose hook is off (UNet.accumulate_step) must not exchange anything
        net.on_bucket_ready = None
        for b in dec[::-1] + enc[::-1]:
            net._ready(b)
        assert log == []
        net.on_bucket_ready = po.bucket_ready
        po.arm(0.01, 0.9, extra_scale)               # UNet.backward -> on_backward_begin
        for b in dec[::-1] + enc[::-1]:
            net._ready(b)
        po.step(net, 0.01, 0.9, extra_scale=extra_scale)
        assert po._armed is None and po._done == []
        return list(log)

    for world in (2, 4, 8):
        for overlap in (False, True):
            for extra in (1.0, 0.125):
                runs = [run_rank(r, world, overlap, extra) for r in range(world)]
                barriers = [[e for e in run if e[0] == "barrier"] for run in runs]
                assert all(b == barriers[0] for b in barriers), "rank-dependent barrier sequence"
                assert barriers[0][-1] == ("barrier", "params", 1, "main")
                assert runs[0][-1] == barriers[0][-1]          # nothing follows the publishing barrier
                n_buckets = len(dp.PeerOptimizer.buckets(types.SimpleNamespace(_buckets=None), net))
                if overlap:
                    assert barriers[0][:-1] == [("barrier", "grads", 0, "side")] * n_buckets
                    assert ("wait_stream", "main", "side") in runs[0]
                else:
                    assert barriers[0][:-1] == [("barrier", "grads", 0, "main")]
                updated = []
                for run in runs:
                    last_barrier = None
                    for e in run:
                        if e[0] == "barrier":
                            last_barrier = e
                        if e[0] == "sgd":
                            _, lo, hi, lr, mu, scale, tag = e
                            assert (lr, mu) == (0.01, 0.9) and abs(scale - extra / world) < 1e-12
                            assert tag == ("side" if overlap else "main")
                            assert last_barrier is not None and last_barrier[1] == "grads"
                            updated.append((lo, hi))
                updated.sort()
                assert all(a2 >= b1 for (_, b1), (a2, _) in zip(updated, updated[1:])), "element updated twice"
                assert sum(b - a for a, b in updated) == sum(b - a for a, b in live)
                assert updated[0][0] == live[0][0] and updated[-1][1] == live[-1][1]
