"""IMAGES -- host-side mirror of the reference's src/images.py geometry helpers.

Same function names, argument names/orders/defaults, assertions and output dtypes as the
reference (NumPy in, NumPy out), but every function runs a hand-written CUDA kernel of
librsu_b200.so (csrc/geometry.cu).  The `*_dev` variants take / return torch CUDA tensors so the
prediction pipeline (tf_aerial_images.ConvolutionalModel.predict) chains them without leaving
the GPU.  Pure data movement (mirror, flips, rot90, crops, patches, nearest-neighbour rotation)
is bit-exact for float32 and float64 inputs (float64 moves as pairs of 32-bit words); the two
averaging helpers accumulate in fp64 on the device from fp32 inputs.

The scoring rules (quantize_mask, labels_for_patches, the label grid of save_submission_csv)
run the `rsu_patch_vote` kernel; the PNG / CSV helpers around the path (load, save_all, overlays,
overlap_pred_true, overlapp_error, save_submission_csv -- SURVEY.md section 8(f)) are host code
on PIL, without matplotlib.
"""
import ctypes as C
import glob
import os

import numpy as np
import torch

from . import _lib
from ._lib import call
from .constants import PIXEL_DEPTH, FOREGROUND_THRESHOLD


# ------------------------------------------------------------------ plumbing
def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _as_words(arr):
    """NumPy array -> (float32 device tensor [N,H,W,Cw], restore) where float64 travels as two
    32-bit words per element so that pure data movement stays bit-exact."""
    arr = np.ascontiguousarray(arr)
    if arr.dtype not in (np.float32, np.float64):
        arr = arr.astype(np.float64)
    dt = arr.dtype
    a4 = arr if arr.ndim == 4 else arr[..., None]
    words = a4.view(np.float32) if dt == np.float64 else a4
    t = torch.from_numpy(np.ascontiguousarray(words)).cuda()

    def restore(out_t, squeeze):
        o = out_t.cpu().numpy()
        if dt == np.float64:
            o = o.view(np.float64)
        return o[..., 0] if squeeze else o

    return t, restore, arr.ndim == 3


# ------------------------------------------------------------------ device-level helpers
def mirror_border_dev(x, n):
    """x: float32 CUDA [N,H,W,C] -> [N,H+2n,W+2n,C] (np.pad 'symmetric', images.py:269-281)."""
    N, H, W, Cc = x.shape
    out = torch.empty(N, H + 2 * n, W + 2 * n, Cc, dtype=torch.float32, device=x.device)
    call("rsu_mirror_pad", _ptr(x), N, H, W, Cc, int(n), _ptr(out))
    return out


def d4_transform_dev(x, ops_u8, repeat=1):
    """Per-image dihedral transform: out[i] = rot90(flipud(x[j]) if op&4 else x[j], k=op&3) with
    j = i % len(x); `repeat` > 1 transforms every input image that many times (ops_u8 has
    repeat * len(x) entries) without materialising the repeated input."""
    n_in, S = x.shape[0], x.shape[1]
    assert x.shape[2] == S, "square images required"
    pixel_bytes = x.element_size() * (x.shape[3] if x.dim() == 4 else 1)
    N = n_in * repeat
    assert ops_u8.numel() == N
    out = torch.empty((N,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    call("rsu_d4_transform", _ptr(x), _ptr(out), N, S, pixel_bytes, _ptr(ops_u8), n_in)
    return out


ENSEMBLE_OPS = (0, 4 | 2, 4, 1, 2, 3)  # orig, fliplr, flipud, rot90 k=1,2,3  (images.py:376-396)


def image_augmentation_ensemble_dev(x):
    n = x.shape[0]
    ops_t = torch.tensor([op for op in ENSEMBLE_OPS for _ in range(n)], dtype=torch.uint8, device=x.device)
    return d4_transform_dev(x.contiguous(), ops_t, repeat=6)  # variant-major, one launch


def extract_patches_dev(x, patch_size, stride, k_begin=0, k_count=-1, out=None):
    N, H, W, Cc = x.shape
    side = (H - patch_size) // stride + 1
    total = N * side * side
    cnt = total - k_begin if k_count < 0 else k_count
    if out is None:
        out = torch.empty(cnt, patch_size, patch_size, Cc, dtype=torch.float32, device=x.device)
    call("rsu_extract_patches", _ptr(x), N, H, W, Cc, int(patch_size), int(stride), int(k_begin),
         int(cnt), _ptr(out))
    return out


def copy_windows_dev(x, win, jobs, out):
    """out[dst] = x[img, y0:y0+win, x0:x0+win, :] (zeros outside the image) for every row
    (img, y0, x0, dst) of the int32 CUDA job table `jobs` [n, 4]  (rsu_copy_windows)."""
    N, H, W, Cc = x.shape
    assert jobs.dtype == torch.int32 and jobs.dim() == 2 and jobs.shape[1] == 4 and jobs.is_contiguous()
    assert out.shape[1:] == (win, win, Cc) and out.is_contiguous() and x.is_contiguous()
    call("rsu_copy_windows", _ptr(x), N, H, W, Cc, int(win), int(jobs.shape[0]), _ptr(jobs), _ptr(out))
    return out


def divide_by_hits_dev(sums, side, patch_size, stride):
    """In place: overlap-add partial sums [N,S,S,C] -> averages (analytic hit counts)."""
    N, S, _, Cc = sums.shape
    call("rsu_divide_by_hits", _ptr(sums), N, S, Cc, int(side), int(patch_size), int(stride))
    return sums


def images_from_patches_dev(patches, num_images, side, stride, k_begin=0, k_count=-1, normalize=True):
    """patches: float32 CUDA [k_count, P, P, C] (slice of the patch list starting at k_begin)."""
    P, Cc = patches.shape[1], patches.shape[3]
    S = (side - 1) * stride + P
    out = torch.empty(num_images, S, S, Cc, dtype=torch.float32, device=patches.device)
    call("rsu_overlap_average", _ptr(patches), num_images, side, P, Cc, int(stride), int(k_begin),
         int(k_count), int(bool(normalize)), _ptr(out))
    return out


def invert_image_augmentation_ensemble_dev(masks):
    """masks: float32 CUDA [6N,S,S] (or [6N,S,S,1]) -> [N,S,S]."""
    assert masks.shape[0] % 6 == 0
    n, S = masks.shape[0] // 6, masks.shape[1]
    out = torch.empty(n, S, S, dtype=torch.float32, device=masks.device)
    call("rsu_ensemble_invert", _ptr(masks.contiguous()), n, S, _ptr(out))
    return out


def _cos_sin_deg(angle):
    try:  # SciPy evaluates the rotation matrix with cosdg / sindg (exact at multiples of 90)
        from scipy import special
        return float(special.cosdg(angle)), float(special.sindg(angle))
    except Exception:  # pragma: no cover
        import math
        return math.cos(math.radians(angle)), math.sin(math.radians(angle))


def rotation_geometry(side, angle):
    """scipy.ndimage.rotate(reshape=True) geometry for a side x side plane: (matrix, offset,
    out_side), evaluated with the same NumPy expressions SciPy uses."""
    c, s = _cos_sin_deg(angle)
    rot = np.array([[c, s], [-s, c]])
    out_bounds = rot @ np.array([[0, 0, side, side], [0, side, 0, side]], dtype=np.float64)
    out_shape = (np.ptp(out_bounds, axis=1) + 0.5).astype(int)
    out_center = rot @ ((out_shape - 1) / 2)
    in_center = (np.array([side, side]) - 1) / 2
    return rot, in_center - out_center, int(out_shape[0])


def rotate_crop_dev(x, angle, crop=None):
    """crop_imgs(rotate_imgs(x, angle), crop) on float32 CUDA [N,H,H,C]; only the centre crop of
    the rotated image is produced (crop=None: the whole rotated image)."""
    N, H, _, Cc = x.shape
    rot, offset, side = rotation_geometry(H, angle)
    if crop is None:
        crop0, crop = 0, side
    else:
        assert side >= crop and crop % 2 == 0
        crop0 = int(side / 2) - crop // 2
    out = torch.empty(N, crop, crop, Cc, dtype=torch.float32, device=x.device)
    m = (C.c_double * 4)(rot[0, 0], rot[0, 1], rot[1, 0], rot[1, 1])
    o = (C.c_double * 2)(offset[0], offset[1])
    call("rsu_rotate_nn_crop", _ptr(x), N, H, Cc, m, o, crop0, int(crop), _ptr(out))
    return out


# ------------------------------------------------------------------ reference API (NumPy)
def img_float_to_uint8(img):
    """Transform an array of float images into uint8 images"""
    return (img * PIXEL_DEPTH).round().astype(np.uint8)


def mirror_border(images, n):
    """mirrors border n border pixels on each side and corner (images.py:269-281)"""
    t, restore, squeeze = _as_words(images)
    return restore(mirror_border_dev(t, n), squeeze)


def extract_patches(images, patch_size, stride=None, predict_patch_size=None):
    """extract square patches from a batch of images (images.py:35-85); float64 output"""
    if not predict_patch_size:
        predict_patch_size = patch_size

    assert (patch_size - predict_patch_size) % 2 == 0 and predict_patch_size <= patch_size

    if not stride:
        stride = patch_size

    num_images, image_height, image_width = images.shape[:3]
    assert image_height == image_width, "Assume square images"
    assert (image_height - patch_size) % stride == 0, "Stride sliding should cover the whole image"

    t, restore, squeeze = _as_words(images)
    out = restore(extract_patches_dev(t, patch_size, stride), squeeze)
    return out.astype(np.float64, copy=False)


def images_from_patches(patches, stride=None):
    """Transform a list of patches into images, averaging overlaps (images.py:131-164)"""
    num_images, num_patches, patch_size, _, num_channel = patches.shape

    if stride is None:
        stride = patch_size

    num_patches_side = int(np.sqrt(num_patches))
    assert np.sqrt(num_patches) == num_patches_side, "Square image assumption broken"

    p = torch.from_numpy(np.ascontiguousarray(patches, dtype=np.float32)).cuda()
    p = p.view(num_images * num_patches, patch_size, patch_size, num_channel)
    out = images_from_patches_dev(p, num_images, num_patches_side, stride)
    return out.cpu().numpy().astype(np.float64)


def rotate_imgs(imgs, angle):
    """safeguard to avoid useless rotation by 0 (images.py:313-317)"""
    if angle == 0:
        return imgs
    t, restore, squeeze = _as_words(imgs)
    return restore(rotate_crop_dev(t, angle), squeeze)


def crop_imgs(imgs, crop_size):
    """centre crop (images.py:354-373) -- a view, like the reference"""
    batch_size, height, width = imgs.shape[:3]
    assert height == width and height >= crop_size
    assert crop_size % 2 == 0
    half_crop = int(crop_size / 2)
    center = int(height / 2)
    return imgs[:, center - half_crop:center + half_crop, center - half_crop:center + half_crop]


def expand_and_rotate(imgs, angles, offset=0):
    """rotate some images by an angle, mirror image for missing part and expanding to output_size
    (images.py:320-351); angle-major float64 output"""
    has_channels = (len(imgs.shape) == 4)
    if not has_channels:
        imgs = np.expand_dims(imgs, -1)

    batch_size, height, width, num_channel = imgs.shape
    assert height == width

    output_size = height + 2 * offset
    padding = int(np.ceil(height * (np.sqrt(2) - 1) / 2)) + int(np.ceil(offset / np.sqrt(2)))

    print("Applying rotations: {} degrees... ".format(", ".join([str(a) for a in angles])))
    t, restore, _ = _as_words(imgs)
    padded = mirror_border_dev(t, padding)
    outs = []
    for angle in angles:
        if angle == 0:
            c0 = int(padded.shape[1] / 2) - output_size // 2
            assert output_size % 2 == 0
            outs.append(padded[:, c0:c0 + output_size, c0:c0 + output_size].contiguous())
        else:
            outs.append(rotate_crop_dev(padded, angle, output_size))
    rotated_imgs = restore(torch.cat(outs, dim=0), False).astype(np.float64, copy=False)
    print("Done")

    if not has_channels:
        rotated_imgs = np.squeeze(rotated_imgs, -1)

    return rotated_imgs


def image_augmentation_ensemble(imgs):
    """create ensemble of images to be predicted (images.py:376-396); float64 [6N,H,W,C]"""
    t, restore, squeeze = _as_words(imgs)
    return restore(image_augmentation_ensemble_dev(t), squeeze).astype(np.float64, copy=False)


def invert_image_augmentation_ensemble(masks):
    """assemble masks of prediction images created by `image_augmentation_ensemble`
    (images.py:399-417).  Unlike the reference the argument is not modified in place."""
    assert masks.shape[0] % 6 == 0
    m = np.ascontiguousarray(masks, dtype=np.float32)
    has_c = m.ndim == 4
    assert not has_c or m.shape[3] == 1, "masks have one channel"
    t = torch.from_numpy(m.reshape(m.shape[0], m.shape[1], m.shape[2])).cuda()
    out = invert_image_augmentation_ensemble_dev(t).cpu().numpy().astype(np.float64)
    return out[..., None] if has_c else out


# ------------------------------------------------------------------ scoring rules (device)
RULE_VOTE, RULE_MEAN = 0, 1


def patch_vote_dev(masks, patch_size, rule, vote_threshold, pixel_threshold=0.5, quantized=False,
                   labels=True):
    """masks: float32 / float64 CUDA [N,S,S] -> (quantized masks or None, uint8 labels
    [N, cells_x, cells_y] or None).  rule RULE_VOTE = quantize_mask's 'mean(v >= 0.5) > t',
    RULE_MEAN = labels_for_patches' 'mean(v) > t'."""
    assert masks.dim() == 3 and masks.shape[1] == masks.shape[2], "square single-channel masks"
    assert masks.dtype in (torch.float32, torch.float64)
    masks = masks.contiguous()
    N, S = masks.shape[0], masks.shape[1]
    g = -(-S // patch_size)
    q = torch.empty_like(masks) if quantized else None
    lab = torch.empty(N, g, g, dtype=torch.uint8, device=masks.device) if labels else None
    call("rsu_patch_vote", _ptr(masks), masks.element_size(), N, S, int(patch_size), int(rule),
         float(pixel_threshold), float(vote_threshold), _ptr(q) if quantized else None,
         _ptr(lab) if labels else None)
    return q, lab


def _masks_to_dev(masks):
    m = np.ascontiguousarray(masks)
    if m.dtype not in (np.float32, np.float64):
        m = m.astype(np.float64)
    has_c = m.ndim == 4
    assert not has_c or m.shape[3] == 1, "masks have one channel"
    return torch.from_numpy(m.reshape(m.shape[0], m.shape[1], m.shape[2])).cuda(), has_c


def labels_for_patches(patches):
    """label 1 = road when the patch mean exceeds FOREGROUND_THRESHOLD (images.py:88-99);
    patches [num, p, p] -> int64 [num]"""
    t, _ = _masks_to_dev(patches)
    _, lab = patch_vote_dev(t, t.shape[1], RULE_MEAN, FOREGROUND_THRESHOLD)
    return lab.reshape(-1).cpu().numpy().astype(np.int64)


def predictions_to_patches(predictions, patch_size):
    """Expand each prediction to a square patch (images.py:167-180)"""
    flat = np.asarray(predictions).reshape(-1)
    return np.broadcast_to(flat[:, None, None, None], (flat.shape[0], patch_size, patch_size, 1))


def quantize_mask(masks, threshold, patch_size):
    """patch_size x patch_size vote: mean(prob >= 0.5) > threshold, written back to every pixel
    of the cell (images.py:256-266); [N,S,S,1] in, same shape and dtype out"""
    t, has_c = _masks_to_dev(masks)
    q, _ = patch_vote_dev(t, patch_size, RULE_VOTE, threshold, quantized=True, labels=False)
    out = q.cpu().numpy()
    return out[..., None] if has_c else out


def patch_labels(masks, patch_size, threshold=FOREGROUND_THRESHOLD, rule=RULE_MEAN):
    """Label grid of a batch of masks: int64 [N, cells, cells] indexed [image, x cell, y cell] --
    what save_submission_csv derives through extract_patches + labels_for_patches
    (images.py:218-224)."""
    t, _ = _masks_to_dev(masks)
    assert t.shape[1] % patch_size == 0, "Stride sliding should cover the whole image"
    _, lab = patch_vote_dev(t, patch_size, rule, threshold)
    return lab.cpu().numpy().astype(np.int64)


def patch_scores(pred_labels, true_labels):
    """accuracy, recall, precision, F1 = 2 / (1/recall + 1/precision) over patch labels
    (summary.py:141-147)."""
    p = np.asarray(pred_labels).reshape(-1) > 0
    t = np.asarray(true_labels).reshape(-1) > 0
    tp, fp, fn = float(np.sum(p & t)), float(np.sum(p & ~t)), float(np.sum(~p & t))
    accuracy = float(np.mean(p == t))
    recall = tp / (tp + fn) if tp + fn > 0 else 0.0
    precision = tp / (tp + fp) if tp + fp > 0 else 0.0
    f1 = 2.0 / (1.0 / recall + 1.0 / precision) if recall > 0 and precision > 0 else 0.0
    return accuracy, recall, precision, f1


def save_submission_csv(masks, path, patch_size):
    """Save the masks in the expected format for submission (images.py:206-237): one row
    'NNN_x_y,label' per patch_size cell, image-major, x outer, y inner."""
    masks = np.asarray(masks)
    if masks.ndim == 4:
        masks = masks.squeeze(-1)
    num_mask, mask_height, mask_width = masks.shape
    assert mask_height == mask_width, "images should be square"
    labels = patch_labels(masks, patch_size)
    os.makedirs(path, exist_ok=True)
    filename = os.path.abspath(os.path.join(path, "submission.csv"))
    print("Saving predictions in {}".format(filename))
    cells = np.arange(labels.shape[1]) * patch_size
    rows = ["id,prediction"]
    for n in range(num_mask):
        rows.extend("{:03d}_{}_{},{}".format(n + 1, x, y, labels[n, j, i])
                    for j, x in enumerate(cells) for i, y in enumerate(cells))
    with open(filename, "w") as f:
        f.write("\n".join(rows) + "\n")
    print("Done")


# ------------------------------------------------------------------ image files and visual dumps
def _png_to_float(img):
    """PIL image -> float32 array in [0, 1] with matplotlib.image.imread's PNG conventions
    (8-bit / 255, 16-bit / 65535, palettes expanded, channels kept)."""
    if img.mode == "P":
        img = img.convert("RGBA" if "transparency" in img.info else "RGB")
    if img.mode in ("I;16", "I;16B", "I;16L", "I"):
        return (np.asarray(img, dtype=np.float64) / 65535.0).astype(np.float32)
    if img.mode == "1":
        img = img.convert("L")
    return np.asarray(img, dtype=np.float32) / np.float32(255.0)


def load(directory):
    """Extract the images in `directory` into a tensor [num_images, height, width(, channels)]
    (images.py:24-32), float32 in [0, 1], files in sorted order"""
    from PIL import Image
    print('Loading images from {} ...'.format(directory))
    files = sorted(glob.glob(os.path.join(directory, '*.png')))
    loaded = []
    for file_path in files:
        with Image.open(file_path) as img:
            loaded.append(_png_to_float(img))
    print("Loaded {} images from {}".format(len(loaded), directory))
    return np.asarray(loaded)


def load_train_data(directory):
    """images of `directory`/images and masks of `directory`/groundtruth (images.py:240-253)"""
    return (load(os.path.abspath(os.path.join(directory, 'images/'))),
            load(os.path.abspath(os.path.join(directory, 'groundtruth/'))))


def overlays(imgs, masks, fade=0.95):
    """Add the masks on top of the images with red transparency (images.py:102-128):
    RGBA uint8 [N,H,W,4]; the red layer's alpha is uint8(mask) * fade, truncated."""
    from PIL import Image
    num_images, im_height, im_width, num_channel = imgs.shape
    assert num_channel == 3, 'Predict image should be colored'
    base = img_float_to_uint8(imgs)
    alpha = (img_float_to_uint8(np.asarray(masks).reshape(num_images, im_height, im_width)) * fade).astype(np.uint8)
    red = np.zeros((im_height, im_width, 4), dtype=np.uint8)
    red[..., 0] = 255
    out = np.empty((num_images, im_height, im_width, 4), dtype=np.uint8)
    for n in range(num_images):
        red[..., 3] = alpha[n]
        out[n] = np.asarray(Image.alpha_composite(Image.fromarray(base[n]).convert('RGBA'),
                                                  Image.fromarray(red)))
    return out


def overlap_pred_true(pred, true):
    """confusion image: prediction in the red channel, ground truth in the green one
    (images.py:284-294)"""
    stacked = np.zeros(pred.shape + (3,), dtype=np.uint8)
    stacked[..., 0] = img_float_to_uint8(pred)
    stacked[..., 1] = img_float_to_uint8(true)
    return stacked


def overlapp_error(pred, true):
    """white where thresholded prediction and ground truth agree, black where they differ
    (images.py:297-310; 'agree' = both zero or both non-zero after the uint8 conversion)"""
    agree = (img_float_to_uint8(true) != 0) == (img_float_to_uint8(pred) != 0)
    return np.repeat((agree * np.uint8(PIXEL_DEPTH))[..., None], 3, axis=-1)


def _grey_bytes(img):
    """matplotlib 2.1's imsave of a 2-D array with cmap='gray' (requirements.txt:6; matplotlib is
    absent here, so this follows its published Normalize + Colormap.__call__(bytes=True)): min/max
    normalisation, index int(v * 256) into a 256-entry linear table whose byte form is
    (linspace(0, 1, 256) * 255) truncated."""
    a = np.asarray(img, dtype=np.float64)
    lo, hi = float(a.min()), float(a.max())
    norm = np.zeros_like(a) if hi == lo else (a - lo) / (hi - lo)
    table = (np.linspace(0.0, 1.0, 256) * 255).astype(np.uint8)
    return table[np.minimum((norm * 256).astype(np.int64), 255)]


def save_all(images, directory, format_="images_{:03d}.png", greyscale=False):
    """Save the `images` in the `directory` as RGBA PNG files numbered from 1
    (images.py:183-203).  2-D images need greyscale=True (matplotlib's default colour map is not
    reproduced); float RGB(A) images must lie in [0, 1]."""
    from PIL import Image
    os.makedirs(directory, exist_ok=True)
    images = np.asarray(images)
    if images.ndim == 4 and images.shape[-1] == 1:
        images = images.squeeze(-1)
    for n in range(images.shape[0]):
        img = images[n]
        if img.ndim == 2:
            if not greyscale:
                raise NotImplementedError("save_all: 2-D images are only written with greyscale=True")
            g = _grey_bytes(img)
            rgba = np.stack([g, g, g, np.full_like(g, 255)], axis=-1)
        else:
            if img.dtype != np.uint8:
                if img.min() < 0 or img.max() > 1:
                    raise ValueError("Floating point image RGB values must be in the 0..1 range.")
                img = (img * 255).astype(np.uint8)
            rgba = img if img.shape[-1] == 4 else np.concatenate(
                [img, np.full(img.shape[:2] + (1,), 255, np.uint8)], axis=-1)
        Image.fromarray(np.ascontiguousarray(rgba), "RGBA").save(os.path.join(directory, format_.format(n + 1)))
