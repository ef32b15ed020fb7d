"""CPU oracle for the NumPy/SciPy half of the hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module; the product path never does.

NumPy restatement of the geometry helpers of the reference's src/images.py, each function citing
the lines it follows.  PINNED: tests/test_oracle_cpu.py checks every function against golden
vectors produced by importing the reference's own, unmodified src/images.py in the build
container (tests/golden/make_golden.py -> tests/golden/images_golden.npz).  The rotation is the
one function whose arithmetic lives in a third-party dependency (scipy.ndimage.rotate, pinned
scipy==1.0.0 in requirements.txt:15); `rotate_nn` restates SciPy's published algorithm and is
pinned against the SciPy installed here.
"""
import numpy as np


def mirror_border(images, n):
    """images.py:269-281 -- np.pad(..., 'symmetric') on H and W, written as an index gather."""
    h, w = images.shape[1], images.shape[2]

    def sym(idx, length):
        m = np.mod(idx, 2 * length)
        return np.where(m < length, m, 2 * length - 1 - m)

    iy = sym(np.arange(-n, h + n), h)
    ix = sym(np.arange(-n, w + n), w)
    return images[:, iy][:, :, ix]


def extract_patches(images, patch_size, stride=None, predict_patch_size=None):
    """images.py:35-85 -- x (column) is the OUTER loop, y the inner; float64 output."""
    if not predict_patch_size:
        predict_patch_size = patch_size
    assert (patch_size - predict_patch_size) % 2 == 0 and predict_patch_size <= patch_size
    if not stride:
        stride = patch_size
    num_images, h, w = images.shape[:3]
    assert h == w, "Assume square images"
    assert (h - patch_size) % stride == 0, "Stride sliding should cover the whole image"
    side = (h - patch_size) // stride + 1
    out = np.zeros((num_images * side * side, patch_size, patch_size) + images.shape[3:])
    k = 0
    for n in range(num_images):
        for x in range(0, w - patch_size + 1, stride):
            for y in range(0, h - patch_size + 1, stride):
                out[k] = images[n, y:y + patch_size, x:x + patch_size]
                k += 1
    return out


def images_from_patches(patches, stride=None):
    """images.py:131-164 -- accumulate + hit count in the same x-outer order, then divide."""
    num_images, num_patches, p, _, c = patches.shape
    if stride is None:
        stride = p
    side = int(np.sqrt(num_patches))
    assert np.sqrt(num_patches) == side, "Square image assumption broken"
    size = (side - 1) * stride + p
    images = np.zeros((num_images, size, size, c), dtype=patches.dtype)
    hits = np.zeros((num_images, size, size, c), dtype=np.uint64)
    for n in range(num_images):
        k = 0
        for x in range(0, size - p + 1, stride):
            for y in range(0, size - p + 1, stride):
                images[n, y:y + p, x:x + p] += patches[n, k]
                hits[n, y:y + p, x:x + p] += 1
                k += 1
    return images / hits


def rotate_params(in_side, cos_a, sin_a):
    """scipy.ndimage.rotate(reshape=True) geometry: matrix, offset and output side."""
    rot = np.array([[cos_a, sin_a], [-sin_a, cos_a]])
    iy = ix = in_side
    out_bounds = rot @ np.array([[0, 0, iy, iy], [0, ix, 0, ix]], dtype=np.float64)
    out_shape = (np.ptp(out_bounds, axis=1) + 0.5).astype(int)
    out_center = rot @ ((out_shape - 1) / 2)
    in_center = (np.array([iy, ix]) - 1) / 2
    return rot, in_center - out_center, out_shape


def cosdg_sindg(angle):
    """SciPy evaluates the rotation matrix with special.cosdg / sindg (exact at multiples of 90)."""
    from scipy import special
    return float(special.cosdg(angle)), float(special.sindg(angle))


def rotate_nn(imgs, angle):
    """images.py:313-317 -- scipy.ndimage.rotate(imgs, angle, axes=(1, 2), order=0) with the
    defaults reshape=True, mode='constant', cval=0.  Nearest neighbour: the input coordinate
    c = R*o + offset is rounded with floor(c + 0.5); samples whose (unrounded) coordinate lies
    outside [0, side-1] on either axis are 0 (mode='constant' is decided before rounding).
    Each (n, c) plane independently; dtype preserved."""
    if angle == 0:
        return imgs
    c, s = cosdg_sindg(angle)
    side = imgs.shape[1]
    rot, offset, out_shape = rotate_params(side, c, s)
    oy, ox = np.meshgrid(np.arange(out_shape[0], dtype=np.float64),
                         np.arange(out_shape[1], dtype=np.float64), indexing="ij")
    # SciPy's NI_GeometricTransform accumulates shift first, then one product per axis
    iy = (offset[0] + oy * rot[0, 0]) + ox * rot[0, 1]
    ix = (offset[1] + oy * rot[1, 0]) + ox * rot[1, 1]
    ry = np.floor(iy + 0.5).astype(np.int64)
    rx = np.floor(ix + 0.5).astype(np.int64)
    ok = (iy >= 0) & (iy <= side - 1) & (ix >= 0) & (ix <= side - 1)
    ryc, rxc = np.clip(ry, 0, side - 1), np.clip(rx, 0, side - 1)
    out = imgs[:, ryc, rxc]
    mask = ok.reshape((1,) + ok.shape + (1,) * (imgs.ndim - 3))
    return np.where(mask, out, np.zeros((), dtype=imgs.dtype)).astype(imgs.dtype)


def crop_imgs(imgs, crop_size):
    """images.py:354-373 -- centre crop [c-h : c+h], c = int(H/2), h = crop/2."""
    h, w = imgs.shape[1], imgs.shape[2]
    assert h == w and h >= crop_size
    assert crop_size % 2 == 0
    half, center = crop_size // 2, int(h / 2)
    return imgs[:, center - half:center + half, center - half:center + half]


def expand_and_rotate(imgs, angles, offset=0):
    """images.py:320-351 -- mirror-pad, then for each angle (angle-major output)
    crop_imgs(rotate_imgs(padded, angle), H + 2*offset); float64 output."""
    has_channels = imgs.ndim == 4
    if not has_channels:
        imgs = np.expand_dims(imgs, -1)
    b, h, w, c = imgs.shape
    assert h == w
    out_size = h + 2 * offset
    padding = int(np.ceil(h * (np.sqrt(2) - 1) / 2)) + int(np.ceil(offset / np.sqrt(2)))
    padded = mirror_border(imgs, padding)
    out = np.zeros((b * len(angles), out_size, out_size, c))
    for i, angle in enumerate(angles):
        out[i * b:(i + 1) * b] = crop_imgs(rotate_nn(padded, angle), out_size)
    if not has_channels:
        out = np.squeeze(out, -1)
    return out


def image_augmentation_ensemble(imgs):
    """images.py:376-396 -- [orig | flip axis 2 | flip axis 1 | rot90 k=1,2,3], variant-major."""
    n = imgs.shape[0]
    out = np.zeros((n * 6,) + imgs.shape[1:])
    out[:n] = imgs
    out[n:2 * n] = imgs[:, :, ::-1]
    out[2 * n:3 * n] = imgs[:, ::-1]
    for i, k in enumerate([1, 2, 3]):
        out[(3 + i) * n:(4 + i) * n] = np.rot90(imgs, k=k, axes=(1, 2))
    return out


def invert_image_augmentation_ensemble(masks):
    """images.py:399-417 -- undo the six variants and average (does not mutate its argument)."""
    assert masks.shape[0] % 6 == 0
    n = masks.shape[0] // 6
    result = masks[:n].copy()
    result += masks[n:2 * n][:, :, ::-1]
    result += masks[2 * n:3 * n][:, ::-1]
    for i, k in enumerate([-1, -2, -3]):
        result += np.rot90(masks[(3 + i) * n:(4 + i) * n], k=k, axes=(1, 2))
    return result / 6


def d4(x, op):
    """out = rot90(flipud(x) if op & 4 else x, k = op & 3) on the (H, W) axes of one image."""
    if op & 4:
        x = x[::-1]
    return np.rot90(x, k=op & 3, axes=(0, 1))


def quantize_mask(masks, threshold, patch_size):
    """images.py:256-266 -- 16x16 vote: mean(prob >= 0.5) > threshold."""
    n, size = masks.shape[0], masks.shape[1]
    out = masks.copy()
    for i in range(n):
        for y in range(0, size, patch_size):
            for x in range(0, size, patch_size):
                label = (masks[i, y:y + patch_size, x:x + patch_size, 0] >= 0.5).mean() > threshold
                out[i, y:y + patch_size, x:x + patch_size, 0] = label
    return out


def patch_f1(pred_masks, true_masks, patch_size=16, threshold=0.25):
    """Patch-level F1 with the quantize_mask rule (images.py:256-266) and the F1 definition of
    summary.py:141-147: 2 / (1/recall + 1/precision)."""
    def labels(m):
        n, s = m.shape[0], m.shape[1]
        g = s // patch_size
        v = (m[:, :g * patch_size, :g * patch_size] >= 0.5).reshape(n, g, patch_size, g, patch_size)
        return v.mean(axis=(2, 4)) > threshold
    p, t = labels(np.asarray(pred_masks).squeeze(-1) if np.asarray(pred_masks).ndim == 4 else pred_masks), \
        labels(np.asarray(true_masks).squeeze(-1) if np.asarray(true_masks).ndim == 4 else true_masks)
    tp = float(np.sum(p & t))
    fp = float(np.sum(p & ~t))
    fn = float(np.sum(~p & t))
    if tp == 0:
        return 0.0
    recall, precision = tp / (tp + fn), tp / (tp + fp)
    return 2.0 / (1.0 / recall + 1.0 / precision)


def labels_for_patches(patches, threshold=0.25):
    """images.py:88-99 -- label 1 when the patch mean exceeds FOREGROUND_THRESHOLD."""
    return (np.asarray(patches, np.float64).mean(axis=(1, 2)) > threshold).astype(np.int64)


def submission_rows(masks, patch_size, threshold=0.25):
    """images.py:206-237 -- the text of submission.csv: header, then per image one row
    'NNN_x_y,label' per cell with x (column offset) outer and y inner; the label of cell (x, y)
    is labels_for_patches of mask[y:y+p, x:x+p] (extract_patches order, images.py:76-77)."""
    m = np.asarray(masks)
    if m.ndim == 4:
        m = m[..., 0]
    n, s = m.shape[0], m.shape[1]
    assert m.shape[2] == s and s % patch_size == 0
    rows = ["id,prediction"]
    for i in range(n):
        for x in range(0, s, patch_size):
            for y in range(0, s, patch_size):
                lab = int(np.float64(m[i, y:y + patch_size, x:x + patch_size]).mean() > threshold)
                rows.append("%03d_%d_%d,%d" % (i + 1, x, y, lab))
    return "\n".join(rows) + "\n"


def to_uint8(img):
    """images.py:19-21 -- round(img * 255) as uint8."""
    return np.round(np.asarray(img) * 255).astype(np.uint8)


def overlays(imgs, masks, fade=0.95):
    """images.py:102-128 -- red layer (255, 0, 0, trunc(uint8(mask) * fade)) composited over the
    opaque image with Pillow's integer 'over' operator (Image.alpha_composite): per channel
    out = (src * a + dst * (255 - a) + rounding) / 255 on 8-bit values, alpha stays 255."""
    base = to_uint8(imgs).astype(np.int64)
    a = (to_uint8(np.asarray(masks).reshape(base.shape[:3])) * fade).astype(np.uint8).astype(np.int64)
    out = np.empty(base.shape[:3] + (4,), np.uint8)
    src = np.array([255, 0, 0], np.int64)
    # Pillow's composite for an opaque destination: blend = a * 255 (outA = 255*255 scale),
    # coef1 = a * 255 * 255 / outA255, ... reduces to the rounded 8-bit lerp below
    for c in range(3):
        t = src[c] * a[..., None][..., 0] + base[..., c] * (255 - a) + 128
        out[..., c] = ((t + (t >> 8)) >> 8).astype(np.uint8)
    out[..., 3] = 255
    return out


def overlap_pred_true(pred, true):
    """images.py:284-294 -- R = uint8(pred), G = uint8(true), B = 0."""
    out = np.zeros(np.asarray(pred).shape + (3,), np.uint8)
    out[..., 0] = to_uint8(pred)
    out[..., 1] = to_uint8(true)
    return out


def overlapp_error(pred, true):
    """images.py:297-310 -- 255 on all channels where (uint8(pred) != 0) == (uint8(true) != 0)."""
    agree = (to_uint8(pred) != 0) == (to_uint8(true) != 0)
    return np.repeat((agree.astype(np.uint8) * 255)[..., None], 3, axis=-1)
