"""The five constants of the reference (src/constants.py:1-5).  The names and values are part of
the API surface (scripts import them by name), so they are the reference's; each one says where
the path uses it."""

# images in [0, 1] <-> 8-bit PNG levels (images.img_float_to_uint8, images.py:19-21)
PIXEL_DEPTH = 255

# RGB aerial images; two classes (background, road) out of the 1x1 head (unet.py:95)
NUM_CHANNELS, NUM_LABELS = 3, 2

# submission rule: a 16 x 16 cell is road when more than a quarter of it is
# (images.labels_for_patches / quantize_mask, images.py:88-99, 256-266)
IMG_PATCH_SIZE = 16
FOREGROUND_THRESHOLD = 0.25
