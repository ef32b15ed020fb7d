"""Constants of the reference (src/constants.py:1-5), values kept verbatim."""
FOREGROUND_THRESHOLD = .25
IMG_PATCH_SIZE = 16
NUM_CHANNELS = 3
NUM_LABELS = 2
PIXEL_DEPTH = 255
