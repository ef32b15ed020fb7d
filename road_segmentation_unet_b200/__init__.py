"""B200-native U-Net training / sliding-window prediction hot path of
aschneuw/road-segmentation-unet.  Host modules mirror the reference's src/ modules
(unet, images, tf_aerial_images, constants); the arithmetic lives in librsu_b200.so
(hand-written sm_100a CUDA behind the C ABI of include/rsu_b200.h)."""
__version__ = "0.1.0"
