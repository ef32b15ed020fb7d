"""Training steps per second at small batch sizes (the README recipe trains at batch 1), eager
launches against CUDA-graph replay (ConvolutionalModel._step_graphs), through train_batch with
pinned host batches in and loss + probabilities out.  Usage: python tools/time_small_batch.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from road_segmentation_unet_b200 import tf_aerial_images as tfa  # noqa: E402


def run(batch, graphs, steps=40):
    opts = tfa.Options()
    opts.num_layers, opts.root_size, opts.dilated_layers = 6, 64, True
    opts.patch_size, opts.batch_size, opts.dropout = 388, batch, 1.0
    opts.cuda_graphs = graphs
    model = tfa.ConvolutionalModel(opts, None)
    S, P = model.input_size, opts.patch_size
    rs = np.random.RandomState(0)
    x = torch.from_numpy(rs.rand(batch, S, S, 3).astype(np.float32)).pin_memory()
    y = torch.from_numpy((rs.rand(batch, P, P) > 0.7).astype(np.uint8)).pin_memory()
    for _ in range(5):
        model.train_batch(x, y)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        model.train_batch(x, y)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    del model
    torch.cuda.empty_cache()
    return dt


if __name__ == "__main__":
    for batch in (1, 2, 4, 8):
        e, g = run(batch, "0"), run(batch, "1")
        print("batch %d: eager %.2f ms/step (%.0f patches/s), graphs %.2f ms/step (%.0f patches/s)"
              % (batch, e * 1e3, batch / e, g * 1e3, batch / g))
