"""Data-parallel plumbing of the training step (one process per GPU).

Two ways to turn per-rank gradients into one identical update on every replica:

  * `PeerOptimizer` (default on NVLink-connected GPUs): the flat gradient and parameter vectors of
    every rank live in symmetric memory (torch.distributed._symmetric_memory: cudaMalloc'd buffers
    exchanged between the processes, plus an NVLS multicast mapping when the fabric has one), and
    ONE kernel per rank -- `rsu_dp_momentum_sgd` -- sums the peers' gradients for the rank's own
    1/world slice, applies the momentum update there and stores the new weights into every peer's
    parameter vector: reduce-scatter + ApplyMomentum + all-gather in the update's own memory pass.
    Optimizer state is sharded (each rank keeps the momentum of its slice).  Two stream-ordered
    barriers on the symmetric-memory signal pads order the step across ranks.
  * `GradientAllReducer` (tf_aerial_images.py, fallback and CPU/gloo tests): bucketed NCCL
    all-reduce of the flat gradient overlapped with the backward pass, then the local update.

torch.distributed is plumbing here (process group, rendezvous of the buffers); the arithmetic and
the data movement are the library's own kernel.
"""
import ctypes as C
import os

import torch

from . import _lib


def rank_slices(live_ranges, rank, world, align=4):
    """The pieces of the flat vector owned by `rank`: every live range [a, b) is cut into `world`
    contiguous parts at multiples of `align` elements (16 bytes: the kernel's access width)."""
    out = []
    for a, b in live_ranges:
        assert a % align == 0 and b % align == 0
        units = (b - a) // align
        lo = a + units * rank // world * align
        hi = a + units * (rank + 1) // world * align
        if hi > lo:
            out.append((lo, hi))
    return out


class PeerOptimizer:
    """Symmetric-memory gradient / parameter buffers of one data-parallel replica and the fused
    reduce + momentum + broadcast step over them."""

    def __init__(self, dist_module, world, rank, n_flat, device="cuda"):
        import torch.distributed._symmetric_memory as symm
        self.dist, self.world, self.rank = dist_module, world, rank
        group = dist_module.group.WORLD
        name = group.group_name
        self.params = symm.empty(n_flat, dtype=torch.float32, device=device)
        self.grads = symm.empty(n_flat, dtype=torch.float32, device=device)
        self.params.zero_()
        self.grads.zero_()
        self._hp = symm.rendezvous(self.params, name)
        self._hg = symm.rendezvous(self.grads, name)
        # RSU_DP_MULTICAST: 0 = peer loads / stores only (default: as fast as NVLS at 2 and at 8
        # GPUs, profiles/r2_dp_n2.txt / r2_dp_n8.txt -- every rank has to receive (world-1)/world
        # of the fp32 weights either way), 1 = NVLS for both directions, 2 = multimem.ld_reduce for
        # the gradients only, 3 = multimem.st for the weights only
        mc_mode = int(os.environ.get("RSU_DP_MULTICAST", "0"))
        want_mc = mc_mode != 0
        peers = _lib.DpPeers()
        peers.world, peers.rank = world, rank
        for r in range(world):
            peers.grads[r] = self._peer_ptr(self._hg, self.grads, r)
            peers.params[r] = self._peer_ptr(self._hp, self.params, r)
        assert peers.grads[rank] == self.grads.data_ptr() and peers.params[rank] == self.params.data_ptr()
        mc_g, mc_p = self._mc_ptr(self._hg, self.grads), self._mc_ptr(self._hp, self.params)
        self.multicast = bool(want_mc and mc_g and mc_p)
        peers.grads_mc = mc_g if self.multicast and mc_mode in (1, 2) else None
        peers.params_mc = mc_p if self.multicast and mc_mode in (1, 3) else None
        self.mc_mode = mc_mode if self.multicast else 0
        self._peers = peers

    @staticmethod
    def _peer_ptr(handle, tensor, r):
        off = tensor.data_ptr() - int(handle.buffer_ptrs[handle.rank])
        return int(handle.buffer_ptrs[r]) + off

    @staticmethod
    def _mc_ptr(handle, tensor):
        try:
            if not handle.has_multicast_support(torch.device("cuda").type, tensor.device.index):
                return 0
        except Exception:
            pass
        base = int(getattr(handle, "multicast_ptr", 0) or 0)
        if not base:
            return 0
        return base + (tensor.data_ptr() - int(handle.buffer_ptrs[handle.rank]))

    def step(self, net, lr, momentum, extra_scale=1.0):
        """All ranks' gradients -> one update of this rank's slices -> everybody's weights.
        extra_scale: additional factor on the summed gradient (1 / micro-batches)."""
        self._hg.barrier(channel=0)          # every rank's backward pass has finished
        lib = _lib.load()
        stream = _lib.stream_ptr()
        for lo, hi in rank_slices(net.live_ranges(), self.rank, self.world):
            _lib.check(lib.rsu_dp_momentum_sgd(C.byref(self._peers), C.c_void_p(net.momentum.data_ptr()),
                                               lo, hi, float(lr), float(momentum), float(extra_scale) / self.world, stream))
        self._hp.barrier(channel=1)          # every rank's weights (and gradient reads) are complete

    def full_momentum(self, net):
        """The complete momentum vector (every rank holds its own slices): one all-reduce of the
        slices into a scratch vector, for checkpoints."""
        full = torch.zeros_like(net.momentum)
        for lo, hi in rank_slices(net.live_ranges(), self.rank, self.world):
            full[lo:hi].copy_(net.momentum[lo:hi])
        self.dist.all_reduce(full)
        return full
