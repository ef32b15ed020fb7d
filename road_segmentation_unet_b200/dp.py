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
        # RSU_DP_MULTICAST: 0 = peer loads / stores only, 1 = NVLS for both directions
        # (multimem.ld_reduce for the gradients, multimem.st for the weights), 2 / 3 = loads / stores
        # only; default "auto" = 1 from four ranks on (profiles/r2_dp_n2.txt / r2_dp_n8.txt: equal at
        # 2 GPUs, ~0.2 ms per step better at 8 -- the switch carries 1/world of the traffic per rank)
        mc_env = os.environ.get("RSU_DP_MULTICAST", "auto")
        mc_mode = (1 if world >= 4 else 0) if mc_env == "auto" else int(mc_env)
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
        # RSU_DP_OVERLAP: 1 = buckets are exchanged on a side stream while the backward pass runs,
        # 0 = one exchange after the whole backward pass.  Default "auto": overlapped from four ranks
        # on.  At two ranks the exchange after the backward pass already costs only 0.1 ms (each
        # rank updates half of the parameters instead of all of them), and that is the combination
        # the full 2-GPU benchmark has run with: the one 2-GPU bench.py run with the overlapped
        # peer-load variant did not finish (cause not established before the round's GPU budget
        # ended; tools/test_dp.py passes with it at 2 and at 8 GPUs, bench.py at 4 and at 8).
        ov_env = os.environ.get("RSU_DP_OVERLAP", "auto")
        self.overlap = (world >= 4) if ov_env == "auto" else ov_env != "0"
        # a rank that never arrives must not hang its peers for ever: the barrier kernels trap
        # after this long (milliseconds; 10 minutes, the default of NCCL's own watchdog: a rank may be
        # busy on the host -- evaluation dumps, a checkpoint -- while its peers already wait)
        self.barrier_timeout_ms = int(os.environ.get("RSU_DP_BARRIER_TIMEOUT_MS", "600000"))
        self._side = torch.cuda.Stream()
        self._armed, self._done, self._acc = None, [], None

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

    # ---- overlapped exchange: buckets are exchanged on a side stream while the backward pass runs
    def arm(self, lr, momentum, extra_scale=1.0):
        """Called before a backward pass: from now on every finished bucket of the flat gradient
        (UNet.on_bucket_ready) is exchanged right away on the side stream -- the kernel needs no
        shared memory and few registers, so its blocks co-reside with the persistent tensor-core
        kernels of the layers that are still being differentiated."""
        self._armed = (float(lr), float(momentum), float(extra_scale))
        self._done = []

    def bucket_ready(self, lo, hi):
        if getattr(self, "_armed", None) is None or not self.overlap:
            return
        lr, momentum, scale = self._armed
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        with torch.cuda.stream(self._side):
            self._side.wait_event(ev)
            self._hg.barrier(channel=0, timeout_ms=self.barrier_timeout_ms)  # all ranks have this bucket
            self._launch([(lo, hi)], lr, momentum, scale)
        self._done.append((lo, hi))

    def _launch(self, ranges, lr, momentum, scale):
        lib = _lib.load()
        stream = _lib.stream_ptr()
        for lo, hi in rank_slices(ranges, self.rank, self.world):
            _lib.check(lib.rsu_dp_momentum_sgd(C.byref(self._peers), C.c_void_p(self._acc.data_ptr()),
                                               lo, hi, lr, momentum, scale / self.world, stream))

    def buckets(self, net):
        """The exchange units: the live pieces of the backward pass's gradient buckets, in flat
        order.  Ownership (which rank updates which elements, and holds their momentum) is always
        the per-bucket split of rank_slices, whether a bucket is exchanged during the backward pass
        or afterwards."""
        if getattr(self, "_buckets", None) is None:
            enc, dec = net._bucket_bounds()
            out = []
            for lo, hi in enc + dec:
                for a, b in net.live_ranges():
                    if min(b, hi) > max(a, lo):
                        out.append((max(a, lo), min(b, hi)))
            self._buckets = sorted(out)
        return self._buckets

    def owned(self, net):
        return [s for b in self.buckets(net) for s in rank_slices([b], self.rank, self.world)]

    def step(self, net, lr, momentum, extra_scale=1.0):
        """All ranks' gradients -> one update of this rank's slices -> everybody's weights: the
        buckets that were not already exchanged during the backward pass, then the barrier after
        which every replica holds the new weights."""
        self._acc = net.momentum
        done = set(getattr(self, "_done", []))
        rest = [b for b in self.buckets(net) if b not in done]
        assert all(d in self.buckets(net) for d in done)
        cur = torch.cuda.current_stream()
        if done:
            cur.wait_stream(self._side)
        if rest:
            self._hg.barrier(channel=0, timeout_ms=self.barrier_timeout_ms)  # every backward pass finished
            for b in rest:
                self._launch([b], float(lr), float(momentum), float(extra_scale))
        self._hp.barrier(channel=1, timeout_ms=self.barrier_timeout_ms)  # all weights (and gradient reads) done
        self._armed, self._done = None, []

    def full_momentum(self, net):
        """The complete momentum vector (every rank holds its own slices): one all-reduce of the
        slices into a scratch vector, for checkpoints."""
        full = torch.zeros_like(net.momentum)
        for lo, hi in self.owned(net):
            full[lo:hi].copy_(net.momentum[lo:hi])
        self.dist.all_reduce(full)
        return full
