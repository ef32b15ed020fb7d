import os, sys, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank = int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
name = dist.group.WORLD.group_name
hs = []
for i, n in enumerate((1000, 1 << 20, 212403278, 212403278)):
    t = symm.empty(n, dtype=torch.float32, device="cuda")
    h = symm.rendezvous(t, name)
    hs.append((t, h))
    print(rank, "alloc", i, "n", n, "data_ptr %x" % t.data_ptr(), "buffer_ptrs", ["%x" % p for p in h.buffer_ptrs],
          "offset", getattr(h, "offset", None), "buffer_size", h.buffer_size, "mc %x" % (h.multicast_ptr or 0),
          "signal_pads", ["%x" % p for p in h.signal_pad_ptrs], flush=True)
# write a pattern through the multicast address of alloc 1 with torch ops? check get_buffer views
t, h = hs[1]
t.fill_(rank + 1)
h.barrier(channel=0)
other = h.get_buffer(1 - rank, (8,), torch.float32)
print(rank, "peer view first values", other.tolist(), flush=True)
h.barrier(channel=0)
dist.destroy_process_group()
