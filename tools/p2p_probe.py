"""Probe: which peer-memory mechanism works on this box (torchrun --nproc-per-node 2 tools/p2p_probe.py)."""
import os
import traceback

import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
print(rank, "can_access_peer", [torch.cuda.can_device_access_peer(lr, j) for j in range(world) if j != lr], flush=True)

# 1) symmetric memory
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(1024, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    t.fill_(float(rank + 1))
    dist.barrier(); torch.cuda.synchronize()
    peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.float32)
    print(rank, "symm_mem ok: peer value", float(peer[0]), "ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal pads", len(hdl.signal_pad_ptrs), flush=True)
    dist.barrier()
except Exception:
    print(rank, "symm_mem FAILED"); traceback.print_exc()

# 2) classic CUDA IPC through torch storages
try:
    x = torch.full((1024,), float(rank + 10), device=dev)
    h = x.untyped_storage()._share_cuda_()
    hs = [None] * world
    dist.all_gather_object(hs, h)
    j = (rank + 1) % world
    st = torch.UntypedStorage._new_shared_cuda(*hs[j])
    y = torch.empty(0, dtype=torch.float32, device=st.device).set_(st, 0, (1024,))
    torch.cuda.synchronize(); dist.barrier()
    print(rank, "cuda ipc ok: peer value", float(y[0].item()), "peer device", y.device, "ptr", hex(y.data_ptr()), flush=True)
    # write into the peer's buffer from this GPU
    z = torch.full((1024,), float(100 + rank), device=dev)
    y.copy_(z)
    torch.cuda.synchronize(); dist.barrier()
    print(rank, "after peer write my buffer holds", float(x[0]), flush=True)
except Exception:
    print(rank, "cuda ipc FAILED"); traceback.print_exc()
dist.barrier()
dist.destroy_process_group()
