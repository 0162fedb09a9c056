"""Per-stage timing of update_cell_halo! on 2 GPUs (developer tool):
    torchrun --nproc-per-node 2 tools/time_halo.py [--cells 256]
For each split dimension (x, y, z) in turn: pack / NCCL send+recv / unpack of one face, CUDA-event timed."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import justpic.jl_b200 as J
from justpic.jl_b200 import halo as H
from bench import local_grids

ap = argparse.ArgumentParser(); ap.add_argument("--cells", type=int, default=256); a = ap.parse_args()
rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
comm = H.create_comm(device=dev)          # jp_halo_exchange's communicator
n = a.cells
for dim in range(3):
    dims = [1, 1, 1]; dims[dim] = 2
    topo = H.CartesianTopology(tuple(dims), rank)
    p = J.init_particles(J.CUDABackend, 24, 48, 12, *local_grids(n, topo.dims, topo.coords()), seed=42 + rank, device=dev)
    fields = J.init_cell_arrays(p, 3)
    arrays = tuple(p.coords) + tuple(fields)
    nb = H.plane_bytes(p.ncells, p.max_xcell, dim, len(arrays))
    sb, rb = torch.empty(nb, dtype=torch.uint8, device=dev), torch.empty(nb, dtype=torch.uint8, device=dev)
    other = 1 - rank
    plane_s, plane_r = (n - 2, n - 1) if rank == 0 else (1, 0)
    rows = []
    for it in range(6):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        ev[0].record(); H._cuda_pack(p, dim, plane_s, arrays, sb)
        ev[1].record()
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, sb, other), dist.P2POp(dist.irecv, rb, other)]): w.wait()
        ev[2].record(); H._cuda_unpack(p, dim, plane_r, arrays, rb)
        ev[3].record(); H.update_cell_halo(p, fields, topo, comm=comm)          # one jp_halo_exchange call (C: pack, NCCL group, unpack)
        ev[4].record(); torch.cuda.synchronize()
        rows.append([ev[i].elapsed_time(ev[i + 1]) for i in range(4)])
    r = np.array(rows)[2:].mean(axis=0)
    if rank == 0:
        print(f"dim {dim}: face {nb / 1e6:.1f} MB  pack {r[0]:.3f}  nccl {r[1]:.3f}  unpack {r[2]:.3f}  | jp_halo_exchange (pack + ncclSend/Recv + unpack) {r[3]:.3f} ms", flush=True)
    del p, fields, arrays, sb, rb
    torch.cuda.empty_cache()
torch.cuda.synchronize(); comm.destroy()
dist.destroy_process_group()
