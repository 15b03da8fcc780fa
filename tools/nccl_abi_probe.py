"""zutis_allreduce_hist over a raw ncclComm_t created with NCCL's own C API (not torch.distributed's communicator):
run under torchrun on >= 2 GPUs.
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/nccl_abi_probe.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import zutis_b200

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
torch.zeros(1, device="cuda")                                     # CUDA context
dist.init_process_group("gloo")                                   # only to hand the unique id around
nccl = ctypes.CDLL("libnccl.so.2")                                # the NCCL torch has already loaded (same soname)

class UniqueId(ctypes.Structure):
    _fields_ = [("internal", ctypes.c_char * 128)]

uid = UniqueId()
if rank == 0:
    assert nccl.ncclGetUniqueId(ctypes.byref(uid)) == 0
blob = [ctypes.string_at(ctypes.byref(uid), 128)]
dist.broadcast_object_list(blob, src=0)
ctypes.memmove(ctypes.byref(uid), blob[0], 128)
comm = ctypes.c_void_p()
nccl.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, UniqueId, ctypes.c_int]
assert nccl.ncclCommInitRank(ctypes.byref(comm), world, uid, rank) == 0

Q = 81
meter = zutis_b200.RunningScore(Q, device="cuda")
gen = torch.Generator().manual_seed(100 + rank)
gt = torch.randint(0, Q, (2, 64, 64), generator=gen).cuda(); pred = torch.randint(0, Q, (2, 64, 64), generator=gen).cuda()
meter.update(gt, pred)
mine = meter.counts().clone()
meter.all_reduce(nccl_comm=comm.value)
torch.cuda.synchronize()
total = meter.counts().cpu()
parts = [None] * world
dist.all_gather_object(parts, mine.cpu())
want = sum(parts)
ok = bool(torch.equal(total, want))
print(f"rank {rank}: zutis_allreduce_hist over a raw communicator equals the sum of the per-rank matrices: {ok} (sum {int(total.sum())})", flush=True)
nccl.ncclCommDestroy.argtypes = [ctypes.c_void_p]
nccl.ncclCommDestroy(comm)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
