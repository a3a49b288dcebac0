"""Probe: does CUDA IPC (cudaIpcGetMemHandle / cudaIpcOpenMemHandle) + peer stores work between two ranks here?
   torchrun --nproc-per-node 2 tools/ipc_probe.py"""
import ctypes, os, torch, torch.distributed as dist
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank)
dist.init_process_group('gloo')
rt = ctypes.CDLL('libcudart.so.12')
ptr = ctypes.c_void_p()
assert rt.cudaMalloc(ctypes.byref(ptr), 1 << 20) == 0
class Handle(ctypes.Structure):
    _fields_ = [('r', ctypes.c_ubyte * 64)]
rt.cudaIpcOpenMemHandle.argtypes = [ctypes.POINTER(ctypes.c_void_p), Handle, ctypes.c_uint]
handle = Handle()
rc = rt.cudaIpcGetMemHandle(ctypes.byref(handle), ptr)
print(rank, 'get handle rc', rc)
handles = [None] * world
dist.all_gather_object(handles, bytes(handle))
peer = ctypes.c_void_p()
other = (rank + 1) % world
h = Handle.from_buffer_copy(handles[other])
rc = rt.cudaIpcOpenMemHandle(ctypes.byref(peer), h, 1)
print(rank, 'open rc', rc, hex(peer.value or 0))
if rc == 0:
    val = (ctypes.c_int * 4)(rank + 100, 1, 2, 3)
    rc = rt.cudaMemcpy(peer, val, 16, 1)   # H2D into the PEER's buffer
    print(rank, 'copy into peer rc', rc)
    dist.barrier()
    got = (ctypes.c_int * 4)()
    rt.cudaMemcpy(got, ptr, 16, 2)
    print(rank, 'my buffer now holds', list(got))
dist.barrier()
