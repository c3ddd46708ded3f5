"""axb_comm: the library's own NCCL communicator (include/axb200.h, csrc/comm.cuh), one per rank / process / GPU.

The collectives of the distributed cases run INSIDE libaxb200.so on the handle's stream; Python only gets the 128-byte
NCCL id from rank 0 to the other ranks.  Any transport will do: Comm.from_torch() uses an initialised torch.distributed
group (gloo or nccl) for that one broadcast, Comm.from_file() a file both ranks can see (no launcher at all)."""
import ctypes as C
import os
import time

from . import _lib
from ._lib import check

ID_BYTES = 128


def _ensure_nccl_visible():
    """the library binds NCCL with dlopen: prefer the copy torch ships, already mapped once torch.cuda.nccl is touched"""
    if os.environ.get("AXB_NCCL_LIB"):
        return
    try:
        import torch
        sp = os.path.dirname(os.path.dirname(torch.__file__))
        cand = os.path.join(sp, "nvidia", "nccl", "lib", "libnccl.so.2")
        if os.path.exists(cand):
            os.environ["AXB_NCCL_LIB"] = cand
    except Exception:
        pass


def unique_id():
    _ensure_nccl_visible()
    buf = (C.c_uint8 * ID_BYTES)()
    check(_lib.lib().axb_comm_get_unique_id(buf))
    return bytes(buf)


class Comm:
    def __init__(self, nranks, rank, id_bytes, device):
        _ensure_nccl_visible()
        self._L = _lib.lib()
        self._h = None
        assert len(id_bytes) == ID_BYTES
        h = C.c_void_p()
        buf = (C.c_uint8 * ID_BYTES).from_buffer_copy(id_bytes)
        check(self._L.axb_comm_create(C.byref(h), int(nranks), int(rank), buf, int(device)))
        self._h, self.nranks, self.rank, self.device = h, int(nranks), int(rank), int(device)

    @classmethod
    def from_torch(cls, device, group=None):
        """collective over an initialised torch.distributed group: rank 0's id is broadcast as an object"""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        return cls(world, rank, box[0], device)

    @classmethod
    def from_file(cls, path, nranks, rank, device, timeout_s=120.0):
        """rank 0 writes the id to `path` (atomically), the others wait for it"""
        if rank == 0:
            tmp = path + ".tmp"
            with open(tmp, "wb") as f:
                f.write(unique_id())
            os.replace(tmp, path)
        t0 = time.time()
        while not os.path.exists(path):
            if time.time() - t0 > timeout_s:
                raise TimeoutError("no NCCL id at %s" % path)
            time.sleep(0.01)
        with open(path, "rb") as f:
            idb = f.read()
        return cls(nranks, rank, idb, device)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._L.axb_comm_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def traffic(self):
        """(payload bytes this rank contributed to collectives, collectives issued) since creation"""
        b, n = C.c_int64(), C.c_int64()
        check(self._L.axb_comm_get_traffic(self._h, C.byref(b), C.byref(n)))
        return b.value, n.value

    def library(self):
        return self._L.axb_comm_library().decode()

    def allreduce_(self, t, op="min"):
        """in-place elementwise reduction of a float64 CUDA tensor on torch's current stream"""
        import torch
        assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
        code = {"min": 0, "max": 1, "sum": 2}[op]
        check(self._L.axb_comm_allreduce_f64(self._h, t.data_ptr(), t.numel(), code, C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)))
        return t
