"""Plumbing for the multi-GPU grid entry points (include/avif_b200.h: avifB200Grid*): builds the NCCL communicator the C library
uses, by broadcasting the library-generated unique id over an existing torch.distributed group (any backend).  No pixel work here."""
import ctypes


def make_comm(lib, dist, rank, world, device=None):
    """Returns a c_void_p holding an ncclComm_t created by libavif_b200 (collective over all ranks), or None for world == 1.
    The caller's current CUDA device must be the rank's GPU."""
    if world <= 1:
        return None
    import torch
    ident = (ctypes.c_uint8 * 128)()
    if rank == 0:
        res = lib.avifB200GridGetUniqueId(ctypes.byref(ident))
        if res != 0:
            raise RuntimeError(f"avifB200GridGetUniqueId failed ({res}): NCCL not loadable")
    backend = dist.get_backend()
    t = torch.tensor(list(ident), dtype=torch.uint8, device=(device if backend == "nccl" else "cpu"))
    dist.broadcast(t, src=0)
    ident = (ctypes.c_uint8 * 128)(*t.cpu().tolist())
    comm = ctypes.c_void_p()
    res = lib.avifB200GridCommInit(ctypes.byref(comm), world, rank, ctypes.byref(ident))
    if res != 0:
        raise RuntimeError(f"avifB200GridCommInit failed ({res})")
    return comm


def destroy_comm(lib, comm):
    if comm is not None:
        lib.avifB200GridCommDestroy(comm)
