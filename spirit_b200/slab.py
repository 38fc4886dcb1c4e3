"""Slab decomposition helpers (one process per GPU, torch.distributed for the rendezvous only).

The lattice is cut along c, the slowest index, so every slab is one contiguous range of the reference's site order:
rank r owns the planes [c_begin, c_begin + nc_local). The data path has exactly one exchange step per solver stage
(the first / last plane to the neighbouring ranks), done inside the library over NCCL (include/spirit_b200.h (3))."""
import ctypes


def partition(nc, world):
    """-> [(c_begin, nc_local)] * world, as even as possible, lower ranks get the remainder"""
    base, rem = divmod(nc, world)
    out, c = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((c, n))
        c += n
    return out


def broadcast_unique_id(make_id, dist, rank):
    """rank 0 creates the 128-byte NCCL id, everybody receives it (any torch.distributed backend)"""
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    assert isinstance(box[0], (bytes, bytearray)) and len(box[0]) == 128
    return bytes(box[0])


def init_comm(lib, dist, rank, world):
    """Create the library's NCCL communicator. `dist` must be an initialised torch.distributed."""
    def make_id():
        buf = ctypes.create_string_buffer(128)
        if lib.SpiritB200_Comm_Unique_Id(buf) != 0:
            raise RuntimeError("SpiritB200_Comm_Unique_Id failed")
        return buf.raw
    uid = broadcast_unique_id(make_id, dist, rank)
    if lib.SpiritB200_Comm_Init(rank, world, uid) != 0:
        raise RuntimeError("SpiritB200_Comm_Init failed")
