"""Peer-memory communicator for the sharded outer step (``xm_comm_*`` / ``xm_allreduce_adam`` in csrc/comm.cu).

One process per GPU.  Each rank allocates a staging / flag block through the C ABI, the 64-byte CUDA IPC handles are
exchanged with ``torch.distributed.all_gather_object`` (plumbing only -- no tensor ever goes through a collective),
and every rank maps its peers' blocks.  After that the meta-gradient exchange happens inside the Adam kernel, over
NVLink loads / stores."""
import ctypes

import torch.distributed as dist

from . import _lib


class PeerComm:
    def __init__(self, n_floats, device, group=None):
        self.lib = _lib.load()
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n_floats = int(n_floats)
        if self.world > _lib.XM_COMM_MAX_WORLD:
            raise _lib.XmetaError('PeerComm supports up to %d ranks on one NVLink domain' % _lib.XM_COMM_MAX_WORLD)
        handle = (ctypes.c_ubyte * _lib.XM_IPC_HANDLE_BYTES)()
        ptr = ctypes.c_void_p()
        _lib.check(self.lib.xm_comm_create(self.world, self.rank, self.n_floats, ctypes.byref(ptr),
                                           ctypes.cast(handle, ctypes.c_void_p)), 'xm_comm_create')
        self.ptr = ptr
        gathered = [None] * self.world
        dist.all_gather_object(gathered, bytes(handle), group=group)
        blob = b''.join(gathered)
        buf = (ctypes.c_ubyte * len(blob)).from_buffer_copy(blob)
        _lib.check(self.lib.xm_comm_connect(self.ptr, ctypes.cast(buf, ctypes.c_void_p)), 'xm_comm_connect')
        dist.barrier(group=group)          # every rank has mapped every block before the first kernel touches one

    def check(self):
        """Raises if a peer failed to arrive in an earlier ``xm_allreduce_adam`` (synchronises the device)."""
        code = self.lib.xm_comm_error(self.ptr)
        if code != 0:
            raise _lib.XmetaError('xm_allreduce_adam: a peer rank did not arrive within the time-out (code %d)' % code)

    def close(self):
        if self.ptr is not None:
            self.lib.xm_comm_destroy(self.ptr)
            self.ptr = None
