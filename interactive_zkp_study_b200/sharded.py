"""Multi-GPU MSM: one process per GPU, points sharded by contiguous range.

sum_i k_i P_i is a sum of independent per-range partial sums, so every rank runs the full single-GPU
Pippenger on its own range of the static table (resident on its GPU) and emits ONE un-normalised XYZZ
point.  The only exchange step of the path is an all-gather of world x 128 bytes; rank 0 folds the
partials and converts to affine.  There is no other collective: the NTT / quotient / PLONK rounds are
single-GPU work ("replicas only").

Process-group plumbing is torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests).
Under NCCL the partial is written by the library directly into the CUDA send buffer of the gather and
folded directly out of its receive buffer (no host staging); under gloo the same buffers live in host
memory.  The reference has no multi-device path; the call it scales is kzg.commit
(/root/reference/zkp/plonk/kzg.py:32-67) / the proof-element sums of proving.py:23-75.

Partial wire format (G1): x | y | zz | zzz, each 32 bytes little-endian, Montgomery form (x * 2^256 mod p);
the point is (x/zz, y/zzz); zz == 0 encodes the point at infinity.
"""

PARTIAL_BYTES = 128


def shard_range(total, rank, world):
    """Contiguous range [start, start + count) of `total` points owned by `rank` of `world`; the
    remainder goes to the first ranks one point each, so ranges tile [0, total) in rank order."""
    if world <= 0 or not 0 <= rank < world or total < 0:
        raise ValueError("shard_range: need 0 <= rank < world and total >= 0")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


class PartialExchange:
    """The gather of one 128-byte partial per rank.  Buffers are allocated once and reused by every
    MSM (a CUDA tensor pair under NCCL, host tensors under gloo)."""

    def __init__(self, group=None):
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("PartialExchange needs an initialised torch.distributed process group")
        self._torch, self._dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.on_device = "nccl" in str(dist.get_backend(group)).lower()
        device = torch.device("cuda", torch.cuda.current_device()) if self.on_device else torch.device("cpu")
        self.send = torch.zeros(PARTIAL_BYTES, dtype=torch.uint8, device=device)
        self.recv = torch.zeros(PARTIAL_BYTES * self.world, dtype=torch.uint8, device=device)

    @property
    def send_addr(self):
        """Where this rank's partial goes (device address under NCCL, host address under gloo)."""
        return self.send.data_ptr()

    @property
    def recv_addr(self):
        return self.recv.data_ptr()

    def write_partial(self, partial):
        """Host-side fill of the send buffer (tests and callers that hold the partial as bytes)."""
        if len(partial) != PARTIAL_BYTES:
            raise ValueError("a G1 partial is %d bytes" % PARTIAL_BYTES)
        t = self._torch.frombuffer(bytearray(partial), dtype=self._torch.uint8)
        self.send.copy_(t)

    def all_gather(self):
        """The exchange step.  On return every rank's receive buffer holds the world partials in rank
        order and (NCCL) the collective has completed on the device, so another stream may read it."""
        self._dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
        if self.on_device:
            # device-wide: the receive buffer is read next by the library's own stream, which knows nothing
            # about torch's streams
            self._torch.cuda.synchronize()

    def gathered_bytes(self):
        return bytes(self.recv.cpu().numpy().tobytes())


def g1_msm_sharded(exchange, table, scalars, n_local, offset=0, sc_offset=0):
    """This rank's share of a sharded G1 MSM: `table` / `scalars` are device handles of the LOCAL point
    and scalar ranges (shard_range of the global ones).  Returns the affine result on rank 0 and None on
    the other ranks."""
    from . import native
    native.g1_msm_dev_partial(table, offset, scalars, sc_offset, n_local, out_addr=exchange.send_addr)
    exchange.all_gather()
    if exchange.rank != 0:
        return None
    return native.g1_combine_partials(exchange.recv_addr, exchange.world)
