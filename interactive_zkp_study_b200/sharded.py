"""Multi-GPU MSM: one process per GPU, points sharded by contiguous range.

sum_i k_i P_i is a sum of independent per-range partial sums, so every rank runs the full single-GPU
Pippenger on its own range of the static table (resident on its GPU) and emits ONE un-normalised XYZZ
point.  The only exchange step of the path is an all-gather of world x 128 bytes, and it lives INSIDE
the library: ``zkp_g1_msm_multi`` enqueues local MSM -> ``ncclAllGather`` -> fold -> affine on the
library's stream with no host synchronisation in between (csrc/comm.cu, csrc/msm_api.cuh).  There is no
other collective: the NTT / quotient / PLONK rounds are single-GPU work ("replicas only").

This module is the host side of that: the shard arithmetic and the bootstrap of the library's NCCL
communicator.  The 128-byte communicator id travels from rank 0 to the other ranks over a plain TCP
rendezvous (standard library sockets) -- or over any channel the caller already has (``exchange_id``).
The reference has no multi-device path; the call being scaled is kzg.commit
(/root/reference/zkp/plonk/kzg.py:32-67) / the proof-element sums of proving.py:23-75.

Partial wire format (G1): x | y | zz | zzz, each 32 bytes little-endian, Montgomery form (x * 2^256 mod p);
the point is (x/zz, y/zzz); zz == 0 encodes the point at infinity.
"""
import os
import socket
import time

PARTIAL_BYTES = 128
ID_BYTES = 128


def shard_range(total, rank, world):
    """Contiguous range [start, start + count) of `total` points owned by `rank` of `world`; the
    remainder goes to the first ranks one point each, so ranges tile [0, total) in rank order."""
    if world <= 0 or not 0 <= rank < world or total < 0:
        raise ValueError("shard_range: need 0 <= rank < world and total >= 0")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def bind_to_gpu_numa_node():
    """Pins the calling process to the CPUs of the NUMA node its GPU hangs off (sysfs), so that page-locked
    staging buffers allocated afterwards are local to the GPU's PCIe root: with 8 ranks uploading at once,
    scalars that cross the socket interconnect are the slowest part of the end-to-end step.  Returns the
    node number, or None when the topology is not exposed (nothing is changed then)."""
    from . import native
    try:
        bdf = native.device_pci_bus_id()
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (OSError, ValueError):
        return None


def _recv_exact(conn, n):
    out = bytearray()
    while len(out) < n:
        chunk = conn.recv(n - len(out))
        if not chunk:
            raise ConnectionError("rendezvous peer closed the connection")
        out += chunk
    return bytes(out)


def tcp_broadcast(payload, rank, world, addr="127.0.0.1", port=29617, timeout=120.0):
    """Rank 0 hands `payload` (bytes) to the world - 1 other ranks; every rank returns it.

    Rank 0 listens on (addr, port) until each peer has connected, announced its rank and been served;
    peers retry the connection until rank 0 is up.  One-shot, no daemon, standard library only."""
    if world == 1:
        return bytes(payload)
    deadline = time.monotonic() + timeout
    if rank == 0:
        payload = bytes(payload)
        srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        srv.bind((addr, port))
        srv.listen(world)
        srv.settimeout(timeout)
        served = set()
        try:
            while len(served) < world - 1:
                conn, _ = srv.accept()
                with conn:
                    conn.settimeout(timeout)
                    peer = int.from_bytes(_recv_exact(conn, 4), "little")
                    if not 1 <= peer < world:
                        raise ValueError("rendezvous: unexpected peer rank %d" % peer)
                    conn.sendall(len(payload).to_bytes(4, "little") + payload)
                    served.add(peer)
        finally:
            srv.close()
        return payload
    while True:
        try:
            with socket.create_connection((addr, port), timeout=5.0) as conn:
                conn.settimeout(timeout)
                conn.sendall(int(rank).to_bytes(4, "little"))
                size = int.from_bytes(_recv_exact(conn, 4), "little")
                return _recv_exact(conn, size)
        except (ConnectionRefusedError, ConnectionResetError, socket.timeout, OSError):
            if time.monotonic() > deadline:
                raise TimeoutError("rendezvous: rank 0 did not answer on %s:%d" % (addr, port))
            time.sleep(0.05)


class Communicator:
    """The library's NCCL communicator for this process (one per process, as one GPU per process)."""

    def __init__(self, rank, world, exchange_id=None, addr=None, port=None):
        """Collective over all ranks.  `exchange_id(id_or_None) -> id` may replace the TCP rendezvous
        (rank 0 passes the id, the others None; all get the id back)."""
        from . import native
        if world <= 0 or not 0 <= rank < world:
            raise ValueError("Communicator: need 0 <= rank < world")
        self.rank, self.world = rank, world
        mine = native.comm_unique_id() if rank == 0 else None
        if exchange_id is not None:
            comm_id = exchange_id(mine)
        else:
            addr = addr or os.environ.get("MASTER_ADDR", "127.0.0.1")
            port = int(port or os.environ.get("ZKP_B200_RDZV_PORT", 0) or int(os.environ.get("MASTER_PORT", "29500")) + 117)
            comm_id = tcp_broadcast(mine, rank, world, addr, port)
        native.comm_init(rank, world, comm_id)
        self.info = native.comm_info()

    @classmethod
    def from_env(cls, **kw):
        """RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT from the environment, as one-process-per-GPU launchers set them."""
        return cls(int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), **kw)

    def barrier(self):
        from . import native
        native.comm_barrier()

    def close(self):
        from . import native
        native.comm_destroy()


def g1_msm_sharded(comm, table, scalars, n_local, offset=0, sc_offset=0):
    """This rank's share of a sharded G1 MSM: `table` / `scalars` are device handles of the LOCAL point
    and scalar ranges (shard_range of the global ones).  Collective; every rank returns the affine
    result.  `scalars` may also be host bytes / a pinned address (uploaded inside the call)."""
    from . import native
    if comm is None:
        raise ValueError("g1_msm_sharded needs a Communicator")
    if isinstance(scalars, native.DeviceHandle):
        return native.g1_msm_multi(table, offset, scalars, sc_offset, n_local)
    return native.g1_msm_multi_table(table, offset, scalars, n_local)
