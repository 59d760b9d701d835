// Multi-GPU exchange step of the sharded MSM: one process per GPU, one NCCL communicator
// owned by the library, ONE collective on the data path -- the all-gather of one un-normalised partial
// sum per rank (128 B in G1, 256 B in G2) -- enqueued on the library's own stream between the local
// Pippenger and the fold, so nothing on the host sits between them.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 inside zkp_comm_unique_id / zkp_comm_init): a
// single-GPU user of the library needs no NCCL at all, and a process that has already loaded a
// libnccl.so.2 (e.g. the one bundled with PyTorch) shares that copy instead of mapping a second one.
#pragma once
#include "common.cuh"

namespace zkp {

struct CommState {
  bool ready = false;
  int rank = 0, world = 1;
  int nccl_version = 0;
  DevBuf gathered[2];  // world x (largest partial) bytes: receive buffer of the all-gather, one per stream lane
};

CommState& comm_state();
// Enqueues the all-gather of `bytes` bytes per rank (send -> recv[rank * bytes]) on `st`.  world == 1
// degenerates to a device-to-device copy through the same NCCL call path.  Throws if zkp_comm_init has
// not succeeded.  Collectives of the one communicator are chained by an event whatever stream they are
// enqueued on, so two lanes never have two of them in flight at once (every rank issues them in the same order).
void comm_all_gather(const void* send, void* recv, size_t bytes, cudaStream_t st);

}  // namespace zkp
