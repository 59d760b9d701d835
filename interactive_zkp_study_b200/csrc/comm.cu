// libzkp_b200: the NCCL communicator behind zkp_g1_msm_multi / zkp_g2_msm_multi (include/zkp_b200.h).
// See comm.cuh for the design; the reference has no multi-device path, the call being scaled is the
// commit loop /root/reference/zkp/plonk/kzg.py:59-67 (and the proof-element sums of proving.py:23-75).
#include <dlfcn.h>
#include <nccl.h>
#include "comm.cuh"
#include "registry.cuh"

namespace zkp {

namespace {

struct NcclApi {
  void* dl = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;
ncclComm_t g_comm = nullptr;
cudaEvent_t g_ev_last = nullptr;  // completion of the communicator's most recent collective (on whatever stream)
CommState g_state;
DevBuf g_barrier_word;

template <class T>
void bind(T& fn, const char* name) {
  fn = reinterpret_cast<T>(dlsym(g_nccl.dl, name));
  if (!fn) throw std::runtime_error(std::string("libnccl.so.2 does not export ") + name);
}

void nccl_load() {
  if (g_nccl.dl) return;
  // RTLD_NOLOAD first: reuse the copy the process already holds (PyTorch ships its own libnccl.so.2)
  const char* override_path = getenv("ZKP_B200_NCCL_LIB");
  void* h = override_path ? dlopen(override_path, RTLD_NOW | RTLD_GLOBAL) : nullptr;
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) throw std::runtime_error(std::string("multi-GPU entry points need NCCL and libnccl.so.2 could not be loaded: ") + dlerror());
  g_nccl.dl = h;
  bind(g_nccl.GetVersion, "ncclGetVersion");
  bind(g_nccl.GetUniqueId, "ncclGetUniqueId");
  bind(g_nccl.CommInitRank, "ncclCommInitRank");
  bind(g_nccl.CommDestroy, "ncclCommDestroy");
  bind(g_nccl.AllGather, "ncclAllGather");
  bind(g_nccl.AllReduce, "ncclAllReduce");
  bind(g_nccl.GetErrorString, "ncclGetErrorString");
}

void nccl_check(ncclResult_t r, const char* what) {
  if (r != ncclSuccess) throw std::runtime_error(std::string(what) + " failed: " + g_nccl.GetErrorString(r));
}

}  // namespace

CommState& comm_state() { return g_state; }

void comm_all_gather(const void* send, void* recv, size_t bytes, cudaStream_t st) {
  if (!g_state.ready) throw InvalidArgument("multi-GPU MSM: zkp_comm_init has not been called on this rank");
  if (!g_ev_last) CUDA_CHECK(cudaEventCreateWithFlags(&g_ev_last, cudaEventDisableTiming));
  else CUDA_CHECK(cudaStreamWaitEvent(st, g_ev_last, 0));
  nccl_check(g_nccl.AllGather(send, recv, bytes, ncclUint8, g_comm, st), "ncclAllGather");
  CUDA_CHECK(cudaEventRecord(g_ev_last, st));
}

}  // namespace zkp

using namespace zkp;

extern "C" {

int zkp_comm_unique_id(uint8_t out_id[ZKP_COMM_ID_BYTES]) {
  return guarded([&](Context&) {
    if (!out_id) throw InvalidArgument("zkp_comm_unique_id: null output");
    static_assert(sizeof(ncclUniqueId) == ZKP_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    nccl_load();
    ncclUniqueId id;
    nccl_check(g_nccl.GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(out_id, &id, sizeof id);
  });
}

int zkp_comm_init(int rank, int world, const uint8_t id[ZKP_COMM_ID_BYTES]) {
  return guarded([&](Context& c) {
    if (!id || world < 1 || rank < 0 || rank >= world) throw InvalidArgument("zkp_comm_init: need 0 <= rank < world and an id");
    if (g_state.ready) throw InvalidArgument("zkp_comm_init: a communicator already exists (zkp_comm_destroy first)");
    nccl_load();
    nccl_check(g_nccl.GetVersion(&g_state.nccl_version), "ncclGetVersion");
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
    nccl_check(g_nccl.CommInitRank(&g_comm, world, uid, rank), "ncclCommInitRank");
    g_state.rank = rank;
    g_state.world = world;
    for (DevBuf& b : g_state.gathered) b.reserve((size_t)world * 256);
    g_barrier_word.reserve(64);
    CUDA_CHECK(cudaMemsetAsync(g_barrier_word.p, 0, 64, c.stream));
    g_state.ready = true;
  });
}

int zkp_comm_info(int* rank, int* world, int* nccl_version) {
  return guarded([&](Context&) {
    if (!g_state.ready) throw InvalidArgument("zkp_comm_info: no communicator");
    if (rank) *rank = g_state.rank;
    if (world) *world = g_state.world;
    if (nccl_version) *nccl_version = g_state.nccl_version;
  });
}

int zkp_comm_barrier(void) {
  return guarded([&](Context& c) {
    if (!g_state.ready) throw InvalidArgument("zkp_comm_barrier: no communicator");
    if (g_ev_last) CUDA_CHECK(cudaStreamWaitEvent(c.stream, g_ev_last, 0));
    nccl_check(g_nccl.AllReduce(g_barrier_word.p, g_barrier_word.p, 1, ncclUint32, ncclSum, g_comm, c.stream), "ncclAllReduce");
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
  });
}

int zkp_comm_destroy(void) {
  return guarded([&](Context& c) {
    if (!g_state.ready) return;
    CUDA_CHECK(cudaStreamSynchronize(c.stream));
    CUDA_CHECK(cudaStreamSynchronize(c.stream2));
    g_state.ready = false;
    nccl_check(g_nccl.CommDestroy(g_comm), "ncclCommDestroy");
    g_comm = nullptr;
  });
}

}  // extern "C"
