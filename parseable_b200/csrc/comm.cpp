// NCCL communicator owned by the library: one process per GPU, partial
// aggregate tables all-reduced over NVLink 5 / NVSwitch.  This is the GPU
// analogue of DataFusion's Partial -> RepartitionExec(Hash) -> FinalPartitioned
// merge that the reference relies on.
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "engine.hpp"

namespace pqb {

namespace {
std::mutex g_mu;
ncclComm_t g_comm = nullptr;
int g_nranks = 0, g_rank = 0;
uint64_t g_epoch = 0;   // bumped by every comm_init_rank: caches of cross-rank agreements are tagged with it

void check(ncclResult_t r, const char* what) {
  if (r != ncclSuccess) throw Error(PQ_ERR_CUDA, std::string(what) + ": " + ncclGetErrorString(r));
}
}  // namespace

static_assert(sizeof(ncclUniqueId) <= PQ_COMM_ID_BYTES, "unique id does not fit the ABI buffer");

int comm_unique_id(uint8_t* id) {
  ncclUniqueId u;
  check(ncclGetUniqueId(&u), "ncclGetUniqueId");
  std::memset(id, 0, PQ_COMM_ID_BYTES);
  std::memcpy(id, &u, sizeof(u));
  return PQ_OK;
}

int comm_init_rank(const uint8_t* id, int nranks, int rank) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_comm) throw Error(PQ_ERR_INVALID_ARG, "communicator already initialised");
  if (nranks < 1 || rank < 0 || rank >= nranks) throw Error(PQ_ERR_INVALID_ARG, "bad rank / nranks");
  ncclUniqueId u;
  std::memcpy(&u, id, sizeof(u));
  check(ncclCommInitRank(&g_comm, nranks, u, rank), "ncclCommInitRank");
  g_nranks = nranks;
  g_rank = rank;
  g_epoch++;
  return PQ_OK;
}

int comm_destroy() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_comm) {
    ncclCommDestroy(g_comm);
    g_comm = nullptr;
  }
  g_nranks = 0;
  return PQ_OK;
}

bool comm_active() { return g_comm != nullptr; }
uint64_t comm_epoch() { return g_epoch; }
void comm_group_begin() { check(ncclGroupStart(), "ncclGroupStart"); }
void comm_group_end() { check(ncclGroupEnd(), "ncclGroupEnd"); }
int comm_nranks() { return g_nranks; }
int comm_rank() { return g_rank; }

void comm_allreduce_u64(void* buf, size_t count, int op, cudaStream_t s) {
  if (!g_comm) throw Error(PQ_ERR_INVALID_ARG, "no communicator");
  ncclDataType_t dt = op == 3 ? ncclFloat64 : ncclInt64;
  ncclRedOp_t ro = op == 1 ? ncclMin : (op == 2 ? ncclMax : ncclSum);
  // int64 sum wraps in two's complement exactly like SUM(Int64) on one device
  check(ncclAllReduce(buf, buf, count, dt, ro, g_comm, s), "ncclAllReduce");
}

void comm_allgather_bytes(const void* send, void* recv, size_t bytes_per_rank, cudaStream_t s) {
  if (!g_comm) throw Error(PQ_ERR_INVALID_ARG, "no communicator");
  check(ncclAllGather(send, recv, bytes_per_rank, ncclUint8, g_comm, s), "ncclAllGather");
}

}  // namespace pqb
