// Multi-GPU exchange layer: thin, stream-ordered wrappers over NCCL (one process per GPU).
// NCCL is dlopen'ed at first use (CNB_NCCL_LIB, else libnccl.so.2 on the loader path — the copy
// torch already mapped when the host program imported it), so single-GPU users need no NCCL.
#include "cnb_common.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <mutex>

namespace cnb {
namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  const char* (*GetErrorString)(ncclResult_t);
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

int load_nccl()
{
  std::lock_guard<std::mutex> g(g_nccl_mu);
  if (g_nccl.handle != nullptr) return CNB_OK;
  const char* env = getenv("CNB_NCCL_LIB");
  void* h         = nullptr;
  if (env != nullptr && env[0] != 0) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  if (h == nullptr) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (h == nullptr) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (h == nullptr) return set_error(CNB_ERR_COMM, "cannot load NCCL: %s", dlerror());
#define LOAD(field, sym)                                                            \
  *reinterpret_cast<void**>(&g_nccl.field) = dlsym(h, sym);                         \
  if (g_nccl.field == nullptr) return set_error(CNB_ERR_COMM, "NCCL symbol %s missing", sym);
  LOAD(GetUniqueId, "ncclGetUniqueId")
  LOAD(CommInitRank, "ncclCommInitRank")
  LOAD(CommDestroy, "ncclCommDestroy")
  LOAD(GroupStart, "ncclGroupStart")
  LOAD(GroupEnd, "ncclGroupEnd")
  LOAD(Send, "ncclSend")
  LOAD(Recv, "ncclRecv")
  LOAD(AllReduce, "ncclAllReduce")
  LOAD(AllGather, "ncclAllGather")
  LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
  g_nccl.handle = h;
  return CNB_OK;
}

int check_nccl(ncclResult_t r, const char* what)
{
  if (r == ncclSuccess) return CNB_OK;
  return set_error(CNB_ERR_COMM, "NCCL error %d (%s) in %s", (int)r,
                   g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?", what);
}
#define CNB_NCCL(expr)                       \
  do {                                       \
    int _rc = check_nccl((expr), #expr);     \
    if (_rc != CNB_OK) return _rc;           \
  } while (0)

}  // namespace
}  // namespace cnb

using namespace cnb;

extern "C" {

int cnb_comm_unique_id(void* id_out)
{
  static_assert(sizeof(ncclUniqueId) == CNB_COMM_ID_BYTES, "ncclUniqueId size");
  int rc = load_nccl();
  if (rc != CNB_OK) return rc;
  CNB_NCCL(g_nccl.GetUniqueId(static_cast<ncclUniqueId*>(id_out)));
  return CNB_OK;
}

void* cnb_comm_init(const void* id, int32_t nranks, int32_t rank)
{
  if (load_nccl() != CNB_OK) return nullptr;
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  ncclComm_t comm = nullptr;
  if (check_nccl(g_nccl.CommInitRank(&comm, nranks, uid, rank), "ncclCommInitRank") != CNB_OK)
    return nullptr;
  return comm;
}

int cnb_comm_destroy(void* comm)
{
  if (comm == nullptr) return CNB_OK;
  CNB_NCCL(g_nccl.CommDestroy(static_cast<ncclComm_t>(comm)));
  return CNB_OK;
}

int cnb_comm_group_start(void)
{
  int rc = load_nccl();
  if (rc != CNB_OK) return rc;
  CNB_NCCL(g_nccl.GroupStart());
  return CNB_OK;
}
int cnb_comm_group_end(void)
{
  CNB_NCCL(g_nccl.GroupEnd());
  return CNB_OK;
}

int cnb_comm_send(void* comm, const void* buf, size_t nbytes, int32_t peer, void* stream)
{
  CNB_NCCL(g_nccl.Send(buf, nbytes, ncclUint8, peer, static_cast<ncclComm_t>(comm),
                       (cudaStream_t)stream));
  return CNB_OK;
}
int cnb_comm_recv(void* comm, void* buf, size_t nbytes, int32_t peer, void* stream)
{
  CNB_NCCL(g_nccl.Recv(buf, nbytes, ncclUint8, peer, static_cast<ncclComm_t>(comm),
                       (cudaStream_t)stream));
  return CNB_OK;
}

int cnb_comm_allreduce(void* comm, const void* send, void* recv, size_t count, int32_t dtype,
                       int32_t red_op, void* stream)
{
  ncclDataType_t dt;
  size_t mult = 1;
  switch (dtype) {
    case CNB_BOOL:
    case CNB_UINT8: dt = ncclUint8; break;
    case CNB_INT8: dt = ncclInt8; break;
    case CNB_INT32: dt = ncclInt32; break;
    case CNB_UINT32: dt = ncclUint32; break;
    case CNB_INT64: dt = ncclInt64; break;
    case CNB_UINT64: dt = ncclUint64; break;
    case CNB_FLOAT16: dt = ncclFloat16; break;
    case CNB_FLOAT32: dt = ncclFloat32; break;
    case CNB_FLOAT64: dt = ncclFloat64; break;
    case CNB_COMPLEX64:
      dt   = ncclFloat32;
      mult = 2;
      break;
    case CNB_COMPLEX128:
      dt   = ncclFloat64;
      mult = 2;
      break;
    default: return set_error(CNB_ERR_UNSUPPORTED, "allreduce: dtype %d has no NCCL type", dtype);
  }
  ncclRedOp_t op;
  switch (red_op) {
    case CNB_RED_SUM:
    case CNB_RED_NANSUM:
    case CNB_RED_SUM_SQUARES:
    case CNB_RED_VARIANCE:
    case CNB_RED_COUNT_NONZERO: op = (dtype == CNB_BOOL) ? ncclMax : ncclSum; break;
    case CNB_RED_PROD:
    case CNB_RED_NANPROD: op = (dtype == CNB_BOOL) ? ncclMin : ncclProd; break;
    case CNB_RED_MAX:
    case CNB_RED_NANMAX:
    case CNB_RED_ANY:
    case CNB_RED_CONTAINS: op = ncclMax; break;
    case CNB_RED_MIN:
    case CNB_RED_NANMIN:
    case CNB_RED_ALL: op = ncclMin; break;
    default: return set_error(CNB_ERR_UNSUPPORTED, "allreduce: reduction %d not supported", red_op);
  }
  if (mult == 2 && op != ncclSum)
    return set_error(CNB_ERR_UNSUPPORTED, "allreduce: complex supports SUM only");
  CNB_NCCL(g_nccl.AllReduce(send, recv, count * mult, dt, op, static_cast<ncclComm_t>(comm),
                            (cudaStream_t)stream));
  return CNB_OK;
}

int cnb_comm_allgather(void* comm, const void* send, void* recv, size_t nbytes_per_rank,
                       void* stream)
{
  CNB_NCCL(g_nccl.AllGather(send, recv, nbytes_per_rank, ncclUint8, static_cast<ncclComm_t>(comm),
                            (cudaStream_t)stream));
  return CNB_OK;
}
}
