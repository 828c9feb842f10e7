// Device runtime behind the C ABI: error plumbing, device selection, stream-ordered memory pool,
// streams/events, reduction scratch.  (What legate.core's allocator / StreamPool give the reference.)
#include "cnb_common.cuh"

#include <atomic>
#include <cstdarg>
#include <cstring>
#include <mutex>
#include <vector>

namespace cnb {

namespace {
thread_local char g_error[1024] = "";
std::atomic<uint64_t> g_launches{0};

struct DeviceState {
  bool initialised = false;
  int device       = -1;
  int sms          = 0;
  // scalar-reduction scratch: SLOTS independent {partials, ticket} pairs handed out round-robin so
  // that reductions in flight on different streams never share a ticket
  static constexpr int SLOTS     = 64;
  static constexpr int MAX_GRID  = 4096;
  char* red_partials             = nullptr;  // SLOTS * MAX_GRID * 16 bytes
  unsigned int* red_tickets      = nullptr;  // SLOTS
  std::atomic<unsigned int> next_slot{0};
};
DeviceState g_dev;
std::mutex g_init_mu;
}  // namespace

int set_error(int code, const char* fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}

int check_cuda(cudaError_t e, const char* what)
{
  if (e == cudaSuccess) return CNB_OK;
  return set_error(CNB_ERR_CUDA, "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
}

// ---- launch accounting / tracing ---------------------------------------------------------------
namespace {
struct TraceRecord {
  int task, op, dtype, kernel_kind;
  long long elems, bytes;
  cudaEvent_t start, stop;
  float ms;
};
struct TraceState {
  std::mutex mu;
  bool active = false;
  int capacity = 0, count = 0;
  std::vector<TraceRecord> records;
};
TraceState g_trace;
thread_local int g_tag_task = 0, g_tag_op = 0, g_tag_dtype = 0;
}  // namespace

void set_task_tag(int task, int op, int dtype)
{
  g_tag_task  = task;
  g_tag_op    = op;
  g_tag_dtype = dtype;
}

LaunchScope::LaunchScope(cudaStream_t stream, int kernel_kind, long long elems, long long bytes)
  : stream_(stream), slot_(-1)
{
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (!g_trace.active) return;
  std::lock_guard<std::mutex> g(g_trace.mu);
  if (!g_trace.active || g_trace.count >= g_trace.capacity) return;
  slot_          = g_trace.count++;
  TraceRecord& r = g_trace.records[slot_];
  r.task         = g_tag_task;
  r.op           = g_tag_op;
  r.dtype        = g_tag_dtype;
  r.kernel_kind  = kernel_kind;
  r.elems        = elems;
  r.bytes        = bytes;
  r.ms           = -1.0f;
  cudaEventRecord(r.start, stream_);
}

LaunchScope::~LaunchScope()
{
  if (slot_ >= 0) cudaEventRecord(g_trace.records[slot_].stop, stream_);
}

int init_device(int device)
{
  std::lock_guard<std::mutex> g(g_init_mu);
  if (g_dev.initialised && g_dev.device == device) return CNB_OK;
  int count = 0;
  CNB_CUDA(cudaGetDeviceCount(&count));
  if (count <= 0) return set_error(CNB_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  if (device < 0 || device >= count)
    return set_error(CNB_ERR_BAD_ARG, "device %d out of range (%d visible)", device, count);
  CNB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  CNB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return set_error(CNB_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only",
                     device, prop.major, prop.minor);
  g_dev.sms = prop.multiProcessorCount;
  // keep freed blocks in the pool (stream-ordered reuse, no cudaFree round trips)
  cudaMemPool_t pool;
  CNB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t threshold = UINT64_MAX;
  CNB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
  if (g_dev.red_partials == nullptr) {
    CNB_CUDA(cudaMalloc(&g_dev.red_partials, (size_t)DeviceState::SLOTS * DeviceState::MAX_GRID * 16));
    CNB_CUDA(cudaMalloc(&g_dev.red_tickets, DeviceState::SLOTS * sizeof(unsigned int)));
    CNB_CUDA(cudaMemset(g_dev.red_tickets, 0, DeviceState::SLOTS * sizeof(unsigned int)));
  }
  g_dev.device      = device;
  g_dev.initialised = true;
  return CNB_OK;
}

int ensure_init()
{
  if (g_dev.initialised) return CNB_OK;
  return init_device(0);
}

int sm_count() { return g_dev.sms > 0 ? g_dev.sms : 148; }

struct RedScratch {
  char* partials;
  unsigned int* ticket;
  int max_grid;
};
int red_acquire_scratch(RedScratch& s, cudaStream_t)
{
  int rc = ensure_init();
  if (rc != CNB_OK) return rc;
  const unsigned int slot = g_dev.next_slot.fetch_add(1) % DeviceState::SLOTS;
  s.partials              = g_dev.red_partials + (size_t)slot * DeviceState::MAX_GRID * 16;
  s.ticket                = g_dev.red_tickets + slot;
  s.max_grid              = DeviceState::MAX_GRID;
  return CNB_OK;
}

void* pool_alloc(size_t nbytes, cudaStream_t stream)
{
  if (ensure_init() != CNB_OK) return nullptr;
  void* p = nullptr;
  if (nbytes == 0) nbytes = 1;
  if (check_cuda(cudaMallocAsync(&p, nbytes, stream), "cudaMallocAsync") != CNB_OK) return nullptr;
  return p;
}
int pool_free(void* p, cudaStream_t stream)
{
  if (p == nullptr) return CNB_OK;
  return check_cuda(cudaFreeAsync(p, stream), "cudaFreeAsync");
}

}  // namespace cnb

using namespace cnb;

extern "C" {

int cnb_device_count(void)
{
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return count;
}
int cnb_init(int32_t device) { return init_device(device); }
int cnb_sm_count(void) { return ensure_init() == CNB_OK ? sm_count() : 0; }

void* cnb_malloc(size_t nbytes, void* stream) { return pool_alloc(nbytes, (cudaStream_t)stream); }
int cnb_free(void* ptr, void* stream) { return pool_free(ptr, (cudaStream_t)stream); }

void* cnb_host_alloc(size_t nbytes)
{
  void* p = nullptr;
  if (check_cuda(cudaMallocHost(&p, nbytes ? nbytes : 1), "cudaMallocHost") != CNB_OK) return nullptr;
  return p;
}
int cnb_host_free(void* ptr) { return check_cuda(cudaFreeHost(ptr), "cudaFreeHost"); }

int cnb_memcpy_h2d(void* dst, const void* src, size_t n, void* stream)
{
  return check_cuda(cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, (cudaStream_t)stream), "H2D");
}
int cnb_memcpy_d2h(void* dst, const void* src, size_t n, void* stream)
{
  return check_cuda(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, (cudaStream_t)stream), "D2H");
}
int cnb_memcpy_d2d(void* dst, const void* src, size_t n, void* stream)
{
  return check_cuda(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream), "D2D");
}
int cnb_memset(void* dst, int value, size_t n, void* stream)
{
  return check_cuda(cudaMemsetAsync(dst, value, n, (cudaStream_t)stream), "memset");
}

void* cnb_stream_create(void)
{
  if (ensure_init() != CNB_OK) return nullptr;
  cudaStream_t s = nullptr;
  if (check_cuda(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "stream create") != CNB_OK)
    return nullptr;
  return s;
}
int cnb_stream_destroy(void* s) { return check_cuda(cudaStreamDestroy((cudaStream_t)s), "stream destroy"); }
int cnb_stream_synchronize(void* s)
{
  return check_cuda(cudaStreamSynchronize((cudaStream_t)s), "stream synchronize");
}
int cnb_device_synchronize(void) { return check_cuda(cudaDeviceSynchronize(), "device synchronize"); }

void* cnb_event_create(void)
{
  cudaEvent_t e = nullptr;
  if (check_cuda(cudaEventCreate(&e), "event create") != CNB_OK) return nullptr;
  return e;
}
int cnb_event_destroy(void* e) { return check_cuda(cudaEventDestroy((cudaEvent_t)e), "event destroy"); }
int cnb_event_record(void* e, void* s)
{
  return check_cuda(cudaEventRecord((cudaEvent_t)e, (cudaStream_t)s), "event record");
}
int cnb_event_synchronize(void* e) { return check_cuda(cudaEventSynchronize((cudaEvent_t)e), "event sync"); }
int cnb_stream_wait_event(void* s, void* e)
{
  return check_cuda(cudaStreamWaitEvent((cudaStream_t)s, (cudaEvent_t)e, 0), "stream wait event");
}
int cnb_event_elapsed_ms(void* a, void* b, float* ms)
{
  return check_cuda(cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b), "event elapsed");
}
int cnb_mem_info(size_t* free_bytes, size_t* total_bytes)
{
  return check_cuda(cudaMemGetInfo(free_bytes, total_bytes), "mem info");
}
uint64_t cnb_launch_count(void) { return g_launches.load(); }

int cnb_trace_start(int32_t capacity)
{
  std::lock_guard<std::mutex> g(g_trace.mu);
  if (capacity <= 0) return set_error(CNB_ERR_BAD_ARG, "trace capacity must be positive");
  while ((int)g_trace.records.size() < capacity) {
    TraceRecord r{};
    CNB_CUDA(cudaEventCreate(&r.start));
    CNB_CUDA(cudaEventCreate(&r.stop));
    g_trace.records.push_back(r);
  }
  g_trace.capacity = capacity;
  g_trace.count    = 0;
  g_trace.active   = true;
  return CNB_OK;
}

int cnb_trace_stop(void)
{
  std::lock_guard<std::mutex> g(g_trace.mu);
  g_trace.active = false;
  for (int i = 0; i < g_trace.count; ++i) {
    TraceRecord& r = g_trace.records[i];
    CNB_CUDA(cudaEventSynchronize(r.stop));
    CNB_CUDA(cudaEventElapsedTime(&r.ms, r.start, r.stop));
  }
  return g_trace.count;
}

int cnb_trace_get(int32_t index, cnb_trace_record_t* out)
{
  std::lock_guard<std::mutex> g(g_trace.mu);
  if (index < 0 || index >= g_trace.count || out == nullptr)
    return set_error(CNB_ERR_BAD_ARG, "trace index %d out of range", index);
  const TraceRecord& r = g_trace.records[index];
  out->task        = r.task;
  out->op          = r.op;
  out->dtype       = r.dtype;
  out->kernel_kind = r.kernel_kind;
  out->elems       = r.elems;
  out->bytes       = r.bytes;
  out->ms          = r.ms;
  return CNB_OK;
}
const char* cnb_last_error(void) { return g_error; }
const char* cnb_version(void) { return "cunumeric_b200 0.1 (sm_100a)"; }
}
