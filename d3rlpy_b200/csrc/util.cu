// Library plumbing: error text, launch counter, stream/graph helpers, staged copies.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace d3b {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

char* err_buf() { return g_err; }

int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }

}  // namespace d3b

using namespace d3b;

namespace d3b {
static bool g_pdl = true;
bool pdl_enabled() { return g_pdl; }
}  // namespace d3b

// programmatic dependent launch between the kernels of an update (on by default); 0 restores plain stream order
extern "C" int d3b_set_pdl(int enabled) {
  d3b::g_pdl = enabled != 0;
  return D3B_OK;
}

extern "C" const char* d3b_last_error(void) { return err_buf(); }
extern "C" int d3b_abi_version(void) { return D3B_ABI_VERSION; }
extern "C" int64_t d3b_launch_count(void) { return launches(); }

extern "C" int d3b_device_info(int device, int* sm_count, int* cc_major, int* cc_minor) {
  cudaDeviceProp p;
  D3B_CUDA(cudaGetDeviceProperties(&p, device));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return D3B_OK;
}

// Small staged copies as a KERNEL between device memory and pinned host memory (either direction): pinned
// allocations are device-addressable under unified addressing, and inside the update graph a kernel node chained by
// programmatic dependent launch starts sooner than a copy-engine node (the minibatch upload in front of the update
// and the 256-byte metric read-back behind it are latency, not bandwidth).  Large transfers keep cudaMemcpyAsync.
namespace d3b {
__global__ void __launch_bounds__(256) copy_mapped_kernel(unsigned char* __restrict__ dst,
                                                          const unsigned char* __restrict__ src, long long bytes,
                                                          int vec16) {
  pdl_trigger();
  pdl_wait();
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long done = 0;
  if (vec16) {
    const long long n16 = bytes >> 4;
    for (long long i = tid; i < n16; i += stride) ((uint4*)dst)[i] = ((const uint4*)src)[i];
    done = n16 << 4;
  }
  for (long long i = done + tid; i < bytes; i += stride) dst[i] = src[i];
}
__global__ void __launch_bounds__(256) zero_words_kernel(uint32_t* __restrict__ dst, long long n_words) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_words;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = 0u;
}
}  // namespace d3b

namespace {
constexpr long long kSmallCopy = 256 << 10;  // below this a copy / fill is latency: run it as a PDL-chained kernel node
int copy_by_kernel(void* dst, const void* src, long long bytes, cudaStream_t st, const char* what) {
  const int vec16 = (((uintptr_t)dst | (uintptr_t)src) & 15) == 0;
  long long blocks = ceil_div_ll(vec16 ? ceil_div_ll(bytes, 16) : bytes, 256);
  if (blocks > kNumSM * 4) blocks = kNumSM * 4;
  launch_pdl(copy_mapped_kernel, dim3((unsigned)blocks), dim3(256), 0, st, (unsigned char*)dst,
             (const unsigned char*)src, bytes, vec16);
  return check_launch(what);
}
}  // namespace


extern "C" int d3b_memset_zero(void* ptr, int64_t bytes, void* stream) {
  D3B_REQUIRE(bytes >= 0 && (ptr || bytes == 0), "memset_zero: bad arguments");
  if (bytes == 0) return D3B_OK;
  if (bytes <= kSmallCopy && (((uintptr_t)ptr | (uintptr_t)bytes) & 3) == 0) {
    const long long words = bytes >> 2;
    long long blocks = ceil_div_ll(words, 256);
    if (blocks > kNumSM * 4) blocks = kNumSM * 4;
    launch_pdl(zero_words_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, (uint32_t*)ptr, words);
    return check_launch("memset_zero");
  }
  D3B_CUDA(cudaMemsetAsync(ptr, 0, (size_t)bytes, (cudaStream_t)stream));
  count_launch();
  return D3B_OK;
}

extern "C" int d3b_copy_h2d(void* dst, const void* src_pinned, int64_t bytes, void* stream) {
  D3B_REQUIRE(bytes >= 0, "copy_h2d: bytes < 0");
  if (bytes == 0) return D3B_OK;
  D3B_CUDA(cudaMemcpyAsync(dst, src_pinned, (size_t)bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return D3B_OK;
}

extern "C" int d3b_copy_d2h(void* dst_pinned, const void* src, int64_t bytes, void* stream) {
  D3B_REQUIRE(bytes >= 0, "copy_d2h: bytes < 0");
  if (bytes == 0) return D3B_OK;
  D3B_CUDA(cudaMemcpyAsync(dst_pinned, src, (size_t)bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return D3B_OK;
}

extern "C" int d3b_copy_mapped(void* dst, const void* src, int64_t bytes, void* stream) {
  D3B_REQUIRE(bytes >= 0 && ((dst && src) || bytes == 0), "copy_mapped: bad arguments");
  if (bytes == 0) return D3B_OK;
  // a pinned host buffer is addressed through its device alias (the same pointer under unified addressing); pageable
  // memory is refused here instead of faulting in the kernel
  void* ends[2] = {dst, const_cast<void*>(src)};
  for (int k = 0; k < 2; ++k) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, ends[k]) != cudaSuccess) {
      cudaGetLastError();
      continue;
    }
    D3B_REQUIRE(at.type != cudaMemoryTypeUnregistered, "copy_mapped: %s is pageable host memory (pin it)",
                k == 0 ? "dst" : "src");
    if (at.type == cudaMemoryTypeHost) {
      D3B_REQUIRE(at.devicePointer != nullptr, "copy_mapped: pinned %s has no device alias", k == 0 ? "dst" : "src");
      ends[k] = at.devicePointer;
    }
  }
  dst = ends[0];
  src = ends[1];
  return copy_by_kernel(dst, src, (long long)bytes, (cudaStream_t)stream, "copy_mapped");
}

extern "C" int d3b_copy_d2d(void* dst, const void* src, int64_t bytes, void* stream) {
  D3B_REQUIRE(bytes >= 0, "copy_d2d: bytes < 0");
  if (bytes == 0) return D3B_OK;
  if (bytes <= kSmallCopy) return copy_by_kernel(dst, src, (long long)bytes, (cudaStream_t)stream, "copy_d2d");
  D3B_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  count_launch();
  return D3B_OK;
}

extern "C" int d3b_stream_sync(void* stream) {
  D3B_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return D3B_OK;
}

namespace {
__global__ void spin_kernel(long long ns) {
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  } while ((long long)(t1 - t0) < ns);
}
}  // namespace

// measurement helper: keeps the stream busy for `ns` so that the host can enqueue a whole update behind it and
// per-kernel CUDA-event durations are not inflated by launch starvation (bench.py kernel_profile)
extern "C" int d3b_spin(int64_t ns, void* stream) {
  D3B_REQUIRE(ns >= 0 && ns <= 50000000, "spin: ns out of range");
  spin_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((long long)ns);
  return check_launch("spin");
}

// fork/join of a side stream (also inside stream capture, where they become graph branches): work enqueued on
// `side` between fork and join runs concurrently with the work on `main_stream`.
namespace {
cudaEvent_t g_fork_evt = nullptr, g_join_evt = nullptr;
int ensure_events() {
  if (!g_fork_evt) {
    if (cudaEventCreateWithFlags(&g_fork_evt, cudaEventDisableTiming) != cudaSuccess) return -1;
    if (cudaEventCreateWithFlags(&g_join_evt, cudaEventDisableTiming) != cudaSuccess) return -1;
  }
  return 0;
}
}  // namespace

extern "C" int d3b_stream_fork(void* main_stream, void* side_stream) {
  D3B_REQUIRE(ensure_events() == 0, "stream_fork: cannot create events");
  D3B_CUDA(cudaEventRecord(g_fork_evt, (cudaStream_t)main_stream));
  D3B_CUDA(cudaStreamWaitEvent((cudaStream_t)side_stream, g_fork_evt, 0));
  return D3B_OK;
}

extern "C" int d3b_stream_join(void* main_stream, void* side_stream) {
  D3B_REQUIRE(ensure_events() == 0, "stream_join: cannot create events");
  D3B_CUDA(cudaEventRecord(g_join_evt, (cudaStream_t)side_stream));
  D3B_CUDA(cudaStreamWaitEvent((cudaStream_t)main_stream, g_join_evt, 0));
  return D3B_OK;
}

extern "C" int d3b_graph_begin(void* stream) {
  D3B_CUDA(cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeThreadLocal));
  return D3B_OK;
}

extern "C" int d3b_graph_end(void* stream, void** graph_exec, int* n_nodes) {
  D3B_REQUIRE(graph_exec, "graph_end: null out pointer");
  cudaGraph_t graph = nullptr;
  D3B_CUDA(cudaStreamEndCapture((cudaStream_t)stream, &graph));
  size_t n = 0;
  cudaGraphGetNodes(graph, nullptr, &n);
  if (n_nodes) *n_nodes = (int)n;
  cudaGraphExec_t exec = nullptr;
  cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return set_err(D3B_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
  *graph_exec = (void*)exec;
  return D3B_OK;
}

extern "C" int d3b_graph_launch(void* graph_exec, void* stream) {
  D3B_REQUIRE(graph_exec, "graph_launch: null graph");
  D3B_CUDA(cudaGraphLaunch((cudaGraphExec_t)graph_exec, (cudaStream_t)stream));
  return D3B_OK;
}

extern "C" int d3b_graph_destroy(void* graph_exec) {
  if (graph_exec) D3B_CUDA(cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
  return D3B_OK;
}
