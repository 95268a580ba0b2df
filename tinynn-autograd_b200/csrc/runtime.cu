// Runtime plumbing of libtnn_b200: device context, the stream-ordered caching pool that stands
// in for numpy's implicit malloc/free (every ops.py call allocates its result), host<->device
// copies, events and per-kernel-family profiling.
#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace tnn {

static Context g_ctx;
static thread_local std::string g_err;
static std::string g_err_shared;

Context& ctx() { return g_ctx; }

void set_error(const std::string& msg) {
  g_err = msg;
  g_err_shared = msg;
}

int fail(const char* file, int line, const std::string& msg) {
  char buf[64];
  snprintf(buf, sizeof(buf), ":%d: ", line);
  const char* base = file;
  for (const char* p = file; *p; ++p)
    if (*p == '/') base = p + 1;
  set_error(std::string(base) + buf + msg);
  return 1;
}

// ------------------------------------------------------------------------------------------
// Caching pool.  All device work is queued on one compute stream, so a block released by the
// host can be handed to a later launch without an event: stream order already guarantees the
// earlier kernels that touched it have run.  (The copy stream only ever writes into buffers the
// host has fenced with tnn_copy_wait_compute.)
// ------------------------------------------------------------------------------------------
struct Pool {
  std::mutex mu;
  std::unordered_map<size_t, std::vector<void*>> free_lists;  // rounded size -> blocks
  std::unordered_map<void*, size_t> live;                     // block -> rounded size
  std::unordered_map<void*, size_t> owned;                    // every cudaMalloc'd block
  size_t reserved = 0, in_use = 0, n_malloc = 0;
};
static Pool g_pool;

// A captured step (tnn_graph_*) replays kernels that have the addresses of their temporaries baked
// in.  Every block handed out while the capture runs therefore belongs to the graph: it can be
// reused by later allocations of the SAME capture (stream order inside the graph is the same as
// it was eagerly), but it never goes back to the general free lists until the graph is destroyed.
struct Graph {
  int id = 0;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  std::vector<void*> blocks;                                   // every block the capture touched
  std::unordered_map<size_t, std::vector<void*>> free_lists;   // only meaningful while capturing
  size_t nodes = 0, kernel_nodes = 0;
};
static Graph* g_capture = nullptr;                              // capture in progress, if any
static std::unordered_map<void*, Graph*> g_block_graph;         // block -> owning graph
static int g_graph_ids = 0;
static int g_live_graphs = 0;
static std::vector<void*> g_retired_scratch;                    // old scratch a live graph may still name
static std::vector<Graph*> g_deferred_destroy;                  // destroyed while another capture was running

static size_t round_size(size_t n) {
  if (n == 0) n = 1;
  if (n <= (1u << 20)) return (n + 511) & ~size_t(511);
  const size_t g = size_t(2) << 20;
  return (n + g - 1) / g * g;
}

static void pool_release_all_free_locked() {
  for (auto& kv : g_pool.free_lists) {
    for (void* p : kv.second) {
      cudaFree(p);
      g_pool.reserved -= kv.first;
      g_pool.owned.erase(p);
    }
    kv.second.clear();
  }
}

int get_scratch(size_t nbytes, void** out) {
  Context& c = ctx();
  if (nbytes > c.scratch_bytes) {
    if (g_capture)
      TNN_FAIL("reduction scratch would have to grow during graph capture (run the step once "
               "eagerly first)");
    // the old scratch may still be read by queued kernels: drain before replacing it
    if (c.scratch) {
      TNN_CUDA(cudaStreamSynchronize(c.stream));
      if (g_live_graphs > 0) g_retired_scratch.push_back(c.scratch);  // a captured step may name it
      else TNN_CUDA(cudaFree(c.scratch));
      c.scratch = nullptr;
      c.scratch_bytes = 0;
    }
    size_t want = nbytes < (size_t(1) << 20) ? (size_t(1) << 20) : nbytes * 2;
    TNN_CUDA(cudaMalloc(&c.scratch, want));
    c.scratch_bytes = want;
  }
  *out = c.scratch;
  return 0;
}

// ------------------------------------------------------------------------------------------
// profiling
// ------------------------------------------------------------------------------------------
struct ProfPair {
  cudaEvent_t a, b;
};
static std::vector<ProfPair> g_prof_used;
static std::vector<ProfPair> g_prof_free;
static ProfPair g_prof_cur;

void prof_begin(int family) {
  Context& c = ctx();
  if (c.prof_family != family) return;
  if (g_prof_free.empty()) {
    ProfPair p;
    cudaEventCreate(&p.a);
    cudaEventCreate(&p.b);
    g_prof_free.push_back(p);
  }
  g_prof_cur = g_prof_free.back();
  g_prof_free.pop_back();
  cudaEventRecord(g_prof_cur.a, c.stream);
}

void prof_end(int family) {
  Context& c = ctx();
  if (c.prof_family != family) return;
  cudaEventRecord(g_prof_cur.b, c.stream);
  g_prof_used.push_back(g_prof_cur);
}

}  // namespace tnn

using namespace tnn;

extern "C" {

const char* tnn_last_error(void) { return g_err_shared.c_str(); }

int tnn_device_count(int* n) {
  TNN_CUDA(cudaGetDeviceCount(n));
  return 0;
}

int tnn_init(int device) {
  Context& c = ctx();
  if (c.inited) {
    if (c.device == device) return 0;
    TNN_FAIL("tnn_init: already initialised on another device");
  }
  TNN_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  TNN_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    TNN_FAIL(std::string("libtnn_b200 is built for sm_100a only; device is ") + prop.name);
  c.device = device;
  c.sm_count = prop.multiProcessorCount;
  c.l2_bytes = (size_t)prop.l2CacheSize;
  TNN_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  TNN_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
  TNN_CUDA(cudaStreamCreateWithFlags(&c.comm_stream, cudaStreamNonBlocking));
  TNN_CUDA(cudaStreamCreateWithFlags(&c.d2h_stream, cudaStreamNonBlocking));
  TNN_CUDA(cudaEventCreateWithFlags(&c.ev_comm, cudaEventDisableTiming));
  TNN_CUDA(cudaEventCreateWithFlags(&c.ev_copy, cudaEventDisableTiming));
  TNN_CUDA(cudaEventCreateWithFlags(&c.ev_compute, cudaEventDisableTiming));
  c.inited = true;
  return 0;
}

int tnn_shutdown(void) {
  Context& c = ctx();
  if (!c.inited) return 0;
  cudaStreamSynchronize(c.stream);
  cudaStreamSynchronize(c.copy_stream);
  cudaStreamSynchronize(c.comm_stream);
  cudaStreamSynchronize(c.d2h_stream);
  {
    std::lock_guard<std::mutex> lk(g_pool.mu);
    for (auto& kv : g_pool.owned) cudaFree(kv.first);
    g_pool.owned.clear();
    g_pool.live.clear();
    g_pool.free_lists.clear();
    g_pool.reserved = g_pool.in_use = 0;
    g_block_graph.clear();
  }
  for (void* p : g_retired_scratch) cudaFree(p);
  g_retired_scratch.clear();
  if (c.scratch) cudaFree(c.scratch);
  if (c.l2_flush_buf) cudaFree(c.l2_flush_buf);
  c.scratch = c.l2_flush_buf = nullptr;
  c.scratch_bytes = c.l2_flush_bytes = 0;
  cudaEventDestroy(c.ev_copy);
  cudaEventDestroy(c.ev_compute);
  cudaStreamDestroy(c.stream);
  cudaStreamDestroy(c.copy_stream);
  cudaEventDestroy(c.ev_comm);
  cudaStreamDestroy(c.comm_stream);
  cudaStreamDestroy(c.d2h_stream);
  c.inited = false;
  return 0;
}

int tnn_sync(void) {
  TNN_REQUIRE_INIT();
  if (g_capture) TNN_FAIL("tnn_sync: a graph capture is in progress (nothing is running yet)");
  TNN_CUDA(cudaStreamSynchronize(ctx().stream));
  return 0;
}

int tnn_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem,
                    size_t* l2_bytes) {
  TNN_REQUIRE_INIT();
  cudaDeviceProp prop;
  TNN_CUDA(cudaGetDeviceProperties(&prop, ctx().device));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (total_mem) *total_mem = prop.totalGlobalMem;
  if (l2_bytes) *l2_bytes = (size_t)prop.l2CacheSize;
  return 0;
}

void* tnn_stream(void) { return (void*)ctx().stream; }
uint64_t tnn_launch_count(void) { return ctx().launches; }

// ---- pool ----------------------------------------------------------------------------------
int tnn_alloc(size_t nbytes, void** out) {
  TNN_REQUIRE_INIT();
  size_t r = round_size(nbytes);
  std::lock_guard<std::mutex> lk(g_pool.mu);
  void* p = nullptr;
  if (g_capture) {
    auto git = g_capture->free_lists.find(r);
    if (git != g_capture->free_lists.end() && !git->second.empty()) {
      p = git->second.back();
      git->second.pop_back();
      g_pool.live[p] = r;
      g_pool.in_use += r;
      *out = p;
      return 0;
    }
  }
  auto it = g_pool.free_lists.find(r);
  if (it != g_pool.free_lists.end() && !it->second.empty()) {
    p = it->second.back();
    it->second.pop_back();
  } else {
    cudaError_t e = cudaMalloc(&p, r);
    if (e != cudaSuccess && g_capture) {
      cudaGetLastError();
      TNN_FAIL(std::string("tnn_alloc: cudaMalloc(") + std::to_string(r) +
               ") failed during graph capture: " + cudaGetErrorString(e));
    }
    if (e != cudaSuccess) {
      // out of memory: drop every cached block (after draining the stream) and retry once
      cudaGetLastError();
      cudaStreamSynchronize(ctx().stream);
      pool_release_all_free_locked();
      e = cudaMalloc(&p, r);
      if (e != cudaSuccess) {
        cudaGetLastError();
        TNN_FAIL(std::string("tnn_alloc: cudaMalloc(") + std::to_string(r) +
                 ") failed: " + cudaGetErrorString(e));
      }
    }
    g_pool.owned[p] = r;
    g_pool.reserved += r;
    g_pool.n_malloc++;
  }
  if (g_capture) {
    g_capture->blocks.push_back(p);
    g_block_graph[p] = g_capture;
  }
  g_pool.live[p] = r;
  g_pool.in_use += r;
  *out = p;
  return 0;
}

// same as tnn_alloc, returning the pointer (NULL on failure): one scalar argument, no out-parameter,
// which halves the host cost of the call that every op result goes through
void* tnn_alloc_ptr(size_t nbytes) {
  void* p = nullptr;
  if (tnn_alloc(nbytes, &p)) return nullptr;
  return p;
}

int tnn_free(void* p) {
  if (!p) return 0;
  std::lock_guard<std::mutex> lk(g_pool.mu);
  auto it = g_pool.live.find(p);
  if (it == g_pool.live.end()) {
    if (!ctx().inited) return 0;  // after shutdown everything is already gone
    TNN_FAIL("tnn_free: pointer was not allocated by tnn_alloc (or double free)");
  }
  size_t r = it->second;
  g_pool.live.erase(it);
  g_pool.in_use -= r;
  if (!g_block_graph.empty()) {
    auto og = g_block_graph.find(p);
    if (og != g_block_graph.end()) {
      // owned by a captured step: reusable inside that capture only, otherwise parked until
      // tnn_graph_destroy
      if (og->second == g_capture) g_capture->free_lists[r].push_back(p);
      return 0;
    }
  }
  g_pool.free_lists[r].push_back(p);
  return 0;
}

int tnn_pool_stats(size_t* reserved_bytes, size_t* in_use_bytes, size_t* n_cuda_malloc) {
  std::lock_guard<std::mutex> lk(g_pool.mu);
  if (reserved_bytes) *reserved_bytes = g_pool.reserved;
  if (in_use_bytes) *in_use_bytes = g_pool.in_use;
  if (n_cuda_malloc) *n_cuda_malloc = g_pool.n_malloc;
  return 0;
}

int tnn_pool_trim(void) {
  TNN_REQUIRE_INIT();
  if (g_capture) TNN_FAIL("tnn_pool_trim: a graph capture is in progress");
  TNN_CUDA(cudaStreamSynchronize(ctx().stream));
  std::lock_guard<std::mutex> lk(g_pool.mu);
  pool_release_all_free_locked();
  return 0;
}

// ---- copies --------------------------------------------------------------------------------
int tnn_h2d(void* dst, const void* src, size_t nbytes) {
  TNN_REQUIRE_INIT();
  if (nbytes == 0) return 0;
  // a captured copy would re-read this (by then dangling) host address at every replay
  if (g_capture) TNN_FAIL("tnn_h2d: host upload inside a graph capture");
  TNN_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyHostToDevice, ctx().stream));
  return 0;
}

int tnn_d2h(void* dst, const void* src, size_t nbytes) {
  TNN_REQUIRE_INIT();
  if (nbytes == 0) return 0;
  if (g_capture) TNN_FAIL("tnn_d2h: host read-back inside a graph capture (nothing has run yet)");
  TNN_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToHost, ctx().stream));
  TNN_CUDA(cudaStreamSynchronize(ctx().stream));
  return 0;
}

int tnn_d2h_async(void* pinned_dst, const void* src, size_t nbytes, void* done_event) {
  TNN_REQUIRE_INIT();
  if (g_capture) TNN_FAIL("tnn_d2h_async: host read-back inside a graph capture (nothing has run yet)");
  Context& c = ctx();
  // ordered after what is queued on the compute stream NOW; later compute work does not delay it
  TNN_CUDA(cudaEventRecord((cudaEvent_t)done_event, c.stream));
  TNN_CUDA(cudaStreamWaitEvent(c.d2h_stream, (cudaEvent_t)done_event, 0));
  if (nbytes) TNN_CUDA(cudaMemcpyAsync(pinned_dst, src, nbytes, cudaMemcpyDeviceToHost, c.d2h_stream));
  TNN_CUDA(cudaEventRecord((cudaEvent_t)done_event, c.d2h_stream));
  return 0;
}

int tnn_event_sync(void* ev) {
  TNN_REQUIRE_INIT();
  TNN_CUDA(cudaEventSynchronize((cudaEvent_t)ev));
  return 0;
}

int tnn_d2d(void* dst, const void* src, size_t nbytes) {
  TNN_REQUIRE_INIT();
  if (nbytes == 0) return 0;
  TNN_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToDevice, ctx().stream));
  return 0;
}

int tnn_memset(void* dst, int byte, size_t nbytes) {
  TNN_REQUIRE_INIT();
  if (nbytes == 0) return 0;
  TNN_CUDA(cudaMemsetAsync(dst, byte, nbytes, ctx().stream));
  return 0;
}

int tnn_host_alloc(size_t nbytes, void** out) {
  TNN_REQUIRE_INIT();
  TNN_CUDA(cudaHostAlloc(out, nbytes ? nbytes : 1, cudaHostAllocDefault));
  return 0;
}

int tnn_host_free(void* p) {
  if (!p) return 0;
  TNN_CUDA(cudaFreeHost(p));
  return 0;
}

// pin / unpin caller-owned host memory in place, so batches can be DMA'd straight out of a numpy
// data set without a staging copy
int tnn_host_register(void* p, size_t nbytes) {
  TNN_REQUIRE_INIT();
  TNN_CUDA(cudaHostRegister(p, nbytes, cudaHostRegisterDefault));
  return 0;
}

int tnn_host_unregister(void* p) {
  if (!p || !ctx().inited) return 0;
  TNN_CUDA(cudaHostUnregister(p));
  return 0;
}

int tnn_h2d_async_copy_stream(void* dst, const void* pinned_src, size_t nbytes) {
  TNN_REQUIRE_INIT();
  if (nbytes == 0) return 0;
  TNN_CUDA(cudaMemcpyAsync(dst, pinned_src, nbytes, cudaMemcpyHostToDevice, ctx().copy_stream));
  return 0;
}

int tnn_copy_stream_sync(void) {
  TNN_REQUIRE_INIT();
  TNN_CUDA(cudaStreamSynchronize(ctx().copy_stream));
  return 0;
}

int tnn_copy_wait_compute(void) {
  TNN_REQUIRE_INIT();
  Context& c = ctx();
  TNN_CUDA(cudaEventRecord(c.ev_compute, c.stream));
  TNN_CUDA(cudaStreamWaitEvent(c.copy_stream, c.ev_compute, 0));
  return 0;
}

int tnn_compute_wait_copy(void) {
  TNN_REQUIRE_INIT();
  Context& c = ctx();
  TNN_CUDA(cudaEventRecord(c.ev_copy, c.copy_stream));
  TNN_CUDA(cudaStreamWaitEvent(c.stream, c.ev_copy, 0));
  return 0;
}

// ---- events --------------------------------------------------------------------------------
int tnn_event_create(void** ev) {
  TNN_REQUIRE_INIT();
  cudaEvent_t e;
  TNN_CUDA(cudaEventCreate(&e));
  *ev = (void*)e;
  return 0;
}

int tnn_event_destroy(void* ev) {
  if (ev) TNN_CUDA(cudaEventDestroy((cudaEvent_t)ev));
  return 0;
}

int tnn_event_record(void* ev) {
  TNN_REQUIRE_INIT();
  TNN_CUDA(cudaEventRecord((cudaEvent_t)ev, ctx().stream));
  return 0;
}

int tnn_event_elapsed_ms(void* start, void* stop, float* ms) {
  TNN_REQUIRE_INIT();
  TNN_CUDA(cudaEventSynchronize((cudaEvent_t)stop));
  TNN_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
  return 0;
}

int tnn_prof_enable(int family) {
  TNN_REQUIRE_INIT();
  ctx().prof_family = family;
  return 0;
}

int tnn_prof_collect(double* total_ms, uint64_t* n_launches) {
  TNN_REQUIRE_INIT();
  TNN_CUDA(cudaStreamSynchronize(ctx().stream));
  double tot = 0.0;
  for (auto& p : g_prof_used) {
    float ms = 0.f;
    TNN_CUDA(cudaEventElapsedTime(&ms, p.a, p.b));
    tot += ms;
    g_prof_free.push_back(p);
  }
  if (total_ms) *total_ms = tot;
  if (n_launches) *n_launches = g_prof_used.size();
  g_prof_used.clear();
  return 0;
}

static void drain_deferred_destroys();

// ---- captured steps (CUDA graphs) --------------------------------------------------------------
// The MNIST-shaped training step is ~35 launches of a few microseconds each: issued one by one
// from Python it is bound by launch overhead.  tnn_graph_begin/end record everything the host
// queues on the compute stream in between -- kernels, memsets, device copies, NCCL collectives --
// into a CUDA graph; tnn_graph_launch replays the whole step with one call.
int tnn_graph_begin(void) {
  TNN_REQUIRE_INIT();
  Context& c = ctx();
  if (g_capture) TNN_FAIL("tnn_graph_begin: a capture is already in progress");
  if (c.prof_family) TNN_FAIL("tnn_graph_begin: per-kernel profiling is on (tnn_prof_enable(0) first)");
  Graph* g = new Graph();
  g->id = ++g_graph_ids;
  cudaError_t e = cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeRelaxed);
  if (e != cudaSuccess) {
    delete g;
    TNN_FAIL(std::string("cudaStreamBeginCapture: ") + cudaGetErrorString(e));
  }
  std::lock_guard<std::mutex> lk(g_pool.mu);
  g_capture = g;
  return 0;
}

static void graph_release_blocks_locked(Graph* g) {
  for (void* p : g->blocks) {
    g_block_graph.erase(p);
    if (g_pool.live.find(p) != g_pool.live.end()) continue;  // still held by the host: now ordinary
    auto ow = g_pool.owned.find(p);
    if (ow != g_pool.owned.end()) g_pool.free_lists[ow->second].push_back(p);
  }
  g->blocks.clear();
  g->free_lists.clear();
}

int tnn_graph_end(void** graph_out) {
  TNN_REQUIRE_INIT();
  Context& c = ctx();
  if (!g_capture) TNN_FAIL("tnn_graph_end: no capture in progress");
  Graph* g = g_capture;
  cudaError_t e = cudaStreamEndCapture(c.stream, &g->graph);
  {
    std::lock_guard<std::mutex> lk(g_pool.mu);
    g_capture = nullptr;
    g->free_lists.clear();
  }
  if (e == cudaSuccess) e = cudaGraphInstantiate(&g->exec, g->graph, 0);
  if (e != cudaSuccess) {
    cudaGetLastError();
    if (g->graph) cudaGraphDestroy(g->graph);
    {
      std::lock_guard<std::mutex> lk(g_pool.mu);
      graph_release_blocks_locked(g);
    }
    delete g;
    drain_deferred_destroys();
    TNN_FAIL(std::string("graph capture failed: ") + cudaGetErrorString(e));
  }
  size_t n = 0;
  cudaGraphGetNodes(g->graph, nullptr, &n);
  g->nodes = n;
  if (n) {
    std::vector<cudaGraphNode_t> nodes(n);
    cudaGraphGetNodes(g->graph, nodes.data(), &n);
    for (auto nd : nodes) {
      cudaGraphNodeType t;
      if (cudaGraphNodeGetType(nd, &t) == cudaSuccess && t == cudaGraphNodeTypeKernel) g->kernel_nodes++;
    }
  }
  g_live_graphs++;
  *graph_out = (void*)g;
  drain_deferred_destroys();
  return 0;
}

// abandon a capture in progress (the host raised in the middle of the step)
int tnn_graph_abort(void) {
  if (!g_capture) return 0;
  Graph* g = g_capture;
  cudaGraph_t tmp = nullptr;
  cudaStreamEndCapture(ctx().stream, &tmp);
  cudaGetLastError();
  if (tmp) cudaGraphDestroy(tmp);
  {
    std::lock_guard<std::mutex> lk(g_pool.mu);
    g_capture = nullptr;
    graph_release_blocks_locked(g);
  }
  delete g;
  drain_deferred_destroys();
  return 0;
}

int tnn_graph_launch(void* graph) {
  TNN_REQUIRE_INIT();
  Graph* g = (Graph*)graph;
  if (!g || !g->exec) TNN_FAIL("tnn_graph_launch: bad graph handle");
  TNN_CUDA(cudaGraphLaunch(g->exec, ctx().stream));
  ctx().launches += g->kernel_nodes;
  return 0;
}

int tnn_graph_info(void* graph, size_t* n_nodes, size_t* n_kernel_nodes, size_t* n_blocks) {
  Graph* g = (Graph*)graph;
  if (!g) TNN_FAIL("tnn_graph_info: bad graph handle");
  if (n_nodes) *n_nodes = g->nodes;
  if (n_kernel_nodes) *n_kernel_nodes = g->kernel_nodes;
  if (n_blocks) *n_blocks = g->blocks.size();
  return 0;
}

static int graph_destroy_now(Graph* g);

// graphs whose owner let go of them in the middle of another capture (Python's garbage collector
// runs whenever it likes): destroyed once the stream may be synchronised again
static void drain_deferred_destroys() {
  std::vector<Graph*> todo;
  todo.swap(g_deferred_destroy);
  for (Graph* g : todo) graph_destroy_now(g);
}

int tnn_graph_destroy(void* graph) {
  Graph* g = (Graph*)graph;
  if (!g) return 0;
  if (!ctx().inited) {  // after shutdown everything is already gone
    delete g;
    return 0;
  }
  if (g_capture) {      // synchronising the stream now would invalidate the capture in progress
    g_deferred_destroy.push_back(g);
    return 0;
  }
  return graph_destroy_now(g);
}

static int graph_destroy_now(Graph* g) {
  cudaStreamSynchronize(ctx().stream);  // a replay may still be running
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  {
    std::lock_guard<std::mutex> lk(g_pool.mu);
    graph_release_blocks_locked(g);
  }
  delete g;
  if (--g_live_graphs == 0) {
    for (void* p : g_retired_scratch) cudaFree(p);
    g_retired_scratch.clear();
  }
  return 0;
}

__global__ void l2_flush_kernel(float4* buf, size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) buf[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

int tnn_l2_flush(void) {
  TNN_REQUIRE_INIT();
  Context& c = ctx();
  if (!c.l2_flush_buf) {
    size_t want = c.l2_bytes ? c.l2_bytes * 2 : (size_t(256) << 20);
    TNN_CUDA(cudaMalloc(&c.l2_flush_buf, want));
    c.l2_flush_bytes = want;
  }
  size_t n4 = c.l2_flush_bytes / sizeof(float4);
  l2_flush_kernel<<<c.sm_count * 8, 256, 0, c.stream>>>((float4*)c.l2_flush_buf, n4);
  TNN_POST_LAUNCH();
  return 0;
}

}  // extern "C"
