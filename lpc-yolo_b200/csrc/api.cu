// api.cu - error string, version and device probes of the C ABI (include/lpcyolo.h).
#include <cstdlib>
#include <stdarg.h>

#include <atomic>
#include <vector>

#include "common.cuh"

static thread_local char g_err[512] = "";

void lpc_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool lpc_pdl_enabled() {
  static const bool on = [] { const char* e = getenv("LPC_PDL"); return !(e && e[0] == '0'); }();
  return on;
}

static std::atomic<unsigned long long> g_launches{0};
void lpc_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" unsigned long long lpc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" const char* lpc_last_error(void) { return g_err; }
extern "C" int lpc_abi_version(void) { return LPC_ABI_VERSION; }

extern "C" int lpc_device_arch(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) LPC_FAIL(LPC_E_CUDA, "device_arch: no CUDA device");
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  return major * 10 + minor;
}

// ---- launch plans: record a step once, replay it from C ---------------------------------------------------------------
// The reference has no native executor (its layer loop is Python, nn/tasks.py:83-111); SURVEY.md section 8(b) proposes a
// plan-level entry next to the per-op ones.  A plan here is the recorded launch sequence of whatever the host code issued
// between lpc_plan_begin() and lpc_plan_end() on the calling thread (e.g. one YOLOv10DetectionModel.detect call): every
// kernel with its grid, shared memory and argument block, the stream it was issued on, and the cross-stream dependencies the
// host code declares through lpc_plan_wait (its side-stream forks and joins).  lpc_plan_run re-issues the sequence with the
// same stream structure (the first recorded stream maps to the caller's stream, the others to streams the plan owns);
// lpc_plan_run_graph captures that into a CUDA graph on its first call and launches the graph afterwards.  All device buffers the recorded step used must still be alive and at the
// same addresses (the caller records inside a private memory pool); the plan owns only host-side copies of the arguments.
struct lpc_plan_op {
  int stream;                 // index of the stream the op was recorded on (0 = the first stream seen = the replay stream)
  int wait_for;               // >= 0: no launch - stream `stream` waits for everything recorded so far on stream `wait_for`
  std::function<cudaError_t(cudaStream_t)> launch;
};
struct lpc_plan {
  std::vector<lpc_plan_op> ops;
  std::vector<cudaStream_t> recorded;      // stream handles seen while recording (index = op.stream)
  std::vector<cudaStream_t> side;          // replay streams for indices >= 1
  std::vector<cudaEvent_t> events;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  cudaStream_t capture_stream = nullptr;
  int stream_index(cudaStream_t s) {
    for (size_t i = 0; i < recorded.size(); ++i)
      if (recorded[i] == s) return (int)i;
    recorded.push_back(s);
    return (int)recorded.size() - 1;
  }
};
static thread_local lpc_plan* g_rec = nullptr;
bool lpc_plan_recording() { return g_rec != nullptr; }
void lpc_plan_push(cudaStream_t recorded_on, std::function<cudaError_t(cudaStream_t)> op) {
  if (g_rec) g_rec->ops.push_back({g_rec->stream_index(recorded_on), -1, std::move(op)});
}

// Stream dependencies of the recorded step (the host code forks independent chains onto side streams: functional.fork_join):
// "stream `waiter` waits for everything issued so far on `signaler`".  Outside a recording this is a no-op.
extern "C" int lpc_plan_wait(void* waiter, void* signaler) {
  if (!g_rec) return LPC_OK;
  const int w = g_rec->stream_index((cudaStream_t)waiter), sg = g_rec->stream_index((cudaStream_t)signaler);
  g_rec->ops.push_back({w, sg, nullptr});
  return LPC_OK;
}

static int plan_issue(lpc_plan* p, cudaStream_t primary) {
  while (p->side.size() + 1 < p->recorded.size()) {
    cudaStream_t st;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) LPC_FAIL(LPC_E_CUDA, "plan: side stream");
    p->side.push_back(st);
  }
  auto map = [&](int i) { return i == 0 ? primary : p->side[i - 1]; };
  size_t ev = 0;
  for (size_t i = 0; i < p->ops.size(); ++i) {
    const lpc_plan_op& op = p->ops[i];
    if (op.wait_for >= 0) {
      if (ev == p->events.size()) {
        cudaEvent_t e;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) LPC_FAIL(LPC_E_CUDA, "plan: event");
        p->events.push_back(e);
      }
      cudaError_t e = cudaEventRecord(p->events[ev], map(op.wait_for));
      if (e == cudaSuccess) e = cudaStreamWaitEvent(map(op.stream), p->events[ev], 0);
      ++ev;
      if (e != cudaSuccess) LPC_FAIL(LPC_E_CUDA, "plan: stream dependency %zu: %s", i, cudaGetErrorString(e));
      continue;
    }
    const cudaError_t e = op.launch(map(op.stream));
    if (e != cudaSuccess) LPC_FAIL(LPC_E_CUDA, "plan: launch %zu of %zu: %s", i, p->ops.size(), cudaGetErrorString(e));
  }
  return LPC_OK;
}

extern "C" int lpc_plan_begin(void) {
  if (g_rec) LPC_FAIL(LPC_E_ARG, "plan_begin: this thread is already recording");
  g_rec = new lpc_plan();
  return LPC_OK;
}

extern "C" int lpc_plan_end(lpc_plan** out) {
  if (!g_rec) LPC_FAIL(LPC_E_ARG, "plan_end: this thread is not recording");
  lpc_plan* p = g_rec;
  g_rec = nullptr;
  if (!out) {
    delete p;
    LPC_FAIL(LPC_E_ARG, "plan_end: null output");
  }
  *out = p;
  return LPC_OK;
}

extern "C" int lpc_plan_size(const lpc_plan* p) {
  if (!p) return LPC_E_ARG;
  int n = 0;
  for (const lpc_plan_op& op : p->ops) n += op.wait_for < 0;
  return n;
}

extern "C" int lpc_plan_run(lpc_plan* p, void* stream) {
  LPC_REQUIRE(p, "plan_run: null plan");
  LPC_REQUIRE(!g_rec, "plan_run: cannot replay while recording");
  return plan_issue(p, (cudaStream_t)stream);
}

extern "C" int lpc_plan_run_graph(lpc_plan* p, void* stream) {
  LPC_REQUIRE(p, "plan_run_graph: null plan");
  LPC_REQUIRE(!g_rec, "plan_run_graph: cannot replay while recording");
  if (!p->exec) {
    if (cudaStreamCreateWithFlags(&p->capture_stream, cudaStreamNonBlocking) != cudaSuccess) LPC_FAIL(LPC_E_CUDA, "plan_run_graph: stream");
    cudaError_t e = cudaStreamBeginCapture(p->capture_stream, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) LPC_FAIL(LPC_E_CUDA, "plan_run_graph: begin capture: %s", cudaGetErrorString(e));
    const int rc = plan_issue(p, p->capture_stream);
    e = cudaStreamEndCapture(p->capture_stream, &p->graph);
    if (rc != LPC_OK || e != cudaSuccess) {
      if (p->graph) cudaGraphDestroy(p->graph);
      p->graph = nullptr;
      if (rc != LPC_OK) return rc;
      LPC_FAIL(LPC_E_CUDA, "plan_run_graph: capture: %s", cudaGetErrorString(e));
    }
    e = cudaGraphInstantiate(&p->exec, p->graph, 0);
    if (e != cudaSuccess) LPC_FAIL(LPC_E_CUDA, "plan_run_graph: instantiate: %s", cudaGetErrorString(e));
  }
  const cudaError_t e = cudaGraphLaunch(p->exec, (cudaStream_t)stream);
  if (e != cudaSuccess) LPC_FAIL(LPC_E_CUDA, "plan_run_graph: launch: %s", cudaGetErrorString(e));
  return LPC_OK;
}

extern "C" void lpc_plan_destroy(lpc_plan* p) {
  if (!p) return;
  if (p->exec) cudaGraphExecDestroy(p->exec);
  if (p->graph) cudaGraphDestroy(p->graph);
  if (p->capture_stream) cudaStreamDestroy(p->capture_stream);
  for (cudaStream_t st : p->side) cudaStreamDestroy(st);
  for (cudaEvent_t e : p->events) cudaEventDestroy(e);
  delete p;
}
