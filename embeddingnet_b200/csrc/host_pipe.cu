// Host-buffer batch-hard step: the call a reference-side training loop makes with host arrays (NumPy / pinned
// staging buffers), i.e. the loss callable of embedding_net/losses_and_accuracies.py:14-44 applied to a batch that
// lives in host memory, returning the loss and d loss / d embeddings to host memory.
//
// A pipe owns `depth` slots of device buffers (carved from the caller's device block), three streams (host->device |
// compute | device->host) and, per slot, a CUDA graph of the en_batch_hard_fwd_bwd kernels captured once at creation
// (device addresses of a slot never change, so neither do its tensor maps).  Steps in different slots overlap: step
// i+1's upload and step i-1's download run under step i's kernels, which is what keeps the PCIe link -- the bound of
// this call at 8.4 MB each way per 65 us of compute -- busy in both directions.
//
// The hand-offs upload -> kernels -> download are made by the CALLING thread inside submit() / wait() (it waits for
// the slot's event and then launches the next stage); no stream ever waits on another stream's event.  Measured on
// the B200 box: with cudaStreamWaitEvent between the three streams a step takes 0.242 ms however the schedule is
// expressed (tools/pipe_probe.cu: three streams, a stream per slot, lock step, chunked copies, a flag-polling
// kernel: 0.217 - 0.233 ms with a 70 us sleep kernel in place of the step), with the hand-offs on the host 0.205 -
// 0.213 ms (probe: 0.198; the copies alone, both directions at once: 0.176).  A helper thread polling the events
// instead of the calling thread was slower than either (0.262 - 0.267 ms, tools/time_pipe.py A/B on one box).
#include <new>
#include "common.cuh"

namespace en {
namespace {

constexpr int kMaxDepth = 8;

struct Slot {
  float* emb = nullptr;
  int32_t* labels = nullptr;
  float* grad = nullptr;
  float* loss = nullptr;
  int32_t *hp_idx = nullptr, *hn_idx = nullptr;
  float *hp = nullptr, *hn = nullptr, *coef = nullptr;
  void* ws = nullptr;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  cudaEvent_t in_done = nullptr, cmp_done = nullptr, out_done = nullptr;
  // host destinations of the step currently in this slot
  float* h_loss = nullptr;
  float* h_grad = nullptr;
  int32_t *h_hp = nullptr, *h_hn = nullptr;
};

struct Pipe {
  int64_t B = 0;
  int d = 0, depth = 0, device = 0;
  size_t ws_bytes = 0;
  int64_t next_ticket = 0;
  int64_t launched = 0;      // tickets whose kernels have been launched
  int64_t downloading = 0;   // tickets whose download has been issued
  int64_t kernels_per_step = 0;
  cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
  Slot slot[kMaxDepth];
};

size_t slot_bytes(int64_t B, int d) {
  const size_t row = static_cast<size_t>(B) * static_cast<size_t>(d) * sizeof(float);
  return 2 * align_up(row) + 6 * align_up(static_cast<size_t>(B) * 4) + align_up(sizeof(float)) +
         align_up(en_ws_bytes_batch_hard(B, d));
}

void destroy(Pipe* p) {
  if (!p) return;
  for (int k = 0; k < p->depth; ++k) {
    Slot& s = p->slot[k];
    if (s.exec) cudaGraphExecDestroy(s.exec);
    if (s.graph) cudaGraphDestroy(s.graph);
    if (s.in_done) cudaEventDestroy(s.in_done);
    if (s.cmp_done) cudaEventDestroy(s.cmp_done);
    if (s.out_done) cudaEventDestroy(s.out_done);
  }
  if (p->s_in) cudaStreamDestroy(p->s_in);
  if (p->s_cmp) cudaStreamDestroy(p->s_cmp);
  if (p->s_out) cudaStreamDestroy(p->s_out);
  delete p;
}

}  // namespace
}  // namespace en

using namespace en;

extern "C" {

size_t en_bh_host_pipe_device_bytes(int64_t B, int d, int depth) {
  if (B <= 0 || d <= 0 || depth < 1 || depth > kMaxDepth) return 0;
  return static_cast<size_t>(depth) * slot_bytes(B, d);
}

int en_bh_host_pipe_create(int64_t B, int d, float margin, int squared, int soft, int depth, void* device_mem,
                           size_t device_bytes, void** pipe_out) {
  EN_REQUIRE(pipe_out != nullptr, "en_bh_host_pipe_create: pipe_out is null");
  *pipe_out = nullptr;
  EN_REQUIRE(B > 0 && d > 0 && depth >= 1 && depth <= kMaxDepth,
             "en_bh_host_pipe_create: bad arguments (B=%lld d=%d depth=%d, depth must be 1..%d)", (long long)B, d,
             depth, kMaxDepth);
  if (int rc = check_sm100()) return rc;
  if (!device_mem || device_bytes < en_bh_host_pipe_device_bytes(B, d, depth) ||
      (reinterpret_cast<uintptr_t>(device_mem) & 255) != 0)
    return fail(EN_ERR_WORKSPACE, "en_bh_host_pipe_create: device block too small or misaligned (%zu < %zu)",
                device_bytes, en_bh_host_pipe_device_bytes(B, d, depth));
  Pipe* p = new (std::nothrow) Pipe();
  EN_REQUIRE(p != nullptr, "en_bh_host_pipe_create: out of host memory");
  p->B = B;
  p->d = d;
  p->depth = depth;
  p->ws_bytes = en_ws_bytes_batch_hard(B, d);
  cudaError_t e = cudaGetDevice(&p->device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_cmp, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    destroy(p);
    return cuda_fail(e, "en_bh_host_pipe_create: streams");
  }
  Workspace w(device_mem, device_bytes);
  const int64_t launches_before = launch_counter();
  for (int k = 0; k < depth; ++k) {
    Slot& s = p->slot[k];
    s.emb = w.take<float>(static_cast<size_t>(B) * d);
    s.grad = w.take<float>(static_cast<size_t>(B) * d);
    s.labels = w.take<int32_t>(B);
    s.hp_idx = w.take<int32_t>(B);
    s.hn_idx = w.take<int32_t>(B);
    s.hp = w.take<float>(B);
    s.hn = w.take<float>(B);
    s.coef = w.take<float>(B);
    s.loss = w.take<float>(1);
    s.ws = w.take<uint8_t>(p->ws_bytes);
    for (cudaEvent_t* ev : {&s.in_done, &s.cmp_done, &s.out_done}) {
      e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
      if (e != cudaSuccess) {
        destroy(p);
        return cuda_fail(e, "en_bh_host_pipe_create: events");
      }
    }
    // the slot's step as a graph: operand split (+ gradient / counter zeroing), distance GEMM, two finalize kernels
    e = cudaStreamBeginCapture(p->s_cmp, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) {
      destroy(p);
      return cuda_fail(e, "en_bh_host_pipe_create: cudaStreamBeginCapture");
    }
    const int rc = en_batch_hard_fwd_bwd(s.emb, s.labels, B, d, margin, squared, soft, s.loss, s.hp_idx, s.hn_idx, s.hp,
                                         s.hn, s.coef, nullptr, s.grad, s.ws, p->ws_bytes, p->s_cmp);
    e = cudaStreamEndCapture(p->s_cmp, &s.graph);
    if (rc != EN_OK) {
      destroy(p);
      return rc;  // en_last_error() holds the step's message
    }
    if (e == cudaSuccess) e = cudaGraphInstantiate(&s.exec, s.graph, 0);
    if (e != cudaSuccess) {
      destroy(p);
      return cuda_fail(e, "en_bh_host_pipe_create: graph capture / instantiate");
    }
    if (k == 0) p->kernels_per_step = launch_counter() - launches_before;
  }
  launch_counter() = launches_before;  // capturing launched nothing
  if (!w.ok()) {
    destroy(p);
    return fail(EN_ERR_WORKSPACE, "en_bh_host_pipe_create: device block too small");
  }
  *pipe_out = p;
  return EN_OK;
}

// The stage hand-offs are made by the CALLING thread (no stream ever waits on another stream's event: see the file
// header): launch the kernels of every ticket below `launch_upto` (each once its upload has landed) and issue the
// download of every ticket below `download_upto` (each once its kernels have finished).
static int drive(Pipe* p, int64_t launch_upto, int64_t download_upto) {
  const size_t row_bytes = static_cast<size_t>(p->B) * p->d * sizeof(float), b4 = static_cast<size_t>(p->B) * 4;
  if (launch_upto < download_upto) launch_upto = download_upto;
  while (p->launched < launch_upto || p->downloading < download_upto) {
    if (p->launched < launch_upto) {
      Slot& s = p->slot[p->launched % p->depth];
      EN_CUDA(cudaEventSynchronize(s.in_done));
      EN_CUDA(cudaGraphLaunch(s.exec, p->s_cmp));
      EN_CUDA(cudaEventRecord(s.cmp_done, p->s_cmp));
      ++p->launched;
    }
    if (p->downloading < download_upto && p->downloading < p->launched) {
      Slot& s = p->slot[p->downloading % p->depth];
      EN_CUDA(cudaEventSynchronize(s.cmp_done));
      EN_CUDA(cudaMemcpyAsync(s.h_grad, s.grad, row_bytes, cudaMemcpyDeviceToHost, p->s_out));
      if (s.h_hp) EN_CUDA(cudaMemcpyAsync(s.h_hp, s.hp_idx, b4, cudaMemcpyDeviceToHost, p->s_out));
      if (s.h_hn) EN_CUDA(cudaMemcpyAsync(s.h_hn, s.hn_idx, b4, cudaMemcpyDeviceToHost, p->s_out));
      EN_CUDA(cudaMemcpyAsync(s.h_loss, s.loss, sizeof(float), cudaMemcpyDeviceToHost, p->s_out));
      EN_CUDA(cudaEventRecord(s.out_done, p->s_out));
      ++p->downloading;
    }
  }
  return EN_OK;
}

int en_bh_host_pipe_submit(void* pipe, const float* emb_host, const int32_t* labels_host, float* loss_host,
                           float* grad_host, int32_t* hp_idx_host, int32_t* hn_idx_host, int64_t* ticket_out) {
  Pipe* p = static_cast<Pipe*>(pipe);
  EN_REQUIRE(p && emb_host && labels_host && loss_host && grad_host, "en_bh_host_pipe_submit: null argument");
  const int64_t t = p->next_ticket;
  Slot& s = p->slot[t % p->depth];
  const size_t row_bytes = static_cast<size_t>(p->B) * p->d * sizeof(float), b4 = static_cast<size_t>(p->B) * 4;
  // the step that used this slot `depth` submits ago must have delivered its results before the slot is reused
  if (t >= p->depth) {
    if (int rc = drive(p, t - p->depth + 1, t - p->depth + 1)) return rc;
    EN_CUDA(cudaEventSynchronize(s.out_done));
  }
  s.h_loss = loss_host;
  s.h_grad = grad_host;
  s.h_hp = hp_idx_host;
  s.h_hn = hn_idx_host;
  EN_CUDA(cudaMemcpyAsync(s.emb, emb_host, row_bytes, cudaMemcpyHostToDevice, p->s_in));
  EN_CUDA(cudaMemcpyAsync(s.labels, labels_host, b4, cudaMemcpyHostToDevice, p->s_in));
  EN_CUDA(cudaEventRecord(s.in_done, p->s_in));
  launch_counter() += p->kernels_per_step;
  p->next_ticket = t + 1;
  if (ticket_out) *ticket_out = t;
  // with this step's upload queued: start the previous step's kernels (its upload has had a whole upload time) and
  // the download of the step before that
  return drive(p, t, t - 1);
}

int en_bh_host_pipe_wait(void* pipe, int64_t ticket) {
  Pipe* p = static_cast<Pipe*>(pipe);
  EN_REQUIRE(p != nullptr, "en_bh_host_pipe_wait: null pipe");
  EN_REQUIRE(ticket >= 0 && ticket < p->next_ticket, "en_bh_host_pipe_wait: ticket %lld was never issued",
             (long long)ticket);
  // slot already reused: the submit that reused it waited for this step's download first (and the event now
  // belongs to the later step)
  if (ticket + p->depth <= p->next_ticket) return EN_OK;
  if (int rc = drive(p, ticket + 1, ticket + 1)) return rc;
  EN_CUDA(cudaEventSynchronize(p->slot[ticket % p->depth].out_done));
  return EN_OK;
}

int en_bh_host_pipe_destroy(void* pipe) {
  Pipe* p = static_cast<Pipe*>(pipe);
  if (!p) return EN_OK;
  int rc = drive(p, p->next_ticket, p->next_ticket);  // everything submitted is carried through
  cudaError_t e = cudaSuccess;
  if (p->s_out) e = cudaStreamSynchronize(p->s_out);  // results of every submitted step are in host memory
  destroy(p);
  if (rc) return rc;
  if (e != cudaSuccess) return cuda_fail(e, "en_bh_host_pipe_destroy");
  return EN_OK;
}

}  // extern "C"
