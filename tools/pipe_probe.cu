// Probe (developer aid): how do 8.4 MB up + 70 us of kernel + 8.4 MB down per step overlap on this box, for several
// ways of expressing the schedule?  nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/pipe_probe.cu -o build/pipe_probe
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void sleep_kernel(unsigned ns_total) {
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 >= ns_total) break;
    __nanosleep(1000);
  }
}

// waits until the copy engine has delivered generation `want` of the flag word (written by a 4-byte H2D copy queued
// behind the data copy), then sleeps: the upload -> kernel dependency without any stream event
__global__ void wait_flag_then_sleep(const volatile int* flag, int want, unsigned ns_total) {
  while (*flag < want) __nanosleep(200);
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 >= ns_total) break;
    __nanosleep(1000);
  }
}

// grid-stride 16-byte copy (zero-copy upload / download when one side is mapped host memory)
__global__ void copy_kernel(const float4* __restrict__ src, float4* __restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(src + i));
    dst[i] = v;
  }
}

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
  const size_t bytes = 4096ull * 512 * 4;
  const int DEPTH = argc > 1 ? atoi(argv[1]) : 3;
  const int N = 300;
  float *h_in, *h_out[8], *d_in[8], *d_out[8];
  CK(cudaHostAlloc(&h_in, bytes, cudaHostAllocMapped));
  for (int k = 0; k < DEPTH; ++k) {
    CK(cudaHostAlloc(&h_out[k], bytes, cudaHostAllocMapped));
    CK(cudaMalloc(&d_in[k], bytes));
    CK(cudaMalloc(&d_out[k], bytes));
  }
  cudaStream_t s_in, s_cmp, s_out, s_slot[8];
  CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&s_cmp, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
  cudaEvent_t in_done[8], cmp_done[8], out_done[8];
  for (int k = 0; k < DEPTH; ++k) {
    CK(cudaStreamCreateWithFlags(&s_slot[k], cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&in_done[k], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&cmp_done[k], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&out_done[k], cudaEventDisableTiming));
  }
  const size_t n16 = bytes / 16;
  int* h_gen;   // pinned generation words, one per step (each is copied once)
  int* d_flag;
  CK(cudaHostAlloc(&h_gen, 4096 * sizeof(int), cudaHostAllocDefault));
  for (int i = 0; i < 4096; ++i) h_gen[i] = i + 1;
  CK(cudaMalloc(&d_flag, 8 * sizeof(int)));
  int gen = 0;
  for (int variant = (argc > 2 ? 8 : 0); variant < 15; ++variant) {
    for (int kern_us : {0, 70}) {
      double t0 = 0;
      for (int i = -20; i < N; ++i) {
        if (i == 0) { CK(cudaDeviceSynchronize()); t0 = now(); }
        const int k = ((i % DEPTH) + DEPTH) % DEPTH;
        if (i + 20 >= DEPTH + 2) CK(cudaEventSynchronize(out_done[k]));
        switch (variant) {
          case 0:  // three streams, device-side event waits (the en_bh_host_pipe schedule)
            CK(cudaMemcpyAsync(d_in[k], h_in, bytes, cudaMemcpyHostToDevice, s_in));
            CK(cudaEventRecord(in_done[k], s_in));
            CK(cudaStreamWaitEvent(s_cmp, in_done[k], 0));
            if (kern_us) sleep_kernel<<<1, 32, 0, s_cmp>>>(kern_us * 1000);
            CK(cudaEventRecord(cmp_done[k], s_cmp));
            CK(cudaStreamWaitEvent(s_out, cmp_done[k], 0));
            CK(cudaMemcpyAsync(h_out[k], d_out[k], bytes, cudaMemcpyDeviceToHost, s_out));
            CK(cudaEventRecord(out_done[k], s_out));
            break;
          case 1:  // one stream per slot, everything in order inside it (classic)
            CK(cudaMemcpyAsync(d_in[k], h_in, bytes, cudaMemcpyHostToDevice, s_slot[k]));
            if (kern_us) sleep_kernel<<<1, 32, 0, s_slot[k]>>>(kern_us * 1000);
            CK(cudaMemcpyAsync(h_out[k], d_out[k], bytes, cudaMemcpyDeviceToHost, s_slot[k]));
            CK(cudaEventRecord(out_done[k], s_slot[k]));
            break;
          case 2:  // upload by copy engine, kernel, download by an SM copy kernel into mapped host memory (own stream)
            CK(cudaMemcpyAsync(d_in[k], h_in, bytes, cudaMemcpyHostToDevice, s_in));
            CK(cudaEventRecord(in_done[k], s_in));
            CK(cudaStreamWaitEvent(s_cmp, in_done[k], 0));
            if (kern_us) sleep_kernel<<<1, 32, 0, s_cmp>>>(kern_us * 1000);
            CK(cudaEventRecord(cmp_done[k], s_cmp));
            CK(cudaStreamWaitEvent(s_out, cmp_done[k], 0));
            copy_kernel<<<64, 256, 0, s_out>>>((const float4*)d_out[k], (float4*)h_out[k], n16);
            CK(cudaEventRecord(out_done[k], s_out));
            break;
          case 3:  // both directions by SM copy kernels (zero-copy), three streams
            copy_kernel<<<64, 256, 0, s_in>>>((const float4*)h_in, (float4*)d_in[k], n16);
            CK(cudaEventRecord(in_done[k], s_in));
            CK(cudaStreamWaitEvent(s_cmp, in_done[k], 0));
            if (kern_us) sleep_kernel<<<1, 32, 0, s_cmp>>>(kern_us * 1000);
            CK(cudaEventRecord(cmp_done[k], s_cmp));
            CK(cudaStreamWaitEvent(s_out, cmp_done[k], 0));
            copy_kernel<<<64, 256, 0, s_out>>>((const float4*)d_out[k], (float4*)h_out[k], n16);
            CK(cudaEventRecord(out_done[k], s_out));
            break;
          case 4:  // zero-copy upload kernel, copy-engine download
            copy_kernel<<<64, 256, 0, s_in>>>((const float4*)h_in, (float4*)d_in[k], n16);
            CK(cudaEventRecord(in_done[k], s_in));
            CK(cudaStreamWaitEvent(s_cmp, in_done[k], 0));
            if (kern_us) sleep_kernel<<<1, 32, 0, s_cmp>>>(kern_us * 1000);
            CK(cudaEventRecord(cmp_done[k], s_cmp));
            CK(cudaStreamWaitEvent(s_out, cmp_done[k], 0));
            CK(cudaMemcpyAsync(h_out[k], d_out[k], bytes, cudaMemcpyDeviceToHost, s_out));
            CK(cudaEventRecord(out_done[k], s_out));
            break;
          case 5:  // uploads only
            CK(cudaMemcpyAsync(d_in[k], h_in, bytes, cudaMemcpyHostToDevice, s_in));
            CK(cudaEventRecord(out_done[k], s_in));
            break;
          case 6:  // zero-copy uploads only (SM kernel, 64 blocks)
            copy_kernel<<<64, 256, 0, s_in>>>((const float4*)h_in, (float4*)d_in[k], n16);
            CK(cudaEventRecord(out_done[k], s_in));
            break;
          case 8:  // lock step: upload(i) may only start when kernel(i-2) has finished, i.e. together with download(i-2)
            if (i + 20 >= 2) CK(cudaStreamWaitEvent(s_in, cmp_done[((i - 2) % DEPTH + DEPTH) % DEPTH], 0));
            CK(cudaMemcpyAsync(d_in[k], h_in, bytes, cudaMemcpyHostToDevice, s_in));
            CK(cudaEventRecord(in_done[k], s_in));
            CK(cudaStreamWaitEvent(s_cmp, in_done[k], 0));
            if (kern_us) sleep_kernel<<<1, 32, 0, s_cmp>>>(kern_us * 1000);
            CK(cudaEventRecord(cmp_done[k], s_cmp));
            CK(cudaStreamWaitEvent(s_out, cmp_done[k], 0));
            CK(cudaMemcpyAsync(h_out[k], d_out[k], bytes, cudaMemcpyDeviceToHost, s_out));
            CK(cudaEventRecord(out_done[k], s_out));
            break;
          case 9:  // copies in 4 chunks each (finer interleaving on the link)
            for (int c = 0; c < 4; ++c)
              CK(cudaMemcpyAsync((char*)d_in[k] + c * (bytes / 4), (char*)h_in + c * (bytes / 4), bytes / 4, cudaMemcpyHostToDevice, s_in));
            CK(cudaEventRecord(in_done[k], s_in));
            CK(cudaStreamWaitEvent(s_cmp, in_done[k], 0));
            if (kern_us) sleep_kernel<<<1, 32, 0, s_cmp>>>(kern_us * 1000);
            CK(cudaEventRecord(cmp_done[k], s_cmp));
            CK(cudaStreamWaitEvent(s_out, cmp_done[k], 0));
            for (int c = 0; c < 4; ++c)
              CK(cudaMemcpyAsync((char*)h_out[k] + c * (bytes / 4), (char*)d_out[k] + c * (bytes / 4), bytes / 4, cudaMemcpyDeviceToHost, s_out));
            CK(cudaEventRecord(out_done[k], s_out));
            break;
          case 10:  // kernel in the upload stream (no event between upload and kernel), download on its own stream
            CK(cudaMemcpyAsync(d_in[k], h_in, bytes, cudaMemcpyHostToDevice, s_in));
            if (kern_us) sleep_kernel<<<1, 32, 0, s_in>>>(kern_us * 1000);
            CK(cudaEventRecord(cmp_done[k], s_in));
            CK(cudaStreamWaitEvent(s_out, cmp_done[k], 0));
            CK(cudaMemcpyAsync(h_out[k], d_out[k], bytes, cudaMemcpyDeviceToHost, s_out));
            CK(cudaEventRecord(out_done[k], s_out));
            break;
          case 11:  // host-driven download: the host waits for kernel(i-1) and only then issues its download (no
                    // copy-engine channel ever sits on a semaphore waiting for a kernel)
            CK(cudaMemcpyAsync(d_in[k], h_in, bytes, cudaMemcpyHostToDevice, s_in));
            CK(cudaEventRecord(in_done[k], s_in));
            CK(cudaStreamWaitEvent(s_cmp, in_done[k], 0));
            if (kern_us) sleep_kernel<<<1, 32, 0, s_cmp>>>(kern_us * 1000);
            CK(cudaEventRecord(cmp_done[k], s_cmp));
            if (i + 20 >= 1) {
              const int kp = ((i - 1) % DEPTH + DEPTH) % DEPTH;
              CK(cudaEventSynchronize(cmp_done[kp]));
              CK(cudaMemcpyAsync(h_out[kp], d_out[kp], bytes, cudaMemcpyDeviceToHost, s_out));
              CK(cudaEventRecord(out_done[kp], s_out));
            }
            break;
          case 12:  // host-driven on both sides: kernel(i) is launched by the host once upload(i) has landed
            CK(cudaMemcpyAsync(d_in[k], h_in, bytes, cudaMemcpyHostToDevice, s_in));
            CK(cudaEventRecord(in_done[k], s_in));
            if (i + 20 >= 1) {
              const int kp = ((i - 1) % DEPTH + DEPTH) % DEPTH;
              CK(cudaEventSynchronize(in_done[kp]));
              if (kern_us) sleep_kernel<<<1, 32, 0, s_cmp>>>(kern_us * 1000);
              CK(cudaEventRecord(cmp_done[kp], s_cmp));
            }
            if (i + 20 >= 2) {
              const int kq = ((i - 2) % DEPTH + DEPTH) % DEPTH;
              CK(cudaEventSynchronize(cmp_done[kq]));
              CK(cudaMemcpyAsync(h_out[kq], d_out[kq], bytes, cudaMemcpyDeviceToHost, s_out));
              CK(cudaEventRecord(out_done[kq], s_out));
            }
            break;
          case 13:  // upload -> kernel through a flag the kernel polls (no event between them); download waits on an event
          case 14:  // ... and the download is issued by the host once the kernel has finished (no device-side wait at all)
            if (i == -20) { CK(cudaDeviceSynchronize()); CK(cudaMemset(d_flag, 0, 8 * sizeof(int))); gen = 0; }
            CK(cudaMemcpyAsync(d_in[k], h_in, bytes, cudaMemcpyHostToDevice, s_in));
            CK(cudaMemcpyAsync(d_flag + k, h_gen + gen, sizeof(int), cudaMemcpyHostToDevice, s_in));
            wait_flag_then_sleep<<<1, 32, 0, s_cmp>>>(d_flag + k, gen + 1, kern_us * 1000);
            ++gen;
            CK(cudaEventRecord(cmp_done[k], s_cmp));
            if (variant == 13) {
              CK(cudaStreamWaitEvent(s_out, cmp_done[k], 0));
              CK(cudaMemcpyAsync(h_out[k], d_out[k], bytes, cudaMemcpyDeviceToHost, s_out));
              CK(cudaEventRecord(out_done[k], s_out));
            } else if (i + 20 >= 1) {
              const int kp = ((i - 1) % DEPTH + DEPTH) % DEPTH;
              CK(cudaEventSynchronize(cmp_done[kp]));
              CK(cudaMemcpyAsync(h_out[kp], d_out[kp], bytes, cudaMemcpyDeviceToHost, s_out));
              CK(cudaEventRecord(out_done[kp], s_out));
            }
            break;
          case 7:  // zero-copy downloads only
            copy_kernel<<<64, 256, 0, s_out>>>((const float4*)d_out[k], (float4*)h_out[k], n16);
            CK(cudaEventRecord(out_done[k], s_out));
            break;
        }
      }
      CK(cudaDeviceSynchronize());
      const double dt = (now() - t0) / N * 1e3;
      static const char* names[] = {"3 streams + events (CE up, CE down)", "stream per slot (CE up, CE down)",
                                    "CE up, SM-kernel down (mapped host)", "SM-kernel up, SM-kernel down",
                                    "SM-kernel up, CE down", "CE uploads only", "SM-kernel uploads only",
                                    "SM-kernel downloads only", "lock step (up(i) starts with down(i-2))",
                                    "3 streams, copies in 4 chunks", "kernel in the upload stream",
                                    "host-driven download", "host-driven kernel and download",
                                    "flag-polling kernel, event download", "flag-polling kernel, host download"};
      printf("depth %d  kernel %2d us  %-38s  %.4f ms/step\n", DEPTH, kern_us, names[variant], dt);
    }
  }
  // ---- timeline of variant 0 with timing events (who waits for whom?)
  {
    const int NT = 40;
    static cudaEvent_t ev[64][6];
    for (int i = 0; i < NT; ++i) for (int j = 0; j < 6; ++j) CK(cudaEventCreate(&ev[i][j]));
    for (int kern_us : {0, 70}) {
      CK(cudaDeviceSynchronize());
      for (int i = 0; i < NT; ++i) {
        const int k = i % DEPTH;
        if (i >= DEPTH) CK(cudaEventSynchronize(ev[i - DEPTH][5]));
        CK(cudaEventRecord(ev[i][0], s_in));
        CK(cudaMemcpyAsync(d_in[k], h_in, bytes, cudaMemcpyHostToDevice, s_in));
        CK(cudaEventRecord(ev[i][1], s_in));
        CK(cudaStreamWaitEvent(s_cmp, ev[i][1], 0));
        CK(cudaEventRecord(ev[i][2], s_cmp));
        if (kern_us) sleep_kernel<<<1, 32, 0, s_cmp>>>(kern_us * 1000);
        CK(cudaEventRecord(ev[i][3], s_cmp));
        CK(cudaStreamWaitEvent(s_out, ev[i][3], 0));
        CK(cudaEventRecord(ev[i][4], s_out));
        CK(cudaMemcpyAsync(h_out[k], d_out[k], bytes, cudaMemcpyDeviceToHost, s_out));
        CK(cudaEventRecord(ev[i][5], s_out));
      }
      CK(cudaDeviceSynchronize());
      printf("timeline (us from the first traced upload), kernel %d us: up[start,end] kernel[start,end] down[start,end]\n", kern_us);
      for (int i = NT - 8; i < NT; ++i) {
        float t[6];
        for (int j = 0; j < 6; ++j) CK(cudaEventElapsedTime(&t[j], ev[NT - 8][0], ev[i][j]));
        printf("  step %2d  up [%7.1f %7.1f]  k [%7.1f %7.1f]  down [%7.1f %7.1f]\n", i, t[0] * 1e3, t[1] * 1e3, t[2] * 1e3,
               t[3] * 1e3, t[4] * 1e3, t[5] * 1e3);
      }
    }
  }
  return 0;
}
