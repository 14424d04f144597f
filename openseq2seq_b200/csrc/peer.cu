// Data-parallel gradient sum over NVLink peer memory (one process per GPU, one node).
//
// Reference being replaced: hvd.allreduce(grad) per variable, open_seq2seq/optimizers/optimizers.py:77-104
// (reduce_gradients) -- NCCL ring kernels in the reference's Horovod build.  Here the flat fp32 gradient buffer of
// every rank is mapped into every other rank (CUDA IPC) and one bucket of it is summed in two phases:
//
//   phase 0  rank r copies its contribution to slice p of the bucket into rank p's staging slot r   (copy engines)
//            and raises flag[0][bucket][r] on rank p;
//   sum      rank r waits for the N-1 flags, then adds the N-1 staged slices onto its own slice r   (one small kernel)
//   phase 1  rank r copies the summed slice r into the gradient buffer of every other rank          (copy engines)
//            and raises flag[1][bucket][r] there; the optimizer waits for all phase-1 flags.
//
// The transport runs on the copy engines, so the only SM work is the slice sum (1/N of the bucket per rank) and the
// one-warp flag kernels: the persistent one-CTA-per-SM convolution kernels of the backward pass that the exchange
// overlaps keep their SMs, which is what an SM-resident ring all-reduce takes from them.  Every rank ends with the
// SAME bits (slice r is summed once, by rank r).  Flags only ever count up (one increment per step and source), the
// expected value lives in device memory, so the sequence can be replayed from a CUDA graph.
//
// Buffer reuse is safe without extra barriers: a peer overwrites my gradient slice (phase 1) only after it has seen
// my phase-0 flag, i.e. after my contribution left; a peer refills my staging slot in the next step only after its
// optimizer ran, which waited for my phase-1 flag, which I raise after my slice sum has read the slot.
#include <cstring>
#include <vector>

#include "common.h"
#include "kernels.h"

namespace os2s {

namespace {

constexpr int kMaxPeers = 16;
struct PtrPack {
  void* p[kMaxPeers];
};

__device__ __forceinline__ unsigned long long peer_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// flags.p[i] (i != self) += 1, system scope.  The copies that precede this kernel on the stream have completed.
__global__ void peer_signal_kernel(PtrPack flags, int n, int self) {
  const int i = threadIdx.x;
  if (i < n && i != self) {
    __threadfence_system();
    atomicAdd_system(reinterpret_cast<unsigned*>(flags.p[i]), 1u);
  }
}

// Wait until flags[g * n + s] >= expected[g] + 1 for every group g < n_groups and source s != self, then advance
// expected[g].  A wait that lasts longer than timeout_ns raises *timed_out and returns (the host turns that into an
// error); it never spins forever.
__global__ void peer_wait_kernel(const volatile unsigned* flags, unsigned* expected, int n_groups, int n, int self,
                                 unsigned long long timeout_ns, unsigned* timed_out) {
  const int t = threadIdx.x;
  if (t < n_groups * n) {
    const int g = t / n, s = t - g * n;
    if (s != self) {
      const unsigned target = expected[g] + 1u;
      const unsigned long long t0 = peer_globaltimer();
      while ((int)(flags[t] - target) < 0) {
        __nanosleep(256);
        // (after the first time-out the exchange is broken anyway: later waits return at once)
        if (peer_globaltimer() - t0 > timeout_ns || *reinterpret_cast<volatile unsigned*>(timed_out)) {
          atomicExch(timed_out, 1u);
          break;
        }
      }
    }
  }
  __syncthreads();
  if (t < n_groups) expected[t] += 1u;
  __threadfence_system();
}

// g[i] += sum over the staged contributions (rank order, own slot skipped by the caller's pointer list)
__global__ void __launch_bounds__(256) peer_slice_sum_kernel(float* __restrict__ g, PtrPack staged, int n,
                                                             long long count) {
  const long long n4 = count >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  constexpr int U = 4;      // independent 16-byte loads per pointer and thread in flight
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += stride * U) {
    float4 a[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      a[u] = i < n4 ? reinterpret_cast<const float4*>(g)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int p = 0; p < n; ++p) {
      float4 b[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long i = i0 + u * stride;
        b[u] = i < n4 ? __ldcs(reinterpret_cast<const float4*>(staged.p[p]) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        a[u].x += b[u].x;
        a[u].y += b[u].y;
        a[u].z += b[u].z;
        a[u].w += b[u].w;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < n4) reinterpret_cast<float4*>(g)[i] = a[u];
    }
  }
  // (slices are multiples of 4 floats except the last one of a bucket)
  const long long tail = n4 << 2;
  for (long long i = tail + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    float a = g[i];
    for (int p = 0; p < n; ++p) a += reinterpret_cast<const float*>(staged.p[p])[i];
    g[i] = a;
  }
}

typedef CUresult (*GetAddressRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
GetAddressRangeFn get_address_range_fn() {
  static GetAddressRangeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (GetAddressRangeFn)p;
  }
  return fn;
}

}  // namespace

struct PeerExchange {
  int rank, world, n_buckets;
  float* grad[kMaxPeers];        // gradient buffer of every rank (own entry: local pointer)
  char* stage[kMaxPeers];        // staging allocation of every rank: [flags | expected | slots]
  std::vector<long long> start, end, chunk, slot_off;   // per bucket, in floats
  long long flags_bytes, header_bytes;
  unsigned* timed_out;           // device word inside the local header
  unsigned long long timeout_ns;

  unsigned* flag(int on_rank, int phase, int b, int src) const {
    return reinterpret_cast<unsigned*>(stage[on_rank]) + ((long long)phase * n_buckets + b) * world + src;
  }
  unsigned* expected(int phase, int b) const {
    return reinterpret_cast<unsigned*>(stage[rank] + flags_bytes) + (long long)phase * n_buckets + b;
  }
  float* slot(int on_rank, int b, int src) const {
    return reinterpret_cast<float*>(stage[on_rank] + header_bytes) + slot_off[b] + (long long)src * chunk[b];
  }
  // slice of bucket b owned by rank r: [lo, hi) in floats of the flat buffer
  void slice(int b, int r, long long* lo, long long* hi) const {
    long long l = start[b] + (long long)r * chunk[b], h = l + chunk[b];
    if (l > end[b]) l = end[b];
    if (h > end[b]) h = end[b];
    *lo = l;
    *hi = h;
  }
};

static long long peer_chunk(long long n, int world) {
  long long c = (n + world - 1) / world;
  return (c + 3) & ~3ll;     // 16-byte aligned slices
}

long long peer_stage_bytes(int world, int n_buckets, const long long* start, const long long* end) {
  long long slots = 0;
  for (int b = 0; b < n_buckets; ++b) slots += peer_chunk(end[b] - start[b], world) * world;
  const long long flags = (((long long)2 * n_buckets * world * 4) + 255) & ~255ll;
  const long long expected = (((long long)2 * n_buckets * 4 + 4) + 255) & ~255ll;
  return flags + expected + slots * 4;
}

int ipc_export(const void* ptr, unsigned char* handle, long long* offset) {
  GetAddressRangeFn fn = get_address_range_fn();
  if (!fn) return fail(ERR_UNSUPPORTED, "ipc_export: cuMemGetAddressRange is not available");
  CUdeviceptr base = 0;
  size_t size = 0;
  if (fn(&base, &size, (CUdeviceptr)ptr) != CUDA_SUCCESS) return fail(ERR_CUDA, "ipc_export: cuMemGetAddressRange failed");
  cudaIpcMemHandle_t h;
  OS2S_CUDA(cudaIpcGetMemHandle(&h, (void*)base));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  std::memcpy(handle, &h, 64);
  *offset = (long long)((CUdeviceptr)ptr - base);
  return OK;
}

int ipc_open(const unsigned char* handle, void** base) {
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, 64);
  OS2S_CUDA(cudaIpcOpenMemHandle(base, h, cudaIpcMemLazyEnablePeerAccess));
  return OK;
}

int ipc_close(void* base) {
  OS2S_CUDA(cudaIpcCloseMemHandle(base));
  return OK;
}

int peer_create(int rank, int world, void* const* grad, void* const* stage, int n_buckets, const long long* start,
                const long long* end, double timeout_s, PeerExchange** out) {
  if (world < 2 || world > kMaxPeers || rank < 0 || rank >= world) return fail(ERR_INVALID, "peer_create: bad rank / world");
  if (n_buckets < 1 || 2 * n_buckets * world > 1024) return fail(ERR_INVALID, "peer_create: too many buckets");
  PeerExchange* px = new PeerExchange();
  px->rank = rank;
  px->world = world;
  px->n_buckets = n_buckets;
  long long off = 0;
  for (int b = 0; b < n_buckets; ++b) {
    if (end[b] < start[b] || (start[b] & 3)) {
      delete px;
      return fail(ERR_INVALID, "peer_create: bucket bounds must be ordered and 16-byte aligned");
    }
    px->start.push_back(start[b]);
    px->end.push_back(end[b]);
    px->chunk.push_back(peer_chunk(end[b] - start[b], world));
    px->slot_off.push_back(off);
    off += px->chunk.back() * world;
  }
  px->flags_bytes = (((long long)2 * n_buckets * world * 4) + 255) & ~255ll;
  px->header_bytes = px->flags_bytes + ((((long long)2 * n_buckets * 4 + 4) + 255) & ~255ll);
  for (int r = 0; r < world; ++r) {
    px->grad[r] = reinterpret_cast<float*>(grad[r]);
    px->stage[r] = reinterpret_cast<char*>(stage[r]);
    if (!grad[r] || !stage[r] || (reinterpret_cast<uintptr_t>(grad[r]) & 15) || (reinterpret_cast<uintptr_t>(stage[r]) & 255)) {
      delete px;
      return fail(ERR_INVALID, "peer_create: null or misaligned peer pointer");
    }
  }
  px->timed_out = reinterpret_cast<unsigned*>(px->stage[rank] + px->flags_bytes) + 2 * n_buckets;
  px->timeout_ns = (unsigned long long)(timeout_s * 1e9);
  *out = px;
  return OK;
}

void peer_destroy(PeerExchange* px) { delete px; }

int peer_set_timeout(PeerExchange* px, double timeout_s) {
  if (!px || !(timeout_s > 0.0)) return fail(ERR_INVALID, "peer_set_timeout: bad argument");
  px->timeout_ns = (unsigned long long)(timeout_s * 1e9);
  return OK;
}

int peer_exchange_bucket(PeerExchange* px, int b, cudaStream_t st) {
  if (!px || b < 0 || b >= px->n_buckets) return fail(ERR_INVALID, "peer_exchange_bucket: bad bucket");
  const int N = px->world, r = px->rank;
  // phase 0: my contribution to slice p -> staging slot r on rank p
  PtrPack f0, f1, staged;
  int n_staged = 0;
  for (int p = 0; p < N; ++p) {
    f0.p[p] = px->flag(p, 0, b, r);
    f1.p[p] = px->flag(p, 1, b, r);
    if (p == r) continue;
    long long lo, hi;
    px->slice(b, p, &lo, &hi);
    if (hi > lo)
      OS2S_CUDA(cudaMemcpyAsync(px->slot(p, b, r), px->grad[r] + lo, (size_t)(hi - lo) * 4, cudaMemcpyDeviceToDevice, st));
    staged.p[n_staged++] = px->slot(r, b, p);
  }
  peer_signal_kernel<<<1, 32, 0, st>>>(f0, N, r);
  peer_wait_kernel<<<1, 32, 0, st>>>(px->flag(r, 0, b, 0), px->expected(0, b), 1, N, r, px->timeout_ns, px->timed_out);
  long long lo, hi;
  px->slice(b, r, &lo, &hi);
  if (hi > lo) {
    // HBM-bound and short; a narrow grid keeps it out of the way of the convolution kernels it overlaps
    const long long n4 = (hi - lo + 3) >> 2;
    int grid = (int)((n4 + 255) / 256);
    if (grid > 32) grid = 32;
    peer_slice_sum_kernel<<<grid, 256, 0, st>>>(px->grad[r] + lo, staged, n_staged, hi - lo);
    // phase 1: the summed slice -> the gradient buffer of every other rank
    for (int p = 0; p < N; ++p)
      if (p != r)
        OS2S_CUDA(cudaMemcpyAsync(px->grad[p] + lo, px->grad[r] + lo, (size_t)(hi - lo) * 4, cudaMemcpyDeviceToDevice, st));
  }
  peer_signal_kernel<<<1, 32, 0, st>>>(f1, N, r);
  return check_launch("peer_exchange_bucket");
}

int peer_finish(PeerExchange* px, cudaStream_t st) {
  if (!px) return fail(ERR_INVALID, "peer_finish: null context");
  const int threads = ((px->n_buckets * px->world + 31) / 32) * 32;
  peer_wait_kernel<<<1, threads, 0, st>>>(px->flag(px->rank, 1, 0, 0), px->expected(1, 0), px->n_buckets, px->world,
                                          px->rank, px->timeout_ns, px->timed_out);
  return check_launch("peer_finish");
}

int peer_timed_out(PeerExchange* px, int* flag) {
  if (!px || !flag) return fail(ERR_INVALID, "peer_timed_out: null argument");
  unsigned v = 0;
  OS2S_CUDA(cudaMemcpy(&v, px->timed_out, 4, cudaMemcpyDeviceToHost));
  *flag = (int)v;
  return OK;
}

}  // namespace os2s
