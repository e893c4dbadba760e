// comm.cu -- multi-GPU plumbing of the C ABI for a C++ host (no Python, no torch): ONE process drives the G devices of a
// box.  Instances shard by contiguous range, one handle per device; there is no exchange inside a step.
// The two real exchanges of the path run over NVLink peer memory, written by this library's own kernels:
//   * decimated trajectory gather (config 4): every device holds the FULL buffer [capacity][13][N_total]; each step kernel
//     stores the snapshots of its instances into the column range of its shard in EVERY device's buffer (peer stores,
//     cdpr_set_snapshot_peers) -- the gather rides along with the compute, the only synchronisation is the end of the pass;
//   * rollout cost vector (config 5): k_sum_peers on every device adds the G partial vectors in RANK ORDER, reading the
//     peers' copies over NVLink -- a deterministic all-reduce (every device ends with the same bits; 4096 doubles are
//     latency-bound, a ring would not help).
// The multi-process form of the same two exchanges (one rank per GPU under torchrun, NVLS multimem.st through symmetric
// memory, NCCL all-reduce) lives in cdpr_simulation_b200/distributed.py.
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/cdpr_b200.h"
#include "common.cuh"

struct cdpr_comm {
  std::vector<int> devices;
  std::vector<void *> gather;  // one buffer per device
  std::vector<cudaStream_t> streams;
  int64_t total = 0, capacity = 0;
  std::string err;
};

static thread_local std::string g_comm_error;
static int cfail(cdpr_comm *c, int code, const std::string &msg) {
  if (c) c->err = msg; else g_comm_error = msg;
  return code;
}

// out[e] = sum over ranks r = 0 .. n-1 (ascending) of in[r][e]
__global__ void k_sum_peers(const double *const *in, int n, long long n_elem, double *out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elem) return;
  double acc = 0.0;
  for (int r = 0; r < n; ++r) acc += in[r][e];
  out[e] = acc;
}

extern "C" const char *cdpr_comm_last_error(cdpr_comm_t c) { return c ? c->err.c_str() : g_comm_error.c_str(); }

extern "C" int cdpr_comm_create(int n_devices, const int *devices, cdpr_comm_t *out) {
  if (!out || n_devices < 1 || n_devices > 8) return cfail(nullptr, CDPR_ERR_BAD_ARG, "1 to 8 devices");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return cfail(nullptr, CDPR_ERR_NO_DEVICE, "no CUDA device");
  cdpr_comm *c = new (std::nothrow) cdpr_comm();
  if (!c) return cfail(nullptr, CDPR_ERR_NOMEM, "out of host memory");
  for (int r = 0; r < n_devices; ++r) {
    const int d = devices ? devices[r] : r;
    if (d < 0 || d >= ndev) { delete c; return cfail(nullptr, CDPR_ERR_NO_DEVICE, "device ordinal out of range"); }
    c->devices.push_back(d);
  }
  for (int a : c->devices) {
    cudaSetDevice(a);
    cudaStream_t st = nullptr;
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    c->streams.push_back(st);
    for (int b : c->devices) {
      if (a == b) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, a, b);
      if (!can) { cdpr_comm_destroy(c); return cfail(nullptr, CDPR_ERR_UNSUPPORTED, "no peer access between the devices (NVLink / PCIe P2P needed)"); }
      const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cdpr_comm_destroy(c); return cfail(nullptr, CDPR_ERR_CUDA, cudaGetErrorString(e)); }
      cudaGetLastError();
    }
  }
  *out = c;
  return CDPR_OK;
}

extern "C" int cdpr_comm_destroy(cdpr_comm_t c) {
  if (!c) return CDPR_ERR_BAD_ARG;
  for (size_t r = 0; r < c->devices.size(); ++r) {
    cudaSetDevice(c->devices[r]);
    if (r < c->gather.size() && c->gather[r]) cudaFree(c->gather[r]);
    if (r < c->streams.size() && c->streams[r]) cudaStreamDestroy(c->streams[r]);
  }
  delete c;
  return CDPR_OK;
}

extern "C" int cdpr_comm_size(cdpr_comm_t c) { return c ? (int)c->devices.size() : -1; }

extern "C" int cdpr_comm_attach_gather(cdpr_comm_t c, cdpr_handle *handles, const int64_t *instances, int64_t every, int64_t capacity) {
  if (!c || !handles || !instances || every < 1 || capacity < 1) return CDPR_ERR_BAD_ARG;
  const int g = (int)c->devices.size();
  int64_t total = 0;
  for (int r = 0; r < g; ++r) { if (!handles[r] || instances[r] < 1) return cfail(c, CDPR_ERR_BAD_ARG, "one handle with >= 1 instances per device"); total += instances[r]; }
  for (int r = 0; r < (int)c->gather.size(); ++r) { cudaSetDevice(c->devices[r]); if (c->gather[r]) cudaFree(c->gather[r]); }
  c->gather.assign(g, nullptr);
  const size_t bytes = sizeof(double) * 13 * (size_t)total * (size_t)capacity;
  for (int r = 0; r < g; ++r) {
    cudaSetDevice(c->devices[r]);
    if (cudaMalloc(&c->gather[r], bytes) != cudaSuccess) return cfail(c, CDPR_ERR_NOMEM, "cudaMalloc(gather buffer) failed");
    cudaMemset(c->gather[r], 0, bytes);
    cudaDeviceSynchronize();
  }
  c->total = total; c->capacity = capacity;
  int64_t offset = 0;
  for (int r = 0; r < g; ++r) {  // handle r stores its columns [offset, offset + n_r) into every device's buffer
    const int rc = cdpr_set_snapshot_peers(handles[r], every, c->gather.data(), g, offset, total, capacity);
    if (rc != CDPR_OK) return cfail(c, rc, std::string("cdpr_set_snapshot_peers: ") + cdpr_last_error(handles[r]));
    offset += instances[r];
  }
  return CDPR_OK;
}

extern "C" void *cdpr_comm_gather_buffer(cdpr_comm_t c, int rank) {
  return (c && rank >= 0 && rank < (int)c->gather.size()) ? c->gather[rank] : nullptr;
}

// One pass on every device: the launches are enqueued back to back (they run concurrently, one per GPU), then every device
// is synchronised -- after that every gather buffer holds the pass's snapshots of ALL shards.
extern "C" int cdpr_comm_step(cdpr_comm_t c, cdpr_handle *handles, int64_t k_steps) {
  if (!c || !handles) return CDPR_ERR_BAD_ARG;
  const int g = (int)c->devices.size();
  for (int r = 0; r < g; ++r) {
    const int rc = cdpr_step(handles[r], k_steps);
    if (rc != CDPR_OK) return cfail(c, rc, std::string("cdpr_step: ") + cdpr_last_error(handles[r]));
  }
  for (int r = 0; r < g; ++r) {
    const int rc = cdpr_synchronize(handles[r]);
    if (rc != CDPR_OK) return cfail(c, rc, std::string("cdpr_synchronize: ") + cdpr_last_error(handles[r]));
  }
  return CDPR_OK;
}

extern "C" int cdpr_comm_allreduce(cdpr_comm_t c, void *const *dev_vectors, int64_t n_elements) {
  if (!c || !dev_vectors || n_elements < 1) return CDPR_ERR_BAD_ARG;
  const int g = (int)c->devices.size();
  // every device: its own scratch copy of the pointer table and of the result, then the rank-ordered sum over peer memory
  std::vector<double *> out(g, nullptr);
  std::vector<const double **> table(g, nullptr);
  int rc = CDPR_OK;
  for (int r = 0; r < g && rc == CDPR_OK; ++r) {
    cudaSetDevice(c->devices[r]);
    if (cudaMalloc((void **)&out[r], sizeof(double) * (size_t)n_elements) != cudaSuccess || cudaMalloc((void **)&table[r], sizeof(double *) * g) != cudaSuccess) {
      rc = cfail(c, CDPR_ERR_NOMEM, "cudaMalloc(all-reduce scratch) failed");
      break;
    }
    cudaMemcpyAsync(table[r], dev_vectors, sizeof(double *) * g, cudaMemcpyHostToDevice, c->streams[r]);
    k_sum_peers<<<(unsigned)((n_elements + 255) / 256), 256, 0, c->streams[r]>>>(table[r], g, n_elements, out[r]);
    if (cudaGetLastError() != cudaSuccess) rc = cfail(c, CDPR_ERR_CUDA, "k_sum_peers launch failed");
  }
  // nobody may overwrite its input before every peer has read it: finish all sums first, then copy back
  for (int r = 0; r < g; ++r) { cudaSetDevice(c->devices[r]); if (cudaStreamSynchronize(c->streams[r]) != cudaSuccess && rc == CDPR_OK) rc = cfail(c, CDPR_ERR_CUDA, "all-reduce kernel failed"); }
  for (int r = 0; r < g && rc == CDPR_OK; ++r) {
    cudaSetDevice(c->devices[r]);
    cudaMemcpyAsync(dev_vectors[r], out[r], sizeof(double) * (size_t)n_elements, cudaMemcpyDeviceToDevice, c->streams[r]);
  }
  for (int r = 0; r < g; ++r) {
    cudaSetDevice(c->devices[r]);
    cudaStreamSynchronize(c->streams[r]);
    if (out[r]) cudaFree(out[r]);
    if (table[r]) cudaFree((void *)table[r]);
  }
  return rc;
}
