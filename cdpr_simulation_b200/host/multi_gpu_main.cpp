// multi_gpu_main.cpp -- configs 4 and 5 of the north_star driven by a C++ host through the C ABI alone (no Python, no torch):
// one process, G devices.  Instances are sharded by contiguous range; the step kernels store the decimated
// trajectory straight into every device's gather buffer over NVLink peer memory; the rollout cost vector is all-reduced by
// the library's rank-ordered peer sum.  Both results are compared with a single-GPU run over all instances.
// usage: cdpr_multigpu_check [G] [instances per device] [steps] [snapshot every]     exit code 0 = all checks passed
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../include/cdpr_b200.h"

#define CHECK(call)                                                                                   \
  do {                                                                                                \
    const int rc_ = (call);                                                                           \
    if (rc_ != CDPR_OK) { std::fprintf(stderr, "%s failed: %d\n", #call, rc_); return 2; }            \
  } while (0)

int main(int argc, char **argv) {
  int ndev = 0;
  cudaGetDeviceCount(&ndev);
  const int G = argc > 1 ? std::atoi(argv[1]) : ndev;
  const int64_t nper = argc > 2 ? std::atoll(argv[2]) : 4099;  // not a multiple of the block size on purpose
  const int steps = argc > 3 ? std::atoi(argv[3]) : 300, every = argc > 4 ? std::atoi(argv[4]) : 100;
  if (G < 1 || G > ndev || G > 8) { std::fprintf(stderr, "need 1..%d devices\n", ndev < 8 ? ndev : 8); return 2; }
  const int nc = 8, n_snap = steps / every;
  const int64_t total = nper * G;
  cdpr_config cfg;
  cdpr_config_default(&cfg, nc);

  // C3-style inputs for all instances: sine parameters and a pose near home
  std::mt19937_64 gen(12345);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  std::vector<double> amp(total), freq(total), phase(total), pose(7 * total), twist(6 * total, 0.0);
  for (int64_t i = 0; i < total; ++i) {
    amp[i] = 0.01 + 0.05 * U(gen); freq[i] = 0.05 + 0.15 * U(gen); phase[i] = 6.283185307179586 * U(gen);
    double q[4] = {0.02 * (U(gen) - 0.5), 0.02 * (U(gen) - 0.5), 0.02 * (U(gen) - 0.5), 1.0};
    const double nq = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    pose[7 * i + 0] = 0.02 * (U(gen) - 0.5); pose[7 * i + 1] = 0.02 * (U(gen) - 0.5); pose[7 * i + 2] = 0.3 + 0.02 * (U(gen) - 0.5);
    for (int k = 0; k < 4; ++k) pose[7 * i + 3 + k] = q[k] / nq;
  }

  // ---- config 4: sharded run with the fused gather ----------------------------------------------------------------
  cdpr_comm_t comm = nullptr;
  CHECK(cdpr_comm_create(G, nullptr, &comm));
  std::vector<cdpr_handle> h(G, nullptr);
  std::vector<int64_t> inst(G, nper);
  for (int r = 0; r < G; ++r) {
    CHECK(cdpr_create(&cfg, nper, r, &h[r]));
    const int64_t lo = r * nper;  // global id = offset + local id
    CHECK(cdpr_set_platform_state(h[r], pose.data() + 7 * lo, twist.data() + 6 * lo));
    CHECK(cdpr_set_sine_cmd(h[r], amp.data() + lo, freq.data() + lo, phase.data() + lo, nper));
  }
  CHECK(cdpr_comm_attach_gather(comm, h.data(), inst.data(), every, n_snap));
  CHECK(cdpr_comm_step(comm, h.data(), steps));

  // single-GPU run over all instances with local snapshots
  cdpr_handle one = nullptr;
  CHECK(cdpr_create(&cfg, total, 0, &one));
  CHECK(cdpr_set_platform_state(one, pose.data(), twist.data()));
  CHECK(cdpr_set_sine_cmd(one, amp.data(), freq.data(), phase.data(), total));
  const size_t traj_bytes = sizeof(double) * 13 * (size_t)total * n_snap;
  void *ref_dev = nullptr;
  cudaSetDevice(0);
  cudaMalloc(&ref_dev, traj_bytes);
  cudaMemset(ref_dev, 0, traj_bytes);
  cudaDeviceSynchronize();
  CHECK(cdpr_set_snapshots(one, every, ref_dev, n_snap));
  CHECK(cdpr_step(one, steps));
  CHECK(cdpr_synchronize(one));
  std::vector<unsigned char> ref(traj_bytes), got(traj_bytes);
  cudaMemcpy(ref.data(), ref_dev, traj_bytes, cudaMemcpyDeviceToHost);
  bool ok_gather = true;
  for (int r = 0; r < G; ++r) {
    cudaSetDevice(r);
    cudaMemcpy(got.data(), cdpr_comm_gather_buffer(comm, r), traj_bytes, cudaMemcpyDeviceToHost);
    const bool same = std::memcmp(ref.data(), got.data(), traj_bytes) == 0;
    std::printf("[config 4] device %d: gathered trajectory (%d snapshots x 13 x %lld instances, %d shards) bitwise equal to the 1-GPU run: %s\n", r, n_snap,
                (long long)total, G, same ? "True" : "False");
    ok_gather = ok_gather && same;
  }
  for (int r = 0; r < G; ++r) cdpr_destroy(h[r]);
  cdpr_destroy(one);
  cudaSetDevice(0);
  cudaFree(ref_dev);

  // ---- config 5: rollouts, robots sharded over the devices, cost vector all-reduced -------------------------------
  const int64_t robots_per_dev = 3, n_seq = 256, n_cmd = 8, spc = 8;
  std::vector<float> cmds((size_t)n_seq * n_cmd * nc);
  std::normal_distribution<double> Nn(0.0, 0.03);
  for (auto &c : cmds) c = (float)Nn(gen);
  const double target[3] = {0.0, 0.0, 0.32};
  std::vector<void *> cost(G, nullptr);
  for (int r = 0; r < G; ++r) {
    cdpr_handle hr = nullptr;
    CHECK(cdpr_create(&cfg, robots_per_dev * n_seq, r, &hr));
    cudaSetDevice(r);
    cudaMalloc(&cost[r], sizeof(double) * n_seq);
    cudaDeviceSynchronize();
    const int64_t lo = r * robots_per_dev;
    CHECK(cdpr_rollout(hr, robots_per_dev, n_seq, pose.data() + 7 * lo, twist.data() + 6 * lo, cmds.data(), n_cmd, spc, target, 0.05, cost[r], nullptr));
    CHECK(cdpr_synchronize(hr));
    cdpr_destroy(hr);
  }
  CHECK(cdpr_comm_allreduce(comm, cost.data(), n_seq));
  std::vector<std::vector<double>> sums(G, std::vector<double>(n_seq));
  for (int r = 0; r < G; ++r) { cudaSetDevice(r); cudaMemcpy(sums[r].data(), cost[r], sizeof(double) * n_seq, cudaMemcpyDeviceToHost); cudaFree(cost[r]); }
  bool ok_same = true;
  for (int r = 1; r < G; ++r) ok_same = ok_same && std::memcmp(sums[0].data(), sums[r].data(), sizeof(double) * n_seq) == 0;
  cdpr_handle all = nullptr;
  CHECK(cdpr_create(&cfg, robots_per_dev * G * n_seq, 0, &all));
  cudaSetDevice(0);
  void *cost_one = nullptr;
  cudaMalloc(&cost_one, sizeof(double) * n_seq);
  cudaDeviceSynchronize();
  CHECK(cdpr_rollout(all, robots_per_dev * G, n_seq, pose.data(), twist.data(), cmds.data(), n_cmd, spc, target, 0.05, cost_one, nullptr));
  CHECK(cdpr_synchronize(all));
  std::vector<double> refc(n_seq);
  cudaMemcpy(refc.data(), cost_one, sizeof(double) * n_seq, cudaMemcpyDeviceToHost);
  cudaFree(cost_one);
  cdpr_destroy(all);
  double worst = 0.0;
  for (int64_t s = 0; s < n_seq; ++s) worst = std::fmax(worst, std::fabs(sums[0][s] - refc[s]) / std::fabs(refc[s]));
  std::printf("[config 5] %lld robots x %lld sequences x %lld steps over %d devices: all-reduced cost identical on every device: %s; max rel diff vs 1-GPU run %.2e\n",
              (long long)(robots_per_dev * G), (long long)n_seq, (long long)(n_cmd * spc), G, ok_same ? "True" : "False", worst);
  cdpr_comm_destroy(comm);
  const bool ok = ok_gather && ok_same && worst < 1e-13;
  std::printf("%s\n", ok ? "MULTI-GPU C++ HOST CHECK PASSED" : "MULTI-GPU C++ HOST CHECK FAILED");
  return ok ? 0 : 1;
}
