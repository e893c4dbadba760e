// sinevelocitytest_main.cpp -- the reference's sinevelocitytest driver (src/sinevelocitytest.cpp:33-49) run headless
// against the batch plugin shim: a 100 Hz publisher of float32 sine velocity commands into a 1 kHz stepping plugin.
// usage: cdpr_sinevelocitytest [instances] [steps] [device]   -> prints the platform pose of robot 0 every 100 steps
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "CdprBatchPlugin.h"

int main(int argc, char **argv) {
  const double cPublishFrequency = 100.0, cVelocityAmplitude = 0.05, cVelocityFrequency = 0.1;
  const int64_t n = argc > 1 ? atoll(argv[1]) : 4;
  const int steps = argc > 2 ? atoi(argv[2]) : 1000;
  const int device = argc > 3 ? atoi(argv[3]) : 0;
  cdpr_config cfg;
  cdpr_config_default(&cfg, 4);
  cdpr_host::CdprBatchPlugin plugin;
  try {
    plugin.Load(cfg, n, device);
  } catch (const std::exception &e) {
    std::fprintf(stderr, "Load failed: %s\n", e.what());
    return 2;
  }
  const int stepsPerCommand = (int)std::llround((1.0 / cPublishFrequency) / cfg.dt);
  cdpr_host::Joy velocityCommand;
  velocityCommand.axes.resize((size_t)n * plugin.wireCount());
  double time = 0.0;
  for (int step = 0; step < steps; ++step) {
    if (step % stepsPerCommand == 0) {
      const double velocity = cVelocityAmplitude * sin(time * cVelocityFrequency * 2 * M_PI);
      for (auto &a : velocityCommand.axes) a = (float)velocity;
      plugin.cableVelocityCommandCallback(velocityCommand);
      time += 1.0 / cPublishFrequency;
    }
    plugin.update();
    if ((step + 1) % 100 == 0) {
      const auto &ps = plugin.platformState();
      std::printf("%d %.17g %.17g %.17g %.17g\n", step + 1, ps.pose[0], ps.pose[1], ps.pose[2], plugin.jointStates().effort[0]);
    }
  }
  return 0;
}
