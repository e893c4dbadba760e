"""The reference's three manual command drivers restated as headless generators:

  SineVelocity    src/sinevelocitytest.cpp:33-49    100 Hz, v = amp * sin(time * freq * 2 * pi)
  SquareVelocity  src/squarevelocitytest.cpp:19-33   10 Hz, +-0.06 m/s with a dead band (|sin| < sqrt(0.5) -> 0), 0.05 Hz
  SquarePosition  src/squarepositiontest.cpp:19-34   10 Hz, bias + copysign(0.05 m, sin), 0.1 Hz

Every driver stores a double into float32 Joy.axes (the same value on all cables) and accumulates its publisher time
with `time += 1.0 / publish_frequency`.  `run` plays a driver against anything with set_velocity_cmd /
set_position_cmd / step (the CUDA batch, or any adapter with the same three methods)."""
from __future__ import annotations

import math

import numpy as np


class _Driver:
    publish_hz = 100.0
    topic = "jointVelocities"

    def __init__(self):
        self.time = 0.0

    def value(self) -> float:
        raise NotImplementedError

    def publish(self) -> np.float32:
        """One loop iteration of the reference node: compute, store as float32, advance the publisher clock."""
        v = np.float32(self.value())
        self.time += 1.0 / self.publish_hz
        return v


class SineVelocity(_Driver):
    publish_hz = 100.0

    def __init__(self, amp: float = 0.05, freq: float = 0.1):
        super().__init__()
        self.amp, self.freq = amp, freq

    def value(self):
        return self.amp * math.sin(self.time * self.freq * 2 * math.pi)


class SquareVelocity(_Driver):
    publish_hz = 10.0

    def __init__(self, amp: float = 0.06, freq: float = 0.05):
        super().__init__()
        self.amp, self.freq = amp, freq

    def value(self):
        sine = math.sin(self.time * self.freq * 2 * math.pi)
        return math.copysign(self.amp, sine) if abs(sine) >= math.sqrt(0.5) else 0.0


class SquarePosition(_Driver):
    publish_hz = 10.0
    topic = "jointPositions"

    def __init__(self, amp: float = 0.05, bias: float = 0.0, freq: float = 0.1):
        super().__init__()
        self.amp, self.bias, self.freq = amp, bias, freq

    def value(self):
        sine = math.sin(self.time * self.freq * 2 * math.pi)
        return self.bias + math.copysign(self.amp, sine)


def run(target, driver: _Driver, n_instances: int, n_cables: int, steps: int, dt: float = 0.001, on_publish=None):
    """Headless schedule of SURVEY.md App. A.4: the k-th command is latched for physics steps k*P+1 .. (k+1)*P,
    P = (1 / publish_hz) / dt.  Steps between publishes run as ONE launch."""
    per = int(round((1.0 / driver.publish_hz) / dt))
    done = 0
    while done < steps:
        v = driver.publish()
        axes = np.full((n_instances, n_cables), v, dtype=np.float32)
        if driver.topic == "jointPositions":
            target.set_position_cmd(axes)
        else:
            target.set_velocity_cmd(axes)
        if on_publish:
            on_publish(axes)
        k = min(per, steps - done)
        target.step(k)
        done += k
