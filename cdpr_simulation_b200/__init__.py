"""cdpr_simulation_b200 -- B200-native batched CDPR step (kinematics, PID cable-force law, rigid-body
update of the cdpr_gazebo plugin) behind a C ABI.  See DESIGN.md."""
from .api import (CdprBatch, CdprError, Config, PidParams, default_config, load, lib_path, measure_fp64_tflops, dterm_weights,
                  MODE_FORCE, MODE_POSITION, MODE_VELOCITY, EXPORTS)

__all__ = ["CdprBatch", "CdprError", "Config", "PidParams", "default_config", "load", "lib_path", "measure_fp64_tflops", "dterm_weights",
           "MODE_FORCE", "MODE_POSITION", "MODE_VELOCITY", "EXPORTS"]
