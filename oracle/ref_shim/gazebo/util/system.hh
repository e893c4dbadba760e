// empty stand-in for gazebo/util/system.hh -- TEST INFRASTRUCTURE ONLY
