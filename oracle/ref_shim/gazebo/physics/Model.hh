#include "gazebo/physics/World.hh"
