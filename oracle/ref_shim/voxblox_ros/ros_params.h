// Shim: stands in for <voxblox_ros/ros_params.h> (oracle/_ref, test infrastructure only).
#include "../voxblox_shim.h"
